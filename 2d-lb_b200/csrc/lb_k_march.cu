// Instantiations of the two-update kernels.  Default build: the marching kernel (lb_march.cuh).
// -DLB_EXPERIMENTS adds its first version, which gathered the columns beside each strip with scalar loads
// (lb_march_rim.cuh), and the round-1 shared-memory tiles (lb_tb2v.cuh row-per-warp, lb_tb2.cuh
// cell-per-thread), kept for A/B measurements (profiles/README.md).
#include "lb_host.h"
#include "lb_march.cuh"
#ifdef LB_EXPERIMENTS
#include "lb_march_rim.cuh"
#include "lb_tb2v.cuh"
#endif

using namespace lb;

// BF: branch-free obstacle handling (three instantiations: no mask / mask / mask + run-time velocity zeroing)
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED, bool PF = false, bool SH = false, bool BF = false>
static void launch_march(const StepParams &p_in, cudaStream_t st)
{
    StepParams p = p_in;
    constexpr int OUT = 30 * V;
    const int nstrips = (p.nx + OUT - 1) / OUT;
    const int nseg = lb_march_segments(p, nstrips, NW, MINB);
    // strips next to a halo edge: 0 on a single slab, else strip 0 and / or the last one
    int ne = 0;
    if (p.edge_first) {
        ne = lb_march_edge_strips(p.nx, OUT, nstrips, p.west == EDGE_HALO, p.east == EDGE_HALO);
    }
    p.tiles_x = nstrips;
    p.tiles_y = nseg;
    p.edge_first = ne;
    p.edge_tiles_y = (ne * nseg + NW - 1) / NW;      // edge CTAs (each serves both sides when there are two)
    const unsigned grid = (unsigned)p.edge_tiles_y + (unsigned)(((long long)(nstrips - ne) * nseg + NW - 1) / NW);
    if (BF) {
        if (p.mask == nullptr) fused_march_kernel<T, V, MATH, NW, MINB, PACKED, PF, 0, SH, 0><<<grid, 32 * NW, 0, st>>>(p);
        else if (p.zero_obstacle_velocity) fused_march_kernel<T, V, MATH, NW, MINB, PACKED, PF, -1, SH, 1><<<grid, 32 * NW, 0, st>>>(p);
        else fused_march_kernel<T, V, MATH, NW, MINB, PACKED, PF, 0, SH, 1><<<grid, 32 * NW, 0, st>>>(p);
        return;
    }
    // handles that do not zero obstacle velocities run the instantiation without that code
    if (p.zero_obstacle_velocity) fused_march_kernel<T, V, MATH, NW, MINB, PACKED, PF, -1, SH><<<grid, 32 * NW, 0, st>>>(p);
    else fused_march_kernel<T, V, MATH, NW, MINB, PACKED, PF, 0, SH><<<grid, 32 * NW, 0, st>>>(p);
}
// K updates per launch (fused_march_k_kernel): kept rows in shared memory, branch-free obstacle code; fp64 with two
// overlap lanes per side (its lane carries two columns, three levels need three)
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED, int K, int OVL = 1>
static void launch_march_k(const StepParams &p_in, cudaStream_t st)
{
    StepParams p = p_in;
    constexpr int OUT = (32 - 2 * OVL) * V;
    const int nstrips = (p.nx + OUT - 1) / OUT;
    const int nseg = lb_march_segments(p, nstrips, NW, MINB);
    int ne = 0;
    if (p.edge_first) {
        ne = lb_march_edge_strips(p.nx, OUT, nstrips, p.west == EDGE_HALO, p.east == EDGE_HALO);
    }
    p.tiles_x = nstrips;
    p.tiles_y = nseg;
    p.edge_first = ne;
    p.edge_tiles_y = (ne * nseg + NW - 1) / NW;
    const unsigned grid = (unsigned)p.edge_tiles_y + (unsigned)(((long long)(nstrips - ne) * nseg + NW - 1) / NW);
    if (p.mask == nullptr) fused_march_k_kernel<T, V, MATH, NW, MINB, PACKED, K, 0, 0, OVL><<<grid, 32 * NW, 0, st>>>(p);
    else if (p.zero_obstacle_velocity) fused_march_k_kernel<T, V, MATH, NW, MINB, PACKED, K, -1, 1, OVL><<<grid, 32 * NW, 0, st>>>(p);
    else fused_march_k_kernel<T, V, MATH, NW, MINB, PACKED, K, 0, 1, OVL><<<grid, 32 * NW, 0, st>>>(p);
}
#define MARCH3(NW, MINB, S)                                                                                      \
    {"march3.w" #NW "b" #MINB ".s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                               \
     {{launch_march_k<float, 4, MATH_STRICT, NW, MINB, true, 3>, launch_march_k<float, 4, MATH_FAST, NW, MINB, true, 3>}, \
      {launch_march_k<double, 2, MATH_STRICT, NW, MINB, false, 3, 2>, launch_march_k<double, 2, MATH_FAST, NW, MINB, false, 3, 2>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, 3, MINB}
#define MARCH3_S(NW, MINB) MARCH3(NW, MINB, 8), MARCH3(NW, MINB, 16), MARCH3(NW, MINB, 32), MARCH3(NW, MINB, 64), MARCH3(NW, MINB, 128)

// name: march.w<warps per CTA>b<CTAs per SM>[.sh | .scalar | .pf].s<rows per segment>
//   .sh      the kept rows of the intermediate level in shared memory (thread-private slots): the shipped form
//   .sh.bf   ... and the bounce-back swap as selects instead of branches (one instantiation per "has a mask" /
//            "zeroes obstacle velocities"): the shipped form on lattices WITH an obstacle mask
//   (none)   ... in registers
//   .scalar  registers, scalar fp32 collision instead of the packed one (A/B)
//   .pf      registers, software-prefetched loads (LB_EXPERIMENTS; measured slower)
#define MARCH(NW, MINB, PACKED, PN, S)                                                                           \
    {"march.w" #NW "b" #MINB PN ".s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                             \
     {{launch_march<float, 4, MATH_STRICT, NW, MINB, PACKED>, launch_march<float, 4, MATH_FAST, NW, MINB, PACKED>},       \
      {launch_march<double, 2, MATH_STRICT, NW, MINB, false>, launch_march<double, 2, MATH_FAST, NW, MINB, false>}},      \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, 0, MINB}
#define MARCHPF(NW, MINB, S)                                                                                     \
    {"march.w" #NW "b" #MINB ".pf.s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                            \
     {{launch_march<float, 4, MATH_STRICT, NW, MINB, true, true>, launch_march<float, 4, MATH_FAST, NW, MINB, true, true>},   \
      {launch_march<double, 2, MATH_STRICT, NW, MINB, false, true>, launch_march<double, 2, MATH_FAST, NW, MINB, false, true>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, 0, MINB}
#define MARCHSH_S(NW, MINB) MARCHSH(NW, MINB, 8), MARCHSH(NW, MINB, 16), MARCHSH(NW, MINB, 32), MARCHSH(NW, MINB, 64), MARCHSH(NW, MINB, 128), MARCHSH(NW, MINB, 256)
#define MARCHSH(NW, MINB, S)                                                                                     \
    {"march.w" #NW "b" #MINB ".sh.s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                            \
     {{launch_march<float, 4, MATH_STRICT, NW, MINB, true, false, true>, launch_march<float, 4, MATH_FAST, NW, MINB, true, false, true>},   \
      {launch_march<double, 2, MATH_STRICT, NW, MINB, false, false, true>, launch_march<double, 2, MATH_FAST, NW, MINB, false, false, true>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, 0, MINB}
#define MARCHBF(NW, MINB, S)                                                                                     \
    {"march.w" #NW "b" #MINB ".sh.bf.s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                         \
     {{launch_march<float, 4, MATH_STRICT, NW, MINB, true, false, true, true>, launch_march<float, 4, MATH_FAST, NW, MINB, true, false, true, true>},   \
      {launch_march<double, 2, MATH_STRICT, NW, MINB, false, false, true, true>, launch_march<double, 2, MATH_FAST, NW, MINB, false, false, true, true>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, 0, MINB}
#define MARCH_S(NW, MINB, PACKED, PN) MARCH(NW, MINB, PACKED, PN, 32), MARCH(NW, MINB, PACKED, PN, 64), MARCH(NW, MINB, PACKED, PN, 128), MARCH(NW, MINB, PACKED, PN, 256)

#ifdef LB_EXPERIMENTS
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED>
static void launch_rim(const StepParams &p_in, cudaStream_t st)
{
    StepParams p = p_in;
    constexpr int SPAN = 32 * V;
    const int nstrips = p.pitch / SPAN;                // the pitch is a multiple of SPAN
    p.tiles_x = (nstrips + NW - 1) / NW;
    p.tiles_y = (p.ny + p.seg_rows - 1) / p.seg_rows;  // segments
    p.edge_tiles_y = p.tiles_y;                        // one edge CTA per segment and side
    const unsigned grid = (unsigned)p.tiles_x * (unsigned)p.tiles_y;
    fused_march_rim_kernel<T, V, MATH, NW, MINB, PACKED><<<grid, 32 * NW, 0, st>>>(p);
}

#define RIM(NW, MINB, PACKED, PN, S)                                                                         \
    {"rim.w" #NW "b" #MINB PN ".s" #S, LB_TB_MARCH, S, NW, 0, 0, 0,                                           \
     {{launch_rim<float, 4, MATH_STRICT, NW, MINB, PACKED>, launch_rim<float, 4, MATH_FAST, NW, MINB, PACKED>},   \
      {launch_rim<double, 2, MATH_STRICT, NW, MINB, false>, launch_rim<double, 2, MATH_FAST, NW, MINB, false>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}}
#define RIM_S(NW, MINB, PACKED, PN) RIM(NW, MINB, PACKED, PN, 32), RIM(NW, MINB, PACKED, PN, 64), RIM(NW, MINB, PACKED, PN, 128), RIM(NW, MINB, PACKED, PN, 256)

template <typename T, int MATH, int BX, int BY, int NT, int MINB>
static void launch_tb2(const Tb2Params &p, dim3 grid, size_t smem, cudaStream_t st)
{
    static bool configured[64] = {};                  // one opt-in per instantiation and device (dynamic smem > 48 KB)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(fused_two_step_kernel<T, MATH, BX, BY, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = true;
    }
    fused_two_step_kernel<T, MATH, BX, BY, NT, MINB><<<grid, NT, smem, st>>>(p);
}
template <typename T, int V, int MATH, int BY, int NW, int MINB>
static void launch_tb2v(const StepParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(fused_two_step_v2_kernel<T, V, MATH, BY, NW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = true;
    }
    fused_two_step_v2_kernel<T, V, MATH, BY, NW, MINB><<<grid, 32 * NW, smem, st>>>(p);
}
#define TB2(BX, BY, NT, MINB)                                                                          \
    {#BX "x" #BY ".t" #NT, LB_TB_CELLS, 0, 0, BX, BY, NT, {{nullptr, nullptr}, {nullptr, nullptr}},         \
     {{nullptr, nullptr}, {nullptr, nullptr}},                                                          \
     {{launch_tb2<float, MATH_STRICT, BX, BY, NT, MINB>, launch_tb2<float, MATH_FAST, BX, BY, NT, MINB>},  \
      {launch_tb2<double, MATH_STRICT, BX, BY, NT, MINB>, launch_tb2<double, MATH_FAST, BX, BY, NT, MINB>}}}
#define TB2V(BY, NW, MINB)                                                                             \
    {"rows" #BY ".w" #NW, LB_TB_ROWS, 0, 0, 0, BY, 32 * NW, {{nullptr, nullptr}, {nullptr, nullptr}},       \
     {{launch_tb2v<float, 4, MATH_STRICT, BY, NW, MINB>, launch_tb2v<float, 4, MATH_FAST, BY, NW, MINB>},  \
      {launch_tb2v<double, 2, MATH_STRICT, BY, NW, MINB>, launch_tb2v<double, 2, MATH_FAST, BY, NW, MINB>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}}
#endif

const LbTbShape g_tb_shapes[] = {
    {"off", LB_TB_OFF, 0, 0, 0, 0, 0, {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}},
    MARCHSH_S(4, 5),
    MARCHBF(4, 5, 8), MARCHBF(4, 5, 16), MARCHBF(4, 5, 32), MARCHBF(4, 5, 64), MARCHBF(4, 5, 128),
    MARCHBF(4, 6, 8), MARCHBF(4, 6, 16), MARCHBF(4, 6, 32), MARCHBF(4, 6, 64), MARCHBF(4, 6, 128),
    MARCH_S(4, 4, true, ""),
    MARCH(4, 4, false, ".scalar", 64),
    MARCH3_S(4, 5), MARCH3_S(4, 4), MARCH3_S(4, 6),
#ifdef LB_EXPERIMENTS
    MARCHSH_S(4, 4), MARCHSH_S(4, 6), MARCHSH_S(8, 3), MARCHSH_S(2, 10),
    MARCH_S(8, 2, true, ""),
    MARCHPF(4, 3, 64),
    MARCH_S(4, 3, true, ""),
    MARCH_S(2, 8, true, ""),
    RIM_S(4, 4, true, ""),
    RIM_S(4, 4, false, ".scalar"),
    TB2(128, 16, 256, 2),
    TB2(128, 8, 256, 4),
    TB2V(6, 8, 3),
    TB2V(14, 8, 2),
    TB2V(6, 4, 6),
#endif
};
const int g_ntb = (int)(sizeof(g_tb_shapes) / sizeof(g_tb_shapes[0]));
// measured best per case (profiles/README.md section 9.3): [0] lattices without an obstacle mask, [1] with one;
// lb_step appends the segment height (".s<rows>", chosen from the lattice size)
const char *const g_tb_auto_f32[2] = {"march.w4b5.sh", "march.w4b6.sh.bf"};
const char *const g_tb_auto_f64[2] = {"march.w4b5.sh", "march.w4b5.sh.bf"};
const char *const g_tb_auto_f32_3 = "march3.w4b4";   // three updates per launch (large lattices)
const char *const g_tb_auto_f64_3 = "march3.w4b5";
