// Host-side tables of compiled kernel instantiations, shared between the translation units of the library:
//   lb_k_step.cu   the one-update kernel's tile variants (lb_fused.cuh, lb_tma.cuh)
//   lb_k_march.cu  the two-update kernels (lb_march.cuh; with -DLB_EXPERIMENTS also the round-1 shared-memory
//                  tiles of lb_tb2.cuh / lb_tb2v.cuh)
//   lb_d2q9.cu     the C ABI, which picks from these tables
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "lb_fused.cuh"
#include "lb_tb2.cuh"

struct LbVariant {
    const char *name;
    int dtype, math, model, V, WX, WY, R;
    void (*launch)(const lb::StepParams &, cudaStream_t);
    bool is_default;
    void (*launch_tma)(const CUtensorMap &, const CUtensorMap &, const lb::StepParams &, cudaStream_t);   // non-null: TMA-staged kernel
    int tma_ty;                                                                                       // its box height
};
extern const LbVariant g_variants[];
extern const int g_nvariants;

// two updates per launch
enum LbTbKind { LB_TB_OFF = 0, LB_TB_MARCH = 1, LB_TB_ROWS = 2, LB_TB_CELLS = 3 };
struct LbTbShape {
    const char *name;
    int kind;
    int seg_rows;              // LB_TB_MARCH: rows per segment
    int nw;                    // LB_TB_MARCH: warps (= strips) per CTA
    int bx, by, nt;            // LB_TB_ROWS / LB_TB_CELLS: tile geometry (round-1 kernels, -DLB_EXPERIMENTS)
    // [dtype][math]; LB_TB_MARCH and LB_TB_ROWS take StepParams, LB_TB_CELLS takes Tb2Params
    void (*launch_march[2][2])(const lb::StepParams &, cudaStream_t);
    void (*launch_rows[2][2])(const lb::StepParams &, dim3, size_t, cudaStream_t);
    void (*launch_cells[2][2])(const lb::Tb2Params &, dim3, size_t, cudaStream_t);
    int depth;                 // LB_TB_MARCH: lattice updates per launch (0 = the default, two)
    int minb;                  // LB_TB_MARCH: CTAs per SM the kernel is compiled for (__launch_bounds__)
};
extern const LbTbShape g_tb_shapes[];
extern const int g_ntb;
extern const char *const g_tb_auto_f32[2];   // names of the shapes lb_step picks on its own: [no mask, mask]
extern const char *const g_tb_auto_f64[2];
extern const char *const g_tb_auto_f32_3;    // ... for three updates per launch
extern const char *const g_tb_auto_f64_3;

// ---- launch geometry of the marching kernels, shared by the launchers (lb_k_march.cu) and lb_plan_march_launch ----
// Strips next to a halo edge (their CTAs come first, wait for the neighbours' flags and count into the hand-shake): strip 0
// and / or the last one -- and the one before it when the last strip is narrower than the GHOST_COLS columns a slab
// publishes, which then belong to two strips.
inline int lb_march_edge_strips(int nx, int out, int nstrips, bool west_halo, bool east_halo)
{
    const int ne_e = !east_halo ? 0 : (nstrips > 1 && nx - (nstrips - 1) * out < lb::GHOST_COLS) ? 2 : 1;
    const int ne = (west_halo ? 1 : 0) + ne_e;
    return ne > nstrips ? nstrips : ne;
}

// Segments of a launch: p.seg_rows rows each; when the caller names a shorter height too (p.seg_rows2), the last rows of
// the range -- two waves of short work items' worth -- are cut into segments of that height, so that the launch does
// not end with half the SMs waiting for a few tall items (lattices too small for that stay uniform).  Sets p.seg_tall
// and p.seg_rows2, returns the number of segments.
inline int lb_march_segments(lb::StepParams &p, int nstrips, int nw, int minb)
{
    const int rows = p.y_end - p.y_begin, S1 = p.seg_rows;
    int n1 = (rows + S1 - 1) / S1, n2 = 0, S2 = S1;
    if (p.seg_rows2 > 0 && p.seg_rows2 < S1 && p.sm_count > 0) {
        const long long short_rows = 2ll * p.sm_count * minb * nw * p.seg_rows2 / nstrips;
        if (short_rows >= p.seg_rows2 && short_rows < rows / 2) {
            S2 = p.seg_rows2;
            n1 = (int)((rows - short_rows) / S1);
            n2 = (rows - n1 * S1 + S2 - 1) / S2;
        }
    }
    p.seg_tall = n1;
    p.seg_rows2 = S2;
    return n1 + n2;
}
