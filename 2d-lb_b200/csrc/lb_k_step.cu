// Instantiations of the one-update kernel (lb_fused.cuh) and its TMA-staged twin (lb_tma.cuh).
// The default build carries the shipped tile (2x2 warps, one row per warp, 6 CTAs per SM) for every
// dtype / math / model, a few alternates that exercise the other vector widths in the tests, and one
// TMA-staged tile per dtype.  -DLB_EXPERIMENTS adds the round-1 tuning sweep (tools/sweep.py).
#include "lb_host.h"
#include "lb_tma.cuh"
#include "../../include/lb_d2q9.h"

using namespace lb;

template <typename T, int V, int MATH, int TY, int MINB, int STP, int MODEL>
static void launch_tma_variant(const CUtensorMap &map_n, const CUtensorMap &map_w, const StepParams &p_in, cudaStream_t st)
{
    StepParams p = p_in;
    p.tiles_x = (p.pitch + 32 * V - 1) / (32 * V);
    p.tiles_y = (p.ny + TY - 1) / TY;
    const unsigned gy = p.tiles_y < 65535 ? p.tiles_y : 65535;
    const dim3 grid((unsigned)p.tiles_x, gy, ((unsigned)p.tiles_y + gy - 1) / gy);
    fused_step_tma_kernel<T, V, MATH, TY, MINB, STP, MODEL><<<grid, 32 * TY, 0, st>>>(map_n, map_w, p);
}
#define VART(T, TN, DT, V, M, MN, TY, MINB)                                                         \
    {TN "." MN ".tma.v" #V ".ty" #TY ".b" #MINB, DT, M, MODEL_D2Q9, V, 1, TY, 1, nullptr, false,       \
     &launch_tma_variant<T, V, M, TY, MINB, 0, MODEL_D2Q9>, TY}

template <typename T, int V, int MATH, int WX, int WY, int R, int MINB, int LDP, int STP, int MODEL = MODEL_D2Q9>
static void launch_variant(const StepParams &p_in, cudaStream_t st)
{
    StepParams p = p_in;
    constexpr int SPAN = 32 * V;
    p.tiles_x = (p.pitch + SPAN * WX - 1) / (SPAN * WX);
    p.tiles_y = (p.y_end - p.y_begin + WY * R - 1) / (WY * R);
    dim3 grid;
    if (p.edge_first) {
        // edge tiles are TALL (edge_rows rows per warp): few CTAs take part in the hand-shake
        p.edge_tiles_y = (p.ny + WY * p.edge_rows - 1) / (WY * p.edge_rows);
        const unsigned n_edge = (p.tiles_x < 2 ? 1u : 2u) * (unsigned)p.edge_tiles_y;
        const unsigned n_int = p.tiles_x > 2 ? (unsigned)(p.tiles_x - 2) * (unsigned)p.tiles_y : 0u;
        grid = dim3(n_edge + n_int, 1, 1);
        // a last tile column narrower than the published columns shares them with the column before it
        p.east_pair = p.east == EDGE_HALO && p.tiles_x >= 2 && p.nx - (p.tiles_x - 1) * SPAN * WX < GHOST_COLS;
        p.done_e_count = p.edge_tiles_y + (p.east_pair ? (p.tiles_x == 2 ? p.edge_tiles_y : p.tiles_y) : 0);
    }
    else {
        const unsigned gy = p.tiles_y < 65535 ? p.tiles_y : 65535;
        grid = dim3((unsigned)p.tiles_x, gy, ((unsigned)p.tiles_y + gy - 1) / gy);
    }
    fused_step_kernel<T, V, MATH, WX, WY, R, MINB, LDP, STP, MODEL><<<grid, 32 * WX * WY, 0, st>>>(p);
}

#define VAR(T, TN, DT, V, M, MN, WX, WY, R, MINB, LDP, STP, DEF)                                    \
    {TN "." MN ".v" #V ".wx" #WX ".wy" #WY ".r" #R ".b" #MINB ".ld" #LDP ".st" #STP, DT, M, MODEL_D2Q9, V, \
     WX, WY, R, &launch_variant<T, V, M, WX, WY, R, MINB, LDP, STP>, DEF, nullptr, 0}
// incompressible model (D2Q9i.cl): the default tile configuration only
#define VARI(T, TN, DT, V, M, MN)                                                                   \
    {TN "." MN ".d2q9i.v" #V ".wx2.wy2.r1.b6.ld1.st0", DT, M, MODEL_D2Q9I, V, 2, 2, 1,               \
     &launch_variant<T, V, M, 2, 2, 1, 6, 1, 0, MODEL_D2Q9I>, true, nullptr, 0}

// shipped tile + the alternates the parity tests walk through (other vector width, other CTA shape)
#define VARS_CORE(T, TN, DT, VMAX, VHALF)                                                           \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 6, 1, 0, true),                                \
    VAR(T, TN, DT, VHALF, MATH_FAST, "fast", 2, 2, 1, 8, 1, 0, false),                              \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 6, 1, 0, true),                            \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 4, 1, 1, 6, 1, 0, false),                           \
    VAR(T, TN, DT, VHALF, MATH_STRICT, "strict", 2, 2, 1, 8, 1, 0, false),                          \
    VAR(T, TN, DT, VHALF, MATH_STRICT, "strict", 2, 2, 2, 6, 1, 0, false)

#ifdef LB_EXPERIMENTS
#define VARS_SWEEP(T, TN, DT, VMAX, VHALF)                                                          \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 4, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 5, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 7, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 8, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 4, 1, 1, 6, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 4, 1, 1, 4, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 4, 1, 2, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 4, 1, 3, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 4, 1, 4, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 4, 2, 1, 4, 1, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 6, 0, 0, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 6, 2, 1, false),                               \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 6, 1, 1, false),                               \
    VAR(T, TN, DT, VHALF, MATH_FAST, "fast", 4, 2, 1, 4, 1, 0, false),                              \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 5, 1, 0, false),                           \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 4, 1, 0, false),                           \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 4, 1, 3, 1, 0, false),
#endif

const LbVariant g_variants[] = {
    VARS_CORE(float, "f32", LB_F32, 4, 2),
    VARS_CORE(double, "f64", LB_F64, 2, 1),
    VARI(float, "f32", LB_F32, 4, MATH_STRICT, "strict"), VARI(float, "f32", LB_F32, 4, MATH_FAST, "fast"),
    VARI(double, "f64", LB_F64, 2, MATH_STRICT, "strict"), VARI(double, "f64", LB_F64, 2, MATH_FAST, "fast"),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 4, 6),
    VART(double, "f64", LB_F64, 2, MATH_STRICT, "strict", 4, 6),
#ifdef LB_EXPERIMENTS
    VARS_SWEEP(float, "f32", LB_F32, 4, 2)
    VARS_SWEEP(double, "f64", LB_F64, 2, 1)
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 4, 4),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 8, 3), VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 8, 2),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 2, 8), VART(float, "f32", LB_F32, 4, MATH_FAST, "fast", 4, 6),
    VART(float, "f32", LB_F32, 4, MATH_FAST, "fast", 8, 3),
    VART(double, "f64", LB_F64, 2, MATH_STRICT, "strict", 8, 3),
    VART(double, "f64", LB_F64, 2, MATH_FAST, "fast", 4, 6),
#endif
};
const int g_nvariants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));
