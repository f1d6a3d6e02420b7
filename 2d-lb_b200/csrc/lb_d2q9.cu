// C-ABI implementation of include/lb_d2q9.h: handle management, the fused-step launcher
// (CUDA-graph batched), single-stage kernels, device-side initialisers and the NVLink
// peer-memory halo plumbing.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false
// (see __graft_entry__.build()).
#include "../../include/lb_d2q9.h"
#include "lb_fused.cuh"
#include "lb_cython.cuh"
#include "lb_oldcl.cuh"
#include "lb_tma.cuh"
#include "lb_tb2.cuh"
#include "lb_tb2v.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace lb;

// =====================================================================================
// handle
// =====================================================================================
struct HaloLayout {
    size_t ghost_bytes;     // one ghost column: 3*(ny+2) elements, rounded to 256 B
    size_t off_ghost_w[2], off_ghost_e[2];
    size_t off_flag_w, off_flag_e, off_done_w, off_done_e, off_error;
    size_t total;
};

static HaloLayout halo_layout(int ny, int elem)
{
    HaloLayout h;
    h.ghost_bytes = (((size_t)3 * (ny + 2) * elem) + 255) / 256 * 256;
    size_t o = 0;
    for (int p = 0; p < 2; ++p) { h.off_ghost_w[p] = o; o += h.ghost_bytes; }
    for (int p = 0; p < 2; ++p) { h.off_ghost_e[p] = o; o += h.ghost_bytes; }
    h.off_flag_w = o; o += 128;
    h.off_flag_e = o; o += 128;
    h.off_done_w = o; o += 128;
    h.off_done_e = o; o += 128;
    h.off_error = o; o += 128;
    h.total = o;
    return h;
}

struct lb_sim {
    lb_config cfg;
    int elem = 4;                 // bytes per population value
    int uv_elem = 4;              // bytes per u / v value (8 for the cython schemes: float64 like the reference)
    bool prestream_done = false;  // cython / opencl_old schemes: is the next step's BC + swap already applied to `cur`
    float *frozen = nullptr;      // opencl_old: the populations `move` never writes (lb_oldcl.cuh)
    int tb2_shape = -1;           // temporal blocking: -1 = automatic, 0 = off, else index into g_tb2_shapes
    int pitch = 0;                // row pitch in elements (multiple of 512 B)
    long long plane = 0;          // elements per plane
    size_t buf_bytes = 0;         // bytes of one guarded 9-plane buffer
    char *buf_base[2] = {nullptr, nullptr};
    void *buf[2] = {nullptr, nullptr};   // plane 0 / row 0 of each ping-pong buffer
    int cur = 0;                  // buffer holding the current post-collision state
    void *rho = nullptr, *u = nullptr, *v = nullptr, *feq = nullptr;
    uint8_t *mask = nullptr, *span_solid = nullptr;
    int mask_pitch = 0, nspans = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int variant = -1;
    int64_t launches = 0;
    uint32_t state_index = 0;     // number of fused steps taken (halo parity / flag value)
    // CUDA graphs of `graph_len` moment-free steps starting from buffer `cur` == index
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    int graph_len[2] = {0, 0};
    int graph_variant[2] = {-2, -2};
    // halo
    char *halo = nullptr;         // my arena
    HaloLayout hl{};
    char *peer[2] = {nullptr, nullptr};   // neighbour arenas (mapped)
    bool peer_ipc[2] = {false, false};
    double *mass_scratch = nullptr;
    CUtensorMap tmap[2][3][2];    // [buffer][box height 2,4,8][plain | haloed box]: TMA descriptors of the ping-pong buffers
    bool tmap_ok = false;
    std::string err;
};

static thread_local std::string g_create_error;

static int fail(lb_sim *s, int code, const std::string &msg)
{
    if (s) s->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(sim, LB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

// =====================================================================================
// fused kernel variants
// =====================================================================================
struct Variant {
    const char *name;
    int dtype, math, model, V, WX, WY, R;
    void (*launch)(const StepParams &, cudaStream_t);
    bool is_default;
    void (*launch_tma)(const CUtensorMap &, const CUtensorMap &, const StepParams &, cudaStream_t);   // non-null: TMA-staged kernel
    int tma_ty;                                                                   // its box height
};

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
        p.edge_rows = 16;
        p.edge_tiles_y = (p.ny + WY * p.edge_rows - 1) / (WY * p.edge_rows);
        const unsigned n_edge = (p.tiles_x < 2 ? 1u : 2u) * (unsigned)p.edge_tiles_y;
        const unsigned n_int = p.tiles_x > 2 ? (unsigned)(p.tiles_x - 2) * (unsigned)p.tiles_y : 0u;
        grid = dim3(n_edge + n_int, 1, 1);
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

#define VARS_FOR(T, TN, DT, VMAX, VHALF)                                                            \
    VAR(T, TN, DT, VMAX, MATH_FAST, "fast", 2, 2, 1, 6, 1, 0, true),                                \
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
    VAR(T, TN, DT, VHALF, MATH_FAST, "fast", 2, 2, 1, 8, 1, 0, false),                              \
    VAR(T, TN, DT, VHALF, MATH_FAST, "fast", 4, 2, 1, 4, 1, 0, false),                              \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 6, 1, 0, true),                            \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 5, 1, 0, false),                           \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 2, 1, 4, 1, 0, false),                           \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 2, 4, 1, 3, 1, 0, false),                           \
    VAR(T, TN, DT, VMAX, MATH_STRICT, "strict", 4, 1, 1, 6, 1, 0, false),                           \
    VAR(T, TN, DT, VHALF, MATH_STRICT, "strict", 2, 2, 1, 8, 1, 0, false),                          \
    VAR(T, TN, DT, VHALF, MATH_STRICT, "strict", 2, 2, 2, 6, 1, 0, false)

static const Variant g_variants[] = {
    VARS_FOR(float, "f32", LB_F32, 4, 2),
    VARS_FOR(double, "f64", LB_F64, 2, 1),
    VARI(float, "f32", LB_F32, 4, MATH_STRICT, "strict"), VARI(float, "f32", LB_F32, 4, MATH_FAST, "fast"),
    VARI(double, "f64", LB_F64, 2, MATH_STRICT, "strict"), VARI(double, "f64", LB_F64, 2, MATH_FAST, "fast"),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 4, 6), VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 4, 4),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 8, 3), VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 8, 2),
    VART(float, "f32", LB_F32, 4, MATH_STRICT, "strict", 2, 8), VART(float, "f32", LB_F32, 4, MATH_FAST, "fast", 4, 6),
    VART(float, "f32", LB_F32, 4, MATH_FAST, "fast", 8, 3),
    VART(double, "f64", LB_F64, 2, MATH_STRICT, "strict", 4, 6), VART(double, "f64", LB_F64, 2, MATH_STRICT, "strict", 8, 3),
    VART(double, "f64", LB_F64, 2, MATH_FAST, "fast", 4, 6),
};
static const int g_nvariants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));

static int default_variant(int dtype, int math, int model)
{
    for (int i = 0; i < g_nvariants; ++i)
        if (g_variants[i].dtype == dtype && g_variants[i].math == math && g_variants[i].model == model &&
            g_variants[i].is_default) return i;
    return -1;
}

// =====================================================================================
// auxiliary kernels (not on the hot path)
// =====================================================================================
template <typename T>
__global__ void k_feq_from_moments(int nx, int ny, int pitch, long long plane, const T *rho, const T *u,
                                   const T *v, T *feq, Consts<T> c, int model)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    T e[9];
    if (model == MODEL_D2Q9I) feq_strict_i<T>(c, rho[i], u[i], v[i], e);
    else feq_strict<T>(c, rho[i], u[i], v[i], e);
#pragma unroll
    for (int j = 0; j < 9; ++j) feq[j * plane + i] = e[j];
}

// D2Q9.cl `move` (+ `copy_buffer`): pull form into the other buffer; destinations whose source
// lies outside the domain keep whatever that buffer held (the reference's stale f_streamed slot).
template <typename T>
__global__ void k_stage_move(int nx, int ny, int pitch, long long plane, int periodic, const T *src, T *dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        int sx = x - ex[j], sy = y - ey[j];
        if (periodic) {
            if (sx < 0) sx += nx;
            if (sx >= nx) sx -= nx;
            if (sy < 0) sy += ny;
            if (sy >= ny) sy -= ny;
        } else if (sx < 0 || sx >= nx || sy < 0 || sy >= ny) continue;
        dst[j * plane + (long long)y * pitch + x] = src[j * plane + (long long)sy * pitch + sx];
    }
}

template <typename T>
__global__ void k_stage_bcs(int nx, int ny, int pitch, long long plane, int gnx, int x_off, int do_pipe,
                            const uint8_t *mask, int mask_pitch, T *f, Consts<T> c, int model)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    const bool solid = mask && mask[(long long)y * mask_pitch + x] == 1;
    const int gx = x_off + x;
    const bool bnd = do_pipe && (gx == 0 || gx == gnx - 1 || y == 0 || y == ny - 1);
    if (!solid && !bnd) return;
    T g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * plane + i];
    if (bnd) {
        if (model == MODEL_D2Q9I) pipe_bc<T, MODEL_D2Q9I>(c, gx, y, gnx, ny, g);
        else pipe_bc<T, MODEL_D2Q9>(c, gx, y, gnx, ny, g);
    }
    if (solid) bounce_back<T>(g);
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j * plane + i] = g[j];
}

template <typename T>
__global__ void k_stage_hydro(int nx, int ny, int pitch, long long plane, const T *f, T *rho, T *u, T *v, int model)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    T g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * plane + i];
    T r, a, b;
    if (model == MODEL_D2Q9I) moments<T, MATH_STRICT, MODEL_D2Q9I>(g, r, a, b);
    else moments<T, MATH_STRICT, MODEL_D2Q9>(g, r, a, b);
    rho[i] = r; u[i] = a; v[i] = b;
}

template <typename T>
__global__ void k_stage_collide(int nx, int ny, int pitch, long long plane, T *f, const T *feq, Consts<T> c)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j * plane + i] = f[j * plane + i] * c.keep + c.omega * feq[j * plane + i];
}

template <typename T>
__global__ void k_zero_velocity(int nx, int ny, int pitch, const uint8_t *mask, int mask_pitch, T *u, T *v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    if (mask[(long long)y * mask_pitch + x] == 1) {
        u[(long long)y * pitch + x] = (T)0;
        v[(long long)y * pitch + x] = (T)0;
    }
}

template <typename T>
__global__ void k_subsample(int nx, int ny, int pitch, int sx, int sy, int ox, const T *src, T *dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= ox) return;
    dst[(long long)j * ox + i] = src[(long long)(j * sy) * pitch + (long long)i * sx];
}

__global__ void k_selftest_rcp(uint32_t first, uint32_t last, unsigned long long *bad)
{
    unsigned long long local = 0;
    const unsigned long long n = (unsigned long long)last - first + 1ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float(first + (uint32_t)i);
        if (__float_as_uint(rcp_rn_nobranch(x)) != __float_as_uint(1.0f / x)) ++local;
        if (__float_as_uint(rcp_rn_nobranch(-x)) != __float_as_uint(1.0f / -x)) ++local;
    }
    if (local) atomicAdd(bad, local);
}

// one flag byte per 32 cells of a row: 0 = no solid node in the group, 1 = some, 2 = all 32 solid
__global__ void k_span_solid(int nx, int ny, const uint8_t *mask, int mask_pitch, uint8_t *span_solid, int nspans)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (s >= nspans || y >= ny) return;
    int n = 0;
    for (int e = 0; e < 32; ++e) {
        const int x = s * 32 + e;
        if (x < nx && mask[(long long)y * mask_pitch + x] == 1) ++n;
    }
    span_solid[(long long)y * nspans + s] = n == 32 ? 2 : (n ? 1 : 0);
}

// arithmetic-free twin of the fused kernel's memory traffic (see lb_selftest_copy in the header)
template <typename T, int V>
__global__ void __launch_bounds__(128, 6) k_copy_pattern(const T *__restrict__ src, T *__restrict__ dst, int pitch, int ny, long long plane)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = ((blockIdx.x * 2 + (warp & 1)) * 32 + lane) * V;
    const int y = (blockIdx.z * gridDim.y + blockIdx.y) * 2 + (warp >> 1);
    if (x0 >= pitch || y >= ny) return;
    const int ym = y > 0 ? y - 1 : ny - 1, yp = y < ny - 1 ? y + 1 : 0;
    const int rows[9] = {y, y, ym, y, yp, ym, ym, yp, yp};
    Pack<T, V> q[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = load_pack<T, V, 1>(src + j * plane + (long long)rows[j] * pitch + x0);
#pragma unroll
    for (int j = 0; j < 9; ++j) store_pack<T, V, 0>(dst + j * plane + (long long)y * pitch + x0, q[j]);
}

__global__ void k_mask_disk(int nx, int ny, int x_off, double cx, double cy, double r2, uint8_t *mask, int mask_pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const double dx = (double)(x_off + x) - cx, dy = (double)y - cy;
    mask[(long long)y * mask_pitch + x] = (dx * dx + dy * dy < r2) ? 1 : 0;
}

// counter-based N(0,1): splitmix64 of (seed, global cell, population) -> Box-Muller
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double normal01(unsigned long long seed, unsigned long long cell, int j)
{
    const unsigned long long h = mix64(mix64(seed ^ (cell * 9ull + (unsigned long long)j)));
    const double u1 = ((double)(h >> 40) + 0.5) * (1.0 / 16777216.0);            // (0,1)
    const double u2 = ((double)((h >> 16) & 0xFFFFFFull) + 0.5) * (1.0 / 16777216.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

template <typename T>
__global__ void k_init_synth(int nx, int ny, int pitch, long long plane, int gnx, int x_off, int kind, double u0,
                             double amplitude, unsigned long long seed, double inlet_rho, double outlet_rho,
                             const uint8_t *mask, int mask_pitch, T *f0, T *f1, T *rho, T *u, T *v, Consts<T> c, int model)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    const int gx = x_off + x;
    T r, a, b;
    if (kind == LB_SYNTH_PIPE_RAMP) {
        r = (T)(inlet_rho - (double)gx * (inlet_rho - outlet_rho) / (double)gnx);
        a = (T)0; b = (T)0;
    } else {
        const double yy = (double)y / (double)ny;
        r = (T)1;
        a = (T)(yy < 0.5 ? u0 * tanh(80.0 * (yy - 0.25)) : u0 * tanh(80.0 * (0.75 - yy)));
        b = (T)(0.05 * u0 * sinpi(2.0 * ((double)gx / (double)gnx + 0.25)));
    }
    if (mask && mask[(long long)y * mask_pitch + x] == 1) { a = (T)0; b = (T)0; }
    rho[i] = r; u[i] = a; v[i] = b;
    T e[9];
    if (model == MODEL_D2Q9I) feq_strict_i<T>(c, r, a, b, e);
    else feq_strict<T>(c, r, a, b, e);
    const unsigned long long cell = (unsigned long long)y * (unsigned long long)gnx + (unsigned long long)gx;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        T val = e[j];
        if (amplitude != 0.0) val = (T)((double)val * (1.0 + amplitude * normal01(seed, cell, j)));
        f0[j * plane + i] = val;
        f1[j * plane + i] = val;
    }
}

template <typename T>
__global__ void k_mass(int nx, int ny, int pitch, long long plane, const T *f, double *out)
{
    __shared__ double sh[256];
    double acc = 0.0;
    const int y = blockIdx.x;
    for (int x = threadIdx.x; x < nx; x += blockDim.x) {
        const long long i = (long long)y * pitch + x;
#pragma unroll
        for (int j = 0; j < 9; ++j) acc += (double)f[j * plane + i];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[y] = sh[0];
}

template <typename T>
__global__ void k_checksum(int nx, int ny, int pitch, long long plane, const T *f, unsigned long long *out)
{
    unsigned long long acc = 0;
    const int y = blockIdx.x;
    for (int x = threadIdx.x; x < nx; x += blockDim.x) {
        const long long i = (long long)y * pitch + x;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const T val = f[j * plane + i];
            if (sizeof(T) == 4) acc += (unsigned long long)__float_as_uint((float)val);
            else acc += (unsigned long long)__double_as_longlong((double)val);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// copies the current boundary columns into the neighbours' ghost columns and publishes the flag
template <typename T>
__global__ void k_halo_prime(int nx, int ny, int pitch, long long plane, const T *f, T *out_w, T *out_e,
                             unsigned int *flag_w_remote, unsigned int *flag_e_remote, unsigned int value)
{
    for (int y = threadIdx.x; y < ny; y += blockDim.x) {
        const long long row = (long long)y * pitch;
        if (out_w) {
            out_w[0 * (ny + 2) + y + 1] = f[3 * plane + row];
            out_w[1 * (ny + 2) + y + 1] = f[6 * plane + row];
            out_w[2 * (ny + 2) + y + 1] = f[7 * plane + row];
        }
        if (out_e) {
            out_e[0 * (ny + 2) + y + 1] = f[1 * plane + row + nx - 1];
            out_e[1 * (ny + 2) + y + 1] = f[5 * plane + row + nx - 1];
            out_e[2 * (ny + 2) + y + 1] = f[8 * plane + row + nx - 1];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (flag_w_remote) st_release_sys(flag_w_remote, value);
        if (flag_e_remote) st_release_sys(flag_e_remote, value);
    }
}

// =====================================================================================
// helpers
// =====================================================================================
static inline dim3 grid2d(const lb_sim *s, int bx = 128) { return dim3((s->cfg.nx + bx - 1) / bx, s->cfg.ny); }

template <typename T>
static Consts<T> consts_of(const lb_sim *s)
{
    return make_consts<T>(s->cfg.omega, s->cfg.inlet_rho, s->cfg.outlet_rho, s->cfg.cs2, s->cfg.cs22, s->cfg.two_cs4);
}

static CyConsts cy_consts_of(const lb_sim *s)
{
    return make_cy_consts(s->cfg.omega, s->cfg.inlet_rho, s->cfg.outlet_rho, s->cfg.cs2, s->cfg.cs22,
                          s->cfg.u_west, s->cfg.u_east);
}

static inline bool is_cython(const lb_sim *s) { return s->cfg.scheme == LB_SCHEME_CYTHON || s->cfg.scheme == LB_SCHEME_CYTHON_OLD; }
static inline bool is_oldcl(const lb_sim *s) { return s->cfg.scheme == LB_SCHEME_OPENCL_OLD; }

static OcParams oc_params_of(const lb_sim *s)
{
    OcParams p{};
    p.plane = s->plane; p.nx = s->cfg.nx; p.ny = s->cfg.ny; p.pitch = s->pitch;
    p.mask = s->mask; p.mask_pitch = s->mask_pitch;
    p.rho = (float *)s->rho; p.u = (float *)s->u; p.v = (float *)s->v;
    p.frozen = s->frozen;
    p.c = consts_of<float>(s);
    p.u_w = (float)s->cfg.u_west; p.u_e = (float)s->cfg.u_east;
    p.kw = 1. / (1. - (double)p.u_w); p.ke = 1. / (1. + (double)p.u_e);
    return p;
}

static void drop_graphs(lb_sim *s)
{
    for (int i = 0; i < 2; ++i) {
        if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
        s->graph[i] = nullptr; s->graph_len[i] = 0; s->graph_variant[i] = -2;
    }
}

static bool uses_halo(const lb_sim *s) { return s->cfg.west_edge == LB_EDGE_HALO || s->cfg.east_edge == LB_EDGE_HALO; }

static void fill_params(lb_sim *s, StepParams &p, int src_idx, int write_moments, uint32_t state_index)
{
    memset(&p, 0, sizeof(p));
    p.src = s->buf[src_idx];
    p.dst = s->buf[src_idx ^ 1];
    p.plane = s->plane;
    p.nx = s->cfg.nx; p.ny = s->cfg.ny; p.pitch = s->pitch;
    p.gnx = s->cfg.global_nx; p.x_off = s->cfg.x_offset;
    p.bc = s->cfg.bc; p.west = s->cfg.west_edge; p.east = s->cfg.east_edge;
    p.write_moments = write_moments;
    p.zero_obstacle_velocity = s->cfg.zero_obstacle_velocity;
    p.mask = s->mask; p.span_solid = s->span_solid; p.mask_pitch = s->mask_pitch; p.nspans = s->nspans;
    p.rho = s->rho; p.u = s->u; p.v = s->v;
    p.cf = consts_of<float>(s);
    p.cd = consts_of<double>(s);
    p.y_begin = 0; p.y_end = s->cfg.ny;
    if (uses_halo(s)) {
        const int rp = state_index & 1, wp = (state_index + 1) & 1;
        const HaloLayout &h = s->hl;
        p.ghost_w = s->halo + h.off_ghost_w[rp];
        p.ghost_e = s->halo + h.off_ghost_e[rp];
        p.flag_w_local = (unsigned int *)(s->halo + h.off_flag_w);
        p.flag_e_local = (unsigned int *)(s->halo + h.off_flag_e);
        p.done_w = (unsigned int *)(s->halo + h.off_done_w);
        p.done_e = (unsigned int *)(s->halo + h.off_done_e);
        p.error_word = (unsigned int *)(s->halo + h.off_error);
        if (s->peer[LB_WEST]) {   // my westward populations land in the west neighbour's EAST ghost
            p.out_w = s->peer[LB_WEST] + h.off_ghost_e[wp];
            p.flag_w_remote = (unsigned int *)(s->peer[LB_WEST] + h.off_flag_e);
        }
        if (s->peer[LB_EAST]) {
            p.out_e = s->peer[LB_EAST] + h.off_ghost_w[wp];
            p.flag_e_remote = (unsigned int *)(s->peer[LB_EAST] + h.off_flag_w);
        }
        p.step_id = state_index + 1;
        p.edge_first = 1;
    }
}

// TMA descriptors: the guarded 9-plane buffer as one [4 + 9*ny][pitch] tensor, box = (32*V) x TY
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ensure_tmaps(lb_sim *sim)
{
    if (sim->tmap_ok) return LB_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(sim, LB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    const encode_tiled_fn encode = (encode_tiled_fn)fn;
    const int V = sim->cfg.dtype == LB_F32 ? 4 : 2;
    const int heights[3] = {2, 4, 8};
    const int halo = 16 / sim->elem;
    for (int b = 0; b < 2; ++b)
        for (int h = 0; h < 3; ++h)
            for (int wide = 0; wide < 2; ++wide) {
            const cuuint64_t dims[2] = {(cuuint64_t)sim->pitch, (cuuint64_t)9 * sim->cfg.ny + 4};
            const cuuint64_t strides[1] = {(cuuint64_t)sim->pitch * sim->elem};
            const cuuint32_t box[2] = {(cuuint32_t)(32 * V + (wide ? 2 * halo : 0)), (cuuint32_t)heights[h]};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = encode(&sim->tmap[b][h][wide], sim->cfg.dtype == LB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                      2, sim->buf_base[b], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(sim, LB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        }
    sim->tmap_ok = true;
    return LB_OK;
}

static int launch_step(lb_sim *sim, int src_idx, int write_moments, uint32_t state_index)
{
    StepParams p;
    fill_params(sim, p, src_idx, write_moments, state_index);
    const Variant &var = g_variants[sim->variant];
    if (var.launch_tma) {
        int rc = ensure_tmaps(sim);
        if (rc) return rc;
        const int hi = var.tma_ty == 2 ? 0 : var.tma_ty == 4 ? 1 : 2;
        var.launch_tma(sim->tmap[src_idx][hi][0], sim->tmap[src_idx][hi][1], p, sim->stream);
    } else var.launch(p, sim->stream);
    CU(cudaGetLastError());
    sim->launches++;
    return LB_OK;
}

// ---- temporal blocking (lb_tb2.cuh, lb_tb2v.cuh): two steps per launch ---------------------------------
#define LB_TB2_AUTO_F32 "rows6.w8"
#define LB_TB2_AUTO_F64 "rows6.w8"
struct Tb2Shape {
    const char *name;
    int bx, by, nt;            // bx == 0: the row-per-warp version (lb_tb2v.cuh), tile width = 32*V
    void (*launch[2][2])(const Tb2Params &, dim3, size_t, cudaStream_t);    // [dtype][math], lb_tb2.cuh
    void (*launch_v[2][2])(const StepParams &, dim3, size_t, cudaStream_t); // [dtype][math], lb_tb2v.cuh
};

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
    {#BX "x" #BY ".t" #NT, BX, BY, NT,                                                                     \
     {{launch_tb2<float, MATH_STRICT, BX, BY, NT, MINB>, launch_tb2<float, MATH_FAST, BX, BY, NT, MINB>},  \
      {launch_tb2<double, MATH_STRICT, BX, BY, NT, MINB>, launch_tb2<double, MATH_FAST, BX, BY, NT, MINB>}}, \
     {{nullptr, nullptr}, {nullptr, nullptr}}}
#define TB2V(BY, NW, MINB)                                                                             \
    {"rows" #BY ".w" #NW, 0, BY, 32 * NW, {{nullptr, nullptr}, {nullptr, nullptr}},                         \
     {{launch_tb2v<float, 4, MATH_STRICT, BY, NW, MINB>, launch_tb2v<float, 4, MATH_FAST, BY, NW, MINB>},  \
      {launch_tb2v<double, 2, MATH_STRICT, BY, NW, MINB>, launch_tb2v<double, 2, MATH_FAST, BY, NW, MINB>}}}
static const Tb2Shape g_tb2_shapes[] = {
    {"off", 0, 0, 0, {{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}},
    TB2(128, 16, 256, 2),
    TB2(64, 32, 256, 2),
    TB2(128, 8, 256, 4),
    TB2(256, 8, 256, 2),
    TB2(128, 32, 512, 1),
    TB2(64, 16, 256, 3),
    TB2(64, 8, 128, 6),
    TB2V(6, 8, 3),
    TB2V(14, 8, 2),
    TB2V(22, 8, 1),
    TB2V(30, 8, 1),
    TB2V(6, 4, 6),
    TB2V(14, 4, 4),
};
static const int g_ntb2 = (int)(sizeof(g_tb2_shapes) / sizeof(g_tb2_shapes[0]));

static inline bool tb2_is_rows(int shape) { return shape > 0 && g_tb2_shapes[shape].bx == 0; }

static size_t tb2_smem_bytes(const lb_sim *sim, int shape)
{
    const Tb2Shape &t = g_tb2_shapes[shape];
    if (tb2_is_rows(shape)) {
        const int off = 16 / sim->elem, span = sim->elem == 4 ? 128 : 64;
        return (size_t)9 * (span + 2 * off) * (t.by + 2) * sim->elem;
    }
    return (size_t)9 * (t.bx + 2) * (t.by + 2) * sim->elem;
}

static bool tb2_servable(const lb_sim *sim)
{
    return sim->cfg.scheme == LB_SCHEME_OPENCL && sim->cfg.model == LB_MODEL_D2Q9 && !uses_halo(sim) &&
           sim->cfg.global_nx == sim->cfg.nx;
}

static int tb2_find(const char *name)
{
    for (int k = 1; k < g_ntb2; ++k)
        if (!strcmp(g_tb2_shapes[k].name, name)) return k;
    return 0;
}

// The tile used when the caller did not choose (tb2_shape == -1): the measured best on B200
// (profiles/README.md section 7) for lattices large enough to be HBM-bound; small lattices keep the
// graph-batched one-step kernel.  0 = one-step kernel.
static int tb2_auto_shape(const lb_sim *sim)
{
    if (!tb2_servable(sim) || g_variants[sim->variant].launch_tma) return 0;
    if ((long long)sim->cfg.nx * sim->cfg.ny < (1ll << 22) || sim->cfg.ny < 64) return 0;
    const int span = sim->elem == 4 ? 128 : 64;
    if (sim->cfg.bc == LB_BC_PERIODIC && sim->cfg.nx % span) return 0;
    const int k = tb2_find(sim->elem == 4 ? LB_TB2_AUTO_F32 : LB_TB2_AUTO_F64);
    if (k > 0 && sim->cfg.ny > 65535 * g_tb2_shapes[k].by) return 0;      // grid.y limit: one-step kernel (3-D grid)
    return k;
}

static int tb2_effective_shape(const lb_sim *sim)
{
    if (sim->tb2_shape >= 0) return tb2_servable(sim) ? sim->tb2_shape : 0;
    return tb2_auto_shape(sim);
}

// two moment-free steps: reads buffer src_idx, writes the other one
static int launch_two_steps(lb_sim *sim, int src_idx, int shape)
{
    const Tb2Shape &t = g_tb2_shapes[shape];
    const size_t smem = tb2_smem_bytes(sim, shape);
    const int di = sim->cfg.dtype == LB_F64, mi = sim->cfg.math == LB_MATH_FAST;
    if (tb2_is_rows(shape)) {
        StepParams p;
        fill_params(sim, p, src_idx, 0, sim->state_index);
        const int span = sim->elem == 4 ? 128 : 64;
        const dim3 grid(sim->pitch / span, (sim->cfg.ny + t.by - 1) / t.by);
        t.launch_v[di][mi](p, grid, smem, sim->stream);
    } else {
        Tb2Params p{};
        p.src = sim->buf[src_idx]; p.dst = sim->buf[src_idx ^ 1];
        p.plane = sim->plane; p.nx = sim->cfg.nx; p.ny = sim->cfg.ny; p.pitch = sim->pitch;
        p.bc = sim->cfg.bc == LB_BC_PERIODIC ? BC_PERIODIC : BC_PIPE;
        p.zero_obstacle_velocity = sim->cfg.zero_obstacle_velocity;
        p.mask = sim->mask; p.mask_pitch = sim->mask_pitch;
        p.cf = consts_of<float>(sim); p.cd = consts_of<double>(sim);
        const dim3 grid((sim->cfg.nx + t.bx - 1) / t.bx, (sim->cfg.ny + t.by - 1) / t.by);
        t.launch[di][mi](p, grid, smem, sim->stream);
    }
    CU(cudaGetLastError());
    sim->launches++;
    return LB_OK;
}

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int lb_temporal_blocking(const lb_sim *sim) { return sim ? tb2_effective_shape(sim) : 0; }
int lb_tb2_shape_count(void) { return g_ntb2; }
const char *lb_tb2_shape_name(int shape) { return (shape >= 0 && shape < g_ntb2) ? g_tb2_shapes[shape].name : nullptr; }

int lb_set_temporal_blocking(lb_sim *sim, int shape)
{
    if (!sim) return LB_ERR_INVALID;
    if (shape == -1) { sim->tb2_shape = -1; return LB_OK; }
    if (shape < 0 || shape >= g_ntb2) return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: unknown tile shape");
    if (shape > 0) {
        if (sim->cfg.scheme != LB_SCHEME_OPENCL || sim->cfg.model != LB_MODEL_D2Q9 || uses_halo(sim) || sim->cfg.global_nx != sim->cfg.nx)
            return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: serves single-slab LB_SCHEME_OPENCL / LB_MODEL_D2Q9 lattices");
        if (sim->cfg.ny > 65535 * g_tb2_shapes[shape].by) return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: lattice too tall for this tile");
        if (tb2_smem_bytes(sim, shape) > 227 * 1024) return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: tile does not fit shared memory for this dtype");
        if (tb2_is_rows(shape) && sim->cfg.bc == LB_BC_PERIODIC && sim->cfg.nx % (sim->elem == 4 ? 128 : 64))
            return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: the row-per-warp tiles need nx to be a multiple of the tile width on a periodic box");
    }
    sim->tb2_shape = shape;
    return LB_OK;
}

int lb_abi_version(void) { return LB_ABI_VERSION; }

int lb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *lb_last_error(const lb_sim *sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int lb_selftest_rcp(int device, uint32_t first_bits, uint32_t last_bits, uint64_t *mismatches)
{
    if (!mismatches || last_bits < first_bits) return LB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return LB_ERR_CUDA;
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc((void **)&d, sizeof(h)) != cudaSuccess) return LB_ERR_CUDA;
    cudaMemset(d, 0, sizeof(h));
    k_selftest_rcp<<<148 * 16, 256>>>(first_bits, last_bits, d);
    const cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return LB_ERR_CUDA;
    *mismatches = h;
    return LB_OK;
}

int lb_selftest_copy(lb_sim *sim, int reps, double *ms_per_launch)
{
    if (!sim || !ms_per_launch || reps < 1) return fail(sim, LB_ERR_INVALID, "lb_selftest_copy: bad argument");
    CU(cudaSetDevice(sim->cfg.device));
    const int ny = sim->cfg.ny;
    const int span = sim->elem == 4 ? 128 : 64;
    const unsigned tx = (sim->pitch + 2 * span - 1) / (2 * span), ty = (ny + 1) / 2;
    const unsigned gy = ty < 65535 ? ty : 65535;
    const dim3 grid(tx, gy, (ty + gy - 1) / gy);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    auto launch = [&]() {
        if (sim->elem == 4) k_copy_pattern<float, 4><<<grid, 128, 0, sim->stream>>>((const float *)sim->buf[sim->cur], (float *)sim->buf[sim->cur ^ 1], sim->pitch, ny, sim->plane);
        else k_copy_pattern<double, 2><<<grid, 128, 0, sim->stream>>>((const double *)sim->buf[sim->cur], (double *)sim->buf[sim->cur ^ 1], sim->pitch, ny, sim->plane);
    };
    launch();                                            // warm-up
    cudaEventRecord(e0, sim->stream);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1, sim->stream);
    cudaError_t e = cudaStreamSynchronize(sim->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CU(e);
    *ms_per_launch = (double)ms / reps;
    return LB_OK;
}

int lb_variant_count(void) { return g_nvariants; }
const char *lb_variant_name(int v) { return (v >= 0 && v < g_nvariants) ? g_variants[v].name : nullptr; }

int lb_create(const lb_config *cfg, lb_sim **out)
{
    lb_sim *sim = nullptr;
    if (!cfg || !out) return fail(nullptr, LB_ERR_INVALID, "lb_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(lb_config))
        return fail(nullptr, LB_ERR_INVALID, "lb_create: lb_config size mismatch (ABI)");
    if (cfg->nx < 2 || cfg->ny < 2) return fail(nullptr, LB_ERR_INVALID, "lb_create: nx, ny must be >= 2");
    if (cfg->dtype != LB_F32 && cfg->dtype != LB_F64) return fail(nullptr, LB_ERR_INVALID, "lb_create: bad dtype");
    if (cfg->bc != LB_BC_PIPE && cfg->bc != LB_BC_PERIODIC && cfg->bc != LB_BC_VELOCITY_YPERIODIC)
        return fail(nullptr, LB_ERR_INVALID, "lb_create: bad bc");
    if (cfg->bc == LB_BC_VELOCITY_YPERIODIC &&
        ((cfg->scheme != LB_SCHEME_CYTHON_OLD && cfg->scheme != LB_SCHEME_OPENCL_OLD) || cfg->ny < 4))
        return fail(nullptr, LB_ERR_INVALID, "lb_create: LB_BC_VELOCITY_YPERIODIC needs LB_SCHEME_CYTHON_OLD or LB_SCHEME_OPENCL_OLD and ny >= 4");
    if (cfg->scheme == LB_SCHEME_OPENCL_OLD && cfg->bc != LB_BC_VELOCITY_YPERIODIC)
        return fail(nullptr, LB_ERR_INVALID, "lb_create: LB_SCHEME_OPENCL_OLD serves LB_BC_VELOCITY_YPERIODIC only (the pressure-driven classes of OLD/opencl.py diverge as shipped)");
    if (cfg->math != LB_MATH_STRICT && cfg->math != LB_MATH_FAST) return fail(nullptr, LB_ERR_INVALID, "lb_create: bad math");
    for (int e : {cfg->west_edge, cfg->east_edge}) {
        if (e < LB_EDGE_BOUNDARY || e > LB_EDGE_HALO) return fail(nullptr, LB_ERR_INVALID, "lb_create: bad edge kind");
        if (cfg->bc == LB_BC_PERIODIC && e == LB_EDGE_BOUNDARY)
            return fail(nullptr, LB_ERR_INVALID, "lb_create: a periodic box needs WRAP or HALO edges");
        if (cfg->bc != LB_BC_PERIODIC && e == LB_EDGE_WRAP)
            return fail(nullptr, LB_ERR_INVALID, "lb_create: pipe flow cannot wrap in x");
    }
    if (cfg->global_nx < cfg->nx || cfg->x_offset < 0 || cfg->x_offset + cfg->nx > cfg->global_nx)
        return fail(nullptr, LB_ERR_INVALID, "lb_create: slab does not fit the global lattice");
    if (!(cfg->omega > 0.0 && cfg->omega < 2.0)) return fail(nullptr, LB_ERR_INVALID, "lb_create: omega must be in (0,2)");
    if (cfg->scheme < LB_SCHEME_OPENCL || cfg->scheme > LB_SCHEME_OPENCL_OLD)
        return fail(nullptr, LB_ERR_INVALID, "lb_create: bad scheme");
    if (cfg->model != LB_MODEL_D2Q9 && cfg->model != LB_MODEL_D2Q9I) return fail(nullptr, LB_ERR_INVALID, "lb_create: bad model");
    if (cfg->model == LB_MODEL_D2Q9I && (cfg->scheme != LB_SCHEME_OPENCL || cfg->bc != LB_BC_PIPE))
        return fail(nullptr, LB_ERR_INVALID, "lb_create: the D2Q9i model exists for LB_SCHEME_OPENCL pipe flow only");
    if (cfg->scheme != LB_SCHEME_OPENCL &&
        (cfg->dtype != LB_F32 || cfg->bc == LB_BC_PERIODIC || cfg->west_edge != LB_EDGE_BOUNDARY ||
         cfg->east_edge != LB_EDGE_BOUNDARY || cfg->global_nx != cfg->nx))
        return fail(nullptr, LB_ERR_INVALID, "lb_create: the cython and opencl_old schemes need dtype F32, a non-periodic bc and a single slab");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, LB_ERR_CUDA, "lb_create: no CUDA device (this library has no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LB_ERR_INVALID, "lb_create: bad device ordinal");

    sim = new lb_sim();
    sim->cfg = *cfg;
    sim->elem = cfg->dtype == LB_F32 ? 4 : 8;
    sim->uv_elem = is_cython(sim) ? 8 : sim->elem;
    const int per512 = 512 / sim->elem;
    sim->pitch = (cfg->nx + per512 - 1) / per512 * per512;
    sim->plane = (long long)sim->pitch * cfg->ny;
    sim->variant = default_variant(cfg->dtype, cfg->math, cfg->model);
    auto bail = [&](int code, const std::string &m) { g_create_error = m; lb_destroy(sim); return code; };
#define CUC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return bail(LB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
    CUC(cudaSetDevice(cfg->device));
    if (cfg->stream) sim->stream = (cudaStream_t)cfg->stream;
    else { CUC(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking)); sim->own_stream = true; }
    const size_t guard = (size_t)2 * sim->pitch * sim->elem;
    sim->buf_bytes = guard * 2 + (size_t)9 * sim->plane * sim->elem;
    for (int i = 0; i < 2; ++i) {
        CUC(cudaMalloc((void **)&sim->buf_base[i], sim->buf_bytes));
        CUC(cudaMemsetAsync(sim->buf_base[i], 0, sim->buf_bytes, sim->stream));
        sim->buf[i] = sim->buf_base[i] + guard;
    }
    const size_t mom = (size_t)sim->plane * sim->elem, mom_uv = (size_t)sim->plane * sim->uv_elem;
    CUC(cudaMalloc(&sim->rho, mom)); CUC(cudaMalloc(&sim->u, mom_uv)); CUC(cudaMalloc(&sim->v, mom_uv));
    CUC(cudaMemsetAsync(sim->rho, 0, mom, sim->stream));
    CUC(cudaMemsetAsync(sim->u, 0, mom_uv, sim->stream));
    CUC(cudaMemsetAsync(sim->v, 0, mom_uv, sim->stream));
    CUC(cudaMalloc((void **)&sim->mass_scratch, sizeof(double) * cfg->ny));
    if (is_oldcl(sim)) {
        CUC(cudaMalloc((void **)&sim->frozen, sizeof(float) * oc_frozen_floats(cfg->nx, cfg->ny)));
        CUC(cudaMemsetAsync(sim->frozen, 0, sizeof(float) * oc_frozen_floats(cfg->nx, cfg->ny), sim->stream));
    }
    if (cfg->west_edge == LB_EDGE_HALO || cfg->east_edge == LB_EDGE_HALO) {
        sim->hl = halo_layout(cfg->ny, sim->elem);
        CUC(cudaMalloc((void **)&sim->halo, sim->hl.total));
        CUC(cudaMemsetAsync(sim->halo, 0, sim->hl.total, sim->stream));
    }
    CUC(cudaStreamSynchronize(sim->stream));
#undef CUC
    *out = sim;
    return LB_OK;
}

int lb_destroy(lb_sim *sim)
{
    if (!sim) return LB_OK;
    cudaSetDevice(sim->cfg.device);
    if (sim->stream) cudaStreamSynchronize(sim->stream);
    drop_graphs(sim);
    for (int side = 0; side < 2; ++side)
        if (sim->peer[side] && sim->peer_ipc[side]) cudaIpcCloseMemHandle(sim->peer[side]);
    for (int i = 0; i < 2; ++i) cudaFree(sim->buf_base[i]);
    cudaFree(sim->rho); cudaFree(sim->u); cudaFree(sim->v); cudaFree(sim->feq);
    cudaFree(sim->mask); cudaFree(sim->span_solid); cudaFree(sim->halo); cudaFree(sim->mass_scratch); cudaFree(sim->frozen);
    if (sim->own_stream && sim->stream) cudaStreamDestroy(sim->stream);
    cudaGetLastError();
    delete sim;
    return LB_OK;
}

int lb_set_variant(lb_sim *sim, int variant)
{
    if (!sim) return LB_ERR_INVALID;
    const bool back_to_default = variant < 0;
    if (variant < 0) variant = default_variant(sim->cfg.dtype, sim->cfg.math, sim->cfg.model);
    if (variant >= g_nvariants || g_variants[variant].dtype != sim->cfg.dtype || g_variants[variant].math != sim->cfg.math ||
        g_variants[variant].model != sim->cfg.model)
        return fail(sim, LB_ERR_INVALID, "lb_set_variant: variant does not match the handle's dtype/math/model");
    if (g_variants[variant].launch_tma && (uses_halo(sim) || sim->cfg.bc == LB_BC_PERIODIC || sim->cfg.scheme != LB_SCHEME_OPENCL))
        return fail(sim, LB_ERR_INVALID, "lb_set_variant: the TMA-staged kernel serves single-slab, non-periodic lattices");
    sim->variant = variant;
    if (back_to_default) sim->tb2_shape = -1;
    else if (sim->tb2_shape < 0) sim->tb2_shape = 0; // a hand-picked one-step variant is what runs
    drop_graphs(sim);
    return LB_OK;
}

int64_t lb_launch_count(const lb_sim *sim) { return sim ? sim->launches : 0; }

int lb_set_mask(lb_sim *sim, const void *host_mask, int elem_bytes)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    drop_graphs(sim);
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    if (!host_mask) {
        CU(cudaStreamSynchronize(sim->stream));
        cudaFree(sim->mask); cudaFree(sim->span_solid);
        sim->mask = nullptr; sim->span_solid = nullptr;
        return LB_OK;
    }
    if (elem_bytes != 1 && elem_bytes != 4) return fail(sim, LB_ERR_INVALID, "lb_set_mask: elem_bytes must be 1 or 4");
    std::vector<uint8_t> packed((size_t)nx * ny);
    if (elem_bytes == 1) {
        const uint8_t *m = (const uint8_t *)host_mask;
        for (size_t i = 0; i < packed.size(); ++i) packed[i] = (m[i] == 1) ? 1 : 0;
    } else {
        const int32_t *m = (const int32_t *)host_mask;
        for (size_t i = 0; i < packed.size(); ++i) packed[i] = (m[i] == 1) ? 1 : 0;
    }
    if (sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC && is_cython(sim))      // lb_cython.cuh folds the row exchange into the pull
        for (int x = 0; x < nx; ++x)
            if (packed[x] || packed[(size_t)(ny - 1) * nx + x])
                return fail(sim, LB_ERR_INVALID, "lb_set_mask: with LB_BC_VELOCITY_YPERIODIC the exchanged rows y=0 and y=ny-1 must be free of solid nodes");
    if (!sim->mask) {
        sim->mask_pitch = sim->pitch;
        sim->nspans = sim->pitch / 32;
        CU(cudaMalloc((void **)&sim->mask, (size_t)sim->mask_pitch * ny));
        CU(cudaMalloc((void **)&sim->span_solid, (size_t)sim->nspans * ny));
    }
    CU(cudaMemsetAsync(sim->mask, 0, (size_t)sim->mask_pitch * ny, sim->stream));
    CU(cudaMemcpy2DAsync(sim->mask, sim->mask_pitch, packed.data(), nx, nx, ny, cudaMemcpyHostToDevice, sim->stream));
    k_span_solid<<<dim3((sim->nspans + 63) / 64, ny), 64, 0, sim->stream>>>(nx, ny, sim->mask, sim->mask_pitch,
                                                                            sim->span_solid, sim->nspans);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sim->stream));   // `packed` is about to go out of scope
    return LB_OK;
}

int lb_set_mask_disk(lb_sim *sim, double cx, double cy, double r)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    drop_graphs(sim);
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    if (!sim->mask) {
        sim->mask_pitch = sim->pitch;
        sim->nspans = sim->pitch / 32;
        CU(cudaMalloc((void **)&sim->mask, (size_t)sim->mask_pitch * ny));
        CU(cudaMalloc((void **)&sim->span_solid, (size_t)sim->nspans * ny));
    }
    CU(cudaMemsetAsync(sim->mask, 0, (size_t)sim->mask_pitch * ny, sim->stream));
    k_mask_disk<<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->cfg.x_offset, cx, cy, r * r, sim->mask, sim->mask_pitch);
    k_span_solid<<<dim3((sim->nspans + 63) / 64, ny), 64, 0, sim->stream>>>(nx, ny, sim->mask, sim->mask_pitch,
                                                                            sim->span_solid, sim->nspans);
    CU(cudaGetLastError());
    return LB_OK;
}

int lb_upload_f(lb_sim *sim, const void *host_f)
{
    if (!sim || !host_f) return fail(sim, LB_ERR_INVALID, "lb_upload_f: null argument");
    CU(cudaSetDevice(sim->cfg.device));
    const size_t w = (size_t)sim->cfg.nx * sim->elem, dp = (size_t)sim->pitch * sim->elem;
    // planes are contiguous (plane = ny*pitch), so one 2-D copy of 9*ny rows does all nine
    CU(cudaMemcpy2DAsync(sim->buf[sim->cur], dp, host_f, w, w, (size_t)9 * sim->cfg.ny, cudaMemcpyHostToDevice, sim->stream));
    // the reference seeds f_streamed with the same data (opencl_dim.py:324-327)
    CU(cudaMemcpyAsync(sim->buf_base[sim->cur ^ 1], sim->buf_base[sim->cur], sim->buf_bytes, cudaMemcpyDeviceToDevice, sim->stream));
    if (is_oldcl(sim)) {
        const int n = sim->cfg.nx > sim->cfg.ny ? sim->cfg.nx : sim->cfg.ny;
        oc_capture_frozen_kernel<<<(n + 127) / 128, 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
                                                                            (const float *)sim->buf[sim->cur], sim->frozen);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(sim->stream));
    sim->prestream_done = false;
    return LB_OK;
}

int lb_upload_moments(lb_sim *sim, const void *host_rho, const void *host_u, const void *host_v)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    const void *hs[3] = {host_rho, host_u, host_v};
    void *ds[3] = {sim->rho, sim->u, sim->v};
    for (int i = 0; i < 3; ++i) {
        const int eb = i == 0 ? sim->elem : sim->uv_elem;      // u, v are float64 for the cython schemes
        const size_t w = (size_t)sim->cfg.nx * eb, dp = (size_t)sim->pitch * eb;
        if (hs[i]) CU(cudaMemcpy2DAsync(ds[i], dp, hs[i], w, w, sim->cfg.ny, cudaMemcpyHostToDevice, sim->stream));
    }
    CU(cudaStreamSynchronize(sim->stream));
    return LB_OK;
}

static int ensure_feq(lb_sim *sim)
{
    if (!sim->feq) {
        CU(cudaMalloc(&sim->feq, (size_t)9 * sim->plane * sim->elem));
        CU(cudaMemsetAsync(sim->feq, 0, (size_t)9 * sim->plane * sim->elem, sim->stream));
    }
    return LB_OK;
}

static int compute_feq(lb_sim *sim)
{
    int rc = ensure_feq(sim);
    if (rc) return rc;
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    if (is_cython(sim))
        cy_feq_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, (const float *)sim->rho,
            (const double *)sim->u, (const double *)sim->v, (float *)sim->feq, cy_consts_of(sim));
    else if (sim->cfg.dtype == LB_F32)
        k_feq_from_moments<float><<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, (const float *)sim->rho,
            (const float *)sim->u, (const float *)sim->v, (float *)sim->feq, consts_of<float>(sim), sim->cfg.model);
    else
        k_feq_from_moments<double><<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, (const double *)sim->rho,
            (const double *)sim->u, (const double *)sim->v, (double *)sim->feq, consts_of<double>(sim), sim->cfg.model);
    CU(cudaGetLastError());
    return LB_OK;
}

// ---- the hot path ---------------------------------------------------------------------
// scheme "cython": see lb_cython.cuh for the fusion order
static int cython_steps(lb_sim *sim, int n_steps)
{
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    const CyConsts c = cy_consts_of(sim);
    if (!sim->prestream_done) {
        cy_prestream_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, (float *)sim->buf[sim->cur],
                                                                   (const double *)sim->u, sim->mask, sim->mask_pitch, c,
                                                                   sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC);
        CU(cudaGetLastError());
    }
    constexpr int WX = 2, WY = 2;
    const unsigned tiles_x = (sim->pitch + 128 * WX - 1) / (128 * WX), tiles_y = (ny + WY - 1) / WY;
    const unsigned gy = tiles_y < 65535 ? tiles_y : 65535;
    const dim3 grid(tiles_x, gy, (tiles_y + gy - 1) / gy);
    for (int i = 0; i < n_steps; ++i) {
        const bool last = (i == n_steps - 1);
        CyParams p;
        p.src = (const float *)sim->buf[sim->cur];
        p.dst = (float *)sim->buf[sim->cur ^ 1];
        p.plane = sim->plane; p.nx = nx; p.ny = ny; p.pitch = sim->pitch;
        p.write_moments = last; p.apply_next_bc = !last;
        p.mask = sim->mask; p.mask_pitch = sim->mask_pitch;
        p.rho = (float *)sim->rho; p.u = (double *)sim->u; p.v = (double *)sim->v;
        p.c = c;
        if (sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC) fused_step_cython_kernel<true, true, WX, WY, 4><<<grid, 32 * WX * WY, 0, sim->stream>>>(p);
        else if (sim->cfg.scheme == LB_SCHEME_CYTHON_OLD) fused_step_cython_kernel<true, false, WX, WY, 4><<<grid, 32 * WX * WY, 0, sim->stream>>>(p);
        else fused_step_cython_kernel<false, false, WX, WY, 4><<<grid, 32 * WX * WY, 0, sim->stream>>>(p);
        CU(cudaGetLastError());
        sim->launches++;
        sim->cur ^= 1; sim->state_index++;
    }
    sim->prestream_done = false;      // the last launch left plain post-collision populations
    return LB_OK;
}

// scheme "opencl_old": see lb_oldcl.cuh for the fusion order
static int oldcl_steps(lb_sim *sim, int n_steps)
{
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    OcParams p = oc_params_of(sim);
    const unsigned row_blocks = (nx + 127) / 128;
    if (!sim->prestream_done) {
        oc_prestream_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(p, (float *)sim->buf[sim->cur]);
        oc_rows_kernel<<<row_blocks, 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, (float *)sim->buf[sim->cur],
                                                            sim->mask, sim->mask_pitch);
        CU(cudaGetLastError());
    }
    constexpr int WX = 2, WY = 2;
    const unsigned tiles_x = (sim->pitch + 128 * WX - 1) / (128 * WX), tiles_y = (ny + WY - 1) / WY;
    const unsigned gy = tiles_y < 65535 ? tiles_y : 65535;
    const dim3 grid(tiles_x, gy, (tiles_y + gy - 1) / gy);
    for (int i = 0; i < n_steps; ++i) {
        const bool last = (i == n_steps - 1);
        p.src = (const float *)sim->buf[sim->cur];
        p.dst = (float *)sim->buf[sim->cur ^ 1];
        p.write_moments = last; p.apply_next_bc = !last;
        fused_step_oldcl_kernel<WX, WY, 4><<<grid, 32 * WX * WY, 0, sim->stream>>>(p);
        if (!last)
            oc_rows_kernel<<<row_blocks, 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, p.dst, sim->mask, sim->mask_pitch);
        CU(cudaGetLastError());
        sim->launches++;
        sim->cur ^= 1; sim->state_index++;
    }
    sim->prestream_done = false;      // the last launch left plain post-collision populations
    return LB_OK;
}

static const int GRAPH_LEN = 32;    // steps per captured graph (even: a graph returns to its start buffer)

// (re)build the graph of GRAPH_LEN moment-free steps that starts from buffer `sim->cur`
static int ensure_graph(lb_sim *sim)
{
    const int gi = sim->cur;
    if (sim->graph[gi] && sim->graph_variant[gi] == sim->variant) return LB_OK;
    if (sim->graph[gi]) { cudaGraphExecDestroy(sim->graph[gi]); sim->graph[gi] = nullptr; }
    cudaGraph_t g = nullptr;
    CU(cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal));
    int idx = sim->cur, rc = LB_OK;
    const int64_t l0 = sim->launches;
    for (int i = 0; i < GRAPH_LEN && rc == LB_OK; ++i) { rc = launch_step(sim, idx, 0, 0); idx ^= 1; }
    sim->launches = l0;                  // captured, not executed
    cudaError_t ce = cudaStreamEndCapture(sim->stream, &g);
    if (rc != LB_OK) { if (g) cudaGraphDestroy(g); return rc; }
    CU(ce);
    ce = cudaGraphInstantiate(&sim->graph[gi], g, 0);
    cudaGraphDestroy(g);
    CU(ce);
    sim->graph_len[gi] = GRAPH_LEN;
    sim->graph_variant[gi] = sim->variant;
    return LB_OK;
}

int lb_step(lb_sim *sim, int n_steps)
{
    if (!sim) return LB_ERR_INVALID;
    if (n_steps < 0) return fail(sim, LB_ERR_INVALID, "lb_step: negative step count");
    if (n_steps == 0) return LB_OK;
    CU(cudaSetDevice(sim->cfg.device));
    if (uses_halo(sim)) {
        for (int side = 0; side < 2; ++side) {
            const int e = side == LB_WEST ? sim->cfg.west_edge : sim->cfg.east_edge;
            if (e == LB_EDGE_HALO && !sim->peer[side]) return fail(sim, LB_ERR_STATE, "lb_step: halo edge not connected");
        }
    }
    if (is_oldcl(sim)) return oldcl_steps(sim, n_steps);
    if (sim->cfg.scheme != LB_SCHEME_OPENCL) return cython_steps(sim, n_steps);
    int remaining = n_steps - 1;            // all but the last step skip the moment stores
    if (const int tb2 = tb2_effective_shape(sim)) {   // temporal blocking: moment-free steps two at a time
        for (; remaining >= 2; remaining -= 2) {
            int rc = launch_two_steps(sim, sim->cur, tb2);
            if (rc) return rc;
            sim->cur ^= 1; sim->state_index += 2;
        }
    }
    const bool graphs_ok = !uses_halo(sim); // halo launches carry a per-step flag value
    while (graphs_ok && remaining >= GRAPH_LEN) {
        int rc = ensure_graph(sim);
        if (rc) return rc;
        CU(cudaGraphLaunch(sim->graph[sim->cur], sim->stream));
        sim->launches += GRAPH_LEN;
        sim->state_index += GRAPH_LEN;
        remaining -= GRAPH_LEN;
    }
    for (; remaining > 0; --remaining) {
        int rc = launch_step(sim, sim->cur, 0, sim->state_index);
        if (rc) return rc;
        sim->cur ^= 1; sim->state_index++;
    }
    int rc = launch_step(sim, sim->cur, 1, sim->state_index);
    if (rc) return rc;
    sim->cur ^= 1; sim->state_index++;
    return LB_OK;
}

// ---- L2-level temporal blocking (experimental) ---------------------------------
// `depth` consecutive steps travel down the lattice together: step s works on row band b - s, each band
// shifted up by s rows, so stream order alone satisfies "row y of step s+1 needs rows y-1..y+1 of step s" and
// "step s+1 may overwrite a row of the buffer step s reads only after step s is done with it".  What step s
// wrote is read by step s+1 a band later, while it is still in L2, and the intermediate time levels are
// overwritten in L2 (the ping-pong buffers alias them) before they are ever written back.
int lb_step_banded(lb_sim *sim, int n_steps, int band_rows, int depth)
{
    if (!sim) return LB_ERR_INVALID;
    if (n_steps < 0 || band_rows < 1 || depth < 1) return fail(sim, LB_ERR_INVALID, "lb_step_banded: bad argument");
    if (sim->cfg.scheme != LB_SCHEME_OPENCL || uses_halo(sim) || sim->cfg.bc == LB_BC_PERIODIC || g_variants[sim->variant].launch_tma)
        return fail(sim, LB_ERR_INVALID, "lb_step_banded: serves single-slab, non-periodic LB_SCHEME_OPENCL lattices (register-shuffle kernel)");
    if (n_steps == 0) return LB_OK;
    CU(cudaSetDevice(sim->cfg.device));
    const int ny = sim->cfg.ny;
    int remaining = n_steps - 1;                       // the last step stores the moments: whole-lattice launch
    const Variant &var = g_variants[sim->variant];
    while (remaining > 0) {
        const int k = remaining < depth ? remaining : depth;
        const int nb = (ny + k + band_rows - 1) / band_rows;          // bands of the most shifted step reach row ny
        for (int b = 0; b < nb + k - 1; ++b) {
            for (int s = 0; s < k; ++s) {
                const int band = b - s;
                if (band < 0 || band >= nb) continue;
                int y0 = band * band_rows - s, y1 = y0 + band_rows;
                if (y0 < 0) y0 = 0;
                if (y1 > ny) y1 = ny;
                if (band == nb - 1) y1 = ny;
                if (y0 >= y1) continue;
                StepParams p;
                fill_params(sim, p, sim->cur ^ (s & 1), 0, sim->state_index + s);
                p.y_begin = y0; p.y_end = y1;
                var.launch(p, sim->stream);
                sim->launches++;
            }
        }
        CU(cudaGetLastError());
        sim->cur ^= (k & 1);
        sim->state_index += k;
        remaining -= k;
    }
    int rc = launch_step(sim, sim->cur, 1, sim->state_index);
    if (rc) return rc;
    sim->cur ^= 1; sim->state_index++;
    return LB_OK;
}

int lb_sync(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    CU(cudaStreamSynchronize(sim->stream));
    if (sim->halo) {
        unsigned int e = 0;
        CU(cudaMemcpy(&e, sim->halo + sim->hl.off_error, sizeof(e), cudaMemcpyDeviceToHost));
        if (e) return fail(sim, LB_ERR_HALO, "halo hand-shake timed out (neighbour slab did not publish its boundary column)");
    }
    return LB_OK;
}

int lb_download(lb_sim *sim, int field, void *host_out)
{
    if (!sim || !host_out) return fail(sim, LB_ERR_INVALID, "lb_download: null argument");
    CU(cudaSetDevice(sim->cfg.device));
    size_t w = (size_t)sim->cfg.nx * sim->elem, dp = (size_t)sim->pitch * sim->elem;
    if (field == LB_FIELD_U || field == LB_FIELD_V) { w = (size_t)sim->cfg.nx * sim->uv_elem; dp = (size_t)sim->pitch * sim->uv_elem; }
    const void *src = nullptr;
    size_t rows = sim->cfg.ny;
    switch (field) {
    case LB_FIELD_F: src = sim->buf[sim->cur]; rows *= 9; break;
    case LB_FIELD_FEQ: { int rc = compute_feq(sim); if (rc) return rc; src = sim->feq; rows *= 9; break; }
    case LB_FIELD_RHO: src = sim->rho; break;
    case LB_FIELD_U: src = sim->u; break;
    case LB_FIELD_V: src = sim->v; break;
    default: return fail(sim, LB_ERR_INVALID, "lb_download: unknown field");
    }
    CU(cudaMemcpy2DAsync(host_out, w, src, dp, w, rows, cudaMemcpyDeviceToHost, sim->stream));
    return lb_sync(sim);
}

int lb_download_strided(lb_sim *sim, int field, int stride_x, int stride_y, void *host_out)
{
    if (!sim || !host_out) return fail(sim, LB_ERR_INVALID, "lb_download_strided: null argument");
    if (stride_x < 1 || stride_y < 1) return fail(sim, LB_ERR_INVALID, "lb_download_strided: strides must be >= 1");
    const void *src = field == LB_FIELD_RHO ? sim->rho : field == LB_FIELD_U ? sim->u : field == LB_FIELD_V ? sim->v : nullptr;
    if (!src) return fail(sim, LB_ERR_INVALID, "lb_download_strided: field must be rho, u or v");
    CU(cudaSetDevice(sim->cfg.device));
    const int eb = field == LB_FIELD_RHO ? sim->elem : sim->uv_elem;
    const int ox = (sim->cfg.nx + stride_x - 1) / stride_x, oy = (sim->cfg.ny + stride_y - 1) / stride_y;
    void *tmp = nullptr;
    CU(cudaMallocAsync(&tmp, (size_t)ox * oy * eb, sim->stream));
    const dim3 grid((ox + 127) / 128, oy);
    if (eb == 4) k_subsample<float><<<grid, 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, stride_x, stride_y, ox, (const float *)src, (float *)tmp);
    else k_subsample<double><<<grid, 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, stride_x, stride_y, ox, (const double *)src, (double *)tmp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_out, tmp, (size_t)ox * oy * eb, cudaMemcpyDeviceToHost, sim->stream));
    CU(cudaFreeAsync(tmp, sim->stream));
    return lb_sync(sim);
}

void *lb_stream(lb_sim *sim) { return sim ? (void *)sim->stream : nullptr; }

int lb_device_ptr(lb_sim *sim, int field, void **ptr, int64_t *pitch_elems)
{
    if (!sim || !ptr) return LB_ERR_INVALID;
    switch (field) {
    case LB_FIELD_F: *ptr = sim->buf[sim->cur]; break;
    case LB_FIELD_FEQ: *ptr = sim->feq; break;
    case LB_FIELD_RHO: *ptr = sim->rho; break;
    case LB_FIELD_U: *ptr = sim->u; break;
    case LB_FIELD_V: *ptr = sim->v; break;
    default: return fail(sim, LB_ERR_INVALID, "lb_device_ptr: unknown field");
    }
    if (pitch_elems) *pitch_elems = sim->pitch;
    return LB_OK;
}

// ---- single stages --------------------------------------------------------------------
#define DISPATCH(KERNEL, GRID, ...)                                                                 \
    do {                                                                                           \
        if (sim->cfg.dtype == LB_F32) { typedef float T; KERNEL<T><<<GRID, 128, 0, sim->stream>>>(__VA_ARGS__); } \
        else { typedef double T; KERNEL<T><<<GRID, 128, 0, sim->stream>>>(__VA_ARGS__); }           \
        CU(cudaGetLastError());                                                                    \
    } while (0)

int lb_stage_move(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    if (uses_halo(sim)) return fail(sim, LB_ERR_STATE, "single stages are not available on halo-connected slabs");
    CU(cudaSetDevice(sim->cfg.device));
    if (is_cython(sim) || is_oldcl(sim)) {
        const float *src = (const float *)sim->buf[sim->cur];
        float *dst = (float *)sim->buf[sim->cur ^ 1];
        if (is_cython(sim)) cy_stage_move_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, src, dst);
        else oc_stage_move_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, src, dst, sim->frozen);
        CU(cudaGetLastError());
        sim->cur ^= 1;
        sim->prestream_done = false;
        return LB_OK;
    }
    DISPATCH(k_stage_move, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, sim->cfg.bc == LB_BC_PERIODIC,
             (const T *)sim->buf[sim->cur], (T *)sim->buf[sim->cur ^ 1]);
    sim->cur ^= 1;
    return LB_OK;
}

int lb_stage_move_bcs(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    if (is_cython(sim)) {          // cython_dim.pyx:204-269 (+ :468-513 with a mask), OLD/cython.pyx:278-316
        const bool vin = sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC;
        cy_prestream_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
            (float *)sim->buf[sim->cur], (const double *)sim->u, sim->mask, sim->mask_pitch, cy_consts_of(sim), vin);
        if (vin) cyv_rows_kernel<<<(sim->cfg.nx + 127) / 128, 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
                                                                                       (float *)sim->buf[sim->cur]);
        CU(cudaGetLastError());
        return LB_OK;
    }
    if (is_oldcl(sim)) {           // D2Q9.cl:263-321 + :398-433 as OLD/opencl.py:290-297, :365-371 launches them
        oc_prestream_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(oc_params_of(sim), (float *)sim->buf[sim->cur]);
        oc_rows_kernel<<<(sim->cfg.nx + 127) / 128, 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
                                                                           (float *)sim->buf[sim->cur], sim->mask, sim->mask_pitch);
        CU(cudaGetLastError());
        return LB_OK;
    }
    DISPATCH(k_stage_bcs, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, sim->cfg.global_nx, sim->cfg.x_offset,
             sim->cfg.bc == LB_BC_PIPE, sim->mask, sim->mask_pitch, (T *)sim->buf[sim->cur], consts_of<T>(sim), sim->cfg.model);
    return LB_OK;
}

int lb_stage_update_hydro(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    if (is_cython(sim)) {
        const int nx = sim->cfg.nx, ny = sim->cfg.ny;
        const float *f = (const float *)sim->buf[sim->cur];
        float *rho = (float *)sim->rho;
        double *u = (double *)sim->u, *v = (double *)sim->v;
        const CyConsts c = cy_consts_of(sim);
        if (sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC)
            cy_stage_hydro_kernel<true, true><<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, f, rho, u, v, sim->mask, sim->mask_pitch, c);
        else if (sim->cfg.scheme == LB_SCHEME_CYTHON_OLD)
            cy_stage_hydro_kernel<true, false><<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, f, rho, u, v, sim->mask, sim->mask_pitch, c);
        else
            cy_stage_hydro_kernel<false, false><<<grid2d(sim), 128, 0, sim->stream>>>(nx, ny, sim->pitch, sim->plane, f, rho, u, v, sim->mask, sim->mask_pitch, c);
        CU(cudaGetLastError());
        return LB_OK;
    }
    if (is_oldcl(sim)) {
        oc_stage_hydro_kernel<<<grid2d(sim), 128, 0, sim->stream>>>(oc_params_of(sim), (const float *)sim->buf[sim->cur]);
        CU(cudaGetLastError());
        return LB_OK;
    }
    DISPATCH(k_stage_hydro, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const T *)sim->buf[sim->cur],
             (T *)sim->rho, (T *)sim->u, (T *)sim->v, sim->cfg.model);
    if (sim->cfg.zero_obstacle_velocity && sim->mask) return lb_stage_zero_velocity(sim);
    return LB_OK;
}

int lb_stage_zero_velocity(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    if (!sim->mask) return LB_OK;
    CU(cudaSetDevice(sim->cfg.device));
    if (is_cython(sim)) {
        k_zero_velocity<double><<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->mask, sim->mask_pitch,
                                                                       (double *)sim->u, (double *)sim->v);
        CU(cudaGetLastError());
        return LB_OK;
    }
    DISPATCH(k_zero_velocity, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->mask, sim->mask_pitch, (T *)sim->u, (T *)sim->v);
    return LB_OK;
}

int lb_stage_update_feq(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    return compute_feq(sim);
}

int lb_stage_collide(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    if (!sim->feq) return fail(sim, LB_ERR_STATE, "lb_stage_collide: call lb_stage_update_feq first");
    if (is_cython(sim)) {
        if (sim->cfg.scheme == LB_SCHEME_CYTHON_OLD)
            cy_stage_collide_kernel<true><<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
                (float *)sim->buf[sim->cur], (const float *)sim->feq, cy_consts_of(sim));
        else
            cy_stage_collide_kernel<false><<<grid2d(sim), 128, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane,
                (float *)sim->buf[sim->cur], (const float *)sim->feq, cy_consts_of(sim));
        CU(cudaGetLastError());
        sim->prestream_done = false;
        return LB_OK;
    }
    DISPATCH(k_stage_collide, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (T *)sim->buf[sim->cur],
             (const T *)sim->feq, consts_of<T>(sim));
    return LB_OK;
}

// ---- synthetic initialisers / diagnostics -----------------------------------------------
int lb_init_synthetic(lb_sim *sim, int kind, double u0, double amplitude, uint64_t seed)
{
    if (!sim) return LB_ERR_INVALID;
    if (kind != LB_SYNTH_PIPE_RAMP && kind != LB_SYNTH_SHEAR_LAYERS) return fail(sim, LB_ERR_INVALID, "lb_init_synthetic: bad kind");
    if (sim->cfg.scheme != LB_SCHEME_OPENCL) return fail(sim, LB_ERR_STATE, "lb_init_synthetic: LB_SCHEME_OPENCL only");
    CU(cudaSetDevice(sim->cfg.device));
    DISPATCH(k_init_synth, grid2d(sim), sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, sim->cfg.global_nx, sim->cfg.x_offset, kind,
             u0, amplitude, (unsigned long long)seed, sim->cfg.inlet_rho, sim->cfg.outlet_rho, sim->mask, sim->mask_pitch,
             (T *)sim->buf[sim->cur], (T *)sim->buf[sim->cur ^ 1], (T *)sim->rho, (T *)sim->u, (T *)sim->v, consts_of<T>(sim), sim->cfg.model);
    return LB_OK;
}

int lb_total_mass(lb_sim *sim, double *out)
{
    if (!sim || !out) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    const int ny = sim->cfg.ny;
    if (sim->cfg.dtype == LB_F32)
        k_mass<float><<<ny, 256, 0, sim->stream>>>(sim->cfg.nx, ny, sim->pitch, sim->plane, (const float *)sim->buf[sim->cur], sim->mass_scratch);
    else
        k_mass<double><<<ny, 256, 0, sim->stream>>>(sim->cfg.nx, ny, sim->pitch, sim->plane, (const double *)sim->buf[sim->cur], sim->mass_scratch);
    CU(cudaGetLastError());
    std::vector<double> rows(ny);
    CU(cudaMemcpyAsync(rows.data(), sim->mass_scratch, sizeof(double) * ny, cudaMemcpyDeviceToHost, sim->stream));
    CU(cudaStreamSynchronize(sim->stream));
    long double acc = 0;
    for (double r : rows) acc += r;
    *out = (double)acc;
    return LB_OK;
}

int lb_checksum(lb_sim *sim, uint64_t *out)
{
    if (!sim || !out) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    unsigned long long *d = (unsigned long long *)sim->mass_scratch;      // ny doubles: room for one u64
    CU(cudaMemsetAsync(d, 0, sizeof(unsigned long long), sim->stream));
    if (sim->cfg.dtype == LB_F32)
        k_checksum<float><<<sim->cfg.ny, 256, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const float *)sim->buf[sim->cur], d);
    else
        k_checksum<double><<<sim->cfg.ny, 256, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const double *)sim->buf[sim->cur], d);
    CU(cudaGetLastError());
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, sim->stream));
    CU(cudaStreamSynchronize(sim->stream));
    *out = h;
    return LB_OK;
}

// ---- halo -----------------------------------------------------------------------------
int lb_halo_ipc_handle(lb_sim *sim, void *out_handle)
{
    if (!sim || !out_handle) return LB_ERR_INVALID;
    if (!sim->halo) return fail(sim, LB_ERR_STATE, "lb_halo_ipc_handle: this slab has no halo edge");
    static_assert(sizeof(cudaIpcMemHandle_t) == LB_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(sim->cfg.device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, sim->halo));
    memcpy(out_handle, &h, sizeof(h));
    return LB_OK;
}

int lb_halo_connect_ipc(lb_sim *sim, int side, const void *peer_handle, int peer_device)
{
    if (!sim || !peer_handle || (side != LB_WEST && side != LB_EAST)) return fail(sim, LB_ERR_INVALID, "lb_halo_connect_ipc: bad argument");
    (void)peer_device;
    CU(cudaSetDevice(sim->cfg.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, peer_handle, sizeof(h));
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    sim->peer[side] = (char *)p;
    sim->peer_ipc[side] = true;
    return LB_OK;
}

int lb_halo_connect_local(lb_sim *sim, int side, lb_sim *peer)
{
    if (!sim || !peer || (side != LB_WEST && side != LB_EAST)) return fail(sim, LB_ERR_INVALID, "lb_halo_connect_local: bad argument");
    if (!peer->halo) return fail(sim, LB_ERR_STATE, "lb_halo_connect_local: peer has no halo arena");
    if (peer->cfg.ny != sim->cfg.ny || peer->cfg.dtype != sim->cfg.dtype)
        return fail(sim, LB_ERR_INVALID, "lb_halo_connect_local: neighbour slabs must share ny and dtype");
    CU(cudaSetDevice(sim->cfg.device));
    if (peer->cfg.device != sim->cfg.device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, sim->cfg.device, peer->cfg.device));
        if (!can) return fail(sim, LB_ERR_CUDA, "lb_halo_connect_local: devices cannot access each other");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->cfg.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
    }
    sim->peer[side] = peer->halo;
    sim->peer_ipc[side] = false;
    return LB_OK;
}

int lb_halo_prime(lb_sim *sim)
{
    if (!sim) return LB_ERR_INVALID;
    if (!uses_halo(sim)) return LB_OK;
    CU(cudaSetDevice(sim->cfg.device));
    const HaloLayout &h = sim->hl;
    const int par = sim->state_index & 1;
    char *ow = (sim->cfg.west_edge == LB_EDGE_HALO && sim->peer[LB_WEST]) ? sim->peer[LB_WEST] + h.off_ghost_e[par] : nullptr;
    char *oe = (sim->cfg.east_edge == LB_EDGE_HALO && sim->peer[LB_EAST]) ? sim->peer[LB_EAST] + h.off_ghost_w[par] : nullptr;
    unsigned int *fw = ow ? (unsigned int *)(sim->peer[LB_WEST] + h.off_flag_e) : nullptr;
    unsigned int *fe = oe ? (unsigned int *)(sim->peer[LB_EAST] + h.off_flag_w) : nullptr;
    if ((sim->cfg.west_edge == LB_EDGE_HALO && !ow) || (sim->cfg.east_edge == LB_EDGE_HALO && !oe))
        return fail(sim, LB_ERR_STATE, "lb_halo_prime: halo edge not connected");
    if (sim->cfg.dtype == LB_F32)
        k_halo_prime<float><<<1, 1024, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const float *)sim->buf[sim->cur],
                                                          (float *)ow, (float *)oe, fw, fe, sim->state_index + 1);
    else
        k_halo_prime<double><<<1, 1024, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const double *)sim->buf[sim->cur],
                                                           (double *)ow, (double *)oe, fw, fe, sim->state_index + 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sim->stream));
    return LB_OK;
}

}  // extern "C"
