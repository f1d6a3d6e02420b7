// C-ABI implementation of include/lb_d2q9.h: handle management, the fused-step launcher
// (CUDA-graph batched), single-stage kernels, device-side initialisers and the NVLink
// peer-memory halo plumbing.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false
// (see __graft_entry__.build()).
#include "../../include/lb_d2q9.h"
#include "lb_host.h"
#include "lb_cython.cuh"
#include "lb_oldcl.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace lb;

// =====================================================================================
// handle
// =====================================================================================
struct HaloLayout {
    size_t ghost_bytes;     // one ghost column: GHOST_SLOTS*(ny+2) elements, rounded to 256 B
    size_t mask_bytes;      // the mask of one neighbour's GHOST_COLS-1 outermost columns: [c][ny] bytes, rounded to 256 B
    size_t off_ghost_w[2], off_ghost_e[2];
    size_t off_mask_w, off_mask_e;
    size_t off_flag_w, off_flag_e, off_done_w, off_done_e, off_error;
    size_t total;
};

static HaloLayout halo_layout(int ny, int elem)
{
    HaloLayout h;
    h.ghost_bytes = (((size_t)GHOST_SLOTS * (ny + 2) * elem) + 255) / 256 * 256;
    h.mask_bytes = ((size_t)(GHOST_COLS - 1) * ny + 255) / 256 * 256;
    size_t o = 0;
    for (int p = 0; p < 2; ++p) { h.off_ghost_w[p] = o; o += h.ghost_bytes; }
    for (int p = 0; p < 2; ++p) { h.off_ghost_e[p] = o; o += h.ghost_bytes; }
    h.off_mask_w = o; o += h.mask_bytes;
    h.off_mask_e = o; o += h.mask_bytes;
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
    bool seed_pending = false;    // lb_upload_f happened and the other ping-pong buffer has not been seeded yet (lb_stage_move)
    bool prestream_done = false;  // cython / opencl_old schemes: is the next step's BC + swap already applied to `cur`
    float *frozen = nullptr;      // opencl_old: the populations `move` never writes (lb_oldcl.cuh)
    int tb2_shape = -1;           // two updates per launch: -1 = automatic, 0 = off, else index into g_tb_shapes
    int sm_count = 148;           // SMs of the device (segment heights that fill whole waves of CTAs)
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
    uint32_t state_index = 0;     // number of fused steps taken
    uint32_t halo_epoch = 0;      // number of halo-exchanging launches taken (ghost parity / flag value)
    bool peer_has_mask[2] = {false, false};   // did the last lb_halo_prime find a mask on the neighbour slab
    unsigned long long halo_timeout_ns = 30000000000ull;   // bound of the in-kernel wait for a neighbour
    int edge_rows = 16;           // rows per warp of the one-update kernel's edge tiles
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
    cudaStream_t s_up = nullptr, s_down = nullptr;     // lb_run_streamed: copy streams
    std::vector<cudaEvent_t> ev_pool;                   // lb_run_streamed: per-band events
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j * plane + i] = f[j * plane + i] * c.keep + c.omega * feq[j * plane + i];
}

template <typename T>
__global__ void k_zero_velocity(int nx, int ny, int pitch, const uint8_t *mask, int mask_pitch, T *u, T *v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    if (mask[(long long)y * mask_pitch + x] == 1) {
        u[(long long)y * pitch + x] = (T)0;
        v[(long long)y * pitch + x] = (T)0;
    }
}

template <typename T>
__global__ void k_subsample(int nx, int ny, int pitch, int sx, int sy, int ox, const T *src, T *dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = lb_grid_row();
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

// host mask rows (uint8 or int32, any values) -> device mask bytes: 1 where the host value is exactly 1 (D2Q9.cl:410)
template <typename M>
__global__ void k_mask_pack(int nx, int rows, const M *staged, size_t staged_pitch, uint8_t *mask, int mask_pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= rows) return;
    mask[(long long)y * mask_pitch + x] = staged[(size_t)y * staged_pitch + x] == (M)1 ? 1 : 0;
}

// one flag byte per 32 cells of a row: 0 = no solid node in the group, 1 = some, 2 = all 32 solid
__global__ void k_span_solid(int nx, int ny, const uint8_t *mask, int mask_pitch, uint8_t *span_solid, int nspans)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
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

// copies the current state's GHOST_COLS outermost columns (all nine populations; layout: lb_fused.cuh StepParams)
// and the obstacle mask of the outermost GHOST_COLS-1 columns into the neighbours' arenas, then publishes the flag
template <typename T>
__global__ void k_halo_prime(int nx, int ny, int pitch, long long plane, const T *f, const uint8_t *mask, int mask_pitch,
                             T *out_w, T *out_e, uint8_t *mask_w, uint8_t *mask_e,
                             unsigned int *flag_w_remote, unsigned int *flag_e_remote, unsigned int *error_word, unsigned int value)
{
    const int gs = ny + 2;
    for (int y = threadIdx.x; y < ny; y += blockDim.x) {
        const long long row = (long long)y * pitch;
        for (int c = 0; c < GHOST_COLS; ++c) {
            if (c >= nx) break;                       // a slab narrower than the ghost: those columns never get read
            for (int j = 0; j < 9; ++j) {
                if (out_w) out_w[(c * 9 + j) * gs + y + 1] = f[j * plane + row + c];
                if (out_e) out_e[(c * 9 + j) * gs + y + 1] = f[j * plane + row + (nx - 1 - c)];
            }
            if (c < GHOST_COLS - 1) {
                if (out_w) mask_w[c * ny + y] = mask ? mask[(long long)y * mask_pitch + c] : 0;
                if (out_e) mask_e[c * ny + y] = mask ? mask[(long long)y * mask_pitch + (nx - 1 - c)] : 0;
            }
        }
    }
    if (threadIdx.x == 0) *error_word = 0u;           // a re-primed handle starts from a clean slate
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
// (columns / bx) x rows, rows beyond the 65535 limit of gridDim.y folded over gridDim.z (kernels: lb_grid_row())
static inline dim3 rows_grid(unsigned nbx, int rows)
{
    const unsigned gy = rows < 65535 ? (unsigned)rows : 65535u;
    return dim3(nbx, gy, ((unsigned)rows + gy - 1) / gy);
}
static inline dim3 grid2d(const lb_sim *s, int bx = 128) { return rows_grid((s->cfg.nx + bx - 1) / bx, s->cfg.ny); }

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

static void fill_params(lb_sim *s, StepParams &p, int src_idx, int write_moments)
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
    p.c2 = pack_consts(p.cf);
    p.y_begin = 0; p.y_end = s->cfg.ny;
    p.edge_rows = s->edge_rows;
    p.sm_count = s->sm_count;
    if (uses_halo(s)) {
        // launch number `halo_epoch` reads the ghost columns of parity epoch&1 once the flag says epoch+1,
        // writes the neighbours' columns of the other parity and publishes epoch+2 there
        const uint32_t epoch = s->halo_epoch;
        const int rp = epoch & 1, wp = (epoch + 1) & 1;
        const HaloLayout &h = s->hl;
        p.ghost_w = s->halo + h.off_ghost_w[rp];
        p.ghost_e = s->halo + h.off_ghost_e[rp];
        p.gmask_w = s->peer_has_mask[LB_WEST] ? (const uint8_t *)(s->halo + h.off_mask_w) : nullptr;
        p.gmask_e = s->peer_has_mask[LB_EAST] ? (const uint8_t *)(s->halo + h.off_mask_e) : nullptr;
        p.flag_w_local = (unsigned int *)(s->halo + h.off_flag_w);
        p.flag_e_local = (unsigned int *)(s->halo + h.off_flag_e);
        p.done_w = (unsigned int *)(s->halo + h.off_done_w);
        p.done_e = (unsigned int *)(s->halo + h.off_done_e);
        p.error_word = (unsigned int *)(s->halo + h.off_error);
        p.halo_timeout_ns = s->halo_timeout_ns;
        if (s->peer[LB_WEST]) {   // my westward populations land in the west neighbour's EAST ghost
            p.out_w = s->peer[LB_WEST] + h.off_ghost_e[wp];
            p.flag_w_remote = (unsigned int *)(s->peer[LB_WEST] + h.off_flag_e);
        }
        if (s->peer[LB_EAST]) {
            p.out_e = s->peer[LB_EAST] + h.off_ghost_w[wp];
            p.flag_e_remote = (unsigned int *)(s->peer[LB_EAST] + h.off_flag_w);
        }
        p.step_id = epoch + 1;
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

// one lattice update: reads buffer src_idx, writes the other one
static int launch_step(lb_sim *sim, int src_idx, int write_moments, int y_begin = 0, int y_end = -1)
{
    StepParams p;
    fill_params(sim, p, src_idx, write_moments);
    if (y_end >= 0) { p.y_begin = y_begin; p.y_end = y_end; }
    const LbVariant &var = g_variants[sim->variant];
    if (var.launch_tma) {
        int rc = ensure_tmaps(sim);
        if (rc) return rc;
        const int hi = var.tma_ty == 2 ? 0 : var.tma_ty == 4 ? 1 : 2;
        var.launch_tma(sim->tmap[src_idx][hi][0], sim->tmap[src_idx][hi][1], p, sim->stream);
    } else var.launch(p, sim->stream);
    CU(cudaGetLastError());
    sim->launches++;
    if (uses_halo(sim)) sim->halo_epoch++;
    return LB_OK;
}

// ---- two lattice updates per launch (lb_march.cuh; with -DLB_EXPERIMENTS also lb_tb2.cuh, lb_tb2v.cuh) -------
static inline int tb_kind(int shape) { return (shape > 0 && shape < g_ntb) ? g_tb_shapes[shape].kind : LB_TB_OFF; }
// lattice updates per launch of a marching shape
static inline int tb_depth(int shape) { return (tb_kind(shape) == LB_TB_MARCH && g_tb_shapes[shape].depth > 2) ? g_tb_shapes[shape].depth : 2; }

static size_t tb2_smem_bytes(const lb_sim *sim, int shape)
{
    const LbTbShape &t = g_tb_shapes[shape];
    if (t.kind == LB_TB_ROWS) {
        const int off = 16 / sim->elem, span = sim->elem == 4 ? 128 : 64;
        return (size_t)9 * (span + 2 * off) * (t.by + 2) * sim->elem;
    }
    if (t.kind == LB_TB_CELLS) return (size_t)9 * (t.bx + 2) * (t.by + 2) * sim->elem;
    return 0;
}

// why `shape` cannot serve this handle (nullptr: it can)
static const char *tb_refusal(const lb_sim *sim, int shape)
{
    const int kind = tb_kind(shape);
    if (kind == LB_TB_OFF) return nullptr;
    if (sim->cfg.scheme != LB_SCHEME_OPENCL || sim->cfg.model != LB_MODEL_D2Q9)
        return "two-update kernels serve LB_SCHEME_OPENCL / LB_MODEL_D2Q9 lattices";
    if (g_variants[sim->variant].launch_tma) return "two-update kernels do not combine with a TMA-staged one-update variant";
    const int span = sim->elem == 4 ? 128 : 64;
    if (kind == LB_TB_MARCH) {
        if (!g_tb_shapes[shape].launch_march[sim->cfg.dtype == LB_F64][sim->cfg.math == LB_MATH_FAST])
            return "this shape is not compiled for the handle's dtype";
        if (uses_halo(sim) && sim->cfg.nx < tb_depth(shape))
            return "a halo-connected slab must be at least as many columns wide as the launch is updates deep";
        if (sim->cfg.ny < tb_depth(shape)) return "the lattice must be at least as many rows high as the launch is updates deep";
        const bool rim = !strncmp(g_tb_shapes[shape].name, "rim", 3);       // -DLB_EXPERIMENTS: lb_march_rim.cuh
        const int need = rim ? span : (sim->elem == 4 ? 4 : 2);
        if (sim->cfg.west_edge == LB_EDGE_WRAP && sim->cfg.nx % need)
            return rim ? "the rim-gather marching kernel needs nx to be a multiple of the strip width (128 fp32 / 64 fp64 cells) on a single-slab periodic box"
                       : "the marching kernel needs nx to be a multiple of the vector width (4 fp32 / 2 fp64 cells) on a single-slab periodic box";
        if (sim->cfg.west_edge == LB_EDGE_WRAP && sim->elem == 8 && tb_depth(shape) > 2 && sim->cfg.nx < 4)
            return "three fp64 updates per launch need a periodic box at least 4 cells wide";
        return nullptr;
    }
    // the round-1 shared-memory tiles: single slab only
    if (uses_halo(sim) || sim->cfg.global_nx != sim->cfg.nx) return "the shared-memory tiles serve single-slab lattices";
    if (sim->cfg.ny > 65535 * g_tb_shapes[shape].by) return "lattice too tall for this tile";
    if (tb2_smem_bytes(sim, shape) > 227 * 1024) return "tile does not fit shared memory for this dtype";
    if (kind == LB_TB_ROWS && sim->cfg.bc == LB_BC_PERIODIC && sim->cfg.nx % span)
        return "the row-per-warp tiles need nx to be a multiple of the tile width on a periodic box";
    return nullptr;
}

static int tb2_find(const char *name)
{
    for (int k = 1; k < g_ntb; ++k)
        if (!strcmp(g_tb_shapes[k].name, name)) return k;
    return 0;
}

// The shape used when the caller did not choose (tb2_shape == -1): the marching kernel on lattices large
// enough to be HBM-bound; small lattices keep the graph-batched one-update kernel.  The decision uses only
// what every slab of a decomposed lattice knows (global width, height), so all slabs decide alike.
#ifndef LB_MARCH3_MAX_SEG
#define LB_MARCH3_MAX_SEG 128         // with short segments at the end of a launch tall ones pay: C4 +1 % over 64 rows
#endif
static int tb2_auto_shape(const lb_sim *sim, bool size_gate = true, bool allow3 = true)
{
    if (size_gate && ((long long)sim->cfg.global_nx * sim->cfg.ny < (1ll << 22) || sim->cfg.ny < 64)) return 0;
    // (slabs of one lattice may run different SHAPES -- only the launch sequence has to agree -- so the choice
    // may depend on what this slab looks like)
    // Segment height: a warp walks `seg` rows of one strip.  Short segments cost 2 (K-1) extra level-1 rows each, long
    // ones leave too few (strip, segment) work items to fill and balance 148 SMs x 16-24 warps (profiles/
    // r2_march_segment_height_*.txt, r2_march3_*.txt).  Lattices with enough rows for segments of 16 or more run
    // THREE updates per launch (fp64 with two overlap lanes per side: 56 stored columns per strip); smaller ones two,
    // with segments down to 8 rows (8 on a 4096 x 1024 lattice, 128 on C4 and C5, 64 on C4's N=8 slabs).
    const bool f32 = sim->elem == 4;
    auto items_per_row = [&](int out) { return (long long)((sim->cfg.nx + out - 1) / out); };
    auto pow2_floor = [](long long want, int lo, int hi) { int s = lo; while (s < hi && 2 * s <= want) s *= 2; return s; };
    std::string name;
    const long long want3 = items_per_row(f32 ? 120 : 56) * sim->cfg.ny / 12288;     // three updates per launch like long segments
    if (allow3 && want3 >= 16) name = std::string(f32 ? g_tb_auto_f32_3 : g_tb_auto_f64_3) + ".s" + std::to_string(pow2_floor(want3, 16, LB_MARCH3_MAX_SEG));
    else name = std::string((f32 ? g_tb_auto_f32 : g_tb_auto_f64)[sim->mask != nullptr]) + ".s" +
                std::to_string(pow2_floor(items_per_row(f32 ? 120 : 60) * sim->cfg.ny / 49152, 8, 64));
    int k = tb2_find(name.c_str());
    if (k > 0 && tb_refusal(sim, k) && tb_depth(k) > 2) return tb2_auto_shape(sim, size_gate, false);   // e.g. a 2-column slab
    return (k > 0 && !tb_refusal(sim, k)) ? k : 0;
}

static int tb2_effective_shape(const lb_sim *sim)
{
    if (sim->tb2_shape >= 0) return tb_refusal(sim, sim->tb2_shape) ? 0 : sim->tb2_shape;
    return tb2_auto_shape(sim);
}

// Rows per segment of a marching launch.  A hand-picked shape runs with the height in its name.  The automatic choice
// starts from that height too; on lattices of a few waves of CTAs it takes the height (6 .. 64 rows) with the best
// "fill of the last wave x useful rows per segment" instead: on C2 (4096 x 1024) 11-row segments make 823 CTAs for 888
// slots, one full wave, where 8-row ones make 1.26 waves -- +10 % (profiles/r2_seg_sweep.txt; lattices of many waves
// show no such effect there and keep the power of two).
static int tb2_segment_rows(const lb_sim *sim, int shape)
{
    const LbTbShape &t = g_tb_shapes[shape];
    if (t.kind != LB_TB_MARCH || sim->tb2_shape >= 0 || t.minb <= 0 || t.nw <= 0) return t.seg_rows;
    const int K = tb_depth(shape), ny = sim->cfg.ny;
    const int out = sim->elem == 4 ? 120 : (K > 2 ? 56 : 60);
    const long long nstrips = (sim->cfg.nx + out - 1) / out;
    const long long slots = (long long)sim->sm_count * t.minb;               // CTAs resident at once
    auto ctas = [&](int S) { return (nstrips * ((ny + S - 1) / S) + t.nw - 1) / t.nw; };
    if (ctas(t.seg_rows) > 3 * slots) return t.seg_rows;
    auto fill = [&](int S) {
        const long long c = ctas(S), waves = (c + slots - 1) / slots;
        return (double)c / (double)(waves * slots) * (double)S / (double)(S + 2 * (K - 1));
    };
    int best = t.seg_rows;
    double best_fill = fill(best) * 1.02;                                    // a clear gain, or the power of two stays
    for (int S = 6; S <= 64; ++S)
        if (fill(S) > best_fill) { best = S; best_fill = fill(S); }
    return best;
}

// ... and the height of the segments at the END of such a launch (0: all alike).  A launch of many waves ends with a
// tail in which ever fewer SMs still work on their last 64-row items (half an item's duration on average: 6 % of
// a launch on a 4096 x 32768 slab, 1 % on C4); cutting the last rows -- two waves of work items' worth, see
// march_segments in lb_k_march.cu -- into segments a quarter as high shortens that tail four times.
static int tb2_short_segment_rows(const lb_sim *sim, int shape)
{
    const LbTbShape &t = g_tb_shapes[shape];
    if (t.kind != LB_TB_MARCH || sim->tb2_shape >= 0 || t.minb <= 0 || t.nw <= 0) return 0;
    const int S1 = tb2_segment_rows(sim, shape);
    if (S1 != t.seg_rows || S1 < 32) return 0;        // few waves: the height already fills them
    return S1 / 4;
}

// one launch of a two-update (three-update: tb_depth) shape: reads buffer src_idx, writes the other one.  The marching
// kernel can store the moments of its last step; the round-1 tiles cannot (write_moments must be 0 for them).
static int launch_two_steps(lb_sim *sim, int src_idx, int shape, int write_moments, int y_begin = 0, int y_end = -1)
{
    const LbTbShape &t = g_tb_shapes[shape];
    const int di = sim->cfg.dtype == LB_F64, mi = sim->cfg.math == LB_MATH_FAST;
    if (t.kind == LB_TB_MARCH) {
        StepParams p;
        fill_params(sim, p, src_idx, write_moments);
        if (y_end >= 0) { p.y_begin = y_begin; p.y_end = y_end; }
        p.seg_rows = tb2_segment_rows(sim, shape);
        p.seg_rows2 = tb2_short_segment_rows(sim, shape);
#ifdef LB_SEG_ROWS_ENV                                 // side builds of tools/seg_sweep.sh: any segment height
        if (const char *e = getenv("LB_SEG_ROWS")) p.seg_rows = atoi(e) > 0 ? atoi(e) : p.seg_rows;
        if (const char *e = getenv("LB_SEG_ROWS2")) p.seg_rows2 = atoi(e);
#endif
        t.launch_march[di][mi](p, sim->stream);
    } else if (t.kind == LB_TB_ROWS) {
        StepParams p;
        fill_params(sim, p, src_idx, 0);
        const int span = sim->elem == 4 ? 128 : 64;
        const dim3 grid(sim->pitch / span, (sim->cfg.ny + t.by - 1) / t.by);
        t.launch_rows[di][mi](p, grid, tb2_smem_bytes(sim, shape), sim->stream);
    } else {
        Tb2Params p{};
        p.src = sim->buf[src_idx]; p.dst = sim->buf[src_idx ^ 1];
        p.plane = sim->plane; p.nx = sim->cfg.nx; p.ny = sim->cfg.ny; p.pitch = sim->pitch;
        p.bc = sim->cfg.bc == LB_BC_PERIODIC ? BC_PERIODIC : BC_PIPE;
        p.zero_obstacle_velocity = sim->cfg.zero_obstacle_velocity;
        p.mask = sim->mask; p.mask_pitch = sim->mask_pitch;
        p.cf = consts_of<float>(sim); p.cd = consts_of<double>(sim);
        const dim3 grid((sim->cfg.nx + t.bx - 1) / t.bx, (sim->cfg.ny + t.by - 1) / t.by);
        t.launch_cells[di][mi](p, grid, tb2_smem_bytes(sim, shape), sim->stream);
    }
    CU(cudaGetLastError());
    sim->launches++;
    if (uses_halo(sim)) sim->halo_epoch++;
    return LB_OK;
}

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int lb_temporal_blocking(const lb_sim *sim) { return sim ? tb2_effective_shape(sim) : 0; }
int lb_tb2_shape_count(void) { return g_ntb; }

// diagnostic (no device needed): the launch geometry the marching kernels' launchers compute
int lb_plan_march_launch(int nx, int rows, int elem_bytes, int depth, int nw, int minb, int seg_rows, int seg_rows2, int sm_count,
                         int west_halo, int east_halo, int *n_strips, int *n_edge_strips, int *n_tall, int *n_short, int *short_rows)
{
    if (nx < 1 || rows < 1 || (elem_bytes != 4 && elem_bytes != 8) || depth < 2 || depth > 3 || nw < 1 || minb < 1 || seg_rows < 1)
        return LB_ERR_INVALID;
    const int out = elem_bytes == 4 ? 120 : (depth > 2 ? 56 : 60);
    const int nstrips = (nx + out - 1) / out;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.y_begin = 0; p.y_end = rows; p.seg_rows = seg_rows; p.seg_rows2 = seg_rows2; p.sm_count = sm_count;
    const int nseg = lb_march_segments(p, nstrips, nw, minb);
    if (n_strips) *n_strips = nstrips;
    if (n_edge_strips) *n_edge_strips = lb_march_edge_strips(nx, out, nstrips, west_halo != 0, east_halo != 0);
    if (n_tall) *n_tall = p.seg_tall;
    if (n_short) *n_short = nseg - p.seg_tall;
    if (short_rows) *short_rows = p.seg_rows2;
    return LB_OK;
}
int lb_segment_rows(const lb_sim *sim) { const int k = sim ? tb2_effective_shape(sim) : 0; return k > 0 ? tb2_segment_rows(sim, k) : 0; }
const char *lb_tb2_shape_name(int shape) { return (shape >= 0 && shape < g_ntb) ? g_tb_shapes[shape].name : nullptr; }

int lb_set_temporal_blocking(lb_sim *sim, int shape)
{
    if (!sim) return LB_ERR_INVALID;
    if (shape == -1) { sim->tb2_shape = -1; return LB_OK; }
    if (shape < 0 || shape >= g_ntb) return fail(sim, LB_ERR_INVALID, "lb_set_temporal_blocking: unknown shape");
    if (const char *why = tb_refusal(sim, shape)) return fail(sim, LB_ERR_INVALID, std::string("lb_set_temporal_blocking: ") + why);
    sim->tb2_shape = shape;
    return LB_OK;
}

int lb_set_halo_timeout(lb_sim *sim, double seconds)
{
    if (!sim || !(seconds > 0.0)) return fail(sim, LB_ERR_INVALID, "lb_set_halo_timeout: seconds must be positive");
    sim->halo_timeout_ns = seconds > 1.8e10 ? ~0ull : (unsigned long long)(seconds * 1e9);
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
    { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && n > 0) sim->sm_count = n; else cudaGetLastError(); }
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
    for (cudaEvent_t e : sim->ev_pool) cudaEventDestroy(e);
    if (sim->s_up) cudaStreamDestroy(sim->s_up);
    if (sim->s_down) cudaStreamDestroy(sim->s_down);
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

}  // extern "C"

// host arrays may be wider than the slab (lb_multi_*: a slab's columns inside the global array): `host_row`
// is the host row length in elements, the pointer already points at the slab's first column
static int set_mask_impl(lb_sim *sim, const void *host_mask, int elem_bytes, size_t host_row)
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
    if (sim->cfg.bc == LB_BC_VELOCITY_YPERIODIC && is_cython(sim))      // lb_cython.cuh folds the row exchange into the pull
        for (int row : {0, ny - 1})
            for (int x = 0; x < nx; ++x) {
                const size_t i = (size_t)row * host_row + x;
                const bool solid = elem_bytes == 1 ? ((const uint8_t *)host_mask)[i] == 1 : ((const int32_t *)host_mask)[i] == 1;
                if (solid) return fail(sim, LB_ERR_INVALID, "lb_set_mask: with LB_BC_VELOCITY_YPERIODIC the exchanged rows y=0 and y=ny-1 must be free of solid nodes");
            }
    if (!sim->mask) {
        sim->mask_pitch = sim->pitch;
        sim->nspans = sim->pitch / 32;
        CU(cudaMalloc((void **)&sim->mask, (size_t)sim->mask_pitch * ny));
        CU(cudaMalloc((void **)&sim->span_solid, (size_t)sim->nspans * ny));
    }
    CU(cudaMemsetAsync(sim->mask, 0, (size_t)sim->mask_pitch * ny, sim->stream));
    // staged through a bounded device buffer (<= 64 MB), normalised to {0,1} on the device
    const size_t row_bytes = (size_t)nx * elem_bytes;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ny, ((size_t)64 << 20) / row_bytes));
    void *stage = nullptr;
    CU(cudaMalloc(&stage, row_bytes * chunk));
    cudaError_t e = cudaSuccess;
    for (int y0 = 0; y0 < ny && e == cudaSuccess; y0 += chunk) {
        const int rows = std::min(chunk, ny - y0);
        e = cudaMemcpy2DAsync(stage, row_bytes, (const char *)host_mask + (size_t)y0 * host_row * elem_bytes, host_row * elem_bytes,
                              row_bytes, rows, cudaMemcpyHostToDevice, sim->stream);
        if (e != cudaSuccess) break;
        uint8_t *dst = sim->mask + (size_t)y0 * sim->mask_pitch;
        if (elem_bytes == 1) k_mask_pack<uint8_t><<<rows_grid((nx + 127) / 128, rows), 128, 0, sim->stream>>>(nx, rows, (const uint8_t *)stage, nx, dst, sim->mask_pitch);
        else k_mask_pack<int32_t><<<rows_grid((nx + 127) / 128, rows), 128, 0, sim->stream>>>(nx, rows, (const int32_t *)stage, nx, dst, sim->mask_pitch);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        k_span_solid<<<rows_grid((sim->nspans + 63) / 64, ny), 64, 0, sim->stream>>>(nx, ny, sim->mask, sim->mask_pitch, sim->span_solid, sim->nspans);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(sim->stream);   // the host array may go away after the call
    cudaFree(stage);
    CU(e);
    return LB_OK;
}

extern "C" {

int lb_set_mask(lb_sim *sim, const void *host_mask, int elem_bytes)
{
    return set_mask_impl(sim, host_mask, elem_bytes, sim ? (size_t)sim->cfg.nx : 0);
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
    k_span_solid<<<rows_grid((sim->nspans + 63) / 64, ny), 64, 0, sim->stream>>>(nx, ny, sim->mask, sim->mask_pitch,
                                                                            sim->span_solid, sim->nspans);
    CU(cudaGetLastError());
    return LB_OK;
}

}  // extern "C"

static int download_impl(lb_sim *sim, int field, void *host_out, size_t host_row, bool sync);

static int upload_f_impl(lb_sim *sim, const void *host_f, size_t host_row)
{
    if (!sim || !host_f) return fail(sim, LB_ERR_INVALID, "lb_upload_f: null argument");
    CU(cudaSetDevice(sim->cfg.device));
    const size_t w = (size_t)sim->cfg.nx * sim->elem, dp = (size_t)sim->pitch * sim->elem;
    // planes are contiguous (plane = ny*pitch), so one 2-D copy of 9*ny rows does all nine
    CU(cudaMemcpy2DAsync(sim->buf[sim->cur], dp, host_f, host_row * sim->elem, w, (size_t)9 * sim->cfg.ny, cudaMemcpyHostToDevice, sim->stream));
    // The reference seeds f_streamed with the same data (opencl_dim.py:324-327).  Only `move` as a single stage
    // can see that seed (destinations without an upstream node keep it); the fused step overwrites or closes
    // every such slot, so the copy is deferred until lb_stage_move asks for it.
    sim->seed_pending = true;
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

static int upload_moments_impl(lb_sim *sim, const void *host_rho, const void *host_u, const void *host_v, size_t host_row)
{
    if (!sim) return LB_ERR_INVALID;
    CU(cudaSetDevice(sim->cfg.device));
    const void *hs[3] = {host_rho, host_u, host_v};
    void *ds[3] = {sim->rho, sim->u, sim->v};
    for (int i = 0; i < 3; ++i) {
        const int eb = i == 0 ? sim->elem : sim->uv_elem;      // u, v are float64 for the cython schemes
        const size_t w = (size_t)sim->cfg.nx * eb, dp = (size_t)sim->pitch * eb;
        if (hs[i]) CU(cudaMemcpy2DAsync(ds[i], dp, hs[i], host_row * eb, w, sim->cfg.ny, cudaMemcpyHostToDevice, sim->stream));
    }
    CU(cudaStreamSynchronize(sim->stream));
    return LB_OK;
}

extern "C" {

int lb_upload_f(lb_sim *sim, const void *host_f) { return upload_f_impl(sim, host_f, sim ? (size_t)sim->cfg.nx : 0); }

int lb_upload_moments(lb_sim *sim, const void *host_rho, const void *host_u, const void *host_v)
{
    return upload_moments_impl(sim, host_rho, host_u, host_v, sim ? (size_t)sim->cfg.nx : 0);
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
    for (int i = 0; i < GRAPH_LEN && rc == LB_OK; ++i) { rc = launch_step(sim, idx, 0); idx ^= 1; }
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

}  // extern "C"

// the two-update shape that accompanies a three-update one (a run of 3a + 2 steps starts with one pair)
static int pair_shape(const lb_sim *sim, int shape)
{
    if (tb_depth(shape) == 2) return shape;
    const int k = tb2_auto_shape(sim, false, false);
    return (k > 0 && tb_depth(k) == 2) ? k : 0;
}

// How n steps are cut into launches with marching shape `shape` (d = 2 or 3 updates per launch): the remainder
// first -- one single-update launch, or one pair -- then n / d full launches.  Depths never decrease (lb_run_streamed
// relies on it), and every slab of a lattice computes the same plan.
static std::vector<int> plan_launches(const lb_sim *sim, int shape, int n_steps)
{
    std::vector<int> plan;
    const int d = tb_depth(shape);
    int lead = n_steps % d;
    if (lead == 2 && !pair_shape(sim, shape)) { plan.push_back(1); lead = 1; }
    if (lead) plan.push_back(lead);
    for (int k = 0; k < n_steps / d; ++k) plan.push_back(d);
    return plan;
}

// n_steps lattice updates.  `final`: the run ends here, so its last launch stores rho, u, v; lb_multi_step
// enqueues long runs in chunks and passes false for all but the last one.
static int step_impl(lb_sim *sim, int n_steps, bool final)
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
    sim->seed_pending = false;              // every launch below rewrites the other ping-pong buffer completely
    if (is_oldcl(sim)) return oldcl_steps(sim, n_steps);
    if (sim->cfg.scheme != LB_SCHEME_OPENCL) return cython_steps(sim, n_steps);
    const int tb = tb2_effective_shape(sim);
    if (tb_kind(tb) == LB_TB_MARCH) {
        // every step of the run inside multi-update launches (plan_launches); the launch that ends the run also
        // stores rho, u, v (the moments of the run's last step)
        const std::vector<int> plan = plan_launches(sim, tb, n_steps);
        int remaining = n_steps;
        for (size_t j = 0; j < plan.size(); ++j) {
            const int d = plan[j];
            const bool last = final && j + 1 == plan.size();
            int rc;
            if (d == 1) rc = launch_step(sim, sim->cur, last);
            else rc = launch_two_steps(sim, sim->cur, d == tb_depth(tb) ? tb : pair_shape(sim, tb), last);
            if (rc) return rc;
            sim->cur ^= 1; sim->state_index += d; remaining -= d;
        }
        return LB_OK;
    }
    int remaining = n_steps - 1;            // all but the last step skip the moment stores
    if (tb) {                               // round-1 tiles (-DLB_EXPERIMENTS): moment-free steps two at a time
        for (; remaining >= 2; remaining -= 2) {
            int rc = launch_two_steps(sim, sim->cur, tb, 0);
            if (rc) return rc;
            sim->cur ^= 1; sim->state_index += 2;
        }
    }
    const bool graphs_ok = !uses_halo(sim); // halo launches carry a per-launch flag value
    while (graphs_ok && remaining >= GRAPH_LEN) {
        int rc = ensure_graph(sim);
        if (rc) return rc;
        CU(cudaGraphLaunch(sim->graph[sim->cur], sim->stream));
        sim->launches += GRAPH_LEN;
        sim->state_index += GRAPH_LEN;
        remaining -= GRAPH_LEN;
    }
    for (; remaining > 0; --remaining) {
        int rc = launch_step(sim, sim->cur, 0);
        if (rc) return rc;
        sim->cur ^= 1; sim->state_index++;
    }
    int rc = launch_step(sim, sim->cur, final ? 1 : 0);
    if (rc) return rc;
    sim->cur ^= 1; sim->state_index++;
    return LB_OK;
}

extern "C" {

int lb_step(lb_sim *sim, int n_steps) { return step_impl(sim, n_steps, true); }

// ---- end to end with host buffers: upload, n_steps, read-back -- pipelined by row bands -----------------------
// The three stages of the reference's user code -- init_pop's upload (opencl_dim.py:324-327), run (:372-387),
// get_fields' read-back (:390-407) -- touch every row once each, and row y of step s only needs rows y-1..y+1 of
// step s-1.  So the lattice is uploaded in bands of rows; as soon as band b has arrived, every launch of the run
// is issued for the rows whose dependency cone lies inside what has arrived (launch j, which brings the rows to
// time level D_j, covers the rows up to D_j short of the band's end: a skewed wavefront), and the finished rows
// of rho, u, v start their way back while later bands are still being uploaded.  H2D, compute and D2H overlap;
// the call takes about as long as the upload alone.  Stream order alone guarantees correctness: a launch only
// reads rows that earlier launches of the same stream (or the awaited upload) completed, and with non-decreasing
// launch depths a launch never overwrites rows that a later-issued launch still reads.  Bit-identical to
// lb_upload_f + lb_step + lb_download.
int lb_run_streamed(lb_sim *sim, const void *host_f, int n_steps, void *host_rho, void *host_u, void *host_v)
{
    if (!sim || !host_f) return fail(sim, LB_ERR_INVALID, "lb_run_streamed: null argument");
    if (n_steps < 1) return fail(sim, LB_ERR_INVALID, "lb_run_streamed: n_steps must be >= 1");
    CU(cudaSetDevice(sim->cfg.device));
    const int nx = sim->cfg.nx, ny = sim->cfg.ny;
    const int tb = tb2_effective_shape(sim);
    const bool can_stream = sim->cfg.scheme == LB_SCHEME_OPENCL && !uses_halo(sim) && sim->cfg.bc == LB_BC_PIPE &&
                            tb_kind(tb) == LB_TB_MARCH && strncmp(g_tb_shapes[tb].name, "rim", 3) != 0 &&
                            !g_variants[sim->variant].launch_tma && n_steps <= 256 && ny >= 1024;
    void *outs[3] = {host_rho, host_u, host_v};
    const int fields[3] = {LB_FIELD_RHO, LB_FIELD_U, LB_FIELD_V};
    if (!can_stream) {                     // same result, stage after stage
        int rc = upload_f_impl(sim, host_f, (size_t)nx);
        if (rc == LB_OK) rc = step_impl(sim, n_steps, true);
        for (int i = 0; i < 3 && rc == LB_OK; ++i)
            if (outs[i]) rc = download_impl(sim, fields[i], outs[i], (size_t)nx, false);
        return rc == LB_OK ? lb_sync(sim) : rc;
    }
    if (!sim->s_up) {
        CU(cudaStreamCreateWithFlags(&sim->s_up, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&sim->s_down, cudaStreamNonBlocking));
    }
    // launch depths d_j (non-decreasing: the remainder first) and D_j = d_0 + ... + d_j
    const std::vector<int> depth = plan_launches(sim, tb, n_steps);
    std::vector<int> reach;
    for (int d : depth) reach.push_back((reach.empty() ? 0 : reach.back()) + d);
    const int L = (int)depth.size();
    int band = ((ny + 23) / 24 + 63) / 64 * 64;        // about 24 bands, whole segments
    const int nb = (ny + band - 1) / band;
    while ((int)sim->ev_pool.size() < 2 * nb + 1) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sim->ev_pool.push_back(e);
    }
    const int A = sim->cur;                // level 0 is uploaded into the current buffer
    const size_t w = (size_t)nx * sim->elem, dp = (size_t)sim->pitch * sim->elem;
    // uploads start when everything already queued on the compute stream is done with the buffers
    CU(cudaEventRecord(sim->ev_pool[2 * nb], sim->stream));
    CU(cudaStreamWaitEvent(sim->s_up, sim->ev_pool[2 * nb], 0));
    std::vector<int> done_to(L, 0);        // rows [0, done_to[j]) of launch j have been issued
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * band, r1 = std::min(ny, r0 + band);
        for (int j = 0; j < 9; ++j)
            CU(cudaMemcpy2DAsync((char *)sim->buf[A] + ((size_t)j * sim->plane + (size_t)r0 * sim->pitch) * sim->elem, dp,
                                 (const char *)host_f + ((size_t)j * ny + r0) * w, w, w, (size_t)(r1 - r0), cudaMemcpyHostToDevice, sim->s_up));
        CU(cudaEventRecord(sim->ev_pool[b], sim->s_up));
        CU(cudaStreamWaitEvent(sim->stream, sim->ev_pool[b], 0));
        int lo_last = 0, hi_last = 0;
        for (int j = 0; j < L; ++j) {
            const int lo = done_to[j];
            const int hi = (b == nb - 1) ? ny : std::max(0, r1 - reach[j]);
            if (j == L - 1) { lo_last = lo; hi_last = std::max(lo, hi); }
            if (hi <= lo) continue;
            const int src_idx = A ^ (j & 1);
            const int rc = depth[j] == 1 ? launch_step(sim, src_idx, j == L - 1, lo, hi)
                                         : launch_two_steps(sim, src_idx, depth[j] == tb_depth(tb) ? tb : pair_shape(sim, tb), j == L - 1, lo, hi);
            if (rc) return rc;
            done_to[j] = hi;
        }
        if (hi_last > lo_last) {           // the rows of this band's last launch are final: send their moments home
            CU(cudaEventRecord(sim->ev_pool[nb + b], sim->stream));
            CU(cudaStreamWaitEvent(sim->s_down, sim->ev_pool[nb + b], 0));
            void *dev[3] = {sim->rho, sim->u, sim->v};
            for (int i = 0; i < 3; ++i) {
                if (!outs[i]) continue;
                const int eb = i == 0 ? sim->elem : sim->uv_elem;
                CU(cudaMemcpy2DAsync((char *)outs[i] + (size_t)lo_last * nx * eb, (size_t)nx * eb,
                                     (const char *)dev[i] + (size_t)lo_last * sim->pitch * eb, (size_t)sim->pitch * eb,
                                     (size_t)nx * eb, (size_t)(hi_last - lo_last), cudaMemcpyDeviceToHost, sim->s_down));
            }
        }
    }
    sim->cur = A ^ (L & 1);
    sim->state_index += (uint32_t)n_steps;
    sim->seed_pending = false;
    sim->prestream_done = false;
    CU(cudaStreamSynchronize(sim->s_up));
    CU(cudaStreamSynchronize(sim->s_down));
    return lb_sync(sim);
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

}  // extern "C"

static int download_impl(lb_sim *sim, int field, void *host_out, size_t host_row, bool sync)
{
    if (!sim || !host_out) return fail(sim, LB_ERR_INVALID, "lb_download: null argument");
    CU(cudaSetDevice(sim->cfg.device));
    int eb = sim->elem;
    if (field == LB_FIELD_U || field == LB_FIELD_V) eb = sim->uv_elem;
    const size_t w = (size_t)sim->cfg.nx * eb, dp = (size_t)sim->pitch * eb;
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
    CU(cudaMemcpy2DAsync(host_out, host_row * eb, src, dp, w, rows, cudaMemcpyDeviceToHost, sim->stream));
    return sync ? lb_sync(sim) : LB_OK;
}

extern "C" {

int lb_download(lb_sim *sim, int field, void *host_out) { return download_impl(sim, field, host_out, sim ? (size_t)sim->cfg.nx : 0, true); }

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
    const dim3 grid = rows_grid((ox + 127) / 128, oy);
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
    if (sim->seed_pending) {       // opencl_dim.py:324-327: f_streamed starts as a copy of f
        CU(cudaMemcpyAsync(sim->buf_base[sim->cur ^ 1], sim->buf_base[sim->cur], sim->buf_bytes, cudaMemcpyDeviceToDevice, sim->stream));
        sim->seed_pending = false;
    }
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
    const int par = sim->halo_epoch & 1;
    const bool w = sim->cfg.west_edge == LB_EDGE_HALO, e = sim->cfg.east_edge == LB_EDGE_HALO;
    if ((w && !sim->peer[LB_WEST]) || (e && !sim->peer[LB_EAST])) return fail(sim, LB_ERR_STATE, "lb_halo_prime: halo edge not connected");
    // my westward columns land in the west neighbour's EAST ghost / mask column, and vice versa
    char *ow = w ? sim->peer[LB_WEST] + h.off_ghost_e[par] : nullptr;
    char *oe = e ? sim->peer[LB_EAST] + h.off_ghost_w[par] : nullptr;
    uint8_t *mw = w ? (uint8_t *)(sim->peer[LB_WEST] + h.off_mask_e) : nullptr;
    uint8_t *me = e ? (uint8_t *)(sim->peer[LB_EAST] + h.off_mask_w) : nullptr;
    unsigned int *fw = w ? (unsigned int *)(sim->peer[LB_WEST] + h.off_flag_e) : nullptr;
    unsigned int *fe = e ? (unsigned int *)(sim->peer[LB_EAST] + h.off_flag_w) : nullptr;
    unsigned int *err = (unsigned int *)(sim->halo + h.off_error);
    // a handle that timed out may have left its edge-tile counters half way
    CU(cudaMemsetAsync(sim->halo + h.off_done_w, 0, 4, sim->stream));
    CU(cudaMemsetAsync(sim->halo + h.off_done_e, 0, 4, sim->stream));
    if (sim->cfg.dtype == LB_F32)
        k_halo_prime<float><<<1, 1024, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const float *)sim->buf[sim->cur],
                                                          sim->mask, sim->mask_pitch, (float *)ow, (float *)oe, mw, me, fw, fe, err, sim->halo_epoch + 1);
    else
        k_halo_prime<double><<<1, 1024, 0, sim->stream>>>(sim->cfg.nx, sim->cfg.ny, sim->pitch, sim->plane, (const double *)sim->buf[sim->cur],
                                                           sim->mask, sim->mask_pitch, (double *)ow, (double *)oe, mw, me, fw, fe, err, sim->halo_epoch + 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sim->stream));
    // the neighbours' mask columns arrive in MY arena by THEIR prime; always consult them (all zeros = no solids)
    sim->peer_has_mask[LB_WEST] = w;
    sim->peer_has_mask[LB_EAST] = e;
    return LB_OK;
}

}  // extern "C"

// =====================================================================================
// lb_multi: one lattice on several slabs / devices behind one handle (SURVEY.md section 8b:
// the handle owns streams, peer mappings and the step loop; the reference binds one queue to devices[0],
// opencl_dim.py:229-240)
// =====================================================================================
struct lb_multi {
    lb_config cfg;                       // the whole lattice
    std::vector<lb_sim *> slabs;
    std::vector<int> x0, w, device;
    std::vector<cudaStream_t> shared_streams;   // one per device that carries more than one slab
    bool concurrent = false;             // one slab per device: slabs run asynchronously, synchronised by the in-kernel flags
    bool primed = false;
    std::string err;
};

static thread_local std::string g_multi_create_error;

static int mfail(lb_multi *m, int code, const std::string &msg)
{
    if (m) m->err = msg; else g_multi_create_error = msg;
    return code;
}
// a failed slab call: copy its message up
static int mslab(lb_multi *m, size_t k, int rc)
{
    if (rc != LB_OK) m->err = "slab " + std::to_string(k) + ": " + m->slabs[k]->err;
    return rc;
}
#define MS(k, call) do { int rc__ = mslab(m, (k), (call)); if (rc__ != LB_OK) return rc__; } while (0)

static int multi_prime(lb_multi *m)
{
    if (m->slabs.size() > 1) {
        for (size_t k = 0; k < m->slabs.size(); ++k) MS(k, lb_sync(m->slabs[k]));          // every state final ...
        for (size_t k = 0; k < m->slabs.size(); ++k) MS(k, lb_halo_prime(m->slabs[k]));    // ... before any ghost is written
    }
    m->primed = true;
    return LB_OK;
}

extern "C" {

const char *lb_multi_last_error(const lb_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int lb_multi_slab_count(const lb_multi *m) { return m ? (int)m->slabs.size() : 0; }

int lb_multi_slab(lb_multi *m, int k, lb_sim **slab, int *x_offset, int *nx)
{
    if (!m || k < 0 || k >= (int)m->slabs.size()) return mfail(m, LB_ERR_INVALID, "lb_multi_slab: bad index");
    if (slab) *slab = m->slabs[k];
    if (x_offset) *x_offset = m->x0[k];
    if (nx) *nx = m->w[k];
    return LB_OK;
}

int lb_multi_destroy(lb_multi *m)
{
    if (!m) return LB_OK;
    // a neighbour's last launch may still be storing into a slab's arena: drain everything first
    for (lb_sim *s : m->slabs) if (s) { cudaSetDevice(s->cfg.device); cudaStreamSynchronize(s->stream); }
    for (size_t k = m->slabs.size(); k-- > 0;) lb_destroy(m->slabs[k]);
    for (size_t d = 0; d < m->shared_streams.size(); ++d) if (m->shared_streams[d]) cudaStreamDestroy(m->shared_streams[d]);
    cudaGetLastError();
    delete m;
    return LB_OK;
}

int lb_multi_create(const lb_config *cfg, int n_slabs, const int *device_ids, lb_multi **out)
{
    if (!cfg || !out || n_slabs < 1 || !device_ids) return mfail(nullptr, LB_ERR_INVALID, "lb_multi_create: bad argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(lb_config)) return mfail(nullptr, LB_ERR_INVALID, "lb_multi_create: lb_config size mismatch (ABI)");
    if (cfg->global_nx != cfg->nx || cfg->x_offset != 0) return mfail(nullptr, LB_ERR_INVALID, "lb_multi_create: cfg describes the WHOLE lattice (global_nx = nx, x_offset = 0)");
    if (n_slabs > 1 && cfg->scheme != LB_SCHEME_OPENCL) return mfail(nullptr, LB_ERR_INVALID, "lb_multi_create: only LB_SCHEME_OPENCL lattices decompose into slabs");
    if (cfg->nx / n_slabs < 2) return mfail(nullptr, LB_ERR_INVALID, "lb_multi_create: each slab needs at least 2 columns");
    lb_multi *m = new lb_multi();
    m->cfg = *cfg;
    // remainder columns go to the first slabs: widths differ by at most one
    const int base = cfg->nx / n_slabs, rem = cfg->nx % n_slabs;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    cudaGetLastError();
    std::vector<int> per_device(ndev > 0 ? ndev : 1, 0);
    for (int k = 0; k < n_slabs; ++k)
        if (device_ids[k] >= 0 && device_ids[k] < ndev) per_device[device_ids[k]]++;
    m->shared_streams.assign(per_device.size(), nullptr);
    bool distinct = n_slabs > 1;
    for (int c : per_device) if (c > 1) distinct = false;
    m->concurrent = distinct;
    int x = 0;
    for (int k = 0; k < n_slabs; ++k) {
        lb_config c = *cfg;
        c.device = device_ids[k];
        c.nx = base + (k < rem ? 1 : 0);
        c.x_offset = x;
        c.global_nx = cfg->nx;
        if (n_slabs > 1) {
            const bool periodic = cfg->bc == LB_BC_PERIODIC;
            c.west_edge = (periodic || k > 0) ? LB_EDGE_HALO : LB_EDGE_BOUNDARY;
            c.east_edge = (periodic || k < n_slabs - 1) ? LB_EDGE_HALO : LB_EDGE_BOUNDARY;
        }
        // slabs that share a device share a stream and advance in lock-step, one launch at a time
        c.stream = nullptr;
        if (c.device >= 0 && c.device < ndev && per_device[c.device] > 1) {
            if (!m->shared_streams[c.device]) {
                if (cudaSetDevice(c.device) != cudaSuccess || cudaStreamCreateWithFlags(&m->shared_streams[c.device], cudaStreamNonBlocking) != cudaSuccess) {
                    cudaGetLastError();
                    lb_multi_destroy(m);
                    return mfail(nullptr, LB_ERR_CUDA, "lb_multi_create: cannot create a stream");
                }
            }
            c.stream = m->shared_streams[c.device];
        }
        lb_sim *s = nullptr;
        const int rc = lb_create(&c, &s);
        if (rc != LB_OK) {
            const std::string why = "lb_multi_create: slab " + std::to_string(k) + ": " + g_create_error;
            lb_multi_destroy(m);
            return mfail(nullptr, rc, why);
        }
        m->slabs.push_back(s); m->x0.push_back(x); m->w.push_back(c.nx); m->device.push_back(c.device);
        x += c.nx;
    }
    for (int k = 0; k < n_slabs && n_slabs > 1; ++k) {
        lb_sim *s = m->slabs[k];
        int rc = LB_OK;
        if (s->cfg.west_edge == LB_EDGE_HALO) rc = lb_halo_connect_local(s, LB_WEST, m->slabs[(k + n_slabs - 1) % n_slabs]);
        if (rc == LB_OK && s->cfg.east_edge == LB_EDGE_HALO) rc = lb_halo_connect_local(s, LB_EAST, m->slabs[(k + 1) % n_slabs]);
        if (rc != LB_OK) {
            const std::string why = "lb_multi_create: connecting slab " + std::to_string(k) + ": " + s->err;
            lb_multi_destroy(m);
            return mfail(nullptr, rc, why);
        }
    }
    *out = m;
    return LB_OK;
}

/* global host arrays in, global host arrays out: every slab copies its own columns (strided 2-D copies) */
int lb_multi_set_mask(lb_multi *m, const void *host_mask, int elem_bytes)
{
    if (!m) return LB_ERR_INVALID;
    for (size_t k = 0; k < m->slabs.size(); ++k)
        MS(k, set_mask_impl(m->slabs[k], host_mask ? (const char *)host_mask + (size_t)m->x0[k] * elem_bytes : nullptr, elem_bytes, (size_t)m->cfg.nx));
    m->primed = false;                   // the neighbours hold a copy of each slab's boundary mask column
    return LB_OK;
}

int lb_multi_set_mask_disk(lb_multi *m, double cx, double cy, double r)
{
    if (!m) return LB_ERR_INVALID;
    for (size_t k = 0; k < m->slabs.size(); ++k) MS(k, lb_set_mask_disk(m->slabs[k], cx, cy, r));
    m->primed = false;
    return LB_OK;
}

int lb_multi_upload_f(lb_multi *m, const void *host_f)
{
    if (!m || !host_f) return mfail(m, LB_ERR_INVALID, "lb_multi_upload_f: null argument");
    for (size_t k = 0; k < m->slabs.size(); ++k)
        MS(k, upload_f_impl(m->slabs[k], (const char *)host_f + (size_t)m->x0[k] * m->slabs[k]->elem, (size_t)m->cfg.nx));
    return multi_prime(m);
}

int lb_multi_upload_moments(lb_multi *m, const void *host_rho, const void *host_u, const void *host_v)
{
    if (!m) return LB_ERR_INVALID;
    for (size_t k = 0; k < m->slabs.size(); ++k) {
        lb_sim *s = m->slabs[k];
        auto at = [&](const void *p, int eb) { return p ? (const void *)((const char *)p + (size_t)m->x0[k] * eb) : nullptr; };
        MS(k, upload_moments_impl(s, at(host_rho, s->elem), at(host_u, s->uv_elem), at(host_v, s->uv_elem), (size_t)m->cfg.nx));
    }
    return LB_OK;
}

int lb_multi_init_synthetic(lb_multi *m, int kind, double u0, double amplitude, uint64_t seed)
{
    if (!m) return LB_ERR_INVALID;
    for (size_t k = 0; k < m->slabs.size(); ++k) MS(k, lb_init_synthetic(m->slabs[k], kind, u0, amplitude, seed));
    return multi_prime(m);
}

int lb_multi_set_temporal_blocking(lb_multi *m, int shape)
{
    if (!m) return LB_ERR_INVALID;
    for (size_t k = 0; k < m->slabs.size(); ++k) MS(k, lb_set_temporal_blocking(m->slabs[k], shape));
    return LB_OK;
}

int lb_multi_temporal_blocking(const lb_multi *m) { return (m && !m->slabs.empty()) ? tb2_effective_shape(m->slabs[0]) : 0; }

int lb_multi_prime(lb_multi *m) { return m ? multi_prime(m) : LB_ERR_INVALID; }

/* the hot path: Pipe_Flow.run (opencl_dim.py:372-387) on every slab; no host synchronisation inside */
int lb_multi_step(lb_multi *m, int n_steps)
{
    if (!m) return LB_ERR_INVALID;
    if (n_steps < 0) return mfail(m, LB_ERR_INVALID, "lb_multi_step: negative step count");
    if (n_steps == 0) return LB_OK;
    const size_t n = m->slabs.size();
    if (n == 1) { MS(0, step_impl(m->slabs[0], n_steps, true)); return LB_OK; }
    if (!m->primed) { int rc = multi_prime(m); if (rc) return rc; }
    if (m->concurrent) {
        // one slab per device: enqueue bounded chunks round-robin; the kernels of neighbouring devices
        // synchronise among themselves through the peer-memory flags.  Every slab cuts a chunk into the same
        // sequence of launches (plan_launches); 30 is a multiple of both launch depths, the remainder goes first.
        const int CHUNK = 30;
        int done = 0;
        while (done < n_steps) {
            const int chunk = (done == 0 && n_steps % CHUNK) ? n_steps % CHUNK : std::min(CHUNK, n_steps - done);
            const bool last = done + chunk == n_steps;
            for (size_t k = 0; k < n; ++k) MS(k, step_impl(m->slabs[k], chunk, last));
            done += chunk;
        }
        return LB_OK;
    }
    // slabs sharing a device share its stream: advance in lock-step, ONE LAUNCH per slab at a time -- a slab's
    // launch waits in-kernel for what its neighbour's previous launch published, and on one stream that
    // launch must already be enqueued ahead of it
    const int tb = tb2_effective_shape(m->slabs[0]);
    std::vector<int> plan;
    if (tb_kind(tb) == LB_TB_MARCH) plan = plan_launches(m->slabs[0], tb, n_steps);
    else plan.assign(n_steps, 1);
    for (size_t j = 0; j < plan.size(); ++j)
        for (size_t k = 0; k < n; ++k) MS(k, step_impl(m->slabs[k], plan[j], j + 1 == plan.size()));
    return LB_OK;
}

int lb_multi_sync(lb_multi *m)
{
    if (!m) return LB_ERR_INVALID;
    int first = LB_OK;
    for (size_t k = 0; k < m->slabs.size(); ++k) {
        const int rc = mslab(m, k, lb_sync(m->slabs[k]));
        if (rc != LB_OK && first == LB_OK) first = rc;
    }
    return first;
}

int lb_multi_download(lb_multi *m, int field, void *host_out)
{
    if (!m || !host_out) return mfail(m, LB_ERR_INVALID, "lb_multi_download: null argument");
    for (size_t k = 0; k < m->slabs.size(); ++k) {       // all copies in flight together, one sync each afterwards
        lb_sim *s = m->slabs[k];
        const int eb = (field == LB_FIELD_U || field == LB_FIELD_V) ? s->uv_elem : s->elem;
        MS(k, download_impl(s, field, (char *)host_out + (size_t)m->x0[k] * eb, (size_t)m->cfg.nx, false));
    }
    return lb_multi_sync(m);
}

int lb_multi_total_mass(lb_multi *m, double *out)
{
    if (!m || !out) return LB_ERR_INVALID;
    long double acc = 0;
    for (size_t k = 0; k < m->slabs.size(); ++k) { double v = 0; MS(k, lb_total_mass(m->slabs[k], &v)); acc += v; }
    *out = (double)acc;
    return LB_OK;
}

int lb_multi_checksum(lb_multi *m, uint64_t *out)
{
    if (!m || !out) return LB_ERR_INVALID;
    uint64_t acc = 0;
    for (size_t k = 0; k < m->slabs.size(); ++k) { uint64_t v = 0; MS(k, lb_checksum(m->slabs[k], &v)); acc += v; }
    *out = acc;
    return LB_OK;
}

int64_t lb_multi_launch_count(const lb_multi *m)
{
    int64_t n = 0;
    if (m) for (const lb_sim *s : m->slabs) n += s->launches;
    return n;
}

}  // extern "C"
