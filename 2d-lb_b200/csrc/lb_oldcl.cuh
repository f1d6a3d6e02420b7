// Scheme "opencl_old": LB_D2Q9/OLD/opencl.py's velocity-inlet / y-periodic classes
// (Pipe_Flow_PeriodicBC_VelocityInlet, Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet, :281-371) as one
// fused CUDA kernel per step -- SURVEY.md 8f-2, OpenCL flavour.
//
// Those classes are the only callers of D2Q9.cl's move_bcs_PeriodicBC_VelocityInlet (:263-321) and
// update_hydro_PeriodicBC_VelocityInlet (:323-374), and they run the kernels in the OLD order
// (OLD/opencl.py:246-255):
//     move_bcs -> [bounce-back] -> move + copy_buffer -> update_hydro -> [u=v=0 in the obstacle]
//     -> update_feq -> collide_particles
// which differs from the `opencl` scheme of lb_fused.cuh in three observable ways:
//   * the boundary pass acts on POST-collision populations, before streaming;
//   * `move` never writes the slots that have no upstream node, so after copy_buffer those hold what
//     f_streamed was created with -- the initial populations -- at every step ("frozen" lines:
//     1,5,8 on x=0; 3,6,7 on x=nx-1; 2,5,6 on y=0; 4,7,8 on y=ny-1), and they do enter the moments;
//   * update_hydro leaves v on the inlet/outlet columns and u in the four corners untouched.
// All arithmetic is float32 with the double-literal promotions of the OpenCL C source, mirrored so
// that the result is BIT-IDENTICAL to the reference's own kernels (tests/golden/oldcl_*.npz).
//
// Fusion order, as in lb_cython.cuh: a launch does  stream(pull) -> moments -> feq -> collide ->
// [inlet/outlet closure + bounce-back for the NEXT step, rows 1..ny-2], and `oc_rows_kernel` then
// exchanges the two periodic rows (and bounces their solid nodes): the row copies read the other
// wall row, which another CTA owns.  The last launch of a run skips the bracket; the next run starts
// with `oc_prestream_kernel` + `oc_rows_kernel`.  Storage float32, 72 B per lattice update.
#pragma once
#include "lb_fused.cuh"

namespace lb {

struct OcParams {
    const float *src;
    float *dst;
    long long plane;
    int nx, ny, pitch;
    int write_moments;       // last launch of a run: store rho, u, v
    int apply_next_bc;       // every launch but the last
    const uint8_t *mask;     // [ny][mask_pitch] or nullptr
    int mask_pitch;
    float *rho, *u, *v;      // [ny][pitch] float32
    const float *frozen;     // west[3][ny] | east[3][ny] | south[3][nx] | north[3][nx]
    Consts<float> c;
    float u_w, u_e;          // np.float32(self.u_w), np.float32(self.u_e)   (OLD/opencl.py:293-294)
    double kw, ke;           // 1./(1.-u_w), 1./(1.+u_e) evaluated in double   (D2Q9.cl:292, :299)
};

__host__ __device__ inline const float *oc_frozen_w(const float *fr, int, int) { return fr; }
__host__ __device__ inline const float *oc_frozen_e(const float *fr, int, int ny) { return fr + 3 * ny; }
__host__ __device__ inline const float *oc_frozen_s(const float *fr, int, int ny) { return fr + 6 * ny; }
__host__ __device__ inline const float *oc_frozen_n(const float *fr, int nx, int ny) { return fr + 6 * ny + 3 * nx; }
inline size_t oc_frozen_floats(int nx, int ny) { return (size_t)6 * ny + (size_t)6 * nx; }

// the populations `move` will never overwrite, taken from the freshly uploaded f (lb_upload_f seeds
// f_streamed with the same data, OLD/opencl.py:221-222)
__global__ void oc_capture_frozen_kernel(int nx, int ny, int pitch, long long plane, const float *f, float *fr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ny) {
        const long long w = (long long)i * pitch, e = w + nx - 1;
        float *fw = fr, *fe = fr + 3 * ny;
        fw[i] = f[1 * plane + w]; fw[ny + i] = f[5 * plane + w]; fw[2 * ny + i] = f[8 * plane + w];
        fe[i] = f[3 * plane + e]; fe[ny + i] = f[6 * plane + e]; fe[2 * ny + i] = f[7 * plane + e];
    }
    if (i < nx) {
        const long long s = i, n = (long long)(ny - 1) * pitch + i;
        float *fs = fr + 6 * ny, *fn = fr + 6 * ny + 3 * nx;
        fs[i] = f[2 * plane + s]; fs[nx + i] = f[5 * plane + s]; fs[2 * nx + i] = f[6 * plane + s];
        fn[i] = f[4 * plane + n]; fn[nx + i] = f[7 * plane + n]; fn[2 * nx + i] = f[8 * plane + n];
    }
}

// ---- inlet / outlet closure of one node, rows 1..ny-2 (D2Q9.cl:291-303), in place on g ----------
__device__ __forceinline__ void oc_velocity_bc(const OcParams &p, int x, int y, float (&g)[9])
{
    if (y < 1 || y >= p.ny - 1) return;
    const float f0 = g[0], f1 = g[1], f2 = g[2], f3 = g[3], f4 = g[4], f5 = g[5], f6 = g[6], f7 = g[7], f8 = g[8];
    if (x == 0) {                // `2*(...)`: an int factor, the sum stays float
        const float rho_w = (float)(p.kw * (double)(((f0 + f2) + f4) + 2.0f * ((f3 + f6) + f7)));
        g[1] = (float)((double)f3 + ((2. / 3.) * (double)rho_w) * (double)p.u_w);
        g[5] = (float)(((double)f7 - (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)p.u_w);
        g[8] = (float)(((double)f6 + (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)p.u_w);
    }
    if (x == p.nx - 1) {         // `2.*(...)`: promotes the sum to double
        const float rho_e = (float)(p.ke * ((double)((f0 + f2) + f4) + 2. * (double)((f1 + f5) + f8)));
        g[3] = (float)((double)f1 - ((2. / 3.) * (double)rho_e) * (double)p.u_e);
        g[6] = (float)(((double)f5 + (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)p.u_e);
        g[7] = (float)(((double)f8 - (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)p.u_e);
    }
}

// start of a run: closure + bounce-back of rows 1..ny-2, in place (rows 0 and ny-1: oc_rows_kernel)
__global__ void oc_prestream_kernel(OcParams p, float *f)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= p.nx || y < 1 || y >= p.ny - 1) return;
    const bool solid = p.mask && p.mask[(long long)y * p.mask_pitch + x] == 1;
    if (!solid && x != 0 && x != p.nx - 1) return;
    const long long i = (long long)y * p.pitch + x;
    float g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * p.plane + i];
    oc_velocity_bc(p, x, y, g);
    if (solid) bounce_back<float>(g);
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j * p.plane + i] = g[j];
}

// the periodic rows (D2Q9.cl:305-318): 4,8,7 of row ny-1 <- row 0 and 2,6,5 of row 0 <- row ny-1, both
// from the values before the pass; then bounce-back of the solid nodes of the two rows (:410-431)
__global__ void oc_rows_kernel(int nx, int ny, int pitch, long long plane, float *f, const uint8_t *mask, int mask_pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    const long long s = x, n = (long long)(ny - 1) * pitch + x;
    float a[9], b[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { a[j] = f[j * plane + s]; b[j] = f[j * plane + n]; }
    a[2] = b[2]; a[6] = b[6]; a[5] = b[5];                 // b's 2, 6, 5 are still the original values
    b[4] = a[4]; b[8] = a[8]; b[7] = a[7];                 // a's 4, 8, 7 were not touched above
    if (mask && mask[x] == 1) bounce_back<float>(a);
    if (mask && mask[(long long)(ny - 1) * mask_pitch + x] == 1) bounce_back<float>(b);
#pragma unroll
    for (int j = 1; j < 9; ++j) { f[j * plane + s] = a[j]; f[j * plane + n] = b[j]; }
}

// ---- single stages (OLD/opencl.py's move / update_hydro as separate calls; move_bcs is
//      oc_prestream_kernel + oc_rows_kernel, update_feq and collide_particles are the opencl scheme's) ----

// `move` + `copy_buffer` (D2Q9.cl:139-171, :123-137): slots without an upstream node take the frozen value
__global__ void oc_stage_move_kernel(int nx, int ny, int pitch, long long plane, const float *src, float *dst,
                                     const float *frozen)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const int lx = nx - 1, ly = ny - 1;
    const long long i = (long long)y * pitch + x;
    float g[9];
    const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int sx = x - cx[j], sy = y - cy[j];
        g[j] = (sx >= 0 && sx <= lx && sy >= 0 && sy <= ly) ? src[j * plane + (long long)sy * pitch + sx] : 0.f;
    }
    const float *fw = oc_frozen_w(frozen, nx, ny), *fe = oc_frozen_e(frozen, nx, ny);
    const float *fs = oc_frozen_s(frozen, nx, ny), *fn = oc_frozen_n(frozen, nx, ny);
    if (x == 0) { g[1] = fw[y]; g[5] = fw[ny + y]; g[8] = fw[2 * ny + y]; }
    if (x == lx) { g[3] = fe[y]; g[6] = fe[ny + y]; g[7] = fe[2 * ny + y]; }
    if (y == 0) { g[2] = fs[x]; g[5] = fs[nx + x]; g[6] = fs[2 * nx + x]; }
    if (y == ly) { g[4] = fn[x]; g[7] = fn[nx + x]; g[8] = fn[2 * nx + x]; }
#pragma unroll
    for (int j = 0; j < 9; ++j) dst[j * plane + i] = g[j];
}

// update_hydro_PeriodicBC_VelocityInlet (D2Q9.cl:323-374) [+ set_zero_velocity_in_obstacle with a mask]
__global__ void oc_stage_hydro_kernel(OcParams p, const float *f)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= p.nx || y >= p.ny) return;
    const int lx = p.nx - 1, ly = p.ny - 1;
    const long long i = (long long)y * p.pitch + x;
    float g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * p.plane + i];
    float rho, u, v;
    moments<float, MATH_STRICT>(g, rho, u, v);
    if (x == 0 || x == lx) {
        u = p.u[i]; v = p.v[i];
        if (y != 0 && y < ly) {
            if (x == 0) {
                rho = (float)(p.kw * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[3] + g[6]) + g[7])));
                u = p.u_w;
            }
            if (x == lx) {
                rho = (float)(p.ke * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[1] + g[5]) + g[8])));
                u = p.u_e;
            }
        }
    }
    if (p.mask && p.mask[(long long)y * p.mask_pitch + x] == 1) { u = 0.f; v = 0.f; }
    p.rho[i] = rho; p.u[i] = u; p.v[i] = v;
}

// ---- the fused step ---------------------------------------------------------------------------
template <int WX, int WY, int MINB>
__global__ void __launch_bounds__(32 * WX * WY, MINB) fused_step_oldcl_kernel(const OcParams p)
{
    constexpr int V = 4, SPAN = 32 * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wx = warp % WX, wy = warp / WX;
    const int span0 = (blockIdx.x * WX + wx) * SPAN;
    const int x0 = span0 + lane * V;
    const int y = (blockIdx.z * gridDim.y + blockIdx.y) * WY + wy;
    if (span0 >= p.pitch || y >= p.ny) return;             // warp-uniform

    const float *__restrict__ src = p.src;
    float *__restrict__ dst = p.dst;
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch, lx = nx - 1, ly = ny - 1;
    const long long rc = (long long)y * pitch + x0;
    const float *pc = src + rc;
    const float *pm = pc - pitch + 2 * plane;              // row y-1: populations 2, 5, 6 (guard rows below y=0)
    const float *pp = pc + pitch + 4 * plane;              // row y+1: populations 4, 7, 8

    Pack<float, V> q[9];
    q[0] = load_pack<float, V, 1>(pc);
    q[1] = load_pack<float, V, 1>(pc + plane);
    q[3] = load_pack<float, V, 1>(pc + 3 * plane);
    q[2] = load_pack<float, V, 1>(pm);
    q[5] = load_pack<float, V, 1>(pm + 3 * plane);
    q[6] = load_pack<float, V, 1>(pm + 4 * plane);
    q[4] = load_pack<float, V, 1>(pp);
    q[7] = load_pack<float, V, 1>(pp + 3 * plane);
    q[8] = load_pack<float, V, 1>(pp + 4 * plane);
    float l1 = 0.f, l5 = 0.f, l8 = 0.f, r3 = 0.f, r6 = 0.f, r7 = 0.f;
    if (lane == 0) { l1 = pc[plane - 1]; l5 = pm[3 * plane - 1]; l8 = pp[4 * plane - 1]; }
    if (lane == 31) { r3 = pc[3 * plane + V]; r6 = pm[4 * plane + V]; r7 = pp[3 * plane + V]; }
    {
        const float s1 = __shfl_up_sync(0xffffffffu, q[1].v[V - 1], 1);
        const float s5 = __shfl_up_sync(0xffffffffu, q[5].v[V - 1], 1);
        const float s8 = __shfl_up_sync(0xffffffffu, q[8].v[V - 1], 1);
        const float s3 = __shfl_down_sync(0xffffffffu, q[3].v[0], 1);
        const float s6 = __shfl_down_sync(0xffffffffu, q[6].v[0], 1);
        const float s7 = __shfl_down_sync(0xffffffffu, q[7].v[0], 1);
        if (lane != 0) { l1 = s1; l5 = s5; l8 = s8; }
        if (lane != 31) { r3 = s3; r6 = s6; r7 = s7; }
    }
#pragma unroll
    for (int e = V - 1; e > 0; --e) { q[1].v[e] = q[1].v[e - 1]; q[5].v[e] = q[5].v[e - 1]; q[8].v[e] = q[8].v[e - 1]; }
    q[1].v[0] = l1; q[5].v[0] = l5; q[8].v[0] = l8;
#pragma unroll
    for (int e = 0; e < V - 1; ++e) { q[3].v[e] = q[3].v[e + 1]; q[6].v[e] = q[6].v[e + 1]; q[7].v[e] = q[7].v[e + 1]; }
    q[3].v[V - 1] = r3; q[6].v[V - 1] = r6; q[7].v[V - 1] = r7;

    // --- slots without an upstream node hold the frozen initial populations (see the header) ---
    const int el_east = lx - x0;                           // element index of the outlet column, if in [0, V)
    const bool on_boundary = (y == 0 || y == ly || x0 == 0 || (el_east >= 0 && el_east < V));
    float eu[V], ev[V];                                    // stored u, v of the inlet/outlet columns
#pragma unroll
    for (int e = 0; e < V; ++e) { eu[e] = 0.f; ev[e] = 0.f; }
    if (on_boundary) {
        const float *fw = oc_frozen_w(p.frozen, nx, ny), *fe = oc_frozen_e(p.frozen, nx, ny);
        const float *fs = oc_frozen_s(p.frozen, nx, ny), *fn = oc_frozen_n(p.frozen, nx, ny);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int x = x0 + e;
            if (x > lx) continue;
            if (x == 0) { q[1].v[e] = fw[y]; q[5].v[e] = fw[ny + y]; q[8].v[e] = fw[2 * ny + y]; }
            if (x == lx) { q[3].v[e] = fe[y]; q[6].v[e] = fe[ny + y]; q[7].v[e] = fe[2 * ny + y]; }
            if (y == 0) { q[2].v[e] = fs[x]; q[5].v[e] = fs[nx + x]; q[6].v[e] = fs[2 * nx + x]; }
            if (y == ly) { q[4].v[e] = fn[x]; q[7].v[e] = fn[nx + x]; q[8].v[e] = fn[2 * nx + x]; }
            if (x == 0 || x == lx) { eu[e] = p.u[rc + e]; ev[e] = p.v[rc + e]; }
        }
    }

    uint32_t solid_bits = 0;
    if (p.mask != nullptr) {
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (x0 + e < nx && p.mask[(long long)y * p.mask_pitch + x0 + e] == 1) solid_bits |= (1u << e);
    }

    float mrho[V], mu[V], mv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int x = x0 + e;
        float g[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
        const bool solid = (solid_bits >> e) & 1u;
        // update_hydro_PeriodicBC_VelocityInlet (D2Q9.cl:351-371) [+ set_zero_velocity_in_obstacle]
        float rho, u, v;
        moments<float, MATH_STRICT>(g, rho, u, v);
        if (x == 0 || x == lx) {
            u = eu[e]; v = ev[e];                          // never written there ...
            if (y != 0 && y < ly) {                        // ... except u (and rho) on rows 1..ny-2
                if (x == 0) {
                    rho = (float)(p.kw * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[3] + g[6]) + g[7])));
                    u = p.u_w;
                }
                if (x == lx) {
                    rho = (float)(p.ke * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[1] + g[5]) + g[8])));
                    u = p.u_e;
                }
            }
        }
        if (solid) { u = 0.f; v = 0.f; }
        mrho[e] = rho; mu[e] = u; mv[e] = v;
        float feq[9];
        feq_strict<float>(p.c, rho, u, v, feq);
#pragma unroll
        for (int j = 0; j < 9; ++j) g[j] = g[j] * p.c.keep + p.c.omega * feq[j];
        if (p.apply_next_bc && y >= 1 && y < ly && x <= lx) {
            if (on_boundary) oc_velocity_bc(p, x, y, g);
            if (solid) bounce_back<float>(g);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) q[j].v[e] = g[j];
    }

    if (x0 + V <= nx) {
        float *pd = dst + rc;
#pragma unroll
        for (int j = 0; j < 9; ++j) store_pack<float, V, 0>(pd + j * plane, q[j]);
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (x0 + e < nx) {
#pragma unroll
                for (int j = 0; j < 9; ++j) dst[j * plane + rc + e] = q[j].v[e];
            }
    }
    if (p.write_moments) {
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (x0 + e < nx) {
                p.rho[rc + e] = mrho[e];
                p.u[rc + e] = mu[e];
                p.v[rc + e] = mv[e];
            }
    }
}

}  // namespace lb
