// Scheme "cython": the reference's CPU path (LB_D2Q9/dimensionless/cython_dim.pyx:204-359,
// :459-513 and LB_D2Q9/OLD/cython.pyx) as ONE fused CUDA kernel per step -- SURVEY.md 8f-1 / A.3.
//
// Why a second kernel: the Cython classes are a different algorithm from the OpenCL ones
// (SURVEY.md F3): boundary closures are applied BEFORE streaming and use the velocity stored by
// the previous step, walls are plain reflections, four populations do not stream along one
// boundary line each, the moments are overridden on the boundary, and the arithmetic is NumPy's
// mixed float32/float64 (f, feq, rho float32; u, v and the whole equilibrium in float64).
// This kernel reproduces all of that so that its output can be compared BIT FOR BIT with vectors
// produced by the unmodified, compiled reference (tests/golden/*.npz).
//
// Fusion order.  The reference's step is  BC -> obstacle swap -> stream -> moments -> feq -> collide.
// A pull kernel cannot apply another node's BC while streaming, so the closure is moved to the END
// of the previous step: a launch does  stream(pull) -> moments -> feq -> collide -> [BC + swap for
// the NEXT step], and stores that.  The last launch of a run skips the bracketed part (so f on the
// device is the reference's post-collision f), and the next run starts with `cy_prestream_kernel`,
// which applies BC + swap in place from the stored u.  Same layout, thread mapping, vector loads and
// shuffles as lb_fused.cuh; storage is float32, 72 B per lattice update.
#pragma once
#include "lb_fused.cuh"

namespace lb {

struct CyConsts {
    double omega, keep;            // np.float64 omega, 1.-omega          (cython_dim.pyx:88, :342)
    double rin, rout;              // inlet_rho, outlet_rho as the host computed them
    double cs2, cs22, cssq;        // cs**2, 2*cs2, 2.0/9.0                (cython_dim.pyx:20-23)
    double i_cs2, i_cs22, i_cssq;  // RN(1/c) for div_const
    double one_m_cs2;              // 1.-cs2                                (:174)
    double tw_rin, tw_rout;        // (2./3.)*inlet_rho, (2./3.)*outlet_rho (:217, :222)
    float rin_f, rout_f;           // cdef float inlet_rho / outlet_rho      (:227-228)
    float keep_f, om_f;            // OLD flavour: omega is a Python float -> float32 relaxation
    float w0, w1, w2;
    // velocity-inlet / y-periodic family (OLD/cython.pyx:268-360)
    double u_w, u_e;               // Python floats self.u_w, self.u_e
    float u_w_f, u_e_f;            // `cdef float u_w, u_e` in move_bcs (:283-284)
    double kw_d, ke_d;             // 1./(1.-u_w), 1./(1.+u_e) with the float u_w / u_e (C code, :295, :302)
    float kw_f, ke_f;              // the same factors as NumPy's float32 scalars in update_hydro (:355, :358)
};

inline CyConsts make_cy_consts(double omega, double rin, double rout, double cs2, double cs22,
                               double u_w = 0.0, double u_e = 0.0)
{
    CyConsts c;
    c.omega = omega; c.keep = 1. - omega;
    c.rin = rin; c.rout = rout;
    c.cs2 = cs2; c.cs22 = cs22; c.cssq = 2.0 / 9.0;
    c.i_cs2 = 1.0 / c.cs2; c.i_cs22 = 1.0 / c.cs22; c.i_cssq = 1.0 / c.cssq;
    c.one_m_cs2 = 1. - cs2;
    c.tw_rin = (2. / 3.) * rin; c.tw_rout = (2. / 3.) * rout;
    c.rin_f = (float)rin; c.rout_f = (float)rout;
    c.keep_f = (float)c.keep; c.om_f = (float)omega;
    c.w0 = (float)(4. / 9.); c.w1 = (float)(1. / 9.); c.w2 = (float)(1. / 36.);
    c.u_w = u_w; c.u_e = u_e;
    c.u_w_f = (float)u_w; c.u_e_f = (float)u_e;
    c.kw_d = 1. / (1. - (double)c.u_w_f); c.ke_d = 1. / (1. + (double)c.u_e_f);
    c.kw_f = (float)(1. / (1. - u_w)); c.ke_f = (float)(1. / (1. + u_e));
    return c;
}

struct CyParams {
    const float *src;
    float *dst;
    long long plane;
    int nx, ny, pitch;
    int write_moments;       // last launch of a run: store rho (f32), u, v (f64)
    int apply_next_bc;       // every launch but the last: BC + swap for the next step before storing
    const uint8_t *mask;     // [ny][mask_pitch] or nullptr
    int mask_pitch;
    float *rho;
    double *u, *v;           // [ny][pitch] float64
    CyConsts c;
};

// ---- the pre-stream boundary closure of one node (cython_dim.pyx:204-269), in place on g.
//      U = this node's stored x-velocity (float64).
__device__ __forceinline__ void cy_prestream_bc(const CyConsts &c, int x, int y, int lx, int ly, double U, float (&g)[9])
{
    const bool west = (x == 0), east = (x == lx), south = (y == 0), north = (y == ly);
    if (!(west || east || south || north)) return;
    const float f0 = g[0], f1 = g[1], f2 = g[2], f3 = g[3], f4 = g[4], f5 = g[5], f6 = g[6], f7 = g[7], f8 = g[8];
    if (west && !south && !north) {                       // :217-219
        g[1] = (float)((double)f3 + c.tw_rin * U);
        g[5] = (float)((double)(((-.5f * f2) + (.5f * f4)) + f7) + ((1. / 6.) * U) * c.rin);
        g[8] = (float)((double)(((.5f * f2) - (.5f * f4)) + f6) + ((1. / 6.) * U) * c.rin);
    } else if (east && !south && !north) {                // :222-224
        g[3] = (float)((double)f1 - c.tw_rout * U);
        g[6] = (float)((double)(((-.5f * f2) + (.5f * f4)) + f8) - ((1. / 6.) * U) * c.rout);
        g[7] = (float)((double)(((.5f * f2) - (.5f * f4)) + f5) - ((1. / 6.) * U) * c.rout);
    } else if (north && !west && !east) {                 // :232-235 plain reflection
        g[4] = f2; g[8] = f6; g[7] = f5;
    } else if (south && !west && !east) {                 // :237-240
        g[2] = f4; g[6] = f8; g[5] = f7;
    } else if (west && south) {                           // :244-248 (Cython emits `2` as 2.0: double expression)
        const double t = ((((double)(-f0) - 2.0 * (double)f3) - 2.0 * (double)f4) - 2.0 * (double)f7) + (double)c.rin_f;
        g[1] = f3; g[2] = f4; g[5] = f7;
        g[6] = (float)(.5 * t); g[8] = g[6];
    } else if (west && north) {                           // :251-255
        const double t = ((((double)(-f0) - 2.0 * (double)f2) - 2.0 * (double)f3) - 2.0 * (double)f6) + (double)c.rin_f;
        g[1] = f3; g[4] = f2; g[8] = f6;
        g[5] = (float)(.5 * t); g[7] = g[5];
    } else if (east && south) {                           // :258-262
        const double t = ((((double)(-f0) - 2.0 * (double)f1) - 2.0 * (double)f4) - 2.0 * (double)f8) + (double)c.rout_f;
        g[3] = f1; g[2] = f4; g[6] = f8;
        g[5] = (float)(.5 * t); g[7] = g[5];
    } else {                                              // :265-269
        const double t = ((((double)(-f0) - 2.0 * (double)f1) - 2.0 * (double)f2) - 2.0 * (double)f5) + (double)c.rout_f;
        g[3] = f1; g[4] = f2; g[7] = f5;
        g[6] = (float)(.5 * t); g[8] = g[6];
    }
}

// ---- velocity-inlet closure of one node (OLD/cython.pyx:291-303): inlet / outlet nodes with
//      1 <= y < ly only; the row exchange of :305-316 is folded into the pull (see the kernel).
__device__ __forceinline__ void cyv_prestream_bc(const CyConsts &c, int x, int y, int lx, int ly, float (&g)[9])
{
    if (y < 1 || y >= ly) return;
    const float f0 = g[0], f1 = g[1], f2 = g[2], f3 = g[3], f4 = g[4], f5 = g[5], f6 = g[6], f7 = g[7], f8 = g[8];
    if (x == 0) {
        const float rho_w = (float)(c.kw_d * ((double)((f0 + f2) + f4) + 2. * (double)((f3 + f6) + f7)));
        g[1] = (float)((double)f3 + ((2. / 3.) * (double)rho_w) * (double)c.u_w_f);
        g[5] = (float)(((double)f7 - (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)c.u_w_f);
        g[8] = (float)(((double)f6 + (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)c.u_w_f);
    } else if (x == lx) {
        const float rho_e = (float)(c.ke_d * ((double)((f0 + f2) + f4) + 2. * (double)((f1 + f5) + f8)));
        g[3] = (float)((double)f1 - ((2. / 3.) * (double)rho_e) * (double)c.u_e_f);
        g[7] = (float)(((double)f5 + (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)c.u_e_f);
        g[6] = (float)(((double)f8 - (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)c.u_e_f);
    }
}

// ---- moments of the velocity-inlet class (OLD/cython.pyx:331-360, :375-378) ---------------------
__device__ __forceinline__ void cyv_moments(const CyConsts &c, const float (&g)[9], int x, int y, int lx, int ly,
                                            bool solid, float &rho, double &u, double &v)
{
    float r = g[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) r = r + g[j];
    const float inv = 1.0f / r;
    u = (double)((((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv);
    v = (double)((((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv);
    rho = r;
    if (y >= 1 && y < ly) {
        if (x == 0) {
            u = c.u_w;
            rho = c.kw_f * (((g[0] + g[2]) + g[4]) + 2.0f * ((g[3] + g[6]) + g[7]));
        }
        if (x == lx) {
            u = c.u_e;
            rho = c.ke_f * (((g[0] + g[2]) + g[4]) + 2.0f * ((g[1] + g[5]) + g[8]));
        }
    }
    if (solid) { u = 0.0; v = 0.0; }
}

// ---- moments with the boundary overrides (cython_dim.pyx:302-333, :459-466) -------------------
template <bool OLD>
__device__ __forceinline__ void cy_moments(const CyConsts &c, const float (&g)[9], int x, int y, int lx, int ly,
                                           bool solid, float &rho, double &u, double &v)
{
    float r = g[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) r = r + g[j];                      // np.sum(f, axis=0), float32
    const float inv = 1.0f / r;                                    // 1./rho, float32
    u = (double)((((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv);
    v = (double)((((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv);
    rho = r;
    if (!OLD && (y == 0 || y == ly)) { u = 0.0; v = 0.0; }         // :317-320 (absent from OLD/cython.pyx)
    if (x == 0) {                                                  // :325, :328
        rho = c.rin_f;
        const float a = (g[0] + g[2]) + g[4];
        const float b = 2.0f * ((g[3] + g[6]) + g[7]);
        // dimensionless: inlet_rho is np.float64 -> float64 expression; OLD: Python float -> float32
        u = OLD ? (double)(1.0f - (a + b) / c.rin_f) : 1. - (double)(a + b) / c.rin;
    }
    if (x == lx) {                                                 // :326, :331
        rho = c.rout_f;
        const float a = (g[0] + g[2]) + g[4];
        const float b = 2.0f * ((g[1] + g[5]) + g[8]);
        // dimensionless: outlet_rho is the Python float 1. -> float32; OLD: np.float64 -> float64
        u = OLD ? -1. + (double)(a + b) / c.rout : (double)(-1.0f + (a + b) / c.rout_f);
    }
    if (solid) { u = 0.0; v = 0.0; }
}

// ---- Succi's factored equilibrium in float64 (cython_dim.pyx:160-189) + relaxation (:336-344) ----
template <bool OLD>
__device__ __forceinline__ void cy_collide(const CyConsts &c, float (&g)[9], float rho, double u, double v)
{
    const double ul = div_const(u, c.cs2, c.i_cs2), vl = div_const(v, c.cs2, c.i_cs2);
    const double uv = ul * vl;
    const double usq = u * u, vsq = v * v;
    const double sumsq = div_const(usq + vsq, c.cs22, c.i_cs22);
    const double sumsq2 = div_const(sumsq * c.one_m_cs2, c.cs2, c.i_cs2);
    const double u2 = div_const(usq, c.cssq, c.i_cssq), v2 = div_const(vsq, c.cssq, c.i_cssq);
    const double r0 = (double)(c.w0 * rho), r1 = (double)(c.w1 * rho), r2 = (double)(c.w2 * rho);
    float feq[9];
    feq[0] = (float)(r0 * (1. - sumsq));
    feq[1] = (float)(r1 * (((1. - sumsq) + u2) + ul));
    feq[2] = (float)(r1 * (((1. - sumsq) + v2) + vl));
    feq[3] = (float)(r1 * (((1. - sumsq) + u2) - ul));
    feq[4] = (float)(r1 * (((1. - sumsq) + v2) - vl));
    feq[5] = (float)(r2 * ((((1. + sumsq2) + ul) + vl) + uv));
    feq[6] = (float)(r2 * ((((1. + sumsq2) - ul) + vl) - uv));
    feq[7] = (float)(r2 * ((((1. + sumsq2) - ul) - vl) + uv));
    feq[8] = (float)(r2 * ((((1. + sumsq2) + ul) - vl) - uv));
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        if (OLD) g[j] = g[j] * c.keep_f + c.om_f * feq[j];
        else g[j] = (float)((double)g[j] * c.keep + c.omega * (double)feq[j]);
    }
}

// equilibrium only (init path: cython_dim.pyx:105 update_feq before init_pop)
__global__ void cy_feq_kernel(int nx, int ny, int pitch, long long plane, const float *rho, const double *u,
                              const double *v, float *feq, CyConsts c)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    float g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = 0.0f;
    // reuse cy_collide's algebra with keep = 0, omega = 1: g <- feq exactly ((double)0*0 + 1*(double)feq)
    CyConsts c1 = c;
    c1.keep = 0.0; c1.omega = 1.0;
    cy_collide<false>(c1, g, rho[i], u[i], v[i]);
#pragma unroll
    for (int j = 0; j < 9; ++j) feq[j * plane + i] = g[j];
}

// BC + obstacle swap in place, from the stored u (start of a run; cython_dim.pyx:204-269, :486-513)
__global__ void cy_prestream_kernel(int nx, int ny, int pitch, long long plane, float *f, const double *u,
                                    const uint8_t *mask, int mask_pitch, CyConsts c, int velocity_inlet)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const bool solid = mask && mask[(long long)y * mask_pitch + x] == 1;
    const bool bnd = (x == 0 || x == nx - 1 || y == 0 || y == ny - 1);
    if (!solid && !bnd) return;
    const long long i = (long long)y * pitch + x;
    float g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * plane + i];
    if (bnd) {
        if (velocity_inlet) cyv_prestream_bc(c, x, y, nx - 1, ny - 1, g);
        else cy_prestream_bc(c, x, y, nx - 1, ny - 1, u[i], g);
    }
    if (solid) bounce_back<float>(g);
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j * plane + i] = g[j];
}

// ---- single stages (the Cython classes' move_bcs / move / update_hydro / collide_particles are plain
//      methods a user can call one by one; cython_dim.pyx:204-344).  Per-node kernels, not hot. --------

// row exchange of the velocity-inlet class, physically (OLD/cython.pyx:305-316); the fused kernel folds
// it into the pull instead.  Idempotent: its sources are never its destinations.
__global__ void cyv_rows_kernel(int nx, int ny, int pitch, long long plane, float *f)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    const long long s = x, n = (long long)(ny - 1) * pitch + x;
    const float a4 = f[4 * plane + s], a8 = f[8 * plane + s], a7 = f[7 * plane + s];
    const float b2 = f[2 * plane + n], b6 = f[6 * plane + n], b5 = f[5 * plane + n];
    f[4 * plane + n] = a4; f[8 * plane + n] = a8; f[7 * plane + n] = a7;
    f[2 * plane + s] = b2; f[6 * plane + s] = b6; f[5 * plane + s] = b5;
}

// `move` (cython_dim.pyx:271-299): the in-place sweeps as a pull from a snapshot; slots the sweeps never
// write keep the node's own value (no upstream node, or one of the four non-streaming lines).
__global__ void cy_stage_move_kernel(int nx, int ny, int pitch, long long plane, const float *src, float *dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const int lx = nx - 1, ly = ny - 1;
    const bool w_ = (x == 0), e_ = (x == lx), s_ = (y == 0), n_ = (y == ly);
    const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    const bool keep[9] = {true, w_ || s_, e_ || s_, e_ || n_, w_ || n_, w_ || s_, e_ || s_, e_ || n_, w_ || n_};
    const long long i = (long long)y * pitch + x;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const long long from = keep[j] ? i : (long long)(y - cy[j]) * pitch + (x - cx[j]);
        dst[j * plane + i] = src[j * plane + from];
    }
}

// `update_hydro` (+ the obstacle classes' zeroing override)
template <bool OLD, bool VIN>
__global__ void cy_stage_hydro_kernel(int nx, int ny, int pitch, long long plane, const float *f, float *rho, double *u,
                                      double *v, const uint8_t *mask, int mask_pitch, CyConsts c)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
    float g[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) g[j] = f[j * plane + i];
    const bool solid = mask && mask[(long long)y * mask_pitch + x] == 1;
    float r;
    double uu, vv;
    if (VIN) cyv_moments(c, g, x, y, nx - 1, ny - 1, solid, r, uu, vv);
    else cy_moments<OLD>(c, g, x, y, nx - 1, ny - 1, solid, r, uu, vv);
    rho[i] = r; u[i] = uu; v[i] = vv;
}

// `collide_particles` (cython_dim.pyx:336-344; OLD/cython.pyx: float32 because omega is a Python float)
template <bool OLD>
__global__ void cy_stage_collide_kernel(int nx, int ny, int pitch, long long plane, float *f, const float *feq, CyConsts c)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = lb_grid_row();
    if (x >= nx || y >= ny) return;
    const long long i = (long long)y * pitch + x;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const float a = f[j * plane + i], e = feq[j * plane + i];
        f[j * plane + i] = OLD ? a * c.keep_f + c.om_f * e : (float)((double)a * c.keep + c.omega * (double)e);
    }
}

// ---- the fused step -------------------------------------------------------------------------
// VIN = true: OLD/cython.pyx's velocity-inlet / y-periodic family.  Its row exchange
// (f4,f8,f7 of row ly <- row 0 ; f2,f6,f5 of row 0 <- row ly, before streaming) is folded into the
// pull: a population that would be read from row ly (0) after the exchange is read from row 0 (ly)
// of the stored field instead -- a warp-uniform change of the source row.
template <bool OLD, bool VIN, int WX, int WY, int MINB>
__global__ void __launch_bounds__(32 * WX * WY, MINB) fused_step_cython_kernel(const CyParams p)
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
    const CyConsts &c = p.c;
    const long long rc = (long long)y * pitch + x0;
    const int ym = (VIN && y == 1) ? ly : y - 1;          // row read by populations 2,5,6
    const int yp = (VIN && y == ly - 1) ? 0 : y + 1;      // row read by populations 4,7,8
    const float *pc = src + rc;
    const float *pm = src + ((long long)ym * pitch + x0) + 2 * plane;
    const float *pp = src + ((long long)yp * pitch + x0) + 4 * plane;

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

    // --- slots the reference's in-place sweeps never write keep the node's OWN pre-stream value
    //     (cython_dim.pyx:284-299): no upstream node, or one of the four non-streaming lines
    //     f2 @ x=lx, f1 @ y=0, f4 @ x=0, f3 @ y=ly.  Boundary threads only.
    const int el_east = lx - x0;
    const bool on_boundary = (y == 0 || y == ly || x0 == 0 || (el_east >= 0 && el_east < V));
    if (on_boundary) {
        Pack<float, V> own[9];
#pragma unroll
        for (int j = 1; j < 9; ++j) {
            long long ro = rc;
            if (VIN) {                                     // the node's own value AFTER the row exchange
                if (y == 0 && (j == 2 || j == 6 || j == 5)) ro = (long long)ly * pitch + x0;
                if (y == ly && (j == 4 || j == 8 || j == 7)) ro = x0;
            }
            own[j] = load_pack<float, V, 1>(src + j * plane + ro);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int x = x0 + e;
            if (x > lx) continue;
            const bool w_ = (x == 0), e_ = (x == lx), s_ = (y == 0), n_ = (y == ly);
            if (e_ || s_) q[2].v[e] = own[2].v[e];          // f2 written for x in [0,lx), y in [1,ly]
            if (e_ || s_) q[6].v[e] = own[6].v[e];          // f6: same range
            if (w_ || s_) q[1].v[e] = own[1].v[e];          // f1: x in [1,lx], y in [1,ly]
            if (w_ || s_) q[5].v[e] = own[5].v[e];
            if (w_ || n_) q[4].v[e] = own[4].v[e];          // f4: x in [1,lx], y in [0,ly)
            if (w_ || n_) q[8].v[e] = own[8].v[e];
            if (e_ || n_) q[3].v[e] = own[3].v[e];          // f3: x in [0,lx), y in [0,ly)
            if (e_ || n_) q[7].v[e] = own[7].v[e];
        }
    }

    // --- obstacle mask bits of this thread's nodes ---
    uint32_t solid_bits = 0;
    if (p.mask != nullptr) {
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (x0 + e < nx && p.mask[(long long)y * p.mask_pitch + x0 + e] == 1) solid_bits |= (1u << e);
    }

    // --- per node: moments (+overrides), equilibrium, relaxation, then next step's closure ---
    float mrho[V];
    double mu[V], mv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int x = x0 + e;
        float g[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
        const bool solid = (solid_bits >> e) & 1u;
        if (VIN) cyv_moments(c, g, x, y, lx, ly, solid, mrho[e], mu[e], mv[e]);
        else cy_moments<OLD>(c, g, x, y, lx, ly, solid, mrho[e], mu[e], mv[e]);
        cy_collide<OLD>(c, g, mrho[e], mu[e], mv[e]);
        if (p.apply_next_bc) {
            if (on_boundary && x <= lx) {
                if (VIN) cyv_prestream_bc(c, x, y, lx, ly, g);
                else cy_prestream_bc(c, x, y, lx, ly, mu[e], g);
            }
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
