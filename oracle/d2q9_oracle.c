/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the D2Q9 collide-and-stream path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the timed CPU arm -- never as part of the product path.
 *
 * Two restatements live here:
 *   1. scheme "opencl"  (d2q9_oracle_impl.h, instantiated for float and double):
 *      the path being replaced -- LB_D2Q9/D2Q9.cl driven by
 *      LB_D2Q9/dimensionless/opencl_dim.py:372-387.  pyopencl and an OpenCL device
 *      do not exist in this image, so the reference's kernel files are compiled as C
 *      (oracle/clshim) and run on the CPU under the reference's own host classes; this
 *      restatement is pinned to them BIT FOR BIT (tests/golden/opencl_*.npz, oldcl_*.npz,
 *      tests/test_opencl_reference.py), besides the Poiseuille known answer and the
 *      constructor printouts stored in docs/opencl_dimensionless_verification.ipynb.
 *   2. scheme "cython"  (below): LB_D2Q9/dimensionless/cython_dim.pyx:204-359
 *      and :459-513, i.e. the reference's own CPU path, including its mixed
 *      float32/float64 arithmetic as NumPy >= 2 evaluates it.  This one IS
 *      pinned directly: oracle/build_ref.py compiles the unmodified reference
 *      .pyx and tests compare field by field (golden vectors in tests/golden/).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* opencl_dim.py:22-25 / cython_dim.pyx:16-19 */
static const double ORACLE_W[9] = {4. / 9., 1. / 9., 1. / 9., 1. / 9., 1. / 9.,
                                   1. / 36., 1. / 36., 1. / 36., 1. / 36.};
static const int ORACLE_CX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int ORACLE_CY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

/* NB: in the float instantiation w is float32(w_j) exactly as the reference
 * stores it (np.array(..., dtype=np.float32)); (REAL)ORACLE_W[j] does that. */
#define REAL float
#define SUFFIX _f32
#include "d2q9_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX _f64
#include "d2q9_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* ======================================================================== */
/* scheme "cython": cython_dim.pyx, evaluated as NumPy 2.x (NEP 50) does:    */
/*   f, feq, rho are float32; u, v are float64; python-float scalars are     */
/*   "weak" (float32 array op python float -> float32), np.float64 scalars   */
/*   (omega, inlet_rho, cs2, cs22 -- all derived from cs=1/np.sqrt(3)) and   */
/*   float64 arrays promote the expression to float64; every store into a    */
/*   float32 array rounds once.                                              */
/* Layout here is f[9][ny][nx] (x fastest); the reference's (9,nx,ny) C-order */
/* array is the same data with x and y swapped in memory -- index mapping     */
/* only, no arithmetic difference.                                            */
/* ======================================================================== */

typedef struct {
    int nx, ny;
    double omega, inlet_rho, outlet_rho;
    double cs2, cs22, cssq;      /* cs**2, 2*cs2, 2.0/9.0  (cython_dim.pyx:20-23) */
    int velocity_inlet;          /* 1: OLD/cython.pyx Pipe_Flow_PeriodicBC_VelocityInlet (:268-360): imposed
                                    x-velocity u_w / u_e at inlet / outlet, rows y=0 and y=ly exchange their
                                    incoming populations ("periodic"), no wall or corner closure */
    double u_w, u_e;             /* Python floats in the reference */
    int old_api;                 /* 1: LB_D2Q9/OLD/cython.pyx flavour -- no wall zeroing in update_hydro
                                    (OLD/cython.pyx:126-149), and omega / inlet_rho are plain Python floats
                                    there ("weak" under NEP 50), so the inlet velocity and the relaxation
                                    are evaluated in float32 instead of float64 */
} cy_params;

#define IDX(j, x, y) ((size_t)(j) * plane + (size_t)(y) * nx + (size_t)(x))

/* cython_dim.pyx:204-269 -- BCs applied BEFORE streaming, inlet/outlet use the
 * velocity stored by the previous update_hydro. */
void oracle_cy_move_bcs(const cy_params *p, float *f, const double *u)
{
    const int nx = p->nx, ny = p->ny, lx = nx - 1, ly = ny - 1;
    const size_t plane = (size_t)nx * ny;
    const double rin = p->inlet_rho, rout = p->outlet_rho;
    /* :217-219 inlet, y in [1, ly) ; each line reads the array as it then stands
     * (f1,f5,f8 are written, f2,f3,f4,f6,f7 only read, so order does not matter) */
    for (int y = 1; y < ly; ++y) {
        const double U = u[(size_t)y * nx + 0];
        const float f2 = f[IDX(2, 0, y)], f3 = f[IDX(3, 0, y)], f4 = f[IDX(4, 0, y)];
        const float f6 = f[IDX(6, 0, y)], f7 = f[IDX(7, 0, y)];
        f[IDX(1, 0, y)] = (float)((double)f3 + ((2. / 3.) * rin) * U);
        f[IDX(5, 0, y)] = (float)((double)(((-.5f * f2) + (.5f * f4)) + f7) + ((1. / 6.) * U) * rin);
        f[IDX(8, 0, y)] = (float)((double)(((.5f * f2) - (.5f * f4)) + f6) + ((1. / 6.) * U) * rin);
    }
    /* :222-224 outlet */
    for (int y = 1; y < ly; ++y) {
        const double U = u[(size_t)y * nx + lx];
        const float f1 = f[IDX(1, lx, y)], f2 = f[IDX(2, lx, y)], f4 = f[IDX(4, lx, y)];
        const float f5 = f[IDX(5, lx, y)], f8 = f[IDX(8, lx, y)];
        f[IDX(3, lx, y)] = (float)((double)f1 - ((2. / 3.) * rout) * U);
        f[IDX(6, lx, y)] = (float)((double)(((-.5f * f2) + (.5f * f4)) + f8) - ((1. / 6.) * U) * rout);
        f[IDX(7, lx, y)] = (float)((double)(((.5f * f2) - (.5f * f4)) + f5) - ((1. / 6.) * U) * rout);
    }
    /* :226-228  cdef float inlet_rho / outlet_rho */
    const float rin_f = (float)rin, rout_f = (float)rout;
    /* :232-240 walls are plain reflections for x in [1, lx) */
    for (int x = 1; x < lx; ++x) {
        f[IDX(4, x, ly)] = f[IDX(2, x, ly)];
        f[IDX(8, x, ly)] = f[IDX(6, x, ly)];
        f[IDX(7, x, ly)] = f[IDX(5, x, ly)];
    }
    for (int x = 1; x < lx; ++x) {
        f[IDX(2, x, 0)] = f[IDX(4, x, 0)];
        f[IDX(6, x, 0)] = f[IDX(8, x, 0)];
        f[IDX(5, x, 0)] = f[IDX(7, x, 0)];
    }
    /* :244-269 corners.  Cython emits the literal `2` as the C double `2.0`, so the generated C
     * evaluates the whole parenthesis in double and rounds once on the store. */
    {
        const double t = ((((double)(-f[IDX(0, 0, 0)]) - 2.0 * (double)f[IDX(3, 0, 0)]) - 2.0 * (double)f[IDX(4, 0, 0)]) - 2.0 * (double)f[IDX(7, 0, 0)]) + (double)rin_f;
        f[IDX(1, 0, 0)] = f[IDX(3, 0, 0)];
        f[IDX(2, 0, 0)] = f[IDX(4, 0, 0)];
        f[IDX(5, 0, 0)] = f[IDX(7, 0, 0)];
        f[IDX(6, 0, 0)] = (float)(.5 * t);
        f[IDX(8, 0, 0)] = (float)(.5 * t);
    }
    {
        const double t = ((((double)(-f[IDX(0, 0, ly)]) - 2.0 * (double)f[IDX(2, 0, ly)]) - 2.0 * (double)f[IDX(3, 0, ly)]) - 2.0 * (double)f[IDX(6, 0, ly)]) + (double)rin_f;
        f[IDX(1, 0, ly)] = f[IDX(3, 0, ly)];
        f[IDX(4, 0, ly)] = f[IDX(2, 0, ly)];
        f[IDX(5, 0, ly)] = (float)(.5 * t);
        f[IDX(7, 0, ly)] = (float)(.5 * t);
        f[IDX(8, 0, ly)] = f[IDX(6, 0, ly)];
    }
    {
        const double t = ((((double)(-f[IDX(0, lx, 0)]) - 2.0 * (double)f[IDX(1, lx, 0)]) - 2.0 * (double)f[IDX(4, lx, 0)]) - 2.0 * (double)f[IDX(8, lx, 0)]) + (double)rout_f;
        f[IDX(3, lx, 0)] = f[IDX(1, lx, 0)];
        f[IDX(2, lx, 0)] = f[IDX(4, lx, 0)];
        f[IDX(6, lx, 0)] = f[IDX(8, lx, 0)];
        f[IDX(5, lx, 0)] = (float)(.5 * t);
        f[IDX(7, lx, 0)] = (float)(.5 * t);
    }
    {
        const double t = ((((double)(-f[IDX(0, lx, ly)]) - 2.0 * (double)f[IDX(1, lx, ly)]) - 2.0 * (double)f[IDX(2, lx, ly)]) - 2.0 * (double)f[IDX(5, lx, ly)]) + (double)rout_f;
        f[IDX(3, lx, ly)] = f[IDX(1, lx, ly)];
        f[IDX(4, lx, ly)] = f[IDX(2, lx, ly)];
        f[IDX(6, lx, ly)] = (float)(.5 * t);
        f[IDX(7, lx, ly)] = f[IDX(5, lx, ly)];
        f[IDX(8, lx, ly)] = (float)(.5 * t);
    }
}

/* cython_dim.pyx:486-513 -- swap opposite pairs on the obstacle pixel list */
void oracle_cy_bounceback(const cy_params *p, const uint8_t *mask, float *f)
{
    const int nx = p->nx, ny = p->ny;
    const size_t plane = (size_t)nx * ny;
    for (size_t c = 0; c < plane; ++c) {
        if (!mask[c]) continue;
        float t;
        t = f[1 * plane + c]; f[1 * plane + c] = f[3 * plane + c]; f[3 * plane + c] = t;
        t = f[2 * plane + c]; f[2 * plane + c] = f[4 * plane + c]; f[4 * plane + c] = t;
        t = f[5 * plane + c]; f[5 * plane + c] = f[7 * plane + c]; f[7 * plane + c] = t;
        t = f[6 * plane + c]; f[6 * plane + c] = f[8 * plane + c]; f[8 * plane + c] = t;
    }
}

/* cython_dim.pyx:284-299 -- in-place streaming.  Written here in PULL form:
 * every destination takes its upstream value from a snapshot of the pre-stream
 * array, except the entries the reference's loop bounds never write (they keep
 * their pre-stream value): out-of-domain sources, and the four "non-streaming
 * lines" f2 @ x=lx, f1 @ y=0, f4 @ x=0, f3 @ y=ly.  The
 * reference's sweep order makes its in-place update equal to this snapshot
 * form; tests/test_oracle.py proves it against the compiled reference. */
void oracle_cy_move(const cy_params *p, float *f, float *scratch)
{
    const int nx = p->nx, ny = p->ny, lx = nx - 1, ly = ny - 1;
    const size_t plane = (size_t)nx * ny;
    memcpy(scratch, f, sizeof(float) * 9 * plane);
#define S(j, x, y) scratch[IDX(j, x, y)]
    for (int y = 1; y <= ly; ++y)              /* :285-288  j in [ly..1], i in [0,lx) */
        for (int x = 0; x < lx; ++x) {
            f[IDX(2, x, y)] = S(2, x, y - 1);
            f[IDX(6, x, y)] = S(6, x + 1, y - 1);
        }
    for (int y = 1; y <= ly; ++y)              /* :289-292  j in [ly..1], i in [lx..1] */
        for (int x = 1; x <= lx; ++x) {
            f[IDX(1, x, y)] = S(1, x - 1, y);
            f[IDX(5, x, y)] = S(5, x - 1, y - 1);
        }
    for (int y = 0; y < ly; ++y)               /* :293-296  j in [0,ly), i in [lx..1] */
        for (int x = 1; x <= lx; ++x) {
            f[IDX(4, x, y)] = S(4, x, y + 1);
            f[IDX(8, x, y)] = S(8, x - 1, y + 1);
        }
    for (int y = 0; y < ly; ++y)               /* :297-299  j in [0,ly), i in [0,lx) */
        for (int x = 0; x < lx; ++x) {
            f[IDX(3, x, y)] = S(3, x + 1, y);
            f[IDX(7, x, y)] = S(7, x + 1, y + 1);
        }
#undef S
}

/* cython_dim.pyx:302-333 (+ :459-466 when a mask is given) */
void oracle_cy_update_hydro(const cy_params *p, const float *f, float *rho, double *u,
                            double *v, const uint8_t *mask)
{
    const int nx = p->nx, ny = p->ny, lx = nx - 1, ly = ny - 1;
    const size_t plane = (size_t)nx * ny;
    for (size_t c = 0; c < plane; ++c) {
        float g[9];
        for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
        float r = g[0];
        for (int j = 1; j < 9; ++j) r = r + g[j];              /* np.sum(f, axis=0), float32 */
        rho[c] = r;
        const float inv = 1.0f / r;                            /* 1./rho, float32 */
        u[c] = (double)((((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv);
        v[c] = (double)((((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv);
    }
    if (!p->old_api)
        for (int x = 0; x < nx; ++x) {                         /* :317-320 */
            u[(size_t)0 * nx + x] = 0; u[(size_t)ly * nx + x] = 0;
            v[(size_t)0 * nx + x] = 0; v[(size_t)ly * nx + x] = 0;
        }
    for (int y = 0; y < ny; ++y) {                             /* :325-333 */
        rho[(size_t)y * nx + 0] = (float)p->inlet_rho;
        rho[(size_t)y * nx + lx] = (float)p->outlet_rho;
        {
            const float a = (f[IDX(0, 0, y)] + f[IDX(2, 0, y)]) + f[IDX(4, 0, y)];
            const float b = 2 * ((f[IDX(3, 0, y)] + f[IDX(6, 0, y)]) + f[IDX(7, 0, y)]);
            /* cython_dim: inlet_rho = 1. + np.abs(...) is np.float64 -> float64 arithmetic;
             * OLD: inlet_rho is the Python float 1. -> float32 arithmetic */
            if (p->old_api) u[(size_t)y * nx + 0] = (double)(1.0f - (a + b) / (float)p->inlet_rho);
            else u[(size_t)y * nx + 0] = 1 - (double)(a + b) / p->inlet_rho;
        }
        {
            const float a = (f[IDX(0, lx, y)] + f[IDX(2, lx, y)]) + f[IDX(4, lx, y)];
            const float b = 2 * ((f[IDX(1, lx, y)] + f[IDX(5, lx, y)]) + f[IDX(8, lx, y)]);
            /* cython_dim: outlet_rho is the Python float 1. ("weak"), so this line is float32 arithmetic;
             * OLD: outlet_rho = deltaP/cs2 + 1 is np.float64, float64 arithmetic */
            if (p->old_api) u[(size_t)y * nx + lx] = -1 + (double)(a + b) / p->outlet_rho;
            else u[(size_t)y * nx + lx] = (double)(-1.0f + (a + b) / (float)p->outlet_rho);
        }
    }
    if (mask)
        for (size_t c = 0; c < plane; ++c)
            if (mask[c]) { u[c] = 0; v[c] = 0; }
}

/* cython_dim.pyx:160-189 -- Succi's factored equilibrium, float64 expression
 * (u, v are float64 arrays), `w*rho` formed in float32 first. */
void oracle_cy_update_feq(const cy_params *p, const float *rho, const double *u,
                          const double *v, float *feq)
{
    const int nx = p->nx, ny = p->ny;
    const size_t plane = (size_t)nx * ny;
    const double cs2 = p->cs2, cs22 = p->cs22, cssq = p->cssq;
    const float w0 = (float)(4. / 9.), w1 = (float)(1. / 9.), w2 = (float)(1. / 36.);
    for (size_t c = 0; c < plane; ++c) {
        const double uu = u[c], vv = v[c];
        const double ul = uu / cs2, vl = vv / cs2, uv = ul * vl;
        const double usq = uu * uu, vsq = vv * vv;
        const double sumsq = (usq + vsq) / cs22;
        const double sumsq2 = (sumsq * (1. - cs2)) / cs2;
        const double u2 = usq / cssq, v2 = vsq / cssq;
        const double r0 = (double)(w0 * rho[c]), r1 = (double)(w1 * rho[c]), r2 = (double)(w2 * rho[c]);
        feq[0 * plane + c] = (float)(r0 * (1. - sumsq));
        feq[1 * plane + c] = (float)(r1 * (((1. - sumsq) + u2) + ul));
        feq[2 * plane + c] = (float)(r1 * (((1. - sumsq) + v2) + vl));
        feq[3 * plane + c] = (float)(r1 * (((1. - sumsq) + u2) - ul));
        feq[4 * plane + c] = (float)(r1 * (((1. - sumsq) + v2) - vl));
        feq[5 * plane + c] = (float)(r2 * ((((1. + sumsq2) + ul) + vl) + uv));
        feq[6 * plane + c] = (float)(r2 * ((((1. + sumsq2) - ul) + vl) - uv));
        feq[7 * plane + c] = (float)(r2 * ((((1. + sumsq2) - ul) - vl) + uv));
        feq[8 * plane + c] = (float)(r2 * ((((1. + sumsq2) + ul) - vl) - uv));
    }
}

/* cython_dim.pyx:336-344 -- omega is np.float64, so the expression is float64 */
void oracle_cy_collide(const cy_params *p, float *f, const float *feq)
{
    const size_t n = (size_t)9 * p->nx * p->ny;
    const double keep = 1. - p->omega, om = p->omega;
    if (p->old_api) {
        const float keep_f = (float)keep, om_f = (float)om;
        for (size_t i = 0; i < n; ++i) f[i] = f[i] * keep_f + om_f * feq[i];
        return;
    }
    for (size_t i = 0; i < n; ++i) f[i] = (float)((double)f[i] * keep + om * (double)feq[i]);
}

/* OLD/cython.pyx:278-316 -- Pipe_Flow_PeriodicBC_VelocityInlet.move_bcs.  Typed-memoryview code:
 * C arithmetic, double literals, `cdef float u_w, u_e`, rho_w / rho_e stored as float. */
void oracle_cyv_move_bcs(const cy_params *p, float *f)
{
    const int nx = p->nx, ny = p->ny, lx = nx - 1, ly = ny - 1;
    const size_t plane = (size_t)nx * ny;
    const float u_w = (float)p->u_w, u_e = (float)p->u_e;
    for (int y = 1; y < ly; ++y) {
        {
            const float f0 = f[IDX(0, 0, y)], f2 = f[IDX(2, 0, y)], f3 = f[IDX(3, 0, y)], f4 = f[IDX(4, 0, y)];
            const float f6 = f[IDX(6, 0, y)], f7 = f[IDX(7, 0, y)];
            const float rho_w = (float)((1. / (1. - (double)u_w)) * ((double)((f0 + f2) + f4) + 2. * (double)((f3 + f6) + f7)));
            f[IDX(1, 0, y)] = (float)((double)f3 + ((2. / 3.) * (double)rho_w) * (double)u_w);
            f[IDX(5, 0, y)] = (float)(((double)f7 - (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)u_w);
            f[IDX(8, 0, y)] = (float)(((double)f6 + (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)u_w);
        }
        {
            const float f0 = f[IDX(0, lx, y)], f1 = f[IDX(1, lx, y)], f2 = f[IDX(2, lx, y)], f4 = f[IDX(4, lx, y)];
            const float f5 = f[IDX(5, lx, y)], f8 = f[IDX(8, lx, y)];
            const float rho_e = (float)((1. / (1. + (double)u_e)) * ((double)((f0 + f2) + f4) + 2. * (double)((f1 + f5) + f8)));
            f[IDX(3, lx, y)] = (float)((double)f1 - ((2. / 3.) * (double)rho_e) * (double)u_e);
            f[IDX(7, lx, y)] = (float)(((double)f5 + (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)u_e);
            f[IDX(6, lx, y)] = (float)(((double)f8 - (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)u_e);
        }
    }
    for (int x = 0; x <= lx; ++x) {            /* :305-316, north then south, in this order per column */
        f[IDX(4, x, ly)] = f[IDX(4, x, 0)];
        f[IDX(8, x, ly)] = f[IDX(8, x, 0)];
        f[IDX(7, x, ly)] = f[IDX(7, x, 0)];
        f[IDX(2, x, 0)] = f[IDX(2, x, ly)];
        f[IDX(6, x, 0)] = f[IDX(6, x, ly)];
        f[IDX(5, x, 0)] = f[IDX(5, x, ly)];
    }
}

/* OLD/cython.pyx:331-360 -- update_hydro of the velocity-inlet class (NumPy: u_w is a Python float,
 * so the rho expressions are float32) (+ :375-378 mask zeroing in the Obstacles variant) */
void oracle_cyv_update_hydro(const cy_params *p, const float *f, float *rho, double *u, double *v,
                             const uint8_t *mask)
{
    const int nx = p->nx, ny = p->ny, lx = nx - 1, ly = ny - 1;
    const size_t plane = (size_t)nx * ny;
    for (size_t c = 0; c < plane; ++c) {
        float g[9];
        for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
        float r = g[0];
        for (int j = 1; j < 9; ++j) r = r + g[j];
        rho[c] = r;
        const float inv = 1.0f / r;
        u[c] = (double)((((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv);
        v[c] = (double)((((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv);
    }
    const float kw = (float)(1. / (1. - p->u_w)), ke = (float)(1. / (1. + p->u_e));
    for (int y = 1; y < ly; ++y) {
        u[(size_t)y * nx + 0] = p->u_w;
        rho[(size_t)y * nx + 0] = kw * (((f[IDX(0, 0, y)] + f[IDX(2, 0, y)]) + f[IDX(4, 0, y)]) +
                                        2.0f * ((f[IDX(3, 0, y)] + f[IDX(6, 0, y)]) + f[IDX(7, 0, y)]));
        u[(size_t)y * nx + lx] = p->u_e;
        rho[(size_t)y * nx + lx] = ke * (((f[IDX(0, lx, y)] + f[IDX(2, lx, y)]) + f[IDX(4, lx, y)]) +
                                         2.0f * ((f[IDX(1, lx, y)] + f[IDX(5, lx, y)]) + f[IDX(8, lx, y)]));
    }
    if (mask)
        for (size_t c = 0; c < plane; ++c)
            if (mask[c]) { u[c] = 0; v[c] = 0; }
}

/* cython_dim.pyx:346-359 (+ :468-513).  scratch: 9*nx*ny floats. */
void oracle_cy_run(const cy_params *p, int n_steps, float *f, float *feq, float *rho,
                   double *u, double *v, const uint8_t *mask, float *scratch)
{
    for (int it = 0; it < n_steps; ++it) {
        if (p->velocity_inlet) oracle_cyv_move_bcs(p, f);
        else oracle_cy_move_bcs(p, f, u);
        if (mask) oracle_cy_bounceback(p, mask, f);
        oracle_cy_move(p, f, scratch);
        if (p->velocity_inlet) oracle_cyv_update_hydro(p, f, rho, u, v, mask);
        else oracle_cy_update_hydro(p, f, rho, u, v, mask);
        oracle_cy_update_feq(p, rho, u, v, feq);
        oracle_cy_collide(p, f, feq);
    }
}
