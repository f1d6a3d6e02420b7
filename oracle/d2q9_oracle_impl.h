/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the D2Q9 hot path.  Nothing in the
 * product path (2d-lb_b200/) may include, link or call this file.
 *
 * This header is a type-generic body: d2q9_oracle.c includes it twice, with
 *   REAL = float  (suffix _f32)   -- the reference's single-fluid OpenCL path
 *   REAL = double (suffix _f64)   -- the same scheme evaluated in double, the
 *                                    fp64 semantic SURVEY.md F5/A.4 describes.
 *
 * It restates, one function per reference kernel, what the reference's OpenCL
 * path computes ("scheme = opencl"):
 *   LB_D2Q9/D2Q9.cl                     (kernels, cited per function below)
 *   LB_D2Q9/dimensionless/opencl_dim.py (launch order, :372-387 and :510-518)
 * The arithmetic is mirrored operation by operation: same association order,
 * no fused multiply-add (compile with -ffp-contract=off), IEEE division, and
 * C "double literal" promotion exactly where the OpenCL C source has one
 * (`1./rho`, `2./3.`, `.5*`), so that a GPU kernel which mirrors the same
 * order must agree BIT FOR BIT.
 *
 * PINNED against the reference itself: the kernel files D2Q9.cl / D2Q9i.cl are
 * compiled as C where they lie (oracle/clshim/opencl_c.h, oracle/build_ref.py)
 * and run on the CPU under the reference's unmodified host classes
 * (oracle/shims/pyopencl, oracle/refload.py); every function below reproduces
 * them bit for bit -- golden vectors tests/golden/opencl_*.npz, oldcl_*.npz and
 * the kernel-by-kernel checks of tests/test_opencl_reference.py.
 *
 * Layout: f[9][ny][nx], x fastest (the reference's device layout,
 * D2Q9.cl:24-25 / opencl_dim.py:165), no row padding.
 */

#ifndef REAL
#error "include from d2q9_oracle.c only"
#endif

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SUFFIX)

/* -- a7: equilibrium, D2Q9.cl:2-64 ---------------------------------------- */
void FN(oracle_update_feq)(int nx, int ny, const REAL *rho, const REAL *u,
                           const REAL *v, REAL *feq, double cs2_d,
                           double cs22_d, double two_cs4_d)
{
    /* opencl_dim.py:305 passes np.float32(cs2), np.float32(cs22), np.float32(two_cs4) */
    const REAL cs2 = (REAL)cs2_d, two_cs2 = (REAL)cs22_d, two_cs4 = (REAL)two_cs4_d;
    const size_t plane = (size_t)nx * (size_t)ny;
    for (int j = 0; j < 9; ++j) {
        const REAL wj = (REAL)ORACLE_W[j];       /* w is a float32 array, opencl_dim.py:22 */
        const int ex = ORACLE_CX[j], ey = ORACLE_CY[j];
        for (size_t c = 0; c < plane; ++c) {
            const REAL uu = u[c], vv = v[c], r = rho[c];
            /* D2Q9.cl:55  cur_cx*u + cur_cy*v  (int -> REAL conversion, then mul, add) */
            const REAL cu = (REAL)ex * uu + (REAL)ey * vv;
            const REAL usq = uu * uu + vv * vv;                      /* :56 */
            /* :58  1.f + cu/cs2 + cu*cu/two_cs4 - usq/two_cs2, left to right */
            REAL inner = (REAL)1 + cu / cs2;
            inner = inner + (cu * cu) / two_cs4;
            inner = inner - usq / two_cs2;
            feq[(size_t)j * plane + c] = (wj * r) * inner;          /* :60 */
        }
    }
}

/* -- a6: moments, D2Q9.cl:67-100 ------------------------------------------ */
void FN(oracle_update_hydro)(int nx, int ny, const REAL *f, REAL *rho, REAL *u, REAL *v)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (size_t c = 0; c < plane; ++c) {
        REAL g[9];
        for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
        REAL r = g[0];
        for (int j = 1; j < 9; ++j) r = r + g[j];                   /* :92 */
        rho[c] = r;
        /* :94  `1./rho` -- the literal is double in OpenCL C, result stored to REAL */
        const REAL inv = (REAL)(1.0 / (double)r);
        u[c] = (((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv;   /* :96 */
        v[c] = (((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv;   /* :97 */
    }
}

/* -- a8: BGK relaxation, D2Q9.cl:102-121 ---------------------------------- */
void FN(oracle_collide)(int nx, int ny, REAL *f, const REAL *feq, double omega_d)
{
    const REAL omega = (REAL)omega_d;            /* np.float32(self.omega), opencl_dim.py:369 */
    const REAL keep = (REAL)1 - omega;           /* `1-omega`, :119 */
    const size_t n = (size_t)9 * nx * ny;
    for (size_t i = 0; i < n; ++i) f[i] = f[i] * keep + omega * feq[i];
}

/* -- a1+a2: push streaming into the second buffer, then copy back.
 *    D2Q9.cl:139-171 (`move`) followed by :123-137 (`copy_buffer`), as the host
 *    method opencl_dim.py:339-353 issues them.  Out-of-domain destinations are
 *    dropped; slots of f_streamed that receive nothing keep what they held. */
void FN(oracle_move)(int nx, int ny, REAL *f, REAL *f_streamed)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (int j = 0; j < 9; ++j) {
        const REAL *src = f + (size_t)j * plane;
        REAL *dst = f_streamed + (size_t)j * plane;
        for (int y = 0; y < ny; ++y) {
            const int ty = y + ORACLE_CY[j];
            if (ty < 0 || ty >= ny) continue;
            for (int x = 0; x < nx; ++x) {
                const int tx = x + ORACLE_CX[j];
                if (tx < 0 || tx >= nx) continue;
                dst[(size_t)ty * nx + tx] = src[(size_t)y * nx + x];
            }
        }
    }
    memcpy(f, f_streamed, sizeof(REAL) * 9 * plane);
}

/* -- a14: periodic push streaming, rocket_yeast.cl:152-191 / multi.cl:330-369 */
void FN(oracle_move_periodic)(int nx, int ny, REAL *f, REAL *f_streamed)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (int j = 0; j < 9; ++j) {
        const REAL *src = f + (size_t)j * plane;
        REAL *dst = f_streamed + (size_t)j * plane;
        for (int y = 0; y < ny; ++y) {
            int ty = y + ORACLE_CY[j];
            if (ty >= ny) ty -= ny;
            if (ty < 0) ty += ny;
            for (int x = 0; x < nx; ++x) {
                int tx = x + ORACLE_CX[j];
                if (tx >= nx) tx -= nx;
                if (tx < 0) tx += nx;
                dst[(size_t)ty * nx + tx] = src[(size_t)y * nx + x];
            }
        }
    }
    memcpy(f, f_streamed, sizeof(REAL) * 9 * plane);
}

/* -- a3: pressure inlet/outlet, walls, corners, in place.  D2Q9.cl:173-261.
 *    All right-hand sides use the values loaded before any store (:187-195). */
void FN(oracle_move_bcs)(int nx, int ny, REAL *f, double inlet_rho_d, double outlet_rho_d)
{
    const REAL rin = (REAL)inlet_rho_d, rout = (REAL)outlet_rho_d;  /* np.float32(...), opencl_dim.py:336 */
    const size_t plane = (size_t)nx * (size_t)ny;
#define F(j) f[(size_t)(j) * plane + c]
    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) {
            /* only boundary nodes are touched */
            if (x != 0 && x != nx - 1 && y != 0 && y != ny - 1) continue;
            const size_t c = (size_t)y * nx + x;
            REAL g[9];
            for (int j = 0; j < 9; ++j) g[j] = F(j);

            if (x == 0 && y >= 1 && y < ny - 1) {                    /* :198-203 inlet */
                const REAL s = ((((g[0] + g[2]) + (REAL)2 * g[3]) + g[4]) + (REAL)2 * g[6]) + (REAL)2 * g[7];
                const REAL ui = -((s - rin) / rin);
                F(1) = (REAL)((double)g[3] + ((2. / 3.) * (double)rin) * (double)ui);
                F(5) = (REAL)((((-.5 * (double)g[2]) + (.5 * (double)g[4])) + (double)g[7]) + ((1. / 6.) * (double)ui) * (double)rin);
                F(8) = (REAL)((((.5 * (double)g[2]) - (.5 * (double)g[4])) + (double)g[6]) + ((1. / 6.) * (double)ui) * (double)rin);
            }
            if (x == nx - 1 && y >= 1 && y < ny - 1) {               /* :205-210 outlet */
                const REAL s = ((((g[0] + (REAL)2 * g[1]) + g[2]) + g[4]) + (REAL)2 * g[5]) + (REAL)2 * g[8];
                const REAL uo = (REAL)(-1) + s / rout;
                F(3) = (REAL)((double)g[1] - ((2. / 3.) * (double)rout) * (double)uo);
                F(6) = (REAL)((((-.5 * (double)g[2]) + (.5 * (double)g[4])) + (double)g[8]) - ((1. / 6.) * (double)uo) * (double)rout);
                F(7) = (REAL)((((.5 * (double)g[2]) - (.5 * (double)g[4])) + (double)g[5]) - ((1. / 6.) * (double)uo) * (double)rout);
            }
            if (y == ny - 1 && x >= 1 && x < nx - 1) {               /* :213-217 north wall */
                F(4) = g[2];
                F(8) = (REAL)(.5 * (double)((-g[1] + g[3]) + (REAL)2 * g[6]));
                F(7) = (REAL)(.5 * (double)((g[1] - g[3]) + (REAL)2 * g[5]));
            }
            if (y == 0 && x >= 1 && x < nx - 1) {                    /* :219-223 south wall */
                F(2) = g[4];
                F(6) = (REAL)(.5 * (double)((g[1] - g[3]) + (REAL)2 * g[8]));
                F(5) = (REAL)(.5 * (double)((-g[1] + g[3]) + (REAL)2 * g[7]));
            }
            if (x == 0 && y == 0) {                                  /* :228-234 */
                const REAL t = (((-g[0] - (REAL)2 * g[3]) - (REAL)2 * g[4]) - (REAL)2 * g[7]) + rin;
                F(1) = g[3]; F(2) = g[4]; F(5) = g[7];
                F(6) = (REAL)(.5 * (double)t);
                F(8) = (REAL)(.5 * (double)t);
            }
            if (x == 0 && y == ny - 1) {                             /* :236-242 */
                const REAL t = (((-g[0] - (REAL)2 * g[2]) - (REAL)2 * g[3]) - (REAL)2 * g[6]) + rin;
                F(1) = g[3]; F(4) = g[2]; F(8) = g[6];
                F(5) = (REAL)(.5 * (double)t);
                F(7) = (REAL)(.5 * (double)t);
            }
            if (x == nx - 1 && y == 0) {                             /* :245-251 */
                const REAL t = (((-g[0] - (REAL)2 * g[1]) - (REAL)2 * g[4]) - (REAL)2 * g[8]) + rout;
                F(3) = g[1]; F(2) = g[4]; F(6) = g[8];
                F(5) = (REAL)(.5 * (double)t);
                F(7) = (REAL)(.5 * (double)t);
            }
            if (x == nx - 1 && y == ny - 1) {                        /* :253-259 */
                const REAL t = (((-g[0] - (REAL)2 * g[1]) - (REAL)2 * g[2]) - (REAL)2 * g[5]) + rout;
                F(3) = g[1]; F(4) = g[2]; F(7) = g[5];
                F(6) = (REAL)(.5 * (double)t);
                F(8) = (REAL)(.5 * (double)t);
            }
        }
    }
#undef F
}

/* -- a4: full bounce-back on solid nodes, D2Q9.cl:398-433 (mask == 1 only) -- */
void FN(oracle_bounceback)(int nx, int ny, const int32_t *mask, REAL *f)
{
    static const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const size_t plane = (size_t)nx * (size_t)ny;
    for (size_t c = 0; c < plane; ++c) {
        if (mask[c] != 1) continue;
        REAL g[9];
        for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
        for (int j = 1; j < 9; ++j) f[(size_t)j * plane + c] = g[opp[j]];
    }
}

/* -- a5: set_zero_velocity_in_obstacle, D2Q9.cl:377-396 -------------------- */
void FN(oracle_zero_velocity)(int nx, int ny, const int32_t *mask, REAL *u, REAL *v)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (size_t c = 0; c < plane; ++c)
        if (mask[c] == 1) { u[c] = (REAL)0; v[c] = (REAL)0; }
}

/* ======================================================================== */
/* Incompressible variant, LB_D2Q9/D2Q9i.cl.  Same kernels   */
/* except: equilibrium (:58-59), moments (:90-94), pressure inlet/outlet      */
/* (:195-205).  PINNED: bit-identical to the reference's own D2Q9i.cl compiled */
/* through oracle/clshim and driven by opencl_dim_D2Q9i.py                    */
/* (tests/golden/opencl_d2q9i_*.npz, tests/test_opencl_reference.py).         */
/* ======================================================================== */
void FN(oracle_update_feq_i)(int nx, int ny, const REAL *rho, const REAL *u, const REAL *v, REAL *feq)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (int j = 0; j < 9; ++j) {
        const REAL wj = (REAL)ORACLE_W[j];
        const int ex = ORACLE_CX[j], ey = ORACLE_CY[j];
        for (size_t c = 0; c < plane; ++c) {
            const REAL uu = u[c], vv = v[c], r = rho[c];
            const REAL cu = (REAL)ex * uu + (REAL)ey * vv;
            const REAL usq = uu * uu + vv * vv;
            /* D2Q9i.cl:58  rho + 3.*cu + (9./2.)*(cu*cu) - (3./2.)*usq : double literals */
            const REAL inner = (REAL)((((double)r + 3. * (double)cu) + (9. / 2.) * (double)(cu * cu)) - (3. / 2.) * (double)usq);
            feq[(size_t)j * plane + c] = (wj * r) * inner;             /* :59  cur_w*rho*inner_feq */
        }
    }
}

void FN(oracle_update_hydro_i)(int nx, int ny, const REAL *f, REAL *rho, REAL *u, REAL *v)
{
    const size_t plane = (size_t)nx * (size_t)ny;
    for (size_t c = 0; c < plane; ++c) {
        REAL g[9];
        for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
        REAL r = g[0];
        for (int j = 1; j < 9; ++j) r = r + g[j];
        rho[c] = r;                                                     /* D2Q9i.cl:92 */
        u[c] = ((((g[1] + g[5]) + g[8]) - g[6]) - g[3]) - g[7];         /* :93 */
        v[c] = ((((g[6] + g[2]) + g[5]) - g[7]) - g[4]) - g[8];         /* :94 */
    }
}

/* D2Q9i.cl:173-257: only the pressure inlet/outlet differ from D2Q9.cl; walls and corners are
 * delegated to the compressible routine afterwards (it does not touch inlet/outlet nodes twice
 * because the loaded values are re-read: so inlet/outlet are applied here on a copy of the loads). */
void FN(oracle_move_bcs_i)(int nx, int ny, REAL *f, double inlet_rho_d, double outlet_rho_d)
{
    const REAL rin = (REAL)inlet_rho_d, rout = (REAL)outlet_rho_d;
    const size_t plane = (size_t)nx * (size_t)ny;
#define F(j) f[(size_t)(j) * plane + c]
    /* walls and corners first via the shared routine, on a lattice whose inlet/outlet interior
     * nodes are restored afterwards -- simpler: handle inlet/outlet here and skip them there by
     * saving and restoring those columns. */
    REAL *save = (REAL *)malloc(sizeof(REAL) * 9 * 2 * (size_t)ny);
    for (int y = 0; y < ny; ++y)
        for (int j = 0; j < 9; ++j) {
            save[((size_t)j * 2 + 0) * ny + y] = f[(size_t)j * plane + (size_t)y * nx + 0];
            save[((size_t)j * 2 + 1) * ny + y] = f[(size_t)j * plane + (size_t)y * nx + (nx - 1)];
        }
    FN(oracle_move_bcs)(nx, ny, f, inlet_rho_d, outlet_rho_d);
    for (int y = 1; y < ny - 1; ++y) {
        REAL g[9];
        size_t c = (size_t)y * nx + 0;
        for (int j = 0; j < 9; ++j) { g[j] = save[((size_t)j * 2 + 0) * ny + y]; F(j) = g[j]; }
        {   /* :194-198 */
            const REAL ui = ((((((-g[0]) - g[2]) - (REAL)2 * g[3]) - g[4]) - (REAL)2 * g[6]) - (REAL)2 * g[7]) + rin;
            F(1) = (REAL)((1. / 3.) * (double)((REAL)3 * g[3] + (REAL)2 * ui));
            F(5) = (REAL)((1. / 6.) * (double)(((((REAL)(-3) * g[2]) + (REAL)3 * g[4]) + (REAL)6 * g[7]) + ui));
            F(8) = (REAL)((1. / 6.) * (double)((((REAL)3 * g[2] - (REAL)3 * g[4]) + (REAL)6 * g[6]) + ui));
        }
        c = (size_t)y * nx + (nx - 1);
        for (int j = 0; j < 9; ++j) { g[j] = save[((size_t)j * 2 + 1) * ny + y]; F(j) = g[j]; }
        {   /* :201-205 */
            const REAL uo = (((((g[0] + (REAL)2 * g[1]) + g[2]) + g[4]) + (REAL)2 * g[5]) + (REAL)2 * g[8]) - rout;
            F(3) = (REAL)((1. / 3.) * (double)((REAL)3 * g[1] - (REAL)2 * uo));
            F(6) = (REAL)((1. / 6.) * (double)(((((REAL)(-3) * g[2]) + (REAL)3 * g[4]) + (REAL)6 * g[8]) - uo));
            F(7) = (REAL)((1. / 6.) * (double)((((REAL)3 * g[2] - (REAL)3 * g[4]) + (REAL)6 * g[5]) - uo));
        }
    }
    free(save);
#undef F
}

/* -- a9: the step loop, opencl_dim.py:372-387 (+ :510-518 with a mask).
 *    bc: 0 = pipe (pressure inlet/outlet + walls), 1 = doubly periodic box.
 *    mask may be NULL.  f_streamed must start as a copy of f (opencl_dim.py:324-327).
 *    zero_obstacle_velocity: 0 = shipped opencl_dim behaviour (F10), 1 = zero u,v
 *    in the mask after every update_hydro (opencl_dim_D2Q9i / benchmarked revision).
 *    model: 0 = D2Q9.cl, 1 = incompressible D2Q9i.cl. */
void FN(oracle_run)(int nx, int ny, int bc, int n_steps, REAL *f, REAL *f_streamed,
                    const int32_t *mask, REAL *rho, REAL *u, REAL *v, REAL *feq,
                    double omega, double inlet_rho, double outlet_rho,
                    double cs2, double cs22, double two_cs4, int zero_obstacle_velocity, int model)
{
    for (int it = 0; it < n_steps; ++it) {
        if (bc == 1) {
            FN(oracle_move_periodic)(nx, ny, f, f_streamed);
        } else {
            FN(oracle_move)(nx, ny, f, f_streamed);
            if (model == 1) FN(oracle_move_bcs_i)(nx, ny, f, inlet_rho, outlet_rho);
            else FN(oracle_move_bcs)(nx, ny, f, inlet_rho, outlet_rho);
        }
        if (mask) FN(oracle_bounceback)(nx, ny, mask, f);
        if (model == 1) FN(oracle_update_hydro_i)(nx, ny, f, rho, u, v);
        else FN(oracle_update_hydro)(nx, ny, f, rho, u, v);
        if (mask && zero_obstacle_velocity) FN(oracle_zero_velocity)(nx, ny, mask, u, v);
        if (model == 1) FN(oracle_update_feq_i)(nx, ny, rho, u, v, feq);
        else FN(oracle_update_feq)(nx, ny, rho, u, v, feq, cs2, cs22, two_cs4);
        FN(oracle_collide)(nx, ny, f, feq, omega);
    }
}

/* == f-2, OpenCL flavour: the velocity-inlet / y-periodic family of LB_D2Q9/OLD/opencl.py ==========
 *    Only that module launches D2Q9.cl's *_PeriodicBC_VelocityInlet kernels, in the OLD step order
 *    (OLD/opencl.py:246-255):  move_bcs -> [bounce-back] -> move -> update_hydro -> [zero u,v in the
 *    obstacle] -> update_feq -> collide.  The boundary pass therefore acts on POST-COLLISION
 *    populations, and the slots `move` never writes (no upstream node) keep what f_streamed held when
 *    it was created -- the initial populations -- for ever. */

/* move_bcs_PeriodicBC_VelocityInlet, D2Q9.cl:263-321.  u_w, u_e arrive as np.float32
 * (OLD/opencl.py:293-294); `1.`, `2./3.`, `1./2.`, `1./6.` are double literals, `2*` (inlet, :292) is
 * an int and keeps that sum in REAL while `2.*` (outlet, :299) promotes it. */
void FN(oracle_move_bcs_vin)(int nx, int ny, REAL *f, double u_w_d, double u_e_d)
{
    const REAL u_w = (REAL)u_w_d, u_e = (REAL)u_e_d;
    const size_t plane = (size_t)nx * (size_t)ny;
    /* the two row copies read populations no work-item of this kernel writes (4,8,7 of row 0;
       2,6,5 of row ny-1), so the result does not depend on the execution order */
    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) {
            const size_t c = (size_t)y * nx + x;
#define F(j) f[(size_t)(j) * plane + c]
            const REAL f0 = F(0), f1 = F(1), f2 = F(2), f3 = F(3), f4 = F(4), f5 = F(5), f6 = F(6), f7 = F(7), f8 = F(8);
            if (x == 0 && y >= 1 && y < ny - 1) {
                const REAL rho_w = (REAL)((1. / (1. - (double)u_w)) * (double)(((f0 + f2) + f4) + (REAL)2 * ((f3 + f6) + f7)));
                F(1) = (REAL)((double)f3 + ((2. / 3.) * (double)rho_w) * (double)u_w);
                F(5) = (REAL)(((double)f7 - (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)u_w);
                F(8) = (REAL)(((double)f6 + (1. / 2.) * (double)(f2 - f4)) + ((1. / 6.) * (double)rho_w) * (double)u_w);
            }
            if (x == nx - 1 && y >= 1 && y < ny - 1) {
                const REAL rho_e = (REAL)((1. / (1. + (double)u_e)) * ((double)((f0 + f2) + f4) + 2. * (double)((f1 + f5) + f8)));
                F(3) = (REAL)((double)f1 - ((2. / 3.) * (double)rho_e) * (double)u_e);
                F(6) = (REAL)(((double)f5 + (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)u_e);
                F(7) = (REAL)(((double)f8 - (1. / 2.) * (double)(f2 - f4)) - ((1. / 6.) * (double)rho_e) * (double)u_e);
            }
            if (y == ny - 1) {                                   /* "NORTH": takes 4, 8, 7 of row 0 */
                F(4) = f[4 * plane + x];
                F(8) = f[8 * plane + x];
                F(7) = f[7 * plane + x];
            }
            if (y == 0) {                                        /* "SOUTH": takes 2, 6, 5 of row ny-1 */
                F(2) = f[2 * plane + (size_t)(ny - 1) * nx + x];
                F(6) = f[6 * plane + (size_t)(ny - 1) * nx + x];
                F(5) = f[5 * plane + (size_t)(ny - 1) * nx + x];
            }
#undef F
        }
    }
}

/* update_hydro_PeriodicBC_VelocityInlet, D2Q9.cl:323-374: u, v are written for interior columns only;
 * on x = 0 and x = nx-1 the rows 1..ny-2 get the boundary density and u_w / u_e, and everything else
 * there (v; u in the four corners) keeps whatever the arrays held. */
void FN(oracle_update_hydro_vin)(int nx, int ny, const REAL *f, REAL *rho, REAL *u, REAL *v,
                                 double u_w_d, double u_e_d)
{
    const REAL u_w = (REAL)u_w_d, u_e = (REAL)u_e_d;
    const size_t plane = (size_t)nx * (size_t)ny;
    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) {
            const size_t c = (size_t)y * nx + x;
            REAL g[9];
            for (int j = 0; j < 9; ++j) g[j] = f[(size_t)j * plane + c];
            REAL r = g[0];
            for (int j = 1; j < 9; ++j) r = r + g[j];
            rho[c] = r;
            const REAL inv = (REAL)(1.0 / (double)r);
            if (x != 0 && x != nx - 1) {
                u[c] = (((((g[1] - g[3]) + g[5]) - g[6]) - g[7]) + g[8]) * inv;
                v[c] = (((((g[5] + g[2]) + g[6]) - g[7]) - g[4]) - g[8]) * inv;
            }
            if (x == 0 && y != 0 && y < ny - 1) {
                rho[c] = (REAL)((1. / (1. - (double)u_w)) * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[3] + g[6]) + g[7])));
                u[c] = u_w;
            }
            if (x == nx - 1 && y != 0 && y < ny - 1) {
                rho[c] = (REAL)((1. / (1. + (double)u_e)) * ((double)((g[0] + g[2]) + g[4]) + 2. * (double)((g[1] + g[5]) + g[8])));
                u[c] = u_e;
            }
        }
    }
}

/* the step loop of OLD/opencl.py's velocity-inlet classes (:246-255, :290-327, :346-371);
 * with a mask, u and v are zeroed in the obstacle after every update_hydro (:359-363). */
void FN(oracle_run_oldcl_vin)(int nx, int ny, int n_steps, REAL *f, REAL *f_streamed,
                              const int32_t *mask, REAL *rho, REAL *u, REAL *v, REAL *feq,
                              double omega, double u_w, double u_e, double cs2, double cs22, double two_cs4)
{
    for (int it = 0; it < n_steps; ++it) {
        FN(oracle_move_bcs_vin)(nx, ny, f, u_w, u_e);
        if (mask) FN(oracle_bounceback)(nx, ny, mask, f);
        FN(oracle_move)(nx, ny, f, f_streamed);
        FN(oracle_update_hydro_vin)(nx, ny, f, rho, u, v, u_w, u_e);
        if (mask) FN(oracle_zero_velocity)(nx, ny, mask, u, v);
        FN(oracle_update_feq)(nx, ny, rho, u, v, feq, cs2, cs22, two_cs4);
        FN(oracle_collide)(nx, ny, f, feq, omega);
    }
}

#undef FN
#undef CAT
#undef CAT2
