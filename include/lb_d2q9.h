/*
 * lb_d2q9.h -- C ABI of the B200-native D2Q9 collide-and-stream engine.
 *
 * This is the drop-in boundary.  The reference
 * (latticeboltzmann/2d-lb) has no FFI of its own: its seam is the pyopencl
 * layer underneath LB_D2Q9/dimensionless/opencl_dim.py.  Each entry point
 * below names the reference call site(s) it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative lb_status otherwise, and
 *     never throws; lb_last_error() gives the message.
 *   - one host thread per handle; the handle owns its device memory, stream,
 *     CUDA graphs and peer mappings.
 *   - host arrays use the reference's device layout: f[9][ny][nx], x fastest,
 *     no padding (D2Q9.cl:24-25; == the bytes of opencl_dim's Fortran-order
 *     (nx,ny,9) host arrays, opencl_dim.py:165).  Element type = the handle's
 *     dtype (float for LB_F32, double for LB_F64).
 *   - there is NO CPU fallback: without a CUDA device lb_create fails.
 */
#ifndef LB_D2Q9_H
#define LB_D2Q9_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_ABI_VERSION 5

typedef struct lb_sim lb_sim; /* opaque */

enum lb_status {
    LB_OK = 0,
    LB_ERR_INVALID = -1,   /* bad argument / configuration */
    LB_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    LB_ERR_STATE = -3,     /* call not valid in the handle's current state */
    LB_ERR_HALO = -4       /* halo hand-shake with a neighbour slab timed out */
};

enum lb_dtype { LB_F32 = 0, LB_F64 = 1 };

/* Boundary family.
 *   LB_BC_PIPE      pressure (Zou-He) inlet x=0 / outlet x=nx-1, walls y=0 / y=ny-1,
 *                   four corner closures -- D2Q9.cl:173-261 (`move_bcs`).
 *   LB_BC_PERIODIC  doubly periodic box -- rocket_yeast.cl:152-191 /
 *                   multi.cl:330-369 (`move_periodic`).
 *   LB_BC_VELOCITY_YPERIODIC  see below. */
enum lb_bc {
    LB_BC_PIPE = 0,
    LB_BC_PERIODIC = 1,
    /* imposed x-velocity u_west / u_east at inlet / outlet (Zou-He), rows y=0 and y=ny-1 exchange
     * their incoming populations -- LB_D2Q9/OLD/cython.pyx:268-360 (Pipe_Flow_PeriodicBC_VelocityInlet,
     * the CPU twin of D2Q9.cl:263-374) with LB_SCHEME_CYTHON_OLD; the OpenCL kernels themselves
     * (D2Q9.cl:263-374 as LB_D2Q9/OLD/opencl.py:281-371 drives them) with LB_SCHEME_OPENCL_OLD. */
    LB_BC_VELOCITY_YPERIODIC = 2
};

/* Arithmetic contract of the fused kernel.
 *   LB_MATH_STRICT  mirrors D2Q9.cl operation by operation (association order,
 *                   IEEE division, no FMA contraction): bit-identical to the
 *                   CPU oracle.
 *   LB_MATH_FAST    same formulas with reciprocal constants and FMA.  Agrees with
 *                   STRICT to rounding: rho and u within 1e-5 (relative) for runs of
 *                   up to 500 steps in fp32, 1e-12 in fp64 (2000 steps).  Longer fp32
 *                   runs accumulate round-off like a random walk in BOTH arithmetics,
 *                   which then drift apart (rho 1.5e-5 .. 2e-5, u 5e-5 .. 1.4e-4 of
 *                   max|u| at 2000 steps) while each stays about as close to the fp64 solution
 *                   (tests/test_parity_gpu.py::test_fast_math_error_growth_...). */
enum lb_math { LB_MATH_STRICT = 0, LB_MATH_FAST = 1 };

/* Which of the reference's two step algorithms the handle runs.
 *   LB_SCHEME_OPENCL      stream -> BCs -> moments -> feq -> collide, D2Q9.cl driven by
 *                         dimensionless/opencl_dim.py:372-387 (the path being replaced; default).
 *   LB_SCHEME_CYTHON      BCs (from the previous step's u) -> stream -> moments (+boundary
 *                         overrides) -> feq -> collide, dimensionless/cython_dim.pyx:346-359, with
 *                         NumPy's mixed float32/float64 arithmetic.  Bit-identical to the compiled
 *                         reference.  dtype must be LB_F32, bc LB_BC_PIPE, single slab; the u and v
 *                         fields are float64 (as in the reference), rho/f/feq float32.
 *   LB_SCHEME_CYTHON_OLD  the same for LB_D2Q9/OLD/cython.pyx (no wall zeroing of u,v; omega and
 *                         inlet_rho are Python floats there, so those expressions are float32).
 *   LB_SCHEME_OPENCL_OLD  D2Q9.cl's kernels in the step order of LB_D2Q9/OLD/opencl.py:246-255
 *                         (BCs -> bounce-back -> stream -> moments -> feq -> collide), for its
 *                         velocity-inlet / y-periodic classes: bc must be LB_BC_VELOCITY_YPERIODIC,
 *                         dtype LB_F32, single slab.  Populations without an upstream node keep the
 *                         values of the last lb_upload_f (the reference's `move` never writes them);
 *                         u, v, rho are float32; lb_upload_moments supplies the velocity entries the
 *                         reference's update_hydro never rewrites (D2Q9.cl:357-371).  Bit-identical
 *                         to the reference's own kernels (tests/golden/oldcl_*.npz). */
enum lb_scheme { LB_SCHEME_OPENCL = 0, LB_SCHEME_CYTHON = 1, LB_SCHEME_CYTHON_OLD = 2, LB_SCHEME_OPENCL_OLD = 3 };

/* Collision model.  LB_MODEL_D2Q9I is the reference's incompressible variant (LB_D2Q9/D2Q9i.cl behind
 * dimensionless/opencl_dim_D2Q9i.py): u, v are the raw momentum (no division by rho, D2Q9i.cl:92-94),
 * feq = w*rho*(rho + 3 c.u + 4.5 (c.u)^2 - 1.5 u^2) (:58-59), and the pressure inlet/outlet closure
 * of :194-205.  Everything else is unchanged. */
enum lb_model { LB_MODEL_D2Q9 = 0, LB_MODEL_D2Q9I = 1 };

enum lb_field {
    LB_FIELD_F = 0,    /* [9][ny][nx] post-collision populations (opencl_dim.py:394-395) */
    LB_FIELD_FEQ = 1,  /* [9][ny][nx] equilibrium of the last moments (opencl_dim.py:397-398) */
    LB_FIELD_RHO = 2,  /* [ny][nx] */
    LB_FIELD_U = 3,    /* [ny][nx] */
    LB_FIELD_V = 4     /* [ny][nx] */
};

enum lb_side { LB_WEST = 0, LB_EAST = 1 };

/* What lies beyond the slab's first / last column. */
enum lb_edge {
    LB_EDGE_BOUNDARY = 0, /* the domain boundary (inlet / outlet for LB_BC_PIPE)      */
    LB_EDGE_WRAP = 1,     /* periodic wrap inside this slab (single-slab periodic box) */
    LB_EDGE_HALO = 2      /* a neighbour slab: ghost column filled through lb_halo_*    */
};

typedef struct lb_config {
    int32_t struct_size;   /* = sizeof(lb_config), ABI guard */
    int32_t device;        /* CUDA device ordinal */
    int32_t nx, ny;        /* extent of THIS slab */
    int32_t dtype;         /* lb_dtype */
    int32_t bc;            /* lb_bc */
    int32_t math;          /* lb_math */
    int32_t zero_obstacle_velocity; /* 1: u=v=0 on solid nodes after every moment update
                                       (opencl_dim_D2Q9i.py / cython_dim.pyx:459-466); 0: shipped
                                       opencl_dim behaviour (moments of solid nodes kept) */
    /* x-slab decomposition (single slab: global_nx = nx, x_offset = 0, edges per bc) */
    int32_t global_nx;     /* columns of the whole lattice */
    int32_t x_offset;      /* global x of local column 0 */
    int32_t west_edge;     /* lb_edge */
    int32_t east_edge;     /* lb_edge */
    int32_t scheme;        /* lb_scheme */
    int32_t model;         /* lb_model (LB_SCHEME_OPENCL only) */
    /* physics: opencl_dim.py:118 (omega), :273-274 (inlet/outlet rho), :26-30 (lattice constants,
       passed in so that they are the very doubles the host computed) */
    double omega, inlet_rho, outlet_rho;
    double cs2, cs22, two_cs4;
    double u_west, u_east; /* LB_BC_VELOCITY_YPERIODIC: imposed inlet / outlet x-velocity */
    void *stream;          /* cudaStream_t to enqueue on; NULL = the handle creates its own */
} lb_config;

/* -- lifetime: replaces cl.Context/CommandQueue/Program.build and the cl.Buffer
 *    allocations of opencl_dim.py:203-255,165-176 ------------------------------ */
int lb_abi_version(void);
int lb_device_count(void);
int lb_create(const lb_config *cfg, lb_sim **out);
int lb_destroy(lb_sim *sim);
/* message of the last failure on this handle (or of the last failed lb_create when sim == NULL) */
const char *lb_last_error(const lb_sim *sim);

/* -- uploads ------------------------------------------------------------------
 * lb_set_mask: cl.Buffer(..., hostbuf=obstacle_mask_host), opencl_dim.py:502-503.
 *   host_mask[ny][nx], elem_bytes 1 (uint8) or 4 (int32); value 1 = solid
 *   (D2Q9.cl:410 tests `== 1`).  NULL removes the mask.
 * lb_upload_f: init_pop's two cl.Buffer(COPY_HOST_PTR) of f and f_streamed, opencl_dim.py:324-327.
 * lb_upload_moments: init_hydro's rho/u/v buffers, opencl_dim.py:291-293.  Any pointer may be NULL. */
int lb_set_mask(lb_sim *sim, const void *host_mask, int elem_bytes);
int lb_upload_f(lb_sim *sim, const void *host_f);
int lb_upload_moments(lb_sim *sim, const void *host_rho, const void *host_u, const void *host_v);

/* -- the hot path: Pipe_Flow.run, opencl_dim.py:372-387 (+ :510-518 with a mask).
 *    Enqueues n_steps fused stream+BC+bounce-back+moments+feq+collide launches
 *    (CUDA-graph batched); no host synchronisation inside.  After the call
 *    completes, rho/u/v hold the moments of the last step (pre-collision), f is
 *    post-collision -- the reference's get_fields semantics. */
int lb_step(lb_sim *sim, int n_steps);
int lb_sync(lb_sim *sim);

/* -- the reference's whole user sequence in one call: init_pop's upload of f (opencl_dim.py:324-327),
 *    run(n_steps) (:372-387) and get_fields' read-back of rho, u, v (:400-407; any output may be NULL), with the
 *    three stages PIPELINED by row bands: a band is updated through all n_steps as soon as it has arrived (a
 *    skewed wavefront of row-range launches) and its moments travel back while later bands are still being
 *    uploaded, so H2D, compute and D2H overlap and the call lasts about as long as the upload alone.  Host
 *    buffers should be page-locked (cudaHostAlloc / torch pin_memory) for the copies to overlap.  Blocking.
 *    Bit-identical to lb_upload_f + lb_step + lb_download, to which it falls back where it cannot pipeline
 *    (halo-connected slabs, periodic boxes, the other schemes, n_steps > 256, ny < 1024). */
int lb_run_streamed(lb_sim *sim, const void *host_f, int n_steps, void *host_rho, void *host_u, void *host_v);

/* -- readback: cl.enqueue_copy(queue, host, dev, is_blocking=True), opencl_dim.py:394-407.
 *    Blocking.  host_out has the layout stated at the top of this file. */
int lb_download(lb_sim *sim, int field, void *host_out);

/* Down-sampled readback for long visual runs (docs/cs205_movie.ipynb cells 17-23 pull the full
 * field every frame): every stride_x-th column and stride_y-th row of a 2-D field (rho, u, v) is
 * gathered on the device and only ceil(ny/stride_y) x ceil(nx/stride_x) values cross PCIe.  Blocking. */
int lb_download_strided(lb_sim *sim, int field, int stride_x, int stride_y, void *host_out);

/* -- single stages, for the kernel-by-kernel checks the reference's notebooks do
 *    (testing/Bryan/opencl_check_03.ipynb).  Each is one non-fused launch operating on the
 *    handle's current f / rho,u,v / feq and is equivalent to the reference kernel named:
 *    move            = D2Q9.cl `move` + `copy_buffer`        (opencl_dim.py:339-353)
 *    move_bcs        = `move_bcs` (+ `bounceback_in_obstacle`) (:329-337, :510-518)
 *    update_hydro    = `update_hydro`                          (:355-362)
 *    update_feq      = `update_feq`                            (:295-306)
 *    collide         = `collide_particles`                     (:364-370)
 *    zero_velocity   = `set_zero_velocity_in_obstacle`         (:506-508)
 *    With the other schemes the same five calls are the methods of the class the scheme mirrors:
 *    LB_SCHEME_CYTHON[_OLD]: move_bcs / move / update_hydro / update_feq / collide_particles of
 *    cython_dim.pyx:204-344 (OLD/cython.pyx:97-360), obstacle swap and velocity zeroing included when a
 *    mask is set; LB_SCHEME_OPENCL_OLD: D2Q9.cl's move_bcs_PeriodicBC_VelocityInlet (+ bounce-back),
 *    move + copy_buffer, update_hydro_PeriodicBC_VelocityInlet (+ zeroing), update_feq, collide_particles
 *    as OLD/opencl.py:189-255, :290-371 launches them. */
int lb_stage_move(lb_sim *sim);
int lb_stage_move_bcs(lb_sim *sim);
int lb_stage_update_hydro(lb_sim *sim);
int lb_stage_update_feq(lb_sim *sim);
int lb_stage_collide(lb_sim *sim);
int lb_stage_zero_velocity(lb_sim *sim);
/* -- device-side initialisers for grids too large for host init (synthetic benchmark inputs) */
enum lb_synth { LB_SYNTH_PIPE_RAMP = 0, LB_SYNTH_SHEAR_LAYERS = 1 };
/* rho/u/v from an analytic profile in GLOBAL coordinates, f = feq*(1+amplitude*N(0,1)) with a
 * counter-based generator keyed on (seed, global cell, population): slab-decomposition invariant.
 *   PIPE_RAMP:    rho = inlet - gx*(inlet-outlet)/global_nx, u=v=0          (opencl_dim.py:279-288)
 *   SHEAR_LAYERS: rho=1, u=U0*tanh(80(y/ny-1/4)) | U0*tanh(80(3/4-y/ny)), v=0.05*U0*sin(2pi(gx/gnx+1/4)) */
int lb_init_synthetic(lb_sim *sim, int kind, double u0, double amplitude, uint64_t seed);
/* solid disk in GLOBAL coordinates (centre cx,cy, radius r): mask = (gx-cx)^2+(y-cy)^2 < r^2 */
int lb_set_mask_disk(lb_sim *sim, double cx, double cy, double r);

/* -- several lattice updates per pass through HBM (csrc/lb_march.cuh).  lb_step runs a run's steps two or three
 *    at a time in one launch -- the intermediate time levels live on chip, so only 38 B (two updates per launch)
 *    or 26 B (three) per lattice update cross the HBM interface in fp32 instead of 72 (fp64: twice that) -- preceded by one shorter
 *    launch when n_steps is not a multiple of the launch depth; the launch that ends the run also stores
 *    rho, u, v.  Results are bit-identical to the one-update kernel in both math modes, on single slabs and on
 *    halo-connected slabs (which then exchange their three outermost columns once per launch instead of one
 *    column every step).
 *    shape -1 = automatic (default): when the WHOLE lattice (global_nx x ny) has at least 2^22 nodes and
 *    ny >= 64, the measured-best shape for this slab -- three updates per launch where the slab is
 *    large enough for segments of 16+ rows, two otherwise; branch-free obstacle code where there is
 *    a mask; a segment height of 8 to 128 rows that leaves thousands of (strip, segment) work items, the last rows of a large
 *    launch in segments a quarter as high -- and the
 *    graph-batched one-update kernel below that size; 0 = off; 1 .. lb_tb2_shape_count()-1 = a compiled shape
 *    by index (lb_tb2_shape_name: "march.w<warps per CTA>b<CTAs per SM>[.sh[.bf] | .scalar].s<rows per segment>"
 *    = two updates per launch, "march3.w..b...s.." = three).  Serves LB_SCHEME_OPENCL /
 *    LB_MODEL_D2Q9 lattices; a single-slab periodic box needs nx to be a multiple of the vector width (4 fp32 /
 *    2 fp64 cells); three updates per launch: slabs at least 3 columns wide and 3 rows high (fp64 strips then
 *    store 56 of 64 loaded columns, two overlap lanes per side).  All slabs of one lattice must
 *    use shapes of the same depth (the automatic choice does).  lb_temporal_blocking returns the shape lb_step will
 *    use (0 = one-update kernel); lb_segment_rows the rows per segment of its launches (0 with the one-update
 *    kernel): the height in the shape's name, except that the automatic choice, on lattices of three waves of CTAs or
 *    fewer, takes the height between 6 and 64 rows that fills the last wave best (C2: 11 rows, one wave). */
int lb_set_temporal_blocking(lb_sim *sim, int shape);
int lb_temporal_blocking(const lb_sim *sim);
int lb_segment_rows(const lb_sim *sim);
int lb_tb2_shape_count(void);
const char *lb_tb2_shape_name(int shape);

/* -- diagnostics */
/* Device self-test of the branch-free reciprocal used by STRICT fp32 math: compares it with IEEE
 * division for every float whose bit pattern lies in [first_bits, last_bits]; returns the count of
 * mismatches in *mismatches. */
int lb_selftest_rcp(int device, uint32_t first_bits, uint32_t last_bits, uint64_t *mismatches);
/* Practical HBM ceiling of the step's access pattern on THIS device, now: `reps` launches of an
 * arithmetic-free kernel with the fused kernel's thread mapping (nine 128-bit loads with the D2Q9 row
 * offsets, nine 128-bit stores) from the handle's current buffer into its scratch buffer, timed with
 * CUDA events on the handle's stream; *ms_per_launch gets the average.  The populations are not
 * modified (the scratch buffer is the other half of the ping-pong pair, rewritten by the next step).
 * bench.py reports it beside the roofline; tools/stream_ceiling*.cu are the stand-alone versions. */
int lb_selftest_copy(lb_sim *sim, int reps, double *ms_per_launch);
/* Diagnostic, no device needed: the launch geometry the marching kernels' launchers compute for a slab nx columns wide and a
 * row range of `rows` rows -- strips (120 stored columns per warp in fp32, 60 in fp64, 56 in fp64 with depth 3), strips that
 * take part in the halo hand-shake (a last strip narrower than the three published columns makes the one before it an edge
 * strip as well), and the segments: n_tall of seg_rows rows, then n_short of *short_rows (seg_rows2 on entry: 0 = uniform;
 * the launcher falls back to uniform when two waves of short work items would be half the range or more).  The CPU tests
 * check that every row is covered exactly once. */
int lb_plan_march_launch(int nx, int rows, int elem_bytes, int depth, int nw, int minb, int seg_rows, int seg_rows2, int sm_count,
                         int west_halo, int east_halo, int *n_strips, int *n_edge_strips, int *n_tall, int *n_short, int *short_rows);
/* sum over all populations and cells of this slab, accumulated in double (mass check) */
int lb_total_mass(lb_sim *sim, double *out);
/* order-independent exact checksum of the populations: the 64-bit wrap-around sum of the raw bit
 * patterns of all 9*nx*ny values (fp32 patterns are zero-extended).  Equal checksums under a
 * re-decomposition into slabs or a periodic shift of the lattice mean bit-identical multisets. */
int lb_checksum(lb_sim *sim, uint64_t *out);
/* number of fused-kernel launches issued by this handle since creation (a two-update launch counts once) */
int64_t lb_launch_count(const lb_sim *sim);
/* choose one of the compiled tile configurations of the one-update fused kernel (-1 = default);
 * lb_variant_count/lb_variant_name enumerate them.  Tuning only: results do not change.  Picking a
 * variant by hand also switches automatic temporal blocking off (the chosen kernel is what runs);
 * -1 restores both defaults. */
int lb_set_variant(lb_sim *sim, int variant);
int lb_variant_count(void);
const char *lb_variant_name(int variant);
/* device pointers, for zero-copy consumers (torch tensors as buffers) */
int lb_device_ptr(lb_sim *sim, int field, void **ptr, int64_t *pitch_elems);
/* the cudaStream_t this handle enqueues on (so that several handles can share one stream) */
void *lb_stream(lb_sim *sim);

/* -- x-slab halo exchange over NVLink peer memory (new design; the reference is single-device).
 *    Each slab owns a halo arena: two ghost arenas x two parities x all nine populations
 *    of the neighbour's three outermost columns (a one-update launch reads three values
 *    per row of them; a K-update launch patches K columns into its overlap lane and
 *    advances the neighbour's outermost K-1 columns itself), the neighbours' mask
 *    columns, and launch flags.  A slab's boundary threads store those
 *    values straight into the NEIGHBOUR's arena inside the fused kernel and then publish a
 *    flag there; the neighbour's boundary tiles poll their local flag before reading. */
#define LB_IPC_HANDLE_BYTES 64
int lb_halo_ipc_handle(lb_sim *sim, void *out_handle /* LB_IPC_HANDLE_BYTES */);
/* connect `side` to a neighbour slab living in another process (CUDA IPC) ... */
int lb_halo_connect_ipc(lb_sim *sim, int side, const void *peer_handle, int peer_device);
/* ... or in this process (virtual ranks on one device / several devices of one process) */
int lb_halo_connect_local(lb_sim *sim, int side, lb_sim *peer);
/* push the current state's two outermost columns and the boundary column of the obstacle mask to the
 * neighbours (call on every slab after uploads / mask changes / initialisers and before the first lb_step;
 * callers barrier in between).  Also clears a previous LB_ERR_HALO condition of this handle. */
int lb_halo_prime(lb_sim *sim);
/* bound of the in-kernel wait for a neighbour's ghost columns (default 30 s; the wait is a hand-shake between
 * GPUs that normally lasts microseconds).  When it expires the waiting tiles skip their work -- nothing derived
 * from stale ghost data is stored or published -- and the next lb_sync returns LB_ERR_HALO; upload + prime
 * recovers the handle. */
int lb_set_halo_timeout(lb_sim *sim, double seconds);
/* LIFETIME of connected slabs: a neighbour's last launch may still be storing into this handle's halo arena
 * when this handle's own stream is idle.  Before lb_destroy, lb_sync EVERY slab of the lattice (and, across
 * processes, barrier) -- lb_b200.slab.SlabLattice.close and LocalSlabs.close do. */

/* -- one lattice on several slabs / devices behind ONE handle.  The reference binds one command queue to
 *    devices[0] (opencl_dim.py:229-240); this is the multi-device form of the same calls.  `cfg` describes the
 *    WHOLE lattice (global_nx = nx, x_offset = 0; edges are derived).  The lattice is cut into n_slabs x-slabs
 *    (widths differ by at most one column, remainder to the first slabs); slab k lives on device_ids[k].  The
 *    handle owns the per-slab lb_sim handles, their streams and peer mappings, keeps the ghost columns primed,
 *    and runs the step loop: with one slab per device the slabs advance asynchronously (bounded chunks,
 *    synchronised by the in-kernel NVLink flags); slabs that share a device share a stream and advance in
 *    lock-step ("virtual ranks": the arithmetic-neutrality tests).  Host arrays are GLOBAL arrays in the
 *    layout stated at the top of this file; every slab copies its own columns.  Results are bit-identical to
 *    the single-slab lattice.  (One process per GPU -- torch.distributed ranks -- uses lb_create +
 *    lb_halo_ipc_handle / lb_halo_connect_ipc instead: lb_b200.slab.SlabLattice.) */
typedef struct lb_multi lb_multi;
int lb_multi_create(const lb_config *cfg, int n_slabs, const int *device_ids, lb_multi **out);
int lb_multi_destroy(lb_multi *m);     /* drains every slab before freeing any arena */
const char *lb_multi_last_error(const lb_multi *m);   /* m == NULL: the last failed lb_multi_create */
int lb_multi_slab_count(const lb_multi *m);
int lb_multi_slab(lb_multi *m, int k, lb_sim **slab, int *x_offset, int *nx);   /* borrow slab k (tuning, diagnostics) */
int lb_multi_set_mask(lb_multi *m, const void *host_mask, int elem_bytes);       /* lb_set_mask, global [ny][nx] */
int lb_multi_set_mask_disk(lb_multi *m, double cx, double cy, double r);
int lb_multi_upload_f(lb_multi *m, const void *host_f);                           /* lb_upload_f, global; primes the ghosts */
int lb_multi_upload_moments(lb_multi *m, const void *host_rho, const void *host_u, const void *host_v);
int lb_multi_init_synthetic(lb_multi *m, int kind, double u0, double amplitude, uint64_t seed);
int lb_multi_set_temporal_blocking(lb_multi *m, int shape);                      /* the same shape on every slab */
int lb_multi_temporal_blocking(const lb_multi *m);
int lb_multi_prime(lb_multi *m);       /* republish ghost columns after changing slab state through lb_multi_slab */
int lb_multi_step(lb_multi *m, int n_steps);                                      /* Pipe_Flow.run, opencl_dim.py:372-387 */
int lb_multi_sync(lb_multi *m);
int lb_multi_download(lb_multi *m, int field, void *host_out);                    /* lb_download, global */
int lb_multi_total_mass(lb_multi *m, double *out);
int lb_multi_checksum(lb_multi *m, uint64_t *out);
int64_t lb_multi_launch_count(const lb_multi *m);

#ifdef __cplusplus
}
#endif
#endif /* LB_D2Q9_H */
