// The hot path: ONE kernel per lattice update.
//
// Replaces the reference's seven launches per step (opencl_dim.py:372-387, :510-518):
//   move + copy_buffer (D2Q9.cl:139-171, :123-137)  -> pull-streaming from the `src` buffer
//   move_bcs (D2Q9.cl:173-261)                      -> pipe_bc() on boundary threads only
//   bounceback_in_obstacle (D2Q9.cl:398-433)        -> bounce_back() where the mask says solid
//   update_hydro + update_feq + collide_particles   -> collide_node() in registers
// and writes the post-collision populations into the `dst` buffer (ping-pong).
//
// Data layout: f[9][ny][pitch], x fastest, pitch a multiple of 512 B so
// every warp-row is either fully inside or fully outside a row; two guard rows before and
// after the 9 planes make every shifted vector / scalar access land in owned memory.
//
// Thread mapping: a warp owns SPAN = 32*V consecutive cells of a row (V = cells per thread,
// 128-bit accesses for V*sizeof(T) == 16).  The +-1 x-shift of populations 1,5,8 / 3,6,7 is
// resolved in registers: aligned vector load, one warp shuffle for the element that crosses
// the thread boundary, one predicated scalar load on lane 0 / lane 31 for the element that
// crosses the warp boundary.  A CTA is WX warps wide and WY warps tall; each warp walks R
// rows (edge tiles of a halo-connected slab walk `edge_rows` rows instead).  Every plane element is read exactly once per step, so the algorithmic traffic is
// 9 loads + 9 stores per cell: 72 B (fp32) / 144 B (fp64) per lattice update.
#pragma once
#include "lb_device.cuh"

namespace lb {

struct StepParams {
    const void *src;          // plane 0, row 0 of the buffer being read
    void *dst;                // plane 0, row 0 of the buffer being written
    long long plane;          // elements between consecutive population planes
    int nx, ny, pitch;        // slab extent and row pitch (elements)
    int gnx, x_off;           // global width, global x of local column 0
    int bc;                   // BC_PIPE / BC_PERIODIC
    int west, east;           // EDGE_* beyond column 0 / column nx-1
    int write_moments;        // store rho,u,v (last step of a run only)
    int zero_obstacle_velocity;
    const uint8_t *mask;      // [ny][mask_pitch], 1 = solid; nullptr = no obstacles
    const uint8_t *span_solid;// [ny][nspans]: per group of 32 cells of a row: 0 no solid node, 1 some, 2 all
    int mask_pitch, nspans;
    void *rho, *u, *v;        // [ny][pitch]
    Consts<float> cf;         // per-launch constants, precomputed on the host in both types
    Consts<double> cd;
    Consts<F2> c2;            // cf with both lanes of every packed register set (lb_f32x2.cuh)
    // x-slab halo (EDGE_HALO): a ghost arena holds the neighbour's GHOST_COLS outermost columns, all nine
    // post-collision populations of each, as [column c][population j][y + 1]: c = 0 is the neighbour's boundary
    // column (its last column for the west ghost, its first for the east ghost), c = 1 the one behind it, ...  A
    // one-update launch reads three values per row of column 0 (the populations entering this slab); a K-update
    // launch of the marching kernel patches them into its overlap lane "as if memory continued beyond the slab" and
    // advances the neighbour's outermost K-1 columns itself (K <= GHOST_COLS).
    const void *ghost_w;      // filled by the west neighbour (read)
    const void *ghost_e;      // filled by the east neighbour (read)
    void *out_w;              // west neighbour's east ghost arena (written from my columns 0, 1, 2)
    void *out_e;              // east neighbour's west ghost arena (written from my columns nx-1, nx-2, nx-3)
    const uint8_t *gmask_w, *gmask_e;   // [GHOST_COLS - 1][ny] obstacle mask of the neighbours' outermost columns (nullptr: none)
    unsigned int *flag_w_local, *flag_e_local;     // polled: neighbour's data for this step is in
    unsigned int *flag_w_remote, *flag_e_remote;   // published: my data for the next step is out
    unsigned int *done_w, *done_e;                 // edge-tile completion counters (local)
    unsigned int *error_word;                      // set on hand-shake timeout
    unsigned long long halo_timeout_ns;            // bound of one flag wait (lb_set_halo_timeout)
    unsigned int step_id;                          // flag value that must be visible before reading ghosts
    int tiles_x, tiles_y;
    int edge_first;           // 1-D grid with the two edge tile columns first (halo overlap); else 2-D/3-D grid
    int edge_rows;            // rows per warp in an edge tile (tall tiles: few participants in the hand-shake)
    int edge_tiles_y;         // edge tiles per side
    int east_pair;            // one-update kernel: the tile column before the last one holds some of the GHOST_COLS
                              // easternmost columns and takes part in the east hand-shake (done_e_count CTAs in all)
    int done_e_count;
    int y_begin, y_end;       // rows this launch updates (whole lattice: 0, ny); band launches of lb_step_banded
    int seg_rows;             // marching kernel (lb_march.cuh): rows per segment (within y_begin .. y_end)
    int seg_rows2, seg_tall;  // ... the first seg_tall segments are seg_rows high, the rest seg_rows2 (shorter work items at
                              // the end of the grid: a shorter tail; seg_rows2 = 0 on entry to the launcher: uniform)
    int sm_count;
};
enum : int { GHOST_COLS = 3, GHOST_SLOTS = 9 * GHOST_COLS };

template <typename T> __device__ __forceinline__ const Consts<T> &consts_in(const StepParams &p);
template <> __device__ __forceinline__ const Consts<float> &consts_in<float>(const StepParams &p) { return p.cf; }
template <> __device__ __forceinline__ const Consts<double> &consts_in<double>(const StepParams &p) { return p.cd; }

// ---- vector access helpers ----------------------------------------------------------
template <typename T, int V> struct VecOf;
template <> struct VecOf<float, 4> { using type = float4; };
template <> struct VecOf<float, 2> { using type = float2; };
template <> struct VecOf<float, 1> { using type = float; };
template <> struct VecOf<double, 2> { using type = double2; };
template <> struct VecOf<double, 1> { using type = double; };

template <typename T, int V> struct Pack { T v[V]; };

__device__ __forceinline__ void unpack(const float4 &a, Pack<float, 4> &p) { p.v[0] = a.x; p.v[1] = a.y; p.v[2] = a.z; p.v[3] = a.w; }
__device__ __forceinline__ void unpack(const float2 &a, Pack<float, 2> &p) { p.v[0] = a.x; p.v[1] = a.y; }
__device__ __forceinline__ void unpack(const float &a, Pack<float, 1> &p) { p.v[0] = a; }
__device__ __forceinline__ void unpack(const double2 &a, Pack<double, 2> &p) { p.v[0] = a.x; p.v[1] = a.y; }
__device__ __forceinline__ void unpack(const double &a, Pack<double, 1> &p) { p.v[0] = a; }
__device__ __forceinline__ float4 repack(const Pack<float, 4> &p) { return make_float4(p.v[0], p.v[1], p.v[2], p.v[3]); }
__device__ __forceinline__ float2 repack(const Pack<float, 2> &p) { return make_float2(p.v[0], p.v[1]); }
__device__ __forceinline__ float repack(const Pack<float, 1> &p) { return p.v[0]; }
__device__ __forceinline__ double2 repack(const Pack<double, 2> &p) { return make_double2(p.v[0], p.v[1]); }
__device__ __forceinline__ double repack(const Pack<double, 1> &p) { return p.v[0]; }

// LDP: 0 = ld.global, 1 = ld.global.nc (read-only path), 2 = ld.global.cs (streaming / evict-first)
template <int LDP, typename VT>
__device__ __forceinline__ VT ld_vec(const VT *p)
{
    if (LDP == 1) return __ldg(p);
    if (LDP == 2) return __ldcs(p);
    return *p;
}
// STP: 0 = st.global, 1 = st.global.cs (streaming), 2 = st.global.wt
template <int STP, typename VT>
__device__ __forceinline__ void st_vec(VT *p, const VT &v)
{
    if (STP == 1) __stcs(p, v);
    else if (STP == 2) __stwt(p, v);
    else *p = v;
}

// predicated scalar load through the read-only path: `@p ld.global.nc` -- no branch
template <typename T> __device__ __forceinline__ T ld_if(const T *p, bool pred);
template <> __device__ __forceinline__ float ld_if<float>(const float *p, bool pred)
{
    float v = 0.0f;
    asm volatile("{\n.reg .pred P;\nsetp.ne.b32 P, %2, 0;\n@P ld.global.nc.f32 %0, [%1];\n}" : "+f"(v) : "l"(p), "r"((int)pred));
    return v;
}
template <> __device__ __forceinline__ double ld_if<double>(const double *p, bool pred)
{
    double v = 0.0;
    asm volatile("{\n.reg .pred P;\nsetp.ne.b32 P, %2, 0;\n@P ld.global.nc.f64 %0, [%1];\n}" : "+d"(v) : "l"(p), "r"((int)pred));
    return v;
}

template <typename T, int V, int LDP>
__device__ __forceinline__ Pack<T, V> load_pack(const T *p)
{
    using VT = typename VecOf<T, V>::type;
    Pack<T, V> r;
    unpack(ld_vec<LDP>(reinterpret_cast<const VT *>(p)), r);
    return r;
}
template <typename T, int V, int STP>
__device__ __forceinline__ void store_pack(T *p, const Pack<T, V> &r)
{
    using VT = typename VecOf<T, V>::type;
    st_vec<STP>(reinterpret_cast<VT *>(p), repack(r));
}

// ---- halo hand-shake ----------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait (bounded by `timeout_ns`) until *flag >= want.  One thread polls, the CTA follows through a barrier.
// Returns false -- for every thread of the CTA -- when the wait timed out now or an earlier wait of this
// handle did: the caller then skips its tile, so that nothing computed from stale ghost columns is stored,
// published to a neighbour or released by a flag.  The fault stays local; lb_sync reports LB_ERR_HALO, and
// the neighbours, which never see this slab's flag, time out in turn.  lb_halo_prime clears the error.
__device__ __forceinline__ bool wait_flag(const unsigned int *flag, unsigned int want, unsigned int *error_word,
                                          unsigned long long timeout_ns)
{
    __shared__ unsigned int wait_failed;
    if (threadIdx.x == 0) {
        unsigned int failed = *reinterpret_cast<volatile unsigned int *>(error_word);
        const unsigned long long t0 = globaltimer_ns();
        while (!failed && (int)(ld_acquire_sys(flag) - want) < 0) {
            if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(error_word, 1u); failed = 1u; break; }
            __nanosleep(64);
        }
        wait_failed = failed;
    }
    __syncthreads();
    const bool ok = wait_failed == 0u;
    __syncthreads();                                   // the shared word may be reused by the CTA's second wait
    return ok;
}

// ---- everything after the nine shifted populations of a thread's V nodes are in registers:
//      slab-edge fix-ups, boundary closure, obstacles, collision, stores, halo publication.
//      Shared by the register-shuffle kernel below and the TMA kernel (lb_tma.cuh).
//      ROLE (temporal blocking, lb_tb2.cuh): ROW_FULL = everything; ROW_TO_REGISTERS = stop after the
//      collision, the caller keeps q (phase 1 stores it to shared memory); ROW_FROM_TILE = the populations
//      came from a shared-memory tile whose rim already holds the wrapped / ghost values, so the slab-edge
//      fix-up from `src` is skipped (phase 2).
//      The marching kernel (lb_march.cuh) brings its own slab-edge values and obstacle bits: ROW_REGS_NOFIX =
//      ROW_TO_REGISTERS without the fix-up; CALLER_MASK = `solid_in` holds one bit per node of the thread (the
//      per-32-cell flag array is indexed by aligned spans, which that kernel's overlapping strips are not);
//      `store_ok` = false suppresses the stores and the halo publication of a lane that only computes overlap.
//      ZOV = 0 compiles the zeroing of obstacle velocities out for handles that do not ask for it (ZOV = -1:
//      decided at run time by the flag): the marching kernel's loop body is two of these back to back, and ncu
//      showed 10 % of its stall cycles to be instruction fetches.  (Turning the boundary closure into a call as
//      well made the kernel 11 % slower -- the call's register conventions reach into the hot path --
//      profiles/r2_march_regs_vs_smem_window_bc_inline_vs_call.txt.)
enum : int { ROW_FULL = 0, ROW_TO_REGISTERS = 1, ROW_FROM_TILE = 2, ROW_REGS_NOFIX = 3 };

//      MASKED (with CALLER_MASK): -1 = look at p.mask at run time and branch (pack swap for all-solid warps, per-node
//      swaps otherwise); 0 = the handle has no mask, no obstacle code at all; 1 = it has one, and the bounce-back swap
//      is eight selects per node, always executed, no branch.  The branchy form renames whole register packs on one
//      of its paths, and where the paths join the compiler moves them back: ncu attributed 80 % of the marching
//      kernel's executed MOVs -- 11 % of all its instructions -- to those joins (profiles/README.md section 9).
template <typename T, int V, int MATH, int STP, int MODEL, int ROLE = ROW_FULL, bool PACKED = false, bool CALLER_MASK = false, int ZOV = -1,
          int MASKED = -1>
__device__ __forceinline__ void finish_row(const StepParams &p, const Consts<T> &c, Pack<T, V> (&q)[9],
                                           const T *__restrict__ src, T *__restrict__ dst,
                                           int x0, int span0, int y, int ym, int yp,
                                           uint32_t solid_in = 0u, bool store_ok = true)
{
    constexpr int SPAN = 32 * V;
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const bool periodic = (p.bc == BC_PERIODIC);
    const int el_east = (nx - 1) - x0;                // element index of column nx-1 in this thread, if in [0,V)
    const bool has_west = (x0 == 0);
    const bool has_east = (el_east >= 0 && el_east < V);
    const long long rc = (long long)y * pitch + x0;

    // --- threads that own the first / last column of the slab or a wall row (rare): slab-edge
    //     fix-up (wrap, ghost column) and the boundary closure, behind ONE branch
    const bool row_is_wall = (!periodic) && (y == 0 || y == ny - 1);
    if (has_west || has_east || row_is_wall) {
        if (ROLE != ROW_FROM_TILE && ROLE != ROW_REGS_NOFIX && has_west && p.west != EDGE_BOUNDARY) {
            T a1, a5, a8;
            if (p.west == EDGE_WRAP) {
                a1 = src[1 * plane + (long long)y * pitch + (nx - 1)];
                a5 = src[5 * plane + (long long)ym * pitch + (nx - 1)];
                a8 = src[8 * plane + (long long)yp * pitch + (nx - 1)];
            } else {
                const T *gw = static_cast<const T *>(p.ghost_w);          // column 0 of the ghost: populations 1, 5, 8
                a1 = __ldcv(gw + 1 * (ny + 2) + (y + 1));
                a5 = __ldcv(gw + 5 * (ny + 2) + (ym + 1));
                a8 = __ldcv(gw + 8 * (ny + 2) + (yp + 1));
            }
            q[1].v[0] = a1; q[5].v[0] = a5; q[8].v[0] = a8;
        }
        if (ROLE != ROW_FROM_TILE && ROLE != ROW_REGS_NOFIX && has_east && p.east != EDGE_BOUNDARY) {
            T a3, a6, a7;
            if (p.east == EDGE_WRAP) {
                a3 = src[3 * plane + (long long)y * pitch];
                a6 = src[6 * plane + (long long)ym * pitch];
                a7 = src[7 * plane + (long long)yp * pitch];
            } else {
                const T *ge = static_cast<const T *>(p.ghost_e);          // column 0 of the ghost: populations 3, 6, 7
                a3 = __ldcv(ge + 3 * (ny + 2) + (y + 1));
                a6 = __ldcv(ge + 6 * (ny + 2) + (ym + 1));
                a7 = __ldcv(ge + 7 * (ny + 2) + (yp + 1));
            }
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (e == el_east) { q[3].v[e] = a3; q[6].v[e] = a6; q[7].v[e] = a7; }
        }
        // boundary closure: only threads that own a wall / inlet / outlet node
        if ((!periodic) && (row_is_wall || (has_west && p.west == EDGE_BOUNDARY) ||
                            (has_east && p.east == EDGE_BOUNDARY))) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                T g[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
                pipe_bc<T, MODEL>(c, p.x_off + x0 + e, y, p.gnx, ny, g);
#pragma unroll
                for (int j = 0; j < 9; ++j) q[j].v[e] = g[j];
            }
        }
    }

    // --- obstacles: one flag byte per (row, 32 cells): 0 = no solid node, 1 = some, 2 = all 32 solid.
    //     Every lane of a warp reads the same SPAN/32 flags, so the branches below are warp-uniform:
    //     warps without solids issue nothing here, and warps deep inside an obstacle (all flags 2) swap
    //     whole register packs -- a compile-time renaming, no mask bytes loaded, no per-node selects.
    uint32_t solid_bits = 0;
    bool all_solid = false;
    if constexpr (CALLER_MASK && MASKED == 0) {
        // no mask on this handle: nothing to do
    } else if constexpr (CALLER_MASK && MASKED == 1) {
        solid_bits = solid_in;
        const bool zov_ = ZOV < 0 ? (p.zero_obstacle_velocity != 0) : (ZOV != 0);
        if (zov_) all_solid = __all_sync(0xffffffffu, solid_bits == (1u << V) - 1u);
        // fp64 (two instructions per select) skips them by a warp-uniform branch when no node of the warp's row is solid:
        // +3 % on C5; in fp32 the same branch costs 1-4 % (profiles/r2_ab_bf_skip.txt; -DLB_BF_SKIP builds it there too)
#ifdef LB_BF_SKIP
        constexpr bool skip_if_fluid = true;
#else
        constexpr bool skip_if_fluid = sizeof(T) == 8;
#endif
        if (!skip_if_fluid || __any_sync(0xffffffffu, solid_bits != 0))
#pragma unroll
        for (int e = 0; e < V; ++e) {                     // D2Q9.cl:410-431, as selects
            const bool sd = (solid_bits >> e) & 1u;
            const T f1 = q[1].v[e], f2 = q[2].v[e], f3 = q[3].v[e], f4 = q[4].v[e];
            const T f5 = q[5].v[e], f6 = q[6].v[e], f7 = q[7].v[e], f8 = q[8].v[e];
            q[1].v[e] = sd ? f3 : f1; q[3].v[e] = sd ? f1 : f3;
            q[2].v[e] = sd ? f4 : f2; q[4].v[e] = sd ? f2 : f4;
            q[5].v[e] = sd ? f7 : f5; q[7].v[e] = sd ? f5 : f7;
            q[6].v[e] = sd ? f8 : f6; q[8].v[e] = sd ? f6 : f8;
        }
    } else if constexpr (CALLER_MASK) {
        solid_bits = solid_in;
        if (p.mask != nullptr && __any_sync(0xffffffffu, solid_bits != 0)) {
            if (__all_sync(0xffffffffu, solid_bits == (1u << V) - 1u)) {     // D2Q9.cl:410-431 on every node of the warp
                all_solid = true;
                Pack<T, V> t;
                t = q[1]; q[1] = q[3]; q[3] = t;
                t = q[2]; q[2] = q[4]; q[4] = t;
                t = q[5]; q[5] = q[7]; q[7] = t;
                t = q[6]; q[6] = q[8]; q[8] = t;
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if ((solid_bits >> e) & 1u) {
                        T t;
                        t = q[1].v[e]; q[1].v[e] = q[3].v[e]; q[3].v[e] = t;
                        t = q[2].v[e]; q[2].v[e] = q[4].v[e]; q[4].v[e] = t;
                        t = q[5].v[e]; q[5].v[e] = q[7].v[e]; q[7].v[e] = t;
                        t = q[6].v[e]; q[6].v[e] = q[8].v[e]; q[8].v[e] = t;
                    }
                }
            }
        }
    } else if (p.mask != nullptr) {
        const uint8_t *sf = p.span_solid + (long long)y * p.nspans + (span0 >> 5);
        uint32_t flags = 0, all = 0;
        if (SPAN == 128) { flags = *reinterpret_cast<const uint32_t *>(sf); all = 0x02020202u; }
        else if (SPAN == 64) { flags = *reinterpret_cast<const uint16_t *>(sf); all = 0x0202u; }
        else {
#pragma unroll
            for (int k = 0; k < SPAN / 32; ++k) { flags |= (uint32_t)sf[k] << (8 * (k & 3)); all |= 2u << (8 * (k & 3)); }
            if (SPAN / 32 > 4) all = 0xffffffffu;        // wider spans: never take the all-solid shortcut
        }
        if (flags == all) {                               // D2Q9.cl:410-431 on every node of the warp
            all_solid = true;
            solid_bits = (1u << V) - 1u;
            Pack<T, V> t;
            t = q[1]; q[1] = q[3]; q[3] = t;
            t = q[2]; q[2] = q[4]; q[4] = t;
            t = q[5]; q[5] = q[7]; q[7] = t;
            t = q[6]; q[6] = q[8]; q[8] = t;
        } else if (flags != 0) {
            const uint8_t *mrow = p.mask + (long long)y * p.mask_pitch + x0;   // bytes past nx are zero
            if (V == 4) {
                const uint32_t m = *reinterpret_cast<const uint32_t *>(mrow);
#pragma unroll
                for (int e = 0; e < V; ++e) if (((m >> (8 * e)) & 0xffu) == 1u) solid_bits |= (1u << e);
            } else if (V == 2) {
                const uint32_t m = *reinterpret_cast<const uint16_t *>(mrow);
#pragma unroll
                for (int e = 0; e < V; ++e) if (((m >> (8 * e)) & 0xffu) == 1u) solid_bits |= (1u << e);
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) if (mrow[e] == 1) solid_bits |= (1u << e);
            }
            if (__any_sync(0xffffffffu, solid_bits != 0)) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if ((solid_bits >> e) & 1u) {         // D2Q9.cl:410-431
                        T t;
                        t = q[1].v[e]; q[1].v[e] = q[3].v[e]; q[3].v[e] = t;
                        t = q[2].v[e]; q[2].v[e] = q[4].v[e]; q[4].v[e] = t;
                        t = q[5].v[e]; q[5].v[e] = q[7].v[e]; q[7].v[e] = t;
                        t = q[6].v[e]; q[6].v[e] = q[8].v[e]; q[8].v[e] = t;
                    }
                }
            }
        }
    }
    // --- per node: moments + equilibrium + BGK relaxation, in registers.  Three warp-uniform copies of
    //     the loop: no velocity to zero (the common one, flag-free); every node solid with zeroed
    //     velocity (the equilibrium folds to w*rho at compile time); mixed warps with a per-node flag.
    Pack<T, V> mrho, mu, mv;
    const bool zov = ZOV < 0 ? (p.zero_obstacle_velocity != 0) : (ZOV != 0);
    if (zov && all_solid) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            T g[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
            collide_node<T, MATH, MODEL>(c, g, mrho.v[e], mu.v[e], mv.v[e], true);
#pragma unroll
            for (int j = 0; j < 9; ++j) q[j].v[e] = g[j];
        }
    } else if (zov && __any_sync(0xffffffffu, solid_bits != 0)) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            T g[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
            collide_node<T, MATH, MODEL>(c, g, mrho.v[e], mu.v[e], mv.v[e], (solid_bits >> e) & 1u);
#pragma unroll
            for (int j = 0; j < 9; ++j) q[j].v[e] = g[j];
        }
    } else if constexpr (PACKED && sizeof(T) == 4 && V % 2 == 0 && MODEL == MODEL_D2Q9) {
        // two nodes per instruction: FADD2 / FMUL2 / FFMA2, each lane rounded like the scalar code above
#pragma unroll
        for (int e = 0; e < V; e += 2) {
            F2 g[9], r2, u2, v2;
#pragma unroll
            for (int j = 0; j < 9; ++j) g[j] = F2((float)q[j].v[e], (float)q[j].v[e + 1]);
            collide_node<F2, MATH, MODEL>(p.c2, g, r2, u2, v2, false);
#pragma unroll
            for (int j = 0; j < 9; ++j) { q[j].v[e] = (T)g[j].lo(); q[j].v[e + 1] = (T)g[j].hi(); }
            mrho.v[e] = (T)r2.lo(); mrho.v[e + 1] = (T)r2.hi();
            mu.v[e] = (T)u2.lo(); mu.v[e + 1] = (T)u2.hi();
            mv.v[e] = (T)v2.lo(); mv.v[e + 1] = (T)v2.hi();
        }
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            T g[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) g[j] = q[j].v[e];
            collide_node<T, MATH, MODEL>(c, g, mrho.v[e], mu.v[e], mv.v[e], false);
#pragma unroll
            for (int j = 0; j < 9; ++j) q[j].v[e] = g[j];
        }
    }

    if (ROLE == ROW_TO_REGISTERS || ROLE == ROW_REGS_NOFIX) return;
    if (!store_ok) return;

    // --- stores: aligned vectors; the one thread straddling column nx-1 goes scalar ---
    if (x0 + V <= nx) {
        T *pd = dst + rc;
#pragma unroll
        for (int j = 0; j < 9; ++j) store_pack<T, V, STP>(pd + j * plane, q[j]);
        if (p.write_moments) {
            store_pack<T, V, 0>(static_cast<T *>(p.rho) + rc, mrho);
            store_pack<T, V, 0>(static_cast<T *>(p.u) + rc, mu);
            store_pack<T, V, 0>(static_cast<T *>(p.v) + rc, mv);
        }
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (x0 + e < nx) {
#pragma unroll
                for (int j = 0; j < 9; ++j) dst[j * plane + rc + e] = q[j].v[e];
                if (p.write_moments) {
                    static_cast<T *>(p.rho)[rc + e] = mrho.v[e];
                    static_cast<T *>(p.u)[rc + e] = mu.v[e];
                    static_cast<T *>(p.v)[rc + e] = mv.v[e];
                }
            }
        }
    }

    // --- publish my GHOST_COLS outermost columns, all nine populations, into the neighbours' ghost arenas ---
    const int gs = ny + 2;                            // rows per (column, population) of a ghost arena
    if (p.west == EDGE_HALO && x0 < GHOST_COLS) {
        T *ow = static_cast<T *>(p.out_w) + (y + 1);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int col = x0 + e;                   // my column `col` is column `col` of the west neighbour's east ghost
            if (col >= 0 && col < GHOST_COLS && col < nx) {
#pragma unroll
                for (int j = 0; j < 9; ++j) ow[(col * 9 + j) * gs] = q[j].v[e];
            }
        }
    }
    if (p.east == EDGE_HALO && x0 + V > nx - GHOST_COLS && x0 < nx) {
        T *oe = static_cast<T *>(p.out_e) + (y + 1);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int col = (nx - 1) - (x0 + e);      // my column nx-1-col is column `col` of the east neighbour's west ghost
            if (col >= 0 && col < GHOST_COLS && x0 + e >= 0) {
#pragma unroll
                for (int j = 0; j < 9; ++j) oe[(col * 9 + j) * gs] = q[j].v[e];
            }
        }
    }
}

// ---- the fused kernel -----------------------------------------------------------------
template <typename T, int V, int MATH, int WX, int WY, int R, int MINB, int LDP, int STP, int MODEL = MODEL_D2Q9>
__global__ void __launch_bounds__(32 * WX * WY, MINB) fused_step_kernel(const StepParams p)
{
    constexpr int SPAN = 32 * V;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wx = warp % WX, wy = warp / WX;

    // tile decode.  Single-GPU launches use a (tiles_x, tiles_y) grid: no integer division.  Halo
    // launches use a 1-D grid in which the two edge tile columns come first, so the neighbours'
    // ghost data is published as early as possible in the step.
    int bx, by;
    int rows = R;                                     // rows each warp walks
    if (p.edge_first) {
        // edge tiles are TALL (edge_rows rows per warp): few CTAs take part in the hand-shake
        const int b = blockIdx.x;
        const int n_edge = (p.tiles_x < 2 ? 1 : 2) * p.edge_tiles_y;
        if (b < n_edge) {
            rows = p.edge_rows;
            if (p.tiles_x < 2) { bx = 0; by = b; }
            else { bx = (b & 1) ? p.tiles_x - 1 : 0; by = b >> 1; }
        } else {
            const int r = b - n_edge;
            by = r / (p.tiles_x - 2);
            bx = 1 + (r - by * (p.tiles_x - 2));
        }
    } else {
        bx = blockIdx.x;
        by = blockIdx.z * gridDim.y + blockIdx.y;
    }
    const bool halo_w = (p.west == EDGE_HALO) && (bx == 0);
    const bool halo_e = (p.east == EDGE_HALO) && (bx == p.tiles_x - 1 || (p.east_pair && bx == p.tiles_x - 2));
    if (halo_w && !wait_flag(p.flag_w_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;
    if (halo_e && !wait_flag(p.flag_e_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;

    const int span0 = (bx * WX + wx) * SPAN;          // first cell of this warp's span
    const int x0 = span0 + lane * V;                  // first cell of this thread
    const int ybase = p.y_begin + (by * WY + wy) * rows;
    const bool warp_active = span0 < p.pitch;         // warp-uniform (pitch is a multiple of SPAN)

    const T *__restrict__ src = static_cast<const T *>(p.src);
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const long long plane = p.plane;
    const int ny = p.ny, pitch = p.pitch;
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);

    if (warp_active) {
        for (int r = 0; r < rows; ++r) {
            const int y = ybase + r;
            if (y >= p.y_end) break;                  // warp-uniform (also guards by >= tiles_y)
            int ym = y - 1, yp = y + 1;               // source rows of the cy=+1 / cy=-1 populations
            if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
            const long long rc = (long long)y * pitch + x0;
            const long long rm = (long long)ym * pitch + x0;
            const long long rp = (long long)yp * pitch + x0;

            // --- 9 aligned vector loads (guard rows make y=-1 / y=ny addresses valid).
            //     Three row bases + uniform multiples of the plane stride: one 64-bit add per address.
            const T *pc = src + rc;                   // planes 0,1,3 read row y
            const T *pm = src + rm + 2 * plane;       // planes 2,5,6 read row y-1
            const T *pp = src + rp + 4 * plane;       // planes 4,7,8 read row y+1
            Pack<T, V> q[9];
            q[0] = load_pack<T, V, LDP>(pc);
            q[1] = load_pack<T, V, LDP>(pc + plane);
            q[3] = load_pack<T, V, LDP>(pc + 3 * plane);
            q[2] = load_pack<T, V, LDP>(pm);
            q[5] = load_pack<T, V, LDP>(pm + 3 * plane);
            q[6] = load_pack<T, V, LDP>(pm + 4 * plane);
            q[4] = load_pack<T, V, LDP>(pp);
            q[7] = load_pack<T, V, LDP>(pp + 3 * plane);
            q[8] = load_pack<T, V, LDP>(pp + 4 * plane);
            // --- elements that cross the warp boundary (same 32-B sector the neighbour warp loads) ---
            //     written as predicated loads off the vector loads' own addresses: no branch, no new address math
            const T l1 = ld_if<T>(pc + plane - 1, lane == 0);
            const T l5 = ld_if<T>(pm + 3 * plane - 1, lane == 0);
            const T l8 = ld_if<T>(pp + 4 * plane - 1, lane == 0);
            const T r3 = ld_if<T>(pc + 3 * plane + V, lane == 31);
            const T r6 = ld_if<T>(pm + 4 * plane + V, lane == 31);
            const T r7 = ld_if<T>(pp + 3 * plane + V, lane == 31);
            // --- elements that cross the thread boundary ---
            const T s1 = __shfl_up_sync(0xffffffffu, q[1].v[V - 1], 1);
            const T s5 = __shfl_up_sync(0xffffffffu, q[5].v[V - 1], 1);
            const T s8 = __shfl_up_sync(0xffffffffu, q[8].v[V - 1], 1);
            const T s3 = __shfl_down_sync(0xffffffffu, q[3].v[0], 1);
            const T s6 = __shfl_down_sync(0xffffffffu, q[6].v[0], 1);
            const T s7 = __shfl_down_sync(0xffffffffu, q[7].v[0], 1);
            // shift: populations moving +x take the value of the cell to their left, and vice versa
#pragma unroll
            for (int e = V - 1; e > 0; --e) {
                q[1].v[e] = q[1].v[e - 1]; q[5].v[e] = q[5].v[e - 1]; q[8].v[e] = q[8].v[e - 1];
            }
            q[1].v[0] = lane == 0 ? l1 : s1; q[5].v[0] = lane == 0 ? l5 : s5; q[8].v[0] = lane == 0 ? l8 : s8;
#pragma unroll
            for (int e = 0; e < V - 1; ++e) {
                q[3].v[e] = q[3].v[e + 1]; q[6].v[e] = q[6].v[e + 1]; q[7].v[e] = q[7].v[e + 1];
            }
            q[3].v[V - 1] = lane == 31 ? r3 : s3; q[6].v[V - 1] = lane == 31 ? r6 : s6; q[7].v[V - 1] = lane == 31 ? r7 : s7;

            finish_row<T, V, MATH, STP, MODEL>(p, c, q, src, dst, x0, span0, y, ym, yp);
        }
    }

    // --- last edge tile of each side releases the neighbour for its next step.  The barrier orders
    //     every thread's peer stores before thread 0's system-scope fence (cumulativity), so one
    //     thread fences for the CTA -- the cooperative-groups grid-sync pattern.
    if (halo_w || halo_e) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (halo_w) {
                const unsigned int old = atomicAdd(p.done_w, 1u);
                if (old == (unsigned int)p.edge_tiles_y - 1u) {
                    __threadfence_system();      // acquire side: the other edge tiles' peer stores precede the flag
                    *p.done_w = 0u;
                    st_release_sys(p.flag_w_remote, p.step_id + 1u);
                }
            }
            if (halo_e) {
                const unsigned int old = atomicAdd(p.done_e, 1u);
                if (old == (unsigned int)p.done_e_count - 1u) {
                    __threadfence_system();
                    *p.done_e = 0u;
                    st_release_sys(p.flag_e_remote, p.step_id + 1u);
                }
            }
        }
    }
}

}  // namespace lb
