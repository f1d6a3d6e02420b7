// The marching kernel, overlapped strips: TWO (fused_march_kernel) or K = 3 (fused_march_k_kernel, at the end of this
// file) lattice updates per pass through HBM, one warp per column strip, the intermediate time levels in registers or
// in thread-private shared-memory slots, no barrier, no scalar gathers.  The description below is the two-update form.
//
// A warp LOADS SPAN = 32*V consecutive columns (128 fp32 / 64 fp64 cells) of every row it walks over and
// STORES the middle OUT = 30*V of them (120 / 60): strip k loads columns [k*OUT - V, k*OUT + 31*V) and owns
// [k*OUT, (k+1)*OUT).  Lanes 0 and 31 only compute: they carry the overlap with the neighbouring strips.
// For every row y the warp
//   phase 1  pulls the nine populations of row y at level t (aligned 128-bit loads, shuffles for the x-shifts),
//            applies closure / bounce-back / collision (finish_row<ROW_REGS_NOFIX>): level t+1, in registers.
//            Level t+1 is exact on every loaded column except the outermost one on each side, whose pull
//            would need a column that was not loaded;
//   phase 2  assembles the pull for row y-1 at level t+1 from what it holds -- populations 0,1,3 of row y-1
//            (kept one iteration), 2,5,6 of row y-2 (kept two iterations), 4,7,8 of row y (just computed) --
//            collides again (finish_row<ROW_FROM_TILE>) and lanes 1..30 store level t+2 of row y-1.  Level t+2
//            is exact on all loaded columns but the outer two on each side, so the V-wide overlap lanes are
//            enough (for up to V updates per pass).
// The intermediate level lives in 36 registers per thread and never touches shared or global memory: DRAM sees 9
// loads + 9 stores per TWO updates (36 B / 72 B per update, fp32 / fp64) plus the overlap columns, 2 lanes in 32,
// which neighbouring strips -- adjacent warps of one CTA, marching in step -- fetch through L2.  Everything a
// strip needs it loads itself with full-width vector loads in the row it is working on: no value is fetched
// ahead of time, so nothing is evicted from L2 between its two uses (the first version of this kernel gathered
// the columns beside each 128-wide strip with scalar loads 16 rows ahead: +32 % DRAM reads, ncu, profiles/).
// Stores are 480 B per row and population: whole 32-byte sectors.
//
// Slab edges (multi-GPU).  A strip next to a halo edge finds the neighbour slab's two
// outermost columns in its overlap lane: before the x-shift the lane's loaded vectors are patched, as if memory
// continued beyond the slab, with what the neighbour published (lb_fused.cuh, StepParams: all nine populations of
// its GHOST_COLS outermost columns, of which a launch K updates deep reads K); the obstacle bits of those columns
// come from the exchanged mask columns and the closure uses global coordinates.  The strip thus advances the
// neighbour's outermost columns to the intermediate levels exactly as the neighbour does: the decomposition stays
// bit-neutral with one exchange per launch.  Every strip that reads ghost columns or publishes some is an edge strip
// (lb_march_edge_strips in lb_host.h: also the strip before a last one narrower than GHOST_COLS columns).
// A single-slab periodic box wraps the overlap lanes' load addresses instead (nx a multiple of V).
//
// Bit-identical to the one-update path in both math modes: both phases call the same per-node code on the same
// values in the same order.
#pragma once
#include "lb_fused.cuh"

namespace lb {

template <typename T, int V>
__device__ __forceinline__ void lap_shift_x(Pack<T, V> (&q)[9])
{
    // populations moving +x take the value of the cell to their left, and vice versa; what lane 0 / lane 31
    // would need from beyond the strip is not loaded: their outermost column is overlap nobody reads
    const T s1 = __shfl_up_sync(0xffffffffu, q[1].v[V - 1], 1);
    const T s5 = __shfl_up_sync(0xffffffffu, q[5].v[V - 1], 1);
    const T s8 = __shfl_up_sync(0xffffffffu, q[8].v[V - 1], 1);
    const T s3 = __shfl_down_sync(0xffffffffu, q[3].v[0], 1);
    const T s6 = __shfl_down_sync(0xffffffffu, q[6].v[0], 1);
    const T s7 = __shfl_down_sync(0xffffffffu, q[7].v[0], 1);
#pragma unroll
    for (int e = V - 1; e > 0; --e) {
        q[1].v[e] = q[1].v[e - 1]; q[5].v[e] = q[5].v[e - 1]; q[8].v[e] = q[8].v[e - 1];
    }
    q[1].v[0] = s1; q[5].v[0] = s5; q[8].v[0] = s8;
#pragma unroll
    for (int e = 0; e < V - 1; ++e) {
        q[3].v[e] = q[3].v[e + 1]; q[6].v[e] = q[6].v[e + 1]; q[7].v[e] = q[7].v[e + 1];
    }
    q[3].v[V - 1] = s3; q[6].v[V - 1] = s6; q[7].v[V - 1] = s7;
}

// "Memory continues beyond the slab": the elements of this thread that are columns of the neighbour slab take the
// neighbour's published level-t values, every population from the row it streams from.  `e0` is the element index of
// the neighbour's boundary column (ghost column 0), `dir` = -1 towards lower elements (west edge) / +1 (east edge).
template <typename T, int V, int NCOLS>
__device__ __forceinline__ void lap_patch_ghost(Pack<T, V> (&q)[9], const T *__restrict__ G, int gs, int e0, int dir, int gy, int ym, int yp)
{
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int c = (e - e0) * dir;                  // ghost column held by element e
        if (c >= 0 && c < NCOLS) {
            const T *g = G + (long long)c * 9 * gs;
            q[0].v[e] = __ldcv(g + 0 * gs + gy + 1); q[1].v[e] = __ldcv(g + 1 * gs + gy + 1); q[3].v[e] = __ldcv(g + 3 * gs + gy + 1);
            q[2].v[e] = __ldcv(g + 2 * gs + ym + 1); q[5].v[e] = __ldcv(g + 5 * gs + ym + 1); q[6].v[e] = __ldcv(g + 6 * gs + ym + 1);
            q[4].v[e] = __ldcv(g + 4 * gs + yp + 1); q[7].v[e] = __ldcv(g + 7 * gs + yp + 1); q[8].v[e] = __ldcv(g + 8 * gs + yp + 1);
        }
    }
}

// one bit per node of the thread: is (x0 + e, y) solid?  Columns beyond a halo edge take the neighbour's mask
// column; the mask rows are zero-padded up to the pitch.
template <int V>
__device__ __forceinline__ uint32_t lap_solid_bits(const StepParams &p, int x0, int xl, int y)
{
    uint32_t bits = 0;
    if (p.mask == nullptr) return 0u;
    if (xl >= 0 && xl < p.pitch) {
        const uint8_t *mrow = p.mask + (long long)y * p.mask_pitch + xl;
        uint32_t m;
        if (V == 4) m = *reinterpret_cast<const uint32_t *>(mrow);
        else if (V == 2) m = *reinterpret_cast<const uint16_t *>(mrow);
        else m = *mrow;
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (((m >> (8 * e)) & 0xffu) == 1u && xl + e < p.nx) bits |= 1u << e;
    }
    if (p.west == EDGE_HALO && x0 < 0 && p.gmask_w != nullptr) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int c = -1 - (x0 + e);               // the neighbour's column c from its edge
            if (c >= 0 && c < GHOST_COLS - 1 && p.gmask_w[c * p.ny + y] == 1) bits |= 1u << e;
        }
    }
    if (p.east == EDGE_HALO && x0 + V > p.nx && x0 < p.nx + GHOST_COLS - 1 && p.gmask_e != nullptr) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int c = (x0 + e) - p.nx;
            if (c >= 0 && c < GHOST_COLS - 1 && p.gmask_e[c * p.ny + y] == 1) bits |= 1u << e;
        }
    }
    return bits;
}

// grid: 1-D.  CTA = NW warps = NW (strip, segment) work items, mostly adjacent strips of one segment of rows.
// nine aligned vector loads of one row: plane j at the row its population streams from
template <typename T, int V>
__device__ __forceinline__ void lap_load_row(Pack<T, V> (&q)[9], const T *__restrict__ src, long long plane, int pitch, int xl,
                                             int gy, int ym, int yp)
{
    const T *pc = src + (long long)gy * pitch + xl;
    const T *pm = src + (long long)ym * pitch + xl + 2 * plane;
    const T *pp = src + (long long)yp * pitch + xl + 4 * plane;
    q[0] = load_pack<T, V, 1>(pc);
    q[1] = load_pack<T, V, 1>(pc + plane);
    q[3] = load_pack<T, V, 1>(pc + 3 * plane);
    q[2] = load_pack<T, V, 1>(pm);
    q[5] = load_pack<T, V, 1>(pm + 3 * plane);
    q[6] = load_pack<T, V, 1>(pm + 4 * plane);
    q[4] = load_pack<T, V, 1>(pp);
    q[7] = load_pack<T, V, 1>(pp + 3 * plane);
    q[8] = load_pack<T, V, 1>(pp + 4 * plane);
}

// PF: software pipelining -- the loads of row y+1 are issued before row y is worked on (36 more registers in
// flight; the warp never waits for DRAM with nothing to do).
// SH: the kept rows of the intermediate level live in shared memory instead of registers -- nine 16-byte slots per
// thread, private to it (no other thread reads them: no synchronisation, no bank conflicts) -- which brings the
// kernel under 100 registers and 20 instead of 16 warps onto an SM.
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED, bool PF = false, int ZOV = -1, bool SH = false, int MASKED = -1>
__global__ void __launch_bounds__(32 * NW, MINB) fused_march_kernel(const StepParams p)
{
    constexpr int OUT = 30 * V;
    using VT = typename VecOf<T, V>::type;
    __shared__ VT win_s[SH ? NW * 9 * 32 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // Work items are (strip, segment) pairs, one per warp, numbered so that no warp slot is left idle by a strip
    // count that is not a multiple of NW: first the strips next to a halo edge (p.edge_first = 0 .. 3 of them per
    // segment -- a last strip narrower than GHOST_COLS columns makes the one before it an edge strip too; their CTAs come first in the grid, wait for the neighbours' ghost columns and publish the new ones
    // as early as possible in the launch), then all other strips, segment after segment.
    const int nstrips = p.tiles_x, nseg = p.tiles_y, ne = p.edge_first;
    const int n_edge_ctas = (ne * nseg + NW - 1) / NW;
    const bool edge_cta = (int)blockIdx.x < n_edge_ctas;
    int strip, seg;
    bool active;
    if (edge_cta) {
        const int item = blockIdx.x * NW + warp;
        active = item < ne * nseg;
        seg = item / ne;
        const int k = item - seg * ne;
        strip = (k == 0 && p.west == EDGE_HALO) ? 0 : nstrips - (ne - k);      // west strip, then the east one(s)
    } else {
        const int nint = nstrips - ne;
        const int item = (blockIdx.x - n_edge_ctas) * NW + warp;
        active = item < nint * nseg;
        seg = nint > 0 ? item / nint : 0;
        strip = item - seg * nint + ((ne > 0 && p.west == EDGE_HALO) ? 1 : 0);
    }
    const bool halo_w = edge_cta && (p.west == EDGE_HALO);
    const bool halo_e = edge_cta && (p.east == EDGE_HALO);
    if (halo_w && !wait_flag(p.flag_w_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;
    if (halo_e && !wait_flag(p.flag_e_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;

    const int own0 = strip * OUT;                      // first column this strip stores
    const int x0 = own0 - V + lane * V;                // first column of this thread (lane 0: overlap to the west)
    const T *__restrict__ src = static_cast<const T *>(p.src);
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);
    // whole lattice: y_begin = 0, y_end = ny; lb_run_streamed launches row bands (the rows just outside are only read)
    const bool tall = seg < p.seg_tall;
    const int ys = p.y_begin + (tall ? seg * p.seg_rows : p.seg_tall * p.seg_rows + (seg - p.seg_tall) * p.seg_rows2);
    const int ye = min(ys + (tall ? p.seg_rows : p.seg_rows2), p.y_end);
    const int gs = ny + 2;

    if (active) {                                      // warp-uniform
        // where this thread's vector is loaded from: itself, or wrapped around a single-slab periodic box
        int xl = x0;
        if (p.west == EDGE_WRAP) { if (xl < 0) xl += nx; else if (xl >= nx) xl -= nx; }
        const bool store_ok = lane != 0 && lane != 31;
        // does this thread hold columns of a neighbour slab?  (-2, -1 beyond the west edge; nx, nx+1 beyond the east)
        const bool ghost_w = (p.west == EDGE_HALO) && x0 < 0;
        const bool ghost_e = (p.east == EDGE_HALO) && x0 + V > nx && x0 <= nx + 1;
        // level t+1 kept between iterations: 0,1,3 of the previous row; 2,5,6 of the previous two rows
        Pack<T, V> h0, h1, h3, a2, a5, a6, b2, b5, b6;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            h0.v[e] = h1.v[e] = h3.v[e] = a2.v[e] = a5.v[e] = a6.v[e] = b2.v[e] = b5.v[e] = b6.v[e] = (T)0;
        }
        // SH: slot k of this thread; 0-2 = populations 0,1,3 of the previous row, 3-5 / 6-8 = populations 2,5,6 of
        // the rows of even / odd index
        VT *win = win_s + (SH ? warp * 9 * 32 + lane : 0);
        if (SH) {
#pragma unroll
            for (int k = 0; k < 9; ++k) win[k * 32] = repack(h0);
        }
        uint32_t solid_prev = 0u;                      // obstacle bits of the row phase 2 works on
        // rows wrap on a periodic box; outside a pipe there is nothing to load
        auto row_of = [&](int y, int &gy, int &ym, int &yp) -> bool {
            gy = y;
            if (periodic) { if (gy < 0) gy = ny - 1; if (gy == ny) gy = 0; }
            else if (gy < 0 || gy >= ny) return false;
            ym = gy - 1; yp = gy + 1;
            if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
            return true;
        };
        Pack<T, V> qn[9];                              // PF: the row after this one, in flight
        if (PF) {
            int gy, ym, yp;
            if (row_of(ys - 1, gy, ym, yp)) lap_load_row<T, V>(qn, src, plane, pitch, xl, gy, ym, yp);
        }
        for (int y = ys - 1; y <= ye; ++y) {
            // ---- phase 1: level t+1 of row y ----------------------------------------------------------
            int gy, ym = 0, yp = 0;
            const bool valid = row_of(y, gy, ym, yp);
            Pack<T, V> q[9];
            uint32_t solid_now = 0u;
            if (PF) {
#pragma unroll
                for (int j = 0; j < 9; ++j) q[j] = qn[j];
                int g2, m2, p2;
                if (y < ye && row_of(y + 1, g2, m2, p2)) lap_load_row<T, V>(qn, src, plane, pitch, xl, g2, m2, p2);
            }
            if (valid) {
                if (!PF) lap_load_row<T, V>(q, src, plane, pitch, xl, gy, ym, yp);
                if (MASKED != 0) solid_now = lap_solid_bits<V>(p, x0, xl, gy);
                if (ghost_w) lap_patch_ghost<T, V, 2>(q, static_cast<const T *>(p.ghost_w), gs, -1 - x0, -1, gy, ym, yp);
                if (ghost_e) lap_patch_ghost<T, V, 2>(q, static_cast<const T *>(p.ghost_e), gs, nx - x0, +1, gy, ym, yp);
                lap_shift_x<T, V>(q);
                finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_REGS_NOFIX, PACKED, true, ZOV, MASKED>(p, c, q, src, dst, x0, own0, gy, ym, yp, solid_now, false);
            } else {
#pragma unroll
                for (int j = 0; j < 9; ++j)
#pragma unroll
                    for (int e = 0; e < V; ++e) q[j].v[e] = (T)0;      // outside the pipe: never reaches a result
            }
            // ---- phase 2: level t+2 of row y-1 -----------------------------------------------------------
            if (y > ys) {
                const int r = y - 1;
                int rm = r - 1, rp = r + 1;
                if (periodic) { if (rm < 0) rm = ny - 1; if (rp >= ny) rp = 0; }
                Pack<T, V> z[9];
                if (SH) {
                    const int o = 3 + 3 * (y & 1);     // the slot of row y-2, which row y replaces below
                    unpack(win[0 * 32], z[0]); unpack(win[1 * 32], z[1]); unpack(win[2 * 32], z[3]);
                    unpack(win[(o + 0) * 32], z[2]); unpack(win[(o + 1) * 32], z[5]); unpack(win[(o + 2) * 32], z[6]);
                } else {
                    z[0] = h0; z[1] = h1; z[3] = h3;
                    z[2] = b2; z[5] = b5; z[6] = b6;
                }
                z[4] = q[4]; z[7] = q[7]; z[8] = q[8];
                if (SH) {                              // row y takes the slots just read (before z is worked on)
                    const int o = 3 + 3 * (y & 1);
                    win[0 * 32] = repack(q[0]); win[1 * 32] = repack(q[1]); win[2 * 32] = repack(q[3]);
                    win[(o + 0) * 32] = repack(q[2]); win[(o + 1) * 32] = repack(q[5]); win[(o + 2) * 32] = repack(q[6]);
                }
                lap_shift_x<T, V>(z);
                finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_FROM_TILE, PACKED, true, ZOV, MASKED>(p, c, z, src, dst, x0, own0, r, rm, rp, solid_prev, store_ok);
            }
            // ---- rotate the kept rows --------------------------------------------------------------------
            if (SH) {
                if (y <= ys) {                         // no phase 2 yet: nothing was read, just park the row
                    const int o = 3 + 3 * (y & 1);
                    win[0 * 32] = repack(q[0]); win[1 * 32] = repack(q[1]); win[2 * 32] = repack(q[3]);
                    win[(o + 0) * 32] = repack(q[2]); win[(o + 1) * 32] = repack(q[5]); win[(o + 2) * 32] = repack(q[6]);
                }
            } else {
                b2 = a2; b5 = a5; b6 = a6;
                a2 = q[2]; a5 = q[5]; a6 = q[6];
                h0 = q[0]; h1 = q[1]; h3 = q[3];
            }
            solid_prev = solid_now;
        }
    }

    // --- release the neighbours for their next launch (same hand-shake as fused_step_kernel) ---
    if (halo_w || halo_e) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (halo_w) {
                const unsigned int old = atomicAdd(p.done_w, 1u);
                if (old == (unsigned int)p.edge_tiles_y - 1u) {
                    __threadfence_system();
                    *p.done_w = 0u;
                    st_release_sys(p.flag_w_remote, p.step_id + 1u);
                }
            }
            if (halo_e) {
                const unsigned int old = atomicAdd(p.done_e, 1u);
                if (old == (unsigned int)p.edge_tiles_y - 1u) {
                    __threadfence_system();
                    *p.done_e = 0u;
                    st_release_sys(p.flag_e_remote, p.step_id + 1u);
                }
            }
        }
    }
}

// ---- K updates per pass ------------------------------------------------------------------------------------------
// The same march with K time levels in flight (K = 2 or 3; the V-wide overlap lanes carry enough columns for K <= V):
// at iteration y level 1 works on row y, level 2 on row y-1, ..., level K on row y-K+1, which is stored.  Level l takes
// populations 0,1,3 of its row and 2,5,6 of the row above from the kept rows of level l-1 (nine thread-private
// shared-memory slots per level), and 4,7,8 from what level l-1 produced a moment ago in this very iteration.  DRAM
// sees 9 loads + 9 stores per K updates; a segment of S rows costs S + 2(K-1) level-1 rows.  Slab edges: the
// neighbour's K outermost columns are patched into the overlap lane (they are all published, lb_fused.cuh
// StepParams), and levels 1 .. K-1 of them are advanced here exactly as the neighbour advances them.
// OVL = overlap lanes per side: a strip stores (32 - 2 OVL) V columns and has OVL V columns of overlap on each side,
// enough for K <= OVL V levels (fp32: one lane of four columns; fp64 with three levels: two lanes of two).
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED, int K, int ZOV, int MASKED, int OVL = 1>
__global__ void __launch_bounds__(32 * NW, MINB) fused_march_k_kernel(const StepParams p)
{
    static_assert(K >= 2 && K <= OVL * V && K <= GHOST_COLS, "overlap lanes carry OVL*V columns, ghost arenas GHOST_COLS");
    constexpr int OUT = (32 - 2 * OVL) * V;
    using VT = typename VecOf<T, V>::type;
    __shared__ VT win_s[NW * (K - 1) * 9 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const int nstrips = p.tiles_x, nseg = p.tiles_y, ne = p.edge_first;     // work items: see fused_march_kernel
    const int n_edge_ctas = (ne * nseg + NW - 1) / NW;
    const bool edge_cta = (int)blockIdx.x < n_edge_ctas;
    int strip, seg;
    bool active;
    if (edge_cta) {
        const int item = blockIdx.x * NW + warp;
        active = item < ne * nseg;
        seg = item / ne;
        const int k = item - seg * ne;
        strip = (k == 0 && p.west == EDGE_HALO) ? 0 : nstrips - (ne - k);      // west strip, then the east one(s)
    } else {
        const int nint = nstrips - ne;
        const int item = (blockIdx.x - n_edge_ctas) * NW + warp;
        active = item < nint * nseg;
        seg = nint > 0 ? item / nint : 0;
        strip = item - seg * nint + ((ne > 0 && p.west == EDGE_HALO) ? 1 : 0);
    }
    const bool halo_w = edge_cta && (p.west == EDGE_HALO);
    const bool halo_e = edge_cta && (p.east == EDGE_HALO);
    if (halo_w && !wait_flag(p.flag_w_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;
    if (halo_e && !wait_flag(p.flag_e_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;

    const int own0 = strip * OUT;
    const int x0 = own0 - OVL * V + lane * V;
    const T *__restrict__ src = static_cast<const T *>(p.src);
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);
    const bool tall = seg < p.seg_tall;
    const int ys = p.y_begin + (tall ? seg * p.seg_rows : p.seg_tall * p.seg_rows + (seg - p.seg_tall) * p.seg_rows2);
    const int ye = min(ys + (tall ? p.seg_rows : p.seg_rows2), p.y_end);
    const int gs = ny + 2;

    if (active) {
        int xl = x0;
        if (p.west == EDGE_WRAP) { if (xl < 0) xl += nx; else if (xl >= nx) xl -= nx; }
        const bool store_ok = lane >= OVL && lane < 32 - OVL;
        const bool ghost_w = (p.west == EDGE_HALO) && x0 < 0;
        const bool ghost_e = (p.east == EDGE_HALO) && x0 + V > nx && x0 < nx + K;
        // slots of level l (1 .. K-1): 0-2 = populations 0,1,3 of its latest row, 3-5 / 6-8 = populations 2,5,6 of its
        // rows of even / odd index
        VT *win = win_s + warp * (K - 1) * 9 * 32 + lane;
        {
            Pack<T, V> zero;
#pragma unroll
            for (int e = 0; e < V; ++e) zero.v[e] = (T)0;
#pragma unroll
            for (int k = 0; k < (K - 1) * 9; ++k) win[k * 32] = repack(zero);
        }
        // rows wrap on a periodic box (by up to K-1 rows beyond either end); outside a pipe there is nothing
        auto row_of = [&](int y, int &gy, int &ym, int &yp) -> bool {
            gy = y;
            if (periodic) { if (gy < 0) gy += ny; if (gy >= ny) gy -= ny; }
            else if (gy < 0 || gy >= ny) return false;
            ym = gy - 1; yp = gy + 1;
            if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
            return true;
        };
        uint32_t solid[K];                             // solid[l-1]: obstacle bits of the row level l works on
#pragma unroll
        for (int l = 0; l < K; ++l) solid[l] = 0u;
        for (int y = ys - (K - 1); y <= ye + K - 2; ++y) {
            // ---- level 1: row y, from global memory -------------------------------------------------------
            Pack<T, V> cur[9];
            {
                int gy, ym = 0, yp = 0;
                if (row_of(y, gy, ym, yp)) {
                    lap_load_row<T, V>(cur, src, plane, pitch, xl, gy, ym, yp);
                    if (MASKED != 0) solid[0] = lap_solid_bits<V>(p, x0, xl, gy);
                    if (ghost_w) lap_patch_ghost<T, V, K>(cur, static_cast<const T *>(p.ghost_w), gs, -1 - x0, -1, gy, ym, yp);
                    if (ghost_e) lap_patch_ghost<T, V, K>(cur, static_cast<const T *>(p.ghost_e), gs, nx - x0, +1, gy, ym, yp);
                    lap_shift_x<T, V>(cur);
                    finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_REGS_NOFIX, PACKED, true, ZOV, MASKED>(p, c, cur, src, dst, x0, own0, gy, ym, yp, solid[0], false);
                } else {
                    solid[0] = 0u;
#pragma unroll
                    for (int j = 0; j < 9; ++j)
#pragma unroll
                        for (int e = 0; e < V; ++e) cur[j].v[e] = (T)0;
                }
            }
            // ---- levels 2 .. K: row y-l+1, from the kept rows of level l-1 and what it produced just now ----
#pragma unroll
            for (int l = 2; l <= K; ++l) {
                const int r = y - (l - 1);
                const bool act = y >= ys - K + 2 * l - 1;          // warp-uniform: does level l have to produce row r?
                VT *w = win + (l - 2) * 9 * 32;
                const int o = 3 + 3 * ((r + 1) & 1);               // the slot of row r-1, which row r+1 replaces below
                Pack<T, V> z[9];
                unpack(w[0 * 32], z[0]); unpack(w[1 * 32], z[1]); unpack(w[2 * 32], z[3]);
                unpack(w[(o + 0) * 32], z[2]); unpack(w[(o + 1) * 32], z[5]); unpack(w[(o + 2) * 32], z[6]);
                z[4] = cur[4]; z[7] = cur[7]; z[8] = cur[8];
                w[0 * 32] = repack(cur[0]); w[1 * 32] = repack(cur[1]); w[2 * 32] = repack(cur[3]);
                w[(o + 0) * 32] = repack(cur[2]); w[(o + 1) * 32] = repack(cur[5]); w[(o + 2) * 32] = repack(cur[6]);
                int gr = 0, rm = 0, rp = 0;
                const bool live = act && row_of(r, gr, rm, rp);
                if (l < K) {
                    if (live) {
                        lap_shift_x<T, V>(z);
                        finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_REGS_NOFIX, PACKED, true, ZOV, MASKED>(p, c, z, src, dst, x0, own0, gr, rm, rp, solid[l - 1], false);
                    }
#pragma unroll
                    for (int j = 0; j < 9; ++j) cur[j] = z[j];     // (not live: values nobody will use)
                } else if (live) {
                    lap_shift_x<T, V>(z);
                    finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_FROM_TILE, PACKED, true, ZOV, MASKED>(p, c, z, src, dst, x0, own0, gr, rm, rp, solid[K - 1], store_ok);
                }
            }
#pragma unroll
            for (int l = K - 1; l > 0; --l) solid[l] = solid[l - 1];
        }
    }

    if (halo_w || halo_e) {                            // release the neighbours for their next launch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (halo_w) {
                const unsigned int old = atomicAdd(p.done_w, 1u);
                if (old == (unsigned int)p.edge_tiles_y - 1u) {
                    __threadfence_system();
                    *p.done_w = 0u;
                    st_release_sys(p.flag_w_remote, p.step_id + 1u);
                }
            }
            if (halo_e) {
                const unsigned int old = atomicAdd(p.done_e, 1u);
                if (old == (unsigned int)p.edge_tiles_y - 1u) {
                    __threadfence_system();
                    *p.done_e = 0u;
                    st_release_sys(p.flag_e_remote, p.step_id + 1u);
                }
            }
        }
    }
}

}  // namespace lb
