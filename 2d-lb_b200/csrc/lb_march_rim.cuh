// EXPERIMENT (-DLB_EXPERIMENTS), superseded by lb_march.cuh: the first version of the marching kernel.  Same
// register-resident two-update scheme, but strips are 128 columns wide without overlap and the level-t+1 values
// beside a strip come from scalar gathers 16 rows ahead -- which ncu showed to cost +32 % DRAM reads (every
// gathered 4-byte value fetches a 64-byte granule that is evicted from L2 before the neighbouring strip loads
// it again; evict_last hints recovered a fifth of it, L2 prefetches made it worse: profiles/README.md).
//
// TWO lattice updates per pass through HBM, one warp per column strip, no shared tile and no CTA-wide barrier.
//
// A warp owns SPAN = 32*V consecutive columns (128 fp32 / 64 fp64 cells) and walks down a segment of rows.
// For every row y it
//   phase 1  pulls the nine populations of row y at time level t from global memory exactly like
//            `fused_step_kernel` (aligned 128-bit loads, shuffles for the x-shifts), applies closure /
//            bounce-back / collision (finish_row<ROW_TO_REGISTERS>): level t+1 of row y, in registers;
//   phase 2  assembles the pull for row y-1 at level t+1 from what it already holds -- populations 0,1,3 of
//            row y-1 (kept one iteration), 2,5,6 of row y-2 (kept two iterations), 4,7,8 of row y (just
//            computed) -- collides again and stores level t+2 of row y-1 (finish_row<ROW_FROM_TILE>).
// The intermediate level therefore lives in 24 registers per thread and never touches shared or global
// memory; DRAM sees 9 loads + 9 stores per TWO updates (36 B / 72 B per update, fp32 / fp64).  Warps never
// wait for each other: the latency of a row's loads is hidden by the other warps of the SM, as in the
// one-update kernel, instead of being exposed behind a tile barrier (lb_tb2v.cuh).
//
// What a strip cannot produce itself are the level-t+1 values just outside its two edges (column span0-1
// and span0+SPAN), which phase 2 pulls from.  Those "rim" nodes are computed redundantly, 16 rows at a
// time: lanes 0-15 take the west rim, lanes 16-31 the east rim, one node each, scalar path (tb2_node), and
// park the three populations phase 2 needs in a 32-row ring in shared memory (768 B per warp).  That is one
// scalar collision pass per 16 rows against 128 vector passes: < 1 % extra arithmetic.  A segment of S rows
// runs phase 1 on S+2 rows (2/S extra; S = 64 by default).
//
// Slab edges (multi-GPU): the rim node beyond a halo edge is the NEIGHBOUR's boundary
// column.  Its level-t populations come from the nine-slot ghost column the neighbour published
// (StepParams), its obstacle flag from the exchanged mask column, and its closure uses global coordinates, so
// the result is the very value the neighbour computes for that node: the decomposition stays bit-neutral,
// with one exchange per two updates.
//
// Bit-identical to the one-update path in both math modes: both phases call the same per-node code on the
// same values in the same order.
#pragma once
#include "lb_fused.cuh"
#include "lb_tb2.cuh"

namespace lb {

template <typename T, int V>
__device__ __forceinline__ void shift_x_regs(Pack<T, V> (&q)[9], int lane, T l1, T l5, T l8, T r3, T r6, T r7)
{
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
    q[1].v[0] = lane == 0 ? l1 : s1; q[5].v[0] = lane == 0 ? l5 : s5; q[8].v[0] = lane == 0 ? l8 : s8;
#pragma unroll
    for (int e = 0; e < V - 1; ++e) {
        q[3].v[e] = q[3].v[e + 1]; q[6].v[e] = q[6].v[e + 1]; q[7].v[e] = q[7].v[e + 1];
    }
    q[3].v[V - 1] = lane == 31 ? r3 : s3; q[6].v[V - 1] = lane == 31 ? r6 : s6; q[7].v[V - 1] = lane == 31 ? r7 : s7;
}

constexpr int MARCH_RIM_ROWS = 32;     // ring of rim rows per warp (batches of 16)

// Scalar load of a rim node's level-t population.  The same 32-byte sector is part of a 512-byte vector load
// of the neighbouring strip up to 16 rows (tens of microseconds, hundreds of MB of streaming traffic) later;
// RIMLD 1 asks L2 to keep it until then (evict_last) instead of fetching it from DRAM twice.
template <typename T, int RIMLD> __device__ __forceinline__ T rim_ld(const T *p);
template <> __device__ __forceinline__ float rim_ld<float, 0>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ double rim_ld<double, 0>(const double *p) { return __ldg(p); }
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <> __device__ __forceinline__ float rim_ld<float, 1>(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
template <> __device__ __forceinline__ double rim_ld<double, 1>(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
#ifndef LB_MARCH_RIMLD
#define LB_MARCH_RIMLD 0
#endif

// Level t+1 of one rim node (gx, gy), gx = span0-1 (side 0) or span0+SPAN (side 1), possibly outside this
// slab.  Returns the three populations phase 2 pulls from it: 1,5,8 (west rim) or 3,6,7 (east rim).
template <typename T, int MATH>
__device__ __forceinline__ void march_rim_node(const StepParams &p, const Consts<T> &c, const T *__restrict__ src,
                                               int side, int gx, int gy, T &o0, T &o1, T &o2)
{
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const long long plane = p.plane;
    const bool periodic = (p.bc == BC_PERIODIC);
    int ym = gy - 1, yp = gy + 1;
    bool has_ym = true, has_yp = true;
    if (periodic) { if (ym < 0) ym = ny - 1; if (yp == ny) yp = 0; }
    else { has_ym = ym >= 0; has_yp = yp < ny; }
    const long long rc = (long long)gy * pitch, rm = (long long)ym * pitch, rp = (long long)yp * pitch;
    const int gs = ny + 2;
    T g[9];
    bool solid = false;
    const bool beyond_w = gx < 0, beyond_e = gx >= nx;
    if (beyond_w && p.west == EDGE_HALO) {
        // the west neighbour's last column: its own populations from the ghost slots, 3,6,7 from my column 0
        const T *G = static_cast<const T *>(p.ghost_w);
        g[0] = __ldcv(G + 0 * gs + gy + 1);
        g[2] = __ldcv(G + 2 * gs + ym + 1);
        g[4] = __ldcv(G + 4 * gs + yp + 1);
        g[1] = __ldcv(G + (9 + 1) * gs + gy + 1);
        g[5] = __ldcv(G + (9 + 5) * gs + ym + 1);
        g[8] = __ldcv(G + (9 + 8) * gs + yp + 1);
        g[3] = __ldg(src + 3 * plane + rc);
        g[6] = __ldg(src + 6 * plane + rm);
        g[7] = __ldg(src + 7 * plane + rp);
        if (p.gmask_w != nullptr) solid = p.gmask_w[gy] == 1;
    } else if (beyond_e && p.east == EDGE_HALO) {
        const T *G = static_cast<const T *>(p.ghost_e);
        g[0] = __ldcv(G + 0 * gs + gy + 1);
        g[2] = __ldcv(G + 2 * gs + ym + 1);
        g[4] = __ldcv(G + 4 * gs + yp + 1);
        g[3] = __ldcv(G + (9 + 3) * gs + gy + 1);
        g[6] = __ldcv(G + (9 + 6) * gs + ym + 1);
        g[7] = __ldcv(G + (9 + 7) * gs + yp + 1);
        g[1] = __ldg(src + 1 * plane + rc + (nx - 1));
        g[5] = __ldg(src + 5 * plane + rm + (nx - 1));
        g[8] = __ldg(src + 8 * plane + rp + (nx - 1));
        if (p.gmask_e != nullptr) solid = p.gmask_e[gy] == 1;
    } else {
        // a node of this slab (wrapped on a single-slab periodic box)
        int x = gx;
        if (x < 0) x = nx - 1;
        if (x >= nx) x = 0;
        int xm = x - 1, xp = x + 1;
        bool has_xm = true, has_xp = true;
        if (p.west == EDGE_WRAP) { if (xm < 0) xm = nx - 1; if (xp == nx) xp = 0; }
        else { has_xm = xm >= 0; has_xp = xp < nx; }
        g[0] = rim_ld<T, LB_MARCH_RIMLD>(src + rc + x);
        g[1] = has_xm ? rim_ld<T, LB_MARCH_RIMLD>(src + 1 * plane + rc + xm) : (T)0;
        g[3] = has_xp ? rim_ld<T, LB_MARCH_RIMLD>(src + 3 * plane + rc + xp) : (T)0;
        g[2] = has_ym ? rim_ld<T, LB_MARCH_RIMLD>(src + 2 * plane + rm + x) : (T)0;
        g[4] = has_yp ? rim_ld<T, LB_MARCH_RIMLD>(src + 4 * plane + rp + x) : (T)0;
        g[5] = (has_xm && has_ym) ? rim_ld<T, LB_MARCH_RIMLD>(src + 5 * plane + rm + xm) : (T)0;
        g[6] = (has_xp && has_ym) ? rim_ld<T, LB_MARCH_RIMLD>(src + 6 * plane + rm + xp) : (T)0;
        g[7] = (has_xp && has_yp) ? rim_ld<T, LB_MARCH_RIMLD>(src + 7 * plane + rp + xp) : (T)0;
        g[8] = (has_xm && has_yp) ? rim_ld<T, LB_MARCH_RIMLD>(src + 8 * plane + rp + xm) : (T)0;
        // a rim node whose own neighbour column lies beyond a halo edge: that column's entering populations
        // are in ghost slots 0-2 (only when the strip next to the edge is one column wide; kept for safety)
        if (x == 0 && p.west == EDGE_HALO) {
            const T *G = static_cast<const T *>(p.ghost_w);
            g[1] = __ldcv(G + 1 * gs + gy + 1); g[5] = __ldcv(G + 5 * gs + ym + 1); g[8] = __ldcv(G + 8 * gs + yp + 1);
        }
        if (x == nx - 1 && p.east == EDGE_HALO) {
            const T *G = static_cast<const T *>(p.ghost_e);
            g[3] = __ldcv(G + 3 * gs + gy + 1); g[6] = __ldcv(G + 6 * gs + ym + 1); g[7] = __ldcv(G + 7 * gs + yp + 1);
        }
        if (p.mask != nullptr) solid = p.mask[(long long)gy * p.mask_pitch + x] == 1;
        gx = x;
    }
    if (!periodic) pipe_bc<T, MODEL_D2Q9>(c, p.x_off + gx, gy, p.gnx, ny, g);
    if (solid) bounce_back<T>(g);
    T rho, u, v;
    collide_node<T, MATH, MODEL_D2Q9>(c, g, rho, u, v, solid && p.zero_obstacle_velocity);
    if (side == 0) { o0 = g[1]; o1 = g[5]; o2 = g[8]; }
    else { o0 = g[3]; o1 = g[6]; o2 = g[7]; }
}

// grid: 1-D.  CTA = NW warps = NW adjacent strips of one segment.  Halo launches order the two edge CTA
// columns first (they wait for / publish the ghost columns), like fused_step_kernel.
template <typename T, int V, int MATH, int NW, int MINB, bool PACKED>
__global__ void __launch_bounds__(32 * NW, MINB) fused_march_rim_kernel(const StepParams p)
{
    constexpr int SPAN = 32 * V;
    __shared__ T rim_s[NW][2][3][MARCH_RIM_ROWS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    int bx, seg;
    if (p.edge_first) {
        const int b = blockIdx.x;
        const int n_edge = (p.tiles_x < 2 ? 1 : 2) * p.tiles_y;
        if (b < n_edge) {
            if (p.tiles_x < 2) { bx = 0; seg = b; }
            else { bx = (b & 1) ? p.tiles_x - 1 : 0; seg = b >> 1; }
        } else {
            const int r = b - n_edge;
            seg = r / (p.tiles_x - 2);
            bx = 1 + (r - seg * (p.tiles_x - 2));
        }
    } else {
        seg = blockIdx.x / p.tiles_x;
        bx = blockIdx.x - seg * p.tiles_x;
    }
    const bool halo_w = (p.west == EDGE_HALO) && (bx == 0);
    const bool halo_e = (p.east == EDGE_HALO) && (bx == p.tiles_x - 1);
    if (halo_w && !wait_flag(p.flag_w_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;
    if (halo_e && !wait_flag(p.flag_e_local, p.step_id, p.error_word, p.halo_timeout_ns)) return;

    const int span0 = (bx * NW + warp) * SPAN;
    const int x0 = span0 + lane * V;
    const T *__restrict__ src = static_cast<const T *>(p.src);
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);
    const int ys = seg * p.seg_rows;
    const int ye = min(ys + p.seg_rows, ny);

    if (span0 < pitch) {                               // warp-uniform
        T(*rim)[3][MARCH_RIM_ROWS] = rim_s[warp];
        const bool east_inside = (p.east == EDGE_HALO) && span0 < nx && span0 + SPAN > nx;
        // level t+1 kept between iterations: 0,1,3 of the previous row; 2,5,6 of the previous two rows
        Pack<T, V> h0, h1, h3, a2, a5, a6, b2, b5, b6;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            h0.v[e] = h1.v[e] = h3.v[e] = a2.v[e] = a5.v[e] = a6.v[e] = b2.v[e] = b5.v[e] = b6.v[e] = (T)0;
        }
        for (int y = ys - 1; y <= ye; ++y) {
            // ---- rim nodes of rows y .. y+15, one lane each -----------------------------------------
            if (((y - (ys - 1)) & 15) == 0) {
                __syncwarp();                          // the ring slots about to be rewritten were read 16+ rows ago
                const int side = lane >> 4, k = lane & 15;
                const int ry = y + k;
                int gy = ry;
                int gx = side ? min(span0 + SPAN, nx) : span0 - 1;   // a slab's last strip may end before its rim: the
                                                                       // node beyond column nx-1 is what phase 2 needs
                bool live = ry <= ye && gx <= nx;
                if (periodic) { if (gy < 0) gy = ny - 1; if (gy == ny) gy = 0; }
                else if (gy < 0 || gy >= ny) live = false;
                if (gx < 0 && p.west == EDGE_BOUNDARY) live = false;       // beyond the inlet: never pulled from
                if (gx >= nx && p.east == EDGE_BOUNDARY) live = false;     // beyond the outlet (or in the pitch padding)
                if (gx > nx) live = false;
                T o0 = (T)0, o1 = (T)0, o2 = (T)0;
                if (live) march_rim_node<T, MATH>(p, c, src, side, gx, gy, o0, o1, o2);
                const int slot = (ry - (ys - 1)) & (MARCH_RIM_ROWS - 1);
                rim[side][0][slot] = o0; rim[side][1][slot] = o1; rim[side][2][slot] = o2;
                __syncwarp();
            }
            // ---- phase 1: level t+1 of row y ----------------------------------------------------------
            int gy = y;
            bool valid = true;
            if (periodic) { if (gy < 0) gy = ny - 1; if (gy == ny) gy = 0; }
            else valid = (gy >= 0 && gy < ny);
            Pack<T, V> q[9];
            if (valid) {
                int ym = gy - 1, yp = gy + 1;
                if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
                const T *pc = src + (long long)gy * pitch + x0;
                const T *pm = src + (long long)ym * pitch + x0 + 2 * plane;
                const T *pp = src + (long long)yp * pitch + x0 + 4 * plane;
                q[0] = load_pack<T, V, 1>(pc);
                q[1] = load_pack<T, V, 1>(pc + plane);
                q[3] = load_pack<T, V, 1>(pc + 3 * plane);
                q[2] = load_pack<T, V, 1>(pm);
                q[5] = load_pack<T, V, 1>(pm + 3 * plane);
                q[6] = load_pack<T, V, 1>(pm + 4 * plane);
                q[4] = load_pack<T, V, 1>(pp);
                q[7] = load_pack<T, V, 1>(pp + 3 * plane);
                q[8] = load_pack<T, V, 1>(pp + 4 * plane);
                const T l1 = ld_if<T>(pc + plane - 1, lane == 0);
                const T l5 = ld_if<T>(pm + 3 * plane - 1, lane == 0);
                const T l8 = ld_if<T>(pp + 4 * plane - 1, lane == 0);
                const T r3 = ld_if<T>(pc + 3 * plane + V, lane == 31);
                const T r6 = ld_if<T>(pm + 4 * plane + V, lane == 31);
                const T r7 = ld_if<T>(pp + 3 * plane + V, lane == 31);
                shift_x_regs<T, V>(q, lane, l1, l5, l8, r3, r6, r7);
                finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_TO_REGISTERS, PACKED>(p, c, q, src, dst, x0, span0, gy, ym, yp);
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
                z[0] = h0; z[1] = h1; z[3] = h3;
                z[2] = b2; z[5] = b5; z[6] = b6;
                z[4] = q[4]; z[7] = q[7]; z[8] = q[8];
                const int s_c = (r - (ys - 1)) & (MARCH_RIM_ROWS - 1);
                const int s_m = (r - 1 - (ys - 1)) & (MARCH_RIM_ROWS - 1);
                const int s_p = (r + 1 - (ys - 1)) & (MARCH_RIM_ROWS - 1);
                T l1 = (T)0, l5 = (T)0, l8 = (T)0, r3 = (T)0, r6 = (T)0, r7 = (T)0;
                if (lane == 0) { l1 = rim[0][0][s_c]; l5 = rim[0][1][s_m]; l8 = rim[0][2][s_p]; }
                if (lane == 31) { r3 = rim[1][0][s_c]; r6 = rim[1][1][s_m]; r7 = rim[1][2][s_p]; }
                shift_x_regs<T, V>(z, lane, l1, l5, l8, r3, r6, r7);
                if (east_inside) {                     // warp-uniform: the neighbour slab's column sits inside this strip
                    const int el = (nx - 1) - x0;      // the thread that owns column nx-1 pulls from it
                    if (el >= 0 && el < V) {
                        const T e3 = rim[1][0][s_c], e6 = rim[1][1][s_m], e7 = rim[1][2][s_p];
#pragma unroll
                        for (int e = 0; e < V; ++e)
                            if (e == el) { z[3].v[e] = e3; z[6].v[e] = e6; z[7].v[e] = e7; }
                    }
                }
                finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_FROM_TILE, PACKED>(p, c, z, src, dst, x0, span0, r, rm, rp);
            }
            // ---- rotate the kept rows --------------------------------------------------------------------
            b2 = a2; b5 = a5; b6 = a6;
            a2 = q[2]; a5 = q[5]; a6 = q[6];
            h0 = q[0]; h1 = q[1]; h3 = q[3];
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

}  // namespace lb
