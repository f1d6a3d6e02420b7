// Temporal blocking, second version: the same two-updates-per-HBM-pass scheme as lb_tb2.cuh with the
// one-step kernel's own machinery in both phases -- a warp owns whole rows of the tile (SPAN = 32*V
// cells: 128 fp32 / 64 fp64), loads nine aligned 128-bit vectors per thread, resolves the +-1 x shifts
// with shuffles, and runs `finish_row` (merged boundary branch, per-32-cell obstacle flags, flag-free
// collide loop).  Against the first version this removes three quarters of the load/store instructions
// and all per-cell 64-bit address arithmetic, and the only scalar work left is the two rim columns of
// the grown tile (2 x (BY+2) cells, one lane each).
//
//   phase 1  rows y0-1 .. y0+BY of the tile (BY+2 rows, warp-strided): pull from global memory exactly like
//            fused_step_kernel, finish_row<ROW_TO_REGISTERS>, nine 128-bit stores to shared memory; then the
//            rim columns x0-1 and x0+SPAN through the scalar path of lb_tb2.cuh (tb2_node);
//   phase 2  rows y0 .. y0+BY-1: nine 128-bit loads from shared memory (rows ey-1 / ey / ey+1 of the block),
//            shuffles, the two warp-edge elements from the rim columns, finish_row<ROW_FROM_TILE> -> global.
// Shared-memory row: [OFF-1] west rim | [OFF .. OFF+SPAN) tile | [OFF+SPAN] east rim, OFF = 16 B so that the
// tile part is 16-byte aligned; row pitch SPAN + 2*OFF elements.
#pragma once
#include "lb_fused.cuh"
#include "lb_tb2.cuh"

namespace lb {

template <typename T, int V, int BY> struct Tb2vTile {
    static constexpr int SPAN = 32 * V;
    static constexpr int OFF = 16 / (int)sizeof(T);
    static constexpr int SP = SPAN + 2 * OFF;          // row pitch (elements)
    static constexpr int EY = BY + 2;
    static constexpr int PS = SP * EY;                 // plane stride (elements)
    static constexpr size_t BYTES = (size_t)9 * PS * sizeof(T);
};

template <typename T, int V>
__device__ __forceinline__ Pack<T, V> lds_pack(const T *p)
{
    using VT = typename VecOf<T, V>::type;
    Pack<T, V> r;
    unpack(*reinterpret_cast<const VT *>(p), r);
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void sts_pack(T *p, const Pack<T, V> &r)
{
    using VT = typename VecOf<T, V>::type;
    *reinterpret_cast<VT *>(p) = repack(r);
}

// shift the x-moving populations by one cell: lanes exchange their edge elements, lane 0 / 31 take l* / r*
template <typename T, int V>
__device__ __forceinline__ void shift_x(Pack<T, V> (&q)[9], int lane, T l1, T l5, T l8, T r3, T r6, T r7)
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

template <typename T, int V, int MATH, int BY, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) fused_two_step_v2_kernel(const StepParams p)
{
    using TL = Tb2vTile<T, V, BY>;
    extern __shared__ __align__(16) unsigned char tb2v_smem[];
    T *s = reinterpret_cast<T *>(tb2v_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span0 = blockIdx.x * TL::SPAN;           // first column of the tile (pitch is a multiple of SPAN)
    const int x0 = span0 + lane * V;
    const int y0 = blockIdx.y * BY;
    const T *__restrict__ src = static_cast<const T *>(p.src);
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const long long plane = p.plane;
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);

    // ---- phase 1, tile columns: time level t+1 of rows y0-1 .. y0+BY ---------------------------------
    for (int ey = warp; ey < TL::EY; ey += NW) {
        int gy = y0 - 1 + ey;
        if (gy > ny) break;                            // warp-uniform; rows beyond the rim of the last tile
        if (periodic) { if (gy < 0) gy = ny - 1; if (gy == ny) gy = 0; }
        else if (gy < 0 || gy == ny) continue;         // outside the pipe: never pulled from
        int ym = gy - 1, yp = gy + 1;
        if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
        const T *pc = src + (long long)gy * pitch + x0;
        const T *pm = src + (long long)ym * pitch + x0 + 2 * plane;
        const T *pp = src + (long long)yp * pitch + x0 + 4 * plane;
        Pack<T, V> q[9];
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
        shift_x<T, V>(q, lane, l1, l5, l8, r3, r6, r7);
        finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_TO_REGISTERS>(p, c, q, src, dst, x0, span0, gy, ym, yp);
        T *row = s + ey * TL::SP + TL::OFF + lane * V;
#pragma unroll
        for (int j = 0; j < 9; ++j) sts_pack<T, V>(row + j * TL::PS, q[j]);
    }
    // ---- phase 1, the two rim columns (x0-1 and x0+SPAN of the tile): one lane per cell ----------------
    if (threadIdx.x < 2 * TL::EY) {
        const int side = threadIdx.x >= TL::EY, ey = threadIdx.x - side * TL::EY;
        int gx = side ? span0 + TL::SPAN : span0 - 1, gy = y0 - 1 + ey;
        bool live = gx <= nx && gy <= ny;
        if (periodic) {
            if (gx < 0) gx = nx - 1;
            if (gx == nx) gx = 0;
            if (gy < 0) gy = ny - 1;
            if (gy == ny) gy = 0;
        } else if (gx < 0 || gx == nx || gy < 0 || gy == ny) live = false;
        if (live) {
            Tb2Params tp;
            tp.src = p.src; tp.dst = p.dst; tp.plane = plane; tp.nx = nx; tp.ny = ny; tp.pitch = pitch;
            tp.bc = p.bc; tp.zero_obstacle_velocity = p.zero_obstacle_velocity; tp.mask = p.mask; tp.mask_pitch = p.mask_pitch;
            int xm = gx - 1, xp = gx + 1, ym = gy - 1, yp = gy + 1;
            bool has_xm = true, has_xp = true, has_ym = true, has_yp = true;
            if (periodic) {
                if (xm < 0) xm = nx - 1;
                if (xp == nx) xp = 0;
                if (ym < 0) ym = ny - 1;
                if (yp == ny) yp = 0;
            } else { has_xm = xm >= 0; has_xp = xp < nx; has_ym = ym >= 0; has_yp = yp < ny; }
            const long long rc = (long long)gy * pitch, rm = (long long)ym * pitch, rp = (long long)yp * pitch;
            T g[9];
            g[0] = __ldg(src + rc + gx);
            g[1] = has_xm ? __ldg(src + 1 * plane + rc + xm) : (T)0;
            g[3] = has_xp ? __ldg(src + 3 * plane + rc + xp) : (T)0;
            g[2] = has_ym ? __ldg(src + 2 * plane + rm + gx) : (T)0;
            g[4] = has_yp ? __ldg(src + 4 * plane + rp + gx) : (T)0;
            g[5] = (has_xm && has_ym) ? __ldg(src + 5 * plane + rm + xm) : (T)0;
            g[6] = (has_xp && has_ym) ? __ldg(src + 6 * plane + rm + xp) : (T)0;
            g[7] = (has_xp && has_yp) ? __ldg(src + 7 * plane + rp + xp) : (T)0;
            g[8] = (has_xm && has_yp) ? __ldg(src + 8 * plane + rp + xm) : (T)0;
            tb2_node<T, MATH>(tp, c, gx, gy, g);
            T *cell = s + ey * TL::SP + (side ? TL::OFF + TL::SPAN : TL::OFF - 1);
#pragma unroll
            for (int j = 0; j < 9; ++j) cell[j * TL::PS] = g[j];
        }
    }
    __syncthreads();

    // ---- phase 2: time level t+2 of rows y0 .. y0+BY-1, from the block to global memory --------------------
    for (int r = warp; r < BY; r += NW) {
        const int y = y0 + r;
        if (y >= ny) break;
        int ym = y - 1, yp = y + 1;                    // only finish_row's wall / halo logic looks at these
        if (periodic) { if (ym < 0) ym = ny - 1; if (yp >= ny) yp = 0; }
        const T *bc_ = s + (r + 1) * TL::SP + TL::OFF + lane * V;       // block row of y, this thread's cells
        const T *bm = bc_ - TL::SP + 2 * TL::PS;                           // row y-1: populations 2, 5, 6
        const T *bp = bc_ + TL::SP + 4 * TL::PS;                           // row y+1: populations 4, 7, 8
        Pack<T, V> q[9];
        q[0] = lds_pack<T, V>(bc_);
        q[1] = lds_pack<T, V>(bc_ + TL::PS);
        q[3] = lds_pack<T, V>(bc_ + 3 * TL::PS);
        q[2] = lds_pack<T, V>(bm);
        q[5] = lds_pack<T, V>(bm + 3 * TL::PS);
        q[6] = lds_pack<T, V>(bm + 4 * TL::PS);
        q[4] = lds_pack<T, V>(bp);
        q[7] = lds_pack<T, V>(bp + 3 * TL::PS);
        q[8] = lds_pack<T, V>(bp + 4 * TL::PS);
        T l1 = (T)0, l5 = (T)0, l8 = (T)0, r3 = (T)0, r6 = (T)0, r7 = (T)0;
        if (lane == 0) { l1 = bc_[TL::PS - 1]; l5 = bm[3 * TL::PS - 1]; l8 = bp[4 * TL::PS - 1]; }
        if (lane == 31) { r3 = bc_[3 * TL::PS + V]; r6 = bm[4 * TL::PS + V]; r7 = bp[3 * TL::PS + V]; }
        shift_x<T, V>(q, lane, l1, l5, l8, r3, r6, r7);
        finish_row<T, V, MATH, 0, MODEL_D2Q9, ROW_FROM_TILE>(p, c, q, src, dst, x0, span0, y, ym, yp);
    }
}

}  // namespace lb
