// TMA-staged variant of the fused step (sm_100a: cp.async.bulk.tensor + mbarrier).
//
// One elected thread issues nine 2-D box loads into shared memory; the -cy_j row displacement of
// each population is absorbed by the box's row coordinate.  The -cx_j displacement can NOT be
// absorbed the same way: UTMALDG requires the box start to be 16-byte aligned in the innermost
// dimension -- a box at x0-1 raises "illegal instruction" (compute-sanitizer, round 1) -- so the six
// x-moving populations are staged with an aligned halo of 16 B on each side (box start x0-H, width
// 32*V + 2H, H = 16/sizeof(T)) and every thread takes its own aligned vector with LDS.128 plus ONE
// neighbouring element with LDS.32/64.  Compared with the register-shuffle kernel this removes the
// shuffles, the predicated lane-0/31 global loads and most of the 64-bit address arithmetic.
// Everything after the loads is the same code as the other kernel (finish_row).
//
// The whole ping-pong buffer -- guard rows and all nine planes, which are contiguous in y -- is one
// 2-D tensor [2 + 9*ny + 2][pitch]; plane j / row y is tensor row 2 + j*ny + y.  Rows "above" y=0 or
// "below" y=ny-1 of a plane are a neighbouring plane's rows, and columns left of x=0 / right of the
// pitch are zero-filled by the TMA unit: garbage that the boundary closure or the slab-edge fix-up
// overwrites, exactly as in the register kernel.  Used for single-slab, non-periodic lattices; other
// configurations take the register-shuffle kernel.
#pragma once
#include <cuda.h>

#include "lb_fused.cuh"

namespace lb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// CTA = TY warps; tile = TY rows x (32*V) cells; one thread = V cells of one row.
// tmap_n: box (32*V) x TY for the populations with cx = 0; tmap_w: box (32*V + 2H) x TY for the others.
template <typename T, int V, int MATH, int TY, int MINB, int STP, int MODEL>
__global__ void __launch_bounds__(32 * TY, MINB)
fused_step_tma_kernel(const __grid_constant__ CUtensorMap tmap_n, const __grid_constant__ CUtensorMap tmap_w,
                      const StepParams p)
{
    constexpr int TX = 32 * V;
    constexpr int H = 16 / (int)sizeof(T);                     // halo elements = 16 bytes
    constexpr int TW = TX + 2 * H;
    struct alignas(128) NTile { T v[TY][TX]; };                // every TMA destination must be 128-byte aligned
    struct alignas(128) WTile { T v[TY][TW]; };
    __shared__ NTile tile_n[3];                                // populations 0, 2, 4
    __shared__ WTile tile_w[6];                                // populations 1, 5, 8 (from x-1) and 3, 6, 7 (from x+1)
    __shared__ alignas(8) unsigned long long bar_storage;

    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * TX;
    const int ty0 = (blockIdx.z * gridDim.y + blockIdx.y) * TY;
    if (ty0 >= p.ny) return;                                   // CTA-uniform
    const uint32_t bar = smem_u32(&bar_storage);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)((3 * TY * TX + 6 * TY * TW) * sizeof(T)));
        const int r0 = 2 + ty0, ny = p.ny;                     // 2 guard rows precede plane 0
        const int xw = tx0 - H;                                // aligned start of the haloed boxes
        tma_load_2d(smem_u32(&tile_n[0].v[0][0]), &tmap_n, tx0, r0 + 0 * ny,     bar);   // f0: row y
        tma_load_2d(smem_u32(&tile_n[1].v[0][0]), &tmap_n, tx0, r0 + 2 * ny - 1, bar);   // f2: row y-1
        tma_load_2d(smem_u32(&tile_n[2].v[0][0]), &tmap_n, tx0, r0 + 4 * ny + 1, bar);   // f4: row y+1
        tma_load_2d(smem_u32(&tile_w[0].v[0][0]), &tmap_w, xw,  r0 + 1 * ny,     bar);   // f1: row y
        tma_load_2d(smem_u32(&tile_w[1].v[0][0]), &tmap_w, xw,  r0 + 5 * ny - 1, bar);   // f5: row y-1
        tma_load_2d(smem_u32(&tile_w[2].v[0][0]), &tmap_w, xw,  r0 + 8 * ny + 1, bar);   // f8: row y+1
        tma_load_2d(smem_u32(&tile_w[3].v[0][0]), &tmap_w, xw,  r0 + 3 * ny,     bar);   // f3: row y
        tma_load_2d(smem_u32(&tile_w[4].v[0][0]), &tmap_w, xw,  r0 + 6 * ny - 1, bar);   // f6: row y-1
        tma_load_2d(smem_u32(&tile_w[5].v[0][0]), &tmap_w, xw,  r0 + 7 * ny + 1, bar);   // f7: row y+1
    }
    mbar_wait(bar, 0);

    const int y = ty0 + wy;
    if (y >= p.ny) return;                                     // warp-uniform; nothing follows the barrier
    using VT = typename VecOf<T, V>::type;
    Pack<T, V> q[9], own;
    unpack(*reinterpret_cast<const VT *>(&tile_n[0].v[wy][lane * V]), q[0]);
    unpack(*reinterpret_cast<const VT *>(&tile_n[1].v[wy][lane * V]), q[2]);
    unpack(*reinterpret_cast<const VT *>(&tile_n[2].v[wy][lane * V]), q[4]);
    constexpr int plus_x[3] = {1, 5, 8}, minus_x[3] = {3, 6, 7};
#pragma unroll
    for (int k = 0; k < 3; ++k) {                              // movers in +x take the cell to their left
        const T *base = &tile_w[k].v[wy][H + lane * V];
        unpack(*reinterpret_cast<const VT *>(base), own);
        q[plus_x[k]].v[0] = base[-1];
#pragma unroll
        for (int e = 1; e < V; ++e) q[plus_x[k]].v[e] = own.v[e - 1];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {                              // movers in -x take the cell to their right
        const T *base = &tile_w[3 + k].v[wy][H + lane * V];
        unpack(*reinterpret_cast<const VT *>(base), own);
#pragma unroll
        for (int e = 0; e < V - 1; ++e) q[minus_x[k]].v[e] = own.v[e + 1];
        q[minus_x[k]].v[V - 1] = base[V];
    }
    finish_row<T, V, MATH, STP, MODEL>(p, consts_in<T>(p), q, static_cast<const T *>(p.src), static_cast<T *>(p.dst),
                                       tx0 + lane * V, tx0, y, y - 1, y + 1);
}

}  // namespace lb
