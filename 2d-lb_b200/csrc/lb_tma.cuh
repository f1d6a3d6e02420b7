// TMA-staged variant of the fused step (sm_100a: cp.async.bulk.tensor + mbarrier).
//
// The register-shuffle kernel (lb_fused.cuh) resolves the +-1 x-shift of the moving populations
// with shuffles, two predicated scalar loads per shifted population and ~60 instructions of 64-bit
// address arithmetic per thread-row.  Here the shift is absorbed by the TMA unit instead: one
// elected thread issues nine 2-D box loads whose start coordinates are already displaced by
// (-cx_j, -cy_j), the hardware zero-fills what lies left / right of the row, and every thread then
// reads its nine vectors from shared memory at ONE common offset (LDS.128, conflict-free).
// Everything after the loads is the same code as the other kernel (finish_row).
//
// The whole ping-pong buffer -- guard rows and all nine planes, which are contiguous in y -- is one
// 2-D tensor [2 + 9*ny + 2][pitch]; plane j / row y is tensor row 2 + j*ny + y.  Rows "above" y=0 or
// "below" y=ny-1 of a plane are a neighbouring plane's rows: garbage that the boundary closure
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
template <typename T, int V, int MATH, int TY, int MINB, int STP, int MODEL>
__global__ void __launch_bounds__(32 * TY, MINB)
fused_step_tma_kernel(const __grid_constant__ CUtensorMap tmap, const StepParams p)
{
    constexpr int TX = 32 * V;
    __shared__ alignas(128) T tile[9][TY][TX];
    __shared__ alignas(8) unsigned long long bar_storage;

    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * TX;
    const int ty0 = (blockIdx.z * gridDim.y + blockIdx.y) * TY;
    if (ty0 >= p.ny) return;                                   // CTA-uniform
    const uint32_t bar = smem_u32(&bar_storage);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 9u * TY * TX * (uint32_t)sizeof(T));
        const int r0 = 2 + ty0, ny = p.ny;                     // 2 guard rows precede plane 0
        tma_load_2d(smem_u32(&tile[0][0][0]), &tmap, tx0,     r0 + 0 * ny,     bar);
        tma_load_2d(smem_u32(&tile[1][0][0]), &tmap, tx0 - 1, r0 + 1 * ny,     bar);
        tma_load_2d(smem_u32(&tile[2][0][0]), &tmap, tx0,     r0 + 2 * ny - 1, bar);
        tma_load_2d(smem_u32(&tile[3][0][0]), &tmap, tx0 + 1, r0 + 3 * ny,     bar);
        tma_load_2d(smem_u32(&tile[4][0][0]), &tmap, tx0,     r0 + 4 * ny + 1, bar);
        tma_load_2d(smem_u32(&tile[5][0][0]), &tmap, tx0 - 1, r0 + 5 * ny - 1, bar);
        tma_load_2d(smem_u32(&tile[6][0][0]), &tmap, tx0 + 1, r0 + 6 * ny - 1, bar);
        tma_load_2d(smem_u32(&tile[7][0][0]), &tmap, tx0 + 1, r0 + 7 * ny + 1, bar);
        tma_load_2d(smem_u32(&tile[8][0][0]), &tmap, tx0 - 1, r0 + 8 * ny + 1, bar);
    }
    mbar_wait(bar, 0);

    const int y = ty0 + wy;
    if (y >= p.ny) return;                                     // warp-uniform; nothing follows the barrier
    using VT = typename VecOf<T, V>::type;
    Pack<T, V> q[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) unpack(*reinterpret_cast<const VT *>(&tile[j][wy][lane * V]), q[j]);
    finish_row<T, V, MATH, STP, MODEL>(p, consts_in<T>(p), q, static_cast<const T *>(p.src), static_cast<T *>(p.dst),
                                       tx0 + lane * V, tx0, y, y - 1, y + 1);
}

}  // namespace lb
