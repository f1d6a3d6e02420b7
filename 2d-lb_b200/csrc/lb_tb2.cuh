// Temporal blocking: TWO lattice updates per pass through HBM.
//
// `fused_step_kernel` already runs at the speed of an arithmetic-free copy (profiles/README.md, section 6):
// 9 loads + 9 stores per update cannot go faster on this device.  The remaining lever is to move fewer
// bytes: this kernel keeps the intermediate time level on chip.  A CTA owns a BX x BY tile of the output.
//   phase 1  every thread-strided cell of the tile grown by one cell on each side (the cells whose
//            step-1 state the tile's step 2 pulls from) is updated exactly as `fused_step_kernel` would --
//            pull from global memory, boundary closure, bounce-back, collision -- and its nine
//            post-collision populations go to shared memory;
//   phase 2  every cell of the tile pulls from that shared-memory block, is updated again and stored.
// DRAM traffic per TWO updates: 36 B read (+ the one-cell rim, mostly L2 hits between neighbouring
// CTAs) and 36 B written, i.e. about 40 B per lattice update instead of 72.  The price is the rim
// recomputed in phase 1, (BX+2)(BY+2)/(BX*BY) - 1 = 14 % at 128 x 16, and a kernel that is bound by
// instruction issue rather than by HBM.
//
// Both phases call the same per-node functions as the one-step kernel (lb_device.cuh), in the same
// order, on the same values, so a run that mixes the two kernels is BIT-IDENTICAL to one that does
// not (tests/test_parity_gpu.py::test_temporal_blocking_is_bit_identical).  The phases are plain
// `__host__ __device__` loops over a thread id: tools/tb2_host.cu replays them CTA by CTA on the CPU
// against the oracle (tests/test_host_logic.py), so the index logic is checked without a GPU.
//
// Scope: single slab (no halo edges), pipe or periodic boundaries, D2Q9 model, moment-free steps; lb_step
// uses it for pairs of steps and the one-step kernel for an odd step and for the last step of a run
// (which stores rho, u, v).
#pragma once
#include "lb_device.cuh"

namespace lb {

struct Tb2Params {
    const void *src;          // plane 0, row 0 of the buffer being read (time level t)
    void *dst;                // plane 0, row 0 of the buffer being written (time level t+2)
    long long plane;
    int nx, ny, pitch;
    int bc;                   // BC_PIPE / BC_PERIODIC
    int zero_obstacle_velocity;
    const uint8_t *mask;      // [ny][mask_pitch] or nullptr
    int mask_pitch;
    Consts<float> cf;
    Consts<double> cd;
};

template <typename T> LB_HD const Consts<T> &consts_in(const Tb2Params &p);
template <> LB_HD const Consts<float> &consts_in<float>(const Tb2Params &p) { return p.cf; }
template <> LB_HD const Consts<double> &consts_in<double>(const Tb2Params &p) { return p.cd; }

template <typename T> LB_HD T tb2_ld(const T *p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// one node: [closure] -> [bounce-back] -> moments + equilibrium + relaxation, in place on g
template <typename T, int MATH>
LB_HD void tb2_node(const Tb2Params &p, const Consts<T> &c, int gx, int gy, T (&g)[9])
{
    if (p.bc != BC_PERIODIC) pipe_bc<T, MODEL_D2Q9>(c, gx, gy, p.nx, p.ny, g);
    bool solid = false;
    if (p.mask != nullptr) solid = p.mask[(long long)gy * p.mask_pitch + gx] == 1;
    if (solid) bounce_back<T>(g);
    T rho, u, v;
    collide_node<T, MATH, MODEL_D2Q9>(c, g, rho, u, v, solid && p.zero_obstacle_velocity);
}

// shared-memory block: nine planes of (BY+2) rows x (BX+2) cells; cell (ex, ey) is the lattice node
// (x0 - 1 + ex, y0 - 1 + ey), wrapped on a periodic box
template <int BX, int BY> struct Tb2Tile {
    static constexpr int EX = BX + 2, EY = BY + 2, CELLS = EX * EY;
};

// ---- phase 1: time level t+1 of the grown tile, into shared memory -------------------------------
template <typename T, int MATH, int BX, int BY>
LB_HD void tb2_phase1(const Tb2Params &p, T *__restrict__ s, int x0, int y0, int tid, int nthreads)
{
    using TL = Tb2Tile<BX, BY>;
    const T *__restrict__ src = static_cast<const T *>(p.src);
    const Consts<T> &c = consts_in<T>(p);
    const bool periodic = (p.bc == BC_PERIODIC);
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const long long plane = p.plane;
    for (int cell = tid; cell < TL::CELLS; cell += nthreads) {
        const int ey = cell / TL::EX, ex = cell - ey * TL::EX;
        int gx = x0 - 1 + ex, gy = y0 - 1 + ey;
        // nodes beyond the rim of the last (partial) tile are never pulled from
        if (gx > nx || gy > ny) continue;
        if (periodic) {
            if (gx < 0) gx = nx - 1;
            if (gx == nx) gx = 0;
            if (gy < 0) gy = ny - 1;
            if (gy == ny) gy = 0;
        } else if (gx < 0 || gx == nx || gy < 0 || gy == ny) {
            continue;                                   // outside the pipe: "unknown", as in the one-step kernel
        }
        // pull: population j comes from (gx - cx_j, gy - cy_j)
        int xm = gx - 1, xp = gx + 1, ym = gy - 1, yp = gy + 1;
        bool has_xm = true, has_xp = true, has_ym = true, has_yp = true;
        if (periodic) {
            if (xm < 0) xm = nx - 1;
            if (xp == nx) xp = 0;
            if (ym < 0) ym = ny - 1;
            if (yp == ny) yp = 0;
        } else {
            has_xm = xm >= 0; has_xp = xp < nx; has_ym = ym >= 0; has_yp = yp < ny;
        }
        const long long rc = (long long)gy * pitch, rm = (long long)ym * pitch, rp = (long long)yp * pitch;
        T g[9];
        g[0] = tb2_ld(src + rc + gx);
        g[1] = has_xm ? tb2_ld(src + 1 * plane + rc + xm) : (T)0;
        g[3] = has_xp ? tb2_ld(src + 3 * plane + rc + xp) : (T)0;
        g[2] = has_ym ? tb2_ld(src + 2 * plane + rm + gx) : (T)0;
        g[4] = has_yp ? tb2_ld(src + 4 * plane + rp + gx) : (T)0;
        g[5] = (has_xm && has_ym) ? tb2_ld(src + 5 * plane + rm + xm) : (T)0;
        g[6] = (has_xp && has_ym) ? tb2_ld(src + 6 * plane + rm + xp) : (T)0;
        g[7] = (has_xp && has_yp) ? tb2_ld(src + 7 * plane + rp + xp) : (T)0;
        g[8] = (has_xm && has_yp) ? tb2_ld(src + 8 * plane + rp + xm) : (T)0;
        tb2_node<T, MATH>(p, c, gx, gy, g);
#pragma unroll
        for (int j = 0; j < 9; ++j) s[j * TL::CELLS + cell] = g[j];
    }
}

// ---- phase 2: time level t+2 of the tile, from shared memory to global memory ----------------------
//      One cell per thread and iteration, lanes along x: conflict-free LDS.32 and fully coalesced
//      128-byte (fp32) / 256-byte (fp64) store segments per warp and population.
template <typename T, int MATH, int BX, int BY>
LB_HD void tb2_phase2(const Tb2Params &p, const T *__restrict__ s, int x0, int y0, int tid, int nthreads)
{
    using TL = Tb2Tile<BX, BY>;
    T *__restrict__ dst = static_cast<T *>(p.dst);
    const Consts<T> &c = consts_in<T>(p);
    const int nx = p.nx, ny = p.ny, pitch = p.pitch;
    const long long plane = p.plane;
    for (int item = tid; item < BX * BY; item += nthreads) {
        const int ty = item / BX, tx = item - ty * BX;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= nx || y >= ny) continue;
        // tile cell (tx, ty) is block cell (tx + 1, ty + 1); its sources are its block neighbours.  A source
        // outside the pipe was never written in phase 1: garbage that the closure overwrites.
        const T *sc = s + (ty + 1) * TL::EX + (tx + 1);
        T g[9];
        g[0] = sc[0];
        g[1] = sc[1 * TL::CELLS - 1];
        g[3] = sc[3 * TL::CELLS + 1];
        g[2] = sc[2 * TL::CELLS - TL::EX];
        g[4] = sc[4 * TL::CELLS + TL::EX];
        g[5] = sc[5 * TL::CELLS - TL::EX - 1];
        g[6] = sc[6 * TL::CELLS - TL::EX + 1];
        g[7] = sc[7 * TL::CELLS + TL::EX + 1];
        g[8] = sc[8 * TL::CELLS + TL::EX - 1];
        tb2_node<T, MATH>(p, c, x, y, g);
        const long long rc = (long long)y * pitch + x;
#pragma unroll
        for (int j = 0; j < 9; ++j) dst[j * plane + rc] = g[j];
    }
}

#ifdef __CUDACC__
template <typename T, int MATH, int BX, int BY, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) fused_two_step_kernel(const Tb2Params p)
{
    extern __shared__ __align__(16) unsigned char tb2_smem[];
    T *s = reinterpret_cast<T *>(tb2_smem);
    const int x0 = blockIdx.x * BX, y0 = blockIdx.y * BY;
    tb2_phase1<T, MATH, BX, BY>(p, s, x0, y0, threadIdx.x, NT);
    __syncthreads();
    tb2_phase2<T, MATH, BX, BY>(p, s, x0, y0, threadIdx.x, NT);
}
#endif

}  // namespace lb
