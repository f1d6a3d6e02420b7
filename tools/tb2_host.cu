// TEST INFRASTRUCTURE -- CPU replay of the temporally blocked kernel's tile logic (csrc/lb_tb2.cuh).
//
//   nvcc -O2 -std=c++17 --shared -Xcompiler -fPIC -o tools/libtb2_host.so tools/tb2_host.cu
//
// The two phases of `fused_two_step_kernel` are `__host__ __device__` loops over a thread id; here every
// CTA of the grid is replayed sequentially (all threads through phase 1, then all through phase 2) with a
// heap block standing in for shared memory.  tests/test_host_logic.py compares the result with two steps
// of the CPU oracle, bit for bit (STRICT math), for pipe / periodic / obstacle cases and several tile
// shapes -- the index logic of the kernel is verified before it ever touches a GPU.
#include <vector>
#include <cstring>
#include "../2d-lb_b200/csrc/lb_tb2.cuh"

using namespace lb;

template <typename T, int BX, int BY>
static void replay(const Tb2Params &p, int nthreads)
{
    using TL = Tb2Tile<BX, BY>;
    std::vector<T> smem((size_t)9 * TL::CELLS);
    for (int by = 0; by < (p.ny + BY - 1) / BY; ++by)
        for (int bx = 0; bx < (p.nx + BX - 1) / BX; ++bx) {
            // poison: a phase-2 read of a block cell phase 1 skipped must not matter
            std::memset(smem.data(), 0xff, smem.size() * sizeof(T));
            for (int t = 0; t < nthreads; ++t) tb2_phase1<T, MATH_STRICT, BX, BY>(p, smem.data(), bx * BX, by * BY, t, nthreads);
            for (int t = 0; t < nthreads; ++t) tb2_phase2<T, MATH_STRICT, BX, BY>(p, smem.data(), bx * BX, by * BY, t, nthreads);
        }
}

extern "C" int tb2_host_run(int nx, int ny, int bc, int is_f64, const void *f_in, void *f_out, const uint8_t *mask,
                            double omega, double rin, double rout, double cs2, double cs22, double two_cs4,
                            int zero_vel, int shape, int nthreads)
{
    Tb2Params p{};
    p.src = f_in; p.dst = f_out;
    p.plane = (long long)nx * ny; p.nx = nx; p.ny = ny; p.pitch = nx;
    p.bc = bc; p.zero_obstacle_velocity = zero_vel;
    p.mask = mask; p.mask_pitch = nx;
    p.cf = make_consts<float>(omega, rin, rout, cs2, cs22, two_cs4);
    p.cd = make_consts<double>(omega, rin, rout, cs2, cs22, two_cs4);
#define SHAPE(ID, BX, BY)                                                        \
    if (shape == ID) {                                                           \
        if (is_f64) replay<double, BX, BY>(p, nthreads); else replay<float, BX, BY>(p, nthreads); \
        return 0;                                                                \
    }
    SHAPE(0, 128, 16)
    SHAPE(1, 64, 32)
    SHAPE(2, 128, 8)
    SHAPE(3, 32, 4)
    SHAPE(4, 8, 8)
    return -1;
}
