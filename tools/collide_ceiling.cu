// Arithmetic ceiling of the D2Q9 collision on this GPU, memory taken out of the picture: every thread keeps
// four nodes in registers and collides them `iters` times with the product's own per-node code
// (lb_device.cuh), once with scalar fp32 instructions and once with the packed f32x2 type (lb_f32x2.cuh:
// FADD2 / FMUL2 / FFMA2).  Prints node updates per second for STRICT and FAST math, and checks that the
// packed lanes are bit-identical to the scalar ones -- after FEW iterations, on a state far from equilibrium:
// the BGK iteration is a contraction, and after thousands of collisions two arithmetics that differ in the last
// bit have long converged to the same fixed point (which is how the first version of this tool missed ptxas
// contracting mul.rn.f32x2 + add.rn.f32x2; see lb_f32x2.cuh).  Also: raw FADD vs FADD2 issue throughput.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o tools/collide_ceiling tools/collide_ceiling.cu
#include <cstdio>
#include <vector>
#include "../2d-lb_b200/csrc/lb_device.cuh"

using namespace lb;

struct KP { Consts<float> cf; Consts<F2> c2; };

template <int MATH, int MINB>
__global__ void __launch_bounds__(128, MINB) k_scalar(KP kp, int iters, float *out)
{
    float g[4][9];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int n = 0; n < 4; ++n)
        for (int j = 0; j < 9; ++j) g[n][j] = (j == 0 ? 4.f / 9 : j < 5 ? 1.f / 9 : 1.f / 36) * (1.0f + 3e-2f * (float)((t * 4 + n + j * 7) % 13));
    float acc = 0.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            float rho, u, v;
            collide_node<float, MATH, MODEL_D2Q9>(kp.cf, g[n], rho, u, v, false);
            // a "stream": rotate the moving populations so that the state keeps changing
            const float t1 = g[n][1]; g[n][1] = g[n][2]; g[n][2] = g[n][3]; g[n][3] = g[n][4]; g[n][4] = t1;
        }
    }
    for (int n = 0; n < 4; ++n)
        for (int j = 0; j < 9; ++j) acc += g[n][j];
    out[t] = acc;
    if (blockIdx.x == 0)
        for (int n = 0; n < 4; ++n)
            for (int j = 0; j < 9; ++j) out[gridDim.x * blockDim.x + (threadIdx.x * 4 + n) * 9 + j] = g[n][j];
}

template <int MATH, int MINB>
__global__ void __launch_bounds__(128, MINB) k_packed(KP kp, int iters, float *out)
{
    F2 g[2][9];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int n = 0; n < 2; ++n)
        for (int j = 0; j < 9; ++j) {
            const float w = (j == 0 ? 4.f / 9 : j < 5 ? 1.f / 9 : 1.f / 36);
            g[n][j] = F2(w * (1.0f + 3e-2f * (float)((t * 4 + 2 * n + j * 7) % 13)), w * (1.0f + 3e-2f * (float)((t * 4 + 2 * n + 1 + j * 7) % 13)));
        }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            F2 rho, u, v;
            collide_node<F2, MATH, MODEL_D2Q9>(kp.c2, g[n], rho, u, v, false);
            const F2 t1 = g[n][1]; g[n][1] = g[n][2]; g[n][2] = g[n][3]; g[n][3] = g[n][4]; g[n][4] = t1;
        }
    }
    float acc = 0.f;
    for (int n = 0; n < 2; ++n)
        for (int j = 0; j < 9; ++j) { acc += g[n][j].lo(); }
    for (int n = 0; n < 2; ++n)
        for (int j = 0; j < 9; ++j) { acc += g[n][j].hi(); }
    out[t] = acc;
    if (blockIdx.x == 0)
        for (int n = 0; n < 2; ++n)
            for (int j = 0; j < 9; ++j) {
                out[gridDim.x * blockDim.x + (threadIdx.x * 4 + 2 * n) * 9 + j] = g[n][j].lo();
                out[gridDim.x * blockDim.x + (threadIdx.x * 4 + 2 * n + 1) * 9 + j] = g[n][j].hi();
            }
}

// raw issue throughput: 8 independent chains per thread
template <int PACKED, int OP>
__global__ void __launch_bounds__(128) k_raw(int iters, float *out, float seed)
{
    float acc = 0.f;
    if (PACKED) {
        F2 a[8], b(seed, seed * 0.5f), c(1.0001f, 0.9999f);
        for (int k = 0; k < 8; ++k) a[k] = F2((float)k + threadIdx.x, (float)k * 0.5f);
    #pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = OP == 0 ? a[k] + b : OP == 1 ? a[k] * c : f2_fma(a[k], c, b);
        }
        for (int k = 0; k < 8; ++k) acc += a[k].lo() + a[k].hi();
    } else {
        float a[16], b = seed, c = 1.0001f;
        for (int k = 0; k < 16; ++k) a[k] = (float)k + threadIdx.x;
    #pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = OP == 0 ? a[k] + b : OP == 1 ? a[k] * c : __fmaf_rn(a[k], c, b);
        }
        for (int k = 0; k < 16; ++k) acc += a[k];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    const double cs = 1. / sqrt(3.), cs2 = cs * cs;
    KP kp;
    kp.cf = make_consts<float>(1.7, 1.003, 1.0, cs2, 2 * cs2, 2 * cs2 * cs2);
    kp.c2 = pack_consts(kp.cf);
    const int iters = 2000;
    float *out;
    {   // bit-identity after 3 and after 17 collisions
        const int blocks = 148 * 4;
        const size_t n = (size_t)blocks * 128;
        cudaMalloc(&out, (n + 128 * 36) * sizeof(float));
        std::vector<float> hs(128 * 36), hp(128 * 36);
        for (int it : {3, 17}) {
            k_scalar<MATH_STRICT, 4><<<blocks, 128>>>(kp, it, out);
            cudaMemcpy(hs.data(), out + n, hs.size() * 4, cudaMemcpyDeviceToHost);
            k_packed<MATH_STRICT, 4><<<blocks, 128>>>(kp, it, out);
            cudaMemcpy(hp.data(), out + n, hp.size() * 4, cudaMemcpyDeviceToHost);
            printf("%2d collisions  STRICT packed == scalar: %s\n", it, memcmp(hs.data(), hp.data(), hs.size() * 4) ? "NO" : "yes");
            k_scalar<MATH_FAST, 4><<<blocks, 128>>>(kp, it, out);
            cudaMemcpy(hs.data(), out + n, hs.size() * 4, cudaMemcpyDeviceToHost);
            k_packed<MATH_FAST, 4><<<blocks, 128>>>(kp, it, out);
            cudaMemcpy(hp.data(), out + n, hp.size() * 4, cudaMemcpyDeviceToHost);
            printf("%2d collisions  FAST   packed == scalar: %s\n", it, memcmp(hs.data(), hp.data(), hs.size() * 4) ? "NO" : "yes");
        }
        cudaFree(out);
    }
    for (int warps_per_sm : {8, 16, 24}) {
        const int blocks = 148 * warps_per_sm / 4;
        const size_t n = (size_t)blocks * 128;
        cudaMalloc(&out, (n + 128 * 36) * sizeof(float));
        std::vector<float> hs(128 * 36), hp(128 * 36);
        const double nodes = (double)n * 4 * iters;
        double ms;
        ms = time_ms([&] { k_scalar<MATH_STRICT, 4><<<blocks, 128>>>(kp, iters, out); });
        cudaMemcpy(hs.data(), out + n, hs.size() * 4, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  STRICT scalar  %8.1f G node-updates/s\n", warps_per_sm, nodes / ms / 1e6);
        ms = time_ms([&] { k_packed<MATH_STRICT, 4><<<blocks, 128>>>(kp, iters, out); });
        cudaMemcpy(hp.data(), out + n, hp.size() * 4, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  STRICT packed  %8.1f G node-updates/s   bit-identical to scalar: %s\n", warps_per_sm, nodes / ms / 1e6,
               memcmp(hs.data(), hp.data(), hs.size() * 4) ? "NO" : "yes");
        ms = time_ms([&] { k_scalar<MATH_FAST, 4><<<blocks, 128>>>(kp, iters, out); });
        cudaMemcpy(hs.data(), out + n, hs.size() * 4, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  FAST   scalar  %8.1f G node-updates/s\n", warps_per_sm, nodes / ms / 1e6);
        ms = time_ms([&] { k_packed<MATH_FAST, 4><<<blocks, 128>>>(kp, iters, out); });
        cudaMemcpy(hp.data(), out + n, hp.size() * 4, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  FAST   packed  %8.1f G node-updates/s   bit-identical to scalar: %s\n", warps_per_sm, nodes / ms / 1e6,
               memcmp(hs.data(), hp.data(), hs.size() * 4) ? "NO" : "yes");
        cudaFree(out);
    }
    {
        const int blocks = 148 * 8, it = 20000;
        cudaMalloc(&out, (size_t)blocks * 128 * sizeof(float));
        const char *names[3] = {"add", "mul", "fma"};
        const double lane_ops = (double)blocks * 128 * 16 * it;
        double ms;
        ms = time_ms([&] { k_raw<0, 0><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw scalar %s  %7.2f T lane-ops/s\n", names[0], lane_ops / ms / 1e9);
        ms = time_ms([&] { k_raw<1, 0><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw packed %s  %7.2f T lane-ops/s\n", names[0], lane_ops / ms / 1e9);
        ms = time_ms([&] { k_raw<0, 1><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw scalar %s  %7.2f T lane-ops/s\n", names[1], lane_ops / ms / 1e9);
        ms = time_ms([&] { k_raw<1, 1><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw packed %s  %7.2f T lane-ops/s\n", names[1], lane_ops / ms / 1e9);
        ms = time_ms([&] { k_raw<0, 2><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw scalar %s  %7.2f T lane-ops/s\n", names[2], lane_ops / ms / 1e9);
        ms = time_ms([&] { k_raw<1, 2><<<blocks, 128>>>(it, out, 1e-3f); }); printf("raw packed %s  %7.2f T lane-ops/s\n", names[2], lane_ops / ms / 1e9);
        cudaFree(out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
