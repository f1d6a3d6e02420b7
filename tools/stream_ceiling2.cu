// Wider search for the HBM ceiling of the D2Q9 access pattern (companion of stream_ceiling.cu): is there
// ANY arithmetic-free way of moving nine planes in and nine planes out that beats the fused kernel's
// own thread mapping?  Variants: CTA shapes, rows per thread, load/store cache hints, and a register-free
// path made of 1-D bulk asynchronous copies (cp.async.bulk global->shared->global, mbarrier pipeline).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_ceiling2 tools/stream_ceiling2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int H> __device__ __forceinline__ float4 ld4(const float *p)
{
    float4 v;
    if (H == 0) v = *(const float4 *)p;
    else if (H == 1) v = __ldg((const float4 *)p);
    else if (H == 2) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int H> __device__ __forceinline__ void st4(float *p, float4 v)
{
    if (H == 0) *(float4 *)p = v;
    else if (H == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// warp = 128 cells of a row; CTA = WX x WY warps; each thread RPT consecutive rows; D2Q9 row offsets
template <int WX, int WY, int RPT, int LH, int SH, int MINB>
__global__ void __launch_bounds__(32 * WX * WY, MINB) tile_copy(const float *__restrict__ src, float *__restrict__ dst, int nx, int ny, long long plane)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = ((blockIdx.x * WX + warp % WX) * 32 + lane) * 4;
    const int yb = ((blockIdx.z * gridDim.y + blockIdx.y) * WY + warp / WX) * RPT;
    if (x0 >= nx || yb >= ny) return;
    float4 q[RPT][9];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = yb + r;
        const int ym = y > 0 ? y - 1 : ny - 1, yp = y < ny - 1 ? y + 1 : 0;
        const int rows[9] = {y, y, ym, y, yp, ym, ym, yp, yp};
#pragma unroll
        for (int j = 0; j < 9; ++j) q[r][j] = ld4<LH>(src + j * plane + (long long)rows[j] * nx + x0);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int j = 0; j < 9; ++j) st4<SH>(dst + j * plane + (long long)(yb + r) * nx + x0, q[r][j]);
}

// ---- register-free: 1-D bulk async copies through shared memory --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra.uni WD;\nbra.uni WL;\nWD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
template <int SEG, int S>
__global__ void __launch_bounds__(32, 1) bulk_copy(const char *__restrict__ src, char *__restrict__ dst, long long plane_bytes, long long row_bytes, long long n_tiles)
{
    extern __shared__ __align__(128) char smem[];
    if (threadIdx.x != 0) return;
    const uint32_t bars = smem_u32(smem), buf = smem_u32(smem + 128);
    for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long n = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // my tiles
    const int shift[9] = {0, 0, -1, 0, 1, -1, -1, 1, 1};
    auto issue_load = [&](long long i) {
        const int s = (int)(i % S);
        const long long off = (blockIdx.x + i * gridDim.x) * (long long)SEG;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"(9 * SEG) : "memory");
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            long long o = off + shift[j] * row_bytes;
            if (o < 0) o += plane_bytes;
            if (o >= plane_bytes) o -= plane_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(buf + (s * 9 + j) * SEG), "l"(src + j * plane_bytes + o), "r"(SEG), "r"(bars + 8 * s) : "memory");
        }
    };
    constexpr int LA = S - 1;
    for (long long i = 0; i < LA && i < n; ++i) issue_load(i);
    for (long long i = 0; i < n; ++i) {
        const int s = (int)(i % S);
        if (i + LA < n) {
            if (i >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stage of iteration i-1 is free
            issue_load(i + LA);
        }
        mbar_wait(bars + 8 * s, (uint32_t)((i / S) & 1));
        const long long off = (blockIdx.x + i * gridDim.x) * (long long)SEG;
#pragma unroll
        for (int j = 0; j < 9; ++j)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst + j * plane_bytes + off), "r"(buf + (s * 9 + j) * SEG), "r"(SEG) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float *A, *B;
static cudaStream_t st;
static cudaEvent_t e0, e1;
static int NX, NY;
static long long PLANE;
static size_t BYTES;

template <typename F> static int timeit(const char *name, F launch)
{
    const int reps = 10;
    float best = 1e30f;
    for (int pass = 0; pass < 3; ++pass) {
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r) launch((r & 1) ? B : A, (r & 1) ? A : B);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        if (pass && ms < best) best = ms;
    }
    printf("%-44s %.4f ms %7.1f GB/s %7.0f MLUPS-eq\n", name, best, 2.0 * BYTES / 1e6 / best, PLANE / best / 1e3);
    fflush(stdout);
    return 0;
}

template <int WX, int WY, int RPT, int LH, int SH, int MINB> static int run_tile()
{
    char name[96];
    snprintf(name, sizeof name, "tile wx%d wy%d rows/thr %d ld%d st%d minb%d", WX, WY, RPT, LH, SH, MINB);
    const int ty = (NY + WY * RPT - 1) / (WY * RPT);
    const dim3 grid((NX / 4 + 32 * WX - 1) / (32 * WX), ty < 65535 ? ty : 65535, (ty + 65534) / 65535);
    return timeit(name, [&](float *s, float *d) { tile_copy<WX, WY, RPT, LH, SH, MINB><<<grid, 32 * WX * WY, 0, st>>>(s, d, NX, NY, PLANE); });
}

template <int SEG, int S> static int run_bulk(int ctas_per_sm)
{
    char name[96];
    snprintf(name, sizeof name, "bulk seg %d B x9, %d stages, %d CTA/SM", SEG, S, ctas_per_sm);
    const int smem = 128 + S * 9 * SEG;
    CK(cudaFuncSetAttribute(bulk_copy<SEG, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long plane_bytes = PLANE * 4, n_tiles = plane_bytes / SEG;
    return timeit(name, [&](float *s, float *d) {
        bulk_copy<SEG, S><<<148 * ctas_per_sm, 32, smem, st>>>((const char *)s, (char *)d, plane_bytes, (long long)NX * 4, n_tiles);
    });
}

int main(int argc, char **argv)
{
    NX = argc > 2 ? atoi(argv[1]) : 16384; NY = argc > 2 ? atoi(argv[2]) : 16384;
    PLANE = (long long)NX * NY; BYTES = (size_t)9 * PLANE * 4;
    CK(cudaMalloc(&A, BYTES)); CK(cudaMalloc(&B, BYTES));
    CK(cudaMemset(A, 1, BYTES)); CK(cudaMemset(B, 0, BYTES));
    CK(cudaStreamCreate(&st)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("grid %d x %d fp32, %.2f GB per buffer; D2Q9 row offsets in every variant\n", NX, NY, BYTES / 1e9);
    timeit("cudaMemcpyAsync D2D", [&](float *s, float *d) { cudaMemcpyAsync(d, s, BYTES, cudaMemcpyDeviceToDevice, st); });
    run_tile<2, 2, 1, 1, 0, 6>();      // the fused kernel's mapping
    run_tile<2, 2, 1, 0, 0, 6>();
    run_tile<2, 2, 1, 2, 1, 6>();
    run_tile<2, 2, 1, 3, 1, 6>();
    run_tile<2, 2, 1, 1, 1, 6>();
    run_tile<2, 2, 1, 1, 2, 6>();
    run_tile<2, 2, 1, 1, 0, 8>();
    run_tile<2, 2, 1, 1, 0, 4>();
    run_tile<2, 2, 1, 1, 0, 2>();
    run_tile<4, 1, 1, 1, 0, 6>();
    run_tile<1, 4, 1, 1, 0, 6>();
    run_tile<1, 1, 1, 1, 0, 16>();
    run_tile<4, 2, 1, 1, 0, 3>();
    run_tile<2, 4, 1, 1, 0, 3>();
    run_tile<8, 1, 1, 1, 0, 3>();
    run_tile<2, 2, 2, 1, 0, 4>();
    run_tile<2, 2, 2, 1, 0, 3>();
    run_tile<1, 2, 2, 1, 0, 6>();
    run_tile<4, 1, 2, 1, 0, 3>();
    run_tile<2, 1, 4, 1, 0, 3>();
    run_bulk<2048, 3>(3);
    run_bulk<2048, 4>(2);
    run_bulk<4096, 2>(2);
    run_bulk<4096, 3>(2);
    run_bulk<4096, 4>(1);
    run_bulk<8192, 2>(1);
    run_bulk<1024, 4>(4);
    run_bulk<1024, 6>(4);
    run_bulk<2048, 3>(4);
    return 0;
}
