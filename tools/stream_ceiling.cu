// What can this GPU's HBM sustain for the D2Q9 access pattern, with the arithmetic removed?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_ceiling tools/stream_ceiling.cu
//   tools/stream_ceiling [nx ny]            (default 16384 x 16384 fp32: 2 x 9.66 GB)
//
// Four kernels over the same 9-plane SoA buffers the library uses (row pitch = nx, 72 B per cell):
//   memcpy     cudaMemcpyAsync device-to-device of the nine planes (the driver's copy engine path)
//   flat       grid-stride float4 copy of the whole buffer, 148 x 8 CTAs
//   tile       the fused kernel's thread mapping (warp = 128 cells of a row, CTA = 2 x 2 warps, nine 128-bit
//              loads + nine 128-bit stores per thread), all planes read at the SAME row
//   pull       the same, with the D2Q9 row offsets (planes 2,5,6 from row y-1; 4,7,8 from row y+1): the
//              fused kernel's exact DRAM access stream without shuffles, boundary work or collision
// The best of these is the practical ceiling for `fused_step_kernel`; profiles/README.md quotes it.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void flat_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __ldg(src + i);
}

template <bool PULL>
__global__ void __launch_bounds__(128, 6) tile_copy(const float *__restrict__ src, float *__restrict__ dst, int nx, int ny, long long plane)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = ((blockIdx.x * 2 + (warp & 1)) * 32 + lane) * 4;
    const int y = (blockIdx.z * gridDim.y + blockIdx.y) * 2 + (warp >> 1);
    if (x0 >= nx || y >= ny) return;
    const int ym = PULL ? (y > 0 ? y - 1 : ny - 1) : y, yp = PULL ? (y < ny - 1 ? y + 1 : 0) : y;
    const int rows[9] = {y, y, ym, y, yp, ym, ym, yp, yp};
    float4 q[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = __ldg((const float4 *)(src + j * plane + (long long)rows[j] * nx + x0));
#pragma unroll
    for (int j = 0; j < 9; ++j) *(float4 *)(dst + j * plane + (long long)y * nx + x0) = q[j];
}

int main(int argc, char **argv)
{
    const int nx = argc > 2 ? atoi(argv[1]) : 16384, ny = argc > 2 ? atoi(argv[2]) : 16384;
    const long long plane = (long long)nx * ny;
    const size_t bytes = (size_t)9 * plane * sizeof(float);
    float *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 20;
    const double gb = 2.0 * bytes / 1e9;
    const dim3 grid(nx / 256, ny / 2 < 65535 ? ny / 2 : 65535, (ny / 2 + 65534) / 65535);
    printf("grid %d x %d fp32, %.2f GB per buffer, %d repetitions each (ping-pong a<->b)\n", nx, ny, bytes / 1e9, reps);
    for (int k = 0; k < 4; ++k) {
        const char *name[4] = {"memcpy", "flat", "tile", "pull"};
        float best = 1e30f, sum = 0;
        for (int pass = 0; pass < 3; ++pass) {
            CK(cudaEventRecord(e0, st));
            for (int r = 0; r < reps; ++r) {
                float *s = (r & 1) ? b : a, *d = (r & 1) ? a : b;
                if (k == 0) CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st));
                else if (k == 1) flat_copy<<<148 * 8, 256, 0, st>>>((const float4 *)s, (float4 *)d, bytes / 16);
                else if (k == 2) tile_copy<false><<<grid, 128, 0, st>>>(s, d, nx, ny, plane);
                else tile_copy<true><<<grid, 128, 0, st>>>(s, d, nx, ny, plane);
            }
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= reps;
            if (pass) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-7s best %.4f ms  %7.1f GB/s   mean %.4f ms  %7.1f GB/s   (%.0f MLUPS-equivalent)\n", name[k], best, gb / best * 1e3,
               sum / 2, gb / (sum / 2) * 1e3, plane / best / 1e3);
    }
    return 0;
}
