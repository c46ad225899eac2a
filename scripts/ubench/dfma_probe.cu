// Micro-benchmarks behind DESIGN.md's latency model (one warp per scheduler):
//  (1) dependent-issue latency of DFMA (1 chain), throughput with 2/4/8 independent chains, full warp;
//  (2) the same with only 16 / 8 / 4 active lanes — does a part-filled warp occupy the FP64 pipe for fewer cycles?
//  (3) SHFL.BFLY + DADD round trip, LDS round trip.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_probe dfma_probe.cu ; run: ./dfma_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void k_dfma(double* out, int iters, int active, double a, double b) {
    if ((int)(threadIdx.x & 31) >= active) return;
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = a + c + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < CH; c++) x[c] = fma(x[c], b, a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 8 * CH);
}
__global__ void k_shfl(double* out, int iters, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) { x += __shfl_xor_sync(0xffffffffu, x, 1); x += __shfl_xor_sync(0xffffffffu, x, 2); }
    }
    long long t1 = clock64();
    out[threadIdx.x + 1] = x;
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 16);
}
__global__ void k_lds(double* out, int iters) {
    __shared__ double sm[64];
    sm[threadIdx.x] = (double)((threadIdx.x + 1) & 31);
    sm[threadIdx.x + 32] = 0;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) idx = (int)sm[idx];
    }
    long long t1 = clock64();
    out[threadIdx.x + 1] = idx;
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 8);
}
template <int CH> static void run(double* d, int blocks, int threads, int active, const char* tag) {
    k_dfma<CH><<<blocks, threads>>>(d, 2000, active, 1.0, 0.999999);
    cudaDeviceSynchronize();
    k_dfma<CH><<<blocks, threads>>>(d, 20000, active, 1.0, 0.999999);
    cudaDeviceSynchronize();
    double h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s chains=%d warps/SM=%d active_lanes=%2d : %.2f cycles per DFMA warp-instruction\n", tag, CH, threads / 32, active, h);
}
int main() {
    double* d;
    cudaMalloc(&d, 1 << 24);
    // one warp per scheduler: 128 threads per block, 1 block per SM (148 blocks)
    run<1>(d, 148, 128, 32, "latency");
    run<2>(d, 148, 128, 32, "ilp2");
    run<4>(d, 148, 128, 32, "ilp4");
    run<8>(d, 148, 128, 32, "ilp8");
    run<8>(d, 148, 128, 16, "ilp8 half warp");
    run<8>(d, 148, 128, 8, "ilp8 quarter warp");
    run<8>(d, 148, 128, 4, "ilp8 4 lanes");
    run<1>(d, 148, 128, 16, "latency half warp");
    // two warps per scheduler
    run<8>(d, 148, 256, 32, "ilp8 2 warps/sched");
    run<8>(d, 148, 256, 16, "ilp8 2 half warps/sched");
    run<4>(d, 148, 256, 16, "ilp4 2 half warps/sched");
    run<1>(d, 148, 256, 32, "lat 2 warps/sched");
    run<1>(d, 148, 512, 32, "lat 4 warps/sched");
    run<1>(d, 148, 1024, 32, "lat 8 warps/sched");
    k_shfl<<<1, 32>>>(d, 20000, 1.0); cudaDeviceSynchronize();
    double h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("shfl.bfly(64-bit)+dadd dependent round trip: %.1f cycles\n", h);
    k_lds<<<1, 32>>>(d, 20000); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("lds.64 + f2i dependent round trip: %.1f cycles\n", h);
    return 0;
}
