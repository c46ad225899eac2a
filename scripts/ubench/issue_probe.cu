// Does a non-FP64 instruction issue in the shadow of a DFMA's second pipe cycle?  One warp per scheduler, 8 independent DFMA
// chains per iteration, plus M independent FP32 FMAs / integer multiply-adds / shared loads per DFMA.  If the scheduler can issue
// another instruction while the half-rate FP64 pipe is busy, the cost per DFMA stays 2 cycles until M > 1; if the DFMA holds the
// issue port for both cycles it is 2 + M.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int M> __global__ void probe(double* out, int iters, double a, double b, float fa, int ia) {
    __shared__ double sm[32 * 8];
    double x[8];
    float f[8 * (M > 0 ? M : 1)];
    int n[8 * (M > 0 ? M : 1)];
#pragma unroll
    for (int c = 0; c < 8; c++) { x[c] = a + c + threadIdx.x; sm[c * 32 + (threadIdx.x & 31)] = c; }
#pragma unroll
    for (int c = 0; c < 8 * (M > 0 ? M : 1); c++) { f[c] = fa + c; n[c] = ia + c; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                x[c] = fma(x[c], b, a);
#pragma unroll
                for (int m = 0; m < M; m++) {
                    if (KIND == 1) f[c * M + m] = fmaf(f[c * M + m], fa, fa);
                    if (KIND == 2) n[c * M + m] = n[c * M + m] * ia + ia;
                    if (KIND == 3) n[c * M + m] += (int)__double_as_longlong(sm[((n[c * M + m] + c) & 7) * 32 + (threadIdx.x & 31)]);
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += x[c];
#pragma unroll
    for (int c = 0; c < 8 * (M > 0 ? M : 1); c++) s += f[c] + n[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 4 * 8);
}
template <int KIND, int M> static void run(double* d, const char* tag) {
    probe<KIND, M><<<148, 128>>>(d, 20000, 1.0, 0.999999, 0.999f, 3);
    cudaDeviceSynchronize();
    double h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s : %.2f cycles per DFMA (+%d companion instructions)\n", tag, h, M);
}
int main() {
    double* d;
    cudaMalloc(&d, 1 << 22);
    run<0, 0>(d, "DFMA only");
    run<1, 1>(d, "DFMA + 1 FFMA");
    run<1, 2>(d, "DFMA + 2 FFMA");
    run<1, 3>(d, "DFMA + 3 FFMA");
    run<2, 1>(d, "DFMA + 1 IMAD");
    run<2, 2>(d, "DFMA + 2 IMAD");
    run<3, 1>(d, "DFMA + 1 (LDS + IADD)");
    return 0;
}
