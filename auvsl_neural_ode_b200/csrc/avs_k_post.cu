// avs_k_post.cu — gradient post-processing (explosion filter + reduction), RMSProp and the fp64 peak probe (sm_100a)
#include "avs_kernels.h"

namespace avs {

// ---- explosion filter + reduction ------------------------------------------------------------
// mode 0 (Trainer::train, Trainer.cpp:169-193): components with |g| > thresh are zeroed, every sample counts
// mode 1 (Trainer::combineResults, Trainer.cpp:661-689): a sample with any exploding component is dropped
__global__ void explode_flag_kernel(const double* __restrict__ gps, int N, double thresh, unsigned char* __restrict__ flags) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    bool bad = false;
    for (int k = 0; k < kNumLive; k++) bad |= fabs(gps[(size_t)k * N + n]) > thresh;
    flags[n] = bad ? 1 : 0;
}

// block k < 104: grad_sum[k]; block 104: loss_sum and n_kept.  Fixed block size and a fixed-shape tree
// make the result independent of scheduling (bitwise reproducible run to run).
__global__ void __launch_bounds__(256) filter_reduce_kernel(const double* __restrict__ gps, const double* __restrict__ loss, int N, double thresh,
                                                           int mode, const unsigned char* __restrict__ flags, double* __restrict__ reduced) {
    __shared__ double sh[256], sh2[256];
    const int k = blockIdx.x;
    double s = 0.0, cnt = 0.0;
    if (k < kNumLive) {
        for (int n = threadIdx.x; n < N; n += 256) {
            double g = gps[(size_t)k * N + n];
            if (mode == 0) { if (fabs(g) > thresh) g = 0.0; }
            else if (flags[n]) g = 0.0;
            s += g;
        }
    } else {
        for (int n = threadIdx.x; n < N; n += 256) {
            const bool keep = (mode == 0) || !flags[n];
            if (keep) { s += loss[n]; cnt += 1.0; }
        }
    }
    sh[threadIdx.x] = s; sh2[threadIdx.x] = cnt;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sh[threadIdx.x] += sh[threadIdx.x + w]; sh2[threadIdx.x] += sh2[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (k < kNumLive) reduced[k] = sh[0];
        else { reduced[kNumNNParams] = sh[0]; reduced[kNumNNParams + 1] = sh2[0]; }
    }
    // networks 1..3 are never evaluated (HybridDynamics.cpp:387): their gradient is exactly 0
    if (k == kNumLive) for (int j = kNumLive + threadIdx.x; j < kNumNNParams; j += 256) reduced[j] = 0.0;
}

cudaError_t launch_filter_reduce(const double* gps, const double* loss, int N, double thresh, int mode, unsigned char* flags, double* reduced, cudaStream_t s) {
    if (mode == 1) {
        explode_flag_kernel<<<(N + 255) / 256, 256, 0, s>>>(gps, N, thresh, flags);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    filter_reduce_kernel<<<kNumLive + 1, 256, 0, s>>>(gps, loss, N, thresh, mode, flags, reduced);
    return cudaGetLastError();
}

// ---- Trainer::updateParams (Trainer.cpp:428-453) ----------------------------------------------
__global__ void rmsprop_kernel(double* __restrict__ params, double* __restrict__ sq, const double* __restrict__ reduced, double lr, double l1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kNumNNParams) return;
    const double cnt = reduced[kNumNNParams + 1];
    const double g = cnt > 0.0 ? reduced[i] / cnt : 0.0;     // updateParams(m_batch_grad / m_cnt)
    const float grad_idx = (float)g;                          // :431,435 float truncation
    const float g2 = grad_idx * grad_idx;                     // float product, then widened (ADF(grad_idx*grad_idx))
    const double s = 0.95 * sq[i] + 0.05 * (double)g2;
    sq[i] = s;
    double p = params[i];
    p -= (lr / (sqrt(s) + 1e-6)) * (double)grad_idx;
    p -= lr * l1 * p;
    params[i] = p;
}
cudaError_t launch_rmsprop(double* params416, double* sq416, const double* reduced418, double lr, double l1, cudaStream_t s) {
    rmsprop_kernel<<<(kNumNNParams + 127) / 128, 128, 0, s>>>(params416, sq416, reduced418, lr, l1);
    return cudaGetLastError();
}

// ---- fp64 peak probe: 8 independent DFMA chains per thread ---------------------------------------
__global__ void fp64_peak_kernel(int iters, double* __restrict__ sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;
}
cudaError_t launch_fp64_peak(int blocks, int threads, int iters, double* sink, cudaStream_t s) {
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace avs
