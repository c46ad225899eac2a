// avs_k_mlp.cu — batched evaluation of the reference's two small auxiliary networks (sm_100a), one sample per thread:
//   MLP_BASE_NETWORK : BaseNetwork::forward (src/BaseNetwork.cpp:41-60), 3-8-8-3 with biases, 131 parameters in the reference
//                      pack order W0 (8x3), b0, W1 (8x8), b1, W2 (3x8), b2, all row-major (:62-152): (wz, vx, vy) -> (nz, fx, fy),
//                      out_k = relu(net(in / 10))_k * (-in_k) * 10.  Dead code in the reference's ODE (HybridDynamics.cpp:433, 557).
//   MLP_LEGACY_TIRE  : the orphan data/nn_pretrained.csv: 5-16-16-3 with biases + scalers, 435 numbers in file
//                      order W0 (16x5), b0, W2 (16x16), b2, W4 (3x16), b4, out_mean[3], out_std[3], in_mean[5], in_std[5]:
//                      out = net((in - in_mean) / in_std) * out_std + out_mean with tanh hidden layers (the layout
//                      scripts/pretrain.py:20-36 prints).  No reference code evaluates this file or defines its five input
//                      features, so it is offered as a plain batched function, not as a tire model of the ODE.
#include "avs_kernels.h"

namespace avs {

template <int NIN, int NH, int NOUT, int KIND>
__global__ void __launch_bounds__(128) mlp_kernel(const double* __restrict__ p, int N, const double* __restrict__ in, double* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double* W0 = p;
    const double* b0 = W0 + NH * NIN;
    const double* W1 = b0 + NH;
    const double* b1 = W1 + NH * NH;
    const double* W2 = b1 + NH;
    const double* b2 = W2 + NOUT * NH;
    const double* tail = b2 + NOUT;  // KIND 1: out_mean[3] | out_std[3] | in_mean[5] | in_std[5]
    double x[NIN], s[NIN], h0[NH], h1[NH];
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        x[i] = in[(size_t)i * N + n];
        s[i] = (KIND == MLP_BASE_NETWORK) ? x[i] * 0.1 : (x[i] - __ldg(tail + 2 * NOUT + i)) / __ldg(tail + 2 * NOUT + NIN + i);
    }
#pragma unroll
    for (int j = 0; j < NH; j++) {
        double a = __ldg(W0 + j * NIN) * s[0];
#pragma unroll
        for (int i = 1; i < NIN; i++) a += __ldg(W0 + j * NIN + i) * s[i];   // Eigen: (W x) + b
        h0[j] = tanh_f64(a + __ldg(b0 + j));
    }
#pragma unroll
    for (int j = 0; j < NH; j++) {
        double a = __ldg(W1 + j * NH) * h0[0];
#pragma unroll
        for (int i = 1; i < NH; i++) a += __ldg(W1 + j * NH + i) * h0[i];
        h1[j] = tanh_f64(a + __ldg(b1 + j));
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        double a = __ldg(W2 + k * NH) * h1[0];
#pragma unroll
        for (int i = 1; i < NH; i++) a += __ldg(W2 + k * NH + i) * h1[i];
        a += __ldg(b2 + k);
        if (KIND == MLP_BASE_NETWORK) a = ((a > 0.0 ? a : 0.0) * (-x[k])) * 10.0;  // relu * (-in) then cwiseProduct(out_std = 10)
        else a = a * __ldg(tail + NOUT + k) + __ldg(tail + k);
        out[(size_t)k * N + n] = a;
    }
}

cudaError_t launch_mlp(int kind, const double* d_params, int N, const double* d_in, double* d_out, cudaStream_t s) {
    const int grid = (N + 127) / 128;
    if (kind == MLP_BASE_NETWORK) mlp_kernel<3, 8, 3, MLP_BASE_NETWORK><<<grid, 128, 0, s>>>(d_params, N, d_in, d_out);
    else if (kind == MLP_LEGACY_TIRE) mlp_kernel<5, 16, 3, MLP_LEGACY_TIRE><<<grid, 128, 0, s>>>(d_params, N, d_in, d_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace avs
