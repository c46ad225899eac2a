// avs_k_misc.cu — single ODE evaluation and settle kernels (terrain selected at run time), sm_100a.
#include "avs_dev.cuh"

namespace avs {

// System::forward for N independent states: x [21][N], u [2][N] -> xd [21][N]
template <int TIRE>
__global__ void __launch_bounds__(kFwdBlock) ode_kernel(int N, const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ xd) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < N;
    if (!valid) n = N - 1;
    double X[13], Xd[13];
#pragma unroll
    for (int c = 0; c < 13; c++) X[c] = x[(size_t)live_to_ref(c) * N + n];
    const double ul = u[n], ur = u[(size_t)N + n];
    double ws[4];
#pragma unroll
    for (int j = 0; j < 4; j++) ws[j] = x[(size_t)(17 + j) * N + n];
    ode_fwd<1, TIRE, TERR_DYN>(c_mc, lane0, X, ul, ur, Xd, ws);
    if (valid && lane0 == 0) {
#pragma unroll
        for (int c = 0; c < 13; c++) xd[(size_t)live_to_ref(c) * N + n] = Xd[c];
        // HybridDynamics.cpp:615-618, 627-630
        xd[(size_t)7 * N + n] = ul; xd[(size_t)8 * N + n] = ur; xd[(size_t)9 * N + n] = ul; xd[(size_t)10 * N + n] = ur;
        xd[(size_t)17 * N + n] = 0.0; xd[(size_t)18 * N + n] = 0.0; xd[(size_t)19 * N + n] = 0.0; xd[(size_t)20 * N + n] = 0.0;
    }
}

// HybridDynamics::settle (HybridDynamics.cpp:142-190): 20 x 50 RK4 steps with u = 0 from
// (0,0,0,1 | 0,0,.16 | 0...), zeroing the world-frame x,y linear velocity before every step.
template <int TIRE>
__global__ void settle_kernel(double* __restrict__ state21) {
    const int lane0 = threadIdx.x & 3;
    const Scratch<32> X = scratch_col<32>(0), ACC = scratch_col<32>(13), T = scratch_col<32>(26), K = scratch_col<32>(39);
    const double x_init[13] = {0, 0, 0, 1, 0, 0, 0.16, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 13; c++) X[c] = x_init[c];
    OdeFwdCall<TIRE, TERR_DYN, 32, false> ode{lane0, 26, 39, 0.0, 0.0, nullptr, 0};
#pragma unroll 1
    for (int it = 0; it < 1000; it++) {
        double q[4] = {X[0], X[1], X[2], X[3]}, R[9];
        quat_to_R(q, R);
        const double w0 = X[7], w1 = X[8], w2 = X[9], v0 = X[10], v1 = X[11], v2 = X[12];
        double aw[3], lw[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            aw[r] = R[3 * r] * w0 + R[3 * r + 1] * w1 + R[3 * r + 2] * w2;
            lw[r] = R[3 * r] * v0 + R[3 * r + 1] * v1 + R[3 * r + 2] * v2;
        }
        lw[0] = 0.0; lw[1] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            X[7 + c] = R[c] * aw[0] + R[3 + c] * aw[1] + R[6 + c] * aw[2];
            X[10 + c] = R[c] * lw[0] + R[3 + c] * lw[1] + R[6 + c] * lw[2];
        }
        NullSink sink;
        double inv_carry = 1.0;
        rk4_fwd(ode, X, ACC, T, K, sink, inv_carry);
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 21; c++) state21[c] = 0.0;
#pragma unroll
        for (int c = 0; c < 13; c++) state21[live_to_ref(c)] = X[c];
    }
}

cudaError_t launch_ode(int tire, const ModelConst& mc, const double* d_live, int N, const double* x, const double* u, double* xd, cudaStream_t s) {
    const int grid = (int)(((long long)N * 4 + kFwdBlock - 1) / kFwdBlock);
    cudaError_t e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    if (tire == TIRE_NN) ode_kernel<TIRE_NN><<<grid, kFwdBlock, 0, s>>>(N, x, u, xd);
    else ode_kernel<TIRE_BEKKER><<<grid, kFwdBlock, 0, s>>>(N, x, u, xd);
    return cudaGetLastError();
}
cudaError_t launch_settle(int tire, const ModelConst& mc, const double* d_live, double* st, cudaStream_t s) {
    cudaError_t e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    const size_t smem = (kFwdScratch + kBekScratch) * 32 * sizeof(double);
    if (tire == TIRE_NN) settle_kernel<TIRE_NN><<<1, 32, smem, s>>>(st);
    else settle_kernel<TIRE_BEKKER><<<1, 32, smem, s>>>(st);
    return cudaGetLastError();
}

}  // namespace avs
