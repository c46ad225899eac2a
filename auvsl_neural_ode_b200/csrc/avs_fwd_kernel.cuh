// avs_fwd_kernel.cuh — forward rollout kernel template (instantiated in avs_k_fwd_*.cu)
#pragma once
#include "avs_kernels.h"

namespace avs {

template <int TIRE, int TERR, bool STORE>
__global__ void __launch_bounds__(kFwdBlock) rollout_fwd_kernel(const __grid_constant__ ModelConst mc, const __grid_constant__ RolloutArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;  // tail lanes mirror the last vehicle so that group shuffles stay converged
    rollout_fwd_body<1, TIRE, TERR, STORE>(mc, a, n, lane0, valid);
}

template <int TIRE, int TERR, bool STORE>
static cudaError_t launch_fwd_t(const ModelConst& mc, const RolloutArgs& a, cudaStream_t s) {
    const long long threads = (long long)a.Nc * 4;
    const int grid = (int)((threads + kFwdBlock - 1) / kFwdBlock);
    rollout_fwd_kernel<TIRE, TERR, STORE><<<grid, kFwdBlock, 0, s>>>(mc, a);
    return cudaGetLastError();
}

}  // namespace avs
