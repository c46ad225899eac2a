// avs_fwd_kernel.cuh — forward rollout kernel template (instantiated in avs_k_fwd_*.cu)
#pragma once
#include "avs_dev.cuh"

namespace avs {

template <int TIRE, int TERR, bool STORE>
__global__ void __launch_bounds__(kFwdBlock) rollout_fwd_kernel(const __grid_constant__ RolloutArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;  // tail lanes mirror the last vehicle so that group shuffles stay converged
    const size_t stage_stride = (size_t)4 * a.Nc * kActLen;
    auto mk_ode = [lane0, stage_stride](double ul, double ur) { return OdeFwdCall<TIRE, TERR, kFwdBlock, STORE && TIRE == TIRE_NN>{lane0, 26, 39, ul, ur, nullptr, stage_stride}; };
    rollout_fwd_body<STORE>(a, n, lane0, valid, scratch_col<kFwdBlock>(0), mk_ode);
}

template <int TIRE, int TERR, bool STORE>
static cudaError_t launch_fwd_t(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    const long long threads = (long long)a.Nc * 4;
    const int grid = (int)((threads + kFwdBlock - 1) / kFwdBlock);
    const size_t smem = (size_t)kFwdScratch * kFwdBlock * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rollout_fwd_kernel<TIRE, TERR, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    rollout_fwd_kernel<TIRE, TERR, STORE><<<grid, kFwdBlock, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace avs
