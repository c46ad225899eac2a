// avs_bwd_kernel.cuh — reverse-sweep (BPTT) kernel template (instantiated in avs_k_bwd_*.cu)
//
// rollout_bwd_kernel replaces the CppAD tape + ADFun::Reverse(1) of Trainer::trainTrajectory
// (/root/reference/src/Trainer.cpp:607-657).  Per RK4 step it reloads the four stage records the forward
// sweep stored (coalesced, 448 B per vehicle-step), recomputes each stage's tire-network activations in
// registers and immediately back-propagates through them.  dL/dW is accumulated per lane in a private
// shared-memory column ([104][kBwdBlock] doubles next to 100 scratch rows for the loop-carried adjoints and the
// double-buffered checkpoint records that cp.async prefetches one stage ahead,
// conflict-free: consecutive lanes -> consecutive banks), summed over the lanes of a vehicle at the end and
// written as the per-sample gradient gps[104][N].
#pragma once
#include "avs_dev.cuh"

namespace avs {

constexpr int kBwdSlotG = kBwdScratch;  // gradient accumulators follow the scratch columns

template <int TERR>
__global__ void __launch_bounds__(kBwdBlock) rollout_bwd_kernel(const __grid_constant__ RolloutArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;
    const Scratch<kBwdBlock> G = scratch_col<kBwdBlock>(kBwdSlotG);
#pragma unroll 8
    for (int k = 0; k < kNumLive; k++) G[k] = 0.0;
    double loss;
    auto mk_odeb = [lane0](double ul, double ur) { return OdeBwdCall<TERR, kBwdBlock>{lane0, 52, 26, 39, kBwdSlotG, ul, ur, 66, 80}; };
    const long nrec = (long)(a.T - 1) * a.substeps * 4;
    // the stream starts at the final-state record; its (non-existent) activation record is one past the last stage
    CkptStreamAsync<kBwdBlock> st{a.ckpt + (size_t)nrec * kCkRows * a.Nc + n, (size_t)a.Nc, nrec, 52, 66, 0,
                                  a.act + ((size_t)nrec * 4 * a.Nc + (size_t)4 * n + lane0) * kActLen, 80};
    st.init();
    rollout_bwd_body(a, n, loss, scratch_col<kBwdBlock>(0), st, mk_odeb);
    const bool writer = valid && lane0 == 0;
    for (int k = 0; k < kNumLive; k++) {
        const double g = Group<1>::sum(G[k]);
        if (writer) a.gps[(size_t)k * a.N + n] = g;
    }
    if (writer) a.loss[n] = loss;
}

template <int TERR> static cudaError_t launch_bwd_t(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    const long long threads = (long long)a.Nc * 4;
    const int grid = (int)((threads + kBwdBlock - 1) / kBwdBlock);
    const size_t smem = (size_t)(kNumLive + kBwdScratch) * kBwdBlock * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rollout_bwd_kernel<TERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    rollout_bwd_kernel<TERR><<<grid, kBwdBlock, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace avs
