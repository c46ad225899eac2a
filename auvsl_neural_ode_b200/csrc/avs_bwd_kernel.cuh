// avs_bwd_kernel.cuh — reverse-sweep (BPTT) kernel template (instantiated in avs_k_bwd_*.cu)
//
// rollout_bwd_kernel replaces the CppAD tape + ADFun::Reverse(1) of Trainer::trainTrajectory
// (/root/reference/src/Trainer.cpp:607-657).  A CTA is two warps serving the same 8 vehicles x 4 wheel lanes:
//   * producer warp — walks the checkpoint stream backwards: per RK4 stage the state record (14 doubles) and the
//     tire-network activation record (20 doubles per wheel) arrive in shared memory by cp.async one stage ahead;
//     it back-propagates through the base link, the ABA wheel passes, the tire network, the contact geometry /
//     terrain gradient and the quaternion normalisations, keeping the loop-carried adjoints in shared-memory
//     scratch columns (conflict-free: consecutive lanes -> consecutive banks);
//   * consumer warp — holds dL/dW (104 doubles per lane) in registers and accumulates the three outer products of
//     every stage from the 37 operands the producer stages for it (double buffer, named barriers).
// At the end the consumer sums the 4 wheel lanes of a vehicle (shuffles) and writes gps[104][N]; the producer
// writes loss[N].
#pragma once
#include "avs_dev.cuh"

namespace avs {

constexpr int kBwdSlotG = kBwdScratch;  // the two staging buffers (2 x kGradOps columns) follow the scratch columns
constexpr int kBwdThreads = 2 * kBwdBlock;  // producer warp + consumer warp

template <int TERR>
__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_kernel(const __grid_constant__ RolloutArgs a) {
    const int lane = threadIdx.x & 31;
    const int gid = blockIdx.x * kBwdBlock + lane;  // 32 (vehicle, wheel) lanes per CTA
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;
    const bool writer = valid && lane0 == 0;
    const long nrec = (long)(a.T - 1) * a.substeps * 4;
    if (threadIdx.x < kBwdBlock) {
        // ---------------- producer warp: state adjoint
        double loss;
        int stage_counter = 0;
        // scratch columns: LAM 0 | YB 13 | KB 26 | record A 39 | record B 53 | activation record 67 ; staging from kBwdSlotG
        auto mk_odeb = [lane0, &stage_counter](double ul, double ur) {
            return OdeBwdCall<TERR, kBwdBlock>{lane0, 39, 53, 0, kBwdSlotG, 67, ul, ur, &stage_counter};
        };
        // the stream starts at the final-state record; its (non-existent) activation record is one past the last stage
        CkptStreamAsync<kBwdBlock> st{a.ckpt + ((size_t)nrec * a.Nc + n) * kCkRows, (size_t)a.Nc * kCkRows, nrec, 39, 53, 0,
                                      a.act + (size_t)nrec * act_stage_doubles(a.Nc) + act_lane_offset(gid), act_stage_doubles(a.Nc), 67};
        st.init();
        rollout_bwd_body(a, n, loss, scratch_col<kBwdBlock>(0), st, mk_odeb);
        if (writer) a.loss[n] = loss;
    } else {
        // ---------------- consumer warp: dL/dW in registers
        double g[kNumLive];
#pragma unroll
        for (int k = 0; k < kNumLive; k++) g[k] = 0.0;
        const double* stage0 = g_smem + (size_t)kBwdSlotG * kBwdBlock + lane;
#pragma unroll 1
        for (long i = 0; i < nrec; i++) {
            const int p = (int)(i & 1);
            bar_sync64(kBarFull + p);
            const double* src = stage0 + (size_t)p * kGradOps * kBwdBlock;
            double op[kGradOps];
#pragma unroll
            for (int k = 0; k < kGradOps; k++) op[k] = src[k * kBwdBlock];
            if (i + 2 < nrec) bar_arrive64(kBarEmpty + p);  // hand the buffer back (the producer waits only while stages remain)
            grad_outer(g, op);
        }
#pragma unroll
        for (int k = 0; k < kNumLive; k++) {
            const double v = Group<1>::sum(g[k]);
            if (writer) a.gps[(size_t)k * a.N + n] = v;
        }
    }
}

template <int TERR> static cudaError_t launch_bwd_t(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    const long long threads = (long long)a.Nc * 4;
    const int grid = (int)((threads + kBwdBlock - 1) / kBwdBlock);
    const size_t smem = (size_t)(kBwdScratch + 2 * kGradOps) * kBwdBlock * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rollout_bwd_kernel<TERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    rollout_bwd_kernel<TERR><<<grid, kBwdThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace avs
