// avs_kernels.h — launch entry points of the sm_100a kernels (internal to libavsim.so)
#pragma once
#include <cuda_runtime.h>
#include "avs_rollout.cuh"

namespace avs {

#ifndef AVS_FWD_BLOCK
#define AVS_FWD_BLOCK 128
#endif
constexpr int kFwdBlock = AVS_FWD_BLOCK;  // threads per CTA, forward rollout (scratch columns are per thread: any multiple of 32 works)
constexpr int kBwdBlock = 32;   // (vehicle, wheel) lanes per CTA of the reverse sweep; the CTA has 2 warps (producer + consumer), 42 KB smem, 255 registers: 4 CTAs per SM

// forward rollout: one vehicle per 4-lane group (lane = wheel)
cudaError_t launch_fwd_nn_flat(bool store, const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s);
cudaError_t launch_fwd_nn_dyn(bool store, const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s);
cudaError_t launch_fwd_bekker(const ModelConst& mc, const RolloutArgs& a, cudaStream_t s);
// reverse sweep (NN tire model only): reads a.ckpt, writes a.loss [Nc] and a.gps [104][stride N]
cudaError_t launch_bwd_flat(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s);
cudaError_t launch_bwd_dyn(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s);
// N independent ODE evaluations: x [21][N], u [2][N] -> xd [21][N]
cudaError_t launch_ode(int tire, const ModelConst& mc, const double* d_live, int N, const double* x, const double* u, double* xd, cudaStream_t s);
// HybridDynamics::settle
cudaError_t launch_settle(int tire, const ModelConst& mc, const double* d_live, double* d_state21, cudaStream_t s);
// explosion filter + deterministic reduction: gps [104][N], loss [N] -> reduced [418]
cudaError_t launch_filter_reduce(const double* gps, const double* loss, int N, double thresh, int mode, unsigned char* flags, double* reduced, cudaStream_t s);
// RMSProp (Trainer::updateParams)
cudaError_t launch_rmsprop(double* params416, double* sq416, const double* reduced418, double lr, double l1, cudaStream_t s);
// the reference's auxiliary networks as batched functions (avs_k_mlp.cu): in [n_in][N] -> out [3][N]
enum { MLP_BASE_NETWORK = 0, MLP_LEGACY_TIRE = 1 };
constexpr int kBaseNetParams = 131, kLegacyNetParams = 435;
cudaError_t launch_mlp(int kind, const double* d_params, int N, const double* d_in, double* d_out, cudaStream_t s);
// DFMA-chain peak probe
cudaError_t launch_fp64_peak(int blocks, int threads, int iters, double* sink, cudaStream_t s);

inline cudaError_t launch_rollout_fwd(int tire, int terr, bool store, const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    // the flat-terrain instantiation is the lean one (no terrain switch, no contact search): everything else runs on TERR_DYN
    if (tire == TIRE_NN) return (terr == TERR_FLAT && !mc.contact_search) ? launch_fwd_nn_flat(store, mc, d_live, a, s) : launch_fwd_nn_dyn(store, mc, d_live, a, s);
    if (store) return cudaErrorInvalidValue;
    return launch_fwd_bekker(mc, a, s);
}
inline cudaError_t launch_rollout_bwd(int terr, const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    return terr == TERR_FLAT ? launch_bwd_flat(mc, d_live, a, s) : launch_bwd_dyn(mc, d_live, a, s);
}

}  // namespace avs
