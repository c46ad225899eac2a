// avs_k_fwd_nn_dyn.cu — forward rollout kernel instantiations (NN tire model, terrain selected at run time), sm_100a.
//
// rollout_fwd_kernel: one vehicle per 4-lane group (lane = wheel), state (13 live doubles + 4 joint
// positions) in registers for the whole rollout; all network weights / model constants arrive as a
// __grid_constant__ struct, i.e. constant-bank operands.  The only HBM traffic is the control stream
// (16 B per macro step), the optional x_list rows (184 B per macro step) and, when STORE, the RK4
// stage records for the reverse sweep (448 B per RK4 step) — all coalesced over the vehicle index.
#include "avs_fwd_kernel.cuh"

namespace avs {

cudaError_t launch_fwd_nn_dyn(bool store, const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    return store ? launch_fwd_t<TIRE_NN, TERR_DYN, true>(mc, d_live, a, s) : launch_fwd_t<TIRE_NN, TERR_DYN, false>(mc, d_live, a, s);
}

}  // namespace avs
