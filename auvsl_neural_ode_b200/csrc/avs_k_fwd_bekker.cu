// avs_k_fwd_bekker.cu — forward rollout kernel instantiations (Bekker tire model, terrain selected at run time), sm_100a.
//
// rollout_fwd_kernel: one vehicle per 4-lane group (lane = wheel), state (13 live doubles + 4 joint
// positions) in registers for the whole rollout; all network weights / model constants arrive as a
// __grid_constant__ struct, i.e. constant-bank operands.  The only HBM traffic is the control stream
// (16 B per macro step), the optional x_list rows (184 B per macro step) and, when STORE, the RK4
// stage records for the reverse sweep (448 B per RK4 step) — all coalesced over the vehicle index.
#include "avs_fwd_kernel.cuh"

namespace avs {

cudaError_t launch_fwd_bekker(const ModelConst& mc, const RolloutArgs& a, cudaStream_t s) {
    if (a.margin) return launch_fwd_t<TIRE_BEKKER, TERR_DYN, false, true>(mc, nullptr, a, s);  // diagnostic: + branch margins
    return launch_fwd_t<TIRE_BEKKER, TERR_DYN, false>(mc, nullptr, a, s);
}

}  // namespace avs
