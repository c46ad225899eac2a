// avs_k_bwd_dyn.cu — reverse-sweep kernel, terrain selected at run time (sm_100a)
#include "avs_bwd_kernel.cuh"
namespace avs {
cudaError_t launch_bwd_dyn(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) { return launch_bwd_t<TERR_DYN>(mc, d_live, a, s); }
}  // namespace avs
