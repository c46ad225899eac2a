// avs_k_bwd_flat.cu — reverse-sweep kernel, flat terrain (sm_100a)
#include "avs_bwd_kernel.cuh"
namespace avs {
cudaError_t launch_bwd_flat(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) { return launch_bwd_t<TERR_FLAT>(mc, d_live, a, s); }
}  // namespace avs
