// avs_fwd_kernel.cuh — forward rollout kernel template (instantiated in avs_k_fwd_*.cu)
#pragma once
#include "avs_dev.cuh"

namespace avs {

static_assert(kFwdBlock <= 4 * 4 * kCkTileVeh, "ck_rec_stride pads the state records to 32 vehicles: a forward CTA must not span more");

// __launch_bounds__(.., 3): the register cap of three resident CTAs (170) is what makes ptxas pick the wide schedule for the
// non-inlined stage callee (163-168 registers, 1 % back-to-back dependent FP64); without it ptxas settles on 128 registers and a
// more serial order (measured: forward+store 14.1 ms vs 11.5 ms at N = 4096 x T = 200).
// MINB = resident CTAs per SM the register cap is computed for.  The instantiation that also stores the adjoint's records runs at
// that cap; with the cap of two CTAs (188 registers; forward only: 176) it is 2 % faster while two CTAs per SM hold the whole
// launch (11.09 -> 10.89 ms at N = 4096, 18.0 -> 17.6 at 8192) and 2 % slower beyond (34.3 against 33.6 ms at 16 384):
// launch_fwd_t picks by size.
// MARGIN (Bekker only): diagnostic instantiation that also reports the branch margins of every trajectory (avs_bekker.cuh)
template <int TIRE, int TERR, bool STORE, bool MARGIN = false, int MINB = 3>
__global__ void __launch_bounds__(kFwdBlock, (TIRE == TIRE_NN ? MINB * (128 / kFwdBlock) : 0)) rollout_fwd_kernel(const __grid_constant__ RolloutArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;  // tail lanes mirror the last vehicle so that group shuffles stay converged
    const Scratch<kFwdBlock> S = scratch_col<kFwdBlock>(0);
    const bool writer = (lane0 == 0) && valid;
    if (TIRE == TIRE_NN) {
        if (STORE) {
            // forward + store: fused stages on vehicle-shared tiles (avs_dev.cuh): T starts as a copy of X, the carried inverse norm as 1
            const TileCol ST = fwd_tile_col((int)threadIdx.x);
            {
                const size_t N = (size_t)a.N;
                const TileCol TT = ST.sub(26);
#pragma unroll
                for (int c = 0; c < 13; c++) TT[c] = __ldg(a.x0 + (size_t)live_to_ref(c) * N + n);
                TT[13] = 1.0;
            }
            auto mk_step = [&a, n, lane0, writer](double ul, double ur) { return FusedStepTiles<TERR, kFwdBlock, STORE>{a, n, lane0, writer, ul, ur}; };
            rollout_fwd_body<STORE>(a, n, lane0, valid, ST, mk_step);
        } else {
            // forward only: fused stages on per-thread scratch columns: T starts as a copy of X, the carried inverse norm as 1
            {
                const size_t N = (size_t)a.N;
#pragma unroll
                for (int c = 0; c < 13; c++) S[26 + c] = __ldg(a.x0 + (size_t)live_to_ref(c) * N + n);
                S[56] = 1.0;
            }
            auto mk_step = [&a, n, lane0, writer](double ul, double ur) { return FusedStep<TIRE_NN, TERR, kFwdBlock>{a, n, lane0, writer, ul, ur}; };
            rollout_fwd_body<STORE>(a, n, lane0, valid, S, mk_step);
        }
    } else {
        const Scratch<kFwdBlock> M = scratch_col<kFwdBlock>(kFwdScratch + kBekScratch);
        if (MARGIN) {
#pragma unroll
            for (int k = 0; k < kNumMargins; k++) M[k] = 1e300;
        }
        // fused stages with the stage input by value, as in the forward-only NN branch (through GenericStep / OdeFwdCall on the
        // T and K scratch columns: 41.7 against 40.9 ms at 8192 x 100 rows): T starts as a copy of X, the carried inverse norm as 1
        {
            const size_t N = (size_t)a.N;
#pragma unroll
            for (int c = 0; c < 13; c++) S[26 + c] = __ldg(a.x0 + (size_t)live_to_ref(c) * N + n);
            S[56] = 1.0;
        }
        auto mk_step = [&a, n, lane0, writer](double ul, double ur) { return FusedStep<TIRE, TERR, kFwdBlock, MARGIN>{a, n, lane0, writer, ul, ur}; };
        rollout_fwd_body<false>(a, n, lane0, valid, S, mk_step);
        if (MARGIN) {  // minimum over the four wheel lanes of the vehicle
#pragma unroll
            for (int k = 0; k < kNumMargins; k++) {
                double m = M[k];
                m = fmin(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmin(m, __shfl_xor_sync(0xffffffffu, m, 2));
                if (writer && a.margin) a.margin[(size_t)k * a.N + n] = m;
            }
        }
    }
}

template <int TIRE, int TERR, bool STORE, bool MARGIN, int MINB>
static cudaError_t launch_fwd_k(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    const long long threads = (long long)a.Nc * 4;
    const int grid = (int)((threads + kFwdBlock - 1) / kFwdBlock);
    // forward + store keeps its loop-carried vectors in vehicle-shared tiles (one set per warp), everything else in per-thread columns
    const size_t smem = (TIRE == TIRE_NN && STORE) ? (size_t)(kFwdBlock / 32) * kFwdTileSet * sizeof(double)
                                                   : (size_t)(kFwdScratch + (TIRE == TIRE_BEKKER ? kBekScratch : 0) + (MARGIN ? kNumMargins : 0)) * kFwdBlock * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rollout_fwd_kernel<TIRE, TERR, STORE, MARGIN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    rollout_fwd_kernel<TIRE, TERR, STORE, MARGIN, MINB><<<grid, kFwdBlock, smem, s>>>(a);
    return cudaGetLastError();
}
template <int TIRE, int TERR, bool STORE, bool MARGIN = false>
static cudaError_t launch_fwd_t(const ModelConst& mc, const double* d_live, const RolloutArgs& a, cudaStream_t s) {
    if (TIRE == TIRE_NN) {
        // the build with the register cap of two CTAs per SM while two CTAs per SM hold the whole launch (forward only: 176
        // registers, 28.04 -> 27.53 ms at 4096 x 600 rows; forward + store: 188 registers), the three-CTA build beyond
        static int sms = 0;
        if (sms == 0) {
            int dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        }
        const long long ctas = ((long long)a.Nc * 4 + kFwdBlock - 1) / kFwdBlock;
        if (ctas <= 2LL * sms * (128 / kFwdBlock)) return launch_fwd_k<TIRE, TERR, STORE, MARGIN, TIRE == TIRE_NN ? 2 : 3>(mc, d_live, a, s);
    }
    return launch_fwd_k<TIRE, TERR, STORE, MARGIN, 3>(mc, d_live, a, s);
}

}  // namespace avs
