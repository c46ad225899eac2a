// avs_bwd_kernel.cuh — reverse-sweep (BPTT) kernel templates (instantiated in avs_k_bwd_*.cu)
//
// They replace the CppAD tape + ADFun::Reverse(1) of Trainer::trainTrajectory (/root/reference/src/Trainer.cpp:607-657).
// A CTA is two warps serving the same 8 vehicles x 4 wheel lanes.  The loop-carried adjoints (LAM, YB, KB) and the staged state
// records are vehicle-level data shared by the four lanes of a vehicle: they live in shared memory as 8-vehicle tiles (seven
// 16-byte pairs per vehicle, read by broadcast).  At the end the consumer sums the 4 wheel lanes of a vehicle (shuffles) and
// writes gps[104][N], the producer writes loss[N].
//
//   rollout_bwd2_kernel:
//     * producer — adjoint of the base link, the ABA wheel passes,
//       contact geometry / terrain gradient, quaternion normalisations AND the layered back-propagation of the tire network
//       (ode_bwd_nn_p), inlined into the stage loop; it reads the activations from the consumer's tile of the stage and stages
//       the 21 operands of the weight-gradient outer products;
//     * consumer — BOTH record streams (cp.async rings of kTiles activation tiles and kSRing state tiles, kLook stages ahead of
//       the producer, which neither issues nor waits for a copy) and the
//       three outer products of the current stage into dL/dW: 104 accumulators per lane in registers.
//   CTAs on the upper warp slots of an SM swap the roles of their warps, so that every scheduler serves one producer and one
//   consumer (AVS_BWD_SWAP): both rollout kernels are bound by issue slots, and the two roles are unequal now.
//   History: (1) producer runs the whole tire adjoint and stages 37 operands, consumer only accumulates: 14.9 ms at N = 4096 x
//   T = 200; (2) consumer forms the 2x2 Jacobian of the xy-net a stage ahead, producer multiplies by it, consumer repeats the
//   layered pass for the weight gradient: 12.3 -> 10.65 ms after the round's tuning — but ~210 FP64 instructions per wheel and
//   stage more than (3), this one.  DESIGN.md §4 / §7b.
#pragma once
#include "avs_dev.cuh"

namespace avs {

constexpr int kBwdThreads = 2 * kBwdBlock;  // producer warp + consumer warp

// shared memory, in columns of kBwdBlock doubles: adjoint tiles LAM | YB | KB (3 x kCkTile doubles) 0 | state-tile ring 11
// (kSRing x kSTileCols columns) | activation tiles (kTiles x kActLen columns) | operand ring (kRing x kFLen columns)
constexpr int kB2SlotRec = 11, kB2SlotT = kB2SlotRec + kSRing * kSTileCols, kLook = kRing - 2, kTiles = kRing, kB2SlotF = kB2SlotT + kTiles * kActLen, kB2Slots = kB2SlotF + kRing * kFLen;
template <int TERR>
__global__ void __launch_bounds__(kBwdThreads) rollout_bwd2_kernel(const __grid_constant__ RolloutArgs a) {
    const int lane = threadIdx.x & 31;
    const int gid = blockIdx.x * kBwdBlock + lane;
    int n = gid >> 2;
    const int lane0 = gid & 3;
    const bool valid = n < a.Nc;
    if (!valid) n = a.Nc - 1;
    const bool writer = valid && lane0 == 0;
    const long nrec = (long)(a.T - 1) * a.substeps * 4;
#if AVS_BWD_SWAP
    // hardware warp slot mod 4 = scheduler, and a 2-warp CTA takes two consecutive slots (profiles/r2_ubench.txt): the CTAs on
    // slots 4..7 of an SM swap roles, so that each scheduler serves one producer and one consumer instead of two of a kind
    __shared__ int s_swap;
    if (threadIdx.x == 0) {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        s_swap = (int)((wid >> 2) & 1u);
    }
    __syncthreads();
    const bool is_producer = (threadIdx.x < kBwdBlock) != (s_swap != 0);
#else
    const bool is_producer = threadIdx.x < kBwdBlock;
#endif
    if (is_producer) {
        // ---------------- producer warp: state adjoint
        double loss;
        int stage_counter = 0;
        const int voff = 2 * (n % kCkTileVeh);  // this lane's vehicle inside the CTA's state tile (tail lanes: the mirrored vehicle)
        auto mk_odeb = [lane0, voff, &stage_counter](double ul, double ur) {
            return OdeBwdCall2<TERR, kBwdBlock>{lane0, voff, 0, kB2SlotT, kB2SlotF, ul, ur, &stage_counter};
        };
        static_assert(kBwdBlock == 4 * kCkTileVeh, "a reverse CTA serves exactly one state tile");
        static_assert(3 * kCkTile <= kB2SlotRec * kBwdBlock && kCkTile <= kSTileCols * kBwdBlock, "tile map");
        CkptStreamRing<kBwdBlock> st{0, kB2SlotRec, voff};
        bar_sync64(0);  // the consumer has fetched the final-state tile (and the tiles of the first stages)
        rollout_bwd_body(a, n, loss, rec_col<kBwdBlock>(0, voff), st, mk_odeb);
        if (writer) a.loss[n] = loss;
    } else {
        // ---------------- consumer warp: activation stream kLook stages ahead, weight gradient
        double g[kNumLive];
#pragma unroll
        for (int k = 0; k < kNumLive; k++) g[k] = 0.0;
        const size_t astride = act_stage_doubles(a.Nc);
        const double* act_last = a.act + (size_t)(nrec - 1) * astride + act_lane_offset(gid);  // reverse stage 0 = last forward stage
        double* const tiles = g_smem + (size_t)kB2SlotT * kBwdBlock + (size_t)lane * 2;
        const double* const fring = g_smem + (size_t)kB2SlotF * kBwdBlock + (size_t)lane * 2;
        auto tile_of = [tiles](int t) { return tiles + (size_t)t * kActLen * kBwdBlock; };
        // state tiles: record number k counted from the final state backwards -> slot k mod kSRing (CkptStreamRing, avs_dev.cuh);
        // lanes 0..31 fetch the first 512 bytes of the 896-byte tile, lanes 0..23 the rest
        const size_t rstride = ck_rec_stride(a.Nc);
        const double* const rec_last = a.ckpt + (size_t)nrec * rstride + (size_t)blockIdx.x * kCkTile;
        auto fetch_state = [&](long k) {
            if (k <= nrec) {
                const double* src = rec_last - (size_t)k * rstride;
                AVS_DBG_RANGE("state tile (reverse prefetch)", src, kCkTile, a.ckpt, a.ckpt_doubles);
                double* dst = g_smem + (size_t)(kB2SlotRec + (int)(k & (kSRing - 1)) * kSTileCols) * kBwdBlock;
                const unsigned s0 = (unsigned)__cvta_generic_to_shared(dst + lane * 2);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(src + lane * 2) : "memory");
                if (lane < kCkTile / 2 - 32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 512u), "l"(src + 64 + lane * 2) : "memory");
            }
        };
        auto prefetch = [&](long j, int t) {  // activation record of reverse stage j -> tile t, and the state tile that stage reads
            fetch_state(j + 1);
            if (j < nrec) {
                const double* src = act_last - (size_t)j * astride;
                AVS_DBG_RANGE("activation record (reverse prefetch)", src, 9 * kActPairStride + 2, a.act, a.act_doubles);
                double* dst = tile_of(t);
#pragma unroll
                for (int i = 0; i < (TERR == TERR_FLAT ? kActLen / 2 - 1 : kActLen / 2); i++) {  // flat: pair 9 (terrain gradient) is not stored
                    const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + (size_t)i * kBwdBlock * 2);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src + (size_t)i * kActPairStride) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // The consumer signals that the tile of a stage has landed kLook = kRing - 2 stages ahead of the stage whose weight
        // gradient it works on: while the producer is held up (every 40th stage it also evaluates the row loss: ~1.7 stage times)
        // the tiles keep arriving.  Ring safety: the tile of stage i is refilled after the operands of stage i have been consumed,
        // and those are staged after the producer has read the tile.  Stage j lives in tile j mod kTiles (kTiles = kLook + 2 =
        // kRing, a power of two: no index registers).
        fetch_state(0);
#pragma unroll 1
        for (int i = 0; i <= kLook; i++) prefetch(i, i);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        bar_sync64(0);  // final-state tile in place: the producer starts with the row loss at it
#pragma unroll 1
        for (int i = 0; i < kLook; i++) if (i < nrec) bar_arrive64(kBarEmpty + i);
#pragma unroll 1
        for (long j = 0; j < nrec; j++) {
            const int t0 = (int)(j & (kTiles - 1));
            prefetch(j + kLook + 1, (int)((j + kLook + 1) & (kTiles - 1)));
            if (j + kLook < nrec) bar_arrive64(kBarEmpty + (int)((j + kLook) & (kRing - 1)));  // its tiles landed before the last wait
            const int q = (int)(j & (kRing - 1));
            bar_sync64(kBarFull + q);
            const ActInSmem<kBwdBlock> rec{tile_of(t0)};
            const ActInSmem<kBwdBlock> ops{fring + (size_t)q * kFLen * kBwdBlock};
            {   // dW4 += (o0, o1) (x) h2
                double o0, o1;
                ops.load2(0, o0, o1);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double ha, hb;
                    rec.load2(4 + jj, ha, hb);
                    g[kOffW4 + 2 * jj] = fma(o0, ha, g[kOffW4 + 2 * jj]); g[kOffW4 + 2 * jj + 1] = fma(o0, hb, g[kOffW4 + 2 * jj + 1]);
                    g[kOffW4 + 8 + 2 * jj] = fma(o1, ha, g[kOffW4 + 8 + 2 * jj]); g[kOffW4 + 8 + 2 * jj + 1] = fma(o1, hb, g[kOffW4 + 8 + 2 * jj + 1]);
                }
            }
            {   // dW2 += d2 (x) h0   (h0 held, d2 streamed pair by pair)
                double h0[8];
#pragma unroll
                for (int kk = 0; kk < 4; kk++) rec.load2(kk, h0[2 * kk], h0[2 * kk + 1]);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double da, db;
                    ops.load2(3 + jj, da, db);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        g[kOffW2 + (2 * jj) * 8 + k] = fma(da, h0[k], g[kOffW2 + (2 * jj) * 8 + k]);
                        g[kOffW2 + (2 * jj + 1) * 8 + k] = fma(db, h0[k], g[kOffW2 + (2 * jj + 1) * 8 + k]);
                    }
                }
            }
            {   // dW0 += d0 (x) (s1, s2, s3)
                double s1, s2, s3, pad;
                ops.load2(1, s1, s2);
                ops.load2(2, s3, pad);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    double da, db;
                    ops.load2(7 + kk, da, db);
                    g[kOffW0 + (2 * kk) * 3 + 0] = fma(da, s1, g[kOffW0 + (2 * kk) * 3 + 0]);
                    g[kOffW0 + (2 * kk) * 3 + 1] = fma(da, s2, g[kOffW0 + (2 * kk) * 3 + 1]);
                    g[kOffW0 + (2 * kk) * 3 + 2] = fma(da, s3, g[kOffW0 + (2 * kk) * 3 + 2]);
                    g[kOffW0 + (2 * kk + 1) * 3 + 0] = fma(db, s1, g[kOffW0 + (2 * kk + 1) * 3 + 0]);
                    g[kOffW0 + (2 * kk + 1) * 3 + 1] = fma(db, s2, g[kOffW0 + (2 * kk + 1) * 3 + 1]);
                    g[kOffW0 + (2 * kk + 1) * 3 + 2] = fma(db, s3, g[kOffW0 + (2 * kk + 1) * 3 + 2]);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
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
    const size_t smem = (size_t)kB2Slots * kBwdBlock * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rollout_bwd2_kernel<TERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = upload_model(mc, d_live, s);
    if (e != cudaSuccess) return e;
    rollout_bwd2_kernel<TERR><<<grid, kBwdThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace avs
