// avs_bwd_kernel.cuh — reverse-sweep (BPTT) kernel templates (instantiated in avs_k_bwd_*.cu)
//
// They replace the CppAD tape + ADFun::Reverse(1) of Trainer::trainTrajectory (/root/reference/src/Trainer.cpp:607-657).
// A CTA is two warps serving the same 8 vehicles x 4 wheel lanes.  The loop-carried adjoints (LAM, YB, KB) and the staged state
// records are vehicle-level data shared by the four lanes of a vehicle: they live in shared memory as 8-vehicle tiles (seven
// 16-byte pairs per vehicle, read by broadcast).  At the end the consumer sums the 4 wheel lanes of a vehicle (shuffles) and
// writes gps[104][N], the producer writes loss[N].
//
//   rollout_bwd2_kernel: tire adjoint split between the warps.
//     * producer — state-tile stream (two contiguous cp.async one stage ahead), adjoint of the base link, the ABA wheel passes,
//       contact geometry / terrain gradient, quaternion normalisations, inlined into the stage loop; through the tire network only
//       J^T F_bar with the 2x2 Jacobian the consumer formed ahead of it; stages 5 values per wheel for the weight gradient;
//     * consumer — activation-record stream (own cp.async ring of kTiles tiles), Jacobians kLook stages AHEAD of the producer,
//       layered back-propagation + three outer products of the current stage into dL/dW (88 accumulators per lane in registers,
//       the 16 of dW4 in shared memory).
//   (The previous cut — producer runs the whole tire adjoint and stages the 37 outer-product operands, consumer only
//   accumulates — took 14.9 ms against 12.3 ms at N = 4096 x T = 200 and was removed)
#pragma once
#include "avs_dev.cuh"

namespace avs {

constexpr int kBwdThreads = 2 * kBwdBlock;  // producer warp + consumer warp

// ---- split tire adjoint ------------------------------------------------------------------------------------------------
// The producer's state chain uses the xy-net Jacobian the consumer formed ahead of it (3x3 worth of FMAs instead of the layered
// back-propagation, 5 staged values instead of 37, no activation traffic); the consumer owns the activation stream, forms J,
// runs the layered back-propagation for the weight gradient and accumulates the outer products.
// shared memory, in columns of kBwdBlock doubles: adjoint tiles LAM | YB | KB (3 x kCkTile doubles) 0 | state tile A 11 |
// state tile B 15 | activation tiles 19 (kTiles x kActLen columns) | J ring | F ring | dW4
constexpr int kB2SlotRecA = 11, kB2SlotRecB = 15, kB2SlotT = 19, kLook = kRing - 2, kTiles = kRing, kB2SlotJ = kB2SlotT + kTiles * kActLen, kB2SlotF = kB2SlotJ + kRing * kJLen, kB2SlotG4 = kB2SlotF + kRing * kFLen, kB2Slots = kB2SlotG4 + 16;  // + dW4 accumulators (16 per lane, as 8 pairs)

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
    if (threadIdx.x < kBwdBlock) {
        // ---------------- producer warp: state adjoint
        double loss;
        int stage_counter = 0;
        const int voff = 2 * (n % kCkTileVeh);  // this lane's vehicle inside the CTA's state tile (tail lanes: the mirrored vehicle)
        auto mk_odeb = [lane0, voff, &stage_counter](double ul, double ur) {
            return OdeBwdCall2<TERR, kBwdBlock>{lane0, voff, kB2SlotRecA, kB2SlotRecB, 0, kB2SlotJ, kB2SlotF, ul, ur, &stage_counter};
        };
        static_assert(kBwdBlock == 4 * kCkTileVeh, "a reverse CTA serves exactly one state tile");
        const size_t rstride = ck_rec_stride(a.Nc);
        CkptStreamRec<kBwdBlock> st{a.ckpt + (size_t)nrec * rstride + (size_t)blockIdx.x * kCkTile, rstride, nrec, kB2SlotRecA, kB2SlotRecB, 0, voff, a.ckpt, a.ckpt_doubles};
        st.init();
        static_assert(3 * kCkTile <= kB2SlotRecA * kBwdBlock && kB2SlotRecA * kBwdBlock + kCkTile <= kB2SlotRecB * kBwdBlock && kB2SlotRecB * kBwdBlock + kCkTile <= kB2SlotT * kBwdBlock, "tile map");
        rollout_bwd_body(a, n, loss, rec_col<kBwdBlock>(0, voff), st, mk_odeb);
        if (writer) a.loss[n] = loss;
    } else {
        // ---------------- consumer warp: activation stream, Jacobians one stage ahead, weight gradient
        double g[kOffW4];
#pragma unroll
        for (int k = 0; k < kOffW4; k++) g[k] = 0.0;
        // dL/dW4 (16 of the 104) lives in shared memory as eight 16-byte pairs per lane: the 88 others take 176 registers and
        // leave the routines below enough room not to spill
        double* const g4 = g_smem + (size_t)kB2SlotG4 * kBwdBlock + (size_t)lane * 2;
#pragma unroll
        for (int i = 0; i < 8; i++) *reinterpret_cast<double2*>(g4 + (size_t)i * kBwdBlock * 2) = make_double2(0.0, 0.0);
        const size_t astride = act_stage_doubles(a.Nc);
        const double* act_last = a.act + (size_t)(nrec - 1) * astride + act_lane_offset(gid);  // reverse stage 0 = last forward stage
        double* const tiles = g_smem + (size_t)kB2SlotT * kBwdBlock + (size_t)lane * 2;
        double* const jring = g_smem + (size_t)kB2SlotJ * kBwdBlock + (size_t)lane * 2;
        const double* const fring = g_smem + (size_t)kB2SlotF * kBwdBlock + lane;
        auto tile_of = [tiles](int t) { return tiles + (size_t)t * kActLen * kBwdBlock; };
        auto prefetch = [&](long j, int t) {  // activation record of reverse stage j -> tile t
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
        // The accumulators take 208 of the 255 registers, so both consumer routines stream the activations from the tile
        // (one 16-byte pair at a time) instead of holding them, and work one output row / one layer at a time.
        auto publish_jac = [&](long j, int t) {  // J of reverse stage j from tile t -> J ring, then signal
            const ActInSmem<kBwdBlock> rec{tile_of(t)};
            double jr[2][2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                double v[8];
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double ha, hb;
                    rec.load2(4 + jj, ha, hb);  // h2[2jj], h2[2jj+1]
                    const double ua = fma(-ha, ha, 1.0) * c_mc.W4[r][2 * jj], ub = fma(-hb, hb, 1.0) * c_mc.W4[r][2 * jj + 1];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = (jj == 0) ? c_mc.W2[0][k] * ua : fma(c_mc.W2[2 * jj][k], ua, v[k]);
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = fma(c_mc.W2[2 * jj + 1][k], ub, v[k]);
                }
                double a0 = 0.0, a2 = 0.0, b0 = 0.0, b2 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    double ha, hb;
                    rec.load2(kk, ha, hb);  // h0[2kk], h0[2kk+1]
                    const double wa = v[2 * kk] * fma(-ha, ha, 1.0), wb = v[2 * kk + 1] * fma(-hb, hb, 1.0);
                    a0 = fma(c_mc.W0[2 * kk][0], wa, a0); a2 = fma(c_mc.W0[2 * kk][2], wa, a2);
                    b0 = fma(c_mc.W0[2 * kk + 1][0], wb, b0); b2 = fma(c_mc.W0[2 * kk + 1][2], wb, b2);
                }
                const double os = (r == 0) ? kOutStd0 : kOutStd1;
                jr[r][0] = -(a0 + b0) * (os / kInStd1);
                jr[r][1] = (a2 + b2) * (os / kInStd3);
            }
            double dzg, pad, dax = 0.0, day = 0.0;
            rec.load2(8, dzg, pad);
            if (TERR != TERR_FLAT) rec.load2(9, dax, day);
            double* dst = jring + (size_t)(j & (kRing - 1)) * kJLen * kBwdBlock;
            *reinterpret_cast<double2*>(dst) = make_double2(jr[0][0], jr[1][0]);                  // jxx jyx
            *reinterpret_cast<double2*>(dst + 2 * kBwdBlock) = make_double2(jr[0][1], jr[1][1]);  // jxy jyy
            *reinterpret_cast<double2*>(dst + 4 * kBwdBlock) = make_double2(dzg, 0.0);
            *reinterpret_cast<double2*>(dst + 6 * kBwdBlock) = make_double2(dax, day);
            bar_arrive64(kBarEmpty + (int)(j & (kRing - 1)));
        };
        // The consumer runs kLook = kRing - 2 stages ahead with J: J depends on the stored activations only, so while the producer
        // is held up (every 40th stage it also evaluates the row loss: ~1.7 stage times) the consumer keeps forming Jacobians, and
        // afterwards the producer finds them ready and catches up while the consumer works off the weight gradients.  With a
        // lead of one stage (round 1) each warp waited on the other in turn (5 % / 12 % of their time).  Ring safety: J(i) is
        // published after F(i - kLook - 1) has been consumed (at least kRing stages would do), F(i) is staged after J(i) has
        // been read.  Stage j lives in tile j mod kTiles (kTiles = kLook + 2 = kRing, a power of two: no index registers).
#pragma unroll 1
        for (int i = 0; i <= kLook; i++) prefetch(i, i);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
        for (int i = 0; i < kLook; i++) if (i < nrec) publish_jac(i, i);
#pragma unroll 1
        for (long j = 0; j < nrec; j++) {
            const int t0 = (int)(j & (kTiles - 1));
            prefetch(j + kLook + 1, (int)((j + kLook + 1) & (kTiles - 1)));
            if (j + kLook < nrec) publish_jac(j + kLook, (int)((j + kLook) & (kTiles - 1)));
            const int q = (int)(j & (kRing - 1));
            bar_sync64(kBarFull + q);
            const double* f = fring + (size_t)q * kFLen * kBwdBlock;
            const double fxb = f[0], fyb = f[kBwdBlock], s1 = f[2 * kBwdBlock], s2 = f[3 * kBwdBlock], s3 = f[4 * kBwdBlock];
            {
                const ActInSmem<kBwdBlock> rec{tile_of(t0)};
                const double o0 = fxb * kOutStd0, o1 = fyb * kOutStd1;
                // layer 4: dW4 += o (x) h2 ; delta2 = (1 - h2^2) . W4^T o
                double d2[8];
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double ha, hb;
                    rec.load2(4 + jj, ha, hb);
                    {   // pair jj = (dW4[0][2jj], dW4[0][2jj+1]), pair 4 + jj = row 1
                        double2* q0 = reinterpret_cast<double2*>(g4 + (size_t)jj * kBwdBlock * 2);
                        double2* q1 = reinterpret_cast<double2*>(g4 + (size_t)(4 + jj) * kBwdBlock * 2);
                        double2 a0 = *q0, a1 = *q1;
                        a0.x = fma(o0, ha, a0.x); a0.y = fma(o0, hb, a0.y);
                        a1.x = fma(o1, ha, a1.x); a1.y = fma(o1, hb, a1.y);
                        *q0 = a0; *q1 = a1;
                    }
                    d2[2 * jj] = fma(c_mc.W4[1][2 * jj], o1, c_mc.W4[0][2 * jj] * o0) * fma(-ha, ha, 1.0);
                    d2[2 * jj + 1] = fma(c_mc.W4[1][2 * jj + 1], o1, c_mc.W4[0][2 * jj + 1] * o0) * fma(-hb, hb, 1.0);
                }
                // layer 2: dW2 += delta2 (x) h0 ; delta0 = (1 - h0^2) . W2^T delta2   (h0 streamed pair by pair)
                double d0[8];
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    double ha, hb;
                    rec.load2(kk, ha, hb);
                    double sa = c_mc.W2[0][2 * kk] * d2[0], sb = c_mc.W2[0][2 * kk + 1] * d2[0];
                    g[kOffW2 + 2 * kk] = fma(d2[0], ha, g[kOffW2 + 2 * kk]); g[kOffW2 + 2 * kk + 1] = fma(d2[0], hb, g[kOffW2 + 2 * kk + 1]);
#pragma unroll
                    for (int jx = 1; jx < 8; jx++) {
                        sa = fma(c_mc.W2[jx][2 * kk], d2[jx], sa); sb = fma(c_mc.W2[jx][2 * kk + 1], d2[jx], sb);
                        g[kOffW2 + jx * 8 + 2 * kk] = fma(d2[jx], ha, g[kOffW2 + jx * 8 + 2 * kk]);
                        g[kOffW2 + jx * 8 + 2 * kk + 1] = fma(d2[jx], hb, g[kOffW2 + jx * 8 + 2 * kk + 1]);
                    }
                    d0[2 * kk] = sa * fma(-ha, ha, 1.0);
                    d0[2 * kk + 1] = sb * fma(-hb, hb, 1.0);
                }
                // layer 0: dW0 += delta0 (x) (s1, s2, s3)
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    g[kOffW0 + k * 3 + 0] = fma(d0[k], s1, g[kOffW0 + k * 3 + 0]);
                    g[kOffW0 + k * 3 + 1] = fma(d0[k], s2, g[kOffW0 + k * 3 + 1]);
                    g[kOffW0 + k * 3 + 2] = fma(d0[k], s3, g[kOffW0 + k * 3 + 2]);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
#pragma unroll
        for (int k = 0; k < kOffW4; k++) {
            const double v = Group<1>::sum(g[k]);
            if (writer) a.gps[(size_t)k * a.N + n] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double2 p = *reinterpret_cast<const double2*>(g4 + (size_t)i * kBwdBlock * 2);
            const int k0 = kOffW4 + (i < 4 ? 2 * i : 8 + 2 * (i - 4));
            const double vx = Group<1>::sum(p.x), vy = Group<1>::sum(p.y);
            if (writer) { a.gps[(size_t)k0 * a.N + n] = vx; a.gps[(size_t)(k0 + 1) * a.N + n] = vy; }
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
