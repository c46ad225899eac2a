// avs_dev.cuh — device-only glue shared by the rollout kernels: the per-TU __constant__ copy of the model,
// the dynamic shared-memory scratch and the NON-inlined ODE / ODE-adjoint entry points.
//
// Why non-inlined: ptxas schedules the body of a rolled loop for minimum register pressure and turns the
// eight interleaved tanh/matvec dependency chains back into serial chains (measured: 30-40 % of all FP64
// instructions issue back-to-back on the result of their predecessor, vs 1-5 % for the same code compiled as a
// stand-alone function).  With one warp per SM sub-partition (N = 4096 vehicles x 4 lanes = 512 warps on 592
// sub-partitions) instruction-level parallelism is the only latency hiding there is.
// Why __constant__ and not a kernel parameter: a non-inlined callee would have to reach a by-reference kernel
// parameter through generic loads; a module-scope __constant__ object keeps every weight a constant-bank
// operand (c[0x3][imm]) of the DFMA that uses it.
#pragma once
#include <cstddef>
#include "avs_kernels.h"

namespace avs {

extern __shared__ double g_smem[];
static __constant__ ModelConst c_mc;

template <int BLOCK> __device__ __forceinline__ Scratch<BLOCK> scratch_col(int slot) {
    return Scratch<BLOCK>{g_smem + slot * BLOCK + threadIdx.x};
}

// ---- Bekker quadrature nodes as ONE non-inlined routine (four nodes per call, both arcs) ---------------------------------
// Everything crosses the call BY VALUE: the 13 node parameters and (th, sin th, cos th) of the four nodes are register
// arguments, the twelve integrand values come back as a returned struct — the device ABI keeps all 37 doubles in registers (no
// local memory in the SASS).  Until the end of round 2 the hand-off went through 37 shared-memory scratch columns; the entry of
// the routine then waited for its 25 LDS (5 % of the kernel's stall samples on that one line) and the caller re-read 12 results
// per call: 44.6 -> 41.7 ms at 8192 x 100 rows, results bit-identical.
constexpr int kBekScratch = 0;  // scratch columns of the hand-off (none any more; the margin columns follow kFwdScratch directly)
struct BekOut12 { double fx[4], fy[4], fz[4]; };
static __device__ __noinline__ BekOut12 bekker_nodes4_call(BekNodeParams q, double t0, double t1, double t2, double t3, double s0, double s1, double s2,
                                                           double s3, double c0, double c1, double c2, double c3) {
    const double th[4] = {t0, t1, t2, t3}, s[4] = {s0, s1, s2, s3}, c[4] = {c0, c1, c2, c3};
    BekOut12 o;
    bekker_nodes_vec<4>(q, th, s, c, o.fx, o.fy, o.fz);
    return o;
}
// MARGIN: the diagnostic rollout keeps the per-lane minima of the kNumMargins branch margins in the scratch columns that
// follow the forward columns
template <int BLOCK, bool MARGIN = false> struct BekNodesCall {
    int slot;
    __device__ __forceinline__ void margin(int kind, double v) const {
        if (MARGIN) {
            const Scratch<BLOCK> M = scratch_col<BLOCK>(slot + kBekScratch);
            M[kind] = fmin(M[kind], v);
        }
    }
    __device__ __forceinline__ void operator()(const BekNodeParams& q, const double* th, const double* s, const double* c, double* fx, double* fy, double* fz) const {
#pragma unroll
        for (int g = 0; g < 3; g++) {
            const BekOut12 o = bekker_nodes4_call(q, th[4 * g], th[4 * g + 1], th[4 * g + 2], th[4 * g + 3], s[4 * g], s[4 * g + 1], s[4 * g + 2], s[4 * g + 3],
                                                  c[4 * g], c[4 * g + 1], c[4 * g + 2], c[4 * g + 3]);
#pragma unroll
            for (int i = 0; i < 4; i++) { fx[4 * g + i] = o.fx[i]; fy[4 * g + i] = o.fy[i]; fz[4 * g + i] = o.fz[i]; }
        }
    }
};

// K = f(T):  T, K = scratch slots (13 columns each); act (STORE_ACT) = this lane's activation record of the stage
template <int TIRE, int TERR, int BLOCK, bool STORE_ACT, bool MARGIN = false>
__device__ __noinline__ void ode_fwd_call(int lane0, int slotT, int slotK, double ul, double ur, double* act) {
    const Scratch<BLOCK> T = scratch_col<BLOCK>(slotT), K = scratch_col<BLOCK>(slotK);
    double X[13], Xd[13];
#pragma unroll
    for (int c = 0; c < 13; c++) X[c] = T[c];
    if (TIRE == TIRE_BEKKER) ode_fwd<1, TIRE, TERR, STORE_ACT, BekNodesCall<BLOCK, MARGIN>>(c_mc, lane0, X, ul, ur, Xd, nullptr, ActOutPtr{act}, BekNodesCall<BLOCK, MARGIN>{kFwdScratch});
    else ode_fwd<1, TIRE, TERR, STORE_ACT>(c_mc, lane0, X, ul, ur, Xd, nullptr, ActOutPtr{act});
#pragma unroll
    for (int c = 0; c < 13; c++) K[c] = Xd[c];
}
template <int TIRE, int TERR, int BLOCK, bool STORE_ACT, bool MARGIN = false> struct OdeFwdCall {
    int lane0, slotT, slotK;
    double ul, ur;
    double* act_step;         // this lane's record of stage 0 of the current RK4 step
    size_t act_stage_stride;  // doubles between consecutive stages
    __device__ __forceinline__ void operator()(Scratch<BLOCK>, Scratch<BLOCK>, int s) const {
        ode_fwd_call<TIRE, TERR, BLOCK, STORE_ACT, MARGIN>(lane0, slotT, slotK, ul, ur, STORE_ACT ? act_step + (size_t)s * act_stage_stride : nullptr);
    }
};

// ---- fused forward stage (forward-only NN kernels): one non-inlined call = evaluate K = f(T) and apply the RK4 stage update
// of HybridDynamics.cpp:470-500 (same arithmetic and order as rk4_fwd in avs_model.cuh), so that the rolled loop in the kernel
// contains nothing but four calls.  The stage input T crosses the call BY VALUE — 13 register arguments in, the next stage
// input returned as a struct, which the device ABI keeps in registers — instead of through its scratch columns: the callee no
// longer starts with 13 LDS and a wait for them (forward only 26.5 -> 25.3 ms at N = 4096 x 600 rows, 44.1 -> 41.6 ms at
// N = 16 384 x 300; round 2, last session).  X and ACC stay in their columns (by value as well: 26.1 ms — they would be live
// across the whole ODE evaluation).  Scratch slots: X 0 | ACC 13 | T 26 (written at the end of a step only) | INV 56.
// tid = threadIdx.x, handed in by the caller: the non-inlined callee would otherwise re-read the special register (S2R, a
// short-scoreboard wait at the head of every call)
struct St13 { double v[13]; };
template <int TIRE, int TERR, int BLOCK, bool MARGIN = false>
__device__ __noinline__ St13 ode_fwd_stage_call(St13 t, int tid, int lane0, double ul, double ur, int s) {
    const Scratch<BLOCK> X{g_smem + tid}, ACC{g_smem + 13 * BLOCK + tid}, T{g_smem + 26 * BLOCK + tid}, INV{g_smem + 56 * BLOCK + tid};
    double k[13];
    if (TIRE == TIRE_BEKKER) ode_fwd<1, TIRE, TERR, false, BekNodesCall<BLOCK, MARGIN>>(c_mc, lane0, t.v, ul, ur, k, nullptr, ActOutPtr{nullptr}, BekNodesCall<BLOCK, MARGIN>{kFwdScratch});
    else ode_fwd<1, TIRE, TERR, false>(c_mc, lane0, t.v, ul, ur, k, nullptr, ActOutPtr{nullptr, kActPairStride});
    const double h = kDt;
    double inv;
    St13 o;
    if (s < 3) {
        const double wgt = (s == 0) ? 1.0 : 2.0;   // k1 + 2 k2 + 2 k3 + k4
        const double cs = (s == 2) ? h : .5 * h;   // X + .5h k1, X + .5h k2, X + h k3
#pragma unroll
        for (int c = 0; c < 13; c++) {
            ACC[c] = (s == 0) ? fma(wgt, k[c], 0.0) : fma(wgt, k[c], ACC[c]);
            o.v[c] = fma(cs, k[c], X[c]);
        }
        renorm(o.v, inv);
    } else {
#pragma unroll
        for (int c = 0; c < 13; c++) o.v[c] = fma(h / 6.0, fma(1.0, k[c], ACC[c]), X[c]);
        renorm(o.v, inv);
#pragma unroll
        for (int c = 0; c < 13; c++) { X[c] = o.v[c]; T[c] = o.v[c]; }
    }
    INV[0] = inv;
    return o;
}
template <int TIRE, int TERR, int BLOCK, bool MARGIN = false> struct FusedStep {  // forward only; the instantiation that stores the adjoint's records uses FusedStepTiles
    const RolloutArgs& a;
    int n, lane0;
    bool writer;
    double ul, ur;
    __device__ __forceinline__ void operator()(size_t, double& inv_carry) const {
        const int tid = (int)threadIdx.x;
        const Scratch<BLOCK> T = scratch_col<BLOCK>(26);
        St13 t;
#pragma unroll
        for (int c = 0; c < 13; c++) t.v[c] = T[c];
#pragma unroll 1
        for (int s = 0; s < 4; s++) t = ode_fwd_stage_call<TIRE, TERR, BLOCK, MARGIN>(t, tid, lane0, ul, ur, s);
        inv_carry = scratch_col<BLOCK>(56)[0];
    }
};


// ---- forward stage on vehicle-shared tiles ---------------------------------------------------------------------------------
// The loop-carried 13-vectors of the forward sweep (X, ACC, T) are vehicle-level data that the four lanes of a vehicle hold as
// bit-identical copies.  Here they live ONCE per vehicle, in tiles of the state-record shape (kCkTile doubles per warp: pair r
// of vehicle v at r*16 + v*2): the four lanes store the same 16 bytes to the same address and read them back by broadcast (one
// shared-memory wavefront per LDS.128 / STS.128 of the warp), and the T tile with the inverse norm in its slot 13 IS the
// stage's state record, so the record store is a coalesced copy of 896 contiguous bytes by all 32 lanes instead of seven
// STG.128 of eight writer lanes in a divergent region.  Per warp: X | ACC | T | - | joint positions.
// Used by the instantiation that stores the adjoint's records (forward + store 10.89 -> 10.59 ms at N = 4096 x T = 200, round 2);
// the forward-only instantiation keeps the per-thread columns below (on tiles it is 3 % slower: 28.47 against 27.53 ms).
constexpr int kFwdTileSet = 5 * kCkTile;
struct TileCol {
    double* p;  // pair 0 of this lane's vehicle
    __device__ __forceinline__ double& operator[](int c) const { return p[(size_t)(c >> 1) * kCkPairStride + (c & 1)]; }
    __device__ __forceinline__ TileCol sub(int c0) const { return TileCol{p + (size_t)(c0 / 13) * kCkTile}; }
};
__device__ __forceinline__ TileCol fwd_tile_col(int tid) { return TileCol{g_smem + (size_t)(tid >> 5) * kFwdTileSet + ((tid & 31) >> 2) * 2}; }
template <int TERR, int BLOCK, bool STORE>
__device__ __noinline__ void ode_fwd_stage_call_tiles(int tid, int lane0, double ul, double ur, int s, double* rec_tile, double* act) {
    double* const wbase = g_smem + (size_t)(tid >> 5) * kFwdTileSet;
    double* const Xt = wbase + ((tid & 31) >> 2) * 2;
    double* const At = Xt + kCkTile;
    double* const Tt = Xt + 2 * kCkTile;
    double x[14], k[13];
#pragma unroll
    for (int i = 0; i < 7; i++) { const double2 v = *reinterpret_cast<const double2*>(Tt + i * kCkPairStride); x[2 * i] = v.x; x[2 * i + 1] = v.y; }
    if (STORE) {  // every warp of the launch owns a tile (ck_rec_stride pads the records to whole CTAs): no branch
        const int l = tid & 31;
        const double* src = wbase + 2 * kCkTile + l * 2;
        *reinterpret_cast<double2*>(rec_tile + l * 2) = *reinterpret_cast<const double2*>(src);
        if (l < kCkTile / 2 - 32) *reinterpret_cast<double2*>(rec_tile + 64 + l * 2) = *reinterpret_cast<const double2*>(src + 64);
    }
    // X and ACC are fetched BEFORE the ODE evaluation (their 28 registers stay live across it; the 2-CTA build has the room), so
    // that the stage update at the tail does not start with a shared-memory round trip (10.21 -> 10.13 ms)
    double X[14], acc[14];
#pragma unroll
    for (int i = 0; i < 7; i++) { const double2 v = *reinterpret_cast<const double2*>(Xt + i * kCkPairStride); X[2 * i] = v.x; X[2 * i + 1] = v.y; }
    if (s > 0) {
#pragma unroll
        for (int i = 0; i < 7; i++) { const double2 v = *reinterpret_cast<const double2*>(At + i * kCkPairStride); acc[2 * i] = v.x; acc[2 * i + 1] = v.y; }
    } else {
#pragma unroll
        for (int c = 0; c < 14; c++) acc[c] = 0.0;
    }
    ode_fwd<1, TIRE_NN, TERR, STORE>(c_mc, lane0, x, ul, ur, k, nullptr, ActOutPtr{act, kActPairStride});
    const double h = kDt;
    double inv;
    if (s < 3) {
        const double wgt = (s == 0) ? 1.0 : 2.0;   // k1 + 2 k2 + 2 k3 + k4
        const double cs = (s == 2) ? h : .5 * h;   // X + .5h k1, X + .5h k2, X + h k3
#pragma unroll
        for (int c = 0; c < 13; c++) acc[c] = fma(wgt, k[c], acc[c]);
#pragma unroll
        for (int i = 0; i < 7; i++) *reinterpret_cast<double2*>(At + i * kCkPairStride) = make_double2(acc[2 * i], acc[2 * i + 1]);
#pragma unroll
        for (int c = 0; c < 13; c++) x[c] = fma(cs, k[c], X[c]);
        renorm(x, inv);
    } else {
#pragma unroll
        for (int c = 0; c < 13; c++) x[c] = fma(h / 6.0, fma(1.0, k[c], acc[c]), X[c]);
        renorm(x, inv);
#pragma unroll
        for (int i = 0; i < 6; i++) *reinterpret_cast<double2*>(Xt + i * kCkPairStride) = make_double2(x[2 * i], x[2 * i + 1]);
        *reinterpret_cast<double2*>(Xt + 6 * kCkPairStride) = make_double2(x[12], 0.0);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) *reinterpret_cast<double2*>(Tt + i * kCkPairStride) = make_double2(x[2 * i], x[2 * i + 1]);
    *reinterpret_cast<double2*>(Tt + 6 * kCkPairStride) = make_double2(x[12], inv);
    __syncwarp();  // the T tile is copied out by all lanes at the head of the next stage
}
template <int TERR, int BLOCK, bool STORE> struct FusedStepTiles {
    const RolloutArgs& a;
    int n, lane0;
    bool writer;
    double ul, ur;
    __device__ __forceinline__ void operator()(size_t step, double& inv_carry) const {
        const size_t NK = (size_t)a.Nc;
        const size_t rec_stride = ck_rec_stride(NK);
        const size_t gidx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
        double* rec = STORE ? a.ckpt + (step * 4) * rec_stride + (gidx >> 5) * kCkTile : nullptr;  // every warp of the launch owns a tile
        const size_t act_stride = act_stage_doubles(NK);
        double* act = STORE ? a.act + (step * 4) * act_stride + act_lane_offset(gidx) : nullptr;
        const int tid = (int)threadIdx.x;
        if (STORE) {
            AVS_DBG_RANGE("state tiles of a step (store)", rec, 3 * rec_stride + kCkTile, a.ckpt, a.ckpt_doubles);
            AVS_DBG_RANGE("activation records of a step (store)", act, 3 * act_stride + 9 * kActPairStride + 2, a.act, a.act_doubles);
        }
#pragma unroll 1
        for (int s = 0; s < 4; s++)
            ode_fwd_stage_call_tiles<TERR, BLOCK, STORE>(tid, lane0, ul, ur, s, rec ? rec + (size_t)s * rec_stride : nullptr, STORE ? act + (size_t)s * act_stride : nullptr);
        inv_carry = fwd_tile_col(tid).sub(26)[13];
    }
};

// ---- named barriers of the warp-specialised reverse kernel (producer warp + consumer warp, 64 participants) -----------
#ifndef AVS_BWD_RING
#define AVS_BWD_RING 4
#endif
constexpr int kRing = AVS_BWD_RING;  // depth of the J and F rings (power of two): the consumer forms J kRing - 2 stages ahead
static_assert(kRing >= 4 && (kRing & (kRing - 1)) == 0 && 2 * kRing + 1 <= 16, "ring depth");
constexpr int kBarFull = 1, kBarEmpty = 1 + kRing;
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// 14-double record staged in shared memory as seven 16-byte pairs: element c of this lane at
// base[((c >> 1) * BLOCK) * 2 + (c & 1)]  (base already includes lane * 2)
template <int BLOCK> struct PairCol {
    double* base;
    __device__ __forceinline__ double& operator[](int c) const { return base[(size_t)(c >> 1) * BLOCK * 2 + (c & 1)]; }
    __device__ __forceinline__ void load2(int i, double& a, double& b) const {
        const double2 v = *reinterpret_cast<const double2*>(base + (size_t)i * BLOCK * 2);
        a = v.x; b = v.y;
    }
    __device__ __forceinline__ void store2(int i, double a, double b) const { *reinterpret_cast<double2*>(base + (size_t)i * BLOCK * 2) = make_double2(a, b); }
    // the 13-vector that starts c0 scalar slots later (vectors are padded to 7 pairs = 14 slots): sub(13) = next vector
    __device__ __forceinline__ PairCol sub(int c0) const { return PairCol{base + (size_t)(c0 / 13) * 14 * BLOCK}; }
};

// activation record staged in shared memory as ten 16-byte pairs: pair i of this lane at base[i * BLOCK * 2]
template <int BLOCK> struct ActInSmem {
    const double* base;
    __device__ __forceinline__ void load2(int i, double& a, double& b) const {
        const double2 v = *reinterpret_cast<const double2*>(base + (size_t)i * BLOCK * 2);
        a = v.x; b = v.y;
    }
    __device__ __forceinline__ ActInSmem wheel(int) const { return *this; }
};
// ---- reverse kernel (rollout_bwd2_kernel): hand-over between the producer warp and the consumer warp ----------------------
// The consumer owns the activation-record stream (a cp.async ring of kRing warp tiles, stage j in tile j mod kRing) and runs
// kRing - 2 stages AHEAD of the producer with it.  The producer reads the activations of its stage straight from that tile,
// runs the layered back-propagation of the tire network on its state chain (ode_bwd_nn_p) and stages the operands of the
// three weight-gradient outer products for the consumer: o0 o1 | s1 s2 | s3 - | d2[8] | d0[8] = kFLen doubles as eleven
// 16-byte pairs, a ring kRing deep (slot q = stage mod kRing).
// Named barriers: kBarEmpty + q = "the activation tile of a slot-q stage has landed" (consumer arrives, producer syncs),
//                 kBarFull + q  = "the operands of a slot-q stage are staged" (producer arrives, consumer syncs).
// Because the consumer refills the tile of stage i only after it has consumed the operands of stage i, and the producer stages
// them only after it has read the tile, those two hand-shakes also protect the reuse of the tiles and of the operand ring.
// (Until the end of round 2 the consumer formed the 2x2 Jacobian of the xy-net a stage ahead and the producer used that; see
// ode_bwd_nn_p in avs_model.cuh for why the layered pass on the state chain is cheaper.)
// AVS_BWD_SWAP: the CTAs that land on the upper four warp slots of an SM swap the roles of their two warps, so that every
// scheduler serves one producer and one consumer; the producer's lane index is then threadIdx.x & 31.
#ifndef AVS_BWD_SWAP
#define AVS_BWD_SWAP 1
#endif
#if AVS_BWD_SWAP
#define AVS_BWD_LANE (threadIdx.x & 31)
#else
#define AVS_BWD_LANE threadIdx.x
#endif
constexpr int kFLen = 22;
template <int TERR, int BLOCK> struct TInSmem {
    const double* tile0;  // pair 0 of this lane in activation tile 0
    int count;
    __device__ __forceinline__ void get(int, double h0[8], double h2[8], double& dzg, double& dax, double& day) const {
        const int q = count & (kRing - 1);  // the stage head (ode_bwd_call2) has already waited for "the tiles of this stage have landed"
        const double* b = tile0 + (size_t)q * kActLen * BLOCK;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 u = *reinterpret_cast<const double2*>(b + (size_t)i * 2 * BLOCK), v = *reinterpret_cast<const double2*>(b + (size_t)(4 + i) * 2 * BLOCK);
            h0[2 * i] = u.x; h0[2 * i + 1] = u.y; h2[2 * i] = v.x; h2[2 * i + 1] = v.y;
        }
        dzg = *(b + (size_t)8 * 2 * BLOCK);
        dax = 0.0; day = 0.0;
        if (TERR != TERR_FLAT) {  // flat: the pair is neither stored nor fetched
            const double2 g = *reinterpret_cast<const double2*>(b + (size_t)9 * 2 * BLOCK);
            dax = g.x; day = g.y;
        }
    }
};
template <int BLOCK> struct FOutSmem {
    double* ring0;  // pair 0 of this lane in operand slot 0
    int count;
    __device__ __forceinline__ void put(int, double o0, double o1, double s1, double s2, double s3, const double d2[8], const double d0[8], const double*, const double*) const {
        const int q = count & (kRing - 1);
        double* dst = ring0 + (size_t)q * kFLen * BLOCK;
        *reinterpret_cast<double2*>(dst) = make_double2(o0, o1);
        *reinterpret_cast<double2*>(dst + 2 * BLOCK) = make_double2(s1, s2);
        *reinterpret_cast<double2*>(dst + 4 * BLOCK) = make_double2(s3, 0.0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            *reinterpret_cast<double2*>(dst + (size_t)(3 + i) * 2 * BLOCK) = make_double2(d2[2 * i], d2[2 * i + 1]);
            *reinterpret_cast<double2*>(dst + (size_t)(7 + i) * 2 * BLOCK) = make_double2(d0[2 * i], d0[2 * i + 1]);
        }
        bar_arrive64(kBarFull + q);
    }
};
// state tile staged in shared memory exactly as it lies in HBM (kCkTile doubles, pair r of vehicle v at r*16 + v*2): the four
// lanes of a vehicle read the same 16 bytes (broadcast), a warp's LDS.128 touches one 128-byte line
typedef PairCol<kCkTileVeh> RecCol;
template <int BLOCK> __device__ __forceinline__ RecCol rec_col(int slot, int voff /* 2 * (vehicle % 8) */) { return RecCol{g_smem + (size_t)slot * BLOCK + voff}; }
// Inlined into the producer's stage loop (round 2): unlike the forward stage — which, inlined into its rolled loop, is
// re-serialised by ptxas and runs 1.9x slower (28.1 -> 52.1 ms, re-measured) — the adjoint stage loses nothing, and the call
// cost goes away: ~40-cycle instruction-fetch misses at the call and at the return, the S2R / CgaCtaId address prologue, and
// the drain of the last STS.128 before RET (11.31 -> 10.91 ms).
template <int TERR, int BLOCK>
__device__ __forceinline__ void ode_bwd_call2(int lane0, int voff, const RecCol XS, int slotLAM, int slotJ, int slotF, double ul, double ur, int stage_count, int s) {
    // XS = the staged state tile; the loop-carried adjoints LAM | YB | KB are vehicle-level 13-vectors, bit-identical in the four
    // lanes of a vehicle (the butterfly sums are commutative), so they live in three tiles of the same shape: the four lanes
    // store the same 16 bytes to the same address and read them back by broadcast — one shared-memory wavefront per
    // LDS.128 / STS.128 of the warp instead of four, and a four times shorter store drain before the callee returns
    const RecCol LAM = rec_col<BLOCK>(slotLAM, voff), YB = LAM.sub(13), KB = LAM.sub(26);
    bar_sync64(kBarEmpty + (stage_count & (kRing - 1)));  // the consumer's cp.async of this stage's state tile and activation tile have landed
    TInSmem<TERR, BLOCK> tin{g_smem + (size_t)slotJ * BLOCK + (size_t)AVS_BWD_LANE * 2, stage_count};
    FOutSmem<BLOCK> fout{g_smem + (size_t)slotF * BLOCK + (size_t)AVS_BWD_LANE * 2, stage_count};
    double X[14], kb[14], Xb[14];
#pragma unroll
    for (int i = 0; i < 7; i++) { XS.load2(i, X[2 * i], X[2 * i + 1]); KB.load2(i, kb[2 * i], kb[2 * i + 1]); }
    ode_bwd_nn_p<1, TERR>(c_mc, lane0, X, ul, ur, kb, Xb, tin, fout);
    // rk4_bwd_stage_update on pairs
    Xb[13] = 0.0;
    const double h = kDt;
    if (s > 0) {
        renorm_bwd(X, X[13], Xb);  // stage input s = N(X + c_s k_s); X[13] = the stored inverse norm
        const double cs = (s == 3) ? h : .5 * h;
        const double wk = (s == 1) ? (h / 6.0) : (h / 3.0);
#pragma unroll
        for (int i = 0; i < 7; i++) {
            double l0, l1, y0, y1;
            LAM.load2(i, l0, l1);
            YB.load2(i, y0, y1);
            LAM.store2(i, l0 + Xb[2 * i], l1 + Xb[2 * i + 1]);
            KB.store2(i, fma(cs, Xb[2 * i], wk * y0), fma(cs, Xb[2 * i + 1], wk * y1));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 7; i++) {
            double l0, l1;
            LAM.load2(i, l0, l1);
            LAM.store2(i, l0 + Xb[2 * i], l1 + Xb[2 * i + 1]);
        }
    }
}
template <int TERR, int BLOCK> struct OdeBwdCall2 {
    int lane0, voff, slotLAM, slotJ, slotF;
    double ul, ur;
    int* stage_counter;
    template <class V> __device__ __forceinline__ void operator()(RecCol XS, V, V, V, int s, const double*) const {
        ode_bwd_call2<TERR, BLOCK>(lane0, voff, XS, slotLAM, slotJ, slotF, ul, ur, (*stage_counter)++, s);
    }
};
// The producer's view of the state-record stream.  The CONSUMER fetches the state tiles (two contiguous cp.async per tile, in the
// same commit group as the activation tile of the stage that reads it) into a ring of kSRing tiles: record number k, counted
// from the final state backwards, lives in slot k mod kSRing; reverse stage j reads record j + 1 as its stage input and, at the
// head of an RK4 step, record j (the step's output) once more.  The producer therefore neither issues nor waits for any copy
// (round-2 end; before, its own cp.async.wait_group + __syncwarp + prefetch at the head of every stage held 7 % of its stall
// samples): `advance` is an increment, and the named barrier at the head of the stage (ode_bwd_call2) covers both tiles.
// kSRing = 8: the slot of record j + 4, which the consumer refills while the producer may still be reading record j, must differ
// from the slots of records j and j + 1.
constexpr int kSRing = 8, kSTileCols = 4;  // a tile of kCkTile = 112 doubles takes four columns of kBwdBlock = 32 doubles
template <int BLOCK> struct CkptStreamRing {
    int k, slot0, voff;
    __device__ __forceinline__ const double* act() const { return nullptr; }
    __device__ __forceinline__ RecCol cur() const { return rec_col<BLOCK>(slot0 + (k & (kSRing - 1)) * kSTileCols, voff); }
    __device__ __forceinline__ void advance() { k++; }
};

// upload the model into this TU's __constant__ copy (stream-ordered, before the launch that reads it).
// d_live (optional) = device copy of the parameter vector: its first 104 entries (network 0: W0, W2, W4 in the
// reference pack order) overwrite the head of ModelConst device-to-device, so a device-side RMSProp step needs
// no host round trip before the next rollout.
static inline cudaError_t upload_model(const ModelConst& mc, const double* d_live, cudaStream_t s) {
    static_assert(offsetof(ModelConst, W0) == 0 && offsetof(ModelConst, W2) == 24 * sizeof(double) && offsetof(ModelConst, W4) == 88 * sizeof(double),
                  "ModelConst must start with W0 | W2 | W4 in the reference pack order");
    cudaError_t e = cudaMemcpyToSymbolAsync(c_mc, &mc, sizeof(ModelConst), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess || !d_live) return e;
    return cudaMemcpyToSymbolAsync(c_mc, d_live, kNumLive * sizeof(double), 0, cudaMemcpyDeviceToDevice, s);
}

}  // namespace avs
