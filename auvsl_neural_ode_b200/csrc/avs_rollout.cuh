// avs_rollout.cuh — per-vehicle rollout drivers (forward sweep, reverse sweep) shared by the CUDA
// kernels (avs_kernels.cu) and the CPU unit-test harness (tests/hostcheck).
//
// Replaces, per vehicle (paths relative to /root/reference):
//   forward : Trainer::evaluateTrajectory / trainTrajectory rollout loops (src/Trainer.cpp:550-575,
//             617-639) -> VehicleSystem::integrate (src/VehicleSystem.cpp:113-148) -> RK4
//   reverse : the CppAD tape + ADFun::Reverse(1) of Trainer::trainTrajectory (src/Trainer.cpp:607-657)
//
// HBM data layout (all fp64, SoA so that lane-adjacent vehicles coalesce; n = vehicle index):
//   x0      [21][N]          initial state, reference layout (HybridDynamics.h:65-69)
//   ctrl    [T-1][2][N]      (vl, vr) applied during macro step i -> i+1 (= traj[i].vl/vr, Trainer.cpp:564-565)
//   x_list  [T][23][N]       optional: the reference's x_list (row 0 = x0 + controls of row 0)
//   x_final [21][N]          optional
//   ckpt    [S*4+1][tiles][7][8][2]  one record per RK4 stage (S = (T-1)*substeps steps) + one for the final state:
//                            13 state entries + the inverse quaternion norm the reverse sweep needs with it, as seven
//                            16-byte pairs.  The records of the 8 vehicles of a warp form one 896-byte "state tile",
//                            pair-major (pair r of vehicle v at r*16 + v*2 doubles): the forward warp's eight writer
//                            lanes store one full 128-byte line per pair, the reverse warp fetches the tile with two
//                            contiguous cp.async.  This stream replaces the tape (448 B per vehicle-step)
//   act     [S*4][4*Nc][20]  per-wheel tire-network activation records of every stage (2560 B per vehicle-step):
//                            the reverse sweep reads them instead of re-evaluating tanh / terrain
//   gt      [T][3][N]        (yaw, x, y) ground truth rows
//   loss    [N], gps [104][N] per-sample loss and gradient w.r.t. the live tire-network weights
#pragma once
#include "avs_model.cuh"

namespace avs {

struct RolloutArgs {
    int N;   // stride (total vehicles) of the caller-facing arrays; pointers are pre-offset to the chunk start
    int Nc;  // vehicles handled by this launch (chunk) = stride of ckpt
    int T, substeps;
    const double* x0;
    const double* ctrl;
    double* x_list;
    double* x_final;
    double* ckpt;
    double* act;  // [S*4][4*Nc][kActLen] tire-network activation records (column 4n + wheel), STORE only
    const double* gt;
    double* loss;
    double* gps;
    double* margin;  // [kNumMargins][N] branch margins (diagnostic Bekker rollout only) or nullptr
    size_t ckpt_doubles, act_doubles;  // extents of ckpt / act (checked by the AVS_DEBUG build on every record access)
};

// AVS_DEBUG build (make debug -> libavsim_debug.so): every address formed on the two record streams is checked against the
// extent of its buffer; a violation prints the site and traps.  compute-sanitizer is closed on the development pool, so this
// is the index check the hand-rolled stream arithmetic gets (tests/test_gpu_soak.py runs the all-kernels case on this build).
#if defined(AVS_DEBUG) && defined(__CUDA_ARCH__)
#define AVS_DBG_RANGE(what, ptr, len, base, total)                                                                         \
    do {                                                                                                                   \
        const double* p__ = (const double*)(ptr);                                                                          \
        if (p__ && !(p__ >= (const double*)(base) && p__ + (len) <= (const double*)(base) + (total))) {                     \
            printf("AVS_DEBUG: %s out of range: offset %lld, length %d, extent %llu (block %d thread %d)\n", what,         \
                   (long long)(p__ - (const double*)(base)), (int)(len), (unsigned long long)(total), blockIdx.x, threadIdx.x); \
            __trap();                                                                                                      \
        }                                                                                                                  \
    } while (0)
#else
#define AVS_DBG_RANGE(what, ptr, len, base, total) do { } while (0)
#endif

AVS_HD double ld(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

constexpr int kCkRows = 14;  // 13 live states + 1 inverse norm per stage record
// state tiles: the records of 8 consecutive vehicles (one warp of either rollout kernel), pair-major
constexpr int kCkTileVeh = 8, kCkPairStride = 2 * kCkTileVeh, kCkTile = kCkRows * kCkTileVeh;  // 16 / 112 doubles
// doubles between consecutive records: whole 128-thread CTAs of the forward kernel (4 tiles), so that every one of its warps owns a
// tile and stores it unconditionally (no branch at the head of the stage callee, avs_dev.cuh)
AVS_HD size_t ck_rec_stride(size_t nc) { return ((nc + 4 * kCkTileVeh - 1) / (4 * kCkTileVeh)) * 4 * kCkTile; }
AVS_HD size_t ck_lane_offset(size_t n) { return (n / kCkTileVeh) * kCkTile + (n % kCkTileVeh) * 2; }  // pair 0 of vehicle n inside a record
// element c of the record whose pair 0 is at p
AVS_HD size_t ck_elem(int c) { return (size_t)(c >> 1) * kCkPairStride + (c & 1); }

// writes stage records straight to HBM as seven 16-byte pairs (only the writer lane of a group stores)
struct CkptSink {
    double* base;  // pair 0 of vehicle n in record step*4
    size_t rec_stride;  // doubles between consecutive records = ck_rec_stride(Nc)
    bool on;
    template <class V> AVS_HD void stage(int s, const V& T, double inv) {
        if (!on) return;
        ckpt_store(base + (size_t)s * rec_stride, T, inv);
    }
    template <class V> AVS_HD static void ckpt_store(double* p, const V& T, double inv) {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int i = 0; i < 6; i++) *reinterpret_cast<double2*>(p + (size_t)i * kCkPairStride) = make_double2(T[2 * i], T[2 * i + 1]);
        *reinterpret_cast<double2*>(p + (size_t)6 * kCkPairStride) = make_double2(T[12], inv);
#else
        for (int c = 0; c < 13; c++) p[ck_elem(c)] = T[c];
        p[ck_elem(13)] = inv;
#endif
    }
};

// reverse record stream on the host / generic path: plain loads, no prefetch
template <class St> struct CkptStreamSimple {
    const double* rec;  // current record in HBM (pair 0 of this vehicle)
    size_t rec_stride;  // doubles between consecutive records
    St buf;             // 14-slot scratch column
    const double* actp; // activation records of the current record's stage (lane-offset applied)
    size_t act_stride;  // doubles between the activation records of consecutive stages
    AVS_HD const double* act() const { return actp; }
    AVS_HD void load() {
#pragma unroll
        for (int c = 0; c < kCkRows; c++) buf[c] = ld(rec + ck_elem(c));
    }
    AVS_HD St cur() const { return buf; }
    AVS_HD void advance() { rec -= rec_stride; actp -= act_stride; load(); }
};

// live-state index -> reference 21-state index
AVS_HD constexpr int live_to_ref(int c) { return c < 7 ? c : c + 4; }

constexpr int kFwdScratch = 57;  // X[13] | ACC[13] | T[13] | K[13] | joint positions[4] | inverse norm carried between stages[1]

// One RK4 step of one vehicle on its scratch columns, generic version: rk4_fwd around an ODE functor (host harness:
// inline ODE; device Bekker / settle: non-inlined ODE call).  The fused device step of the NN kernels lives in
// avs_dev.cuh (FusedStep) and does the same arithmetic with the stage update inside the non-inlined callee.
template <bool STORE, class Ode, class St> struct GenericStep {
    Ode ode;
    const RolloutArgs& a;
    int n, lane0;
    bool writer;
    St S;
    AVS_HD void operator()(size_t step, double& inv_carry) {
        const size_t NK = (size_t)a.Nc;
        St X = S, ACC = S.sub(13), TT = S.sub(26), KK = S.sub(39);
        if (STORE) {
            ode.act_step = a.act + ((step * 4) * (4 * NK) + (size_t)4 * n + lane0) * kActLen;
            CkptSink sink{a.ckpt + (step * 4) * ck_rec_stride(NK) + ck_lane_offset(n), ck_rec_stride(NK), writer};
            rk4_fwd(ode, X, ACC, TT, KK, sink, inv_carry);
        } else {
            NullSink sink;
            rk4_fwd(ode, X, ACC, TT, KK, sink, inv_carry);
        }
    }
};

// S = per-thread scratch column with kFwdScratch slots; mk_step(ul, ur) returns the RK4-step functor for a macro step
template <bool STORE, class MkStep, class St>
AVS_HD void rollout_fwd_body(const RolloutArgs& a, int n, int lane0, bool valid, St S, MkStep mk_step) {
    const size_t N = (size_t)a.N, NK = (size_t)a.Nc;
    const bool writer = (lane0 == 0) && valid;
    St X = S, JP = S.sub(52);
#pragma unroll
    for (int c = 0; c < 13; c++) X[c] = ld(a.x0 + (size_t)live_to_ref(c) * N + n);
#pragma unroll
    for (int j = 0; j < 4; j++) JP[j] = ld(a.x0 + (size_t)(7 + j) * N + n);
    if (a.x_list && writer) {
        // row 0 = initial 23-vector: state as given + the controls of row 0 (Trainer.cpp:561-562, 610-611)
        for (int c = 0; c < 21; c++) a.x_list[(size_t)c * N + n] = ld(a.x0 + (size_t)c * N + n);
        a.x_list[(size_t)21 * N + n] = a.T > 1 ? ld(a.ctrl + n) : 0.0;
        a.x_list[(size_t)22 * N + n] = a.T > 1 ? ld(a.ctrl + N + n) : 0.0;
    }
    double ul = 0.0, ur = 0.0, inv_carry = 1.0;
    for (int i = 1; i < a.T; i++) {
        ul = ld(a.ctrl + ((size_t)(i - 1) * 2 + 0) * N + n);
        ur = ld(a.ctrl + ((size_t)(i - 1) * 2 + 1) * N + n);
        auto step = mk_step(ul, ur);
#pragma unroll 1
        for (int s = 0; s < a.substeps; s++) step((size_t)(i - 1) * a.substeps + s, inv_carry);
        // joint positions: X[7..10] += (ts/6)(k1 + 2k2 + 2k3 + k4) per RK4 step, k = (vl,vr,vl,vr) (HybridDynamics.cpp:495, 615-618)
        const double il = (kDt / 6.0) * (((ul + 2 * ul) + 2 * ul) + ul);
        const double ir = (kDt / 6.0) * (((ur + 2 * ur) + 2 * ur) + ur);
        for (int s = 0; s < a.substeps; s++) { JP[0] += il; JP[1] += ir; JP[2] += il; JP[3] += ir; }
        if (a.x_list && writer) {
            double* row = a.x_list + (size_t)i * 23 * N + n;
#pragma unroll
            for (int c = 0; c < 13; c++) row[(size_t)live_to_ref(c) * N] = X[c];
#pragma unroll
            for (int j = 0; j < 4; j++) row[(size_t)(7 + j) * N] = JP[j];
            row[(size_t)17 * N] = ul; row[(size_t)18 * N] = ur; row[(size_t)19 * N] = ul; row[(size_t)20 * N] = ur;
            row[(size_t)21 * N] = 0.0; row[(size_t)22 * N] = 0.0;  // VehicleSystem.cpp:144-147
        }
    }
    if (STORE && writer) {
        double* last = a.ckpt + ((size_t)(a.T - 1) * a.substeps * 4) * ck_rec_stride(NK) + ck_lane_offset(n);
        AVS_DBG_RANGE("final state record (store)", last, 6 * kCkPairStride + 2, a.ckpt, a.ckpt_doubles);
        CkptSink::ckpt_store(last, X, inv_carry);
    }
    if (a.x_final && writer) {
#pragma unroll
        for (int c = 0; c < 13; c++) a.x_final[(size_t)live_to_ref(c) * N + n] = X[c];
#pragma unroll
        for (int j = 0; j < 4; j++) a.x_final[(size_t)(7 + j) * N + n] = JP[j];
        // joint velocities hold the last applied control (or the caller's initial values when T == 1)
        a.x_final[(size_t)17 * N + n] = a.T > 1 ? ul : ld(a.x0 + (size_t)17 * N + n);
        a.x_final[(size_t)18 * N + n] = a.T > 1 ? ur : ld(a.x0 + (size_t)18 * N + n);
        a.x_final[(size_t)19 * N + n] = a.T > 1 ? ul : ld(a.x0 + (size_t)19 * N + n);
        a.x_final[(size_t)20 * N + n] = a.T > 1 ? ur : ld(a.x0 + (size_t)20 * N + n);
    }
}

constexpr int kBwdScratch = 87;  // host harness: LAM[13] | YB[13] | KB[13] | record buffer[14] | ...

// Reverse sweep over the checkpoints written by rollout_fwd_body<.., STORE=true>.
//   L = (sum_{i=1}^{T-1} l_i) / T / traj_len   (Trainer.cpp:612-649)
// mk_odeb(ul, ur) returns the functor SB = (df/dX)^T KB at XS (accumulating dL/dW); `st` is the reverse record
// stream positioned (and loaded) at the final-state record.
template <class MkOdeB, class Stream, class St>
AVS_HD void rollout_bwd_body(const RolloutArgs& a, int n, double& loss_out, St S, Stream& st, MkOdeB mk_odeb) {
    const size_t N = (size_t)a.N;
    const int T = a.T;
    St LAM = S, YB = S.sub(13), KB = S.sub(26);
    // traj_len = sum |d(x,y)_gt| ; 0 -> 1   (Trainer.cpp:635-647)
    double traj_len = 0.0;
    {
        double px = ld(a.gt + ((size_t)0 * 3 + 1) * N + n), py = ld(a.gt + ((size_t)0 * 3 + 2) * N + n);
        for (int i = 1; i < T; i++) {
            const double x = ld(a.gt + ((size_t)i * 3 + 1) * N + n), y = ld(a.gt + ((size_t)i * 3 + 2) * N + n);
            const double dx = x - px, dy = y - py;
            traj_len += sqrt(dx * dx + dy * dy);
            px = x; py = y;
        }
    }
    if (traj_len == 0.0) traj_len = 1.0;
    const double seed = 1.0 / (double)T / traj_len;
#pragma unroll
    for (int c = 0; c < 13; c++) LAM[c] = 0.0;
    double lsum = 0.0;
    for (int i = T - 1; i >= 1; i--) {
        {
            // st.cur() = state of data row i (the state after the step about to be reversed)
            const auto XR = st.cur();
            double Xr[13], lam[13];
#pragma unroll
            for (int c = 0; c < 13; c++) { Xr[c] = XR[c]; lam[c] = 0.0; }
            const double yaw_gt = ld(a.gt + ((size_t)i * 3 + 0) * N + n);
            const double x_gt = ld(a.gt + ((size_t)i * 3 + 1) * N + n);
            const double y_gt = ld(a.gt + ((size_t)i * 3 + 2) * N + n);
            lsum += row_loss(Xr, yaw_gt, x_gt, y_gt, seed, lam);
#pragma unroll
            for (int c = 0; c < 6; c++) LAM[c] += lam[c];  // the row loss touches q and (x, y) only
        }
        const double ul = ld(a.ctrl + ((size_t)(i - 1) * 2 + 0) * N + n);
        const double ur = ld(a.ctrl + ((size_t)(i - 1) * 2 + 1) * N + n);
        auto odeb = mk_odeb(ul, ur);
#pragma unroll 1
        for (int s = a.substeps - 1; s >= 0; s--) rk4_bwd_nn(odeb, st, LAM, YB, KB);
    }
    loss_out = lsum * seed;
}

}  // namespace avs
