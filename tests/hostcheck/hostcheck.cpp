// CPU unit-test harness for the device math in auvsl_neural_ode_b200/csrc/avs_model.cuh and
// avs_rollout.cuh: the same __host__ __device__ templates the sm_100a kernels are built from,
// compiled by g++ with one thread owning all four wheels (WPT = 4).  Used only by tests/ (CPU
// suite) to check the constant-folded ODE, the RK4 step and the hand-written adjoint against
// oracle/ without a GPU.  Not part of the shipped library.
#include <cstring>
#include <vector>
#include "../../auvsl_neural_ode_b200/csrc/avs_rollout.cuh"
#include "../../auvsl_neural_ode_b200/csrc/avs_weights.h"

using namespace avs;

struct ArraySink {
    double St[4][13], inv[4];
    template <class V> void stage(int s, const V& T, double i) { for (int c = 0; c < 13; c++) St[s][c] = T[c]; inv[s] = i; }
};

struct HostAcc {
    double g[kNumLive];
    HostAcc() { for (int i = 0; i < kNumLive; i++) g[i] = 0; }
    void record(const double* op) { grad_outer(g, op); }
};

static void fill_mc(ModelConst& mc, const double* p416, const double* zw152, const double* bek5, const double* height, const double* soil,
                    int nx, int ny, double x0, double y0, double dx, double dy) {
    memset(&mc, 0, sizeof(mc));
    if (p416) {
        memcpy(mc.W0, p416 + kOffW0, sizeof(mc.W0));
        memcpy(mc.W2, p416 + kOffW2, sizeof(mc.W2));
        memcpy(mc.W4, p416 + kOffW4, sizeof(mc.W4));
    }
    if (zw152) {
        memcpy(mc.Z0, zw152, sizeof(mc.Z0));
        memcpy(mc.Z2, zw152 + 8, sizeof(mc.Z2));
        memcpy(mc.Z4, zw152 + 72, sizeof(mc.Z4));
    }
    if (bek5) memcpy(mc.bek, bek5, sizeof(mc.bek));
    compute_AIinv(mc.AIinv);
    fill_contact_search(mc);
    mc.terr_kind = 0;
    static std::vector<double> ztab;
    if (zw152) { ztab.resize((size_t)kZTabRows * kZTabRow); build_ztab(mc.Z0, mc.Z2, mc.Z4, ztab.data()); mc.ztab = ztab.data(); }
    mc.grid.height = height; mc.grid.soil = soil; mc.grid.nx = nx; mc.grid.ny = ny;
    mc.grid.x0 = x0; mc.grid.y0 = y0; mc.grid.inv_dx = 1.0 / dx; mc.grid.inv_dy = 1.0 / dy;
}

template <int TIRE, int TERR> static void ode13(const ModelConst& mc, const double* X13, const double* u, double* Xd13, const double* ws) {
    ode_fwd<4, TIRE, TERR>(mc, 0, X13, u[0], u[1], Xd13, ws);
}
template <int TIRE, int TERR> static void rk413(const ModelConst& mc, const double* X13, const double* u, double* Xn13, double* stages) {
    ArraySink sink;
    double sc[52];
    Scratch<1> X{sc}, ACC{sc + 13}, TT{sc + 26}, KK{sc + 39};
    for (int c = 0; c < 13; c++) X[c] = X13[c];
    OdeFwdInline<4, TIRE, TERR, false> ode{mc, 0, u[0], u[1], nullptr, 0};
    double inv_carry = 1.0;
    rk4_fwd(ode, X, ACC, TT, KK, sink, inv_carry);
    for (int c = 0; c < 13; c++) Xn13[c] = X[c];
    if (stages) memcpy(stages, sink.St, sizeof(sink.St));
}

#define DISPATCH(FN, tire, terr, ...)                                                          \
    do {                                                                                       \
        if (tire == 0) {                                                                       \
            if (terr == 0) FN<0, 0>(__VA_ARGS__); else if (terr == 1) FN<0, 1>(__VA_ARGS__);   \
            else if (terr == 2) FN<0, 2>(__VA_ARGS__); else FN<0, 3>(__VA_ARGS__);             \
        } else {                                                                               \
            if (terr == 0) FN<1, 0>(__VA_ARGS__); else if (terr == 1) FN<1, 1>(__VA_ARGS__);   \
            else if (terr == 2) FN<1, 2>(__VA_ARGS__); else FN<1, 3>(__VA_ARGS__);             \
        }                                                                                      \
    } while (0)

template <int TIRE, int TERR> static void fwd_body(const ModelConst& mc, const RolloutArgs& a, int n, bool store) {
    double sc[kFwdScratch];
    const size_t stage_stride = (size_t)4 * a.Nc * kActLen;
    if (store) {
        typedef OdeFwdInline<4, TIRE, TERR, TIRE == 0> Ode;
        auto mk = [&](double ul, double ur) { return GenericStep<true, Ode, Scratch<1>>{Ode{mc, 0, ul, ur, nullptr, stage_stride}, a, n, 0, true, Scratch<1>{sc}}; };
        rollout_fwd_body<true>(a, n, 0, true, Scratch<1>{sc}, mk);
    } else {
        typedef OdeFwdInline<4, TIRE, TERR, false> Ode;
        auto mk = [&](double ul, double ur) { return GenericStep<false, Ode, Scratch<1>>{Ode{mc, 0, ul, ur, nullptr, 0}, a, n, 0, true, Scratch<1>{sc}}; };
        rollout_fwd_body<false>(a, n, 0, true, Scratch<1>{sc}, mk);
    }
}
// Bekker rollout with the branch-margin probe (avs_bekker.cuh): same node evaluation as BekNodesInline + a minima sink
struct BekNodesMarginHost {
    double* m;
    void margin(int kind, double v) const { if (v < m[kind]) m[kind] = v; }
    void operator()(const BekNodeParams& q, const double* th, const double* s, const double* c, double* fx, double* fy, double* fz) const {
        BekNodesInline()(q, th, s, c, fx, fy, fz);
    }
};
template <int TERR> struct OdeFwdMarginHost {
    const ModelConst& mc; double ul, ur; double* m;
    double* act_step = nullptr;  // (member GenericStep expects; the STORE path is never instantiated for it)
    template <class St> void operator()(St T, St K, int) const {
        double X[13], Xd[13];
        for (int c = 0; c < 13; c++) X[c] = T[c];
        ode_fwd<4, 1, TERR, false, BekNodesMarginHost>(mc, 0, X, ul, ur, Xd, nullptr, ActOutPtr{nullptr}, BekNodesMarginHost{m});
        for (int c = 0; c < 13; c++) K[c] = Xd[c];
    }
};
template <int TIRE, int TERR> static void fwd_body_margin(const ModelConst& mc, const RolloutArgs& a, int n, double* m) {
    double sc[kFwdScratch];
    typedef OdeFwdMarginHost<TERR> Ode;
    auto mk = [&](double ul, double ur) { return GenericStep<false, Ode, Scratch<1>>{Ode{mc, ul, ur, m}, a, n, 0, true, Scratch<1>{sc}}; };
    rollout_fwd_body<false>(a, n, 0, true, Scratch<1>{sc}, mk);
}

template <int TERR> static void bwd_body(const ModelConst& mc, const RolloutArgs& a, int n, HostAcc& acc, double& loss) {
    double sc[kBwdScratch];
    auto mk = [&mc, &acc](double ul, double ur) { return OdeBwdInline<4, TERR, HostAcc>{mc, 0, ul, ur, acc}; };
    const size_t nrec = (size_t)(a.T - 1) * a.substeps * 4;
    CkptStreamSimple<Scratch<1>> st{a.ckpt + nrec * ck_rec_stride(a.Nc) + ck_lane_offset(n), ck_rec_stride(a.Nc), Scratch<1>{sc + 52},
                                    a.act + (nrec * 4 * a.Nc + (size_t)4 * n) * kActLen, (size_t)4 * a.Nc * kActLen};
    st.load();
    rollout_bwd_body(a, n, loss, Scratch<1>{sc}, st, mk);
}

static int g_contact_search = 0;  // hc_set_contact_search: the TERR_DYN instantiations read ModelConst::contact_search

extern "C" {

void hc_set_contact_search(int on) { g_contact_search = on; }

struct hc_model {
    int tire, terr, nx, ny;
    double x0, y0, dx, dy;
    const double* params;  // 416 (NN) or 5 (Bekker)
    const double* zw;      // 152 fixed z-net weights (Z0[8], Z2[64], Z4[8]) — NN only
    const double* height;
    const double* soil;
};

static void mk(const hc_model* m, ModelConst& mc);
static void mk_base(const hc_model* m, ModelConst& mc) {
    fill_mc(mc, m->tire == 0 ? m->params : nullptr, m->zw, m->tire == 1 ? m->params : nullptr, m->height, m->soil, m->nx, m->ny, m->x0, m->y0,
            m->dx > 0 ? m->dx : 1.0, m->dy > 0 ? m->dy : 1.0);
}

static void mk(const hc_model* m, ModelConst& mc) {
    mk_base(m, mc);
    mc.terr_kind = m->terr;                  // read by the TERR_DYN (= 7) instantiations only
    mc.contact_search = g_contact_search;
}

void hc_zweights(double* out152) { memcpy(out152, kFixedZ0, 64); memcpy(out152 + 8, kFixedZ2, 512); memcpy(out152 + 72, kFixedZ4, 64); }
void hc_default_params(double* p416) { for (int k = 0; k < 4; k++) { memcpy(p416 + 104 * k, kDefaultW0, 192); memcpy(p416 + 104 * k + 24, kDefaultW2, 512); memcpy(p416 + 104 * k + 88, kDefaultW4, 128); } }
void hc_aiinv(double* out36) { double A[6][6]; compute_AIinv(A); memcpy(out36, A, sizeof(A)); }
double hc_tanh(double x) { return tanh_f64(x); }
// the tanh of the kernels' hot path (tanh_vec: 19 FP64 instructions per element on the device; the CPU build divides exactly)
double hc_tanh_vec(double x) {
    double v[1] = {x};
    tanh_vec<1>(v);
    return v[0];
}
double hc_ztab(double s0, double* dz) {
    static std::vector<double> t;
    if (t.empty()) { t.resize((size_t)kZTabRows * kZTabRow); build_ztab((const double*)kFixedZ0, (const double(*)[8])kFixedZ2, (const double*)kFixedZ4, t.data()); }
    // both evaluators of the table: the branch-free value form and the joint value + derivative Horner pass
    const double v = ztab_eval(t.data(), s0);
    if (dz) {
        double d = 0.0;
        const double v2 = s0 > 0.0 ? ztab_eval_d(t.data(), s0, d) : v;
        *dz = d;
        if (fabs(v2 - v) > 4e-16 * (1.0 + fabs(v))) return NAN;  // the two forms must agree to rounding
    }
    return v;
}

void hc_bekker_tire(const hc_model* m, double cvx, double cvy, double w, double sinkage, double* F3) {
    ModelConst mc; mk(m, mc);
    double cp[3] = {0, 0, 0};
    bekker_tire_fwd<0>(mc, cvx, cvy, w, sinkage, cp, F3);
}
void hc_ode(const hc_model* m, const double* X13, const double* u2, double* Xd13, const double* ws4) {
    ModelConst mc; mk(m, mc);
    if (g_contact_search) { if (m->tire == 0) ode13<0, TERR_DYN>(mc, X13, u2, Xd13, ws4); else ode13<1, TERR_DYN>(mc, X13, u2, Xd13, ws4); return; }
    DISPATCH(ode13, m->tire, m->terr, mc, X13, u2, Xd13, ws4);
}
void hc_rk4(const hc_model* m, const double* X13, const double* u2, double* Xn13, double* stages52) {
    ModelConst mc; mk(m, mc);
    DISPATCH(rk413, m->tire, m->terr, mc, X13, u2, Xn13, stages52);
}
// vector-Jacobian product of the ODE: Xb = (df/dX)^T Xdb, gW = (df/dW)^T Xdb
void hc_ode_vjp(const hc_model* m, const double* X13, const double* u2, const double* Xdb13, double* Xb13, double* gW104) {
    ModelConst mc; mk(m, mc);
    HostAcc acc;
    double kb[13], Xd[13], actrec[kActLen * 4];
    memcpy(kb, Xdb13, sizeof(kb));
    Scratch<1> KB{kb};
    switch (m->terr) {  // forward with activation store (four consecutive 20-double records), then the adjoint
        case 0: ode_fwd<4, 0, 0, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn<4, 0>(mc, 0, X13, u2[0], u2[1], KB, Xb13, acc, ActInPtr{actrec}); break;
        case 1: ode_fwd<4, 0, 1, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn<4, 1>(mc, 0, X13, u2[0], u2[1], KB, Xb13, acc, ActInPtr{actrec}); break;
        case 2: ode_fwd<4, 0, 2, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn<4, 2>(mc, 0, X13, u2[0], u2[1], KB, Xb13, acc, ActInPtr{actrec}); break;
        default: ode_fwd<4, 0, 3, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn<4, 3>(mc, 0, X13, u2[0], u2[1], KB, Xb13, acc, ActInPtr{actrec}); break;
    }
    memcpy(gW104, acc.g, sizeof(acc.g));
}
// the same product through ode_bwd_nn_p (layered back-propagation on the state chain, outer-product operands handed over):
// what the reverse kernel's producer warp runs
void hc_ode_vjp_layered(const hc_model* m, const double* X13, const double* u2, const double* Xdb13, double* Xb13, double* gW104) {
    ModelConst mc; mk(m, mc);
    HostAcc acc;
    double kb[13], Xd[13], actrec[kActLen * 4];
    memcpy(kb, Xdb13, sizeof(kb));
    Scratch<1> KB{kb};
    LayeredFromRecords<HostAcc> io{actrec, acc};
    switch (m->terr) {
        case 0: ode_fwd<4, 0, 0, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn_p<4, 0>(mc, 0, X13, u2[0], u2[1], KB, Xb13, io, io); break;
        case 1: ode_fwd<4, 0, 1, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn_p<4, 1>(mc, 0, X13, u2[0], u2[1], KB, Xb13, io, io); break;
        case 2: ode_fwd<4, 0, 2, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn_p<4, 2>(mc, 0, X13, u2[0], u2[1], KB, Xb13, io, io); break;
        default: ode_fwd<4, 0, 3, true>(mc, 0, X13, u2[0], u2[1], Xd, nullptr, ActOutPtr{actrec}); ode_bwd_nn_p<4, 3>(mc, 0, X13, u2[0], u2[1], KB, Xb13, io, io); break;
    }
    memcpy(gW104, acc.g, sizeof(acc.g));
}
// SoA batch rollout, same buffers as the kernels
void hc_rollout(const hc_model* m, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final) {
    ModelConst mc; mk(m, mc);
    RolloutArgs a; memset(&a, 0, sizeof(a));
    a.N = N; a.Nc = N; a.T = T; a.substeps = substeps; a.x0 = x0; a.ctrl = ctrl; a.x_list = x_list; a.x_final = x_final;
    if (g_contact_search) {  // the contact search lives in the run-time-terrain instantiation
        for (int n = 0; n < N; n++) { if (m->tire == 0) fwd_body<0, TERR_DYN>(mc, a, n, false); else fwd_body<1, TERR_DYN>(mc, a, n, false); }
        return;
    }
    for (int n = 0; n < N; n++) DISPATCH(fwd_body, m->tire, m->terr, mc, a, n, false);
}
// Bekker rollout + branch margins [kNumMargins][N] (same definitions as the kernel's diagnostic instantiation)
void hc_rollout_margins(const hc_model* m, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final, double* margins) {
    ModelConst mc; mk(m, mc);
    RolloutArgs a; memset(&a, 0, sizeof(a));
    a.N = N; a.Nc = N; a.T = T; a.substeps = substeps; a.x0 = x0; a.ctrl = ctrl; a.x_list = x_list; a.x_final = x_final;
    for (int n = 0; n < N; n++) {
        double mm[kNumMargins];
        for (int k = 0; k < kNumMargins; k++) mm[k] = 1e300;
        DISPATCH(fwd_body_margin, 1, m->terr, mc, a, n, mm);
        for (int k = 0; k < kNumMargins; k++) margins[(size_t)k * N + n] = mm[k];
    }
}
void hc_rollout_grad(const hc_model* m, int N, int T, int substeps, const double* x0, const double* ctrl, const double* gt, double* loss, double* gps104xN) {
    ModelConst mc; mk(m, mc);
    std::vector<double> ckpt(((size_t)(T - 1) * substeps * 4 + 1) * ck_rec_stride(N));
    RolloutArgs a; memset(&a, 0, sizeof(a));
    std::vector<double> actbuf((size_t)(T - 1) * substeps * 4 * kActLen * 4 * N);
    a.N = N; a.Nc = N; a.T = T; a.substeps = substeps; a.x0 = x0; a.ctrl = ctrl; a.ckpt = ckpt.data(); a.act = actbuf.data(); a.gt = gt;
    for (int n = 0; n < N; n++) {
        DISPATCH(fwd_body, 0, m->terr, mc, a, n, true);
        HostAcc acc;
        double L;
        switch (m->terr) {
            case 0: bwd_body<0>(mc, a, n, acc, L); break;
            case 1: bwd_body<1>(mc, a, n, acc, L); break;
            case 2: bwd_body<2>(mc, a, n, acc, L); break;
            default: bwd_body<3>(mc, a, n, acc, L); break;
        }
        loss[n] = L;
        for (int k = 0; k < kNumLive; k++) gps104xN[(size_t)k * N + n] = acc.g[k];
    }
}

}  // extern "C"
