// avs_model.cuh — constant-folded Jackal vehicle ODE, RK4 step and their hand-written adjoints.
//
// This is the arithmetic core of the sm_100a kernels (avs_kernels.cu).  It is written as
// __host__ __device__ templates so that the *same* source is also compiled by g++ into the CPU
// unit-test harness tests/hostcheck (which checks it against oracle/ without a GPU).  It is NOT a
// CPU fallback: the shipped library (libavsim.so) only reaches it through CUDA kernels.
//
// What the reference does per ODE call (all on CppAD::AD<double>, dense 6x6 Eigen products) and how
// it is folded here — reference paths relative to /root/reference:
//   * joint angles are frozen at q = 0 (src/auvsl_dynamics/HybridDynamics.cpp:520) so the four wheel
//     motion transforms X_i (generated/transforms.cpp:194-239, 286-331, 378-423, 470-515) are the
//     constant matrices [[E,0],[-E t_i^, E]],  E = [[1,0,0],[0,0,-1],[0,1,0]]
//   * the articulated inertia of a wheel, Ia = I_w - U U^T / D (generated/forward_dynamics.cpp:116),
//     and the assembled base inertia AI_b = I_b + sum_i X_i^T Ia X_i (:118-153) are constants; the
//     LLT solve of :156 becomes a product with the precomputed inverse (ModelConst::AIinv)
//   * the third ABA pass (:158-173) only produces qdd, which the caller discards
//     (HybridDynamics.cpp:627-630) — dropped
//   * TireNetwork::forward (src/auvsl_dynamics/TireNetwork.cpp:54-119) always uses network 0
//     (HybridDynamics.cpp:387)
// Differentiable state ("St", 13 doubles): quaternion x,y,z,w | position | body angular vel | body
// linear vel  = reference state entries 0-3, 4-6, 11-13, 14-16 (HybridDynamics.h:65-69).  Joint
// positions (7-10) integrate the controls and feed nothing; joint velocities (17-20) are the controls.
//
// Work split ("lanes"): a vehicle is processed by 4/WPT threads, each owning WPT wheels.  Per-wheel
// work (contact geometry, tire model, ABA passes 1-2 for that wheel) is lane-local; the 6-vector
// bias-force contributions are summed across the lanes of the group (Group<WPT>::sum, warp shuffles
// on the device, identity when WPT == 4) and the small base-link part is evaluated redundantly.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AVS_HD __host__ __device__ __forceinline__
#else
#define AVS_HD inline
#endif

namespace avs {

// ------------------------------------------------------------------------------------------------
// Model constants (generated/model_constants.h:24-100); see DESIGN.md §3 for the derivations
// ------------------------------------------------------------------------------------------------
constexpr double kTireRadius = .098;
constexpr double kTx = 0.13099999725818634;   // |tx| of the wheel joints; FL,FR = +, RL,RR = -
constexpr double kTy = 0.1877949982881546;    // |ty|; FL,RL = +, FR,RR = -
constexpr double kTz = 0.03449999913573265;
constexpr double kRz = kTz + (-kTireRadius);  // contact point z offset (HybridDynamics.cpp:222-232)
constexpr double kMb = 16.52400016784668;
constexpr double kCx = 0.011999273672699928, kCz = 0.06699594855308533;  // base COM (comy forced to 0, :50,:68)
constexpr double kIbxx = 0.38783785700798035, kIbyy = 0.46875107288360596, kIbzz = 0.4509454071521759;
constexpr double kIbxy = 0.0011965520679950714, kIbxz = -0.003115505911409855, kIbyz = 0.0031140821520239115;
constexpr double kMw = 0.47699999809265137;
constexpr double kIwxx = 0.0013000000035390258, kIwyy = 0.0013000000035390258, kIwzz = 0.002400000113993883;
constexpr double kIwyz = 4.808253448174149E-11;
constexpr double kIwyyA = kIwyy - kIwyz * kIwyz / kIwzz;  // Ia(AY,AY) after compute_Ia_revolute
constexpr double kKappa = kIwyz / kIwzz;                  // -U(AY)/D
constexpr double kGravity = -9.81;                        // HybridDynamics.cpp:34
constexpr double kDt = 1e-3;                              // HybridDynamics.cpp:33
// TireNetwork scalers (TireNetwork.cpp:11-13)
constexpr double kOutStd0 = 17.241097080572587, kOutStd1 = 15.9103406661298, kOutStd2 = 12.632573461434308;
constexpr double kInStd0 = 0.0008373582968488336, kInStd1 = 0.5799984931945801, kInStd2 = 0.576934278011322,
                 kInStd3 = 0.5774632096290588;
constexpr double kNNDamp = -20.0;       // HybridDynamics.cpp:403
constexpr double kBekkerDamp = -200.0;  // BekkerDynamics.cpp:106

enum TireKind { TIRE_NN = 0, TIRE_BEKKER = 1 };
enum TerrainKind { TERR_FLAT = 0, TERR_SLOPE = 1, TERR_BUMPY = 2, TERR_GRID = 3, TERR_DYN = 7 /* switch on ModelConst::terr_kind at run time */ };

constexpr int kNumNNParams = 416;  // TireNetwork::getNumParams (TireNetwork.cpp:122-128): 4 nets x 104
constexpr int kNumLive = 104;      // only net 0 is evaluated -> entries 104..415 of the gradient are 0
constexpr int kOffW0 = 0, kOffW2 = 24, kOffW4 = 88;  // pack order W0,W2,W4 row-major (TireNetwork.cpp:130-161)

struct GridTerrain {
    const double* height;  // [ny][nx]
    const double* soil;    // [5][ny][nx] or nullptr
    int nx, ny;
    double x0, y0, inv_dx, inv_dy;
};

// Everything a kernel needs besides per-vehicle data; passed by value (__grid_constant__), so every
// entry is a constant-bank operand with a compile-time offset.
struct ModelConst {
    double W0[8][3], W2[8][8], W4[2][8];  // trainable xy-net (TireNetwork.h:40-42)
    double Z0[8], Z2[8][8], Z4[8];        // fixed z-net (TireNetwork.cpp:17-32)
    double AIinv[6][6];                   // (I_b + sum X^T Ia X)^-1
    double bek[5];                        // Bekker soil 5-vector as passed to setParams (BekkerDynamics.cpp:12-21)
    GridTerrain grid;
    const double* ztab;                   // [kZTabRows][kZTabRow] piecewise polynomials of the fixed z-net and its derivative (build_ztab)
    int terr_kind;                        // used by the TERR_DYN instantiations
    // optional 10-angle contact search (HybridDynamics::get_tire_cpts_sinkages_fancy, HybridDynamics.cpp:258-316), TERR_DYN
    // instantiations only: sin / cos of the candidate pitch angles -0.3 ... +0.3 rad (filled by fill_contact_search)
    int contact_search;
    double cs_sin[10], cs_cos[10];
};
inline void fill_contact_search(ModelConst& mc) {
    const double max_angle = -.3;
    const int max_checks = 10;
    for (int jj = 0; jj < max_checks; jj++) {
        const double a = max_angle - (2 * max_angle * jj / ((double)max_checks - 1));  // :293
        mc.cs_sin[jj] = sin(a);
        mc.cs_cos[jj] = cos(a);
    }
}

// ------------------------------------------------------------------------------------------------
// fp64 tanh.  tanh|x| = (1-e)/(1+e), e = exp(-2|x|): one exp (Cody-Waite reduction + degree-11
// polynomial, |rel err| < 2e-17 before rounding) and one division on d in [1,2] (hardware
// reciprocal seed + cubic Newton step; the scalar tanh_f64 adds one residual correction, the 8-wide tanh_vec of the
// hot path does not: quotient within ~1.5 ulp instead of 0.5, 22 instead of 24 FP64 instructions).  Absolute error <= ~3e-16.
// ------------------------------------------------------------------------------------------------
AVS_HD double exp_negarg(double x /* <= 0 */) {
    const double L2E = 1.4426950408889634e+0, MAGIC = 6755399441055744.0;
    const double LN2_HI = 6.93147180559945286e-01, LN2_LO = 2.31904681384629956e-17;
    double t = fma(x, L2E, MAGIC);
    double kd = t - MAGIC;
    double r = fma(kd, -LN2_HI, x);
    r = fma(kd, -LN2_LO, r);
    double p = 2.5110049204818658e-08;
    p = fma(p, r, 2.763265472252779e-07);
    p = fma(p, r, 2.755724088722987e-06);
    p = fma(p, r, 2.4801485441561313e-05);
    p = fma(p, r, 0.00019841269890076403);
    p = fma(p, r, 0.0013888888952352863);
    p = fma(p, r, 0.008333333333319589);
    p = fma(p, r, 0.04166666666648795);
    p = fma(p, r, 0.1666666666666668);
    p = fma(p, r, 0.5000000000000019);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#if defined(__CUDA_ARCH__)
    int k = __double2loint(t);
    k = max(k, -1000);  // exp(-693) ~ 1e-301: result irrelevant next to 1, keeps the exponent field valid
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    int64_t k = (int64_t)kd;
    if (k < -1000) k = -1000;
    return ldexp(p, (int)k);
#endif
}

AVS_HD double div_1to2(double n, double d /* in [1,2] */) {
#if defined(__CUDA_ARCH__)
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    double e0 = fma(-d, r0, 1.0);
    double e1 = fma(e0, e0, e0);
    double r1 = fma(r0, e1, r0);
    double q = n * r1;
    double rem = fma(-d, q, n);
    return fma(rem, r1, q);
#else
    return n / d;
#endif
}

AVS_HD double tanh_f64(double x) {
    double ax = fabs(x);
    double e = exp_negarg(-2.0 * ax);
    double t = div_1to2(1.0 - e, 1.0 + e);
    return copysign(t, x);
}

// W-wide tanh, in place: every step is issued for all W elements before the next step starts (with one warp per SM
// sub-partition the only latency hiding available is instruction-level parallelism, and ptxas does not interleave separately
// inlined dependent chains).  19 FP64-pipe instructions + 1 MUFU per element — the kernels are bound by issue slots, an FP64
// instruction takes two, and 16 tanh per wheel and stage are 40 % of the forward sweep's FP64 work — against the 24 of tanh_f64:
//   * -2|x| is never formed: with r' = -r/2 = |x| + k ln2/2 the polynomial of exp(r) = exp(-2 r') has the coefficients
//     c_n (-2)^n (exact scalings);
//   * the low word of ln 2 is left out of the reduction: it changes e = exp(-2|x|) by |k| 2.3e-17 relative, i.e. tanh by at most
//     2 e |k| 2.3e-17 <= 2.3e-17 absolute (e <= 2^k, k <= 0) — a fifth of an ulp of the result;
//   * tanh|x| = 1 - 2e / (1 + e) with 2e out of the exponent patch (k + 1), the reciprocal of d = 1 + e in [1, 2] by
//     rcp.approx.ftz.f64 + one cubic Newton step (good to ~2^-60), no residual correction of the quotient.
// Absolute error <= 3e-16 (tests/test_device_math_cpu.py).  Measured (round 2, N = 4096): forward only 27.5 -> 26.6 ms, forward +
// store 10.15 -> 9.9 ms; the folding alone had given 0.5 % earlier in the round and was not kept then.
template <int W> AVS_HD void tanh_vec(double* x) {
    const double L2E = 1.4426950408889634e+0, MAGIC = 6755399441055744.0;
    const double LN2_HI = 6.93147180559945286e-01;
    double a[W], t[W], r[W], p[W];
#pragma unroll
    for (int i = 0; i < W; i++) a[i] = fabs(x[i]);
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = fma(a[i], -2.0 * L2E, MAGIC);
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = t[i] - MAGIC;
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = fma(r[i], 0.5 * LN2_HI, a[i]);
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(-2048.0 * 2.5110049204818658e-08, r[i], 1024.0 * 2.763265472252779e-07);
#define AVS_POLY_STEP(C)           \
    _Pragma("unroll") for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], C);
    AVS_POLY_STEP(-512.0 * 2.755724088722987e-06)
    AVS_POLY_STEP(256.0 * 2.4801485441561313e-05)
    AVS_POLY_STEP(-128.0 * 0.00019841269890076403)
    AVS_POLY_STEP(64.0 * 0.0013888888952352863)
    AVS_POLY_STEP(-32.0 * 0.008333333333319589)
    AVS_POLY_STEP(16.0 * 0.04166666666648795)
    AVS_POLY_STEP(-8.0 * 0.1666666666666668)
    AVS_POLY_STEP(4.0 * 0.5000000000000019)
    AVS_POLY_STEP(-2.0)
    AVS_POLY_STEP(1.0)
#undef AVS_POLY_STEP
    // 2e = p * 2^(k + 1)
#pragma unroll
    for (int i = 0; i < W; i++) {
#if defined(__CUDA_ARCH__)
        int k = __double2loint(t[i]);
        k = max(k, -1000) + 1;  // exp(-693) ~ 1e-301: result irrelevant next to 1, keeps the exponent field valid
        p[i] = __hiloint2double(__double2hiint(p[i]) + (k << 20), __double2loint(p[i]));
#else
        int64_t k = (int64_t)(t[i] - MAGIC);
        if (k < -1000) k = -1000;
        p[i] = ldexp(p[i], (int)k + 1);
#endif
    }
    // 1 - 2e / (1 + e)
#pragma unroll
    for (int i = 0; i < W; i++) a[i] = fma(p[i], 0.5, 1.0);
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < W; i++) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(a[i]));
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = fma(-a[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = fma(t[i], t[i], t[i]);
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = fma(r[i], t[i], r[i]);
#else
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = 1.0 / a[i];
#endif
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = fma(-p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = copysign(t[i], x[i]);
}

AVS_HD double rsqrt_f64(double x) {  // for quaternion re-normalisation, x ~ 1
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// ------------------------------------------------------------------------------------------------
// Lane group
// ------------------------------------------------------------------------------------------------
template <int WPT> struct Group {
    static_assert(WPT == 1 || WPT == 2 || WPT == 4, "wheels per thread");
    static constexpr int LANES = 4 / WPT;
    // sum over the lanes of the group; every lane receives the total
    AVS_HD static double sum(double x) {
#if defined(__CUDA_ARCH__)
        if (WPT <= 2) x += __shfl_xor_sync(0xffffffffu, x, 1);
        if (WPT == 1) x += __shfl_xor_sync(0xffffffffu, x, 2);
#endif
        return x;
    }
};

// Per-thread scratch column: element c lives at p[c * STRIDE].  On the device p points into shared memory
// at &buf[threadIdx.x] with STRIDE = blockDim.x (consecutive lanes -> consecutive banks, conflict-free);
// on the host STRIDE = 1 over a local array.  Loop-carried rollout state lives here so that the registers
// are free for the instruction-level parallelism of the ODE evaluation.
template <int STRIDE> struct Scratch {
    double* p;
    AVS_HD double& operator[](int c) const { return p[c * STRIDE]; }
    AVS_HD Scratch sub(int c0) const { return Scratch{p + c0 * STRIDE}; }
};

// ------------------------------------------------------------------------------------------------
// Terrain lookup: altitude and its gradient at (x, y)
//   (src/types/TerrainMap.h:10, src/TestTerrainMaps.cpp:6-28; grid = extension)
// ------------------------------------------------------------------------------------------------
struct GridCellD { int i, j; double fu, fv, du, dv; };  // du/dv: d(fu)/dx, d(fv)/dy (0 where clamped)
AVS_HD GridCellD grid_cell(const GridTerrain& g, double x, double y) {
    double u = (x - g.x0) * g.inv_dx, v = (y - g.y0) * g.inv_dy;
    const double umax = (double)(g.nx - 1), vmax = (double)(g.ny - 1);
    GridCellD c;
    c.du = g.inv_dx; c.dv = g.inv_dy;
    if (u < 0.0) { u = 0.0; c.du = 0.0; }
    if (u > umax) { u = umax; c.du = 0.0; }
    if (v < 0.0) { v = 0.0; c.dv = 0.0; }
    if (v > vmax) { v = vmax; c.dv = 0.0; }
    int i = (int)floor(u), j = (int)floor(v);
    i = i > g.nx - 2 ? g.nx - 2 : i; j = j > g.ny - 2 ? g.ny - 2 : j;
    i = i < 0 ? 0 : i; j = j < 0 ? 0 : j;
    c.i = i; c.j = j; c.fu = u - (double)i; c.fv = v - (double)j;
    return c;
}
AVS_HD double grid_lerp(const GridTerrain& g, const double* f, const GridCellD& c, double* dfdx, double* dfdy) {
    const double* row0 = f + (size_t)c.j * g.nx + c.i;
    const double* row1 = row0 + g.nx;
#if defined(__CUDA_ARCH__)
    const double h00 = __ldg(row0), h01 = __ldg(row0 + 1), h10 = __ldg(row1), h11 = __ldg(row1 + 1);
#else
    const double h00 = row0[0], h01 = row0[1], h10 = row1[0], h11 = row1[1];
#endif
    const double a = h00 + c.fu * (h01 - h00);
    const double b = h10 + c.fu * (h11 - h10);
    if (dfdx) {
        *dfdx = ((h01 - h00) + c.fv * ((h11 - h10) - (h01 - h00))) * c.du;
        *dfdy = (b - a) * c.dv;
    }
    return a + c.fv * (b - a);
}

template <int TERR> AVS_HD double altitude(const ModelConst& mc, double x, double y, double& dax, double& day) {
    const int kind = (TERR == TERR_DYN) ? mc.terr_kind : TERR;
    if (kind == TERR_FLAT) { dax = 0.0; day = 0.0; return 0.0; }
    if (kind == TERR_SLOPE) { dax = 0.1; day = 0.0; return 0.1 * x; }
    if (kind == TERR_BUMPY) {
        // CondExpLt(arg,1.6,arg,0); CondExpGt(arg,0,arg,0); temp=.2 sin(arg); CondExpGt(temp,0,temp,0)
        double arg = x - 2.0;
        bool live = true;
        if (!(arg < 1.6)) { arg = 0.0; live = false; }
        if (!(arg > 0.0)) { arg = 0.0; live = false; }
        double temp = .2 * sin(arg);
        day = 0.0;
        if (temp > 0.0) { dax = live ? .2 * cos(arg) : 0.0; return temp; }
        dax = 0.0;
        return 0.0;
    }
    GridCellD c = grid_cell(mc.grid, x, y);
    return grid_lerp(mc.grid, mc.grid.height, c, &dax, &day);
}

// ------------------------------------------------------------------------------------------------
// Tire network (TireNetwork.cpp:54-119), network 0.  Activations are kept for the adjoint.
// ------------------------------------------------------------------------------------------------
// ---- the fixed z-branch as a table ---------------------------------------------------------------
// z4(s0) = Z4 . tanh(Z2 tanh(Z0 s0)) has CONSTANT weights (TireNetwork.cpp:17-32; they are not in the
// trainable parameter vector, :130-161) and a scalar input (the scaled sinkage), i.e. it is one fixed analytic
// odd function of one variable.  The reference re-evaluates its 16 tanh per wheel per ODE call; here it is
// constant-folded like AI^-1: a piecewise degree-9 polynomial on 832 uniform intervals of width 1/16 over
// s0 in [0, 52] (beyond which every first-layer unit is saturated and z4 is constant to 1.5e-16), built at
// context creation in long double (build_ztab).  Interpolation error < 1e-16 relative (tests).  For s0 <= 0,
// z4 <= 0 and relu gives exactly 0 with zero derivative, as CondExpGt does.
constexpr int kZTabDeg = 9, kZTabPerUnit = 16, kZTabN = 832;
constexpr double kZTabSMax = (double)kZTabN / kZTabPerUnit;
// Row layout (kZTabRow = 10 doubles = five 16-byte pairs, kZTabRows = kZTabN + 1 rows; row kZTabN is the constant tail):
//   c0..c9 : value polynomial in y = 16 s0 - (i + 1/2), y in [-1/2, 1/2]  (interval i = [i/16, (i+1)/16], centre form)
// Evaluation is branch-free and short: the interval index comes out of the mantissa of (16 s0 - 1/2) + 2^52 + 2^51 (round to
// nearest: no F2I / I2F round trip), the coefficients arrive as five 16-byte loads, and the polynomial is evaluated by
// Estrin's scheme (depth 4 after the loads instead of 9 dependent Horner steps).  The first version sat in a divergent region
// (s0 > 0 ? ... : 0) that ptxas could not interleave with the xy-net: ~300 cycles of serial chain per ODE stage on the single
// resident warp (ncu source page, round 2).  Forward-only rollouts use this form (28.6 -> 28.0 ms at 4096 x 600 rows); the
// instantiation that also stores the adjoint's records keeps the joint value + derivative Horner pass (ztab_eval_d).
constexpr int kZTabRow = kZTabDeg + 1, kZTabRows = kZTabN + 1;

struct ZTabPos { double y; const double* row; };
AVS_HD ZTabPos ztab_pos(const double* tab, double s0) {
    const double MAGIC = 6755399441055744.0;
    double sc = s0 > 0.0 ? s0 : 0.0;
    sc = sc < kZTabSMax ? sc : kZTabSMax;
    const double v = fma(sc, (double)kZTabPerUnit, -0.5);   // in [-1/2, kZTabN - 1/2]
    const double t = v + MAGIC;                             // nearest integer i in [0, kZTabN] in the low mantissa bits
#if defined(__CUDA_ARCH__)
    const int i = __double2loint(t);
#else
    const int i = (int)(int64_t)(t - MAGIC);
#endif
    return ZTabPos{v - (t - MAGIC), tab + (size_t)i * kZTabRow};
}
struct ZPair { double x, y; };
AVS_HD ZPair ztab_ld(const double* row, int i) {
#if defined(__CUDA_ARCH__)
    const double2 v = __ldg(reinterpret_cast<const double2*>(row) + i);
    return ZPair{v.x, v.y};
#else
    return ZPair{row[2 * i], row[2 * i + 1]};
#endif
}
// z4(max(s0, 0)); callers apply the relu gate themselves
AVS_HD double ztab_eval(const double* tab, double s0) {
    const ZTabPos q = ztab_pos(tab, s0);
    const ZPair c01 = ztab_ld(q.row, 0), c23 = ztab_ld(q.row, 1), c45 = ztab_ld(q.row, 2), c67 = ztab_ld(q.row, 3), c89 = ztab_ld(q.row, 4);
    const double y = q.y, y2 = y * y, y4 = y2 * y2, y8 = y4 * y4;
    const double a0 = fma(c01.y, y, c01.x), a1 = fma(c23.y, y, c23.x), a2 = fma(c45.y, y, c45.x), a3 = fma(c67.y, y, c67.x), a4 = fma(c89.y, y, c89.x);
    const double b0 = fma(a1, y2, a0), b1 = fma(a3, y2, a2);
    return fma(a4, y8, fma(b1, y4, b0));
}
// value AND derivative (d z4 / d s0), by one joint Horner pass over the value coefficients — the form the STORE instantiation
// of the forward kernel uses (18 FP64 instructions; Estrin's scheme for both would take 24, and the kernel is bound by issue
// slots).  Branch-free since the end of round 2 (forward + store 10.48 -> 10.23 ms); under the 170-register cap of the earlier
// three-CTA build every branch-free variant had been slower (11.3-11.9 against 11.08 ms).
AVS_HD double ztab_eval_d(const double* tab, double s0, double& dz) {
    // evaluated at max(s0, 0), the caller gates value and derivative by selects: no divergent region, so ptxas interleaves the
    // Horner chain (18 dependent FMAs, ~150 cycles when it ran alone inside `s0 > 0 ? ... : 0`: ncu source page) with the xy-net
    const double sp = s0 > 0.0 ? s0 : 0.0;
    const double sc = sp < kZTabSMax ? sp : kZTabSMax;
    const double u = sc * (double)kZTabPerUnit;
    int i = (int)u;
    i = i > kZTabN - 1 ? kZTabN - 1 : i;
    const double y = (u - (double)i) - 0.5;
    const double* c = tab + (size_t)i * kZTabRow;
    double cc[kZTabDeg + 1];
#pragma unroll
    for (int k = 0; k <= kZTabDeg; k++) {  // ten 8-byte loads: five 16-byte ones were measured slower here (10.29 against 10.21 ms)
#if defined(__CUDA_ARCH__)
        cc[k] = __ldg(c + k);
#else
        cc[k] = c[k];
#endif
    }
    double p = cc[kZTabDeg], d = 0.0;
#pragma unroll
    for (int k = kZTabDeg - 1; k >= 0; k--) { d = fma(d, y, p); p = fma(p, y, cc[k]); }
    dz = s0 < kZTabSMax ? d * (double)kZTabPerUnit : 0.0;
    return p;
}

// host-side construction (long double): Chebyshev-node interpolation per interval, converted to monomials in y in [-1/2, 1/2]
inline void build_ztab(const double Z0[8], const double Z2[8][8], const double Z4[8], double* tab) {
    const int n = kZTabDeg + 1;
    const long double PI = 3.14159265358979323846264338327950288L;
    // monomial coefficients of the Chebyshev polynomials T_0..T_9
    long double Tm[kZTabDeg + 1][kZTabDeg + 1] = {{0}};
    Tm[0][0] = 1;
    Tm[1][1] = 1;
    for (int k = 2; k < n; k++)
        for (int j = 0; j < n; j++) Tm[k][j] = (j > 0 ? 2 * Tm[k - 1][j - 1] : 0) - Tm[k - 2][j];
    auto znet = [&](long double sv) {
        long double g0[8], z = 0;
        for (int j = 0; j < 8; j++) g0[j] = tanhl((long double)Z0[j] * sv);
        for (int j = 0; j < 8; j++) {
            long double acc = 0;
            for (int q = 0; q < 8; q++) acc += (long double)Z2[j][q] * g0[q];
            z += (long double)Z4[j] * tanhl(acc);
        }
        return z;
    };
    for (int i = 0; i < kZTabN; i++) {
        const long double a = (long double)i / kZTabPerUnit, b = (long double)(i + 1) / kZTabPerUnit;
        long double fx[kZTabDeg + 1];
        for (int k = 0; k < n; k++) fx[k] = znet((a + b) / 2 + (b - a) / 2 * cosl(PI * (2 * k + 1) / (2 * n)));
        long double mono[kZTabDeg + 1] = {0};  // in x = 2 y in [-1, 1]
        for (int m = 0; m < n; m++) {  // Chebyshev coefficient a_m
            long double am = 0;
            for (int k = 0; k < n; k++) am += fx[k] * cosl(PI * m * (2 * k + 1) / (2 * n));
            am *= (m == 0 ? 1.0L : 2.0L) / n;
            for (int j = 0; j < n; j++) mono[j] += am * Tm[m][j];
        }
        double* row = tab + (size_t)i * kZTabRow;
        long double scale = 1;  // x^j = 2^j y^j
        for (int j = 0; j < n; j++) { mono[j] *= scale; scale *= 2; row[j] = (double)mono[j]; }
    }
    // tail row: s0 >= kZTabSMax (every first-layer unit saturated): constant value, zero derivative
    double* last = tab + (size_t)kZTabN * kZTabRow;
    for (int j = 0; j < kZTabRow; j++) last[j] = 0.0;
    last[0] = (double)znet((long double)kZTabSMax);
}

struct NNAct {
    double s1, s2, s3;  // scaled xy features
    double h0[8], h2[8];
    double z4, dz4;     // z-branch value and its derivative w.r.t. the scaled sinkage
};

// in: contact velocity (vx, vy), wheel speed w, sinkage zr -> (Fx, Fy, Fz) before the damping term.
// Loop nests are ordered so that consecutive instructions belong to different accumulators (ILP).
template <bool WANT_D = false>
AVS_HD void tire_nn_fwd(const ModelConst& mc, double vx, double vy, double w, double zr, double F[3], NNAct& a) {
    const double s0 = zr * (1.0 / kInStd0);
    a.s1 = (w * kTireRadius - vx) * (1.0 / kInStd1);
    a.s2 = w * (1.0 / kInStd2);
    a.s3 = vy * (1.0 / kInStd3);
    // z-branch (table); for s0 <= 0 relu gives exactly 0 with zero derivative, as CondExpGt does
    a.dz4 = 0.0;
    double z4;
    if (WANT_D) {
        double dz;
        const double zt = ztab_eval_d(mc.ztab, s0, dz);
        z4 = s0 > 0.0 ? zt : 0.0;
        a.dz4 = s0 > 0.0 ? dz : 0.0;
    } else {
        const double zt = ztab_eval(mc.ztab, s0);  // branch-free: evaluated at max(s0, 0)
        z4 = s0 > 0.0 ? zt : 0.0;
    }
    a.z4 = z4;
#pragma unroll
    for (int j = 0; j < 8; j++) a.h0[j] = mc.W0[j][0] * a.s1;
#pragma unroll
    for (int j = 0; j < 8; j++) a.h0[j] = fma(mc.W0[j][1], a.s2, a.h0[j]);
#pragma unroll
    for (int j = 0; j < 8; j++) a.h0[j] = fma(mc.W0[j][2], a.s3, a.h0[j]);
    tanh_vec<8>(a.h0);
#pragma unroll
    for (int j = 0; j < 8; j++) a.h2[j] = mc.W2[j][0] * a.h0[0];
#pragma unroll
    for (int k = 1; k < 8; k++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a.h2[j] = fma(mc.W2[j][k], a.h0[k], a.h2[j]);
    }
    tanh_vec<8>(a.h2);
    // two output rows, each as two interleaved partial sums
    double f0a = mc.W4[0][0] * a.h2[0], f0b = mc.W4[0][1] * a.h2[1];
    double f1a = mc.W4[1][0] * a.h2[0], f1b = mc.W4[1][1] * a.h2[1];
#pragma unroll
    for (int k = 2; k < 8; k += 2) {
        f0a = fma(mc.W4[0][k], a.h2[k], f0a); f0b = fma(mc.W4[0][k + 1], a.h2[k + 1], f0b);
        f1a = fma(mc.W4[1][k], a.h2[k], f1a); f1b = fma(mc.W4[1][k + 1], a.h2[k + 1], f1b);
    }
    F[0] = (f0a + f0b) * kOutStd0;
    F[1] = (f1a + f1b) * kOutStd1;
    F[2] = (z4 > 0.0 ? z4 : 0.0) * kOutStd2;  // relu = CondExpGt(x,0,x,0) (TireNetwork.cpp:47-50)
}

// The weight gradient of one tire-network evaluation is the sum of three outer products
//   dW4 += (o0, o1) (x) h2 ,  dW2 += d2 (x) h0 ,  dW0 += d0 (x) (s1, s2, s3)        (pack order of TireNetwork.cpp:130-161)
// kGradOps = the 37 operands, in this order: o0 o1 | h2[8] | d2[8] | h0[8] | d0[8] | s1 s2 s3.
constexpr int kGradOps = 37;
template <class G> AVS_HD void grad_outer(G&& g, const double* op) {
    const double o0 = op[0], o1 = op[1];
    const double* h2 = op + 2; const double* d2 = op + 10; const double* h0 = op + 18; const double* d0 = op + 26; const double* sv = op + 34;
#pragma unroll
    for (int k = 0; k < 8; k++) { g[kOffW4 + k] = fma(o0, h2[k], g[kOffW4 + k]); g[kOffW4 + 8 + k] = fma(o1, h2[k], g[kOffW4 + 8 + k]); }
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int k = 0; k < 8; k++) g[kOffW2 + j * 8 + k] = fma(d2[j], h0[k], g[kOffW2 + j * 8 + k]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int c = 0; c < 3; c++) g[kOffW0 + j * 3 + c] = fma(d0[j], sv[c], g[kOffW0 + j * 3 + c]);
    }
}

// adjoint: Fbar[3] -> (vx_bar, vy_bar, zr_bar); the operands of the weight-gradient outer products are handed to
// acc.record(op[37]) — a host accumulator applies grad_outer at once, the device producer warp only stages them for
// the consumer warp that owns the gradient registers (see avs_bwd_kernel.cuh).
template <class Acc>
AVS_HD void tire_nn_bwd(const ModelConst& mc, const NNAct& a, const double Fbar[3], double& vxb, double& vyb, double& zrb, Acc& acc) {
    const double o0 = Fbar[0] * kOutStd0, o1 = Fbar[1] * kOutStd1;
    const double oz = (a.z4 > 0.0) ? Fbar[2] * kOutStd2 : 0.0;  // CondExpGt: derivative 0 at z4 <= 0
    double op[kGradOps];
    double* d2 = op + 10; double* d0 = op + 26;
    op[0] = o0; op[1] = o1;
#pragma unroll
    for (int k = 0; k < 8; k++) { op[2 + k] = a.h2[k]; op[18 + k] = a.h0[k]; }
    op[34] = a.s1; op[35] = a.s2; op[36] = a.s3;
#pragma unroll
    for (int k = 0; k < 8; k++) d2[k] = fma(mc.W4[1][k], o1, mc.W4[0][k] * o0);
#pragma unroll
    for (int k = 0; k < 8; k++) d2[k] *= fma(-a.h2[k], a.h2[k], 1.0);
#pragma unroll
    for (int k = 0; k < 8; k++) d0[k] = mc.W2[0][k] * d2[0];
#pragma unroll
    for (int j = 1; j < 8; j++) {
#pragma unroll
        for (int k = 0; k < 8; k++) d0[k] = fma(mc.W2[j][k], d2[j], d0[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) d0[k] *= fma(-a.h0[k], a.h0[k], 1.0);
    acc.record(op);
    double s1a = mc.W0[0][0] * d0[0], s1b = mc.W0[1][0] * d0[1];
    double s3a = mc.W0[0][2] * d0[0], s3b = mc.W0[1][2] * d0[1];
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        s1a = fma(mc.W0[j][0], d0[j], s1a); s1b = fma(mc.W0[j + 1][0], d0[j + 1], s1b);
        s3a = fma(mc.W0[j][2], d0[j], s3a); s3b = fma(mc.W0[j + 1][2], d0[j + 1], s3b);
    }
    // wheel speed (s2) is a control, not differentiated
    vxb = -(s1a + s1b) * (1.0 / kInStd1);
    vyb = (s3a + s3b) * (1.0 / kInStd3);
    zrb = (oz * a.dz4) * (1.0 / kInStd0);
}

// ------------------------------------------------------------------------------------------------
// Rotation helpers (src/auvsl_dynamics/utils.cpp:45-52, 20-28)
// ------------------------------------------------------------------------------------------------
AVS_HD void quat_to_R(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * y * y - 2 * z * z; R[1] = 2 * x * y - 2 * w * z;     R[2] = 2 * x * z + 2 * w * y;
    R[3] = 2 * x * y + 2 * w * z;     R[4] = 1 - 2 * x * x - 2 * z * z; R[5] = 2 * y * z - 2 * w * x;
    R[6] = 2 * x * z - 2 * w * y;     R[7] = 2 * y * z + 2 * w * x;     R[8] = 1 - 2 * x * x - 2 * y * y;
}
AVS_HD void quat_to_R_bwd(const double q[4], const double Rb[9], double qb[4]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double s01 = Rb[1] + Rb[3], s02 = Rb[2] + Rb[6], s12 = Rb[5] + Rb[7];
    const double a01 = Rb[3] - Rb[1], a02 = Rb[2] - Rb[6], a12 = Rb[7] - Rb[5];
    qb[0] += 2.0 * (y * s01 + z * s02 - 2.0 * x * (Rb[4] + Rb[8]) + w * a12);
    qb[1] += 2.0 * (x * s01 + z * s12 - 2.0 * y * (Rb[0] + Rb[8]) + w * a02);
    qb[2] += 2.0 * (x * s02 + y * s12 - 2.0 * z * (Rb[0] + Rb[4]) + w * a01);
    qb[3] += 2.0 * (z * a01 + y * a02 + x * a12);
}

// ---- activation record: what the reverse sweep needs from one tire-network evaluation ------------------
// Stored by the forward sweep (one record per wheel per RK4 stage) so that the adjoint does not re-evaluate the
// 16 tanh: kActLen = 20 contiguous doubles  h0[8] | h2[8] | dz = (z4 > 0 ? dz4/ds0 : 0) | pad | dalt/dx | dalt/dy,
// written and read as ten 16-byte pairs (STG.128 / 16-byte cp.async).  The scaled inputs s1..s3 are cheaper to
// recompute from the stage state than to store.
constexpr int kActLen = 20;
// Record in (global / host) memory as ten 16-byte pairs, pair i at p + i * pstride doubles.  pstride = 2: one
// contiguous 160-byte record (host harness).  The CUDA kernels use pstride = 64: the records of the 32 lanes of a warp
// are interleaved pair by pair ("warp tile", kActTile doubles), so that every 16-byte store / cp.async of a warp
// covers one contiguous 512-byte span instead of 32 sectors 160 bytes apart.
struct ActOutPtr {
    double* p;
    int pstride = 2;
    AVS_HD void store2(int i, double a, double b) const {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<double2*>(p + (size_t)i * pstride) = make_double2(a, b);
#else
        p[(size_t)i * pstride] = a; p[(size_t)i * pstride + 1] = b;
#endif
    }
    AVS_HD ActOutPtr wheel(int k) const { return ActOutPtr{p + (size_t)k * kActLen, pstride}; }  // contiguous records only
};
constexpr int kActPairStride = 64;           // doubles between the pairs of one lane inside a warp tile
constexpr int kActTile = kActLen * 32;       // doubles per warp tile (32 lanes x 20)
// doubles per RK4 stage of the tiled activation buffer for nc vehicles (whole 128-thread CTAs of the forward kernel)
AVS_HD size_t act_stage_doubles(size_t nc) { return ((4 * nc + 127) / 128) * 128 * kActLen; }
// this lane's base inside the tiled buffer of one stage (gid = global lane index = 4 * vehicle + wheel)
AVS_HD size_t act_lane_offset(size_t gid) { return (gid >> 5) * kActTile + (gid & 31) * 2; }
struct ActInPtr {
    const double* p;
    AVS_HD void load2(int i, double& a, double& b) const { a = p[2 * i]; b = p[2 * i + 1]; }
    AVS_HD ActInPtr wheel(int k) const { return ActInPtr{p + (size_t)k * kActLen}; }
};
// WITH_SLOPE = false (flat-terrain instantiation): the terrain gradient is identically 0 and the adjoint never reads it, so pair 9
// is neither written nor fetched (its slot stays in the layout: one STG.128 / cp.async per lane and stage and 10 % of the
// activation traffic less)
template <bool WITH_SLOPE = true, class AOut> AVS_HD void act_store(const NNAct& a, double dax, double day, const AOut& o) {
#pragma unroll
    for (int j = 0; j < 4; j++) { o.store2(j, a.h0[2 * j], a.h0[2 * j + 1]); o.store2(4 + j, a.h2[2 * j], a.h2[2 * j + 1]); }
    o.store2(8, a.z4 > 0.0 ? a.dz4 : 0.0, 0.0);
    if (WITH_SLOPE) o.store2(9, dax, day);
}
template <class AIn> AVS_HD void act_load(NNAct& a, double& dax, double& day, const AIn& in) {
#pragma unroll
    for (int j = 0; j < 4; j++) { in.load2(j, a.h0[2 * j], a.h0[2 * j + 1]); in.load2(4 + j, a.h2[2 * j], a.h2[2 * j + 1]); }
    double pad;
    in.load2(8, a.dz4, pad);
    a.z4 = 1.0;  // the relu gate is already folded into dz
    in.load2(9, dax, day);
}

// ------------------------------------------------------------------------------------------------
// One wheel: contact geometry -> tire force -> ABA passes 1 and 2 -> contribution X_i^T pa_i to the
// base bias force.  (HybridDynamics.cpp:214-253, 321-417; generated/forward_dynamics.cpp:56-153)
// ------------------------------------------------------------------------------------------------
struct WheelAct {
    NNAct nn;
    double dax, day;          // terrain gradient at the contact point
    double w0, w1, w2;        // wheel-frame angular velocity (with qd on the joint axis)
    double u0, u1, u2;        // wheel-frame linear velocity
    double n0, n1, n2;        // I_w * omega_w
};

struct NullAcc { AVS_HD void record(const double*) {} };

// Contact search of HybridDynamics::get_tire_cpts_sinkages_fancy (HybridDynamics.cpp:258-316; roty utils.cpp:5-17): ten
// candidate points on the rim, test_pos = joint + R roty(a) (0, 0, -r) = joint - r (R[:,0] sin a + R[:,2] cos a); the sinkage
// is the largest candidate sinkage (0 when none is positive: max_sinkage starts at 0), the reported point the JOINT centre.
// Forward only (the reference never tapes this function); dax / day are the terrain gradient at the selected candidate.
template <int TERR>
AVS_HD void wheel_contact_search(const ModelConst& mc, double tx, double ty, const double R[9], const double X[13], double& sink, double& dax, double& day, double cp[3]) {
    const double* p = X + 4;
#pragma unroll
    for (int i = 0; i < 3; i++) cp[i] = (R[3 * i] * tx + R[3 * i + 1] * ty + R[3 * i + 2] * kTz) + p[i];
    double best = 0.0;
    dax = 0.0; day = 0.0;
#pragma unroll 1
    for (int jj = 0; jj < 10; jj++) {
        const double sa = mc.cs_sin[jj], ca = mc.cs_cos[jj];
        double tp[3];
#pragma unroll
        for (int i = 0; i < 3; i++) tp[i] = cp[i] + (R[3 * i] * sa + R[3 * i + 2] * ca) * (-kTireRadius);
        double gx, gy;
        const double sk = altitude<TERR>(mc, tp[0], tp[1], gx, gy) - tp[2];
        if (sk > best) { best = sk; dax = gx; day = gy; }
    }
    sink = best;
}

// wheel index i: 0 FL(+,+) 1 FR(+,-) 2 RL(-,+) 3 RR(-,-); qd = commanded wheel speed
template <int TERR>
AVS_HD void wheel_contact(const ModelConst& mc, double tx, double ty, const double R[9], const double X[13], double& sink, double cv[3],
                          double& dax, double& day, double cp[3]) {
    const double* p = X + 4; const double* w = X + 7; const double* v = X + 10;
    if (TERR == TERR_DYN && mc.contact_search) {
        wheel_contact_search<TERR>(mc, tx, ty, R, X, sink, dax, day, cp);
    } else {
        cp[0] = fma(R[0], tx, fma(R[1], ty, fma(R[2], kRz, p[0])));
        cp[1] = fma(R[3], tx, fma(R[4], ty, fma(R[5], kRz, p[1])));
        cp[2] = fma(R[6], tx, fma(R[7], ty, fma(R[8], kRz, p[2])));
        const double alt = altitude<TERR>(mc, cp[0], cp[1], dax, day);
        sink = alt - cp[2];
    }
    // cv = v - r^ w = v + w x r       (HybridDynamics.cpp:336-346)
    cv[0] = v[0] + (w[1] * kRz - w[2] * ty);
    cv[1] = v[1] + (w[2] * tx - w[0] * kRz);
    cv[2] = v[2] + (w[0] * ty - w[1] * tx);
}

// ABA passes 1-2 for one wheel given the tire force (Fx, Fy, Fz) in the contact frame
AVS_HD void wheel_aba_fwd(double tx, double ty, const double X[13], double qd, const double F[3], double C[6], WheelAct& a) {
    const double* w = X + 7; const double* v = X + 10;
    // a = v + w x t ; omega_w = E w + qd z ; v_w = E a       (forward_dynamics.cpp:58-59)
    const double ax = v[0] + (w[1] * kTz - w[2] * ty);
    const double ay = v[1] + (w[2] * tx - w[0] * kTz);
    const double az = v[2] + (w[0] * ty - w[1] * tx);
    a.w0 = w[0]; a.w1 = -w[2]; a.w2 = w[1] + qd;
    a.u0 = ax;   a.u1 = -az;   a.u2 = ay;
    // c = crm(v_i)[:,AZ] * qd   (:62-63)
    const double c0 = qd * a.w1, c1 = -qd * a.w0, c3 = qd * a.u1, c4 = -qd * a.u0;
    // p = v x* (I v) - fext     (:66, :43)   wheel inertia: com = 0, Ixy = Ixz = 0
    a.n0 = kIwxx * a.w0;
    a.n1 = kIwyy * a.w1 - kIwyz * a.w2;
    a.n2 = kIwzz * a.w2 - kIwyz * a.w1;
    const double pA0 = a.w1 * a.n2 - a.w2 * a.n1;
    const double pA1 = a.w2 * a.n0 - a.w0 * a.n2;
    const double pA2 = a.w0 * a.n1 - a.w1 * a.n0;
    // fext (link frame) = (0,0,0, Fx, -Fz, Fy)   (HybridDynamics.cpp:405-411)
    const double pL0 = kMw * (a.w1 * a.u2 - a.w2 * a.u1) - F[0];
    const double pL1 = kMw * (a.w2 * a.u0 - a.w0 * a.u2) + F[2];
    const double pL2 = kMw * (a.w0 * a.u1 - a.w1 * a.u0) - F[1];
    // pa = p + Ia c + U u / D,  u = -p[AZ]     (:112-117)
    const double pa0 = fma(kIwxx, c0, pA0);
    const double pa1 = fma(kKappa, pA2, fma(kIwyyA, c1, pA1));
    const double pa3 = fma(kMw, c3, pL0);
    const double pa4 = fma(kMw, c4, pL1);
    const double pa5 = pL2;
    // X^T pa: force transform to the base frame   (:120)
    const double fb0 = pa3, fb1 = pa5, fb2 = -pa4;
    C[0] = pa0 + (ty * fb2 - kTz * fb1);
    C[1] = (kTz * fb0 - tx * fb2);
    C[2] = -pa1 + (tx * fb1 - ty * fb0);
    C[3] = fb0; C[4] = fb1; C[5] = fb2;
}

// the part of wheel_aba_fwd the adjoint needs (wheel-frame velocities and I_w w)
AVS_HD void wheel_aba_act(double tx, double ty, const double X[13], double qd, WheelAct& a) {
    const double* w = X + 7; const double* v = X + 10;
    a.w0 = w[0]; a.w1 = -w[2]; a.w2 = w[1] + qd;
    a.u0 = v[0] + (w[1] * kTz - w[2] * ty);
    a.u1 = -(v[2] + (w[0] * ty - w[1] * tx));
    a.u2 = v[1] + (w[2] * tx - w[0] * kTz);
    a.n0 = kIwxx * a.w0;
    a.n1 = kIwyy * a.w1 - kIwyz * a.w2;
    a.n2 = kIwzz * a.w2 - kIwyz * a.w1;
}

// adjoint of wheel_aba_fwd: Cb[6] -> Fb[3] (tire force adjoint) and += into wb[3], vb[3]
AVS_HD void wheel_aba_bwd(double tx, double ty, double qd, const WheelAct& a, const double Cb[6], double Fb[3], double wb[3], double vb[3]) {
    // C = (pa0, 0, -pa1) + t x fb ; fb
    const double pa0b = Cb[0], pa1b = -Cb[2];
    // fb_bar = Cb[3..5] + (a x t) with a = Cb[0..2]
    const double fb0b = Cb[3] + (Cb[1] * kTz - Cb[2] * ty);
    const double fb1b = Cb[4] + (Cb[2] * tx - Cb[0] * kTz);
    const double fb2b = Cb[5] + (Cb[0] * ty - Cb[1] * tx);
    const double pa3b = fb0b, pa5b = fb1b, pa4b = -fb2b;
    const double pA0b = pa0b, c0b = kIwxx * pa0b;
    const double pA1b = pa1b, c1b = kIwyyA * pa1b, pA2b = kKappa * pa1b;
    const double pL0b = pa3b, c3b = kMw * pa3b;
    const double pL1b = pa4b, c4b = kMw * pa4b;
    const double pL2b = pa5b;
    Fb[0] = -pL0b; Fb[2] = pL1b; Fb[1] = -pL2b;
    // pL = m w x u - ...:  w_bar += m (u x pLb), u_bar += m (pLb x w)
    double w0b = kMw * (a.u1 * pL2b - a.u2 * pL1b);
    double w1b = kMw * (a.u2 * pL0b - a.u0 * pL2b);
    double w2b = kMw * (a.u0 * pL1b - a.u1 * pL0b);
    double u0b = kMw * (pL1b * a.w2 - pL2b * a.w1);
    double u1b = kMw * (pL2b * a.w0 - pL0b * a.w2);
    double u2b = kMw * (pL0b * a.w1 - pL1b * a.w0);
    // pA = w x n: w_bar += n x pAb ; n_bar = pAb x w
    w0b += a.n1 * pA2b - a.n2 * pA1b;
    w1b += a.n2 * pA0b - a.n0 * pA2b;
    w2b += a.n0 * pA1b - a.n1 * pA0b;
    const double n0b = pA1b * a.w2 - pA2b * a.w1;
    const double n1b = pA2b * a.w0 - pA0b * a.w2;
    const double n2b = pA0b * a.w1 - pA1b * a.w0;
    // n = I_w w (symmetric)
    w0b += kIwxx * n0b;
    w1b += kIwyy * n1b - kIwyz * n2b;
    w2b += kIwzz * n2b - kIwyz * n1b;
    // c0 = qd w1, c1 = -qd w0, c3 = qd u1, c4 = -qd u0
    w1b += qd * c0b; w0b -= qd * c1b; u1b += qd * c3b; u0b -= qd * c4b;
    // omega_w = (wx, -wz, wy + qd) ; v_w = (ax, -az, ay)
    wb[0] += w0b; wb[2] -= w1b; wb[1] += w2b;
    const double axb = u0b, azb = -u1b, ayb = u2b;
    // a = v + w x t:  v_bar += ab ; w_bar += t x ab
    vb[0] += axb; vb[1] += ayb; vb[2] += azb;
    wb[0] += ty * azb - kTz * ayb;
    wb[1] += kTz * axb - tx * azb;
    wb[2] += tx * ayb - ty * axb;
}

// ------------------------------------------------------------------------------------------------
// Base link: bias force of the base + assembled acceleration + kinematics
//   (forward_dynamics.cpp:105, 156, 177; HybridDynamics.cpp:561-631)
// ------------------------------------------------------------------------------------------------
AVS_HD void base_inertia_mul(const double w[3], const double v[3], double nB[3], double fB[3]) {
    // n = Ibar w + m c x v ; f = m (v + w x c),  c = (cx, 0, cz)
    nB[0] = kIbxx * w[0] - kIbxy * w[1] - kIbxz * w[2] + kMb * (0.0 * v[2] - kCz * v[1]);
    nB[1] = -kIbxy * w[0] + kIbyy * w[1] - kIbyz * w[2] + kMb * (kCz * v[0] - kCx * v[2]);
    nB[2] = -kIbxz * w[0] - kIbyz * w[1] + kIbzz * w[2] + kMb * (kCx * v[1] - 0.0 * v[0]);
    fB[0] = kMb * (v[0] + (w[1] * kCz - w[2] * 0.0));
    fB[1] = kMb * (v[1] + (w[2] * kCx - w[0] * kCz));
    fB[2] = kMb * (v[2] + (w[0] * 0.0 - w[1] * kCx));
}

// S = summed wheel contributions; Xd = time derivative of the 13 live states
AVS_HD void base_fwd(const ModelConst& mc, const double R[9], const double X[13], const double S[6], double Xd[13]) {
    const double* q = X; const double* w = X + 7; const double* v = X + 10;
    double nB[3], fB[3], pb[6];
    base_inertia_mul(w, v, nB, fB);
    pb[0] = (w[1] * nB[2] - w[2] * nB[1]) + (v[1] * fB[2] - v[2] * fB[1]) + S[0];
    pb[1] = (w[2] * nB[0] - w[0] * nB[2]) + (v[2] * fB[0] - v[0] * fB[2]) + S[1];
    pb[2] = (w[0] * nB[1] - w[1] * nB[0]) + (v[0] * fB[1] - v[1] * fB[0]) + S[2];
    pb[3] = (w[1] * fB[2] - w[2] * fB[1]) + S[3];
    pb[4] = (w[2] * fB[0] - w[0] * fB[2]) + S[4];
    pb[5] = (w[0] * fB[1] - w[1] * fB[0]) + S[5];
    // quaternion rate (utils.cpp:20-28)
    Xd[0] = .5 * (q[3] * w[0] - q[2] * w[1] + q[1] * w[2]);
    Xd[1] = .5 * (q[2] * w[0] + q[3] * w[1] - q[0] * w[2]);
    Xd[2] = .5 * (-q[1] * w[0] + q[0] * w[1] + q[3] * w[2]);
    Xd[3] = .5 * (-q[0] * w[0] - q[1] * w[1] - q[2] * w[2]);
    // world-frame linear velocity (HybridDynamics.cpp:604)
    Xd[4] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    Xd[5] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    Xd[6] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    // a = -AI^-1 p + g_b,  g_b = R^T (0,0,-9.81)
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double s = mc.AIinv[r][0] * pb[0];
#pragma unroll
        for (int c = 1; c < 6; c++) s = fma(mc.AIinv[r][c], pb[c], s);
        Xd[7 + r] = -s;
    }
    Xd[10] += kGravity * R[6];
    Xd[11] += kGravity * R[7];
    Xd[12] += kGravity * R[8];
}

// adjoint of base_fwd, part 1: Sb[6] = -AIinv^T Xdb[7..12] = adjoint of every wheel contribution.
// XD is anything indexable (array or scratch column).
template <class XD> AVS_HD void base_bwd_S(const ModelConst& mc, const XD& Xdb, double Sb[6]) {
    const double l0 = Xdb[7];
#pragma unroll
    for (int c = 0; c < 6; c++) Sb[c] = mc.AIinv[0][c] * l0;
#pragma unroll
    for (int r = 1; r < 6; r++) {
        const double l = Xdb[7 + r];
#pragma unroll
        for (int c = 0; c < 6; c++) Sb[c] = fma(mc.AIinv[r][c], l, Sb[c]);
    }
#pragma unroll
    for (int c = 0; c < 6; c++) Sb[c] = -Sb[c];
}

// adjoint of base_fwd, part 2: everything that flows into X and R.  Xb[13] and Rb[9] are OVERWRITTEN.
template <class XD>
AVS_HD void base_bwd_rest(const double R[9], const double X[13], const XD& Xdb, const double Sb[6], double Xb[13], double Rb[9]) {
    const double* q = X; const double* w = X + 7; const double* v = X + 10;
    double* qb = Xb; double* wb = Xb + 7; double* vb = Xb + 10;
    double nB[3], fB[3];
    base_inertia_mul(w, v, nB, fB);
    const double* a = Sb; const double* b = Sb + 3;
    // pb_n = w x nB + v x fB ; pb_f = w x fB
    double nBb[3], fBb[3];
    wb[0] = (nB[1] * a[2] - nB[2] * a[1]) + (fB[1] * b[2] - fB[2] * b[1]);
    wb[1] = (nB[2] * a[0] - nB[0] * a[2]) + (fB[2] * b[0] - fB[0] * b[2]);
    wb[2] = (nB[0] * a[1] - nB[1] * a[0]) + (fB[0] * b[1] - fB[1] * b[0]);
    vb[0] = fB[1] * a[2] - fB[2] * a[1];
    vb[1] = fB[2] * a[0] - fB[0] * a[2];
    vb[2] = fB[0] * a[1] - fB[1] * a[0];
    nBb[0] = a[1] * w[2] - a[2] * w[1];
    nBb[1] = a[2] * w[0] - a[0] * w[2];
    nBb[2] = a[0] * w[1] - a[1] * w[0];
    fBb[0] = (a[1] * v[2] - a[2] * v[1]) + (b[1] * w[2] - b[2] * w[1]);
    fBb[1] = (a[2] * v[0] - a[0] * v[2]) + (b[2] * w[0] - b[0] * w[2]);
    fBb[2] = (a[0] * v[1] - a[1] * v[0]) + (b[0] * w[1] - b[1] * w[0]);
    // nB = Ibar w + m c x v  ->  w_bar += Ibar nBb ; v_bar += m (nBb x c)
    wb[0] += kIbxx * nBb[0] - kIbxy * nBb[1] - kIbxz * nBb[2];
    wb[1] += -kIbxy * nBb[0] + kIbyy * nBb[1] - kIbyz * nBb[2];
    wb[2] += -kIbxz * nBb[0] - kIbyz * nBb[1] + kIbzz * nBb[2];
    vb[0] += kMb * (nBb[1] * kCz - nBb[2] * 0.0);
    vb[1] += kMb * (nBb[2] * kCx - nBb[0] * kCz);
    vb[2] += kMb * (nBb[0] * 0.0 - nBb[1] * kCx);
    // fB = m v + m w x c    ->  v_bar += m fBb ; w_bar += m (c x fBb)
    vb[0] += kMb * fBb[0]; vb[1] += kMb * fBb[1]; vb[2] += kMb * fBb[2];
    wb[0] += kMb * (0.0 * fBb[2] - kCz * fBb[1]);
    wb[1] += kMb * (kCz * fBb[0] - kCx * fBb[2]);
    wb[2] += kMb * (kCx * fBb[1] - 0.0 * fBb[0]);
    // quaternion rate
    const double l0 = .5 * Xdb[0], l1 = .5 * Xdb[1], l2 = .5 * Xdb[2], l3 = .5 * Xdb[3];
    wb[0] += q[3] * l0 + q[2] * l1 - q[1] * l2 - q[0] * l3;
    wb[1] += -q[2] * l0 + q[3] * l1 + q[0] * l2 - q[1] * l3;
    wb[2] += q[1] * l0 - q[0] * l1 + q[3] * l2 - q[2] * l3;
    qb[0] = -w[2] * l1 + w[1] * l2 - w[0] * l3;
    qb[1] = w[2] * l0 - w[0] * l2 - w[1] * l3;
    qb[2] = -w[1] * l0 + w[0] * l1 - w[2] * l3;
    qb[3] = w[0] * l0 + w[1] * l1 + w[2] * l2;
    // pdot = R v ; g_b = R^T (0,0,g)
    const double p0 = Xdb[4], p1 = Xdb[5], p2 = Xdb[6];
    vb[0] += R[0] * p0 + R[3] * p1 + R[6] * p2;
    vb[1] += R[1] * p0 + R[4] * p1 + R[7] * p2;
    vb[2] += R[2] * p0 + R[5] * p1 + R[8] * p2;
    Rb[0] = p0 * v[0]; Rb[1] = p0 * v[1]; Rb[2] = p0 * v[2];
    Rb[3] = p1 * v[0]; Rb[4] = p1 * v[1]; Rb[5] = p1 * v[2];
    Rb[6] = fma(p2, v[0], kGravity * Xdb[10]); Rb[7] = fma(p2, v[1], kGravity * Xdb[11]); Rb[8] = fma(p2, v[2], kGravity * Xdb[12]);
    Xb[4] = 0.0; Xb[5] = 0.0; Xb[6] = 0.0;
}

}  // namespace avs

#include "avs_bekker.cuh"

namespace avs {

AVS_HD void wheel_sign(int i, double& tx, double& ty) {
    tx = (i & 2) ? -kTx : kTx;
    ty = (i & 1) ? -kTy : kTy;
}

// ------------------------------------------------------------------------------------------------
// Full ODE:  Xd = f(X; u)      (HybridDynamics::ODE, HybridDynamics.cpp:519-631)
//   lane0 = first wheel owned by this thread; ul, ur = commanded left / right wheel speeds
// ------------------------------------------------------------------------------------------------
//   ws (optional) = the four joint-velocity state entries X[17..20]: the tire models read THOSE as the wheel
//   speed (HybridDynamics.cpp:384, BekkerDynamics.cpp:68) while the ABA uses the control; inside integrate()
//   both are equal (VehicleSystem.cpp:132-133), so rollouts pass nullptr.
//   act_out (STORE_ACT, NN only) = activation record of this thread's first wheel; the records of its WPT
//   wheels are consecutive (kActLen doubles each)
template <int WPT, int TIRE, int TERR, bool STORE_ACT = false, class BekEval = BekNodesInline>
AVS_HD void ode_fwd(const ModelConst& mc, int lane0, const double X[13], double ul, double ur, double Xd[13], const double* ws = nullptr,
                    ActOutPtr act_out = ActOutPtr{nullptr}, const BekEval& bek_eval = BekEval()) {
    double R[9];
    quat_to_R(X, R);
    double S[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < WPT; k++) {
        const int i = lane0 + k;
        double tx, ty;
        wheel_sign(i, tx, ty);
        const double qd = (i & 1) ? ur : ul;  // qd = (vl, vr, vl, vr)  (HybridDynamics.cpp:539-542)
        double sink, cv[3], dax, day, cp[3], F[3], C[6];
        WheelAct act;
        wheel_contact<TERR>(mc, tx, ty, R, X, sink, cv, dax, day, cp);
        const double wsp = ws ? ws[i] : qd;
        if (TIRE == TIRE_NN) {
            tire_nn_fwd<STORE_ACT>(mc, cv[0], cv[1], wsp, sink, F, act.nn);
            F[2] = fma(kNNDamp, cv[2], F[2]);
#if defined(__CUDA_ARCH__)
            if (STORE_ACT) act_store<TERR != TERR_FLAT>(act.nn, dax, day, act_out.wheel(k));
#else
            if (STORE_ACT) act_store(act.nn, dax, day, act_out.wheel(k));  // the host harness reads whole records
#endif
        } else {
            bekker_tire_fwd<TERR>(mc, cv[0], cv[1], wsp, sink, cp, F, bek_eval);
            F[2] = fma(kBekkerDamp, cv[2], F[2]);
        }
        wheel_aba_fwd(tx, ty, X, qd, F, C, act);
#pragma unroll
        for (int c = 0; c < 6; c++) S[c] += C[c];
    }
#pragma unroll
    for (int c = 0; c < 6; c++) S[c] = Group<WPT>::sum(S[c]);
    base_fwd(mc, R, X, S, Xd);
}

// Adjoint of the ODE at stage input X:  Xb = (df/dX)^T Xdb (overwritten) ; weight gradients into acc.
// Xdb is indexable (scratch column).  The tire-network activations and terrain gradients come from the records
// the forward sweep stored, so neither the tanh nor the terrain lookup is
// re-evaluated (act_in.wheel(k) = record of this thread's k-th wheel).  The base-link part that needs X and R again runs after the wheel part so that nothing but Sb[6]
// has to stay live across the tire-network adjoint.
struct NoHook { AVS_HD void operator()() const {} };
// after_load() runs once the activation records have been read (the device uses it to start refilling the
// shared-memory staging slots with the next record while this stage computes)
template <int WPT, int TERR, class XD, class Acc, class AIn, class Hook = NoHook>
AVS_HD void ode_bwd_nn(const ModelConst& mc, int lane0, const double X[13], double ul, double ur, const XD& Xdb, double Xb[13], Acc& acc,
                       const AIn& act_in, const Hook& after_load = Hook()) {
    double Sb[6];
    base_bwd_S(mc, Xdb, Sb);
    // wheel-local adjoints, summed over the group: cp_bar (x) r  ->  Rb, pb ; wb ; vb
    double gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, gs[3] = {0, 0, 0};  // sum sx*cpb, sum sy*cpb, sum cpb
    double wb[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < WPT; k++) {
        const int i = lane0 + k;
        double tx, ty;
        wheel_sign(i, tx, ty);
        const double qd = (i & 1) ? ur : ul;
        double dax, day, Fb[3];
        WheelAct act;
        act_load(act.nn, dax, day, act_in.wheel(k));
        if (k == WPT - 1) after_load();
        {   // scaled tire-network inputs from the stage state (TireNetwork.cpp:73-89; cv = v + w x r)
            const double cvx = X[10] + (X[8] * kRz - X[9] * ty), cvy = X[11] + (X[9] * tx - X[7] * kRz);
            act.nn.s1 = (qd * kTireRadius - cvx) * (1.0 / kInStd1);
            act.nn.s2 = qd * (1.0 / kInStd2);
            act.nn.s3 = cvy * (1.0 / kInStd3);
        }
        wheel_aba_act(tx, ty, X, qd, act);
        wheel_aba_bwd(tx, ty, qd, act, Sb, Fb, wb, vb);
        double cvb[3], sinkb;
        cvb[2] = kNNDamp * Fb[2];
        tire_nn_bwd(mc, act.nn, Fb, cvb[0], cvb[1], sinkb, acc);
        // cv = v + w x r
        vb[0] += cvb[0]; vb[1] += cvb[1]; vb[2] += cvb[2];
        wb[0] += ty * cvb[2] - kRz * cvb[1];
        wb[1] += kRz * cvb[0] - tx * cvb[2];
        wb[2] += tx * cvb[1] - ty * cvb[0];
        // sink = alt(cpx, cpy) - cpz
        const double cpb0 = (TERR == TERR_FLAT) ? 0.0 : sinkb * dax, cpb1 = (TERR == TERR_FLAT) ? 0.0 : sinkb * day, cpb2 = -sinkb;
        const double sx = (i & 2) ? -1.0 : 1.0, sy = (i & 1) ? -1.0 : 1.0;
        gx[0] += sx * cpb0; gx[1] += sx * cpb1; gx[2] += sx * cpb2;
        gy[0] += sy * cpb0; gy[1] += sy * cpb1; gy[2] += sy * cpb2;
        gs[0] += cpb0; gs[1] += cpb1; gs[2] += cpb2;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // flat terrain: d(alt)/dx = d(alt)/dy = 0, so the x and y components of every cp_bar are exact zeros and six of
        // the fifteen group sums (each ~40 cycles of the single resident warp) drop out
        if (!(TERR == TERR_FLAT && c < 2)) { gx[c] = Group<WPT>::sum(gx[c]); gy[c] = Group<WPT>::sum(gy[c]); gs[c] = Group<WPT>::sum(gs[c]); }
        wb[c] = Group<WPT>::sum(wb[c]); vb[c] = Group<WPT>::sum(vb[c]);
    }
    double R[9], Rb[9];
    quat_to_R(X, R);
    base_bwd_rest(R, X, Xdb, Sb, Xb, Rb);
    // cp = R r + p with r = (sx*kTx, sy*kTy, kRz)
#pragma unroll
    for (int c = 0; c < 3; c++) {
        Rb[3 * c + 0] = fma(kTx, gx[c], Rb[3 * c + 0]);
        Rb[3 * c + 1] = fma(kTy, gy[c], Rb[3 * c + 1]);
        Rb[3 * c + 2] = fma(kRz, gs[c], Rb[3 * c + 2]);
        Xb[4 + c] += gs[c];
        Xb[7 + c] += wb[c];
        Xb[10 + c] += vb[c];
    }
    quat_to_R_bwd(X, Rb, Xb);
}

// The same adjoint with the LAYERED back-propagation of the tire network run right here, on the state chain (reverse kernel,
// end of round 2): `tin.get(k, h0, h2, dzg, dax, day)` delivers the stored activations of this thread's k-th wheel,
// `fout.put(k, o0, o1, s1, s2, s3, d2, d0)` hands over the operands of the three weight-gradient outer products
//   dW4 += (o0, o1) (x) h2 ,  dW2 += d2 (x) h0 ,  dW0 += d0 (x) (s1, s2, s3).
// Arithmetic and order are those of tire_nn_bwd, so the result is bit-identical to ode_bwd_nn.  Why not a separately formed
// 2x2 Jacobian of the xy-net for the state chain (the design until late in round 2): both rollout
// kernels turned out to be bound by issue slots (an FP64 instruction takes two, any other one; measured
// by moving the 2x2 Jacobian's ~216 FP64 instructions per stage into the forward kernel: it got slower by exactly their issue
// time wherever they were placed), and forming J on top of the layered pass costs ~210 FP64 instructions per wheel and stage
// more than the layered pass alone, which yields (vx_bar, vy_bar) as a by-product (16 more FMAs).
template <int WPT, int TERR, class XD, class TIn, class FOut>
AVS_HD void ode_bwd_nn_p(const ModelConst& mc, int lane0, const double X[13], double ul, double ur, const XD& Xdb, double Xb[13], TIn& tin, FOut& fout) {
    double Sb[6];
    base_bwd_S(mc, Xdb, Sb);
    double gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, gs[3] = {0, 0, 0};  // sum sx*cpb, sum sy*cpb, sum cpb
    double wb[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < WPT; k++) {
        const int i = lane0 + k;
        double tx, ty;
        wheel_sign(i, tx, ty);
        const double qd = (i & 1) ? ur : ul;
        double Fb[3];
        WheelAct act;
        const double cvx = X[10] + (X[8] * kRz - X[9] * ty), cvy = X[11] + (X[9] * tx - X[7] * kRz);
        const double s1 = (qd * kTireRadius - cvx) * (1.0 / kInStd1), s2 = qd * (1.0 / kInStd2), s3 = cvy * (1.0 / kInStd3);
        wheel_aba_act(tx, ty, X, qd, act);
        wheel_aba_bwd(tx, ty, qd, act, Sb, Fb, wb, vb);
        double h0[8], h2[8], dzg, dax, day;
        tin.get(k, h0, h2, dzg, dax, day);
        const double o0 = Fb[0] * kOutStd0, o1 = Fb[1] * kOutStd1;
        double d2[8], d0[8];
#pragma unroll
        for (int q = 0; q < 8; q++) d2[q] = fma(mc.W4[1][q], o1, mc.W4[0][q] * o0);
#pragma unroll
        for (int q = 0; q < 8; q++) d2[q] *= fma(-h2[q], h2[q], 1.0);
#pragma unroll
        for (int q = 0; q < 8; q++) d0[q] = mc.W2[0][q] * d2[0];
#pragma unroll
        for (int j = 1; j < 8; j++) {
#pragma unroll
            for (int q = 0; q < 8; q++) d0[q] = fma(mc.W2[j][q], d2[j], d0[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) d0[q] *= fma(-h0[q], h0[q], 1.0);
        fout.put(k, o0, o1, s1, s2, s3, d2, d0, h0, h2);
        double s1a = mc.W0[0][0] * d0[0], s1b = mc.W0[1][0] * d0[1];
        double s3a = mc.W0[0][2] * d0[0], s3b = mc.W0[1][2] * d0[1];
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            s1a = fma(mc.W0[j][0], d0[j], s1a); s1b = fma(mc.W0[j + 1][0], d0[j + 1], s1b);
            s3a = fma(mc.W0[j][2], d0[j], s3a); s3b = fma(mc.W0[j + 1][2], d0[j + 1], s3b);
        }
        double cvb[3];
        cvb[0] = -(s1a + s1b) * (1.0 / kInStd1);
        cvb[1] = (s3a + s3b) * (1.0 / kInStd3);
        cvb[2] = kNNDamp * Fb[2];
        const double sinkb = (Fb[2] * kOutStd2 * dzg) * (1.0 / kInStd0);
        // cv = v + w x r
        vb[0] += cvb[0]; vb[1] += cvb[1]; vb[2] += cvb[2];
        wb[0] += ty * cvb[2] - kRz * cvb[1];
        wb[1] += kRz * cvb[0] - tx * cvb[2];
        wb[2] += tx * cvb[1] - ty * cvb[0];
        // sink = alt(cpx, cpy) - cpz
        const double cpb0 = (TERR == TERR_FLAT) ? 0.0 : sinkb * dax, cpb1 = (TERR == TERR_FLAT) ? 0.0 : sinkb * day, cpb2 = -sinkb;
        const double sx = (i & 2) ? -1.0 : 1.0, sy = (i & 1) ? -1.0 : 1.0;
        gx[0] += sx * cpb0; gx[1] += sx * cpb1; gx[2] += sx * cpb2;
        gy[0] += sy * cpb0; gy[1] += sy * cpb1; gy[2] += sy * cpb2;
        gs[0] += cpb0; gs[1] += cpb1; gs[2] += cpb2;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (!(TERR == TERR_FLAT && c < 2)) { gx[c] = Group<WPT>::sum(gx[c]); gy[c] = Group<WPT>::sum(gy[c]); gs[c] = Group<WPT>::sum(gs[c]); }
        wb[c] = Group<WPT>::sum(wb[c]); vb[c] = Group<WPT>::sum(vb[c]);
    }
    double R[9], Rb[9];
    quat_to_R(X, R);
    base_bwd_rest(R, X, Xdb, Sb, Xb, Rb);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        Rb[3 * c + 0] = fma(kTx, gx[c], Rb[3 * c + 0]);
        Rb[3 * c + 1] = fma(kTy, gy[c], Rb[3 * c + 1]);
        Rb[3 * c + 2] = fma(kRz, gs[c], Rb[3 * c + 2]);
        Xb[4 + c] += gs[c];
        Xb[7 + c] += wb[c];
        Xb[10 + c] += vb[c];
    }
    quat_to_R_bwd(X, Rb, Xb);
}
// host-side provider for ode_bwd_nn_p: activations from four consecutive records, weight gradient applied at once
template <class Acc> struct LayeredFromRecords {
    const double* act; Acc& acc;
    AVS_HD void get(int k, double h0[8], double h2[8], double& dzg, double& dax, double& day) {
        NNAct a;
        act_load(a, dax, day, ActInPtr{act + (size_t)k * kActLen});
        dzg = a.dz4;
#pragma unroll
        for (int i = 0; i < 8; i++) { h0[i] = a.h0[i]; h2[i] = a.h2[i]; }
    }
    AVS_HD void put(int, double o0, double o1, double s1, double s2, double s3, const double d2[8], const double d0[8], const double h0[8], const double h2[8]) {
        double op[kGradOps];
        op[0] = o0; op[1] = o1;
#pragma unroll
        for (int i = 0; i < 8; i++) { op[2 + i] = h2[i]; op[10 + i] = d2[i]; op[18 + i] = h0[i]; op[26 + i] = d0[i]; }
        op[34] = s1; op[35] = s2; op[36] = s3;
        acc.record(op);
    }
};
// ------------------------------------------------------------------------------------------------
// RK4 step with per-stage quaternion re-normalisation (HybridDynamics::RK4, HybridDynamics.cpp:461-502)
// ------------------------------------------------------------------------------------------------
template <class V> AVS_HD void renorm(V&& Y, double& inv_norm) {
    const double q0 = Y[0], q1 = Y[1], q2 = Y[2], q3 = Y[3];
    inv_norm = rsqrt_f64(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    Y[0] = q0 * inv_norm; Y[1] = q1 * inv_norm; Y[2] = q2 * inv_norm; Y[3] = q3 * inv_norm;
}
// X = Y/|Y| on the quaternion part: Yb = (Xb - X (X.Xb)) / |Y|, in place
template <class VX, class VB> AVS_HD void renorm_bwd(const VX& Xn, double inv_norm, VB&& Xb) {
    const double x0 = Xn[0], x1 = Xn[1], x2 = Xn[2], x3 = Xn[3];
    const double b0 = Xb[0], b1 = Xb[1], b2 = Xb[2], b3 = Xb[3];
    const double d = x0 * b0 + x1 * b1 + x2 * b2 + x3 * b3;
    Xb[0] = (b0 - x0 * d) * inv_norm; Xb[1] = (b1 - x1 * d) * inv_norm;
    Xb[2] = (b2 - x2 * d) * inv_norm; Xb[3] = (b3 - x3 * d) * inv_norm;
}

// Stage sink: receives every (normalised) stage input and the inverse norm that goes into its record (stages 1-3:
// the norm that produced that stage input; stage 0: the PREVIOUS step's final normalisation).  NullSink for plain rollouts.
struct NullSink {
    template <class V> AVS_HD void stage(int, const V&, double) {}
};

// One RK4 step.  All loop-carried state lives in per-thread scratch columns (X in/out; ACC, T, K scratch,
// 13 slots each); `ode(T, K, s)` evaluates K = f(T) for stage s.  On the device `ode` is a call to a NON-inlined function:
// ptxas schedules the body of a rolled loop for minimum register pressure and re-serialises independent
// dependency chains, whereas a stand-alone function body keeps the interleaved (ILP) order it is written in.
// The four stages run as a rolled loop; the stage weights of HybridDynamics.cpp:470-495 are selected by index.
template <class Ode, class Sink, class St>
AVS_HD void rk4_fwd(Ode& ode, St X, St ACC, St T, St K, Sink& sink, double& inv_carry) {
    // inv_carry: in = inverse norm of the previous step's final normalisation (rides in this step's stage-0 record),
    //            out = this step's
    double inv = inv_carry;
    const double h = kDt;
#pragma unroll
    for (int c = 0; c < 13; c++) { ACC[c] = 0.0; T[c] = X[c]; }
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        sink.stage(s, T, inv);
        ode(T, K, s);
        const double wgt = (s == 0 || s == 3) ? 1.0 : 2.0;   // k1 + 2 k2 + 2 k3 + k4
        const double cs = (s == 2) ? h : .5 * h;             // X + .5h k1, X + .5h k2, X + h k3
#pragma unroll
        for (int c = 0; c < 13; c++) { const double k = K[c]; ACC[c] = fma(wgt, k, ACC[c]); T[c] = fma(cs, k, X[c]); }
        renorm(T, inv);
    }
#pragma unroll
    for (int c = 0; c < 13; c++) T[c] = fma(h / 6.0, ACC[c], X[c]);
    renorm(T, inv);
#pragma unroll
    for (int c = 0; c < 13; c++) X[c] = T[c];
    inv_carry = inv;
}

// Adjoint of one RK4 step over a reverse RECORD STREAM.  The forward sweep stores, per RK4 step, four records
// (one per stage): 13 state entries + slot 13.  Slot 13 of a stage-s record (s = 1..3) holds the inverse norm
// that produced that stage input; slot 13 of a stage-0 record holds the inverse norm of the PREVIOUS step's final
// normalisation (so that the record that serves as "state after the step" also delivers that step's norm).
// On entry st.cur() is the record after this step; st.advance() steps to the previous record (and prefetches the
// one before it).  `odeb(XS, LAM, YB, KB, s, act)` performs the adjoint of stage s (see rk4_bwd_stage_update; it also
// accumulates the weight gradient); act = the stage's activation records.
// LAM (13, in/out) = adjoint of the step output on entry, adjoint of the step input on return; YB, KB scratch.
template <class OdeB, class Stream, class St>
AVS_HD void rk4_bwd_nn(OdeB& odeb, Stream& st, St LAM, St YB, St KB) {
    const double h = kDt;
    {
        const auto XN = st.cur();
        const double inv_n = XN[13];
        renorm_bwd(XN, inv_n, LAM);  // adjoint of Y+ = X + h/6 (k1 + 2k2 + 2k3 + k4) through the final normalisation
    }
#pragma unroll
    for (int c = 0; c < 13; c++) { const double l = LAM[c]; YB[c] = l; KB[c] = (h / 6.0) * l; }
#pragma unroll 1
    for (int s = 3; s >= 0; s--) {
        st.advance();
        odeb(st.cur(), LAM, YB, KB, s, st.act());  // LAM += J^T KB (through the stage normalisation); KB <- adjoint of k_{s-1}
    }
}

// The per-stage update shared by the host functor and the device callee: given sb = (df/dX)^T KB at the stage input
// XS (record with its inverse norm in slot 13), fold it into the running adjoints.
template <class XSv, class St>
AVS_HD void rk4_bwd_stage_update(const XSv& XS, double sb[13], int s, St LAM, St YB, St KB) {
    const double h = kDt;
    if (s > 0) {
        renorm_bwd(XS, XS[13], sb);  // stage input s = N(X + c_s k_s)
        const double cs = (s == 3) ? h : .5 * h;
        const double wk = (s == 1) ? (h / 6.0) : (h / 3.0);  // weight of k_s in the output combination
#pragma unroll
        for (int c = 0; c < 13; c++) { LAM[c] += sb[c]; KB[c] = fma(cs, sb[c], wk * YB[c]); }
    } else {
#pragma unroll
        for (int c = 0; c < 13; c++) LAM[c] += sb[c];
    }
}

// inline ODE functors (host harness; the device kernels use non-inlined calls, see avs_dev.cuh)
template <int WPT, int TIRE, int TERR, bool STORE_ACT> struct OdeFwdInline {
    const ModelConst& mc; int lane0; double ul, ur;
    double* act_step;  // activation records of the current RK4 step (this thread's first wheel, stage 0)
    size_t act_stage_stride;  // doubles between the records of consecutive stages
    template <class St> AVS_HD void operator()(St T, St K, int s) const {
        double X[13], Xd[13];
#pragma unroll
        for (int c = 0; c < 13; c++) X[c] = T[c];
        ode_fwd<WPT, TIRE, TERR, STORE_ACT>(mc, lane0, X, ul, ur, Xd, nullptr, ActOutPtr{STORE_ACT ? act_step + (size_t)s * act_stage_stride : nullptr});
#pragma unroll
        for (int c = 0; c < 13; c++) K[c] = Xd[c];
    }
};
template <int WPT, int TERR, class Acc> struct OdeBwdInline {
    const ModelConst& mc; int lane0; double ul, ur; Acc& acc;
    template <class XSv, class St> AVS_HD void operator()(const XSv& XS, St LAM, St YB, St KB, int s, const double* act) const {
        double X[13], Xb[13];
#pragma unroll
        for (int c = 0; c < 13; c++) X[c] = XS[c];
        LayeredFromRecords<Acc> io{act, acc};  // the adjoint the reverse kernel's producer warp runs (bit-identical to ode_bwd_nn)
        ode_bwd_nn_p<WPT, TERR>(mc, lane0, X, ul, ur, KB, Xb, io, io);
        rk4_bwd_stage_update(XS, Xb, s, LAM, YB, KB);
    }
};

// ------------------------------------------------------------------------------------------------
// Loss of one row (VehicleSystem::loss, src/VehicleSystem.cpp:70-92) and its gradient w.r.t. the
// 13 live states:  l = |wrap(yaw - yaw_gt)| + sqrt((x_gt - x)^2 + (y_gt - y)^2)
// ------------------------------------------------------------------------------------------------
AVS_HD double row_loss(const double X[13], double yaw_gt, double x_gt, double y_gt, double scale, double lam[13]) {
    const double qx = X[0], qy = X[1], qz = X[2], qw = X[3];
    const double xe = x_gt - X[4], ye = y_gt - X[5];
    const double lin = sqrt(xe * xe + ye * ye);
    const double s = 2 * (qw * qz + qx * qy), c = 1 - 2 * (qy * qy + qz * qz);  // utils.cpp:75-77
    const double yaw = atan2(s, c);
    const double e = yaw - yaw_gt;
    const double wr = atan2(sin(e), cos(e));
    if (lam) {
        if (lin > 0.0) {
            lam[4] -= scale * xe / lin;
            lam[5] -= scale * ye / lin;
        }
        const double sg = wr > 0.0 ? 1.0 : (wr < 0.0 ? -1.0 : 0.0);  // CppAD abs'(0) = 0
        const double g = scale * sg / (s * s + c * c);
        const double sb = g * c, cb = -g * s;
        lam[0] += sb * 2 * qy;
        lam[1] += sb * 2 * qx + cb * (-4 * qy);
        lam[2] += sb * 2 * qw + cb * (-4 * qz);
        lam[3] += sb * 2 * qz;
    }
    return fabs(wr) + lin;
}

// AI_b^-1: host-side precomputation (double; called once per context)
inline void compute_AIinv(double out[6][6]) {
    // I_b (generated/inertia_properties.cpp:8-12)
    double AI[6][6] = {{0}};
    AI[0][0] = kIbxx;  AI[0][1] = -kIbxy; AI[0][2] = -kIbxz;
    AI[1][0] = -kIbxy; AI[1][1] = kIbyy;  AI[1][2] = -kIbyz;
    AI[2][0] = -kIbxz; AI[2][1] = -kIbyz; AI[2][2] = kIbzz;
    const double c[3] = {kCx, 0.0, kCz};
    const double cx[3][3] = {{0, -c[2], c[1]}, {c[2], 0, -c[0]}, {-c[1], c[0], 0}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { AI[i][3 + j] = kMb * cx[i][j]; AI[3 + i][j] = -kMb * cx[i][j]; }
    for (int i = 0; i < 3; i++) AI[3 + i][3 + i] = kMb;
    // + sum_i X_i^T Ia X_i, Ia = diag-ish(Ixx, Iyy', 0 | m, m, m) in the wheel frame
    for (int w = 0; w < 4; w++) {
        double tx, ty;
        wheel_sign(w, tx, ty);
        double X[6][6] = {{0}};
        X[0][0] = 1; X[1][2] = -1; X[2][1] = 1;
        X[3][1] = kTz; X[3][2] = -ty; X[3][3] = 1;
        X[4][0] = -ty; X[4][1] = tx;  X[4][5] = -1;
        X[5][0] = -kTz; X[5][2] = tx; X[5][4] = 1;
        const double Ia[6] = {kIwxx, kIwyyA, 0.0, kMw, kMw, kMw};
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += X[k][i] * Ia[k] * X[k][j];
            AI[i][j] += s;
        }
    }
    // Gauss-Jordan inverse with partial pivoting (6x6, SPD, cond ~ 50)
    double M[6][12];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { M[i][j] = AI[i][j]; M[i][6 + j] = (i == j) ? 1.0 : 0.0; }
    for (int col = 0; col < 6; col++) {
        int piv = col;
        for (int r = col + 1; r < 6; r++) if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
        if (piv != col) for (int j = 0; j < 12; j++) { double t = M[col][j]; M[col][j] = M[piv][j]; M[piv][j] = t; }
        const double d = 1.0 / M[col][col];
        for (int j = 0; j < 12; j++) M[col][j] *= d;
        for (int r = 0; r < 6; r++) if (r != col) {
            const double f = M[r][col];
            for (int j = 0; j < 12; j++) M[r][j] -= f * M[col][j];
        }
    }
    // symmetrise
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) out[i][j] = 0.5 * (M[i][6 + j] + M[j][6 + i]);
}

}  // namespace avs
