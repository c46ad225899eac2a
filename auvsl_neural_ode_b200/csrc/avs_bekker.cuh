// avs_bekker.cuh — Bekker/Wong terramechanics tire model, forward only.
// (scalar walk = the reference's structure; the front / rear arcs normally go through the vectorised quadrature below)
//
// Restates BekkerDynamics::get_tire_f_ext (src/auvsl_dynamics/BekkerDynamics.cpp:31-120) and
// BekkerTireModel::{get_features, get_forces, integrate} (src/auvsl_dynamics/BekkerTireModel.cpp:
// 163-278) for one wheel.  Folding applied (results unchanged up to rounding):
//   * the three Ty integrals are computed and thrown away by the reference (:266-275) — skipped
//   * each contact-arc region (front cf, centre cc, rear cr) is walked ONCE: the integrands of
//     Fx, Fy, Fz share sigma/j/tau at every quadrature node, and the trapezoid rule of :163-179
//     evaluates every interior node twice with identical arguments
//   * the centre region has zero width whenever ze == zr (pgc < pg), which the reference integrates
//     to an exact 0 — skipped in that case
// Included from avs_model.cuh (needs ModelConst / GridTerrain / grid_cell / grid_lerp).
#pragma once

namespace avs {

struct BekkerCtx {
    double R, b, K, c, pg;
    double slip_ratio, tan_sa, tan_phi, kk, ku, n, ze;
    double theta_f, theta_r, theta_c, cos_tf, sin_tf, cos_tr, cos_tc, sin_tc;
};

struct Bek3 { double fx, fy, fz; };

// integrand values at one node of the front (REGION 0), centre (1) or rear (2) arc
template <int REGION> AVS_HD Bek3 bekker_node(const BekkerCtx& k, double th) {
    const double s = sin(th), c = cos(th);
    double sigma, jx, jy;
    const double om = 1.0 - k.slip_ratio;
    if (REGION == 0) {
        sigma = k.kk * pow(k.R * (c - k.cos_tf), k.n);                                         // :18-24
        jx = k.R * ((th - k.theta_f) - om * (s - k.sin_tf));                                   // :37-40
        jy = k.R * om * k.tan_sa * (k.theta_f - th);                                           // :51-54
    } else if (REGION == 1) {
        const double tn = tan(th);
        sigma = k.kk * pow(k.ze, k.n);                                                         // :26-29
        jx = k.R * ((k.theta_c - k.theta_f) + om * k.sin_tf - k.sin_tc + (k.slip_ratio * k.cos_tc * tn));  // :42-45
        jy = k.R * om * k.tan_sa * (k.theta_f - th + k.sin_tc - k.cos_tc * tn);                // :55-58
    } else {
        sigma = k.ku * pow(k.R * (c - k.cos_tr), k.n);                                         // :31-35
        jx = k.R * ((2 * k.theta_c + th - k.theta_f) - om * (s - k.sin_tf) - 2 * k.sin_tc);    // :46-49
        jy = k.R * om * k.tan_sa * (k.theta_f - th - 2 * k.theta_c + 2 * k.sin_tc);            // :59-62
    }
    double j = sqrt(jy * jy + jx * jx);                                                        // :64-78
    if (j == 0.0) j = 1e-6;                                                                    // "smart divide" :83-84
    const double shear = (k.c + sigma * k.tan_phi) * (1.0 - exp(-j / k.K));
    const double tau_x = (-jx / j) * shear;                                                    // :80-104
    const double tau_y = (-jy / j) * shear;                                                    // :106-129
    Bek3 r;
    if (REGION == 1) {
        const double sec2 = (1.0 / c) * (1.0 / c);
        r.fx = tau_x * sec2;  // Fx_eqn3 :155-158
        r.fy = tau_y * sec2;  // Fy_eqn3 :132-135
        r.fz = 0.0;
    } else {
        r.fx = tau_x * c - sigma * s;  // Fx_eqn1/2 :146-154
        r.fy = tau_y;                  // tau_y_cf / tau_y_cr
        r.fz = sigma * c + tau_x * s;  // Fz_eqn1/2 :137-144
    }
    return r;
}

// BekkerTireModel::integrate (:163-179) for the three force integrands of one region at once
// On the device this scalar walk is the COLD path (degenerate arcs, the centre arc when the ground pressure saturates): kept out
// of line, so that its ~30 inlined libm calls per region do not sit between the hot blocks of the stage (rollout kernel 41.0 ->
// 40.5 ms at 8192 x 100 rows; with the context passed by value instead of by reference: 40.9)
#if defined(__CUDACC__)
template <int REGION> __host__ __device__ __noinline__ Bek3 bekker_integrate(const BekkerCtx& k, double upper_b, double lower_b) {
#else
template <int REGION> AVS_HD Bek3 bekker_integrate(const BekkerCtx& k, double upper_b, double lower_b) {
#endif
    const double dtheta = (upper_b - lower_b) / 10.0;
    const double eps = dtheta * .1;
    Bek3 sum = {0.0, 0.0, 0.0};
    double theta = lower_b;
    Bek3 fl = bekker_node<REGION>(k, theta);
    const double stop = upper_b - eps - dtheta;
    // value-dependent loop bound kept (9 iterations whenever dtheta > 0)
    for (int it = 0; it < 16 && theta < stop; it++) {
        const Bek3 fr = bekker_node<REGION>(k, theta + dtheta);
        sum.fx += .5 * dtheta * (fr.fx + fl.fx);
        sum.fy += .5 * dtheta * (fr.fy + fl.fy);
        sum.fz += .5 * dtheta * (fr.fz + fl.fz);
        fl = fr;
        theta += dtheta;
    }
    const Bek3 fu = bekker_node<REGION>(k, upper_b);
    const Bek3 fm = bekker_node<REGION>(k, upper_b - dtheta);
    sum.fx += .5 * dtheta * (fu.fx + fm.fx);
    sum.fy += .5 * dtheta * (fu.fy + fm.fy);
    sum.fz += .5 * dtheta * (fu.fz + fm.fz);
    return sum;
}

// ------------------------------------------------------------------------------------------------
// Vectorised quadrature.  The reference's scalar walk is ~30 libm calls per arc on one dependent chain, with
// branches (slow paths of pow / sin / cos) that keep ptxas from interleaving anything: measured 42 % of the FP64
// instructions consume the result of the one before, and the code does not fit the instruction cache.  Here the 12
// node evaluations of an arc (t0 .. t9 of the running sum, upper - dtheta, upper) are issued step by step for four
// nodes at a time (measured best of 4 / 6 / 12) from branch-free elementary functions:
//   * sin / cos of the nodes by rotating (sin, cos)(lower) with (sin, cos)(dtheta); the two end nodes take the
//     context's own values, so the sigma base R (cos - cos_end) is an exact 0 there, as in the reference;
//   * pow(b, n) = exp(n log b) with a table-free log (fdlibm kernel) and the exp of tanh_f64;
//   * 1/j from one rsqrt (MUFU seed + two Newton steps), j = q / sqrt(q).
// Results agree with the scalar walk to ~1e-14 relative (tests/test_device_math_cpu.py compares both with the oracle).
// ------------------------------------------------------------------------------------------------
AVS_HD void split_hi_lo(double x, int& hi, int& lo) {
#if defined(__CUDA_ARCH__)
    hi = __double2hiint(x); lo = __double2loint(x);
#else
    int64_t b; memcpy(&b, &x, 8); hi = (int)(b >> 32); lo = (int)(b & 0xffffffff);
#endif
}
AVS_HD double join_hi_lo(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
AVS_HD double rcp_seed(double d) {
#if defined(__CUDA_ARCH__)
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r;
#else
    return 1.0 / d;
#endif
}
AVS_HD double rsqrt_seed(double d) {
#if defined(__CUDA_ARCH__)
    double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r;
#else
    return 1.0 / sqrt(d);
#endif
}

// W-wide natural logarithm of non-negative normal numbers (0 maps to about -709), in place.  fdlibm's kernel:
// x = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f / (2 + f), log m = f - f^2/2 + s (f^2/2 + R(s^2)).
template <int W> AVS_HD void log_vec(double* x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    double f[W], dk[W], s[W], z[W], w[W], t1[W], t2[W], h[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
        int hi, lo;
        split_hi_lo(x[i], hi, lo);
        hi += 0x3ff00000 - 0x3fe6a09e;
        dk[i] = (double)((hi >> 20) - 0x3ff);
        hi = (hi & 0x000fffff) + 0x3fe6a09e;
        f[i] = join_hi_lo(hi, lo) - 1.0;
    }
    // s = f / (2 + f): reciprocal seed + two Newton steps + residual correction
#pragma unroll
    for (int i = 0; i < W; i++) h[i] = 2.0 + f[i];
#pragma unroll
    for (int i = 0; i < W; i++) w[i] = rcp_seed(h[i]);
#pragma unroll
    for (int i = 0; i < W; i++) t1[i] = fma(-h[i], w[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) t1[i] = fma(t1[i], t1[i], t1[i]);
#pragma unroll
    for (int i = 0; i < W; i++) w[i] = fma(w[i], t1[i], w[i]);
#pragma unroll
    for (int i = 0; i < W; i++) s[i] = f[i] * w[i];
#pragma unroll
    for (int i = 0; i < W; i++) t1[i] = fma(-h[i], s[i], f[i]);
#pragma unroll
    for (int i = 0; i < W; i++) s[i] = fma(t1[i], w[i], s[i]);
#pragma unroll
    for (int i = 0; i < W; i++) z[i] = s[i] * s[i];
#pragma unroll
    for (int i = 0; i < W; i++) w[i] = z[i] * z[i];
#pragma unroll
    for (int i = 0; i < W; i++) { t1[i] = fma(w[i], Lg6, Lg4); t2[i] = fma(w[i], Lg7, Lg5); }
#pragma unroll
    for (int i = 0; i < W; i++) { t1[i] = fma(w[i], t1[i], Lg2); t2[i] = fma(w[i], t2[i], Lg3); }
#pragma unroll
    for (int i = 0; i < W; i++) { t1[i] = w[i] * t1[i]; t2[i] = fma(w[i], t2[i], Lg1); }
#pragma unroll
    for (int i = 0; i < W; i++) t2[i] = fma(z[i], t2[i], t1[i]);  // R
#pragma unroll
    for (int i = 0; i < W; i++) h[i] = 0.5 * f[i] * f[i];
#pragma unroll
    for (int i = 0; i < W; i++) t1[i] = fma(s[i], h[i] + t2[i], dk[i] * LN2_LO);
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = fma(dk[i], LN2_HI, f[i] - (h[i] - t1[i]));
}

// W-wide exp for arguments in about [-700, 700], in place (same reduction and polynomial as exp_negarg)
template <int W> AVS_HD void exp_vec(double* x) {
    const double L2E = 1.4426950408889634e+0, MAGIC = 6755399441055744.0;
    const double LN2_HI = 6.93147180559945286e-01, LN2_LO = 2.31904681384629956e-17;
    double t[W], r[W], p[W];
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = fma(x[i], L2E, MAGIC);
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = t[i] - MAGIC;
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(r[i], -LN2_HI, x[i]);
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = fma(r[i], -LN2_LO, p[i]);
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(2.5110049204818658e-08, r[i], 2.763265472252779e-07);
#define AVS_POLY_STEP(C) \
    _Pragma("unroll") for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], C);
    AVS_POLY_STEP(2.755724088722987e-06)
    AVS_POLY_STEP(2.4801485441561313e-05)
    AVS_POLY_STEP(0.00019841269890076403)
    AVS_POLY_STEP(0.0013888888952352863)
    AVS_POLY_STEP(0.008333333333319589)
    AVS_POLY_STEP(0.04166666666648795)
    AVS_POLY_STEP(0.1666666666666668)
    AVS_POLY_STEP(0.5000000000000019)
    AVS_POLY_STEP(1.0)
    AVS_POLY_STEP(1.0)
#undef AVS_POLY_STEP
#pragma unroll
    for (int i = 0; i < W; i++) {
#if defined(__CUDA_ARCH__)
        int k = __double2loint(t[i]);
        k = min(max(k, -1000), 1000);
        x[i] = __hiloint2double(__double2hiint(p[i]) + (k << 20), __double2loint(p[i]));
#else
        int64_t k = (int64_t)(t[i] - MAGIC);
        if (k < -1000) k = -1000;
        if (k > 1000) k = 1000;
        x[i] = ldexp(p[i], (int)k);
#endif
    }
}
AVS_HD double pow_pos(double b, double e) {  // b >= 0; pow(0, e > 0) comes out as ~1e-246 instead of 0
    double v[1] = {b};
    log_vec<1>(v);
    v[0] *= e;
    exp_vec<1>(v);
    return v[0];
}

// Everything a node of the front / rear arc needs, as run-time values: one routine serves both arcs.
//   sigma = coef (R (cos th - cos_end))^n                                  (:18-35)
//   jx = R ((th + A) - om (sin th - sin_tf) - B)                           (:37-49)   front: A = -theta_f, B = 0
//   jy = RomT (C - th),  RomT = R om tan(slip angle)                       (:51-62)   rear: A = 2 theta_c - theta_f, B = 2 sin theta_c
struct BekNodeParams {
    double R, n, cos_end, coef, A, om, sin_tf, B, RomT, C, invK, c_soil, tan_phi;
};
constexpr int kBekNodeParams = 13;
template <int REGION> AVS_HD BekNodeParams bekker_node_params(const BekkerCtx& k) {
    static_assert(REGION == 0 || REGION == 2, "the centre arc stays on the scalar walk");
    BekNodeParams q;
    q.R = k.R; q.n = k.n; q.om = 1.0 - k.slip_ratio; q.sin_tf = k.sin_tf;
    q.RomT = k.R * q.om * k.tan_sa; q.invK = 1.0 / k.K; q.c_soil = k.c; q.tan_phi = k.tan_phi;
    if (REGION == 0) { q.cos_end = k.cos_tf; q.coef = k.kk; q.A = -k.theta_f; q.B = 0.0; q.C = k.theta_f; }
    else { q.cos_end = k.cos_tr; q.coef = k.ku; q.A = 2 * k.theta_c - k.theta_f; q.B = 2 * k.sin_tc; q.C = k.theta_f - 2 * k.theta_c + 2 * k.sin_tc; }
    return q;
}

// integrand values at W nodes with known th, sin th, cos th
template <int W>
AVS_HD void bekker_nodes_vec(const BekNodeParams& k, const double* th, const double* s, const double* c, double* fx, double* fy, double* fz) {
    double sg[W], jx[W], jy[W], q[W], r[W], e[W];
#pragma unroll
    for (int i = 0; i < W; i++) sg[i] = k.R * (c[i] - k.cos_end);
    log_vec<W>(sg);
#pragma unroll
    for (int i = 0; i < W; i++) sg[i] *= k.n;
    exp_vec<W>(sg);
#pragma unroll
    for (int i = 0; i < W; i++) sg[i] *= k.coef;
#pragma unroll
    for (int i = 0; i < W; i++) {
        jx[i] = k.R * (((th[i] + k.A) - k.om * (s[i] - k.sin_tf)) - k.B);
        jy[i] = k.RomT * (k.C - th[i]);
    }
#pragma unroll
    for (int i = 0; i < W; i++) q[i] = fma(jy[i], jy[i], jx[i] * jx[i]);
    // r = 1/sqrt(q): seed + two Newton steps
#pragma unroll
    for (int i = 0; i < W; i++) e[i] = q[i] > 1e-290 ? q[i] : 1e-290;
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = rsqrt_seed(e[i]);
#pragma unroll
    for (int i = 0; i < W; i++) { const double h = 0.5 * r[i]; r[i] = fma(h, fma(-e[i] * r[i], r[i], 1.0), r[i]); }
#pragma unroll
    for (int i = 0; i < W; i++) { const double h = 0.5 * r[i]; r[i] = fma(h, fma(-e[i] * r[i], r[i], 1.0), r[i]); }
    // j = q r ; j == 0 -> 1e-6 ("smart divide" :83-84), 1/j = r
#pragma unroll
    for (int i = 0; i < W; i++) {
        const bool zero = (q[i] == 0.0);
        e[i] = zero ? -1e-6 * k.invK : -(q[i] * r[i]) * k.invK;
        r[i] = zero ? 1e6 : r[i];
    }
    exp_vec<W>(e);
#pragma unroll
    for (int i = 0; i < W; i++) {
        const double shear = (k.c_soil + sg[i] * k.tan_phi) * (1.0 - e[i]);
        const double tau_x = (-jx[i] * r[i]) * shear;                                           // :80-104
        const double tau_y = (-jy[i] * r[i]) * shear;                                           // :106-129
        fx[i] = tau_x * c[i] - sg[i] * s[i];                                                    // Fx_eqn1/2 :146-154
        fy[i] = tau_y;
        fz[i] = sg[i] * c[i] + tau_x * s[i];                                                    // Fz_eqn1/2 :137-144
    }
}

// How the 12 nodes of an arc are evaluated: inline, four at a time (host harness, single-ODE kernels).  The rollout kernel
// substitutes a policy that runs the same routine as ONE non-inlined function called three times per arc, so that the hot
// quadrature code is ~350 instructions that stay in the instruction cache instead of six unrolled copies (avs_dev.cuh).
// Branch-margin probe (same kinds and definitions as oracle/jackal_oracle.hpp): how far each value-dependent
// predicate of the Bekker path was from flipping.  The node-evaluation policy doubles as the probe sink; the production
// policies ignore it (the calls compile away), the diagnostic instantiation keeps per-lane minima.
//   0 |sinkage| (contact on/off)   1 |w R - vx| (sign of Fx)   2 slip clamps at 1e-3   3 |pgc - pg| / pg
//   4 |R - sinkage| (zr clamp)     5 |dtheta| of the non-empty arcs (trapezoid loop bound)
constexpr int kNumMargins = 6;

struct BekNodesInline {
    AVS_HD void margin(int, double) const {}
    AVS_HD void operator()(const BekNodeParams& q, const double* th, const double* s, const double* c, double* fx, double* fy, double* fz) const {
        bekker_nodes_vec<4>(q, th, s, c, fx, fy, fz);
        bekker_nodes_vec<4>(q, th + 4, s + 4, c + 4, fx + 4, fy + 4, fz + 4);
        bekker_nodes_vec<4>(q, th + 8, s + 8, c + 8, fx + 8, fy + 8, fz + 8);
    }
};

// BekkerTireModel::integrate (:163-179) of the front / rear arc with the 12 node evaluations vectorised.
// (sin, cos) of the two bounds are passed in; degenerate arcs take the scalar walk.
template <int REGION, class Eval>
AVS_HD Bek3 bekker_integrate_vec(const BekkerCtx& k, double upper_b, double lower_b, double sin_lo, double cos_lo, double sin_up, double cos_up, const Eval& eval) {
    const double dtheta = (upper_b - lower_b) / 10.0;
    if (!(dtheta > 1e-9)) return bekker_integrate<REGION>(k, upper_b, lower_b);
    double th[12], s[12], c[12], fx[12], fy[12], fz[12];
    th[0] = lower_b;
#pragma unroll
    for (int i = 1; i < 10; i++) th[i] = th[i - 1] + dtheta;  // the reference's running theta += dtheta
    th[10] = upper_b - dtheta;
    th[11] = upper_b;
    double sd, cd;
#if defined(__CUDA_ARCH__)
    sincos(dtheta, &sd, &cd);
#else
    sd = sin(dtheta); cd = cos(dtheta);
#endif
    s[0] = sin_lo; c[0] = cos_lo;
#pragma unroll
    for (int i = 1; i < 10; i++) {
        s[i] = fma(s[i - 1], cd, c[i - 1] * sd);
        c[i] = fma(c[i - 1], cd, -s[i - 1] * sd);
    }
    s[11] = sin_up; c[11] = cos_up;
    s[10] = fma(sin_up, cd, -cos_up * sd);
    c[10] = fma(cos_up, cd, sin_up * sd);
    eval(bekker_node_params<REGION>(k), th, s, c, fx, fy, fz);
    Bek3 sum = {0.0, 0.0, 0.0};
#pragma unroll
    for (int it = 0; it < 9; it++) {
        sum.fx += .5 * dtheta * (fx[it + 1] + fx[it]);
        sum.fy += .5 * dtheta * (fy[it + 1] + fy[it]);
        sum.fz += .5 * dtheta * (fz[it + 1] + fz[it]);
    }
    sum.fx += .5 * dtheta * (fx[11] + fx[10]);
    sum.fy += .5 * dtheta * (fy[11] + fy[10]);
    sum.fz += .5 * dtheta * (fz[11] + fz[10]);
    return sum;
}

// One wheel: (contact vx, vy, wheel speed w, sinkage) -> (Fx, Fy, Fz) before the damping term
template <int TERR, class Eval = BekNodesInline>
AVS_HD void bekker_tire_fwd(const ModelConst& mc, double cvx, double cvy, double w, double sinkage, const double cp[3], double F[3], const Eval& eval = Eval()) {
    // soil parameters: global 5-vector (BekkerDynamics.cpp:12-21, 59-63) or per-contact-point grids (extension)
    double p0 = mc.bek[0], p1 = mc.bek[1], p2 = mc.bek[2], p3 = mc.bek[3], p4 = mc.bek[4];
    const int terr_kind = (TERR == TERR_DYN) ? mc.terr_kind : TERR;
    if (terr_kind == TERR_GRID && mc.grid.soil != nullptr) {
        const GridCellD cell = grid_cell(mc.grid, cp[0], cp[1]);
        const size_t plane = (size_t)mc.grid.nx * mc.grid.ny;
        p0 = grid_lerp(mc.grid, mc.grid.soil + 0 * plane, cell, nullptr, nullptr);
        p1 = grid_lerp(mc.grid, mc.grid.soil + 1 * plane, cell, nullptr, nullptr);
        p2 = grid_lerp(mc.grid, mc.grid.soil + 2 * plane, cell, nullptr, nullptr);
        p3 = grid_lerp(mc.grid, mc.grid.soil + 3 * plane, cell, nullptr, nullptr);
        p4 = grid_lerp(mc.grid, mc.grid.soil + 4 * plane, cell, nullptr, nullptr);
    }
    p3 = p3 > 0.0 ? p3 : 0.0;  // CondExpGt(m_params[3], 0, m_params[3], 0)
    const double kc = 100.0 * p0, kphi = 1000.0 * p1, n0 = 1.0 * p2, n1 = 1.0 * p3, phi = 1.0 * p4;

    // get_features (:185-209)
    const double vel_x_tan = .098 * w;
    const double vx_c = fabs(cvx) > 1e-3 ? cvx : 1e-3;
    const double wr_c = fabs(vel_x_tan) > 1e-3 ? vel_x_tan : 1e-3;
    const double slip_ratio = 1 - (vx_c / wr_c);
    const double slip_angle = tanh_f64(cvy / fabs(vx_c));

    eval.margin(0, fabs(sinkage));
    eval.margin(1, fabs(vel_x_tan - cvx));
    eval.margin(2, fabs(fabs(cvx) - 1e-3));
    eval.margin(2, fabs(fabs(vel_x_tan) - 1e-3));
    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
    if (sinkage > 0.0) {  // BekkerDynamics.cpp:81
        BekkerCtx k;
        k.R = 0.098; k.b = 0.05; k.K = .0254; k.c = 8.62; k.pg = 100;  // BekkerTireModel.cpp:5-14, .h:34-35
        const double k0 = 2195, Au = 192400;
        eval.margin(4, fabs(k.R - sinkage));
        const double zr = sinkage > k.R ? k.R : sinkage;  // :219
        k.slip_ratio = slip_ratio;
        k.tan_sa = tan(slip_angle);
        k.tan_phi = tan(phi);
        k.n = n0 + (n1 * fabs(slip_ratio));
        k.kk = (kc / k.b) + kphi;
        const double pgc = k.kk * pow_pos(zr, k.n);
        eval.margin(3, fabs(pgc - k.pg) / k.pg);
        if (pgc < k.pg) k.ze = zr;
        else k.ze = pow_pos(k.pg / k.kk, 1.0 / k.n);
        k.ku = k0 + (Au * k.ze);
        const double zu = pow_pos(k.kk / k.ku, 1.0 / k.n) * k.ze;
        // contact angles: theta = acos(1 - u) with u in [0, 1]; their cosines are 1 - u and their sines sqrt(u (2 - u))
        // (the reference calls cos / sin on the acos results; identical up to rounding, five libm calls fewer)
        const double uf = zr / k.R, ur = (zu + zr - k.ze) / k.R, uc = (zr - k.ze) / k.R;
        k.theta_f = acos(1 - uf);
        k.theta_r = -acos(1 - ur);
        k.theta_c = acos(1 - uc);
        k.cos_tf = 1 - uf; k.sin_tf = sqrt(uf * (2 - uf));
        k.cos_tr = 1 - ur;
        k.cos_tc = 1 - uc; k.sin_tc = sqrt(uc * (2 - uc));
        const double sin_tr = -sqrt(ur * (2 - ur));
        {   // arc widths as BekkerTireModel::integrate forms them (:165): front, centre, rear
            const double df = (k.theta_f - k.theta_c) / 10.0, dc = (k.theta_c - (-k.theta_c)) / 10.0, dr = ((-k.theta_c) - k.theta_r) / 10.0;
            if (df != 0.0) eval.margin(5, fabs(df));
            if (dc != 0.0) eval.margin(5, fabs(dc));
            if (dr != 0.0) eval.margin(5, fabs(dr));
        }
        const double bt_R = k.R * k.b;
        const Bek3 If = bekker_integrate_vec<0>(k, k.theta_f, k.theta_c, k.sin_tc, k.cos_tc, k.sin_tf, k.cos_tf, eval);
        const Bek3 Ir = bekker_integrate_vec<2>(k, -k.theta_c, k.theta_r, sin_tr, k.cos_tr, -k.sin_tc, k.cos_tc, eval);
        Bek3 Ic = {0.0, 0.0, 0.0};
        if (k.theta_c != 0.0) Ic = bekker_integrate<1>(k, k.theta_c, -k.theta_c);
        Fx = bt_R * If.fx + bt_R * k.cos_tc * Ic.fx + bt_R * Ir.fx;                // :245-247
        Fy = bt_R * If.fy + bt_R * k.cos_tc * Ic.fy + bt_R * Ir.fy;                // :249-251
        Fz = bt_R * If.fz + bt_R * Ir.fz + 2 * bt_R * k.sin_tc * k.pg;             // :253-255
        Fx *= 1000; Fy *= 1000; Fz *= 1000;                                        // :263-266
    }
    // sign hack (BekkerDynamics.cpp:93)
    F[0] = (vel_x_tan - cvx > 0.0) ? fabs(Fx) : -fabs(Fx);
    F[1] = Fy;
    F[2] = Fz;
}

}  // namespace avs
