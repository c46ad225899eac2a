// avs_bekker.cuh — Bekker/Wong terramechanics tire model, forward only.
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
template <int REGION> AVS_HD Bek3 bekker_integrate(const BekkerCtx& k, double upper_b, double lower_b) {
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

// One wheel: (contact vx, vy, wheel speed w, sinkage) -> (Fx, Fy, Fz) before the damping term
template <int TERR>
AVS_HD void bekker_tire_fwd(const ModelConst& mc, double cvx, double cvy, double w, double sinkage, const double cp[3], double F[3]) {
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
    const double slip_angle = tanh(cvy / fabs(vx_c));

    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
    if (sinkage > 0.0) {  // BekkerDynamics.cpp:81
        BekkerCtx k;
        k.R = 0.098; k.b = 0.05; k.K = .0254; k.c = 8.62; k.pg = 100;  // BekkerTireModel.cpp:5-14, .h:34-35
        const double k0 = 2195, Au = 192400;
        const double zr = sinkage > k.R ? k.R : sinkage;  // :219
        k.slip_ratio = slip_ratio;
        k.tan_sa = tan(slip_angle);
        k.tan_phi = tan(phi);
        k.n = n0 + (n1 * fabs(slip_ratio));
        k.kk = (kc / k.b) + kphi;
        const double pgc = k.kk * pow(zr, k.n);
        if (pgc < k.pg) k.ze = zr;
        else k.ze = pow(k.pg / k.kk, 1.0 / k.n);
        k.ku = k0 + (Au * k.ze);
        const double zu = pow(k.kk / k.ku, 1.0 / k.n) * k.ze;
        k.theta_f = acos(1 - (zr / k.R));
        k.theta_r = -acos(1 - ((zu + zr - k.ze) / k.R));
        k.theta_c = acos(1 - ((zr - k.ze) / k.R));
        k.cos_tf = cos(k.theta_f); k.sin_tf = sin(k.theta_f);
        k.cos_tr = cos(k.theta_r);
        k.cos_tc = cos(k.theta_c); k.sin_tc = sin(k.theta_c);
        const double bt_R = k.R * k.b;
        const Bek3 If = bekker_integrate<0>(k, k.theta_f, k.theta_c);
        const Bek3 Ir = bekker_integrate<2>(k, -k.theta_c, k.theta_r);
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
