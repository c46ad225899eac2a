// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity status: pinned to the reference's own first-party sources, compiled in place behind
// restated third-party headers (oracle/_ref, oracle/refshim/; tests/test_ref_pinning.py) — see oracle/README.md for what that
// does and does not claim.
// CPU restatement of the reference's batched-rollout hot path, templated on the scalar type
// (double / orc::Rev / orc::Dual / orc::Cnt, see scalar.hpp).  It follows the reference
// statement by statement, INCLUDING its dense "wasteful" order (6x6 products with constant
// zeros, LLT of a constant matrix every call, the discarded third ABA pass), so that it is the
// parity authority for the constant-folded CUDA kernels.  Every function cites the reference
// lines it restates (paths relative to /root/reference).
//
// Third-party arithmetic that is NOT vendored in the reference (and absent from this image):
//   iit-rbd (RobCoGen runtime; version unpinned, README.md:8): InertiaMatrix::fill,
//     Utils::buildInertiaTensor, vxIv, motionCrossProductMx, compute_Ia_revolute,
//     ctransform_Ia_revolute — restated below from the published Featherstone/iit-rbd definitions
//   Eigen3 (CMakeLists.txt:4,29): dense products, LLT solve
//   CppAD (src/types/Scalars.h:4-11): CondExp*, abs, pow, tanh, atan2, ADFun::Reverse
// The reference's tests pin no numeric golden vectors for this path; the loose
// physical known-answer tests that do exist are re-run on this oracle in tests/test_oracle.py, and every function below is
// compared with the reference's own code (compiled from /root/reference where it lies) in tests/test_ref_pinning.py.
#pragma once
#include <cstddef>
#include <vector>
#include "scalar.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------
// Constants: include/auvsl_dynamics/generated/model_constants.h:24-100
// ---------------------------------------------------------------------------------------------
constexpr double tire_radius = .098;
constexpr double tx_w[4] = {0.13099999725818634, 0.13099999725818634, -0.13099999725818634, -0.13099999725818634};
constexpr double ty_w[4] = {0.1877949982881546, -0.1877949982881546, 0.1877949982881546, -0.1877949982881546};
constexpr double tz_w[4] = {0.03449999913573265, 0.03449999913573265, 0.03449999913573265, 0.03449999913573265};
constexpr double m_base_link = 16.52400016784668;
constexpr double comx_base_link = 0.011999273672699928;
constexpr double comy_base_link = 0.;
constexpr double comz_base_link = 0.06699594855308533;
constexpr double ix_base_link = 0.38783785700798035;
constexpr double ixy_base_link = 0.0011965520679950714;
constexpr double ixz_base_link = -0.003115505911409855;
constexpr double iy_base_link = 0.46875107288360596;
constexpr double iyz_base_link = 0.0031140821520239115;
constexpr double iz_base_link = 0.4509454071521759;
constexpr double m_wheel_link = 0.47699999809265137;
constexpr double ix_wheel_link = 0.0013000000035390258;
constexpr double iy_wheel_link = 0.0013000000035390258;
constexpr double iyz_wheel_link = 4.808253448174149E-11;
constexpr double iz_wheel_link = 0.002400000113993883;

constexpr int STATE_DIM = 21;  // HybridDynamics.h:30, layout :65-69
constexpr int CNTRL_DIM = 2;
constexpr int NUM_NN_PARAMS = 416;  // TireNetwork.cpp:122-128
enum { AX = 0, AY, AZ, LX, LY, LZ };  // iit-rbd spatial index order (angular first)

template <class S> struct Vec3 { S v[3]; S& operator[](int i) { return v[i]; } const S& operator[](int i) const { return v[i]; } };
template <class S> struct Mat3 { S m[3][3]; };
template <class S> struct Vec6 { S v[6]; S& operator[](int i) { return v[i]; } const S& operator[](int i) const { return v[i]; } };
template <class S> struct Mat6 { S m[6][6]; };

template <class S> Vec3<S> mul(const Mat3<S>& A, const Vec3<S>& x) {
    Vec3<S> r;
    for (int i = 0; i < 3; i++) r[i] = A.m[i][0] * x[0] + A.m[i][1] * x[1] + A.m[i][2] * x[2];
    return r;
}
template <class S> Vec3<S> mulT(const Mat3<S>& A, const Vec3<S>& x) {
    Vec3<S> r;
    for (int i = 0; i < 3; i++) r[i] = A.m[0][i] * x[0] + A.m[1][i] * x[1] + A.m[2][i] * x[2];
    return r;
}
template <class S> Vec6<S> mul(const Mat6<S>& A, const Vec6<S>& x) {
    Vec6<S> r;
    for (int i = 0; i < 6; i++) {
        S s = A.m[i][0] * x[0];
        for (int j = 1; j < 6; j++) s = s + A.m[i][j] * x[j];
        r[i] = s;
    }
    return r;
}
template <class S> Vec6<S> mulT(const Mat6<S>& A, const Vec6<S>& x) {
    Vec6<S> r;
    for (int i = 0; i < 6; i++) {
        S s = A.m[0][i] * x[0];
        for (int j = 1; j < 6; j++) s = s + A.m[j][i] * x[j];
        r[i] = s;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// utils: src/auvsl_dynamics/utils.cpp
// ---------------------------------------------------------------------------------------------
// utils.cpp:45-52 (x,y,z,w quaternion -> body-to-world rotation)
template <class S> Mat3<S> toMatrixRotation(const S& x, const S& y, const S& z, const S& w) {
    Mat3<S> r;
    r.m[0][0] = 1 - 2 * y * y - 2 * z * z; r.m[0][1] = 2 * x * y - 2 * w * z;     r.m[0][2] = 2 * x * z + 2 * w * y;
    r.m[1][0] = 2 * x * y + 2 * w * z;     r.m[1][1] = 1 - 2 * x * x - 2 * z * z; r.m[1][2] = 2 * y * z - 2 * w * x;
    r.m[2][0] = 2 * x * z - 2 * w * y;     r.m[2][1] = 2 * y * z + 2 * w * x;     r.m[2][2] = 1 - 2 * x * x - 2 * y * y;
    return r;
}
// utils.cpp:20-28
template <class S> void calcQuatDot(const S q[4], const Vec3<S>& w, S qd[4]) {
    S m[4][3];
    m[0][0] = q[3];  m[0][1] = -q[2]; m[0][2] = q[1];
    m[1][0] = q[2];  m[1][1] = q[3];  m[1][2] = -q[0];
    m[2][0] = -q[1]; m[2][1] = q[0];  m[2][2] = q[3];
    m[3][0] = -q[0]; m[3][1] = -q[1]; m[3][2] = -q[2];
    for (int i = 0; i < 4; i++) qd[i] = .5 * m[i][0] * w[0] + .5 * m[i][1] * w[1] + .5 * m[i][2] * w[2];
}
// utils.cpp:56-78 — only yaw is consumed on the path (VehicleSystem.cpp:79-83)
template <class S> S eulerYaw(const S& qw, const S& qx, const S& qy, const S& qz) {
    S siny_cosp = 2 * (qw * qz + qx * qy);
    S cosy_cosp = 1 - 2 * (qy * qy + qz * qz);
    return s_atan2(siny_cosp, cosy_cosp);
}

// ---------------------------------------------------------------------------------------------
// Terrain: src/types/TerrainMap.h:6-11, src/TestTerrainMaps.cpp:6-28  (+ grid EXTENSION)
// ---------------------------------------------------------------------------------------------
struct TerrainSpec {
    int kind = 0;  // 0 Flat, 1 Slope, 2 Bumpy, 3 Grid (extension)
    int nx = 0, ny = 0;
    double x0 = 0, y0 = 0, dx = 1, dy = 1;
    const double* height = nullptr;  // [ny][nx]
    const double* soil = nullptr;    // [5][ny][nx] or null (Bekker soil grids; extension)
    int contact_search = 0;          // 1: sinkage from HybridDynamics::get_tire_cpts_sinkages_fancy (10-angle contact search)
};

// Bilinear cell lookup shared by the height and soil grids (extension; the CUDA kernel implements
// the identical clamp / floor / lerp sequence so parity for grids is oracle-vs-kernel only).
template <class S> struct GridCell { int i, j; S fu, fv; };
template <class S> GridCell<S> gridCell(const TerrainSpec& t, const S& x, const S& y) {
    S u = (x - t.x0) / t.dx, v = (y - t.y0) / t.dy;
    const double umax = (double)(t.nx - 1), vmax = (double)(t.ny - 1);
    u = cond_lt(u, S(0.0), S(0.0), u); u = cond_gt(u, S(umax), S(umax), u);
    v = cond_lt(v, S(0.0), S(0.0), v); v = cond_gt(v, S(vmax), S(vmax), v);
    int i = (int)std::floor(value(u)), j = (int)std::floor(value(v));
    if (i > t.nx - 2) i = t.nx - 2;
    if (j > t.ny - 2) j = t.ny - 2;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    GridCell<S> c; c.i = i; c.j = j; c.fu = u - (double)i; c.fv = v - (double)j;
    return c;
}
template <class S> S gridLerp(const TerrainSpec& t, const double* g, const GridCell<S>& c) {
    const double h00 = g[(size_t)c.j * t.nx + c.i], h01 = g[(size_t)c.j * t.nx + c.i + 1];
    const double h10 = g[(size_t)(c.j + 1) * t.nx + c.i], h11 = g[(size_t)(c.j + 1) * t.nx + c.i + 1];
    S a = h00 + c.fu * (h01 - h00);
    S b = h10 + c.fu * (h11 - h10);
    return a + c.fv * (b - a);
}

template <class S> S getAltitude(const TerrainSpec& t, const S& x, const S& y) {
    switch (t.kind) {
        case 0: return S(0.0);   // TestTerrainMaps.cpp:6-9
        case 1: return 0.1 * x;  // :12-15
        case 2: {                // :18-28
            S zero(0.0), mx(1.6);
            S arg = x - 2.0;
            arg = cond_lt(arg, mx, arg, zero);
            arg = cond_gt(arg, zero, arg, zero);
            S temp = .2 * s_sin(arg);
            return cond_gt(temp, zero, temp, zero);
        }
        default: return gridLerp(t, t.height, gridCell(t, x, y));
    }
}

// ---------------------------------------------------------------------------------------------
// TireNetwork: include/auvsl_dynamics/TireNetwork.h:14-58, src/auvsl_dynamics/TireNetwork.cpp
// ---------------------------------------------------------------------------------------------
// Pretrained weights, TireNetwork::load_model() (TireNetwork.cpp:198-226): numeric DATA.
static const double kW0[24] = {
    1.8523e+00, -2.4395e-01, -2.0577e+00, 5.7221e-01, 9.7006e-01, 3.7060e-01, -6.8117e-02, 1.5170e+00,
    6.9176e-01, -1.5698e-01, -1.1519e+00, 4.7048e-01, -1.0036e-02, 1.4627e-01, 1.8264e+00, -1.3942e+00,
    7.5852e-02, 5.7577e-03, 1.8568e+00, -3.1544e-01, 2.0023e+00, -1.6360e+01, -1.1013e+00, 2.8912e-02};
static const double kW2[64] = {
    -0.6545, 1.2153, 0.2281, -0.8833, -1.4218, -2.7572, 0.3027, -2.1109, -2.0394, -0.0442, 0.3384, 0.2103, -1.5120,
    1.2095, 0.1672, 0.1592, 0.5422, 0.3246, -0.0478, 0.3761, 0.7557, -1.2539, 0.8754, -0.4485, -1.3457, -1.3106,
    -0.0682, 1.4183, -0.9145, 4.1784, -0.4422, 1.5920, -0.4213, 0.2675, -0.1125, 0.0500, 1.3408, 1.1443, 0.3911,
    1.5622, -0.6181, -0.3097, 0.4885, 0.4184, 0.8413, 0.4065, 0.7734, 0.3717, 0.2096, -1.8048, -0.1093, -1.7614,
    -0.7847, 0.1165, -0.5894, 0.0641, 0.0843, 1.7857, -0.8906, -1.0495, 0.7428, -2.3685, -0.3676, -1.9639};
static const double kW4[16] = {0.3660, -0.2972, 0.5233, 1.0919, -0.2640, -0.1887, -0.1882, 0.6705,
                               -2.1059, -0.4484, -0.3344, -0.9591, 0.3450, -0.7008, 0.4277, 1.2480};
// Fixed z-branch weights, TireNetwork ctor (TireNetwork.cpp:17-32)
static const double kZ0[8] = {0.4197, 0.4528, 0.4402, -0.7671, -0.4358, 0.4605, 0.4606, 0.8535};
static const double kZ2[64] = {
    -0.7578, 0.2568, 0.2189, 0.2418, 0.0571, 0.3125, 0.1221, 0.3023, -0.7722, 0.0397, 0.3921, 0.1906, -0.2118,
    0.3593, -0.1007, 0.0398, 0.8704, -0.2301, 0.2404, -0.2393, -0.1154, 0.0570, -0.2363, 0.0252, -0.5188, 0.4544,
    -0.2427, 0.0496, 0.0680, 0.1878, 0.1537, 0.1105, 0.5801, -0.1279, 0.2275, -0.0404, -0.0301, -0.1021, -0.4477,
    -0.5217, 1.0052, -0.0991, -0.1396, -0.0447, 0.2628, -0.5366, -0.3278, 0.0602, -0.8997, 0.0576, 0.4188, 0.2293,
    -0.5440, 0.0735, -0.0501, 0.2984, -0.5156, 0.3454, 0.2395, 0.6368, -0.0076, 0.0691, -0.0122, 0.1513};
static const double kZ4[8] = {-26.3910, -26.2104, 26.6124, -26.3831, 26.4911, 26.6735, -26.4100, -26.7209};
static const double kOutStd[3] = {17.241097080572587, 15.9103406661298, 12.632573461434308};            // :11
static const double kInStd[4] = {0.0008373582968488336, 0.5799984931945801, 0.576934278011322, 0.5774632096290588};  // :12

// TireNetwork::load_model + getParams (TireNetwork.cpp:163-226): default 416-vector, 4 identical nets
inline void nnDefaultParams(double* p416) {
    int idx = 0;
    for (int kk = 0; kk < 4; kk++) {
        for (int i = 0; i < 24; i++) p416[idx++] = kW0[i];
        for (int i = 0; i < 64; i++) p416[idx++] = kW2[i];
        for (int i = 0; i < 16; i++) p416[idx++] = kW4[i];
    }
}

template <class S> struct TireNetwork {
    struct Params {
        S weight0[8][3], weight2[8][8], weight4[2][8];
        double z_weight0[8], z_weight2[8][8], z_weight4[8];
    };
    Params m_params[4];
    double in_std_inv[4];
    TireNetwork() {  // TireNetwork.cpp:9-34
        for (int i = 0; i < 4; i++) in_std_inv[i] = 1.0 / kInStd[i];
        for (int kk = 0; kk < 4; kk++) {
            for (int i = 0; i < 8; i++) {
                m_params[kk].z_weight0[i] = kZ0[i];
                m_params[kk].z_weight4[i] = kZ4[i];
                for (int j = 0; j < 8; j++) m_params[kk].z_weight2[i][j] = kZ2[i * 8 + j];
            }
        }
    }
    // TireNetwork.cpp:130-161 (pack order W0,W2,W4 row-major, net after net)
    void setParams(const S* params) {
        int idx = 0;
        for (int kk = 0; kk < 4; kk++) {
            for (int i = 0; i < 8; i++) for (int j = 0; j < 3; j++) m_params[kk].weight0[i][j] = params[idx++];
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) m_params[kk].weight2[i][j] = params[idx++];
            for (int i = 0; i < 2; i++) for (int j = 0; j < 8; j++) m_params[kk].weight4[i][j] = params[idx++];
        }
    }
    void getParams(S* params) const {  // :163-194
        int idx = 0;
        for (int kk = 0; kk < 4; kk++) {
            for (int i = 0; i < 8; i++) for (int j = 0; j < 3; j++) params[idx++] = m_params[kk].weight0[i][j];
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) params[idx++] = m_params[kk].weight2[i][j];
            for (int i = 0; i < 2; i++) for (int j = 0; j < 8; j++) params[idx++] = m_params[kk].weight4[i][j];
        }
    }
    static S relu(const S& x) { S zero(0.0); return cond_gt(x, zero, x, zero); }  // :47-50
    // TireNetwork.cpp:54-119.  in_vec = (vx, vy, w, zr, 0,0,0,0); out = (Fx, Fy, Fz, 0)
    void forward(const S in_vec[8], S out_vec[4], int ii) const {
        const Params& P = m_params[ii];
        S tire_tangent_vel = in_vec[2] * tire_radius;
        S diff = tire_tangent_vel - in_vec[0];
        S bekker_vec[4] = {in_vec[3], diff, in_vec[2], in_vec[1]};
        S scaled[4];
        for (int i = 0; i < 4; i++) scaled[i] = bekker_vec[i] * in_std_inv[i];
        S xy_features[3] = {scaled[1], scaled[2], scaled[3]};
        S xy0[8], xy2[8], xy4[2];
        for (int i = 0; i < 8; i++) {
            S s = P.weight0[i][0] * xy_features[0];
            for (int j = 1; j < 3; j++) s = s + P.weight0[i][j] * xy_features[j];
            xy0[i] = s_tanh(s);
        }
        for (int i = 0; i < 8; i++) {
            S s = P.weight2[i][0] * xy0[0];
            for (int j = 1; j < 8; j++) s = s + P.weight2[i][j] * xy0[j];
            xy2[i] = s_tanh(s);
        }
        for (int i = 0; i < 2; i++) {
            S s = P.weight4[i][0] * xy2[0];
            for (int j = 1; j < 8; j++) s = s + P.weight4[i][j] * xy2[j];
            xy4[i] = s;
        }
        S z0[8], z2[8], z4;
        for (int i = 0; i < 8; i++) z0[i] = s_tanh(P.z_weight0[i] * scaled[0]);
        for (int i = 0; i < 8; i++) {
            S s = P.z_weight2[i][0] * z0[0];
            for (int j = 1; j < 8; j++) s = s + P.z_weight2[i][j] * z0[j];
            z2[i] = s_tanh(s);
        }
        z4 = P.z_weight4[0] * z2[0];
        for (int j = 1; j < 8; j++) z4 = z4 + P.z_weight4[j] * z2[j];
        S forces[3] = {xy4[0], xy4[1], relu(z4)};
        for (int i = 0; i < 3; i++) out_vec[i] = forces[i] * kOutStd[i];
        out_vec[3] = S(0.0);
    }
};

// ---------------------------------------------------------------------------------------------
// Branch-margin probe (SURVEY.md §7 "parity hazards from discontinuities").  The Bekker path switches on values:
// a predicate that sits within rounding distance of its threshold may flip between two correct implementations and
// change the result by O(1e-3).  When a probe array is installed (thread-local), every evaluation records how far each
// predicate was from flipping; the minimum over a rollout is the trajectory's branch margin.  Kinds:
//   0 |sinkage|                       contact on/off                 BekkerDynamics.cpp:81
//   1 |w R - vx|                      sign of Fx                     BekkerDynamics.cpp:93
//   2 ||vx| - 1e-3|, ||w R| - 1e-3|   slip clamps                    BekkerTireModel.cpp:197-198
//   3 |pgc - pg| / pg                 ze = zr  vs  (pg/k)^(1/n)      BekkerTireModel.cpp:232
//   4 |R - sinkage|                   zr = min(zr, R)                BekkerTireModel.cpp:219
//   5 |dtheta| of non-empty arcs      trapezoid loop bound           BekkerTireModel.cpp:166-170
// ---------------------------------------------------------------------------------------------
constexpr int NUM_MARGINS = 6;
inline double*& active_margin() { static thread_local double* p = nullptr; return p; }
inline void margin_note(int kind, double v) {
    double* m = active_margin();
    if (m && v < m[kind]) m[kind] = v;
}

// ---------------------------------------------------------------------------------------------
// BekkerTireModel: include/auvsl_dynamics/BekkerTireModel.h, src/auvsl_dynamics/BekkerTireModel.cpp
// ---------------------------------------------------------------------------------------------
template <class S> struct BekkerTireModel {
    // BekkerTireModel.cpp:5-14, BekkerTireModel.h:34-35,44
    double k0 = 2195, Au = 192400, K = .0254, density = 1681, c = 8.62, pg = 100;
    double R = 0.098, b = 0.05;
    int num_steps = 10;
    S zr, slip_ratio, slip_angle, kc, kphi, n0, n1, phi, n, pgc, ze, zu, ku, theta_f, theta_r, theta_c;

    typedef S (BekkerTireModel::*Fn)(S);
    S sigma_x_cf(S theta) { S diff = s_cos(theta) - s_cos(theta_f); return ((kc / b) + kphi) * s_pow(R * diff, n); }  // :18-24
    S sigma_x_cc(S) { return ((kc / b) + kphi) * s_pow(ze, n); }                                                      // :26-29
    S sigma_x_cr(S theta) { S diff = s_cos(theta) - s_cos(theta_r); return ku * s_pow(R * (diff), n); }               // :31-35
    S jx_cf(S theta) { return R * ((theta - theta_f) - (1 - slip_ratio) * (s_sin(theta) - s_sin(theta_f))); }         // :37-40
    S jx_cc(S theta) { return R * ((theta_c - theta_f) + (1 - slip_ratio) * s_sin(theta_f) - s_sin(theta_c) + (slip_ratio * s_cos(theta_c) * s_tan(theta))); }  // :42-45
    S jx_cr(S theta) { return R * ((2 * theta_c + theta - theta_f) - (1 - slip_ratio) * (s_sin(theta) - s_sin(theta_f)) - 2 * s_sin(theta_c)); }                // :46-49
    S jy_cf(S theta) { return R * (1 - slip_ratio) * s_tan(slip_angle) * (theta_f - theta); }                                                                   // :51-54
    S jy_cc(S theta) { return R * (1 - slip_ratio) * s_tan(slip_angle) * (theta_f - theta + s_sin(theta_c) - s_cos(theta_c) * s_tan(theta)); }                  // :55-58
    S jy_cr(S theta) { return R * (1 - slip_ratio) * s_tan(slip_angle) * (theta_f - theta - 2 * theta_c + 2 * s_sin(theta_c)); }                                // :59-62
    static S sq(const S& x) { return s_pow(x, S(2.0)); }  // CppAD::pow(x, 2)
    S j_cf(S theta) { return s_sqrt(sq(jy_cf(theta)) + sq(jx_cf(theta))); }  // :64-68
    S j_cc(S theta) { return s_sqrt(sq(jy_cc(theta)) + sq(jx_cc(theta))); }  // :70-73
    S j_cr(S theta) { return s_sqrt(sq(jy_cr(theta)) + sq(jx_cr(theta))); }  // :75-78
    static S nz(const S& j) { return cond_eq(j, S(0.0), S(1e-6), j); }       // "smart divide" :83-84
    S tau_x_cf(S theta) { S j = nz(j_cf(theta)); return (-jx_cf(theta) / j) * (c + (sigma_x_cf(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }            // :80-86
    S tau_x_cc(S theta) { S j = nz(j_cc(theta)); S t = jx_cc(theta) / j; return (-t) * (c + (sigma_x_cc(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }   // :88-95
    S tau_x_cr(S theta) { S j = nz(j_cr(theta)); S t = jx_cr(theta) / j; return (-t) * (c + (sigma_x_cr(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }   // :97-104
    S tau_y_cf(S theta) { S j = nz(j_cf(theta)); S t = jy_cf(theta) / j; return (-t) * (c + (sigma_x_cf(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }   // :106-113
    S tau_y_cc(S theta) { S j = nz(j_cc(theta)); S t = jy_cc(theta) / j; return (-t) * (c + (sigma_x_cc(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }   // :114-121
    S tau_y_cr(S theta) { S j = nz(j_cr(theta)); S t = jy_cr(theta) / j; return (-t) * (c + (sigma_x_cr(theta) * s_tan(phi))) * (1 - (s_exp(-j / K))); }   // :122-129
    S Fy_eqn3(S theta) { return tau_y_cc(theta) * sq(1.0 / s_cos(theta)); }                                   // :132-135
    S Fz_eqn1(S theta) { return (sigma_x_cf(theta) * s_cos(theta)) + (tau_x_cf(theta) * s_sin(theta)); }      // :137-140
    S Fz_eqn2(S theta) { return (sigma_x_cr(theta) * s_cos(theta)) + (tau_x_cr(theta) * s_sin(theta)); }      // :141-144
    S Fx_eqn1(S theta) { return (tau_x_cf(theta) * s_cos(theta)) - (sigma_x_cf(theta) * s_sin(theta)); }      // :146-149
    S Fx_eqn2(S theta) { return (tau_x_cr(theta) * s_cos(theta)) - (sigma_x_cr(theta) * s_sin(theta)); }      // :151-154
    S Fx_eqn3(S theta) { return tau_x_cc(theta) * (sq(1.0 / s_cos(theta))); }                                 // :155-158
    S Ty_eqn1(S theta) { return tau_x_cc(theta) * (sq(1.0 / s_cos(theta))); }                                 // :159-162

    // :163-179 trapezoid; the value-dependent loop bound is kept
    S integrate(Fn func, S upper_b, S lower_b) {
        S dtheta = (upper_b - lower_b) / (double)num_steps;
        S eps = dtheta * .1;
        if (value(dtheta) != 0.0) margin_note(5, std::fabs(value(dtheta)));
        S sum(0.0);
        for (S theta = lower_b; theta < (upper_b - eps - dtheta); theta += dtheta) {
            sum += .5 * dtheta * ((this->*func)(theta + dtheta) + (this->*func)(theta));
        }
        sum += .5 * dtheta * ((this->*func)(upper_b) + (this->*func)(upper_b - dtheta));
        return sum;
    }
    // :185-209   inputs = (zr, vx, vy, w*R)
    void get_features(const S inputs[4], S out[3]) {
        S tire_vx = inputs[1], tire_vy = inputs[2], w_R = inputs[3];
        const S small(1e-3);
        margin_note(2, std::fabs(std::fabs(value(tire_vx)) - 1e-3));
        margin_note(2, std::fabs(std::fabs(value(w_R)) - 1e-3));
        S tire_vx_clamped = cond_gt(s_abs(tire_vx), small, tire_vx, small);
        S tangent_vx_clamped = cond_gt(s_abs(w_R), small, w_R, small);
        S sr = 1 - (tire_vx_clamped / tangent_vx_clamped);
        S sa = s_tanh(tire_vy / s_abs(tire_vx_clamped));
        out[0] = inputs[0]; out[1] = sr; out[2] = sa;
    }
    // :212-278
    void get_forces(const S features[8], S wrench[4]) {
        margin_note(4, std::fabs(R - value(features[0])));
        zr = cond_gt(features[0], S(R), S(R), features[0]);
        slip_ratio = features[1];
        slip_angle = features[2];
        kc = features[3]; kphi = features[4]; n0 = features[5]; n1 = features[6]; phi = features[7];
        n = n0 + (n1 * s_abs(slip_ratio));
        pgc = ((kc / b) + kphi) * s_pow(zr, n);
        margin_note(3, std::fabs(value(pgc) - pg) / pg);
        if (value(pgc) < pg) {
            ze = zr;
        } else {
            ze = s_pow(pg / ((kc / b) + kphi), 1.0 / n);
        }
        ku = k0 + (Au * ze);
        zu = s_pow(((kc / b) + kphi) / ku, 1.0 / n) * ze;
        theta_f = s_acos(1 - (zr / R));
        theta_r = -s_acos(1 - ((zu + zr - ze) / R));
        theta_c = s_acos(1 - ((zr - ze) / R));
        const double bt_R = R * b, bt_RR = R * R * b;
        typedef BekkerTireModel M;
        S Fx = bt_R * integrate(&M::Fx_eqn1, theta_f, theta_c) + bt_R * s_cos(theta_c) * integrate(&M::Fx_eqn3, theta_c, -theta_c) +
               bt_R * integrate(&M::Fx_eqn2, -theta_c, theta_r);
        S Fy = bt_R * integrate(&M::tau_y_cf, theta_f, theta_c) + bt_R * s_cos(theta_c) * integrate(&M::Fy_eqn3, theta_c, -theta_c) +
               bt_R * integrate(&M::tau_y_cr, -theta_c, theta_r);
        S Fz = bt_R * integrate(&M::Fz_eqn1, theta_f, theta_c) + bt_R * integrate(&M::Fz_eqn2, -theta_c, theta_r) +
               2 * bt_R * s_sin(theta_c) * pg;
        S Ty = -bt_RR * integrate(&M::tau_x_cf, theta_f, theta_c) + -bt_RR * integrate(&M::tau_x_cr, -theta_c, theta_r) +
               -bt_RR * s_cos(theta_c) * s_cos(theta_c) * integrate(&M::Ty_eqn1, theta_c, -theta_c);
        (void)Ty;  // discarded by the reference (:275)
        wrench[0] = 1000 * Fx; wrench[1] = 1000 * Fy; wrench[2] = 1000 * Fz; wrench[3] = S(0.0);
    }
};

// ---------------------------------------------------------------------------------------------
// RobCoGen-generated rigid-body model (src/auvsl_dynamics/generated/*), at q = 0
// ---------------------------------------------------------------------------------------------
// iit::rbd::Utils::buildInertiaTensor (off-diagonals negated) + InertiaMatrix::fill(m, com, I):
//   [[ I , m c^ ], [ -m c^ , m 1 ]]      (inertia_properties.cpp:6-38)
inline void fillInertia(double M[6][6], double m, const double c[3], double ixx, double iyy, double izz, double ixy, double ixz, double iyz) {
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) M[i][j] = 0.0;
    M[0][0] = ixx;  M[0][1] = -ixy; M[0][2] = -ixz;
    M[1][0] = -ixy; M[1][1] = iyy;  M[1][2] = -iyz;
    M[2][0] = -ixz; M[2][1] = -iyz; M[2][2] = izz;
    const double cx[3][3] = {{0, -c[2], c[1]}, {c[2], 0, -c[0]}, {-c[1], c[0], 0}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        M[i][3 + j] = m * cx[i][j];
        M[3 + i][j] = -m * cx[i][j];
    }
    for (int i = 0; i < 3; i++) M[3 + i][3 + i] = m;
}
// MotionTransforms::fr_<wheel>_link_X_fr_base_link at q=0 (transforms.cpp:194-239, 286-331, 378-423, 470-515)
inline void wheelMotionTransform(double X[6][6], int w) {
    const double tx = tx_w[w], ty = ty_w[w], tz = tz_w[w];
    const double s = 0.0, c = 1.0;  // sin(q), cos(q) with q = 0 (HybridDynamics.cpp:520)
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) X[i][j] = 0.0;
    X[0][0] = c;  X[0][2] = -s;
    X[1][0] = -s; X[1][2] = -c;
    X[2][1] = 1.0;
    X[3][0] = -ty * s; X[3][1] = (tx * s) + (tz * c); X[3][2] = -ty * c; X[3][3] = c;  X[3][5] = -s;
    X[4][0] = -ty * c; X[4][1] = (tx * c) - (tz * s); X[4][2] = ty * s;  X[4][3] = -s; X[4][5] = -c;
    X[5][0] = -tz;     X[5][2] = tx;                  X[5][4] = 1.0;
}
// getWholeBodyCOM (generated/miscellaneous.cpp:6-39); homogeneous translations transforms.cpp:3558-3572
inline void wholeBodyCOM(double com[3]) {
    double sum[3] = {comx_base_link * m_base_link, comy_base_link * m_base_link, comz_base_link * m_base_link};
    double M = m_base_link;
    for (int w = 0; w < 4; w++) {
        sum[0] += m_wheel_link * tx_w[w]; sum[1] += m_wheel_link * ty_w[w]; sum[2] += m_wheel_link * tz_w[w];
        M += m_wheel_link;
    }
    for (int i = 0; i < 3; i++) com[i] = sum[i] / M;
}

// iit::rbd::vxIv(v, I) = v x* (I v)
template <class S> Vec6<S> vxIv(const Vec6<S>& v, const double I[6][6]) {
    Vec6<S> Iv;
    for (int i = 0; i < 6; i++) {
        S s = I[i][0] * v[0];
        for (int j = 1; j < 6; j++) s = s + I[i][j] * v[j];
        Iv[i] = s;
    }
    Vec6<S> r;
    // [ w x n + v x f ; w x f ]
    r[0] = v[1] * Iv[2] - v[2] * Iv[1] + v[4] * Iv[5] - v[5] * Iv[4];
    r[1] = v[2] * Iv[0] - v[0] * Iv[2] + v[5] * Iv[3] - v[3] * Iv[5];
    r[2] = v[0] * Iv[1] - v[1] * Iv[0] + v[3] * Iv[4] - v[4] * Iv[3];
    r[3] = v[1] * Iv[5] - v[2] * Iv[4];
    r[4] = v[2] * Iv[3] - v[0] * Iv[5];
    r[5] = v[0] * Iv[4] - v[1] * Iv[3];
    return r;
}

// Jackal::rcg::ForwardDynamics::fd (generated/forward_dynamics.cpp:30-178; q=0 via forward_dynamics.h:122-139)
template <class S> struct ForwardDynamics {
    double I_base[6][6], I_wheel[6][6], X[4][6][6];
    ForwardDynamics() {
        const double cb[3] = {comx_base_link, comy_base_link, comz_base_link}, c0[3] = {0, 0, 0};
        fillInertia(I_base, m_base_link, cb, ix_base_link, iy_base_link, iz_base_link, ixy_base_link, ixz_base_link, iyz_base_link);
        fillInertia(I_wheel, m_wheel_link, c0, ix_wheel_link, iy_wheel_link, iz_wheel_link, 0.0, 0.0, iyz_wheel_link);
        for (int w = 0; w < 4; w++) wheelMotionTransform(X[w], w);
    }
    // fext[0] = base link, fext[1..4] = FL,FR,RL,RR wheel links, each in its own link frame
    void fd(S qdd[4], Vec6<S>& base_a, const Vec6<S>& base_v, const Vec6<S>& g, const S qd[4], const S tau[4], const Vec6<S> fext[5]) const {
        S AI[6][6];
        Vec6<S> base_p, wp[4], wv[4], wc[4];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) AI[i][j] = S(I_base[i][j]);  // :40
        for (int i = 0; i < 6; i++) base_p[i] = -fext[0][i];                                 // :41
        for (int w = 0; w < 4; w++) for (int i = 0; i < 6; i++) wp[w][i] = -fext[w + 1][i];  // :43-49
        // FIRST PASS :56-102
        for (int w = 0; w < 4; w++) {
            for (int i = 0; i < 6; i++) {
                S s = X[w][i][0] * base_v[0];
                for (int j = 1; j < 6; j++) s = s + X[w][i][j] * base_v[j];
                wv[w][i] = s;
            }
            wv[w][AZ] += qd[w];
            // motionCrossProductMx(v): [[w^,0],[v^,w^]]; column AZ = (wy, -wx, 0, vy, -vx, 0)
            const Vec6<S>& v = wv[w];
            wc[w][0] = v[AY] * qd[w];  wc[w][1] = (-v[AX]) * qd[w]; wc[w][2] = S(0.0) * qd[w];
            wc[w][3] = v[LY] * qd[w];  wc[w][4] = (-v[LX]) * qd[w]; wc[w][5] = S(0.0) * qd[w];
            Vec6<S> b = vxIv(v, I_wheel);
            for (int i = 0; i < 6; i++) wp[w][i] += b[i];
        }
        {
            Vec6<S> b = vxIv(base_v, I_base);  // :105
            for (int i = 0; i < 6; i++) base_p[i] += b[i];
        }
        // SECOND PASS :107-153, order RR, RL, FR, FL
        S wu[4], wD[4];
        Vec6<S> wU[4];
        for (int w = 3; w >= 0; w--) {
            wu[w] = tau[w] - wp[w][AZ];
            for (int i = 0; i < 6; i++) wU[w][i] = S(I_wheel[i][AZ]);
            wD[w] = wU[w][AZ];
            // compute_Ia_revolute: Ia = IA - U U^T / D, row/column AZ left zero
            S Ia[6][6];
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
                if (i == AZ || j == AZ) Ia[i][j] = S(0.0);
                else Ia[i][j] = I_wheel[i][j] - wU[w][i] * wU[w][j] / wD[w];
            }
            Vec6<S> pa;
            for (int i = 0; i < 6; i++) {
                S s = Ia[i][0] * wc[w][0];
                for (int j = 1; j < 6; j++) s = s + Ia[i][j] * wc[w][j];
                pa[i] = wp[w][i] + s + wU[w][i] * wu[w] / wD[w];
            }
            // ctransform_Ia_revolute: IaB = X^T Ia X ; base_AI += IaB ; base_p += X^T pa
            S T[6][6];
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
                S s = Ia[i][0] * X[w][0][j];
                for (int k = 1; k < 6; k++) s = s + Ia[i][k] * X[w][k][j];
                T[i][j] = s;
            }
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
                S s = X[w][0][i] * T[0][j];
                for (int k = 1; k < 6; k++) s = s + X[w][k][i] * T[k][j];
                AI[i][j] += s;
            }
            for (int i = 0; i < 6; i++) {
                S s = X[w][0][i] * pa[0];
                for (int k = 1; k < 6; k++) s = s + X[w][k][i] * pa[k];
                base_p[i] += s;
            }
        }
        // :156  base_a = -AI.llt().solve(base_p)
        S L[6][6];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) L[i][j] = S(0.0);
        for (int j = 0; j < 6; j++) {
            S d = AI[j][j];
            for (int k = 0; k < j; k++) d = d - L[j][k] * L[j][k];
            L[j][j] = s_sqrt(d);
            for (int i = j + 1; i < 6; i++) {
                S s = AI[i][j];
                for (int k = 0; k < j; k++) s = s - L[i][k] * L[j][k];
                L[i][j] = s / L[j][j];
            }
        }
        S y[6], x[6];
        for (int i = 0; i < 6; i++) {
            S s = base_p[i];
            for (int k = 0; k < i; k++) s = s - L[i][k] * y[k];
            y[i] = s / L[i][i];
        }
        for (int i = 5; i >= 0; i--) {
            S s = y[i];
            for (int k = i + 1; k < 6; k++) s = s - L[k][i] * x[k];
            x[i] = s / L[i][i];
        }
        for (int i = 0; i < 6; i++) base_a[i] = -x[i];
        // THIRD PASS :158-173 (results discarded by the caller, HybridDynamics.cpp:627-630)
        for (int w = 0; w < 4; w++) {
            Vec6<S> a;
            for (int i = 0; i < 6; i++) {
                S s = X[w][i][0] * base_a[0];
                for (int j = 1; j < 6; j++) s = s + X[w][i][j] * base_a[j];
                a[i] = s + wc[w][i];
            }
            S dot = wU[w][0] * a[0];
            for (int j = 1; j < 6; j++) dot = dot + wU[w][j] * a[j];
            qdd[w] = (wu[w] - dot) / wD[w];
        }
        for (int i = 0; i < 6; i++) base_a[i] += g[i];  // :177
    }
};

// ---------------------------------------------------------------------------------------------
// HybridDynamics / BekkerDynamics
// ---------------------------------------------------------------------------------------------
enum TireModel { TIRE_NN = 0, TIRE_BEKKER = 1 };

template <class S> struct HybridDynamics {
    static constexpr double timestep = 1e-3;  // HybridDynamics.cpp:33
    TerrainSpec terrain;
    int tire_model = TIRE_NN;
    TireNetwork<S> tire_network;
    BekkerTireModel<S> bekker_tire_model;
    S bekker_params[5];  // BekkerDynamics::m_params
    ForwardDynamics<S> fwd_dynamics;
    S state_[STATE_DIM];

    // BekkerDynamics::setParams (BekkerDynamics.cpp:12-21)
    void setBekkerParams(const S* p) {
        for (int i = 0; i < 5; i++) bekker_params[i] = p[i];
        S zero(0.0);
        bekker_params[3] = cond_gt(bekker_params[3], zero, bekker_params[3], zero);
    }
    void initState() {  // HybridDynamics.cpp:62-65
        const double s0[STATE_DIM] = {0, 0, 0, 1, 0, 0, 0.16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < STATE_DIM; i++) state_[i] = S(s0[i]);
    }
    // HybridDynamics.cpp:111-133
    static void initStateCOM(const S* start_state, S* base_state) {
        double com[3];
        wholeBodyCOM(com);
        const double r_ss[3][3] = {{0., -com[2], com[1]}, {com[2], 0., -com[0]}, {-com[1], com[0], 0.}};
        S com_lin_vel[3] = {start_state[14], start_state[15], start_state[16]};
        S ang_vel[3] = {start_state[11], start_state[12], start_state[13]};
        S base_lin_vel[3];
        for (int i = 0; i < 3; i++)
            base_lin_vel[i] = com_lin_vel[i] + (r_ss[i][0] * ang_vel[0] + r_ss[i][1] * ang_vel[1] + r_ss[i][2] * ang_vel[2]);
        for (int i = 0; i < STATE_DIM; i++) base_state[i] = start_state[i];
        base_state[14] = base_lin_vel[0]; base_state[15] = base_lin_vel[1]; base_state[16] = base_lin_vel[2];
    }
    // HybridDynamics.cpp:214-240
    void get_tire_cpts(const S* X, Vec3<S> cpt_pts[4], Mat3<S>& rot) const {
        rot = toMatrixRotation(X[0], X[1], X[2], X[3]);
        for (int ii = 0; ii < 4; ii++) {
            Vec3<S> e; e[0] = S(tx_w[ii] + 0.0); e[1] = S(ty_w[ii] + 0.0); e[2] = S(tz_w[ii] + (-tire_radius));
            Vec3<S> r = mul(rot, e);
            cpt_pts[ii][0] = r[0] + X[4]; cpt_pts[ii][1] = r[1] + X[5]; cpt_pts[ii][2] = r[2] + X[6];
        }
    }
    // :242-253
    void get_tire_sinkages(const Vec3<S> cpt_pts[4], S sinkages[4]) const {
        for (int i = 0; i < 4; i++) {
            const S altitude = getAltitude(terrain, cpt_pts[i][0], cpt_pts[i][1]);
            sinkages[i] = altitude - cpt_pts[i][2];
        }
    }
    // HybridDynamics::get_tire_cpts_sinkages_fancy (HybridDynamics.cpp:258-316; utils.cpp:5-17 roty): for every wheel, ten
    // candidate contact points on the rim, at pitch angles -0.3 ... +0.3 rad about the joint axis; the sinkage is the largest
    // candidate sinkage (0 if none is positive) and the reported point is the JOINT centre.  (Unused by the reference's own
    // ODE; offered here as an option for rough height fields.)
    void get_tire_cpts_sinkages_fancy(const S* X, Vec3<S> cpt_pts[4], S sinkages[4]) const {
        Mat3<S> base_rot = toMatrixRotation(X[0], X[1], X[2], X[3]);
        const int max_checks = 10;
        const S max_angle(-.3);
        for (int ii = 0; ii < 4; ii++) {
            Vec3<S> e; e[0] = S(tx_w[ii]); e[1] = S(ty_w[ii]); e[2] = S(tz_w[ii]);
            Vec3<S> r = mul(base_rot, e);
            Vec3<S> joint;
            joint[0] = r[0] + X[4]; joint[1] = r[1] + X[5]; joint[2] = r[2] + X[6];
            cpt_pts[ii] = joint;
            S max_sinkage(0.0);
            for (int jj = 0; jj < max_checks; jj++) {
                S test_angle = max_angle - (2 * max_angle * jj / (S((double)max_checks) - 1));
                const S sa = s_sin(test_angle), ca = s_cos(test_angle);
                // test_rot = base_rot * roty(test_angle); test_pos = joint + test_rot * (0, 0, -tire_radius)
                Vec3<S> test_pos;
                for (int i = 0; i < 3; i++) {
                    const S col2 = base_rot.m[i][0] * sa + base_rot.m[i][2] * ca;
                    test_pos[i] = joint[i] + col2 * (-tire_radius);
                }
                const S test_sinkage = getAltitude(terrain, test_pos[0], test_pos[1]) - test_pos[2];
                if (value(test_sinkage) > value(max_sinkage)) max_sinkage = test_sinkage;
            }
            sinkages[ii] = max_sinkage;
        }
    }
    // :321-349
    void get_tire_cpt_vels(const S* X, Vec3<S> cpt_vels[4]) const {
        Vec3<S> lin_vel, ang_vel;
        for (int i = 0; i < 3; i++) { lin_vel[i] = X[14 + i]; ang_vel[i] = X[11 + i]; }
        for (int ii = 0; ii < 4; ii++) {
            const double e[3] = {tx_w[ii] + 0.0, ty_w[ii] + 0.0, tz_w[ii] + (-tire_radius)};
            const double r_ss[3][3] = {{0., -e[2], e[1]}, {e[2], 0., -e[0]}, {-e[1], e[0], 0.}};
            for (int i = 0; i < 3; i++)
                cpt_vels[ii][i] = lin_vel[i] - (r_ss[i][0] * ang_vel[0] + r_ss[i][1] * ang_vel[1] + r_ss[i][2] * ang_vel[2]);
        }
    }
    // NN: HybridDynamics.cpp:351-417 ; Bekker: BekkerDynamics.cpp:31-120
    void get_tire_f_ext(const S* X, Vec6<S> ext_forces[5]) {
        Vec3<S> cpt_points[4]; Mat3<S> rot;
        S sinkages[4];
        if (terrain.contact_search) {
            get_tire_cpts_sinkages_fancy(X, cpt_points, sinkages);
        } else {
            get_tire_cpts(X, cpt_points, rot);
            get_tire_sinkages(cpt_points, sinkages);
        }
        Vec3<S> cpt_vels[4];
        get_tire_cpt_vels(X, cpt_vels);
        for (int ii = 0; ii < 4; ii++) {
            S lin_force[3];
            if (tire_model == TIRE_NN) {
                S features[8], forces[4];
                for (int k = 4; k < 8; k++) features[k] = S(0.0);
                features[0] = cpt_vels[ii][0]; features[1] = cpt_vels[ii][1]; features[2] = X[17 + ii]; features[3] = sinkages[ii];
                tire_network.forward(features, forces, 0);  // always network 0 (:387)
                lin_force[0] = forces[0]; lin_force[1] = forces[1]; lin_force[2] = forces[2];
                lin_force[2] += -20 * cpt_vels[ii][2];      // :403
            } else {
                S features[8], bf[3], inputs[4], forces[4];
                // soil features: BekkerDynamics.cpp:59-63 (global 5-vector) — or per-contact-point soil grids (extension)
                S p[5];
                for (int k = 0; k < 5; k++) p[k] = bekker_params[k];
                if (terrain.kind == 3 && terrain.soil) {
                    GridCell<S> c = gridCell(terrain, cpt_points[ii][0], cpt_points[ii][1]);
                    for (int k = 0; k < 5; k++) p[k] = gridLerp(terrain, terrain.soil + (size_t)k * terrain.nx * terrain.ny, c);
                    S zero(0.0);
                    p[3] = cond_gt(p[3], zero, p[3], zero);
                }
                features[3] = 100.0 * p[0]; features[4] = 1000.0 * p[1]; features[5] = 1.0 * p[2]; features[6] = 1.0 * p[3]; features[7] = 1.0 * p[4];
                S vel_x_tan = .098 * X[17 + ii];
                S tire_vx = cpt_vels[ii][0];
                inputs[0] = sinkages[ii]; inputs[1] = tire_vx; inputs[2] = cpt_vels[ii][1]; inputs[3] = vel_x_tan;
                bekker_tire_model.get_features(inputs, bf);
                features[0] = bf[0]; features[1] = bf[1]; features[2] = bf[2];
                margin_note(0, std::fabs(value(sinkages[ii])));
                margin_note(1, std::fabs(value(vel_x_tan) - value(cpt_vels[ii][0])));
                if (value(sinkages[ii]) > 0.0) {
                    bekker_tire_model.get_forces(features, forces);
                } else {
                    for (int k = 0; k < 4; k++) forces[k] = S(0.0);
                }
                forces[0] = cond_gt(vel_x_tan - cpt_vels[ii][0], S(0.0), s_abs(forces[0]), -s_abs(forces[0]));  // :93
                lin_force[0] = forces[0]; lin_force[1] = forces[1]; lin_force[2] = forces[2];
                lin_force[2] += -200.0 * cpt_vels[ii][2];  // :106
            }
            Vec6<S>& wrench = ext_forces[ii + 1];
            wrench[0] = S(0.0); wrench[1] = S(0.0); wrench[2] = S(0.0);
            wrench[3] = lin_force[0]; wrench[4] = -lin_force[2]; wrench[5] = lin_force[1];
        }
    }
    // HybridDynamics.cpp:519-631
    void ODE(const S* X, S* Xd, const S* u) {
        S qd[4] = {u[0], u[1], u[0], u[1]}, qdd[4], tau[4] = {S(0.0), S(0.0), S(0.0), S(0.0)};
        Vec6<S> base_vel, base_acc;
        for (int i = 0; i < 6; i++) base_vel[i] = X[11 + i];
        Vec6<S> ext_forces[5];
        for (int l = 0; l < 5; l++) for (int i = 0; i < 6; i++) ext_forces[l][i] = S(0.0);
        get_tire_f_ext(X, ext_forces);
        S quat[4] = {X[0], X[1], X[2], X[3]};
        Mat3<S> rot = toMatrixRotation(X[0], X[1], X[2], X[3]);
        Vec3<S> gravity_lin; gravity_lin[0] = S(0.0); gravity_lin[1] = S(0.0); gravity_lin[2] = S(-9.81);
        gravity_lin = mulT(rot, gravity_lin);
        Vec6<S> gravity_b;
        gravity_b[0] = S(0.0); gravity_b[1] = S(0.0); gravity_b[2] = S(0.0);
        gravity_b[3] = gravity_lin[0]; gravity_b[4] = gravity_lin[1]; gravity_b[5] = gravity_lin[2];
        fwd_dynamics.fd(qdd, base_acc, base_vel, gravity_b, qd, tau, ext_forces);
        Vec3<S> ang_vel, lin_vel;
        for (int i = 0; i < 3; i++) { ang_vel[i] = base_vel[i]; lin_vel[i] = base_vel[3 + i]; }
        S quat_dot[4];
        calcQuatDot(quat, ang_vel, quat_dot);
        lin_vel = mul(rot, lin_vel);
        for (int i = 0; i < 4; i++) Xd[i] = quat_dot[i];
        for (int i = 0; i < 3; i++) Xd[4 + i] = lin_vel[i];
        Xd[7] = u[0]; Xd[8] = u[1]; Xd[9] = u[0]; Xd[10] = u[1];
        for (int i = 0; i < 6; i++) Xd[11 + i] = base_acc[i];
        for (int i = 17; i < 21; i++) Xd[i] = S(0.0);
    }
    static void renorm(S* x) {
        S n = s_sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        x[0] /= n; x[1] /= n; x[2] /= n; x[3] /= n;
    }
    // HybridDynamics.cpp:461-502
    void RK4(const S* X, S* Xt1, const S* u) {
        S temp[STATE_DIM], k1[STATE_DIM], k2[STATE_DIM], k3[STATE_DIM], k4[STATE_DIM];
        const double ts = timestep;
        ODE(X, k1, u);
        for (int i = 0; i < STATE_DIM; i++) temp[i] = X[i] + (.5 * ts) * k1[i];
        renorm(temp);
        ODE(temp, k2, u);
        for (int i = 0; i < STATE_DIM; i++) temp[i] = X[i] + (.5 * ts) * k2[i];
        renorm(temp);
        ODE(temp, k3, u);
        for (int i = 0; i < STATE_DIM; i++) temp[i] = X[i] + ts * k3[i];
        renorm(temp);
        ODE(temp, k4, u);
        for (int i = 0; i < STATE_DIM; i++) Xt1[i] = X[i] + (ts / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
        renorm(Xt1);
    }
    // HybridDynamics.cpp:142-190
    void settle() {
        S Xt1[STATE_DIM], u[2] = {S(0.0), S(0.0)};
        for (int j = 0; j < 20; j++) {
            for (int ii = 0; ii < 50; ii++) {
                state_[17] = state_[19] = u[0];
                state_[18] = state_[20] = u[1];
                Mat3<S> rot = toMatrixRotation(state_[0], state_[1], state_[2], state_[3]);
                Vec3<S> ang_vel, lin_vel;
                for (int i = 0; i < 3; i++) { ang_vel[i] = state_[11 + i]; lin_vel[i] = state_[14 + i]; }
                ang_vel = mul(rot, ang_vel);
                lin_vel = mul(rot, lin_vel);
                lin_vel[0] = S(0.0);
                lin_vel[1] = S(0.0);
                ang_vel = mulT(rot, ang_vel);
                lin_vel = mulT(rot, lin_vel);
                for (int i = 0; i < 3; i++) { state_[11 + i] = ang_vel[i]; state_[14 + i] = lin_vel[i]; }
                RK4(state_, Xt1, u);
                for (int i = 0; i < STATE_DIM; i++) state_[i] = Xt1[i];
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// System adapters: src/VehicleSystem.cpp, src/BekkerSystem.cpp
// ---------------------------------------------------------------------------------------------
struct GroundTruthDataRow {  // src/types/GroundTruthDataRow.h:3-15
    double time{0}, vl{0}, vr{0}, x{0}, y{0}, z{0}, yaw{0}, wz{0}, vx{0}, vy{0};
};

inline void bekkerDefaultParams(double* p) {  // BekkerSystem.cpp:21-28
    p[0] = .2976; p[1] = 2.083; p[2] = 0.8; p[3] = 0.0; p[4] = .3927;
}

template <class S> struct VehicleSystem {
    HybridDynamics<S> dyn;
    int substeps = 10;  // VehicleSystem.cpp:129
    VehicleSystem(int tire_model, const TerrainSpec& t) { dyn.tire_model = tire_model; dyn.terrain = t; }
    void setParams(const S* p) {
        if (dyn.tire_model == TIRE_NN) dyn.tire_network.setParams(p);
        else dyn.setBekkerParams(p);
    }
    // VehicleSystem.cpp:113-148 / BekkerSystem.cpp:145-178
    void integrate(const S* Xk, S* Xk1) {
        S x0[STATE_DIM], x1[STATE_DIM], u[2];
        for (int i = 0; i < STATE_DIM; i++) x0[i] = Xk[i];
        for (int i = 0; i < 2; i++) u[i] = Xk[STATE_DIM + i];
        for (int i = 0; i < STATE_DIM; i++) x1[i] = x0[i];
        for (int ii = 0; ii < substeps; ii++) {
            x0[17] = x0[19] = u[0];
            x0[18] = x0[20] = u[1];
            dyn.RK4(x0, x1, u);
            for (int i = 0; i < STATE_DIM; i++) x0[i] = x1[i];
        }
        for (int i = 0; i < STATE_DIM; i++) Xk1[i] = x1[i];
        Xk1[STATE_DIM] = S(0.0); Xk1[STATE_DIM + 1] = S(0.0);
    }
    // VehicleSystem.cpp:70-92 / BekkerSystem.cpp:102-124
    static S loss(const S* gt_vec, const S* vec) {
        S x_err = gt_vec[4] - vec[4];
        S y_err = gt_vec[5] - vec[5];
        S lin_err = s_sqrt((x_err * x_err) + (y_err * y_err));
        S yaw = eulerYaw(vec[3], vec[0], vec[1], vec[2]);
        S yaw_err = yaw - gt_vec[3];
        yaw_err = s_atan2(s_sin(yaw_err), s_cos(yaw_err));
        S ang_err = s_abs(yaw_err);
        return ang_err + lin_err;
    }
    // VehicleSystem.cpp:96-109 — quaternion-order bug at :103 replicated (ang part unused by Trainer)
    static void evaluate(const S* gt_vec, const S* vec, S& ang_mse, S& lin_mse) {
        S x_err = gt_vec[4] - vec[4];
        S y_err = gt_vec[5] - vec[5];
        lin_mse = (x_err * x_err) + (y_err * y_err);
        S yaw = eulerYaw(vec[0], vec[1], vec[2], vec[3]);
        S yaw_err = yaw - gt_vec[3];
        yaw_err = s_atan2(s_sin(yaw_err), s_cos(yaw_err));
        ang_mse = s_abs(yaw_err);
    }
    // VehicleSystem.cpp:187-238 (values only: the reference builds it from doubles)
    static void initializeState(const GroundTruthDataRow& gt, S* xk_robot /*23*/) {
        S xk[STATE_DIM], xb[STATE_DIM];
        for (int i = 0; i < STATE_DIM; i++) xk[i] = S(0.0);
        xk[2] = S(std::sin(gt.yaw / 2.0)); xk[3] = S(std::cos(gt.yaw / 2.0));
        xk[4] = S(gt.x); xk[5] = S(gt.y); xk[6] = S(gt.z);
        xk[13] = S(gt.wz); xk[14] = S(gt.vx); xk[15] = S(gt.vy);
        HybridDynamics<S>::initStateCOM(xk, xb);
        for (int i = 0; i < STATE_DIM; i++) xk_robot[i] = xb[i];
        xk_robot[21] = S(gt.vl); xk_robot[22] = S(gt.vr);
    }
    // VehicleSystem.cpp:151-177: settle from z=.16; returns quat + z (x=y=0), other entries 0
    void getDefaultInitialState(S* state21) {
        for (int i = 0; i < STATE_DIM; i++) state21[i] = S(0.0);
        dyn.initState();
        dyn.settle();
        for (int i = 0; i < 4; i++) state21[i] = dyn.state_[i];
        state21[6] = dyn.state_[6];
    }
};

// ---------------------------------------------------------------------------------------------
// Trainer: src/Trainer.cpp
// ---------------------------------------------------------------------------------------------
// evaluateTrajectory (Trainer.cpp:550-589).  x_list: [T][23] doubles.
inline double evaluateTrajectory(int tire_model, const TerrainSpec& terr, const double* params, const GroundTruthDataRow* traj, int T, double* x_list) {
    VehicleSystem<double> sys(tire_model, terr);
    sys.setParams(params);
    double xk[23], xk1[23];
    VehicleSystem<double>::initializeState(traj[0], xk);
    for (int k = 0; k < 23; k++) x_list[k] = xk[k];
    double traj_len = 0;
    double gt_vec[STATE_DIM];
    for (int k = 0; k < STATE_DIM; k++) gt_vec[k] = 0;
    for (int i = 1; i < T; i++) {
        xk[21] = traj[i - 1].vl; xk[22] = traj[i - 1].vr;
        sys.integrate(xk, xk1);
        for (int k = 0; k < 23; k++) { xk[k] = xk1[k]; x_list[(size_t)i * 23 + k] = xk[k]; }
        for (int k = 0; k < STATE_DIM; k++) gt_vec[k] = 0;
        gt_vec[3] = traj[i].yaw; gt_vec[4] = traj[i].x; gt_vec[5] = traj[i].y;
        double dx = traj[i].x - traj[i - 1].x, dy = traj[i].y - traj[i - 1].y;
        traj_len += std::sqrt(dx * dx + dy * dy);
    }
    double ang_mse, lin_mse;
    VehicleSystem<double>::evaluate(gt_vec, x_list + (size_t)(T - 1) * 23, ang_mse, lin_mse);
    return std::sqrt(lin_mse) / traj_len;
}

// trainTrajectory (Trainer.cpp:591-659) restated with a per-macro-step tape.  CppAD records ONE
// tape for the whole rollout and calls Reverse(1) on it; the reverse sweep of a chain
//   x_i = integrate(x_{i-1}; P),  L = sum_i l_i(x_i) / T / traj_len
// factorises exactly into per-step vector-Jacobian products  lambda_{i-1} = (dx_i/dx_{i-1})^T lambda_i,
// grad += (dx_i/dP)^T lambda_i, so taping each step separately (replaying forward values from
// x_list) performs the same partial-derivative arithmetic with a bounded tape (≈0.3 M nodes).
// Returns loss; grad416 receives dL/dparams.
inline double trainTrajectory(const TerrainSpec& terr, const double* params416, const GroundTruthDataRow* traj, int T, double* x_list, double* grad416) {
    // forward on doubles (values identical to the taped forward)
    VehicleSystem<double> sysd(TIRE_NN, terr);
    sysd.setParams(params416);
    double xk[23], xk1[23];
    VehicleSystem<double>::initializeState(traj[0], xk);
    for (int k = 0; k < 23; k++) x_list[k] = xk[k];
    double loss = 0, traj_len = 0;
    for (int i = 1; i < T; i++) {
        xk[21] = traj[i - 1].vl; xk[22] = traj[i - 1].vr;
        sysd.integrate(xk, xk1);
        for (int k = 0; k < 23; k++) { xk[k] = xk1[k]; x_list[(size_t)i * 23 + k] = xk[k]; }
        double gt_vec[STATE_DIM] = {0};
        gt_vec[3] = traj[i].yaw; gt_vec[4] = traj[i].x; gt_vec[5] = traj[i].y;
        double dx = traj[i].x - traj[i - 1].x, dy = traj[i].y - traj[i - 1].y;
        traj_len += std::sqrt(dx * dx + dy * dy);
        loss += VehicleSystem<double>::loss(gt_vec, xk);
    }
    loss /= (double)T;  // x_list.size() (:641)
    if (traj_len == 0) traj_len = 1;
    loss = loss / traj_len;
    const double seed = 1.0 / (double)T / traj_len;

    // reverse sweep
    for (int k = 0; k < NUM_NN_PARAMS; k++) grad416[k] = 0.0;
    double lambda[23] = {0};
    Tape tape;
    Tape* prev = active_tape();
    active_tape() = &tape;
    std::vector<double> adj;
    for (int i = T - 1; i >= 1; i--) {
        tape.clear();
        tape.nodes.reserve(400000);
        Rev P[NUM_NN_PARAMS], xin[23], xout[23];
        for (int k = 0; k < NUM_NN_PARAMS; k++) P[k] = Rev::independent(params416[k]);
        for (int k = 0; k < 23; k++) xin[k] = Rev::independent(x_list[(size_t)(i - 1) * 23 + k]);
        xin[21] = Rev(traj[i - 1].vl); xin[22] = Rev(traj[i - 1].vr);
        VehicleSystem<Rev> sys(TIRE_NN, terr);
        sys.setParams(P);
        sys.integrate(xin, xout);
        Rev gt_vec[STATE_DIM];
        gt_vec[3] = Rev(traj[i].yaw); gt_vec[4] = Rev(traj[i].x); gt_vec[5] = Rev(traj[i].y);
        Rev li = VehicleSystem<Rev>::loss(gt_vec, xout);
        adj.assign(tape.nodes.size(), 0.0);
        if (li.id >= 0) adj[li.id] += seed;
        for (int k = 0; k < 21; k++) if (xout[k].id >= 0) adj[xout[k].id] += lambda[k];
        tape.reverse(adj);
        for (int k = 0; k < NUM_NN_PARAMS; k++) grad416[k] += adj[P[k].id];
        for (int k = 0; k < 21; k++) lambda[k] = adj[NUM_NN_PARAMS + k];
    }
    active_tape() = prev;
    return loss;
}

// Monolithic-tape variant (exactly the reference's structure: one tape, one Reverse); used to
// validate the per-step factorisation above at small T.
inline double trainTrajectoryMonolithic(const TerrainSpec& terr, const double* params416, const GroundTruthDataRow* traj, int T, double* grad416) {
    Tape tape;
    Tape* prev = active_tape();
    active_tape() = &tape;
    std::vector<Rev> P(NUM_NN_PARAMS);
    for (int k = 0; k < NUM_NN_PARAMS; k++) P[k] = Rev::independent(params416[k]);
    VehicleSystem<Rev> sys(TIRE_NN, terr);
    sys.setParams(P.data());
    Rev xk[23], xk1[23];
    VehicleSystem<Rev>::initializeState(traj[0], xk);
    Rev loss(0.0), traj_len(0.0);
    for (int i = 1; i < T; i++) {
        xk[21] = Rev(traj[i - 1].vl); xk[22] = Rev(traj[i - 1].vr);
        sys.integrate(xk, xk1);
        for (int k = 0; k < 23; k++) xk[k] = xk1[k];
        Rev gt_vec[STATE_DIM];
        gt_vec[3] = Rev(traj[i].yaw); gt_vec[4] = Rev(traj[i].x); gt_vec[5] = Rev(traj[i].y);
        Rev dx(traj[i].x - traj[i - 1].x), dy(traj[i].y - traj[i - 1].y);
        traj_len += s_sqrt(dx * dx + dy * dy);
        loss += VehicleSystem<Rev>::loss(gt_vec, xk);
    }
    loss /= Rev((double)T);
    if (traj_len.v == 0) traj_len = Rev(1.0);
    loss = loss / traj_len;
    std::vector<double> adj(tape.nodes.size(), 0.0);
    if (loss.id >= 0) adj[loss.id] = 1.0;
    tape.reverse(adj);
    for (int k = 0; k < NUM_NN_PARAMS; k++) grad416[k] = adj[P[k].id];
    active_tape() = prev;
    return loss.v;
}

// Trainer::updateParams (Trainer.cpp:428-453): RMSProp with float-truncated gradient
inline void updateParams(double* params, double* squared_grad, const double* grad, int n, double lr = 1e-3, double l1_weight = 0.0) {
    for (int i = 0; i < n; i++) {
        float grad_idx = (float)grad[i];
        squared_grad[i] = 0.95 * squared_grad[i] + 0.05 * (double)(grad_idx * grad_idx);
        params[i] -= (lr / (std::sqrt(squared_grad[i]) + 1e-6)) * (double)grad_idx;
        params[i] -= lr * l1_weight * params[i];
    }
}

}  // namespace orc
