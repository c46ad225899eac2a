// ref_capi.cpp — TEST INFRASTRUCTURE ONLY (oracle/README.md).
//
// C entry points over the reference's OWN classes, for pinning the oracle restatement (oracle/jackal_oracle.hpp) and, via
// golden fixtures, the CUDA path.  This file is glue written for this repository; everything it calls — HybridDynamics,
// TireNetwork, BekkerDynamics, BekkerTireModel, VehicleSystem, BekkerSystem, Trainer, the RobCoGen-generated forward
// dynamics / transforms / inertia properties, TestTerrainMaps — is compiled FROM THE REFERENCE'S SOURCE FILES WHERE THEY LIE
// (/root/reference/src/..., never copied; recipe: oracle/Makefile target `ref`) into oracle/_ref/libjackal_ref.so.
// The three third-party libraries those sources include are absent from this image; they are replaced by the stand-in
// headers under oracle/refshim/ (Eigen dense algebra, CppAD's AD<double> + ADFun::Reverse, iit-rbd's spatial helpers), whose
// semantics are stated in their headers.  So: first-party arithmetic = the reference's own code; third-party arithmetic =
// restated from the published definitions (rounding-level differences only are possible there).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <vector>

#include "BekkerSystem.h"
#include "BekkerTireModel.h"
#include "TestTerrainMaps.h"
#include "Trainer.h"
#include "VehicleSystem.h"
#include "generated/miscellaneous.h"
#include "utils.h"

typedef Eigen::Matrix<ADF, Eigen::Dynamic, 1> VectorAD;

namespace {
std::shared_ptr<const TerrainMap<ADF>> mkMap(int kind) {
    if (kind == 1) return std::make_shared<const SlopeTerrainMap<ADF>>();
    if (kind == 2) return std::make_shared<const BumpyTerrainMap<ADF>>();
    return std::make_shared<const FlatTerrainMap<ADF>>();
}
std::shared_ptr<SystemFactory<ADF>> mkFactory(int tire, int terr) {
    if (tire == 1) return std::make_shared<BekkerSystemFactory<ADF>>(mkMap(terr));
    return std::make_shared<VehicleSystemFactory<ADF>>(mkMap(terr));
}
std::shared_ptr<System<ADF>> mkSystem(int tire, int terr, const double* params) {
    std::shared_ptr<System<ADF>> sys = mkFactory(tire, terr)->makeSystem();
    VectorAD p(sys->getNumParams());
    for (int i = 0; i < sys->getNumParams(); i++) p[i] = params[i];
    sys->setParams(p);
    return sys;
}
// The reference prints progress lines to std::cout.  libstdc++ is linked statically into this library, so its std::cout is
// private to it: it is pointed at a sink ONCE, for good (swapping the buffer per call would race between the host's threads
// and leave a dangling buffer behind for the flush at exit).  The sink is leaked on purpose.
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
struct QuietOnce {
    QuietOnce() { if (!std::getenv("JACKAL_REF_VERBOSE")) std::cout.rdbuf(new NullBuf()); }
};
QuietOnce g_quiet_once;
struct Quiet {};  // (marks the entry points whose callees print)
GroundTruthDataRow mkRow(const double* r) {  // (time, vl, vr, x, y, z, yaw, wz, vx, vy), as oracle_capi.cpp
    GroundTruthDataRow g;
    g.time = r[0]; g.vl = r[1]; g.vr = r[2]; g.x = r[3]; g.y = r[4]; g.z = r[5]; g.yaw = r[6]; g.wz = r[7]; g.vx = r[8]; g.vy = r[9];
    return g;
}
}  // namespace

extern "C" {

int ref_num_params(int tire) { Quiet q; return mkFactory(tire, 0)->makeSystem()->getNumParams(); }

// System::getDefaultParams: TireNetwork::load_model weights / BekkerSystem defaults
void ref_default_params(int tire, double* out) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkFactory(tire, 0)->makeSystem();
    VectorAD p(sys->getNumParams());
    sys->getDefaultParams(p);
    for (int i = 0; i < sys->getNumParams(); i++) out[i] = CppAD::Value(p[i]);
}

void ref_whole_body_com(double* com3) {
    Jackal::rcg::InertiaProperties inertias;
    Jackal::rcg::HomogeneousTransforms ht;
    Jackal::rcg::Vector3 c = Jackal::rcg::getWholeBodyCOM(inertias, ht);
    for (int i = 0; i < 3; i++) com3[i] = CppAD::Value(c[i]);
}

// TerrainMap::getAltitude of the reference's Flat / Slope / Bumpy maps
double ref_altitude(int terr, double x, double y) { return CppAD::Value(mkMap(terr)->getAltitude(ADF(x), ADF(y))); }

// TireNetwork::forward(in8, out4, 0) with the given 416 parameters
void ref_tire_nn_forward(const double* p416, const double* in8, double* out4) {
    TireNetwork net;
    VectorAD p(net.getNumParams());
    for (int i = 0; i < net.getNumParams(); i++) p[i] = p416[i];
    net.setParams(p, 0);
    Eigen::Matrix<ADF, 8, 1> in;
    Eigen::Matrix<ADF, 4, 1> out;
    for (int i = 0; i < 8; i++) in[i] = in8[i];
    net.forward(in, out, 0);
    for (int i = 0; i < 4; i++) out4[i] = CppAD::Value(out[i]);
}

// BekkerTireModel::get_features + get_forces: inputs (zr, vx, vy, w R), soil features (kc, kphi, n0, n1, phi)
void ref_bekker_forces(const double* inputs4, const double* soil5, double* features3, double* forces4) {
    BekkerTireModel m;
    Eigen::Matrix<ADF, 4, 1> in;
    for (int i = 0; i < 4; i++) in[i] = inputs4[i];
    Eigen::Matrix<ADF, 3, 1> f = m.get_features(in);
    Eigen::Matrix<ADF, 8, 1> feat;
    for (int i = 0; i < 3; i++) { feat[i] = f[i]; features3[i] = CppAD::Value(f[i]); }
    for (int i = 0; i < 5; i++) feat[3 + i] = soil5[i];
    Eigen::Matrix<ADF, 4, 1> w = m.get_forces(feat);
    for (int i = 0; i < 4; i++) forces4[i] = CppAD::Value(w[i]);
}

// ForwardDynamics::fd at q = 0: base_v6, g6, qd4, fext [5][6] (base link, FL, FR, RL, RR) -> base_a6, qdd4
void ref_fd(const double* base_v6, const double* g6, const double* qd4, const double* fext30, double* base_a6, double* qdd4) {
    Jackal::rcg::InertiaProperties inertias;
    Jackal::rcg::MotionTransforms transforms;
    Jackal::rcg::ForwardDynamics fd(inertias, transforms);
    Jackal::rcg::JointState q(0, 0, 0, 0), qd, qdd, tau(0, 0, 0, 0);
    Jackal::rcg::Velocity v;
    Jackal::rcg::Acceleration g, a;
    for (int i = 0; i < 6; i++) { v[i] = base_v6[i]; g[i] = g6[i]; }
    for (int i = 0; i < 4; i++) qd[i] = qd4[i];
    Jackal::rcg::LinkDataMap<Jackal::rcg::Force> fext(Jackal::rcg::Force::Zero());
    for (int l = 0; l < 5; l++) {
        Jackal::rcg::Force f;
        for (int i = 0; i < 6; i++) f[i] = fext30[l * 6 + i];
        fext[Jackal::rcg::orderedLinkIDs[l]] = f;
    }
    fd.fd(qdd, a, v, g, q, qd, tau, fext);
    for (int i = 0; i < 6; i++) base_a6[i] = CppAD::Value(a[i]);
    for (int i = 0; i < 4; i++) qdd4[i] = CppAD::Value(qdd[i]);
}

// HybridDynamics::get_tire_cpts_sinkages_fancy (the 10-angle contact search): sinkages [4], reported points [4][3]
void ref_fancy_sinkages(int terr, const double* X21, double* sinkages4, double* pts12) {
    HybridDynamics dyn;
    dyn.setTerrainMap(mkMap(terr));
    Eigen::Matrix<ADF, 21, 1> X;
    for (int i = 0; i < 21; i++) X[i] = X21[i];
    Eigen::Matrix<ADF, 3, 1> pts[4];
    Eigen::Matrix<ADF, 3, 3> rots[4];
    ADF sk[4];
    dyn.get_tire_cpts_sinkages_fancy(X, pts, rots, sk);
    for (int i = 0; i < 4; i++) { sinkages4[i] = CppAD::Value(sk[i]); for (int k = 0; k < 3; k++) pts12[i * 3 + k] = CppAD::Value(pts[i][k]); }
}

// BaseNetwork::setParams + forward (src/BaseNetwork.cpp): (wz, vx, vy) -> (nz, fx, fy)
void ref_base_network_forward(const double* p131, const double* in3, double* out3) {
    BaseNetwork net;
    VectorAD p(net.getNumParams());
    for (int i = 0; i < net.getNumParams(); i++) p[i] = p131[i];
    net.setParams(p, 0);
    Eigen::Matrix<ADF, 3, 1> in, out;
    for (int i = 0; i < 3; i++) in[i] = in3[i];
    net.forward(in, out);
    for (int i = 0; i < 3; i++) out3[i] = CppAD::Value(out[i]);
}

// System::forward: X23 = state + controls -> Xd23
void ref_ode(int tire, const double* params, int terr, const double* X23, double* Xd23) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkSystem(tire, terr, params);
    VectorAD X(23), Xd(23);
    for (int i = 0; i < 23; i++) X[i] = X23[i];
    // System::forward asserts size == state dim in the reference but indexes 23 entries; give it what it indexes
    sys->forward(X, Xd);
    for (int i = 0; i < 23; i++) Xd23[i] = CppAD::Value(Xd[i]);
}

// System::integrate: 10 RK4 steps of 1 ms
void ref_integrate(int tire, const double* params, int terr, const double* X23, double* X23out) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkSystem(tire, terr, params);
    VectorAD X(23), X1(23);
    for (int i = 0; i < 23; i++) X[i] = X23[i];
    sys->integrate(X, X1);
    for (int i = 0; i < 23; i++) X23out[i] = CppAD::Value(X1[i]);
}

// System::getDefaultInitialState (initState + settle): quaternion in [0..3], z in [6]
void ref_default_initial_state(int tire, const double* params, int terr, double* state21) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkSystem(tire, terr, params);
    VectorAD s = VectorAD::Zero(21);
    sys->getDefaultInitialState(s);
    for (int i = 0; i < 21; i++) state21[i] = CppAD::Value(s[i]);
}

// System::initializeState(GroundTruthDataRow)
void ref_initialize_state(const double* row10, double* x23) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkFactory(0, 0)->makeSystem();
    VectorAD x = sys->initializeState(mkRow(row10));
    for (int i = 0; i < 23; i++) x23[i] = CppAD::Value(x[i]);
}

double ref_loss(const double* gt21, const double* x23) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkFactory(0, 0)->makeSystem();
    VectorAD g(21), x(23);
    for (int i = 0; i < 21; i++) g[i] = gt21[i];
    for (int i = 0; i < 23; i++) x[i] = x23[i];
    return CppAD::Value(sys->loss(g, x));
}

// rollout from an explicit 23-vector: x_list [T][23]; row i uses the controls ctrl[i-1] (Trainer.cpp:564-569)
void ref_rollout(int tire, const double* params, int terr, int T, const double* x0_23, const double* ctrl /*[T-1][2]*/, double* x_list) {
    Quiet q;
    std::shared_ptr<System<ADF>> sys = mkSystem(tire, terr, params);
    VectorAD xk(23), xk1(23);
    for (int i = 0; i < 23; i++) { xk[i] = x0_23[i]; x_list[i] = x0_23[i]; }
    for (int i = 1; i < T; i++) {
        xk[21] = ctrl[(i - 1) * 2 + 0];
        xk[22] = ctrl[(i - 1) * 2 + 1];
        sys->integrate(xk, xk1);
        xk = xk1;
        for (int k = 0; k < 23; k++) x_list[(size_t)i * 23 + k] = CppAD::Value(xk[k]);
    }
}

// Trainer::trainTrajectory — the reference's own tape + ADFun::Reverse(1) — on rows [T][10]; returns the loss
double ref_train_trajectory(const double* p416, int terr, const double* rows, int T, double* x_list /*[T][23] or null*/, double* grad416) {
    Quiet q;
    Trainer trainer(mkFactory(0, terr), 1);
    for (int i = 0; i < (int)trainer.m_params.size(); i++) trainer.m_params[i] = p416[i];
    trainer.m_system_adf->setParams(trainer.m_params);
    std::vector<GroundTruthDataRow> traj;
    for (int i = 0; i < T; i++) traj.push_back(mkRow(rows + (size_t)i * 10));
    std::vector<VectorAD> xl(T);
    Trainer::VectorF grad;
    double loss = 0;
    trainer.trainTrajectory(traj, xl, grad, loss);
    for (int i = 0; i < (int)grad.size(); i++) grad416[i] = grad[i];
    if (x_list) for (int i = 0; i < T; i++) for (int k = 0; k < 23; k++) x_list[(size_t)i * 23 + k] = CppAD::Value(xl[i][k]);
    return loss;
}

// Trainer::evaluateTrajectory: loss = sqrt(lin_mse_final) / traj_len
double ref_evaluate_trajectory(int tire, const double* params, int terr, const double* rows, int T, double* x_list) {
    Quiet q;
    Trainer trainer(mkFactory(tire, terr), 1);
    for (int i = 0; i < (int)trainer.m_params.size(); i++) trainer.m_params[i] = params[i];
    std::vector<GroundTruthDataRow> traj;
    for (int i = 0; i < T; i++) traj.push_back(mkRow(rows + (size_t)i * 10));
    std::vector<VectorAD> xl(T);
    double loss = 0;
    trainer.evaluateTrajectory(traj, xl, loss);
    if (x_list) for (int i = 0; i < T; i++) for (int k = 0; k < 23; k++) x_list[(size_t)i * 23 + k] = CppAD::Value(xl[i][k]);
    return loss;
}

// Trainer::updateParams (RMSProp with the float-truncated gradient) on explicit state
void ref_update_params(double* params, double* squared_grad, const double* grad, int n) {
    Quiet q;
    Trainer trainer(mkFactory(0, 0), 1);
    Trainer::VectorF g(n);
    for (int i = 0; i < n; i++) { trainer.m_params[i] = params[i]; trainer.m_squared_grad[i] = squared_grad[i]; g[i] = grad[i]; }
    trainer.updateParams(g);
    for (int i = 0; i < n; i++) { params[i] = CppAD::Value(trainer.m_params[i]); squared_grad[i] = CppAD::Value(trainer.m_squared_grad[i]); }
}

// z and quaternion the reference's Trainer settles to at construction (computeEqState)
void ref_trainer_eq_state(int tire, int terr, double* z_stable, double* quat4) {
    Quiet q;
    Trainer trainer(mkFactory(tire, terr), 1);
    *z_stable = CppAD::Value(trainer.m_z_stable);
    for (int i = 0; i < 4; i++) quat4[i] = CppAD::Value(trainer.m_quat_stable[i]);
}

}  // extern "C"
