// ORACLE — TEST INFRASTRUCTURE ONLY ("parity unpinned", see jackal_oracle.hpp header).
// extern "C" surface over the templated restatement so that tests/ and bench.py (cpu_baseline /
// --impl reference legs) can drive it through ctypes.  Batch entry points use the same SoA
// [component][N] layout as include/avsim.h so GPU results can be compared element for element.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>
#include "jackal_oracle.hpp"

using namespace orc;

template <class F> static void parallelFor(int n, int nthreads, F f) {
    if (nthreads <= 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& t : th) t.join();
}

extern "C" {

struct orc_terrain {
    int kind, nx, ny;
    double x0, y0, dx, dy;
    const double* height;
    const double* soil;
    int contact_search;  // 0: contact point under the hub (reference ODE); 1: get_tire_cpts_sinkages_fancy
};

static TerrainSpec mkTerrain(const orc_terrain* t) {
    TerrainSpec s;
    if (!t) return s;
    s.kind = t->kind; s.nx = t->nx; s.ny = t->ny; s.x0 = t->x0; s.y0 = t->y0; s.dx = t->dx; s.dy = t->dy;
    s.height = t->height; s.soil = t->soil; s.contact_search = t->contact_search;
    return s;
}

int orc_num_nn_params() { return NUM_NN_PARAMS; }
void orc_nn_default_params(double* p416) { nnDefaultParams(p416); }
void orc_bekker_default_params(double* p5) { bekkerDefaultParams(p5); }
void orc_whole_body_com(double* com3) { wholeBodyCOM(com3); }

double orc_altitude(const orc_terrain* t, double x, double y) { return getAltitude(mkTerrain(t), x, y); }

// TireNetwork::forward with network index 0
void orc_tire_nn_forward(const double* p416, const double* in8, double* out4) {
    TireNetwork<double> net;
    net.setParams(p416);
    net.forward(in8, out4, 0);
}
// BekkerTireModel: inputs (zr,vx,vy,wR) + soil (kc,kphi,n0,n1,phi) -> features3 and forces4 (get_forces is
// evaluated unconditionally here; the sinkage gate lives in get_tire_f_ext)
void orc_bekker_forces(const double* inputs4, const double* soil5, double* features3, double* forces4) {
    BekkerTireModel<double> m;
    m.get_features(inputs4, features3);
    double f[8] = {features3[0], features3[1], features3[2], soil5[0], soil5[1], soil5[2], soil5[3], soil5[4]};
    m.get_forces(f, forces4);
}

void orc_ode(int tire_model, const double* params, const orc_terrain* t, const double* X21, const double* u2, double* Xd21) {
    VehicleSystem<double> sys(tire_model, mkTerrain(t));
    sys.setParams(params);
    sys.dyn.ODE(X21, Xd21, u2);
}
void orc_rk4(int tire_model, const double* params, const orc_terrain* t, const double* X21, const double* u2, double* X1) {
    VehicleSystem<double> sys(tire_model, mkTerrain(t));
    sys.setParams(params);
    sys.dyn.RK4(X21, X1, u2);
}
// base acceleration from the dense ABA for arbitrary external wrenches (tests the spatial algebra)
void orc_fd(const double* base_v6, const double* g6, const double* qd4, const double* fext30, double* base_a6, double* qdd4) {
    ForwardDynamics<double> fdyn;
    Vec6<double> v, g, a, fe[5];
    for (int i = 0; i < 6; i++) { v[i] = base_v6[i]; g[i] = g6[i]; }
    for (int l = 0; l < 5; l++) for (int i = 0; i < 6; i++) fe[l][i] = fext30[l * 6 + i];
    double tau[4] = {0, 0, 0, 0};
    fdyn.fd(qdd4, a, v, g, qd4, tau, fe);
    for (int i = 0; i < 6; i++) base_a6[i] = a[i];
}
void orc_integrate(int tire_model, const double* params, const orc_terrain* t, int substeps, const double* X23, double* X23out) {
    VehicleSystem<double> sys(tire_model, mkTerrain(t));
    sys.substeps = substeps;
    sys.setParams(params);
    sys.integrate(X23, X23out);
}
void orc_settle(int tire_model, const double* params, const orc_terrain* t, double* state21_full) {
    VehicleSystem<double> sys(tire_model, mkTerrain(t));
    sys.setParams(params);
    sys.dyn.initState();
    sys.dyn.settle();
    for (int i = 0; i < STATE_DIM; i++) state21_full[i] = sys.dyn.state_[i];
}
void orc_init_state_com(const double* start21, double* base21) { HybridDynamics<double>::initStateCOM(start21, base21); }
// rows: [T][10] = time,vl,vr,x,y,z,yaw,wz,vx,vy
static void toRows(const double* rows, int T, std::vector<GroundTruthDataRow>& out) {
    out.resize(T);
    for (int i = 0; i < T; i++) {
        const double* r = rows + (size_t)i * 10;
        out[i].time = r[0]; out[i].vl = r[1]; out[i].vr = r[2]; out[i].x = r[3]; out[i].y = r[4];
        out[i].z = r[5]; out[i].yaw = r[6]; out[i].wz = r[7]; out[i].vx = r[8]; out[i].vy = r[9];
    }
}
void orc_initialize_state(const double* row10, double* x23) {
    std::vector<GroundTruthDataRow> r;
    toRows(row10, 1, r);
    VehicleSystem<double>::initializeState(r[0], x23);
}
double orc_loss(const double* gt21, const double* x23) { return VehicleSystem<double>::loss(gt21, x23); }
void orc_evaluate(const double* gt21, const double* x23, double* ang_mse, double* lin_mse) {
    VehicleSystem<double>::evaluate(gt21, x23, *ang_mse, *lin_mse);
}
double orc_evaluate_trajectory(int tire_model, const double* params, const orc_terrain* t, const double* rows, int T, double* x_list) {
    std::vector<GroundTruthDataRow> traj;
    toRows(rows, T, traj);
    return evaluateTrajectory(tire_model, mkTerrain(t), params, traj.data(), T, x_list);
}
double orc_train_trajectory(const double* p416, const orc_terrain* t, const double* rows, int T, double* x_list, double* grad416) {
    std::vector<GroundTruthDataRow> traj;
    toRows(rows, T, traj);
    return trainTrajectory(mkTerrain(t), p416, traj.data(), T, x_list, grad416);
}
double orc_train_trajectory_monolithic(const double* p416, const orc_terrain* t, const double* rows, int T, double* grad416) {
    std::vector<GroundTruthDataRow> traj;
    toRows(rows, T, traj);
    return trainTrajectoryMonolithic(mkTerrain(t), p416, traj.data(), T, grad416);
}
// Reverse sweep through ONE BekkerSystem::integrate (BekkerSystem.cpp:145-178) on the tape scalar, as
// train_bekker.cpp would run it inside Trainer::trainTrajectory:  grad5 = (dx1/dp)^T lambda, xbar = (dx1/dx0)^T lambda.
// Returns the number of non-finite entries among grad5 and xbar (see tests/test_oracle.py: the reference's
// Bekker gradient is not finite whenever ze == zr, because theta_c = acos(1) is differentiated at 1).
int orc_bekker_integrate_vjp(const double* p5, const orc_terrain* t, const double* X23, const double* lambda21, double* x1_23, double* grad5,
                             double* xbar21) {
    Tape tape;
    Tape* prev = active_tape();
    active_tape() = &tape;
    Rev P[5], xin[23], xout[23];
    for (int k = 0; k < 5; k++) P[k] = Rev::independent(p5[k]);
    for (int k = 0; k < 21; k++) xin[k] = Rev::independent(X23[k]);
    xin[21] = Rev(X23[21]); xin[22] = Rev(X23[22]);
    VehicleSystem<Rev> sys(TIRE_BEKKER, mkTerrain(t));
    sys.setParams(P);
    sys.integrate(xin, xout);
    std::vector<double> adj(tape.nodes.size(), 0.0);
    for (int k = 0; k < 21; k++) if (xout[k].id >= 0) adj[xout[k].id] += lambda21[k];
    tape.reverse(adj);
    int bad = 0;
    for (int k = 0; k < 5; k++) { grad5[k] = adj[P[k].id]; bad += !std::isfinite(grad5[k]); }
    for (int k = 0; k < 21; k++) { xbar21[k] = adj[xin[k].id]; bad += !std::isfinite(xbar21[k]); }
    for (int k = 0; k < 23; k++) x1_23[k] = xout[k].v;
    active_tape() = prev;
    return bad;
}
// forward-mode directional check: d loss / d p[idx[k]] for k < 8, via Dual<8>
double orc_train_trajectory_dual8(const double* p416, const orc_terrain* t, const double* rows, int T, const int* idx8, double* dloss8) {
    typedef Dual<8> D;
    std::vector<GroundTruthDataRow> traj;
    toRows(rows, T, traj);
    std::vector<D> P(NUM_NN_PARAMS);
    for (int k = 0; k < NUM_NN_PARAMS; k++) P[k] = D(p416[k]);
    for (int k = 0; k < 8; k++) P[idx8[k]].d[k] = 1.0;
    VehicleSystem<D> sys(TIRE_NN, mkTerrain(t));
    sys.setParams(P.data());
    D xk[23], xk1[23];
    VehicleSystem<D>::initializeState(traj[0], xk);
    D loss(0.0);
    double traj_len = 0;
    for (int i = 1; i < T; i++) {
        xk[21] = D(traj[i - 1].vl); xk[22] = D(traj[i - 1].vr);
        sys.integrate(xk, xk1);
        for (int k = 0; k < 23; k++) xk[k] = xk1[k];
        D gt_vec[STATE_DIM];
        gt_vec[3] = D(traj[i].yaw); gt_vec[4] = D(traj[i].x); gt_vec[5] = D(traj[i].y);
        double dx = traj[i].x - traj[i - 1].x, dy = traj[i].y - traj[i - 1].y;
        traj_len += std::sqrt(dx * dx + dy * dy);
        loss += VehicleSystem<D>::loss(gt_vec, xk);
    }
    loss = loss / D((double)T);
    if (traj_len == 0) traj_len = 1;
    loss = loss / D(traj_len);
    for (int k = 0; k < 8; k++) dloss8[k] = loss.d[k];
    return loss.v;
}
void orc_rmsprop(double* params, double* squared_grad, const double* grad, int n, double lr, double l1) {
    updateParams(params, squared_grad, grad, n, lr, l1);
}

// op counts of one ODE evaluation / one RK4 step as the reference executes it
void orc_count_ops(int tire_model, const double* params, const double* X21, const double* u2, int what /*0 ODE, 1 RK4*/, long long* out9) {
    OpCount c;
    active_counter() = &c;
    {
        TerrainSpec terr;
        VehicleSystem<Cnt> sys(tire_model, terr);
        std::vector<Cnt> P(tire_model == TIRE_NN ? NUM_NN_PARAMS : 5);
        for (size_t k = 0; k < P.size(); k++) P[k] = Cnt(params[k]);
        sys.setParams(P.data());
        Cnt X[21], Xd[21], u[2] = {Cnt(u2[0]), Cnt(u2[1])};
        for (int i = 0; i < 21; i++) X[i] = Cnt(X21[i]);
        c = OpCount();
        if (what == 0) sys.dyn.ODE(X, Xd, u); else sys.dyn.RK4(X, Xd, u);
    }
    active_counter() = nullptr;
    out9[0] = c.addsub; out9[1] = c.mul; out9[2] = c.div; out9[3] = c.tanh; out9[4] = c.sqrt;
    out9[5] = c.trig; out9[6] = c.exp; out9[7] = c.pow; out9[8] = c.other;
}

// ------------------------------------------------------------------------------------------
// Batched (SoA) entry points mirroring include/avsim.h; one trajectory per work item, worker
// threads pull items from an atomic counter (the reference's Trainer::Worker pool,
// Trainer.cpp:204-305, 809-830, one trajectory per worker).
// ------------------------------------------------------------------------------------------

// x0 [21][N], ctrl [T-1][2][N], x_list [T][23][N] or null, x_final [21][N] or null
void orc_rollout_batch(int tire_model, const double* params, const orc_terrain* t, int N, int T, int substeps, const double* x0,
                       const double* ctrl, double* x_list, double* x_final, int nthreads) {
    TerrainSpec terr = mkTerrain(t);
    parallelFor(N, nthreads, [&](int n) {
        VehicleSystem<double> sys(tire_model, terr);
        sys.substeps = substeps;
        sys.setParams(params);
        double xk[23], xk1[23];
        for (int k = 0; k < 21; k++) xk[k] = x0[(size_t)k * N + n];
        xk[21] = T > 1 ? ctrl[(size_t)0 * N + n] : 0.0;
        xk[22] = T > 1 ? ctrl[(size_t)1 * N + n] : 0.0;
        if (x_list) for (int k = 0; k < 23; k++) x_list[((size_t)0 * 23 + k) * N + n] = xk[k];
        for (int i = 1; i < T; i++) {
            xk[21] = ctrl[((size_t)(i - 1) * 2 + 0) * N + n];
            xk[22] = ctrl[((size_t)(i - 1) * 2 + 1) * N + n];
            sys.integrate(xk, xk1);
            for (int k = 0; k < 23; k++) xk[k] = xk1[k];
            if (x_list) for (int k = 0; k < 23; k++) x_list[((size_t)i * 23 + k) * N + n] = xk[k];
        }
        if (x_final) for (int k = 0; k < 21; k++) x_final[(size_t)k * N + n] = xk[k];
    });
}

// gt [T][3][N] = (yaw, x, y) per row; loss [N]; grad_per_sample [416][N] or null;
// grad_sum [416], loss_sum, n_kept follow Trainer::train's zero-the-component filter (mode 0,
// Trainer.cpp:169-179) or combineResults' drop-the-sample filter (mode 1, Trainer.cpp:661-689).
void orc_rollout_grad_batch(const double* p416, const orc_terrain* t, int N, int T, int substeps, const double* x0, const double* ctrl,
                            const double* gt, double explode_thresh, int explode_mode, double* loss, double* grad_sum, double* loss_sum,
                            long long* n_kept, double* grad_per_sample, int nthreads) {
    (void)substeps;  // the reference hard-codes 10 (VehicleSystem.cpp:129); kept for signature parity
    TerrainSpec terr = mkTerrain(t);
    std::vector<double> grads((size_t)N * NUM_NN_PARAMS);
    parallelFor(N, nthreads, [&](int n) {
        // re-express the SoA sample as the GroundTruthDataRow window trainTrajectory consumes.  The kernel
        // API passes the initial 21-state directly, so initializeState is bypassed by a row whose
        // fields reproduce x0 (done by the caller for reference-shaped tests); here we tape from x0.
        std::vector<double> x_list((size_t)T * 23);
        // forward on doubles
        VehicleSystem<double> sysd(TIRE_NN, terr);
        sysd.setParams(p416);
        double xk[23], xk1[23];
        for (int k = 0; k < 21; k++) xk[k] = x0[(size_t)k * N + n];
        xk[21] = ctrl[(size_t)0 * N + n]; xk[22] = ctrl[(size_t)1 * N + n];
        for (int k = 0; k < 23; k++) x_list[k] = xk[k];
        double L = 0, traj_len = 0;
        auto G = [&](int i, int c) { return gt[((size_t)i * 3 + c) * N + n]; };
        for (int i = 1; i < T; i++) {
            xk[21] = ctrl[((size_t)(i - 1) * 2 + 0) * N + n];
            xk[22] = ctrl[((size_t)(i - 1) * 2 + 1) * N + n];
            sysd.integrate(xk, xk1);
            for (int k = 0; k < 23; k++) { xk[k] = xk1[k]; x_list[(size_t)i * 23 + k] = xk[k]; }
            double gt_vec[STATE_DIM] = {0};
            gt_vec[3] = G(i, 0); gt_vec[4] = G(i, 1); gt_vec[5] = G(i, 2);
            double dx = G(i, 1) - G(i - 1, 1), dy = G(i, 2) - G(i - 1, 2);
            traj_len += std::sqrt(dx * dx + dy * dy);
            L += VehicleSystem<double>::loss(gt_vec, xk);
        }
        L /= (double)T;
        if (traj_len == 0) traj_len = 1;
        L = L / traj_len;
        const double seed = 1.0 / (double)T / traj_len;
        double* g = grads.data() + (size_t)n * NUM_NN_PARAMS;
        for (int k = 0; k < NUM_NN_PARAMS; k++) g[k] = 0;
        double lambda[23] = {0};
        Tape tape;
        active_tape() = &tape;
        std::vector<double> adj;
        for (int i = T - 1; i >= 1; i--) {
            tape.clear();
            tape.nodes.reserve(400000);
            Rev P[NUM_NN_PARAMS], xin[23], xout[23];
            for (int k = 0; k < NUM_NN_PARAMS; k++) P[k] = Rev::independent(p416[k]);
            for (int k = 0; k < 23; k++) xin[k] = Rev::independent(x_list[(size_t)(i - 1) * 23 + k]);
            xin[21] = Rev(ctrl[((size_t)(i - 1) * 2 + 0) * N + n]);
            xin[22] = Rev(ctrl[((size_t)(i - 1) * 2 + 1) * N + n]);
            VehicleSystem<Rev> sys(TIRE_NN, terr);
            sys.setParams(P);
            sys.integrate(xin, xout);
            Rev gt_vec[STATE_DIM];
            gt_vec[3] = Rev(G(i, 0)); gt_vec[4] = Rev(G(i, 1)); gt_vec[5] = Rev(G(i, 2));
            Rev li = VehicleSystem<Rev>::loss(gt_vec, xout);
            adj.assign(tape.nodes.size(), 0.0);
            if (li.id >= 0) adj[li.id] += seed;
            for (int k = 0; k < 21; k++) if (xout[k].id >= 0) adj[xout[k].id] += lambda[k];
            tape.reverse(adj);
            for (int k = 0; k < NUM_NN_PARAMS; k++) g[k] += adj[P[k].id];
            for (int k = 0; k < 21; k++) lambda[k] = adj[NUM_NN_PARAMS + k];
        }
        active_tape() = nullptr;
        loss[n] = L;
    });
    for (int k = 0; k < NUM_NN_PARAMS; k++) grad_sum[k] = 0;
    *loss_sum = 0; *n_kept = 0;
    for (int n = 0; n < N; n++) {
        double* g = grads.data() + (size_t)n * NUM_NN_PARAMS;
        if (grad_per_sample) for (int k = 0; k < NUM_NN_PARAMS; k++) grad_per_sample[(size_t)k * N + n] = g[k];
        bool explode = false;
        for (int k = 0; k < NUM_NN_PARAMS; k++) if (std::fabs(g[k]) > explode_thresh) explode = true;
        if (explode_mode == 1 && explode) continue;
        for (int k = 0; k < NUM_NN_PARAMS; k++) {
            double v = g[k];
            if (explode_mode == 0 && std::fabs(v) > explode_thresh) v = 0.0;
            grad_sum[k] += v;
        }
        *loss_sum += loss[n];
        *n_kept += 1;
    }
}

// forward rollouts that also report the branch margins (jackal_oracle.hpp): margins [NUM_MARGINS][N] = per trajectory the
// minimum over every ODE evaluation of the rollout
void orc_rollout_batch_margins(int tire_model, const double* params, const orc_terrain* t, int N, int T, int substeps, const double* x0,
                               const double* ctrl, double* x_list, double* x_final, double* margins, int nthreads) {
    TerrainSpec terr = mkTerrain(t);
    parallelFor(N, nthreads, [&](int n) {
        double m[NUM_MARGINS];
        for (int k = 0; k < NUM_MARGINS; k++) m[k] = 1e300;
        active_margin() = m;
        VehicleSystem<double> sys(tire_model, terr);
        sys.substeps = substeps;
        sys.setParams(params);
        double xk[23], xk1[23];
        for (int k = 0; k < 21; k++) xk[k] = x0[(size_t)k * N + n];
        xk[21] = T > 1 ? ctrl[(size_t)0 * N + n] : 0.0;
        xk[22] = T > 1 ? ctrl[(size_t)1 * N + n] : 0.0;
        if (x_list) for (int k = 0; k < 23; k++) x_list[((size_t)0 * 23 + k) * N + n] = xk[k];
        for (int i = 1; i < T; i++) {
            xk[21] = ctrl[((size_t)(i - 1) * 2 + 0) * N + n];
            xk[22] = ctrl[((size_t)(i - 1) * 2 + 1) * N + n];
            sys.integrate(xk, xk1);
            for (int k = 0; k < 23; k++) xk[k] = xk1[k];
            if (x_list) for (int k = 0; k < 23; k++) x_list[((size_t)i * 23 + k) * N + n] = xk[k];
        }
        if (x_final) for (int k = 0; k < 21; k++) x_final[(size_t)k * N + n] = xk[k];
        active_margin() = nullptr;
        for (int k = 0; k < NUM_MARGINS; k++) margins[(size_t)k * N + n] = m[k];
    });
}

// HybridDynamics::get_tire_cpts_sinkages_fancy on one state: sinkages [4], joint centres [4][3]
void orc_fancy_sinkages(const orc_terrain* t, const double* X21, double* sinkages4, double* pts12) {
    HybridDynamics<double> dyn;
    dyn.terrain = mkTerrain(t);
    Vec3<double> pts[4];
    dyn.get_tire_cpts_sinkages_fancy(X21, pts, sinkages4);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) pts12[i * 3 + k] = pts[i][k];
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
