// GpuSystem.h — System<Scalar> implemented over the C ABI of libavsim.so (include/avsim.h).
//
// One class serves both adapters of the reference: VehicleSystem (NN tire model, reference
// src/VehicleSystem.{h,cpp}) and BekkerSystem (reference src/BekkerSystem.{h,cpp}); they differ only in
// the tire model of the context and the parameter vector.  The scalar virtuals (forward / integrate / ...)
// each issue a 1-vehicle launch — correct but latency-bound; Trainer uses the batched virtuals.
// All systems of one tire model share ONE context per process (the reference builds a fresh HybridDynamics per
// trajectory, src/Trainer.cpp:596; a CUDA context per trajectory would be absurd), so a System object is a
// light handle: creating many is cheap, and they are not thread-safe against each other (as include/avsim.h
// says: one context per GPU, caller serialises).
#pragma once
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/avsim.h"
#include "TestTerrainMaps.h"
#include "types/System.h"
#include "types/SystemFactory.h"

namespace avs {

inline void check(avs_ctx* ctx, int rc, const char* what) {
    if (rc != AVS_OK) throw std::runtime_error(std::string(what) + ": " + avs_strerror(rc) + " — " + (ctx ? avs_last_error(ctx) : "no context") +
                                               " (libavsim has no CPU fallback)");
}

// process-wide context per tire model
class SharedContext {
public:
    static std::shared_ptr<SharedContext> get(int tire_model) {
        static std::mutex mu;
        static std::weak_ptr<SharedContext> slots[2];
        std::lock_guard<std::mutex> lk(mu);
        std::shared_ptr<SharedContext> sp = slots[tire_model].lock();
        if (!sp) {
            sp.reset(new SharedContext(tire_model));
            slots[tire_model] = sp;
        }
        return sp;
    }
    ~SharedContext() { avs_destroy(ctx); }
    avs_ctx* ctx = nullptr;
    // which handle's terrain / parameters are currently loaded into the context.  Handle ids are drawn from a
    // process-wide counter and never reused (an address would be: Trainer makes and destroys a System per trajectory)
    unsigned long long terrain_owner = 0, params_owner = 0;
    static unsigned long long nextHandleId() {
        static std::atomic<unsigned long long> next{1};
        return next.fetch_add(1);
    }

private:
    explicit SharedContext(int tire_model) {
        const char* dev = std::getenv("AVSL_DEVICE");
        const int rc = avs_create(&ctx, dev ? std::atoi(dev) : 0, tire_model);
        check(nullptr, rc, "avs_create");
    }
};

template <typename Scalar, int TIRE> class GpuSystem : public System<Scalar> {
public:
    typedef typename System<Scalar>::VectorS VectorS;
    typedef typename System<Scalar>::MatrixS MatrixS;

    explicit GpuSystem(const std::shared_ptr<const TerrainMap<Scalar>>& map)
        : m_map(map), m_shared(SharedContext::get(TIRE)), m_id(SharedContext::nextHandleId()) {
        this->setNumParams(TIRE == AVS_TIRE_NN ? AVS_NUM_NN_PARAMS : AVS_NUM_BEKKER_PARAMS);
        this->setStateDim(AVS_STATE_DIM);
        this->setControlDim(AVS_CNTRL_DIM);
        m_params.resize(this->getNumParams());
        if (TIRE == AVS_TIRE_NN) avs_nn_default_params(m_params.data());
        else avs_bekker_default_params(m_params.data());
    }

    void getDefaultParams(VectorS& params) override {
        if (TIRE == AVS_TIRE_NN) {
            std::cout << "Loading Model\n";  // TireNetwork::load_model
            avs_nn_default_params(m_params.data());
        } else {
            avs_bekker_default_params(m_params.data());
        }
        params.resize(m_params.size());
        for (size_t i = 0; i < m_params.size(); i++) params[i] = m_params[i];
    }
    void setParams(const VectorS& params) override {
        for (size_t i = 0; i < m_params.size(); i++) m_params[i] = params[i];
        if (TIRE == AVS_TIRE_BEKKER) {
            std::cout << "Params: " << m_params[0] << ", " << m_params[1] << ", " << m_params[2] << ", " << m_params[3] << ", " << m_params[4] << "\n";
        }
        m_dirty = true;
    }
    void getParams(VectorS& params) override {
        params.resize(m_params.size());
        for (size_t i = 0; i < m_params.size(); i++) params[i] = m_params[i];
        if (TIRE == AVS_TIRE_BEKKER && params[3] < 0) params[3] = 0;  // BekkerDynamics::setParams clamp
    }

    // state derivative of the 21 states; controls ride in X[21..22] (reference VehicleSystem.cpp:38-66)
    void forward(const VectorS& X, VectorS& Xd) override {
        bind();
        double x[21], u[2] = {X[21], X[22]}, xd[21];
        for (int i = 0; i < 21; i++) x[i] = X[i];
        check(ctx(), avs_ode(ctx(), 1, x, u, xd), "avs_ode");
        for (int i = 0; i < 21; i++) Xd[i] = xd[i];
        Xd[21] = 0; Xd[22] = 0;
    }
    // 10 RK4 steps of 1 ms with the controls in Xk[21..22] (reference VehicleSystem.cpp:113-148)
    void integrate(const VectorS& Xk, VectorS& Xk1) override {
        bind();
        double x[21], u[2] = {Xk[21], Xk[22]}, xf[21];
        for (int i = 0; i < 21; i++) x[i] = Xk[i];
        // integrate() overwrites the joint velocities with the controls before every RK4 step
        check(ctx(), avs_rollout(ctx(), 1, 2, 10, x, u, nullptr, xf), "avs_rollout");
        for (int i = 0; i < 21; i++) Xk1[i] = xf[i];
        Xk1[21] = 0; Xk1[22] = 0;
    }
    // |wrap(yaw - yaw_gt)| + |(x,y) - (x,y)_gt|   (reference VehicleSystem.cpp:70-92)
    Scalar loss(const VectorS& gt_vec, VectorS& vec) override {
        const double xe = gt_vec[4] - vec[4], ye = gt_vec[5] - vec[5];
        const double lin = std::sqrt(xe * xe + ye * ye);
        const double yaw = yawOf(vec[3], vec[0], vec[1], vec[2]);
        const double e = yaw - gt_vec[3];
        return std::fabs(std::atan2(std::sin(e), std::cos(e))) + lin;
    }
    // squared position error + (mis-ordered, unused) heading error (reference VehicleSystem.cpp:96-109)
    void evaluate(const VectorS& gt_vec, const VectorS& vec, Scalar& ang_mse, Scalar& lin_mse) override {
        const double xe = gt_vec[4] - vec[4], ye = gt_vec[5] - vec[5];
        lin_mse = xe * xe + ye * ye;
        const double yaw = yawOf(vec[0], vec[1], vec[2], vec[3]);  // argument order as in the reference (:103)
        const double e = yaw - gt_vec[3];
        ang_mse = std::fabs(std::atan2(std::sin(e), std::cos(e)));
    }
    // yaw-only quaternion, COM velocities -> base-link state, controls appended (reference VehicleSystem.cpp:187-238)
    VectorS initializeState(const GroundTruthDataRow& gt) override {
        double s[21] = {0}, b[21];
        s[2] = std::sin(gt.yaw / 2.0); s[3] = std::cos(gt.yaw / 2.0);
        s[4] = gt.x; s[5] = gt.y; s[6] = gt.z;
        s[13] = gt.wz; s[14] = gt.vx; s[15] = gt.vy;
        avs_init_state_com(s, b);
        VectorS out(23);
        for (int i = 0; i < 21; i++) out[i] = b[i];
        out[21] = gt.vl; out[22] = gt.vr;
        return out;
    }
    // settle from z = 0.16 and report quaternion + z (reference VehicleSystem.cpp:151-177)
    void getDefaultInitialState(VectorS& state) override {
        bind();
        double s[21];
        check(ctx(), avs_settle(ctx(), s), "avs_settle");
        for (size_t i = 0; i < state.size(); i++) state[i] = 0;
        for (int i = 0; i < 4; i++) state[i] = s[i];
        state[6] = s[6];
        if (TIRE == AVS_TIRE_NN) {
            std::cout << "Stabilized z: " << s[6] << "\n";
            std::cout << "Stabilized Quaternion: " << s[0] << ", " << s[1] << ", " << s[2] << ", " << s[3] << "\n";
        }
    }

    // ---- batched path
    void rolloutBatch(int N, int T, const double* x0, const double* ctrl, double* x_list, double* x_final) override {
        bind();
        check(ctx(), avs_rollout(ctx(), N, T, 10, x0, ctrl, x_list, x_final), "avs_rollout");
    }
    void rolloutGradBatch(int N, int T, const double* x0, const double* ctrl, const double* gt, double thresh, int mode, double* loss, double* reduced,
                          double* grad_per_sample) override {
        if (TIRE != AVS_TIRE_NN) throw std::runtime_error("rolloutGradBatch: the adjoint exists for the NN tire model only");
        bind();
        check(ctx(), avs_rollout_grad(ctx(), N, T, 10, x0, ctrl, gt, thresh, mode, loss, reduced, grad_per_sample), "avs_rollout_grad");
    }
    void applyRmsProp(const double* reduced, double lr, double l1) override {
        if (TIRE != AVS_TIRE_NN) throw std::runtime_error("applyRmsProp: NN tire model only");
        bind();
        check(ctx(), avs_rmsprop_step(ctx(), reduced, lr, l1), "avs_rmsprop_step");
        // the update ran on the device copy; fetch it at once (3.3 KB) so that the host copy of THIS handle stays the
        // authoritative one: a later setParams(), or another handle re-binding the shared context, can then never
        // lose or shadow the step
        check(ctx(), avs_get_nn_params(ctx(), m_params.data()), "avs_get_nn_params");
    }
    avs_ctx* context() { bind(); return ctx(); }

private:
    static double yawOf(double qw, double qx, double qy, double qz) {  // toEulerAngles, yaw part (reference utils.cpp:75-77)
        return std::atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
    }
    avs_ctx* ctx() { return m_shared->ctx; }
    // push this system's parameters / terrain into the shared context if another handle touched it
    void bind() {
        if (m_dirty || m_shared->params_owner != m_id) {
            if (TIRE == AVS_TIRE_NN) check(ctx(), avs_set_nn_params(ctx(), m_params.data()), "avs_set_nn_params");
            else check(ctx(), avs_set_bekker_params(ctx(), m_params.data()), "avs_set_bekker_params");
            m_dirty = false;
            m_shared->params_owner = m_id;
        }
        if (m_shared->terrain_owner != m_id) {
            const TerrainMap<Scalar>* map = m_map.get();
            const avs::TerrainKind k = map ? map->deviceKind() : avs::TERRAIN_FLAT;
            if (k == avs::TERRAIN_GRID) {
                const avs::TerrainGrid* g = map->deviceGrid();
                check(ctx(), avs_set_terrain_grid(ctx(), g->height.data(), g->soil.empty() ? nullptr : g->soil.data(), g->nx, g->ny, g->x0, g->y0, g->dx, g->dy),
                      "avs_set_terrain_grid");
            } else if (k == avs::TERRAIN_CUSTOM) {
                // user-defined analytic map: sample it (1024 x 1024 cells of 2 cm around the origin, AVSL_GRID_* to override)
                const int n = envInt("AVSL_GRID_N", 1024);
                const double cell = envDouble("AVSL_GRID_CELL", 0.02);
                avs::TerrainGrid g;
                g.nx = g.ny = n; g.dx = g.dy = cell; g.x0 = g.y0 = -0.5 * (n - 1) * cell;
                g.height.resize((size_t)n * n);
                for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) g.height[(size_t)j * n + i] = map->getAltitude(g.x0 + i * cell, g.y0 + j * cell);
                check(ctx(), avs_set_terrain_grid(ctx(), g.height.data(), nullptr, n, n, g.x0, g.y0, cell, cell), "avs_set_terrain_grid");
            } else {
                check(ctx(), avs_set_terrain_analytic(ctx(), (int)k), "avs_set_terrain_analytic");
            }
            m_shared->terrain_owner = m_id;
        }
    }
    static int envInt(const char* k, int d) { const char* v = std::getenv(k); return v ? std::atoi(v) : d; }
    static double envDouble(const char* k, double d) { const char* v = std::getenv(k); return v ? std::atof(v) : d; }

    std::shared_ptr<const TerrainMap<Scalar>> m_map;
    std::shared_ptr<SharedContext> m_shared;
    std::vector<double> m_params;
    const unsigned long long m_id;
    bool m_dirty = true;  // host copy newer than the context
};

}  // namespace avs
