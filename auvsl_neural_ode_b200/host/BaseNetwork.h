// BaseNetwork.h — API surface of the reference's base-link residual network (reference src/BaseNetwork.{h,cpp}).
// In the reference this 3-8-8-3 MLP with biases (131 parameters) is dead code: its forward call is commented out
// (src/auvsl_dynamics/HybridDynamics.cpp:433) and get_base_f_ext is never invoked (:557).  It is therefore NOT on
// the rollout kernels; this class keeps the type, parameter packing and forward semantics available on the host, and
// forwardBatch() evaluates the same function for N inputs at once on the GPU through the C ABI (avs_mlp_forward).
// LegacyTireNet below is the loader + batched evaluation of the orphan data/nn_pretrained.csv network.
#pragma once
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/avsim.h"
#include "types/VecX.h"

class BaseNetwork {
public:
    typedef avs::VecX<double> VectorS;
    static const int m_num_hidden_units = 8, m_num_inputs = 3, m_num_outputs = 3;

    BaseNetwork() {
        // 0.001 * uniform(-1, 1) initialisation, scalers out = 10, in = 1/10 (reference BaseNetwork.cpp:14-27)
        auto r = []() { return 0.001 * (2.0 * std::rand() / RAND_MAX - 1.0); };
        for (auto& row : m_weight0) for (double& w : row) w = r();
        for (double& b : m_bias0) b = r();
        for (auto& row : m_weight1) for (double& w : row) w = r();
        for (double& b : m_bias1) b = r();
        for (auto& row : m_weight2) for (double& w : row) w = r();
        for (double& b : m_bias2) b = r();
    }
    int getNumInputs() const { return m_num_inputs; }
    int getNumOutputs() const { return m_num_outputs; }
    int getNumParams() { return 8 * 3 + 8 + 8 * 8 + 8 + 3 * 8 + 3; }

    // in: (wz, vx, vy) -> out: (nz, fx, fy) = relu(net) * (-in) * 10   (reference BaseNetwork.cpp:41-60)
    void forward(const double in_vec[3], double out_vec[3]) const {
        double s[3], l0[8], l1[8];
        for (int i = 0; i < 3; i++) s[i] = in_vec[i] * 0.1;
        for (int j = 0; j < 8; j++) {
            double a = m_bias0[j];
            for (int k = 0; k < 3; k++) a += m_weight0[j][k] * s[k];
            l0[j] = std::tanh(a);
        }
        for (int j = 0; j < 8; j++) {
            double a = m_bias1[j];
            for (int k = 0; k < 8; k++) a += m_weight1[j][k] * l0[k];
            l1[j] = std::tanh(a);
        }
        for (int j = 0; j < 3; j++) {
            double a = m_bias2[j];
            for (int k = 0; k < 8; k++) a += m_weight2[j][k] * l1[k];
            out_vec[j] = (a > 0 ? a : 0.0) * (-in_vec[j]) * 10.0;
        }
    }
    // pack order W0, b0, W1, b1, W2, b2, row-major (reference BaseNetwork.cpp:62-152)
    void setParams(const VectorS& p, int idx) { visit([&](double& w) { w = p[idx++]; }); }
    void getParams(VectorS& p, int idx) { visit([&](double& w) { p[idx++] = w; }); }

    // N evaluations on the GPU: in [3][N] = (wz, vx, vy) rows, out [3][N] = (nz, fx, fy) rows; ctx = any libavsim context
    void forwardBatch(avs_ctx* ctx, int N, const double* in, double* out) {
        std::vector<double> p;
        p.reserve(AVS_NUM_BASE_NETWORK_PARAMS);
        visit([&](double& w) { p.push_back(w); });
        const int rc = avs_mlp_forward(ctx, AVS_MLP_BASE_NETWORK, p.data(), N, in, out);
        if (rc != AVS_OK) throw std::runtime_error(std::string("BaseNetwork::forwardBatch: ") + avs_strerror(rc) + " — " + avs_last_error(ctx));
    }

    double m_weight0[8][3], m_bias0[8], m_weight1[8][8], m_bias1[8], m_weight2[3][8], m_bias2[3];

private:
    template <class F> void visit(F f) {
        for (auto& row : m_weight0) for (double& w : row) f(w);
        for (double& b : m_bias0) f(b);
        for (auto& row : m_weight1) for (double& w : row) f(w);
        for (double& b : m_bias1) f(b);
        for (auto& row : m_weight2) for (double& w : row) f(w);
        for (double& b : m_bias2) f(b);
    }
};

// data/nn_pretrained.csv: 5-16-16-3 network with biases + output / input scalers (435 numbers; no reference code reads it)
class LegacyTireNet {
public:
    explicit LegacyTireNet(const std::string& csv_path) : m_params(AVS_NUM_LEGACY_NET_PARAMS) {
        if (avs_load_legacy_net_csv(csv_path.c_str(), m_params.data()) != AVS_OK)
            throw std::runtime_error("LegacyTireNet: " + csv_path + " does not hold exactly 435 numbers");
    }
    const std::vector<double>& params() const { return m_params; }
    // in [5][N] -> out [3][N] = net((in - in_mean) / in_std) * out_std + out_mean
    void forwardBatch(avs_ctx* ctx, int N, const double* in, double* out) const {
        const int rc = avs_mlp_forward(ctx, AVS_MLP_LEGACY_TIRE, m_params.data(), N, in, out);
        if (rc != AVS_OK) throw std::runtime_error(std::string("LegacyTireNet::forwardBatch: ") + avs_strerror(rc) + " — " + avs_last_error(ctx));
    }
private:
    std::vector<double> m_params;
};
