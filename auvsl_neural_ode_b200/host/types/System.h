// types/System.h — the ODE + loss interface Trainer talks to (reference src/types/System.h:8-37), extended
// with BATCHED entry points.  The scalar virtuals keep the reference's names, argument meaning and 23-vector
// convention (21 states + 2 controls); the batched ones are what the B200 path is built for: N independent
// windows per call, SoA [component][N] buffers exactly as in include/avsim.h.  Their default implementations
// loop over the scalar virtuals, so a third-party System keeps working (slowly) under the batched Trainer.
#pragma once
#include <stdexcept>
#include "types/GroundTruthDataRow.h"
#include "types/Scalars.h"
#include "types/VecX.h"

template <typename Scalar> class System {
public:
    typedef avs::VecX<Scalar> VectorS;
    typedef avs::VecX<Scalar> MatrixS;
    virtual ~System() = default;

    int getControlDim() { return m_num_controls; }
    int getStateDim() { return m_num_states; }
    int getNumParams() { return m_num_params; }

    virtual void getDefaultParams(VectorS& params) = 0;
    virtual void getDefaultInitialState(VectorS& state) = 0;
    virtual void setParams(const VectorS& params) = 0;
    virtual void getParams(VectorS& params) = 0;
    virtual void forward(const VectorS& X, VectorS& Xd) = 0;
    virtual Scalar loss(const VectorS& gt_vec, VectorS& vec) = 0;
    virtual void evaluate(const VectorS& gt_vec, const VectorS& vec, Scalar& ang_err, Scalar& lin_err) = 0;
    virtual void integrate(const VectorS& X0, VectorS& X1) = 0;
    virtual VectorS initializeState(const GroundTruthDataRow& gt_state) = 0;

    // ---- batched extensions (SoA, fp64) --------------------------------------------------------------
    // x0 [21][N], ctrl [T-1][2][N] -> x_list [T][23][N] (may be null), x_final [21][N] (may be null)
    virtual void rolloutBatch(int N, int T, const double* x0, const double* ctrl, double* x_list, double* x_final) {
        const int sd = getStateDim(), cd = getControlDim();
        VectorS xk(sd + cd), xk1(sd + cd);
        for (int n = 0; n < N; n++) {
            for (int c = 0; c < sd; c++) xk[c] = x0[(size_t)c * N + n];
            xk[sd] = T > 1 ? ctrl[n] : 0.0; xk[sd + 1] = T > 1 ? ctrl[(size_t)N + n] : 0.0;
            if (x_list) for (int c = 0; c < sd + cd; c++) x_list[(size_t)c * N + n] = xk[c];
            for (int i = 1; i < T; i++) {
                xk[sd] = ctrl[((size_t)(i - 1) * 2) * N + n]; xk[sd + 1] = ctrl[((size_t)(i - 1) * 2 + 1) * N + n];
                integrate(xk, xk1);
                xk = xk1;
                if (x_list) for (int c = 0; c < sd + cd; c++) x_list[((size_t)i * (sd + cd) + c) * N + n] = xk[c];
            }
            if (x_final) for (int c = 0; c < sd; c++) x_final[(size_t)c * N + n] = xk[c];
        }
    }
    // gt [T][3][N] -> loss [N], reduced [numParams + 2] (grad_sum | loss_sum | n_kept), grad_per_sample
    // [numParams][N] (may be null).  A System without an adjoint reports that it cannot train.
    virtual void rolloutGradBatch(int, int, const double*, const double*, const double*, double, int, double*, double*, double*) {
        throw std::runtime_error("System::rolloutGradBatch: this system has no adjoint");
    }
    // RMSProp of Trainer::updateParams on the system's own parameters; reduced as produced above
    virtual void applyRmsProp(const double*, double, double) { throw std::runtime_error("System::applyRmsProp: not supported"); }

protected:
    void setControlDim(int n) { m_num_controls = n; }
    void setStateDim(int n) { m_num_states = n; }
    void setNumParams(int n) { m_num_params = n; }

private:
    int m_num_controls = 0, m_num_states = 0, m_num_params = 0;
};
