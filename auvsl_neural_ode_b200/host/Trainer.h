// Trainer.h — training / evaluation driver with the reference's public surface (reference src/Trainer.h:13-112),
// re-architected around batches: every window of a file (or of all files) is gathered into ONE call of the
// System's batched entry points instead of being rolled out one after another on CppAD scalars.
//
// Kept from the reference: method names and meaning, window sizes (200 / 600 rows, src/Trainer.h:90-93), the
// loss and gradient definitions (src/Trainer.cpp:550-659), both explosion-filter flavours (:169-179, :661-689),
// the RMSProp update with the float-truncated gradient (:428-453), the parameter file format (:695-742) and the
// stdout line formats that scripts/plot_loss.py parses.  Not kept: CppAD taping, the polling worker-thread pool
// (num_threads is accepted and ignored: the batch is the parallelism), matplotlib plotting.
// Paths: the reference hard-codes /home/justin/...; here AVSL_DATA_DIR / AVSL_PARAM_FILE / AVSL_BEST_PARAM_FILE
// override those defaults.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "types/GroundTruthDataRow.h"
#include "types/System.h"
#include "types/SystemFactory.h"
#include "types/TerrainMap.h"

class Trainer {
public:
    typedef avs::VecX<double> VectorF;
    typedef avs::VecX<double> MatrixF;
    typedef avs::VecX<ADF> VectorAD;
    typedef avs::VecX<ADF> MatrixAD;

    Trainer(std::shared_ptr<SystemFactory<ADF>> factory, int num_threads = 2);
    ~Trainer();

    // --- reference surface
    void updateParams(const VectorF& grad);
    void evaluateTrajectory(const std::vector<GroundTruthDataRow>& traj, std::vector<VectorAD>& x_list, double& loss);
    void trainTrajectory(const std::vector<GroundTruthDataRow>& traj, std::vector<VectorAD>& x_list, VectorF& gradient, double& loss);
    void plotTrajectory(const std::vector<GroundTruthDataRow>&, const std::vector<VectorAD>&) {}  // plotting is out of scope
    void loadDataFile(std::string fn);
    void train();          // batch size 1, zero-the-component filter, one RMSProp step per window
    void trainThreads();   // whole data set as one batch, drop-the-sample filter, one RMSProp step
    void evaluate_train3();
    void evaluate_cv3();
    void evaluate_ld3();
    void evaluate_file(const std::string name, const int max_num_segments, std::vector<VectorAD>& sim_traj, std::vector<GroundTruthDataRow>& gt_traj);
    void evaluate_validation_dataset();
    void computeEqState();
    bool saveVec(const VectorAD& params, const std::string& file_name);  // true = failure, as in the reference
    bool loadVec(VectorAD& params, const std::string& file_name);
    void save();
    void load();
    bool combineResults(VectorF& batch_grad, const VectorF& sample_grad, double& batch_loss, const double& sample_loss);

    // --- batched extensions
    // losses of all `window`-row windows (stride = window) of the loaded file, one batched rollout
    std::vector<double> evaluateWindows(int window, int stride);
    // one full-batch training step over explicit windows; returns the mean kept loss
    double trainBatch(const std::vector<std::vector<GroundTruthDataRow>>& windows, int explode_mode);
    // number of GPUs trainBatch / trainThreads shard the batch over (default: AVSL_GPUS or 1).  With G > 1 the Trainer owns
    // one context per GPU, one host thread per context, and ONE ncclAllReduce of the 418 reduced doubles per step behind
    // the C ABI (avs_allreduce); every replica then applies the same device RMSProp (reference seam: src/Trainer.cpp:204-263,
    // 661-689 — worker threads + `batch_grad += sample_grad` on the main thread)
    void setNumGpus(int g);
    int numGpus() const { return m_gpus; }
    void setDataDir(const std::string& d) { m_data_dir = d; }
    void setParamFile(const std::string& f) { m_param_file = f; }

    const int m_train_steps = 200;
    const int m_eval_steps = 600;
    const int m_inc_train_steps = 200;
    const int m_inc_eval_steps = 600;
    const double m_l1_weight = 0.0;
    const double m_lr = 1e-3;

    double m_best_CV3;
    ADF m_z_stable;
    VectorAD m_quat_stable;
    std::string m_param_file, m_best_param_file, m_data_dir;
    int m_cnt;
    std::vector<GroundTruthDataRow> m_data;
    VectorAD m_params;
    VectorAD m_squared_grad;
    VectorF m_batch_grad;
    double m_batch_loss;

    std::shared_ptr<SystemFactory<ADF>> m_factory_adf;
    std::shared_ptr<System<ADF>> m_system_adf;
    const int m_num_threads;

private:
    // page-locked staging buffer (cudaHostAlloc through avs_alloc_pinned): the SoA batches go to the GPU by DMA
    class PinnedBuf {
    public:
        PinnedBuf() = default;
        PinnedBuf(const PinnedBuf&) = delete;
        PinnedBuf& operator=(const PinnedBuf&) = delete;
        PinnedBuf(PinnedBuf&& o) noexcept : p_(o.p_), n_(o.n_), pinned_(o.pinned_) { o.p_ = nullptr; o.n_ = 0; }
        PinnedBuf& operator=(PinnedBuf&& o) noexcept { if (this != &o) { release(); p_ = o.p_; n_ = o.n_; pinned_ = o.pinned_; o.p_ = nullptr; o.n_ = 0; } return *this; }
        ~PinnedBuf() { release(); }
        void assignZero(size_t n);
        double* data() { return p_; }
        const double* data() const { return p_; }
        double& operator[](size_t i) { return p_[i]; }
        size_t size() const { return n_; }
        bool pinned() const { return pinned_; }
    private:
        void release();
        double* p_ = nullptr;
        size_t n_ = 0;
        bool pinned_ = false;
    };
    struct Packed { int N = 0, T = 0; PinnedBuf x0, ctrl, gt; };
    Packed pack(const std::vector<std::vector<GroundTruthDataRow>>& windows) { return pack(windows, 0, windows.size()); }
    Packed pack(const std::vector<std::vector<GroundTruthDataRow>>& windows, size_t begin, size_t end);
    struct MultiGpu;
    std::shared_ptr<MultiGpu> m_multi;
    int m_gpus = 1;
    void adoptDeviceUpdate();
    double trainBatchMultiGpu(const std::vector<std::vector<GroundTruthDataRow>>& windows, int explode_mode);
    std::vector<std::vector<GroundTruthDataRow>> windowsOf(int window, int stride) const;
    std::string dataFile(const char* stem, int idx) const;
    void evaluateSet(const char* stem, int first, int last, const char* label, bool per_window_lines);
};
