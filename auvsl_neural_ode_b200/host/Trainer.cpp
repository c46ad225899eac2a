// Trainer.cpp — see Trainer.h.  Reference citations are to /root/reference/src/Trainer.cpp.
#include "Trainer.h"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>

#include "../../include/avsim.h"
#include "TestTerrainMaps.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace {
std::string envOr(const char* key, const char* fallback) {
    const char* v = std::getenv(key);
    return v ? std::string(v) : std::string(fallback);
}
}  // namespace

Trainer::Trainer(std::shared_ptr<SystemFactory<ADF>> factory, int num_threads) : m_factory_adf{factory}, m_num_threads{num_threads} {
    m_system_adf = m_factory_adf->makeSystem();
    const int np = m_system_adf->getNumParams();
    m_params = VectorAD::Zero(np);
    m_batch_grad = VectorF::Zero(np);
    m_squared_grad = VectorAD::Ones(np);  // :63
    m_system_adf->getDefaultParams(m_params);
    m_system_adf->setParams(m_params);
    computeEqState();
    m_cnt = 0;
    m_batch_loss = 0;
    m_param_file = envOr("AVSL_PARAM_FILE", "/home/justin/tire.net");              // :70
    m_best_param_file = envOr("AVSL_BEST_PARAM_FILE", "/home/justin/best_tire.net");  // :261
    m_data_dir = envOr("AVSL_DATA_DIR", "/home/justin/code/auvsl_dynamics_bptt/scripts/");
    if (!m_data_dir.empty() && m_data_dir.back() != '/') m_data_dir += '/';
    m_best_CV3 = .096;  // :72
    const char* g = std::getenv("AVSL_GPUS");
    m_gpus = g ? std::max(1, std::atoi(g)) : 1;
}

Trainer::~Trainer() {}

// :80-100
void Trainer::computeEqState() {
    VectorAD state = VectorAD::Zero(m_system_adf->getStateDim());
    m_system_adf->getDefaultInitialState(state);
    m_quat_stable = VectorAD::Zero(4);
    for (int i = 0; i < 4; i++) m_quat_stable[i] = state[i];
    m_z_stable = state[6];
    std::cout << "Stable Quaternion: " << m_quat_stable[0] << ", " << m_quat_stable[1] << ", " << m_quat_stable[2] << ", " << m_quat_stable[3] << "\n";
    std::cout << "z_stable: " << m_z_stable << "\n";
}

// CSV schema idx,time,vel_left,vel_right,x,y,yaw,wx,wy,wz,vx,vy  (:102-143; scripts/preprocess.py:100-119)
void Trainer::loadDataFile(std::string fn) {
    std::cout << "Opening " << fn << "\n";
    std::cout.flush();
    std::ifstream f(fn);
    m_data.clear();
    if (!f.is_open()) {
        std::cout << "File was not open\n";
        return;
    }
    std::string line;
    std::getline(f, line);  // column headings
    bool last_row_terminated = false;
    while (std::getline(f, line)) {
        const bool terminated = !f.eof();  // the line ended in '\n' (and not at the end of the file)
        if (line.empty() || line.find_first_not_of(" \t\r") == std::string::npos) continue;
        for (char& c : line) if (c == ',') c = ' ';
        std::istringstream ss(line);
        double idx, wx, wy;
        GroundTruthDataRow r;
        if (!(ss >> idx >> r.time >> r.vl >> r.vr >> r.x >> r.y >> r.yaw >> wx >> wy >> r.wz >> r.vx >> r.vy)) continue;
        r.z = m_z_stable;  // :138
        m_data.push_back(r);
        last_row_terminated = terminated;
    }
    // The reference loops on `while (data_file.peek() != EOF)` (:123): after the last row of a file that ends in a newline
    // peek() still sees '\n', every extraction of the extra pass fails and leaves `row` untouched, and the LAST ROW IS PUSHED
    // A SECOND TIME.  With the strict `j < size - window` bound of the window loops (:163, :353) that extra row decides
    // whether a file of exactly k * stride + window rows yields its last window, so it is reproduced here.
    if (last_row_terminated && !m_data.empty()) m_data.push_back(m_data.back());
}

std::string Trainer::dataFile(const char* stem, int idx) const {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%s%02d.csv", stem, idx);
    return m_data_dir + buf;
}

std::vector<std::vector<GroundTruthDataRow>> Trainer::windowsOf(int window, int stride) const {
    std::vector<std::vector<GroundTruthDataRow>> w;
    // for (j = 0; j < size - window; j += stride)   (:163, :353)
    for (long j = 0; j < (long)m_data.size() - window; j += stride) w.emplace_back(m_data.begin() + j, m_data.begin() + j + window);
    return w;
}

void Trainer::PinnedBuf::assignZero(size_t n) {
    release();
    if (n == 0) return;
    void* p = nullptr;
    if (avs_alloc_pinned(&p, (unsigned long long)(n * sizeof(double))) == AVS_OK) {
        pinned_ = true;
    } else {  // still host memory, only pageable: the copy engine stages it (not a compute fallback)
        p = std::malloc(n * sizeof(double));
        if (!p) throw std::bad_alloc();
        pinned_ = false;
    }
    p_ = static_cast<double*>(p);
    n_ = n;
    std::memset(p_, 0, n * sizeof(double));
}
void Trainer::PinnedBuf::release() {
    if (!p_) return;
    if (pinned_) avs_free_pinned(p_); else std::free(p_);
    p_ = nullptr; n_ = 0;
}

// windows [begin, end) -> SoA buffers of include/avsim.h, in page-locked memory
Trainer::Packed Trainer::pack(const std::vector<std::vector<GroundTruthDataRow>>& windows, size_t begin, size_t end) {
    Packed p;
    p.N = (int)(end - begin);
    if (p.N <= 0) { p.N = 0; return p; }
    p.T = (int)windows[begin].size();
    const size_t N = p.N;
    p.x0.assignZero(21 * N);
    p.ctrl.assignZero((size_t)(p.T - 1) * 2 * N);
    p.gt.assignZero((size_t)p.T * 3 * N);
    for (size_t n = 0; n < N; n++) {
        const auto& tr = windows[begin + n];
        VectorAD xk = m_system_adf->initializeState(tr[0]);  // :610
        for (int c = 0; c < 21; c++) p.x0[c * N + n] = xk[c];
        for (int i = 0; i < p.T; i++) {
            if (i < p.T - 1) { p.ctrl[((size_t)i * 2) * N + n] = tr[i].vl; p.ctrl[((size_t)i * 2 + 1) * N + n] = tr[i].vr; }  // :619-620
            p.gt[((size_t)i * 3) * N + n] = tr[i].yaw; p.gt[((size_t)i * 3 + 1) * N + n] = tr[i].x; p.gt[((size_t)i * 3 + 2) * N + n] = tr[i].y;
        }
    }
    return p;
}

// :550-589  loss = sqrt(lin_mse_final) / traj_len
void Trainer::evaluateTrajectory(const std::vector<GroundTruthDataRow>& traj, std::vector<VectorAD>& x_list, double& loss) {
    m_system_adf->setParams(m_params);
    Packed p = pack({traj});
    const int T = p.T;
    std::vector<double> xl((size_t)T * 23);
    m_system_adf->rolloutBatch(1, T, p.x0.data(), p.ctrl.data(), xl.data(), nullptr);
    x_list.resize(T);
    for (int i = 0; i < T; i++) {
        x_list[i] = VectorAD(23);
        for (int c = 0; c < 23; c++) x_list[i][c] = xl[(size_t)i * 23 + c];
    }
    double traj_len = 0;
    for (int i = 1; i < T; i++) traj_len += std::hypot(traj[i].x - traj[i - 1].x, traj[i].y - traj[i - 1].y);
    VectorAD gt_vec = VectorAD::Zero(m_system_adf->getStateDim());
    gt_vec[3] = traj[T - 1].yaw; gt_vec[4] = traj[T - 1].x; gt_vec[5] = traj[T - 1].y;
    ADF ang_mse, lin_mse;
    m_system_adf->evaluate(gt_vec, x_list.back(), ang_mse, lin_mse);
    loss = std::sqrt(lin_mse) / traj_len;
}

// :591-659  L = (sum_i l_i) / T / traj_len and dL/dparams
void Trainer::trainTrajectory(const std::vector<GroundTruthDataRow>& traj, std::vector<VectorAD>& x_list, VectorF& gradient, double& loss) {
    std::shared_ptr<System<ADF>> sys = m_factory_adf->makeSystem();  // :596 (a light handle onto the shared context)
    sys->setParams(m_params);
    Packed p = pack({traj});
    const int np = sys->getNumParams();
    std::vector<double> red(np + 2), gps(np);
    sys->rolloutGradBatch(1, p.T, p.x0.data(), p.ctrl.data(), p.gt.data(), 1e300, 0, &loss, red.data(), gps.data());
    gradient = VectorF(np);
    for (int k = 0; k < np; k++) gradient[k] = gps[k];
    if (!x_list.empty()) {
        std::vector<double> xl((size_t)p.T * 23);
        sys->rolloutBatch(1, p.T, p.x0.data(), p.ctrl.data(), xl.data(), nullptr);
        x_list.resize(p.T);
        for (int i = 0; i < p.T; i++) {
            x_list[i] = VectorAD(23);
            for (int c = 0; c < 23; c++) x_list[i][c] = xl[(size_t)i * 23 + c];
        }
    }
}

// :145-202  sequential by construction: the parameters change after every window
void Trainer::train() {
    VectorF traj_grad(m_system_adf->getNumParams());
    double loss, avg_loss = 0;
    std::vector<VectorAD> x_list;  // the reference fills x_list but only uses it for (disabled) plots
    int cnt_actual = 0;
    for (int i = 1; i <= 16; i++) {
        loadDataFile(dataFile("Train3_data", i));
        for (long j = 0; j < (long)m_data.size() - m_train_steps; j += m_inc_train_steps) {
            std::vector<GroundTruthDataRow> traj(m_data.begin() + j, m_data.begin() + j + m_train_steps);
            trainTrajectory(traj, x_list, traj_grad, loss);
            for (size_t k = 0; k < m_params.size(); k++) {
                if (std::fabs(traj_grad[k]) > 100.0) {  // :169-179
                    std::cout << "Explosion " << k << ":" << traj_grad[k] << "\n";
                    traj_grad[k] = 0.0;
                }
            }
            m_batch_grad += traj_grad;
            avg_loss += loss;
            std::cout << "Loss: " << loss << "\tdParams: " << traj_grad[0] << "\n";
            std::cout.flush();
            m_cnt++;
            cnt_actual++;
            updateParams(m_batch_grad / m_cnt);
            m_cnt = 0;
        }
        save();
    }
    std::cout << "Average Loss: " << avg_loss / cnt_actual << "\n";
}

// ---- multi-GPU: one context + one host thread per GPU, ONE all-reduce per step behind the C ABI -----------------------
struct Trainer::MultiGpu {
    std::vector<avs_ctx*> ctx;
    ~MultiGpu() { for (avs_ctx* c : ctx) avs_destroy(c); }
};

void Trainer::setNumGpus(int g) {
    if (g < 1) g = 1;
    if (g != m_gpus) m_multi.reset();
    m_gpus = g;
}

static void ckAvs(avs_ctx* c, int rc, const char* what) {
    if (rc != AVS_OK) throw std::runtime_error(std::string(what) + ": " + avs_strerror(rc) + " — " + (c ? avs_last_error(c) : "no context"));
}

double Trainer::trainBatchMultiGpu(const std::vector<std::vector<GroundTruthDataRow>>& windows, int explode_mode) {
    const int G = m_gpus;
    const int np = m_system_adf->getNumParams();
    if (np != AVS_NUM_NN_PARAMS) throw std::runtime_error("Trainer: multi-GPU training exists for the NN tire model only");
    if (!m_multi) {
        auto mg = std::make_shared<MultiGpu>();
        for (int g = 0; g < G; g++) {
            avs_ctx* c = nullptr;
            ckAvs(nullptr, avs_create(&c, g, AVS_TIRE_NN), "avs_create");
            mg->ctx.push_back(c);
        }
        ckAvs(mg->ctx[0], avs_comm_init_all(mg->ctx.data(), G), "avs_comm_init_all");
        // terrain: the factory's map, as the single-GPU systems configure it
        const auto& map = m_factory_adf->terrainMap();
        const avs::TerrainKind k = map ? map->deviceKind() : avs::TERRAIN_FLAT;
        for (avs_ctx* c : mg->ctx) {
            if (k == avs::TERRAIN_GRID) {
                const avs::TerrainGrid* g = map->deviceGrid();
                ckAvs(c, avs_set_terrain_grid(c, g->height.data(), g->soil.empty() ? nullptr : g->soil.data(), g->nx, g->ny, g->x0, g->y0, g->dx, g->dy), "avs_set_terrain_grid");
            } else if (k == avs::TERRAIN_CUSTOM) {
                throw std::runtime_error("Trainer: multi-GPU training needs a Flat/Slope/Bumpy/Grid terrain map");
            } else {
                ckAvs(c, avs_set_terrain_analytic(c, (int)k), "avs_set_terrain_analytic");
            }
        }
        m_multi = mg;
    }
    const size_t N = windows.size();
    std::vector<double> params(np);
    for (int k = 0; k < np; k++) params[k] = m_params[k];
    std::vector<std::vector<double>> red(G, std::vector<double>(AVS_REDUCED_LEN, 0.0));
    std::vector<std::string> errs(G);
    std::vector<Packed> shard(G);
    for (int g = 0; g < G; g++) {  // contiguous shards, sizes differ by at most one (packing touches the shared System: serial)
        const size_t base = N / G, rem = N % G;
        const size_t b = g * base + std::min<size_t>(g, rem), e = b + base + ((size_t)g < rem ? 1 : 0);
        shard[g] = pack(windows, b, e);
    }
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g]() {
            try {
                avs_ctx* c = m_multi->ctx[g];
                ckAvs(c, avs_set_nn_params(c, params.data()), "avs_set_nn_params");
                Packed& p = shard[g];
                std::vector<double> loss(std::max(p.N, 1));
                if (p.N > 0) {
                    ckAvs(c, avs_rollout_grad(c, p.N, p.T, 10, p.x0.data(), p.ctrl.data(), p.gt.data(), 100.0, explode_mode, loss.data(), red[g].data(), nullptr),
                          "avs_rollout_grad");
                } else {  // a rank without windows still joins the collective with zeros
                    ckAvs(c, avs_set_reduced(c, red[g].data()), "avs_set_reduced");
                }
                ckAvs(c, avs_allreduce(c, red[g].data()), "avs_allreduce");
                // identical update on every replica (device RMSProp on the all-reduced buffer)
                ckAvs(c, avs_rmsprop_step_dev(c, nullptr, m_lr, m_l1_weight), "avs_rmsprop_step_dev");
                ckAvs(c, avs_sync(c), "avs_sync");
            } catch (const std::exception& e) {
                errs[g] = e.what();
            }
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++) if (!errs[g].empty()) throw std::runtime_error("GPU " + std::to_string(g) + ": " + errs[g]);
    m_cnt = (int)red[0][np + 1];
    m_batch_loss = red[0][np];
    for (int k = 0; k < np; k++) m_batch_grad[k] = red[0][k];
    return m_cnt ? m_batch_loss / m_cnt : 0.0;
}

// one batched step over explicit windows (the GPU-shaped replacement of the worker pool, :204-305)
double Trainer::trainBatch(const std::vector<std::vector<GroundTruthDataRow>>& windows, int explode_mode) {
    if (m_gpus > 1) return trainBatchMultiGpu(windows, explode_mode);
    std::shared_ptr<System<ADF>> sys = m_system_adf;
    sys->setParams(m_params);
    Packed p = pack(windows);
    if (p.N == 0) return 0.0;
    const int np = sys->getNumParams();
    std::vector<double> loss(p.N), red(np + 2);
    sys->rolloutGradBatch(p.N, p.T, p.x0.data(), p.ctrl.data(), p.gt.data(), 100.0, explode_mode, loss.data(), red.data(), nullptr);
    m_cnt = (int)red[np + 1];
    m_batch_loss = red[np];
    for (int k = 0; k < np; k++) m_batch_grad[k] = red[k];
    return m_cnt ? m_batch_loss / m_cnt : 0.0;
}

// :204-263  whole training set = one batch, exploding samples dropped, one update
void Trainer::trainThreads() {
    std::vector<std::vector<GroundTruthDataRow>> all;
    for (int i = 1; i <= 16; i++) {
        loadDataFile(dataFile("Train3_data", i));
        auto w = windowsOf(m_train_steps, m_inc_train_steps);
        all.insert(all.end(), w.begin(), w.end());
    }
    const double avg = trainBatch(all, 1);
    std::cout << "Avg Loss: " << avg << "\n";
    std::cout.flush();
    save();
    if (m_cnt > 0) {
        if (m_gpus > 1) adoptDeviceUpdate();
        else updateParams(m_batch_grad / m_cnt);
    }
    if (avg < m_best_CV3 && m_cnt > 0) {
        m_best_CV3 = avg;
        std::cout << "dabest " << m_best_CV3 << "\n";
        saveVec(m_params, m_best_param_file);
    }
}

// :661-689
bool Trainer::combineResults(VectorF& batch_grad, const VectorF& sample_grad, double& batch_loss, const double& sample_loss) {
    for (size_t i = 0; i < m_params.size(); i++) {
        if (std::fabs(sample_grad[i]) > 100.0) {
            std::cout << "Explosion " << i << ":" << sample_grad[i] << "\n";
            return false;
        }
    }
    batch_grad += sample_grad;
    batch_loss += sample_loss;
    m_cnt++;
    std::cout << "Loss: " << sample_loss << "\tdParams: " << sample_grad[0] << "\n";
    std::cout.flush();
    return true;
}

// :428-453  RMSProp, gradient truncated to float before use
void Trainer::updateParams(const VectorF& grad) {
    double norm = 0;
    for (size_t i = 0; i < m_params.size(); i++) {
        const float g = (float)grad[i];
        m_squared_grad[i] = 0.95 * m_squared_grad[i] + 0.05 * (double)(g * g);
        m_params[i] -= (m_lr / (std::sqrt(m_squared_grad[i]) + 1e-6)) * (double)g;
        m_params[i] -= m_lr * m_l1_weight * m_params[i];
        norm += std::fabs(grad[i]);
    }
    const double update0 = (m_lr / (std::sqrt(m_squared_grad[0]) + 1e-6)) * grad[0];
    std::cout << "Gradient norm: " << norm << " Param[0]: " << m_params[0] << " grad[0]: " << grad[0] << " update[0]: " << update0
              << " squared_grad[0]: " << std::sqrt(m_squared_grad[0]) << "\n";
    m_batch_grad = VectorF::Zero(m_system_adf->getNumParams());
    m_system_adf->setParams(m_params);
}

// multi-GPU: every replica has already applied updateParams' RMSProp on its device copy (bit-identical across replicas);
// fetch the result from replica 0 and print the reference's log line
void Trainer::adoptDeviceUpdate() {
    const int np = (int)m_params.size();
    std::vector<double> p(np);
    ckAvs(m_multi->ctx[0], avs_get_nn_params(m_multi->ctx[0], p.data()), "avs_get_nn_params");
    double norm = 0;
    for (int i = 0; i < np; i++) {
        const double g = m_batch_grad[i] / m_cnt;
        const float gf = (float)g;
        m_squared_grad[i] = 0.95 * m_squared_grad[i] + 0.05 * (double)(gf * gf);
        norm += std::fabs(g);
        m_params[i] = p[i];
    }
    const double g0 = m_batch_grad[0] / m_cnt;
    std::cout << "Gradient norm: " << norm << " Param[0]: " << m_params[0] << " grad[0]: " << g0 << " update[0]: "
              << (m_lr / (std::sqrt(m_squared_grad[0]) + 1e-6)) * g0 << " squared_grad[0]: " << std::sqrt(m_squared_grad[0]) << "\n";
    m_batch_grad = VectorF::Zero(np);
    m_system_adf->setParams(m_params);
}

std::vector<double> Trainer::evaluateWindows(int window, int stride) {
    m_system_adf->setParams(m_params);
    auto w = windowsOf(window, stride);
    Packed p = pack(w);
    std::vector<double> out(p.N);
    if (p.N == 0) return out;
    std::vector<double> xf((size_t)21 * p.N);
    m_system_adf->rolloutBatch(p.N, p.T, p.x0.data(), p.ctrl.data(), nullptr, xf.data());
    for (int n = 0; n < p.N; n++) {
        double traj_len = 0;
        for (int i = 1; i < p.T; i++) traj_len += std::hypot(w[n][i].x - w[n][i - 1].x, w[n][i].y - w[n][i - 1].y);
        const double ex = w[n][p.T - 1].x - xf[(size_t)4 * p.N + n], ey = w[n][p.T - 1].y - xf[(size_t)5 * p.N + n];
        out[n] = std::sqrt(ex * ex + ey * ey) / traj_len;  // :587-588
    }
    return out;
}

void Trainer::evaluateSet(const char* stem, int first, int last, const char* label, bool per_window_lines) {
    double loss_avg = 0;
    int cnt = 0;
    for (int i = first; i <= last; i++) {
        loadDataFile(dataFile(stem, i));
        std::vector<double> l = evaluateWindows(m_eval_steps, m_inc_eval_steps);
        for (size_t k = 0; k < l.size(); k++) {
            if (per_window_lines) std::cout << label << " " << k * m_inc_eval_steps << " loss: " << l[k] << "\n;";  // :394
            loss_avg += l[k];
            cnt++;
        }
    }
    std::cout << label << " avg loss: " << loss_avg / cnt << "\n";
}

void Trainer::evaluate_train3() { evaluateSet("Train3_data", 1, 17, "Train3", false); }  // :333-365
void Trainer::evaluate_cv3() { evaluateSet("CV3_data", 1, 144, "CV3", true); }           // :368-399
void Trainer::evaluate_ld3() { evaluateSet("LD3_data", 1, 1, "LD3", false); }            // :401-426

// :307-331 (prints the loss of the LAST window, as the reference does)
void Trainer::evaluate_validation_dataset() {
    loadDataFile(dataFile("Train3_data", 17));
    std::vector<double> l = evaluateWindows(m_eval_steps, m_inc_eval_steps);
    std::cout << "Train3_17 validation loss: " << (l.empty() ? 0.0 : l.back()) << "\n";
}

// :744-776
void Trainer::evaluate_file(const std::string name, const int max_num_segments, std::vector<VectorAD>& sim_traj, std::vector<GroundTruthDataRow>& gt_traj) {
    double loss = 0;
    loadDataFile(name);
    for (long j = 0; j < (long)m_data.size() - m_eval_steps; j += m_inc_eval_steps) {
        std::vector<VectorAD> x_list(m_eval_steps);
        std::vector<GroundTruthDataRow> traj(m_data.begin() + j, m_data.begin() + j + m_eval_steps);
        evaluateTrajectory(traj, x_list, loss);
        sim_traj.insert(sim_traj.end(), x_list.begin(), x_list.end());
        gt_traj.insert(gt_traj.end(), traj.begin(), traj.end());
        std::cout << "Name: " << j << " Loss: " << loss << "\n";
        if (j > ((max_num_segments - 1) * m_eval_steps)) return;
    }
}

void Trainer::save() { saveVec(m_params, m_param_file); }

// one line "v,v,...,\n" (:695-710); full precision instead of the reference's 6 significant digits
bool Trainer::saveVec(const VectorAD& params, const std::string& file_name) {
    std::ofstream f(file_name);
    if (!f.is_open()) return true;
    f << std::setprecision(17);
    for (size_t i = 0; i < params.size(); i++) f << params[i] << ",";
    f << "\n";
    return false;
}

void Trainer::load() {  // :712-724
    if (!loadVec(m_params, m_param_file)) {
        m_system_adf->setParams(m_params);
    } else {
        std::cout << "No param save file found\n";
        m_system_adf->getDefaultParams(m_params);
        m_system_adf->setParams(m_params);
    }
}

bool Trainer::loadVec(VectorAD& params, const std::string& file_name) {  // :726-742
    std::ifstream f(file_name);
    if (!f.is_open()) return true;
    char comma;
    for (size_t i = 0; i < params.size(); i++) {
        f >> params[i];
        f >> comma;
    }
    return false;
}
