// avs_capi.cu — the C ABI declared in include/avsim.h: context, device memory, launches.
// No torch, no CPU fallback: every compute entry point ends in a kernel launch on the context's stream.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <nccl.h>  // types and enums only: the library is bound at run time (dlopen), libavsim.so does not link it

#include "../../include/avsim.h"
#include "avs_kernels.h"
#include "avs_weights.h"

using namespace avs;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return (T*)p; }
};

struct avs_ctx {
    int device = 0;
    int tire = AVS_TIRE_NN;
    int terr = AVS_TERRAIN_FLAT;
    unsigned long long ckpt_budget = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host-buffer entry points: H2D of the ground truth overlaps the forward sweep
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_fwd = false, ev_bwd = false;
    ModelConst mc;
    double params[AVS_NUM_NN_PARAMS];
    double* d_params = nullptr;  // [416] device copy (RMSProp operates here)
    double* d_sq = nullptr;      // [416] RMSProp state
    bool params_dirty_on_device = false;
    DevBuf grid_h, grid_s, ztab;
    DevBuf ckpt, act, gps, loss, reduced, flags, settle;
    DevBuf in_x0, in_ctrl, in_gt, out_list, out_final, io_a, io_b, io_c;
    DevBuf margins;
    ncclComm_t comm = nullptr;  // one communicator rank per context (avs_comm_*)
    int world = 1, rank = 0;
    long long launches = 0;
    int plan_N = -1;
    size_t plan_per_vehicle = 0, plan_chunk = 0;
    unsigned long long plan_budget = 0;
    std::string err;
};

static int fail(avs_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) return fail(ctx, AVS_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---- one __constant__ model per translation unit, shared by every context of a device -------------------------------
// upload_model (avs_dev.cuh) and the launch that reads it are ordered within ONE stream only.  Two contexts on the same
// GPU use different streams, so a later context's upload could overtake an earlier context's still-running kernel.  Every
// launch therefore waits for the previous launch on that device when it came from a different stream (a no-op for the
// usual one-context-per-GPU set-up) and leaves an event behind.
namespace {
struct DeviceOrder {
    std::mutex mu;
    cudaEvent_t ev = nullptr;
    cudaStream_t last = nullptr;
    bool any = false;
};
DeviceOrder g_order[64];
}  // namespace
static cudaError_t pre_launch(avs_ctx* c) {
    DeviceOrder& o = g_order[c->device & 63];
    std::lock_guard<std::mutex> lk(o.mu);
    if (o.any && o.last != c->stream) return cudaStreamWaitEvent(c->stream, o.ev, 0);
    return cudaSuccess;
}
static cudaError_t post_launch(avs_ctx* c) {
    DeviceOrder& o = g_order[c->device & 63];
    std::lock_guard<std::mutex> lk(o.mu);
    if (!o.ev) {
        cudaError_t e = cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    o.last = c->stream;
    o.any = true;
    return cudaEventRecord(o.ev, c->stream);
}
#define LAUNCH(call)            \
    do {                        \
        CK(pre_launch(ctx));    \
        CK(call);               \
        CK(post_launch(ctx));   \
    } while (0)

// ---- NCCL, bound at run time --------------------------------------------------------------------------------------
// dlopen("libnccl.so.2") returns the copy the process already maps (a Python process that imported torch: torch's bundled
// NCCL; a plain C++ host: the system library), so libavsim.so never mixes two NCCL builds in one process.
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
#define AVS_BIND(field, sym)                                                                 \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, sym));                    \
    if (!api.field && api.err.empty()) api.err = std::string("libnccl: missing symbol ") + sym;
        AVS_BIND(GetUniqueId, "ncclGetUniqueId")
        AVS_BIND(CommInitRank, "ncclCommInitRank")
        AVS_BIND(CommInitAll, "ncclCommInitAll")
        AVS_BIND(CommDestroy, "ncclCommDestroy")
        AVS_BIND(AllReduce, "ncclAllReduce")
        AVS_BIND(GetErrorString, "ncclGetErrorString")
        AVS_BIND(GetVersion, "ncclGetVersion")
#undef AVS_BIND
    });
    return &api;
}
}  // namespace
#define CKN(call)                                                                                                       \
    do {                                                                                                                \
        ncclResult_t r__ = (call);                                                                                      \
        if (r__ != ncclSuccess) return fail(ctx, AVS_E_COMM, "%s: %s (%s:%d)", #call, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

static void load_weights(avs_ctx* c) {
    memcpy(c->mc.W0, c->params + kOffW0, sizeof(c->mc.W0));
    memcpy(c->mc.W2, c->params + kOffW2, sizeof(c->mc.W2));
    memcpy(c->mc.W4, c->params + kOffW4, sizeof(c->mc.W4));
}

// after a device-side RMSProp step the authoritative parameters live in d_params
static int refresh_params(avs_ctx* ctx) {
    if (ctx->params_dirty_on_device) {
        CK(cudaMemcpyAsync(ctx->params, ctx->d_params, sizeof(ctx->params), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->params_dirty_on_device = false;
        load_weights(ctx);
    }
    return AVS_OK;
}

// device copy of the parameter vector is authoritative for the NN weights (kept in sync by avs_set_nn_params and
// updated in place by the device RMSProp), see upload_model()
static const double* live(const avs_ctx* c) { return c->tire == AVS_TIRE_NN ? c->d_params : nullptr; }

extern "C" {

const char* avs_strerror(int code) {
    switch (code) {
        case AVS_OK: return "ok";
        case AVS_E_ARG: return "invalid argument";
        case AVS_E_CUDA: return "CUDA error or no usable device";
        case AVS_E_STATE: return "call not valid in this configuration";
        case AVS_E_NOMEM: return "out of memory";
        case AVS_E_COMM: return "NCCL error or communicator not initialised";
    }
    return "unknown";
}

int avs_nn_default_params(double* p416) {
    if (!p416) return AVS_E_ARG;
    for (int k = 0; k < 4; k++) {  // TireNetwork::load_model: the same weights for all four stored networks
        memcpy(p416 + 104 * k + kOffW0, kDefaultW0, sizeof(kDefaultW0));
        memcpy(p416 + 104 * k + kOffW2, kDefaultW2, sizeof(kDefaultW2));
        memcpy(p416 + 104 * k + kOffW4, kDefaultW4, sizeof(kDefaultW4));
    }
    return AVS_OK;
}
int avs_bekker_default_params(double* p5) {
    if (!p5) return AVS_E_ARG;
    memcpy(p5, kBekkerDefault, sizeof(kBekkerDefault));
    return AVS_OK;
}

int avs_create(avs_ctx** out, int device, int tire_model) {
    if (!out || (tire_model != AVS_TIRE_NN && tire_model != AVS_TIRE_BEKKER)) return AVS_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return AVS_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AVS_E_CUDA;
    if (prop.major < 10) return AVS_E_CUDA;  // sm_100a SASS only
    if (cudaSetDevice(device) != cudaSuccess) return AVS_E_CUDA;
    avs_ctx* ctx = new avs_ctx();
    ctx->device = device;
    ctx->tire = tire_model;
    memset(&ctx->mc, 0, sizeof(ctx->mc));
    avs_nn_default_params(ctx->params);
    load_weights(ctx);
    memcpy(ctx->mc.Z0, kFixedZ0, sizeof(kFixedZ0));
    memcpy(ctx->mc.Z2, kFixedZ2, sizeof(kFixedZ2));
    memcpy(ctx->mc.Z4, kFixedZ4, sizeof(kFixedZ4));
    memcpy(ctx->mc.bek, kBekkerDefault, sizeof(kBekkerDefault));
    compute_AIinv(ctx->mc.AIinv);
    fill_contact_search(ctx->mc);
    ctx->mc.grid.inv_dx = ctx->mc.grid.inv_dy = 1.0;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->stream = ctx->own_stream;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_params, sizeof(double) * AVS_NUM_NN_PARAMS) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_sq, sizeof(double) * AVS_NUM_NN_PARAMS) == cudaSuccess;
    if (ok) {  // constant-fold the fixed z-network into its table
        std::vector<double> tab((size_t)kZTabRows * kZTabRow);
        build_ztab(ctx->mc.Z0, ctx->mc.Z2, ctx->mc.Z4, tab.data());
        ok = ctx->ztab.reserve(tab.size() * sizeof(double)) == cudaSuccess &&
             cudaMemcpy(ctx->ztab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
        ctx->mc.ztab = ctx->ztab.as<double>();
    }
    if (!ok) { avs_destroy(ctx); return AVS_E_CUDA; }
    if (cudaMemcpy(ctx->d_params, ctx->params, sizeof(ctx->params), cudaMemcpyHostToDevice) != cudaSuccess) { avs_destroy(ctx); return AVS_E_CUDA; }
    *out = ctx;
    return avs_reset_optimizer(ctx);
}

void avs_destroy(avs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(ctx->comm);
    ctx->margins.release();
    DevBuf* bufs[] = {&ctx->grid_h, &ctx->grid_s, &ctx->ztab, &ctx->ckpt, &ctx->act, &ctx->gps, &ctx->loss, &ctx->reduced, &ctx->flags, &ctx->settle,
                      &ctx->in_x0, &ctx->in_ctrl, &ctx->in_gt, &ctx->out_list, &ctx->out_final, &ctx->io_a, &ctx->io_b, &ctx->io_c};
    for (DevBuf* b : bufs) b->release();
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->d_sq) cudaFree(ctx->d_sq);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* avs_last_error(const avs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int avs_set_nn_params(avs_ctx* ctx, const double* p416) {
    if (!ctx || !p416) return fail(ctx, AVS_E_ARG, "avs_set_nn_params: null argument");
    CK(cudaSetDevice(ctx->device));
    memcpy(ctx->params, p416, sizeof(ctx->params));
    load_weights(ctx);
    CK(cudaMemcpyAsync(ctx->d_params, ctx->params, sizeof(ctx->params), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->params_dirty_on_device = false;
    return AVS_OK;
}
int avs_get_nn_params(avs_ctx* ctx, double* p416) {
    if (!ctx || !p416) return fail(ctx, AVS_E_ARG, "avs_get_nn_params: null argument");
    CK(cudaSetDevice(ctx->device));
    int rc = refresh_params(ctx);
    if (rc) return rc;
    memcpy(p416, ctx->params, sizeof(ctx->params));
    return AVS_OK;
}
int avs_set_bekker_params(avs_ctx* ctx, const double* p5) {
    if (!ctx || !p5) return fail(ctx, AVS_E_ARG, "avs_set_bekker_params: null argument");
    memcpy(ctx->mc.bek, p5, sizeof(ctx->mc.bek));
    return AVS_OK;
}
int avs_get_bekker_params(avs_ctx* ctx, double* p5) {
    if (!ctx || !p5) return fail(ctx, AVS_E_ARG, "avs_get_bekker_params: null argument");
    memcpy(p5, ctx->mc.bek, sizeof(ctx->mc.bek));
    // BekkerDynamics::setParams clamps n1 at 0 (BekkerDynamics.cpp:19-20); getParams returns the clamped value
    if (p5[3] < 0.0) p5[3] = 0.0;
    return AVS_OK;
}

int avs_set_terrain_analytic(avs_ctx* ctx, int kind) {
    if (!ctx || kind < AVS_TERRAIN_FLAT || kind > AVS_TERRAIN_BUMPY) return fail(ctx, AVS_E_ARG, "avs_set_terrain_analytic: kind %d", kind);
    ctx->terr = kind;
    ctx->mc.terr_kind = kind;
    return AVS_OK;
}
int avs_set_terrain_grid(avs_ctx* ctx, const double* height, const double* soil, int nx, int ny, double x0, double y0, double dx, double dy) {
    if (!ctx || !height || nx < 2 || ny < 2 || !(dx > 0) || !(dy > 0)) return fail(ctx, AVS_E_ARG, "avs_set_terrain_grid: bad grid");
    CK(cudaSetDevice(ctx->device));
    const size_t plane = (size_t)nx * ny * sizeof(double);
    CK(ctx->grid_h.reserve(plane));
    CK(cudaMemcpyAsync(ctx->grid_h.p, height, plane, cudaMemcpyHostToDevice, ctx->stream));
    if (soil) {
        CK(ctx->grid_s.reserve(5 * plane));
        CK(cudaMemcpyAsync(ctx->grid_s.p, soil, 5 * plane, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->mc.grid.height = ctx->grid_h.as<double>();
    ctx->mc.grid.soil = soil ? ctx->grid_s.as<double>() : nullptr;
    ctx->mc.grid.nx = nx; ctx->mc.grid.ny = ny; ctx->mc.grid.x0 = x0; ctx->mc.grid.y0 = y0;
    ctx->mc.grid.inv_dx = 1.0 / dx; ctx->mc.grid.inv_dy = 1.0 / dy;
    ctx->terr = AVS_TERRAIN_GRID;
    ctx->mc.terr_kind = AVS_TERRAIN_GRID;
    return AVS_OK;
}

int avs_set_contact_search(avs_ctx* ctx, int on) {
    if (!ctx || (on != 0 && on != 1)) return fail(ctx, AVS_E_ARG, "avs_set_contact_search: on must be 0 or 1");
    ctx->mc.contact_search = on;
    return AVS_OK;
}

int avs_set_checkpoint_budget(avs_ctx* ctx, unsigned long long bytes) {
    if (!ctx) return AVS_E_ARG;
    ctx->ckpt_budget = bytes;
    return AVS_OK;
}

int avs_whole_body_com(double* com3) {
    if (!com3) return AVS_E_ARG;
    // getWholeBodyCOM (generated/miscellaneous.cpp:6-39): wheel link COMs are their joint origins
    const double M = kMb + 4 * kMw;
    com3[0] = (kCx * kMb + kMw * (kTx + kTx - kTx - kTx)) / M;
    com3[1] = (0.0 * kMb + kMw * (kTy - kTy + kTy - kTy)) / M;
    com3[2] = (kCz * kMb + kMw * (kTz + kTz + kTz + kTz)) / M;
    return AVS_OK;
}
int avs_init_state_com(const double* s, double* b) {
    if (!s || !b) return AVS_E_ARG;
    double com[3];
    avs_whole_body_com(com);
    const double w[3] = {s[11], s[12], s[13]};
    // base_lin_vel = com_lin_vel + com^ w   (HybridDynamics.cpp:118-124)
    const double lv[3] = {s[14] + (0. * w[0] - com[2] * w[1] + com[1] * w[2]), s[15] + (com[2] * w[0] + 0. * w[1] - com[0] * w[2]),
                          s[16] + (-com[1] * w[0] + com[0] * w[1] + 0. * w[2])};
    for (int i = 0; i < 21; i++) b[i] = s[i];
    b[14] = lv[0]; b[15] = lv[1]; b[16] = lv[2];
    return AVS_OK;
}

int avs_settle(avs_ctx* ctx, double* state21_out) {
    if (!ctx || !state21_out) return fail(ctx, AVS_E_ARG, "avs_settle: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->settle.reserve(21 * sizeof(double)));
    LAUNCH(launch_settle(ctx->tire, ctx->mc, live(ctx), ctx->settle.as<double>(), ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(state21_out, ctx->settle.p, 21 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

int avs_ode(avs_ctx* ctx, int N, const double* x, const double* u, double* xd) {
    if (!ctx || N <= 0 || !x || !u || !xd) return fail(ctx, AVS_E_ARG, "avs_ode: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)N;
    CK(ctx->io_a.reserve(21 * n * 8)); CK(ctx->io_b.reserve(2 * n * 8)); CK(ctx->io_c.reserve(21 * n * 8));
    CK(cudaMemcpyAsync(ctx->io_a.p, x, 21 * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->io_b.p, u, 2 * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(launch_ode(ctx->tire, ctx->mc, live(ctx), N, ctx->io_a.as<double>(), ctx->io_b.as<double>(), ctx->io_c.as<double>(), ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(xd, ctx->io_c.p, 21 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

static int rollout_dev_impl(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, double* d_x_list, double* d_x_final,
                            double* d_margins) {
    if (!ctx || N <= 0 || T < 1 || substeps < 1 || !d_x0 || (T > 1 && !d_ctrl)) return fail(ctx, AVS_E_ARG, "avs_rollout: bad argument (N=%d T=%d substeps=%d)", N, T, substeps);
    if (d_margins && ctx->tire != AVS_TIRE_BEKKER) return fail(ctx, AVS_E_STATE, "avs_rollout_margins: branch margins exist for the Bekker tire model only");
    CK(cudaSetDevice(ctx->device));
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.Nc = N; a.T = T; a.substeps = substeps; a.x0 = d_x0; a.ctrl = d_ctrl; a.x_list = d_x_list; a.x_final = d_x_final; a.margin = d_margins;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    LAUNCH(launch_rollout_fwd(ctx->tire, ctx->terr, false, ctx->mc, live(ctx), a, ctx->stream));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->ev_fwd = true; ctx->ev_bwd = false;
    ctx->launches++;
    return AVS_OK;
}

int avs_rollout_dev(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, double* d_x_list, double* d_x_final) {
    return rollout_dev_impl(ctx, N, T, substeps, d_x0, d_ctrl, d_x_list, d_x_final, nullptr);
}

static int rollout_host_impl(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final, double* margins);
int avs_rollout(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final) {
    return rollout_host_impl(ctx, N, T, substeps, x0, ctrl, x_list, x_final, nullptr);
}
int avs_rollout_margins(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final, double* margins) {
    if (!margins) return fail(ctx, AVS_E_ARG, "avs_rollout_margins: margins is null");
    return rollout_host_impl(ctx, N, T, substeps, x0, ctrl, x_list, x_final, margins);
}

static int rollout_host_impl(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final, double* margins) {
    if (!ctx || N <= 0 || T < 1 || substeps < 1 || !x0 || (T > 1 && !ctrl)) return fail(ctx, AVS_E_ARG, "avs_rollout: bad argument (N=%d T=%d substeps=%d)", N, T, substeps);
    CK(cudaSetDevice(ctx->device));
    if (margins) CK(ctx->margins.reserve((size_t)kNumMargins * N * 8));
    const size_t n = (size_t)N;
    const size_t b_x0 = 21 * n * 8, b_ctrl = (size_t)(T - 1) * 2 * n * 8, b_list = (size_t)T * 23 * n * 8;
    CK(ctx->in_x0.reserve(b_x0));
    if (b_ctrl) CK(ctx->in_ctrl.reserve(b_ctrl));
    if (x_list) CK(ctx->out_list.reserve(b_list));
    if (x_final) CK(ctx->out_final.reserve(b_x0));
    CK(cudaMemcpyAsync(ctx->in_x0.p, x0, b_x0, cudaMemcpyHostToDevice, ctx->stream));
    if (b_ctrl) CK(cudaMemcpyAsync(ctx->in_ctrl.p, ctrl, b_ctrl, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rollout_dev_impl(ctx, N, T, substeps, ctx->in_x0.as<double>(), ctx->in_ctrl.as<double>(), x_list ? ctx->out_list.as<double>() : nullptr,
                              x_final ? ctx->out_final.as<double>() : nullptr, margins ? ctx->margins.as<double>() : nullptr);
    if (rc) return rc;
    if (margins) CK(cudaMemcpyAsync(margins, ctx->margins.p, (size_t)kNumMargins * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (x_list) CK(cudaMemcpyAsync(x_list, ctx->out_list.p, b_list, cudaMemcpyDeviceToHost, ctx->stream));
    if (x_final) CK(cudaMemcpyAsync(x_final, ctx->out_final.p, b_x0, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

// host_gt (optional): ground truth still in host memory; it is copied into d_gt on the copy stream AFTER the first
// forward sweep has been enqueued (only the reverse sweep reads it), so the copy runs under the forward kernel.
static int rollout_grad_impl(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, const double* d_gt, double explode_thresh,
                             int explode_mode, double* d_loss, double* d_reduced, double* d_grad_live, const double* host_gt) {
    if (!ctx || N <= 0 || T < 2 || substeps < 1 || !d_x0 || !d_ctrl || !d_gt || (explode_mode != 0 && explode_mode != 1))
        return fail(ctx, AVS_E_ARG, "avs_rollout_grad: bad argument (N=%d T=%d substeps=%d mode=%d)", N, T, substeps, explode_mode);
    if (ctx->tire != AVS_TIRE_NN) return fail(ctx, AVS_E_STATE, "avs_rollout_grad: the adjoint exists for the NN tire model only");
    if (ctx->mc.contact_search) return fail(ctx, AVS_E_STATE, "avs_rollout_grad: the 10-angle contact search is forward only (the reference never tapes it)");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)N;
    if (!d_loss) { CK(ctx->loss.reserve(n * 8)); d_loss = ctx->loss.as<double>(); }
    if (!d_grad_live) { CK(ctx->gps.reserve((size_t)kNumLive * n * 8)); d_grad_live = ctx->gps.as<double>(); }
    // reverse-sweep buffers: (S*4 + 1) state records of 14 doubles (tiled by 8 vehicles) + S*4 x 4 wheels activation records of 20 doubles per vehicle; process the batch in chunks if it does not fit the budget
    const size_t per_vehicle_ck = ((size_t)(T - 1) * substeps * 4 + 1) * kCkRows * 8;
    const size_t per_vehicle_act = (size_t)(T - 1) * substeps * 4 * kActLen * 4 * 8;  // + tile padding, see act_bytes below
    const size_t per_vehicle = per_vehicle_ck + per_vehicle_act;
    size_t chunk;
    if (ctx->plan_N == N && ctx->plan_per_vehicle == per_vehicle && ctx->plan_budget == ctx->ckpt_budget) {
        chunk = ctx->plan_chunk;  // same problem shape as the last call: no driver queries on the hot path
    } else {
        size_t budget = ctx->ckpt_budget;
        if (budget == 0) {
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            budget = (size_t)((double)(free_b + ctx->ckpt.cap + ctx->act.cap) * 0.6);
        }
        chunk = budget / per_vehicle;
        if (chunk >= n) chunk = n;
        else chunk = (chunk / 256) * 256;  // whole CTAs (aligning chunks to whole resident waves was measured: no gain, round 2)
        ctx->plan_N = N; ctx->plan_per_vehicle = per_vehicle; ctx->plan_budget = ctx->ckpt_budget; ctx->plan_chunk = chunk;
    }
    if (chunk == 0) return fail(ctx, AVS_E_NOMEM, "avs_rollout_grad: the reverse-sweep buffers of one 256-vehicle chunk (%zu B) exceed the memory budget", per_vehicle * 256);
    CK(ctx->ckpt.reserve(((size_t)(T - 1) * substeps * 4 + 1) * ck_rec_stride(chunk) * 8));  // whole 8-vehicle state tiles
    CK(ctx->act.reserve((size_t)(T - 1) * substeps * 4 * act_stage_doubles(chunk) * 8));  // warp-tiled, padded to whole forward CTAs
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    for (size_t begin = 0; begin < n; begin += chunk) {
        const size_t cnt = (n - begin < chunk) ? (n - begin) : chunk;
        RolloutArgs a;
        memset(&a, 0, sizeof(a));
        a.N = N; a.Nc = (int)cnt; a.T = T; a.substeps = substeps;
        a.x0 = d_x0 + begin; a.ctrl = d_ctrl + begin; a.gt = d_gt + begin; a.ckpt = ctx->ckpt.as<double>(); a.act = ctx->act.as<double>();
        a.ckpt_doubles = ctx->ckpt.cap / 8; a.act_doubles = ctx->act.cap / 8;
        a.loss = d_loss + begin; a.gps = d_grad_live + begin;
        LAUNCH(launch_rollout_fwd(AVS_TIRE_NN, ctx->terr, true, ctx->mc, live(ctx), a, ctx->stream));
        if (begin == 0 && cnt == n) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        if (begin == 0 && host_gt) {
            CK(cudaMemcpyAsync(const_cast<double*>(d_gt), host_gt, (size_t)T * 3 * n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
        }
        LAUNCH(launch_rollout_bwd(ctx->terr, ctx->mc, live(ctx), a, ctx->stream));
        ctx->launches += 2;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->ev_fwd = (chunk == n); ctx->ev_bwd = (chunk == n);
    if (d_reduced) {
        if (explode_mode == 1) CK(ctx->flags.reserve(n));
        LAUNCH(launch_filter_reduce(d_grad_live, d_loss, N, explode_thresh, explode_mode, ctx->flags.as<unsigned char>(), d_reduced, ctx->stream));
        ctx->launches += (explode_mode == 1) ? 2 : 1;
    }
    return AVS_OK;
}

int avs_rollout_grad_dev(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, const double* d_gt, double explode_thresh,
                         int explode_mode, double* d_loss, double* d_reduced, double* d_grad_live) {
    return rollout_grad_impl(ctx, N, T, substeps, d_x0, d_ctrl, d_gt, explode_thresh, explode_mode, d_loss, d_reduced, d_grad_live, nullptr);
}

// scatter [104][N] live gradients into the reference's [416][N] layout (rows 104..415 are zero)
int avs_rollout_grad(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, const double* gt, double explode_thresh,
                     int explode_mode, double* loss, double* reduced, double* grad_per_sample) {
    if (!ctx || N <= 0 || T < 2 || !x0 || !ctrl || !gt) return fail(ctx, AVS_E_ARG, "avs_rollout_grad: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)N;
    const size_t b_x0 = 21 * n * 8, b_ctrl = (size_t)(T - 1) * 2 * n * 8, b_gt = (size_t)T * 3 * n * 8;
    CK(ctx->in_x0.reserve(b_x0)); CK(ctx->in_ctrl.reserve(b_ctrl)); CK(ctx->in_gt.reserve(b_gt));
    CK(ctx->loss.reserve(n * 8)); CK(ctx->gps.reserve((size_t)kNumLive * n * 8)); CK(ctx->reduced.reserve(AVS_REDUCED_LEN * 8));
    CK(cudaMemcpyAsync(ctx->in_x0.p, x0, b_x0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in_ctrl.p, ctrl, b_ctrl, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rollout_grad_impl(ctx, N, T, substeps, ctx->in_x0.as<double>(), ctx->in_ctrl.as<double>(), ctx->in_gt.as<double>(), explode_thresh,
                               explode_mode, ctx->loss.as<double>(), ctx->reduced.as<double>(), ctx->gps.as<double>(), gt);
    if (rc) return rc;
    if (loss) CK(cudaMemcpyAsync(loss, ctx->loss.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (reduced) CK(cudaMemcpyAsync(reduced, ctx->reduced.p, AVS_REDUCED_LEN * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (grad_per_sample) {
        CK(cudaMemcpyAsync(grad_per_sample, ctx->gps.p, (size_t)kNumLive * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        memset(grad_per_sample + (size_t)kNumLive * n, 0, (size_t)(AVS_NUM_NN_PARAMS - kNumLive) * n * 8);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

int avs_reset_optimizer(avs_ctx* ctx) {
    if (!ctx) return AVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> ones(AVS_NUM_NN_PARAMS, 1.0);  // m_squared_grad = Ones (Trainer.cpp:63)
    CK(cudaMemcpyAsync(ctx->d_sq, ones.data(), sizeof(double) * AVS_NUM_NN_PARAMS, cudaMemcpyHostToDevice, ctx->stream));
    // only the squared-gradient state is reset; after a device RMSProp step d_params is the authoritative copy and the
    // host snapshot is stale, so it must not be uploaded here (avs_create uploads the initial parameters itself)
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

int avs_rmsprop_step_dev(avs_ctx* ctx, const double* d_reduced, double lr, double l1_weight) {
    if (!ctx) return AVS_E_ARG;
    if (!d_reduced) {  // the context's own buffer: what the host-buffer avs_rollout_grad (+ avs_allreduce) left on the device
        if (!ctx->reduced.p) return fail(ctx, AVS_E_STATE, "avs_rmsprop_step_dev: no reduced gradient on the device yet");
        d_reduced = ctx->reduced.as<double>();
    }
    CK(cudaSetDevice(ctx->device));
    LAUNCH(launch_rmsprop(ctx->d_params, ctx->d_sq, d_reduced, lr, l1_weight, ctx->stream));
    ctx->launches++;
    ctx->params_dirty_on_device = true;
    return AVS_OK;
}

int avs_set_reduced(avs_ctx* ctx, const double* reduced_host) {
    if (!ctx || !reduced_host) return fail(ctx, AVS_E_ARG, "avs_set_reduced: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->reduced.reserve(AVS_REDUCED_LEN * 8));
    CK(cudaMemcpyAsync(ctx->reduced.p, reduced_host, AVS_REDUCED_LEN * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // the caller's buffer may be pageable and short-lived
    return AVS_OK;
}

int avs_rmsprop_step(avs_ctx* ctx, const double* reduced_host, double lr, double l1_weight) {
    if (!ctx || !reduced_host) return fail(ctx, AVS_E_ARG, "avs_rmsprop_step: null argument");
    int rc = avs_set_reduced(ctx, reduced_host);
    if (rc) return rc;
    return avs_rmsprop_step_dev(ctx, ctx->reduced.as<double>(), lr, l1_weight);
}

// ---- the reference's auxiliary networks as batched device functions ------------------------------------------------------
int avs_load_legacy_net_csv(const char* path, double* params435) {
    if (!path || !params435) return AVS_E_ARG;
    FILE* f = fopen(path, "r");
    if (!f) return AVS_E_ARG;
    // one number per line; line 426 of the reference's file carries a junk token after the number ("285.06...;in_mean 0.0",
    // i.e. two numbers): every token that parses as a number counts, anything else is skipped
    int n = 0;
    char tok[256];
    while (n < kLegacyNetParams + 1 && fscanf(f, " %255[^ \t\r\n;,]", tok) == 1) {
        char* end = nullptr;
        const double v = strtod(tok, &end);
        if (end != tok && *end == '\0') { if (n < kLegacyNetParams) params435[n] = v; n++; }
        int c = fgetc(f);  // the delimiter
        (void)c;
    }
    fclose(f);
    return n == kLegacyNetParams ? AVS_OK : AVS_E_ARG;
}
int avs_mlp_forward(avs_ctx* ctx, int kind, const double* params, int N, const double* in, double* out) {
    if (!ctx || !params || N <= 0 || !in || !out || (kind != MLP_BASE_NETWORK && kind != MLP_LEGACY_TIRE)) return fail(ctx, AVS_E_ARG, "avs_mlp_forward: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int np = kind == MLP_BASE_NETWORK ? kBaseNetParams : kLegacyNetParams, nin = kind == MLP_BASE_NETWORK ? 3 : 5;
    const size_t n = (size_t)N;
    CK(ctx->io_a.reserve(nin * n * 8)); CK(ctx->io_b.reserve(np * 8)); CK(ctx->io_c.reserve(3 * n * 8));
    CK(cudaMemcpyAsync(ctx->io_a.p, in, nin * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->io_b.p, params, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_mlp(kind, ctx->io_b.as<double>(), N, ctx->io_a.as<double>(), ctx->io_c.as<double>(), ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->io_c.p, 3 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}

// ---- pinned host staging ----------------------------------------------------------------------------------------------
int avs_alloc_pinned(void** out, unsigned long long bytes) {
    if (!out || bytes == 0) return AVS_E_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault);
    if (e == cudaErrorMemoryAllocation) return AVS_E_NOMEM;
    return e == cudaSuccess ? AVS_OK : AVS_E_CUDA;
}
int avs_free_pinned(void* p) {
    if (!p) return AVS_OK;
    return cudaFreeHost(p) == cudaSuccess ? AVS_OK : AVS_E_CUDA;
}

// ---- multi-GPU: one communicator rank per context, one all-reduce of reduced[418] per training step ---------------------
int avs_comm_unique_id(char* id128) {
    if (!id128) return AVS_E_ARG;
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return AVS_E_COMM;
    static_assert(sizeof(ncclUniqueId) == AVS_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return AVS_E_COMM;
    memcpy(id128, &id, sizeof(id));
    return AVS_OK;
}
int avs_comm_init_rank(avs_ctx* ctx, const char* id128, int world, int rank) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, AVS_E_ARG, "avs_comm_init_rank: bad argument (world=%d rank=%d)", world, rank);
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return fail(ctx, AVS_E_COMM, "%s", api->err.c_str());
    if (ctx->comm) return fail(ctx, AVS_E_STATE, "avs_comm_init_rank: the context already has a communicator");
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    CKN(api->CommInitRank(&ctx->comm, world, id, rank));
    ctx->world = world; ctx->rank = rank;
    return AVS_OK;
}
int avs_comm_init_all(avs_ctx** ctxs, int n) {
    if (!ctxs || n < 1 || n > 64) return AVS_E_ARG;
    avs_ctx* ctx = ctxs[0];
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return fail(ctx, AVS_E_COMM, "%s", api->err.c_str());
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < n; i++) {
        if (!ctxs[i] || ctxs[i]->comm) return fail(ctx, AVS_E_STATE, "avs_comm_init_all: context %d is null or already has a communicator", i);
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(ctx, AVS_E_ARG, "avs_comm_init_all: contexts %d and %d share device %d", j, i, devs[i]);
    }
    CKN(api->CommInitAll(comms, n, devs));
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->world = n; ctxs[i]->rank = i; }
    return AVS_OK;
}
int avs_comm_destroy(avs_ctx* ctx) {
    if (!ctx) return AVS_E_ARG;
    if (ctx->comm) {
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        CKN(nccl_api()->CommDestroy(ctx->comm));
        ctx->comm = nullptr; ctx->world = 1; ctx->rank = 0;
    }
    return AVS_OK;
}
int avs_comm_info(const avs_ctx* ctx, int* world, int* rank, int* nccl_version) {
    if (!ctx) return AVS_E_ARG;
    if (world) *world = ctx->world;
    if (rank) *rank = ctx->rank;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi* api = nccl_api();
        if (api->err.empty()) api->GetVersion(nccl_version);
    }
    return AVS_OK;
}
int avs_allreduce_dev(avs_ctx* ctx, double* d_reduced) {
    if (!ctx) return AVS_E_ARG;
    if (!d_reduced) {
        if (!ctx->reduced.p) return fail(ctx, AVS_E_STATE, "avs_allreduce_dev: no reduced gradient on the device yet");
        d_reduced = ctx->reduced.as<double>();
    }
    if (!ctx->comm) {
        if (ctx->world == 1) return AVS_OK;  // single rank: the local sum is the global sum
        return fail(ctx, AVS_E_COMM, "avs_allreduce_dev: no communicator");
    }
    CK(cudaSetDevice(ctx->device));
    // in place, on the context's stream: ordered after the filter/reduce kernel that produced it and before the RMSProp
    // kernel that consumes it; no host hop
    CKN(nccl_api()->AllReduce(d_reduced, d_reduced, AVS_REDUCED_LEN, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return AVS_OK;
}
int avs_allreduce(avs_ctx* ctx, double* reduced_host_out) {
    int rc = avs_allreduce_dev(ctx, nullptr);
    if (rc) return rc;
    if (reduced_host_out) {
        CK(cudaMemcpyAsync(reduced_host_out, ctx->reduced.p, AVS_REDUCED_LEN * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return AVS_OK;
}

void* avs_stream(avs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int avs_set_stream(avs_ctx* ctx, void* stream) {
    if (!ctx) return AVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)stream;  // NULL is a valid handle: the legacy default stream
    return AVS_OK;
}
int avs_use_own_stream(avs_ctx* ctx) {
    if (!ctx) return AVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return AVS_OK;
}
int avs_sync(avs_ctx* ctx) {
    if (!ctx) return AVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return AVS_OK;
}
long long avs_kernel_launches(const avs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int avs_last_kernel_ms(avs_ctx* ctx, float* fwd_ms, float* bwd_ms) {
    if (!ctx) return AVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (fwd_ms) { *fwd_ms = 0; if (ctx->ev_fwd) CK(cudaEventElapsedTime(fwd_ms, ctx->ev[0], ctx->ev[1])); }
    if (bwd_ms) { *bwd_ms = 0; if (ctx->ev_bwd) CK(cudaEventElapsedTime(bwd_ms, ctx->ev[1], ctx->ev[2])); }
    return AVS_OK;
}

int avs_measure_fp64_peak(avs_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(ctx, AVS_E_ARG, "avs_measure_fp64_peak: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    CK(ctx->reduced.reserve(AVS_REDUCED_LEN * 8));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    CK(launch_fp64_peak(blocks, threads, 1024, ctx->reduced.as<double>(), ctx->stream));  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(ctx->ev[3], ctx->stream));
        CK(launch_fp64_peak(blocks, threads, iters, ctx->reduced.as<double>(), ctx->stream));
        CK(cudaEventRecord(ctx->ev[2], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[2]));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    ctx->launches += 6;
    ctx->ev_bwd = false;
    *tflops = best;
    return AVS_OK;
}

}  // extern "C"
