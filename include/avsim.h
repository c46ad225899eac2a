/* avsim.h — C ABI of libavsim.so: B200-native batched differentiable Jackal rollouts (fp64).
 *
 * Drop-in boundary for the ONE data-parallel hot path of jyurkanin/auvsl_neural_ode: rolling out N
 * independent vehicle trajectories (articulated-body dynamics + tire model + terrain + RK4) and
 * back-propagating Trainer's trajectory loss to the tire-network weights.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - return 0 = OK, non-zero = error (AVS_E_*); message via avs_last_error(ctx).  No exceptions and
 *     no C++ / torch types cross this boundary.
 *   - one context per GPU; a context is NOT thread-safe (caller serialises), it owns its CUDA stream
 *     and all device scratch.  There is no CPU fallback: every call fails with AVS_E_CUDA when no
 *     sm_100-class device is usable.
 *   - all arrays are fp64, struct-of-arrays "[component][N]" (vehicle index fastest) so that
 *     neighbouring lanes touch neighbouring addresses:
 *       x0      [21][N]      state in the reference layout (include/auvsl_dynamics/HybridDynamics.h:65-69)
 *                            qx,qy,qz,qw | x,y,z | q1..q4 | wx,wy,wz | vx,vy,vz | qd1..qd4
 *       ctrl    [T-1][2][N]  (vl, vr) of data row i, applied while integrating row i -> i+1
 *                            (src/Trainer.cpp:564-565, 619-620)
 *       x_list  [T][23][N]   the reference's x_list (src/Trainer.cpp:557-569): row 0 = x0 + controls
 *                            of row 0; rows >= 1 = integrate() output (entries 21,22 zeroed,
 *                            src/VehicleSystem.cpp:144-147)
 *       gt      [T][3][N]    (yaw, x, y) of each data row (src/Trainer.cpp:626-628)
 *   - "T" is the number of data rows of a window (reference: 200 train / 600 eval, src/Trainer.h:90-93);
 *     a rollout performs (T-1) * substeps RK4 steps of 1e-3 s (src/VehicleSystem.cpp:129,
 *     src/auvsl_dynamics/HybridDynamics.cpp:33).
 *   - *_dev variants take DEVICE pointers, enqueue on the context's stream and return without
 *     synchronising; the plain variants take HOST pointers and include the H2D/D2H copies.
 */
#ifndef AVSIM_H_
#define AVSIM_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avs_ctx avs_ctx;

enum {
    AVS_OK = 0,
    AVS_E_ARG = 1,    /* invalid argument */
    AVS_E_CUDA = 2,   /* CUDA runtime error / no usable device */
    AVS_E_STATE = 3,  /* call not valid for the configured tire model / terrain */
    AVS_E_NOMEM = 4,
    AVS_E_COMM = 5    /* NCCL error / communicator missing */
};

enum { AVS_TIRE_NN = 0, AVS_TIRE_BEKKER = 1 };
enum { AVS_TERRAIN_FLAT = 0, AVS_TERRAIN_SLOPE = 1, AVS_TERRAIN_BUMPY = 2, AVS_TERRAIN_GRID = 3 };
enum { AVS_EXPLODE_ZERO_COMPONENT = 0, AVS_EXPLODE_DROP_SAMPLE = 1 };

#define AVS_STATE_DIM 21
#define AVS_CNTRL_DIM 2
#define AVS_NUM_NN_PARAMS 416
#define AVS_NUM_BEKKER_PARAMS 5
#define AVS_REDUCED_LEN 418 /* grad_sum[416] | loss_sum | n_kept */
#define AVS_NUM_MARGINS 6   /* branch-margin kinds of the Bekker path, see avs_rollout_margins */
#define AVS_COMM_ID_BYTES 128

/* ---- life cycle.  Replaces: `new HybridDynamics` / `makeSystem()` per trajectory
 *      (src/Trainer.cpp:596, src/VehicleSystem.h:38-51, src/BekkerSystem.h factory). */
int avs_create(avs_ctx** out, int device, int tire_model);
void avs_destroy(avs_ctx* ctx);
const char* avs_last_error(const avs_ctx* ctx);
/* usable without a context (static error text for a failed avs_create) */
const char* avs_strerror(int code);

/* ---- parameters.  Replaces System::setParams/getParams/getDefaultParams
 *      (src/types/System.h:18-21; src/VehicleSystem.cpp:23-35,180-184; TireNetwork.cpp:130-226;
 *       src/BekkerSystem.cpp:21-28; BekkerDynamics.cpp:12-29).  p416 uses the reference pack order:
 *      for net k = 0..3: W0 (8x3), W2 (8x8), W4 (2x8), row-major.  Only net 0 is evaluated
 *      (HybridDynamics.cpp:387). */
int avs_nn_default_params(double* p416);
int avs_bekker_default_params(double* p5);
int avs_set_nn_params(avs_ctx* ctx, const double* p416);
int avs_get_nn_params(avs_ctx* ctx, double* p416);
int avs_set_bekker_params(avs_ctx* ctx, const double* p5);
int avs_get_bekker_params(avs_ctx* ctx, double* p5);

/* ---- terrain.  Replaces TerrainMap<Scalar>::getAltitude (src/types/TerrainMap.h:10) and the
 *      Flat/Slope/Bumpy maps (src/TestTerrainMaps.cpp:6-28).  The grid (bilinear height field plus
 *      optional per-cell Bekker soil 5-vectors, soil = [5][ny][nx]) is an extension. */
int avs_set_terrain_analytic(avs_ctx* ctx, int kind);
int avs_set_terrain_grid(avs_ctx* ctx, const double* height, const double* soil, int nx, int ny, double x0, double y0, double dx, double dy);

/* ---- contact model.  0 (default): the contact point lies straight under the hub, as in the reference's ODE
 *      (HybridDynamics::get_tire_cpts + get_tire_sinkages, HybridDynamics.cpp:214-253).  1: the 10-angle contact search of
 *      HybridDynamics::get_tire_cpts_sinkages_fancy (HybridDynamics.cpp:258-316, present but not called in the reference): per
 *      wheel, ten rim points at pitch -0.3 ... +0.3 rad, sinkage = the largest candidate sinkage (>= 0), reported point = the
 *      joint centre (per-cell soil, where configured, is looked up there).  Forward paths only: avs_rollout_grad returns
 *      AVS_E_STATE while it is on. */
int avs_set_contact_search(avs_ctx* ctx, int on);

/* ---- tuning: upper bound in bytes for the reverse-sweep checkpoint buffer (0 = auto: 60 % of free HBM) */
int avs_set_checkpoint_budget(avs_ctx* ctx, unsigned long long bytes);

/* ---- helpers of the adapters (host math, no GPU):
 *      HybridDynamics::initStateCOM (HybridDynamics.cpp:111-133) and getWholeBodyCOM
 *      (generated/miscellaneous.cpp:6-39). */
int avs_whole_body_com(double* com3);
int avs_init_state_com(const double* start21, double* base21);

/* ---- HybridDynamics::initState + settle (HybridDynamics.cpp:62-65,142-190), i.e. what
 *      System::getDefaultInitialState runs (src/VehicleSystem.cpp:151-177).  Writes the full
 *      21-state after the 1000 projected RK4 steps. */
int avs_settle(avs_ctx* ctx, double* state21_out);

/* ---- System::forward (src/types/System.h:22): N independent ODE evaluations.  x [21][N], u [2][N] -> xd [21][N] */
int avs_ode(avs_ctx* ctx, int N, const double* x, const double* u, double* xd);

/* ---- forward rollouts.  Replaces the loop of Trainer::evaluateTrajectory (src/Trainer.cpp:550-575)
 *      around System::integrate (src/VehicleSystem.cpp:113-148, src/BekkerSystem.cpp:145-178).
 *      x_list and x_final may each be NULL. */
int avs_rollout(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final);
int avs_rollout_dev(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, double* d_x_list, double* d_x_final);

/* ---- diagnostic forward rollout of the Bekker model that also reports, per trajectory, how far every value-dependent
 *      branch of BekkerDynamics::get_tire_f_ext / BekkerTireModel::{get_features, get_forces, integrate} stayed from flipping
 *      (minimum over all ODE evaluations of the rollout).  margins [6][N]:
 *        0 |sinkage|                      contact on/off            (BekkerDynamics.cpp:81)
 *        1 |w R - vx|                     sign of Fx                (BekkerDynamics.cpp:93)
 *        2 ||vx| - 1e-3|, ||w R| - 1e-3|  slip clamps               (BekkerTireModel.cpp:197-198)
 *        3 |pgc - pg| / pg                ground-pressure branch    (BekkerTireModel.cpp:232)
 *        4 |R - sinkage|                  zr = min(zr, R)           (BekkerTireModel.cpp:219)
 *        5 |dtheta| of non-empty arcs     trapezoid loop bound      (BekkerTireModel.cpp:166-170)
 *      Two correct implementations may differ by O(1e-3) on a trajectory whose margin is within rounding of 0; parity
 *      tests gate on margin > 1e-10 and report the rest.  Tire model must be AVS_TIRE_BEKKER. */
int avs_rollout_margins(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, double* x_list, double* x_final, double* margins);

/* ---- forward + reverse (BPTT).  Replaces Trainer::trainTrajectory (src/Trainer.cpp:591-659: CppAD
 *      tape + ADFun::Reverse(1)) for N windows at once, plus the per-sample explosion filter and sum of
 *      Trainer::train (:169-193, mode 0: zero the exploding component) or Trainer::combineResults
 *      (:661-689, mode 1: drop the sample).
 *        loss            [N]        per-sample L = (sum_i l_i)/T/traj_len          (may be NULL)
 *        reduced         [418]      grad_sum[416] | loss_sum | n_kept (as double)   (may be NULL)
 *        grad_per_sample [416][N]   unfiltered per-sample gradients                 (may be NULL)
 *      Tire model must be AVS_TIRE_NN. */
int avs_rollout_grad(avs_ctx* ctx, int N, int T, int substeps, const double* x0, const double* ctrl, const double* gt,
                     double explode_thresh, int explode_mode, double* loss, double* reduced, double* grad_per_sample);
/* device variant: d_loss [N], d_reduced [418], d_grad_live [104][N] (gradient w.r.t. net 0 only; may be NULL) */
int avs_rollout_grad_dev(avs_ctx* ctx, int N, int T, int substeps, const double* d_x0, const double* d_ctrl, const double* d_gt,
                         double explode_thresh, int explode_mode, double* d_loss, double* d_reduced, double* d_grad_live);

/* ---- Trainer::updateParams (src/Trainer.cpp:428-453): RMSProp on the context's NN parameters with the
 *      float-truncated gradient g = reduced[0..415] / reduced[417].  d_reduced is a DEVICE pointer
 *      (typically the all-reduced buffer); squared-gradient state lives in the context (starts at 1). */
int avs_rmsprop_step_dev(avs_ctx* ctx, const double* d_reduced /* NULL = the context's own reduced buffer */, double lr, double l1_weight);
/* host variant: uploads reduced[418] first (e.g. after a host-side reduction over ranks) */
int avs_rmsprop_step(avs_ctx* ctx, const double* reduced_host, double lr, double l1_weight);
/* upload reduced[418] into the context's own reduced buffer (e.g. zeros on a rank whose shard is empty, before avs_allreduce) */
int avs_set_reduced(avs_ctx* ctx, const double* reduced_host);
/* squared-gradient state back to ones (m_squared_grad = Ones, src/Trainer.cpp:63); the parameters are not touched */
int avs_reset_optimizer(avs_ctx* ctx);

/* ---- multi-GPU.  Replaces the gradient seam of the reference's worker pool (src/Trainer.cpp:204-263 trainThreads,
 *      :661-689 combineResults: `batch_grad += sample_grad` on the main thread): the batch is sharded by trajectory, one
 *      context (= one NCCL rank) per GPU; each rank filters and sums its shard on the device, then ONE ncclAllReduce(sum) of
 *      reduced[418] (grad_sum | loss_sum | n_kept) runs in place on the context's stream, and every rank applies the same
 *      RMSProp update, so the replicas stay bit-identical without a broadcast.  NCCL is bound at run time (dlopen of
 *      libnccl.so.2); these calls fail with AVS_E_COMM when it is absent.
 *        one process per GPU : rank 0 calls avs_comm_unique_id and ships the 128 bytes to the other ranks (any transport);
 *                              every rank calls avs_comm_init_rank
 *        one process, n GPUs : avs_comm_init_all over n contexts on n distinct devices; then one host thread per context */
int avs_comm_unique_id(char* id128);
int avs_comm_init_rank(avs_ctx* ctx, const char* id128, int world, int rank);
int avs_comm_init_all(avs_ctx** ctxs, int n);
int avs_comm_destroy(avs_ctx* ctx);
int avs_comm_info(const avs_ctx* ctx, int* world, int* rank, int* nccl_version);
/* in-place all-reduce of a DEVICE buffer of 418 doubles (NULL = the context's own reduced buffer, i.e. what the host-buffer
 * avs_rollout_grad left on the device); enqueued on the context's stream, no synchronisation, no host hop.  A context
 * without a communicator and world == 1 returns AVS_OK without doing anything. */
int avs_allreduce_dev(avs_ctx* ctx, double* d_reduced);
/* host convenience: all-reduce the context's own reduced buffer and (optionally) copy the result to reduced_host_out[418] */
int avs_allreduce(avs_ctx* ctx, double* reduced_host_out);

/* ---- the reference's two auxiliary networks as batched device functions; in [n_in][N], out [3][N]
 *      AVS_MLP_BASE_NETWORK: BaseNetwork::forward (src/BaseNetwork.cpp:41-60): (wz, vx, vy) -> (nz, fx, fy); params[131] in the
 *        reference pack order W0 (8x3), b0, W1 (8x8), b1, W2 (3x8), b2 (src/BaseNetwork.cpp:62-152).  The reference's ODE never
 *        calls it (HybridDynamics.cpp:433, 557 commented out), so it is not wired into the rollout kernels either.
 *      AVS_MLP_LEGACY_TIRE: the 5-16-16-3 network + scalers stored in data/nn_pretrained.csv (435 numbers, read by
 *        avs_load_legacy_net_csv, which tolerates the junk token on line 426): out = net((in - in_mean) / in_std) * out_std +
 *        out_mean, tanh hidden layers.  No reference code reads that file or defines its five inputs, hence a
 *        plain function and not a tire model of the ODE. */
enum { AVS_MLP_BASE_NETWORK = 0, AVS_MLP_LEGACY_TIRE = 1 };
#define AVS_NUM_BASE_NETWORK_PARAMS 131
#define AVS_NUM_LEGACY_NET_PARAMS 435
int avs_load_legacy_net_csv(const char* path, double* params435);
int avs_mlp_forward(avs_ctx* ctx, int kind, const double* params, int N, const double* in, double* out);

/* ---- pinned (page-locked) host staging for the SoA buffers of the host-pointer entry points
 *      (Trainer::loadDataFile -> windows -> x0 / ctrl / gt, reference src/Trainer.cpp:102-143, 163-165) */
int avs_alloc_pinned(void** out, unsigned long long bytes);
int avs_free_pinned(void* p);

/* ---- stream / bookkeeping */
void* avs_stream(avs_ctx* ctx);                 /* cudaStream_t the *_dev calls are enqueued on */
/* enqueue on a caller-owned cudaStream_t instead (NULL = the legacy default stream); avs_use_own_stream reverts */
int avs_set_stream(avs_ctx* ctx, void* stream);
int avs_use_own_stream(avs_ctx* ctx);
int avs_sync(avs_ctx* ctx);
long long avs_kernel_launches(const avs_ctx* ctx); /* kernels launched by this context so far */
/* average duration in ms of the most recent rollout kernels (CUDA events on the ctx stream; valid after avs_sync) */
int avs_last_kernel_ms(avs_ctx* ctx, float* fwd_ms, float* bwd_ms);
/* DFMA-chain microbenchmark: measured fp64 FMA throughput of this device in TFLOP/s (FMA = 2 flop) */
int avs_measure_fp64_peak(avs_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* AVSIM_H_ */
