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
    AVS_E_NOMEM = 4
};

enum { AVS_TIRE_NN = 0, AVS_TIRE_BEKKER = 1 };
enum { AVS_TERRAIN_FLAT = 0, AVS_TERRAIN_SLOPE = 1, AVS_TERRAIN_BUMPY = 2, AVS_TERRAIN_GRID = 3 };
enum { AVS_EXPLODE_ZERO_COMPONENT = 0, AVS_EXPLODE_DROP_SAMPLE = 1 };

#define AVS_STATE_DIM 21
#define AVS_CNTRL_DIM 2
#define AVS_NUM_NN_PARAMS 416
#define AVS_NUM_BEKKER_PARAMS 5
#define AVS_REDUCED_LEN 418 /* grad_sum[416] | loss_sum | n_kept */

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
int avs_rmsprop_step_dev(avs_ctx* ctx, const double* d_reduced, double lr, double l1_weight);
/* host variant: uploads reduced[418] first (e.g. after a host-side reduction over ranks) */
int avs_rmsprop_step(avs_ctx* ctx, const double* reduced_host, double lr, double l1_weight);
int avs_reset_optimizer(avs_ctx* ctx);

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
