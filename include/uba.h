/* uba.h — C ABI of libuba, the B200-native windowed bundle-adjustment inner loop.
 *
 * This is the drop-in boundary for the hot path of
 *   me::optimisation::BundleAdjuster<M>::optimise
 *   (reference: include/MotionEstimation/optimisation/BundleAdjuster.h:378-476)
 * i.e. everything the reference delegates to Ceres: per-observation stereo/mono
 * reprojection residuals (:78-94, :113-130, :153-171), their 6-DoF pose / 3-D point
 * Jacobians (AutoDiffCostFunction<.,M,6,3>, :97-102,:133-138,:174-179), the robust
 * loss (HuberLoss(1.0), :397,:447; Cauchy is the commented alternative), fixed
 * cameras (:406-407,:452-454), point box bounds (:408-413,:455-460), the
 * SPARSE_SCHUR Levenberg–Marquardt solve (:416-422,:463-469) and the optional pose
 * covariance (:478-528).
 *
 * Plain C: pointers and sizes only.  All floating point is fp64, all indices int32.
 * Every function returns UBA_OK (0) or a negative uba_status; nothing throws across
 * this boundary.  The caller owns every buffer passed in; the handle owns device
 * memory, streams, graphs and (optionally) an NCCL communicator.  One handle per
 * host thread; calls on one handle are synchronous and blocking, like ceres::Solve.
 *
 * There is NO CPU fallback behind these entry points: without a CUDA device (or
 * when the sm_100a kernels cannot be loaded) uba_create fails with UBA_ERR_CUDA.
 */
#ifndef UBA_H_INCLUDED
#define UBA_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UBA_VERSION 100

typedef enum uba_status {
  UBA_OK = 0,
  UBA_ERR_INVALID_ARGUMENT = -1, /* null pointer, negative size, M not in {2,4}, index out of range */
  UBA_ERR_STATE = -2,            /* call order: mirrors the reference's Status state machine (:188) */
  UBA_ERR_CUDA = -3,             /* no device, kernel image missing, launch/runtime error */
  UBA_ERR_INFEASIBLE = -4,       /* a bounded point starts outside its box: Ceres returns FAILURE */
  UBA_ERR_NUMERICAL = -5,        /* LM gave up: too many consecutive invalid steps / non-finite cost */
  UBA_ERR_NCCL = -6,
  UBA_ERR_UNSUPPORTED = -7
} uba_status;

/* Robust loss applied to s = ||r_obs||^2 of each M-row residual block.
 * Reference default: Huber(1.0) (BundleAdjuster.h:397,:447). */
typedef enum uba_loss { UBA_LOSS_TRIVIAL = 0, UBA_LOSS_HUBER = 1, UBA_LOSS_CAUCHY = 2 } uba_loss;

/* Termination, named after ceres::TerminationType as far as the reference reads it
 * through Summary::IsSolutionUsable() (BundleAdjuster.h:427,:474). */
typedef enum uba_termination {
  UBA_TERM_RUNNING = 0,
  UBA_TERM_CONVERGENCE_FUNCTION = 1,
  UBA_TERM_CONVERGENCE_GRADIENT = 2,
  UBA_TERM_CONVERGENCE_PARAMETER = 3,
  UBA_TERM_NO_CONVERGENCE = 4, /* iteration / time cap reached: still "usable" */
  UBA_TERM_FAILURE = 5,        /* not usable: Status::FAILED */
  UBA_TERM_CONVERGENCE_RADIUS = 6
} uba_termination;

/* CalibrationParameters (BundleAdjuster.h:35-45): K[0] (left), K[1] (right),
 * feat_var, baseline.  Only fx1 and cx1 of K[1] are read by the stereo functor
 * (:163); fy1/cy1 are carried for completeness. */
typedef struct uba_calib {
  double fx0, fy0, cx0, cy0;
  double fx1, fy1, cx1, cy1;
  double feat_var;  /* sigma^2; the residual is scaled by 1/sqrt(feat_var) (:400,:448) */
  double baseline;  /* metres; for M=2 a zero baseline becomes 0.5 (:389-390) */
} uba_calib;

/* Solver options.  uba_config_default() fills in the reference's hard-coded values
 * (BundleAdjuster.h:416-420,:463-467) and the Ceres defaults it inherits. */
typedef struct uba_config {
  int32_t loss_kind;               /* uba_loss; default UBA_LOSS_HUBER */
  double loss_scale;               /* default 1.0 */
  int32_t max_iterations;          /* default 50 (Ceres default) */
  double function_tolerance;       /* default 1e-3 (reference) */
  double gradient_tolerance;       /* default 1e-10 */
  double parameter_tolerance;      /* default 1e-8 */
  double initial_radius;           /* default 1e4 */
  double max_radius;               /* default 1e16 */
  double min_radius;               /* default 1e-32 */
  double min_relative_decrease;    /* default 1e-3 */
  double min_lm_diagonal;          /* default 1e-6 */
  double max_lm_diagonal;          /* default 1e32 */
  int32_t max_consecutive_invalid_steps; /* default 5 */
  double max_solver_time_s;        /* default 1.0 (reference); <= 0 disables the wall-clock cap */
  int32_t fixed_iterations;        /* > 0: run exactly this many LM iterations, no convergence tests
                                      (deterministic parity / benchmark mode); default 0 */
  int32_t jacobi_scaling;          /* default 1 */
  int32_t use_bounds;              /* default 1: point box of BundleAdjuster.h:442-443,:455-460 */
  int32_t device;                  /* CUDA device ordinal; default 0 (or LOCAL_RANK under torchrun) */
  int32_t linearizer;              /* 0 auto (tiled: warp-per-camera-slot kernel, Schur products on the FP64 MMA path, for windows
                                      whose tracks span up to 10 keyframes; the wide-part tiled kernel for windows with longer
                                      tracks), 1 generic (thread per point, global fp64 atomics), 2 lane-per-observation tiled
                                      kernel only */
  int32_t compute_covariance;      /* CalibrationParameters::compute_cov (:40); default 0 */
  int32_t solver;                  /* 0 auto (banded LDL^T for large block-banded systems), 1 dense Cholesky only */
  int32_t sliding_window;          /* 1: keep the window's observation rows resident so that uba_window_advance can slide it
                                      (per-frame BA, BASELINE config c2); default 0 */
} uba_config;

/* Per-window result of uba_optimise (ceres::Solver::Summary, as far as the reference uses it). */
typedef struct uba_summary {
  int32_t termination;        /* uba_termination */
  int32_t usable;             /* Summary::IsSolutionUsable(): 1 -> Status::SUCCESSFUL, 0 -> FAILED */
  int32_t iterations;         /* LM iterations executed (excluding iteration 0) */
  int32_t successful_steps;
  int32_t unsuccessful_steps;
  int32_t invalid_steps;
  double initial_cost;
  double final_cost;
  double final_radius;
  double final_gradient_max_norm;
} uba_summary;

/* One record per LM iteration (iteration 0 = initial evaluation). */
typedef struct uba_iteration {
  double cost;             /* cost at the iterate held after this iteration */
  double candidate_cost;   /* cost at x (+) delta */
  double model_cost_change;
  double relative_decrease;
  double radius;           /* trust-region radius used for this step */
  double step_norm;
  double gradient_max_norm;
  int32_t accepted;        /* 1 accepted, 0 rejected, -1 invalid step */
  int32_t pad_;
} uba_iteration;

/* Optional outputs of uba_linearize: every pointer may be NULL.  Host buffers.
 * n_free = number of non-fixed cameras, n = 6*n_free.  Blocks are row-major. */
typedef struct uba_linearization_out {
  double* residuals;   /* [n_obs][M]   raw residuals r (before the loss corrector), caller order */
  double* weights;     /* [n_obs]      sqrt(rho'(s)) */
  double* cost;        /* [n_windows]  0.5*sum rho(s) */
  double* grad_cams;   /* [n_cams][6]  J~^T r~ camera part (fixed cameras: zeros) */
  double* grad_pts;    /* [n_pts][3]   J~^T r~ point part, caller order */
  double* B;           /* [n_cams][36] sum F^T F, undamped (fixed cameras: zeros) */
  double* C;           /* [n_pts][9]   sum E^T E, undamped, caller order */
  double* W;           /* [n_obs][18]  F^T E (6x3) per observation (fixed cameras: zeros), caller order */
  double* S;           /* per window, concatenated: [n][n] damped reduced camera matrix (full symmetric) */
  double* rhs;         /* per window, concatenated: [n] */
  double* lm_diag_cams;/* [n_cams][6]  lambda added to the camera diagonal */
  double* lm_diag_pts; /* [n_pts][3]   lambda added to the point diagonal */
} uba_linearization_out;

typedef struct uba_handle uba_handle;

/* ---- lifecycle -------------------------------------------------------------------- */
void uba_config_default(uba_config* cfg);
int uba_create(const uba_config* cfg, uba_handle** out);
void uba_destroy(uba_handle* h);
const char* uba_last_error(const uba_handle* h); /* h may be NULL: last error of uba_create */
int uba_version(void);

/* ---- problem set-up (replaces the BundleAdjuster constructors + initialise*,
 *      BundleAdjuster.h:196-228, :286-376) ------------------------------------------- *
 * cams6  [n_cams][6]  = [tx,ty,tz, rx,ry,rz], model p_cam = R(r) X + t  (:304-309)
 * pts3   [n_pts][3]
 * feats  [n_obs][M]   AoS as in Observation<M>::data (:25)
 * cam_idx/pt_idx/cam_id [n_obs] as in Observation<M> (:26-28); cam_id may be NULL (all 0)
 * Observations may come in any order; the library orders them point-major
 * (stable), which is the order initialiseObservations produces (:364-374).
 * Copies everything: the caller may free its buffers on return (:273-275). */
int uba_set_problem(uba_handle* h, int M, int n_cams, int n_pts, int n_obs,
                    const double* cams6, const double* pts3, const double* feats,
                    const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                    const uba_calib* calib);

/* Batch of independent windows sharing one calibration: concatenated arrays plus
 * [n_windows+1] offset tables; cam_idx / pt_idx are window-local. */
int uba_set_batch(uba_handle* h, int M, int n_windows,
                  const int32_t* win_cam_off, const int32_t* win_pt_off, const int64_t* win_obs_off,
                  const double* cams6, const double* pts3, const double* feats,
                  const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                  const uba_calib* calib);

/* Sliding window (per-frame BA; replaces re-running initialiseObservations, BundleAdjuster.h:351-376, over the whole track
 * container, core/feature_types.h:121-191, for every frame).  Needs uba_config.sliding_window = 1, a single window whose
 * tracks are runs of consecutive keyframes (what WBA_Point::addMatch asserts) given point-major / frame-ascending, and one
 * camID per track; otherwise UBA_ERR_UNSUPPORTED and the caller re-submits with uba_set_problem.
 *   - the n_drop oldest keyframes leave: their observations are removed, camera indices shift down by n_drop;
 *   - tracks left without observations are ERASED and the surviving points renumbered densely in order (what a caller does
 *     with its std::vector<WBA_Point>); pt_id_map [old n_pts] (may be NULL) receives the new id of every old point or -1;
 *   - n_new_cams keyframes are appended (new_cams6), n_new_pts new points are appended after the survivors (new_pts3,
 *     new_pt_cam_id or NULL), and n_new_obs observations are added: cam_idx in the NEW window numbering and >= the first
 *     new keyframe, pt_idx in the NEW point numbering, each extending its track by consecutive keyframes;
 *   - cams6_all / pts3_all: initial iterate of the new window for ALL its cameras / points, or NULL to keep what the last
 *     uba_optimise left on the device (plus new_cams6 / new_pts3 for the newcomers).
 * Everything happens on the device apart from O(n_pts) index work; the handle is then in the same state as after
 * uba_set_problem of the equivalent window (uba_optimise, uba_get_*, uba_get_tables work as usual). */
int uba_window_advance(uba_handle* h, int n_drop, int n_new_cams, const double* new_cams6, int n_new_pts, const double* new_pts3,
                       const int32_t* new_pt_cam_id, int n_new_obs, const double* feats, const int32_t* cam_idx, const int32_t* pt_idx,
                       const double* cams6_all, const double* pts3_all, int32_t* pt_id_map);

/* ---- the hot path ----------------------------------------------------------------- */
/* One linearisation at the current iterate: residuals, analytic Jacobians, robust
 * weights, per-point Schur elimination into the reduced camera system.  `radius`
 * is the LM trust-region radius used for the damping (<= 0: undamped). */
int uba_linearize(uba_handle* h, int fixed_frames, double radius, uba_linearization_out* out);

/* BundleAdjuster<M>::optimise(fixedFrames) (:378-476).  summaries: [n_windows] or NULL. */
int uba_optimise(uba_handle* h, int fixed_frames, uba_summary* summaries);

/* ---- results (getCameraPoses / getPoints, :231-237; covariance :238, :478-528) ---- */
int uba_get_cameras(uba_handle* h, double* cams6);
int uba_get_points(uba_handle* h, double* pts3);
int uba_get_pose_covariances(uba_handle* h, double* cov36); /* [n_cams][36]; fixed cameras: zeros */
int uba_get_iterations(uba_handle* h, int window, uba_iteration* out, int max_records, int* n_records);
int uba_get_sizes(const uba_handle* h, int* n_windows, int* n_cams, int* n_pts, int64_t* n_obs);

/* ---- index / ordering tables (bit-exact against the oracle) ----------------------- *
 * obs_order [n_obs]   caller observation id at each internal (point-major, camera-ascending) slot
 * pt_obs_off[n_pts+1] CSR offsets of the point-major order, caller point order
 * pt_order  [n_pts]   caller point id at each internal (segment-sorted) point slot
 * free_cam  [n_cams]  -1 for fixed / unobserved cameras, else compact index in the reduced system
 * Any pointer may be NULL.  free_cam needs fixed_frames. */
int uba_get_tables(uba_handle* h, int fixed_frames, int32_t* obs_order, int64_t* pt_obs_off,
                   int32_t* pt_order, int32_t* free_cam);

/* ---- multi-GPU: one process per GPU; point-sharded windows allreduce the reduced
 *      camera system over NCCL.  The unique id is produced on rank 0 and broadcast
 *      by the launcher (torch.distributed / MPI / a file). ---------------------------- */
#define UBA_NCCL_UNIQUE_ID_BYTES 128
int uba_comm_unique_id(uba_handle* h, char id[UBA_NCCL_UNIQUE_ID_BYTES]);
int uba_comm_init(uba_handle* h, const char id[UBA_NCCL_UNIQUE_ID_BYTES], int rank, int n_ranks);
/* After uba_comm_init, uba_set_problem receives this rank's POINT SHARD (all cameras,
 * a subset of points and their observations) and uba_linearize / uba_optimise reduce
 * over ranks.  uba_set_batch problems never communicate. */

/* Host-side sharding helpers (no CUDA; also exported by libuba_host.so).  uba_shard_points assigns every point of a
 * window to one of n_ranks shards by KEYFRAME RANGE — points ordered by (first keyframe of the track, caller index), cut
 * into pieces of equal observation count — so that a rank's Schur products touch one stretch of the block band of the
 * reduced camera system and only neighbouring ranks overlap.  pt_rank [n_pts] out; rank_obs / rank_pts [n_ranks] out or
 * NULL.  uba_shard_extract writes one rank's points (caller order kept, indices renumbered) and observations and returns
 * the number of observations written; pt_ids_out [rank_pts] receives the caller point id of each shard point. */
int uba_shard_points(int n_cams, int n_pts, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx, int n_ranks,
                     int32_t* pt_rank, int64_t* rank_obs, int32_t* rank_pts);
int64_t uba_shard_extract(int M, int n_pts, int64_t n_obs, const double* pts3, const double* feats, const int32_t* cam_idx,
                          const int32_t* pt_idx, const int32_t* cam_id, const int32_t* pt_rank, int rank, double* pts3_out,
                          double* feats_out, int32_t* cam_idx_out, int32_t* pt_idx_out, int32_t* cam_id_out, int32_t* pt_ids_out);

/* ---- device-side timing for benchmarks (CUDA events on the library's stream) ------- */
typedef struct uba_timing {
  double linearize_ms;   /* summed over launches since the last reset */
  double solve_ms;
  double backsub_ms;
  double update_ms;
  double comm_ms;
  double total_ms;       /* whole uba_optimise device time */
  int64_t linearize_launches;
  int64_t kernel_launches; /* all kernels launched by the library since the last reset */
} uba_timing;
int uba_set_profiling(uba_handle* h, int enabled); /* 1: per-phase events (serialises phases) */
int uba_get_timing(uba_handle* h, uba_timing* out, int reset);
/* Runs `repeats` linearise+Schur passes on the resident problem and returns the mean
 * device time per pass in milliseconds (CUDA events, library stream). */
int uba_time_linearize(uba_handle* h, int fixed_frames, double radius, int repeats, int flush_l2,
                       double* ms_per_pass);
int uba_time_iteration(uba_handle* h, int fixed_frames, int iterations, int flush_l2,
                       double* ms_per_iteration);
/* fp64 FMA throughput of the device (TFLOP/s), measured with a DFMA micro-kernel: the second
 * roofline ceiling of the lineariser (MEASURED_PEAKS.json carries no fp64 figure). */
int uba_probe_fp64_tflops(uba_handle* h, double* tflops);

/* ---- pose-only mode: the numerical core of the frame-to-frame stereo visual odometry ---- *
 * Replaces, for me::StereoVisualOdometry (reference src/vo/StereoVisualOdometry.cpp): project3D (:22-32, triangulation from
 * disparity), reproject (:116-143), updateJacobian (:291-329), optimize (:165-283, Gauss-Newton or Levenberg-Marquardt on
 * 3 Euler angles + translation with the points fixed), computeInliers (:94-114) and the RANSAC loop of process (:59-75) —
 * with every hypothesis fitted and scored concurrently on the device.  The caller keeps drawing the random triples (the
 * reference uses rand(), :145-163) and passes them in, so a run is reproducible against the reference's own stream.
 * Deliberate difference: optimize()'s loop ends when a stop condition is set or after max_iter iterations; the reference's
 * condition (:277) compares the iteration counter with the StopCondition enum value and does not terminate on ordinary
 * data. */
typedef struct uba_vo_params {
  double fu1, fv1, cu1, cv1, fu2, fv2, cu2, cv2, baseline;   /* StereoVisualOdometry::parameters (vo/StereoVisualOdometry.h:26-35) */
  int32_t method;             /* 0 Gauss-Newton (reference default), 1 Levenberg-Marquardt (VisualOdometry.h:16) */
  int32_t max_iter;           /* default 100 */
  double e1, e2, e3, e4;      /* mean squared reprojection error, gradient, increment, relative decrease thresholds */
  double inlier_threshold;    /* pixels; default 2.0 */
} uba_vo_params;
void uba_vo_params_default(uba_vo_params* p);
/* quads8 [n][8] float: previous left (x,y), previous right, current left, current right — StereoOdoMatchesf f1..f4.
 * Uploads the matches, triangulates them on the device (project3D) and keeps points + observations resident. */
int uba_vo_set_matches(uba_handle* h, const uba_vo_params* params, int n, const float* quads8);
int uba_vo_get_points(uba_handle* h, double* pts4 /*[n][4] normalised homogeneous*/);
/* Parity / debug: one evaluation at state6 = {roll, pitch, yaw, tx, ty, tz} over `selection`: A = J J^T [6][6], B = J r [6],
 * residuals [n_sel][4] (observed - predicted), J [6][4 n_sel] (the layout of m_J).  Any output may be NULL. */
int uba_vo_linearize(uba_handle* h, const double state6[6], int n_sel, const int32_t* selection, double* A36, double* B6,
                     double* residuals, double* J);
/* All RANSAC hypotheses at once: triples [n_hyp][3] match indices; a hypothesis is skipped (ok = 0) when twice the area of
 * its current-left triangle is <= 1000 px^2 (:66) or its fit does not converge; best_hyp = the EARLIEST hypothesis with the
 * most inliers (-1: none).  The inlier set of the best hypothesis stays on the device for uba_vo_refine. */
int uba_vo_ransac(uba_handle* h, const double init6[6], int n_hyp, const int32_t* triples, int32_t* best_hyp, int32_t* inlier_counts,
                  int32_t* hyp_ok, double* hyp_states);
int uba_vo_get_inliers(uba_handle* h, int32_t* idx, int32_t* n_inliers);
/* optimize() from init6 over `selection` (or, selection == NULL, over the inliers of the last uba_vo_ransac);
 * converged = optimize()'s return value. */
int uba_vo_refine(uba_handle* h, const double init6[6], int n_sel, const int32_t* selection, double state6[6], int32_t* converged,
                  int32_t* iterations);

/* Host-only helpers with the device code's conventions: getMotion() (:331-342) — T16 = row-major [R(euler)^T | t; 0 0 0 1] of
 * state6 = {roll, pitch, yaw, tx, ty, tz}; reproject() (:116-141) — predicted {left x, left y, right x, right y} of n
 * homogeneous points pts4 [n][4] (uba_vo_get_points) under state6 (what getPredictions() returns for the inliers). */
void uba_vo_pose_matrix(const double state6[6], double T16[16]);
void uba_vo_predict(const uba_vo_params* params, const double state6[6], int n, const double* pts4, double* pred4);

/* ---- synthetic stereo-rig generator; host only, deterministic ---- */
typedef struct uba_synth_spec {
  int32_t M;               /* 4 stereo, 2 mono/two-camera */
  int32_t n_cams;
  int32_t n_pts;
  int32_t track_min;       /* track length drawn uniformly in [track_min, track_max] */
  int32_t track_max;
  int32_t full_tracks;     /* 1: every point is seen by every keyframe (C1) */
  double outlier_fraction; /* observations replaced by uniform-random pixels (C5) */
  double pixel_sigma;      /* 0.5 */
  double pose_t_sigma;     /* 0.05 */
  double pose_r_sigma;     /* 0.005 */
  double point_rel_sigma;  /* 0.01 */
  int32_t fixed_frames;    /* cameras [0, fixed_frames) keep their ground-truth pose */
  uint64_t seed;
} uba_synth_spec;
void uba_synth_default_calib(uba_calib* calib);
/* Returns the number of observations written (<= max_obs) or a negative uba_status. */
int64_t uba_synth_generate(const uba_synth_spec* spec, const uba_calib* calib, int64_t max_obs,
                           double* cams_gt, double* cams_init, double* pts_gt, double* pts_init,
                           double* feats, int32_t* cam_idx, int32_t* pt_idx, int32_t* cam_id);

/* ---- boundary maths shared with the adapter (rotation_utils.h:190-204) ------------- */
void uba_log_map_quat(const double q_wxyz[4], double r[3]);
void uba_exp_map_quat(const double r[3], double q_wxyz[4]);

#ifdef __cplusplus
}
#endif
#endif /* UBA_H_INCLUDED */
