// uba_device.h — device-side views and kernel launchers shared by uba_kernels.cu and uba_host.cu.
#ifndef UBA_DEVICE_H_INCLUDED
#define UBA_DEVICE_H_INCLUDED

#include <cuda_runtime.h>
#include <stdint.h>

#include "uba_math.h"

namespace uba {

// Per-window scalars accumulated by the kernels of one LM iteration (zeroed before each).  They
// are grouped by how a point-sharded multi-GPU run reduces them:
//   w_lin  [nW][2]  cost, failures of the lineariser      -> summed over ranks with S (after linearise)
//   w_post [nW][5]  candidate cost, point parts of the model change / step norm / x norm, stop requests (wall-clock cap)
//                                                          -> summed over ranks (after back-substitution)
//   w_loc  [nW][4]  camera parts + solver failures          -> identical on every rank, never reduced
//   w_max  [nW]     projected-gradient max norm             -> max over ranks
enum WLin { WL_COST = 0, WL_FAIL = 1, WL_COUNT = 2 };
enum WPost { WP_COSTNEW = 0, WP_MCPT = 1, WP_STEP2 = 2, WP_X2 = 3, WP_STOP = 4, WP_COUNT = 5 };
enum WLoc { WC_MCCAM = 0, WC_STEP2 = 1, WC_X2 = 2, WC_FAIL = 3, WC_COUNT = 4 };

// Levenberg–Marquardt controller state of one window (device resident).
struct WinState {
  double radius;
  double decrease_factor;
  double cost;            // cost at the current iterate (from the last linearisation)
  double gmax;
  double initial_cost;
  int32_t cur;            // parity of the current iterate in the double-buffered cams/pts
  int32_t done;           // 0 running, else uba_termination
  int32_t iter;           // LM iterations executed
  int32_t n_success, n_unsuccess, n_invalid, consecutive_invalid;
  int32_t scale_ready;    // Jacobi scaling (iteration 0) captured
  int32_t pad_[2];
};

struct IterRec {  // mirrors uba_iteration
  double cost, candidate_cost, model_cost_change, relative_decrease, radius, step_norm, gradient_max_norm;
  int32_t accepted, pad_;
};

struct SolverCfg {
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int32_t max_consecutive_invalid_steps, fixed_iterations, max_iterations, jacobi_scaling, use_bounds;
};

// Segment-local tiling plan of the fast lineariser (see uba_kernels.cu, k_lin_tile).  A part is a
// contiguous range of internal point slots whose cameras all belong to one short, ascending "local
// camera list"; the fixed cameras (camIdx < fixedFrames) are its first n_fixed entries.
struct TilePart {
  int32_t window;
  int32_t pt_begin, pt_end;   // internal point slots
  int32_t cam_list_off;       // offset into tile_cams (window-local camera indices, ascending)
  int32_t n_local;            // local cameras, <= kTileMaxLocal
  int32_t n_fixed;            // leading fixed cameras; free ones: n_local - n_fixed <= kTileMaxFree
  int32_t pad_[2];
};
constexpr int kTileThreads = 256;
constexpr int kTileMaxLocal = 32;
constexpr int kTileMaxFree = 21;   // 21*22/2 = 231 camera-pair blocks <= kTileThreads

struct DevView {
  int32_t M, nW, NC, NP;
  int64_t NO;
  // windows
  const int32_t* w_cam_off;   // [nW+1]
  const int32_t* w_pt_off;    // [nW+1] internal point slots
  const int32_t* w_free_off;  // [nW+1] offsets into free_list
  const int64_t* w_red_off;   // [nW+1] offsets into Sacc / A (sum of n^2)
  const int32_t* free_list;   // global camera index of each free camera, window by window
  const int32_t* w_beta;      // [nW] > 0: the reduced system is banded with this scalar half-bandwidth and
                              //      is solved by k_chol_banded (k_assemble then only builds lambda and rhs)
  // cameras (global index = w_cam_off[w] + local)
  double* cams[2];            // [NC][6]
  double* camR[2];            // [NC][kCamStride]
  const int32_t* free_cam;    // [NC] compact index in the window's reduced system, or -1
  const int32_t* cam_win;     // [NC]
  double* cam_s2;             // [NC][6] Jacobi scale^2
  double* cam_lam;            // [NC][6] camera damping of this iteration
  double* cam_y;              // [NC][6] camera solution y_c (step = -y), zeros for fixed cameras
  // points (internal order)
  double* pts[2];             // [NP][3]
  double* pt_is2;             // [NP][3] 1 / (Jacobi scale^2) of the point columns, captured at the first linearisation
  double* pt_rec;             // [NP][kPtRec]
  const int32_t* pt_obs_off;  // [NP+1]
  const int32_t* pt_win;      // [NP]
  // observations (internal order), SoA
  const double* feat;         // [M][NO]
  const int32_t* obs_cam;     // [NO] window-local camera index | (camID != 0) << 30
  // accumulators
  double* Sacc;               // per window [n][n], upper block triangle: sum_j Z Z^T
  double* Bacc;               // [NC][36]
  double* vacc;               // [NC][6]
  double* zh;                 // [NC][6]
  double* w_lin;              // [nW][WL_COUNT]
  double* w_post;             // [nW][WP_COUNT]
  double* w_loc;              // [nW][WC_COUNT]
  double* w_max;              // [nW]
  double* A;                  // per window [n][n] assembled damped reduced matrix (full), then its factor
  double* rhs;                // [6 * n_free_total]
  double* Zbuf;               // [NO][18] scratch of the generic lineariser
  // tiled lineariser plan
  const TilePart* parts;      // [n_parts]
  int32_t n_parts;
  int32_t tile_threads;       // 128 or 256 threads per CTA of k_lin_tile2
  const int32_t* tile_cams;   // local camera lists
  const uint32_t* pt_mask;    // [NP] bit s: the point observes slot s of its part's list (0: not a tile point)
  const int32_t* gen_pts;     // [n_gen] internal point slots handled by the generic lineariser
  int32_t n_gen;
  // pipelined solve (single banded window on one GPU): the band solver runs UNDER the lineariser.  A part's CTA counts itself
  // done for each of its cameras; whoever completes a camera assembles that camera's rows of the reduced system and raises
  // its flag; the solver waits on the flags of the rows it loads.
  int32_t pipe_on;
  const int32_t* cam_expect;  // [NC] parts whose camera list holds the camera
  int32_t* cam_done;          // [NC] parts that have flushed (zeroed per iteration)
  int32_t* row_ready;         // [free cameras] rows 6 f .. 6 f + 5 of the band and the rhs are assembled (zeroed per iteration)
  WinState* ws;               // [nW]
  IterRec* recs;              // [nW][rec_stride]
  int32_t rec_stride;
  int32_t* n_active;          // [1] windows still running
  Calib calib;
  LossCfg loss;
  SolverCfg cfg;
};

// Optional debug / parity outputs of a linearisation (device pointers, internal order; may be null).
struct DebugOut {
  double* residuals;  // [NO][M]
  double* weights;    // [NO]
  double* C;          // [NP][9]
  double* W;          // [NO][18]
  double* grad_pts;   // [NP][3]
  double* lam_pts;    // [NP][3]
};

// ---- point-sharded windows: the cross-rank sums over NVLink PEER MEMORY (uba_peer.cu) --------------------------------
// Every rank maps every other rank's accumulator block, flag words and inbox (CUDA IPC, one process per GPU).  Two
// exchange points per LM iteration, both plain kernels inside the captured iteration graph:
//   k_peer_reduce  after the lineariser: arrive (a flag word stored into every peer), wait for the peers whose shards
//                  touch the same cameras, then PULL their partial {B, v, Z h, cost, band of S_acc} straight out of their
//                  HBM over NVLink and sum in rank order into the local consumer copy (bitwise identical on every rank);
//   k_peer_post    after the back-substitution: PUSH this rank's {candidate cost, model change, norms, gradient max,
//                  stop request} into every peer's inbox, arrive, wait, sum the inbox in rank order.
// A peer that does not arrive within timeout_ns raises *err (the host turns it into UBA_ERR_NCCL) instead of spinning.
constexpr int kMaxRanks = 16;
constexpr int kInboxSlots = WP_COUNT + 1;      // w_post sums (incl. the stop request), gradient max-norm
struct PeerView {
  int32_t rank, n_ranks;
  double* acc[kMaxRanks];                  // accumulator block of every rank (own entry: the local pointer)
  unsigned long long* flags[kMaxRanks];    // arrival words of every rank: flags[p][r] = last exchange rank r reached
  double* inbox[kMaxRanks];                // [2][n_ranks][nW][kInboxSlots] of every rank
  unsigned long long* epoch;               // local: [0] exchanges passed, [1] post exchanges passed, [2] CTA completion counter
  int32_t* err;                            // local: a peer did not arrive in time
  const int32_t* cam_lo;                   // [n_ranks] first / last global camera index each rank's shard observes
  const int32_t* cam_hi;
  const double* stop_req;                  // local scalar set by the host: this rank wants to stop (wall-clock cap)
  long long timeout_ns;
};
// layout of the accumulator block (doubles), identical on every rank
struct AccLayout {
  int64_t off_Bacc, off_vacc, off_zh, off_wlin, off_sacc, sum_end;   // [0, sum_end): what the lineariser produces
  int64_t off_wpost, off_wmax, off_wloc, total;
};
int launch_peer_reduce(const DevView& V, const PeerView& P, const AccLayout& L, double* acc_red, int dense, cudaStream_t st);
int launch_peer_post(const DevView& V, const DevView& Vc, const PeerView& P, cudaStream_t st);
int launch_peer_barrier(const PeerView& P, cudaStream_t st);

// launchers (uba_kernels.cu); all asynchronous on `st`; return the number of kernels launched
int launch_cam_prep(const DevView& V, int parity, cudaStream_t st);
// only_listed: process V.gen_pts instead of every point
int launch_lin_generic(const DevView& V, const DebugOut& dbg, bool only_listed, cudaStream_t st);
constexpr int kSlotMaxLocal = 10;  // widest part (local cameras) the slot-per-warp lineariser takes: 320-thread CTAs
constexpr int kLinVariants = 8;    // launch classes of the tiled linearisers: 2 of k_lin_slot, 5 of k_lin_tile2, k_lin_wide
int launch_lin_tiled(const DevView& V, const int* variant_off, cudaStream_t st);
int lin_part_variant(int n_local, int n_free_local, bool slot_kernel);
int launch_assemble(const DevView& V, int max_n, cudaStream_t st);
// h_win_n: host array [nW] of reduced-system sizes (6 * free cameras); h_win_beta: [nW] banded half-bandwidth or 0
// parts: 3 = factorisation + epilogue (default), 1 = factorisation only, 2 = epilogue only
int launch_solve(const DevView& V, const int* h_win_n, const int* h_win_beta, int max_small_n, cudaStream_t st, bool keep_factor = false, int parts = 3);
constexpr int kBandMaxBeta = 63;
int solve_small_limit();
int launch_backsub(const DevView& V, cudaStream_t st);
int launch_lm_update(const DevView& V, cudaStream_t st);
int launch_init_state(const DevView& V, double initial_radius, cudaStream_t st);
int launch_cov_state(const DevView& V, int enter, cudaStream_t st);
int launch_cov_blocks(const DevView& V, int n_free_total, int max_n, double* cov36, cudaStream_t st);
int launch_l2_flush(double* buf, size_t n, cudaStream_t st);
int launch_rank_max(double* w_max, double* w_rmax, double* w_post, const double* stop_req, int nW, int rank, int n_ranks, int gather, cudaStream_t st);
int launch_ingest_feats(const void* raw, const int32_t* src, double* feat, int64_t NO, int M, int raw_is_f32, cudaStream_t st);
int launch_dfma_probe(double* out, int iters, cudaStream_t st);
// sliding window (uba_window_advance)
int launch_win_shift(const double* old_rows, const int32_t* old_off, const int32_t* dropped, const int32_t* id_map, const int32_t* new_off,
                     int old_np, int M, double* new_rows, cudaStream_t st);
int launch_win_append(const double* feats, const int32_t* pt_idx, const int32_t* cam_idx, const int32_t* lo, const int32_t* off, int n, int M,
                      double* rows, cudaStream_t st);
int launch_win_pack(const double* rows, const int32_t* off_c, const int32_t* lo_c, const unsigned char* cid_c, const int32_t* pt_order,
                    const int32_t* off_i, int NP, int64_t NO, int M, double* feat, int32_t* obs_cam, cudaStream_t st);
int launch_win_points(const double* old_pts, const int32_t* src_slot, const double* fresh, int NP, double* out, cudaStream_t st);
int launch_win_rows(const void* raw, int raw_is_f32, double* rows, int64_t n, cudaStream_t st);

}  // namespace uba
#endif
