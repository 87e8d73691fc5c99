// uba_host.cu — host side of libuba: the C ABI of include/uba.h.
//
// Mirrors what me::optimisation::BundleAdjuster<M> does around ceres::Solve
// (reference include/MotionEstimation/optimisation/BundleAdjuster.h):
//   uba_set_problem / uba_set_batch  ~ constructors + initialiseParameters/Observations (:196-228,:286-376)
//   uba_optimise                     ~ optimise(fixedFrames) (:378-476)
//   uba_get_cameras / uba_get_points ~ getCameraPoses / getPoints (:231-237)
// Everything numerical runs in the kernels of uba_kernels.cu; this file only builds
// index tables, moves buffers and sequences launches.  No CPU fallback.
#include <cuda_runtime.h>
#include <omp.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <numeric>
#include <string>
#include <map>
#include <vector>

#include "../../include/uba.h"
#include "uba_device.h"
#ifndef UBA_EMU
#include "uba_vo.h"
#endif

using namespace uba;

namespace {

thread_local std::string g_create_error;

// ---- minimal NCCL binding, resolved at run time (libnccl is only needed for sharded runs) ------
struct Id128 { char internal[UBA_NCCL_UNIQUE_ID_BYTES]; };  // ncclUniqueId
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommGetAsyncError)(void*, int*) = nullptr;
  int (*CommAbort)(void*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclDouble = 8, kNcclChar = 0, kNcclSum = 0, kNcclMax = 2;

bool load_nccl(NcclApi& api, std::string& err) {
  if (api.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) { err = "cannot dlopen libnccl.so.2"; return false; }
  api.GetUniqueId = (int (*)(void*))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(api.lib, "ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
  api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllGather");
  api.CommGetAsyncError = (int (*)(void*, int*))dlsym(api.lib, "ncclCommGetAsyncError");
  api.CommAbort = (int (*)(void*))dlsym(api.lib, "ncclCommAbort");
  api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { err = "libnccl lacks required symbols"; return false; }
  return true;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  // grows with a quarter of headroom: a sliding window's sizes creep up and down from call to call, and every
  // cudaFree / cudaMalloc pair costs milliseconds
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    const size_t want = cap ? n + n / 4 + 256 : std::max<size_t>(n, 1);
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e != cudaSuccess && want > n) { cudaGetLastError(); e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) cap = n; return e; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    const size_t want = cap ? n + n / 4 + 256 : std::max<size_t>(n, 1);
    cap = 0;
    cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
    if (e != cudaSuccess && want > n) { cudaGetLastError(); e = cudaMallocHost((void**)&p, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) cap = n; return e; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct uba_handle {
  uba_config cfg;
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  // problem (host tables)
  int state = 0;  // 0 uninitialised, 1 problem set, 2 optimised
  bool device_dirty = false;   // a timing / parity call iterated on the device copy: uba_optimise restores the staged iterate first
  int M = 4, nW = 0, NC = 0, NP = 0;
  int64_t NO = 0;
  uba_calib calib_in{};
  std::vector<int32_t> w_cam_off, w_pt_off;        // [nW+1]
  std::vector<int64_t> w_obs_off;                  // [nW+1]
  std::vector<int32_t> obs_order;                  // canonical table (global caller obs id per point-major slot)
  std::vector<int64_t> pt_obs_off_caller;          // [NP+1] canonical CSR (caller point order)
  std::vector<int32_t> pt_order;                   // [NP] caller point id (global) per internal slot
  PinBuf<int32_t> h_obs_internal;                  // [NO] caller obs id per internal obs slot (pinned: uploaded for the device-side gather)
  std::vector<int32_t> pt_obs_off_int;             // [NP+1]
  std::vector<char> cam_seen;                      // [NC]
  std::vector<int32_t> pt_lo, pt_hi;               // [NP] lowest / highest camera of each internal point slot (-1: none)
  std::vector<char> pt_contig;                     // [NP] the track is a run of consecutive cameras, strictly ascending
  std::vector<int32_t> cam_win_h, pt_win_h;
  // prepared for a given fixed_frames
  int prepared_fixed = -1;
  std::vector<int32_t> free_cam_h, free_list_h, w_free_off_h, win_n, win_beta;
  bool dense_override = false;                // covariance / parity dumps: every window uses the dense accumulator layout
  // ingest scratch, kept across calls: fresh multi-megabyte vectors page-fault on every call (~1 ms per c4 window on a VM)
  std::vector<int32_t> sc_cnt, sc_lo, sc_hi, sc_hist;
  std::vector<int64_t> sc_first;
  std::vector<char> sc_contig;
  std::vector<std::pair<size_t, size_t>> sum_ranges, zero_ranges;   // (offset, count) pieces of the accumulator block summed over ranks
  std::vector<char> win_infeasible;           // a window whose start violates the point bounds
  DevBuf<int32_t> d_w_beta;
  std::vector<int64_t> w_red_off_h;
  int max_n = 0;
  bool use_tile = false;
  bool use_slot = true;                       // the slot-per-warp lineariser takes the parts that are narrow enough
  int tile_threads = 256;
  int variant_off[kLinVariants + 1] = {};
  std::vector<TilePart> parts_h;
  std::vector<int32_t> tile_cams_h, gen_pts_h;
  std::vector<uint32_t> pt_mask_h;
  DevBuf<TilePart> d_parts;
  DevBuf<uint32_t> d_pt_mask;
  DevBuf<int32_t> d_tile_cams, d_gen_pts;
  // pinned staging (internal order)
  PinBuf<double> h_cams, h_pts, h_feat, h_out;
  PinBuf<int32_t> h_obs_cam;
  PinBuf<int32_t> h_adv;                      // uba_window_advance: every index table of one advance, in the device's layout
  PinBuf<double> h_cams_prev;                 // uba_window_advance: the previous window's cameras (fetched while the host plans)
  // device
  DevBuf<double> d_cams, d_camR, d_cam_s2, d_cam_lam, d_cam_y, d_pts, d_pt_s2, d_pt_rec, d_feat, d_acc, d_A, d_rhs, d_Zbuf, d_dbg, d_export, d_flush;
  DevBuf<int32_t> d_w_cam_off, d_w_pt_off, d_w_free_off, d_free_list, d_free_cam, d_cam_win, d_pt_obs_off, d_pt_win, d_obs_cam, d_obs_src, d_n_active, d_pt_order;
  DevBuf<int64_t> d_w_red_off;
  DevBuf<double> d_cov;
  std::vector<double> cov_h;   // [NC][36], filled when cfg.compute_covariance
  DevBuf<WinState> d_ws;
  DevBuf<IterRec> d_recs;
  size_t acc_sum1 = 0;   // doubles reduced (sum) after linearise: Sacc|Bacc|vacc|zh|w_lin
  size_t acc_total = 0;  // all accumulator doubles (one memset)
  size_t off_Bacc = 0, off_vacc = 0, off_zh = 0, off_wlin = 0, off_sacc = 0, off_wpost = 0, off_wrmax = 0, off_wmax = 0, off_wloc = 0;
  DevView V{};
  std::vector<WinState> ws_h;
  // comm: NCCL for the set-up exchanges (and as the fallback data path, UBA_PEER=0); the per-iteration sums of a
  // point-sharded window go over NVLink peer memory (uba_peer.cu)
  NcclApi nccl;
  void* comm = nullptr;
  int rank = 0, n_ranks = 1;
  bool any_rank_infeasible = false;           // some rank's point shard starts outside the box: the window fails on every rank
  bool peer_wanted = true;                    // UBA_PEER=0 keeps the NCCL data path
  bool peer_on = false;                       // peer mappings are valid for the current buffers
  bool in_optimise = false;                   // run_iteration_fast is being driven by uba_optimise (not by a timing entry point)
  DevBuf<unsigned char> d_ctl;                // [flags n u64 | epoch 4 u64 | err (8 B) | stop_req double | cam_lo n i32 | cam_hi n i32 | inbox]
  size_t ctl_inbox_off = 0, ctl_bytes = 0;
  DevBuf<double> d_acc_red;                   // consumer copy of the accumulator block (sums over ranks)
  DevBuf<char> d_xchg;                        // staging of the set-up exchanges
  void* mapped_acc[kMaxRanks] = {};           // cudaIpcOpenMemHandle results (own rank: null)
  void* mapped_ctl[kMaxRanks] = {};
  void* exchanged_acc = nullptr;              // local pointers the current mappings were exchanged for
  void* exchanged_ctl = nullptr;
  std::vector<int32_t> rank_cam_lo, rank_cam_hi;
  int local_cam_lo = 0, local_cam_hi = -1;    // cameras this rank's own observations touch (before the OR over ranks)
  PeerView P{};
  AccLayout L{};
  DevView Vc{};                               // consumer view: V with the accumulator pointers into d_acc_red
  double* d_stop_req = nullptr;               // device scalar (inside d_ctl, or d_stop_local without peers)
  DevBuf<double> d_stop_local;
  // CUDA graph of one LM iteration (re-captured whenever the device view changes)
#ifndef UBA_EMU
  cudaGraphExec_t graph_exec = nullptr;
#endif
  int64_t graph_kernels = 0;
  std::vector<char> graph_sig;                // host-side launch geometry the captured graph bakes in (see prepare)
  DevView graph_V, graph_Vc;                  // ... and the device views it was captured with (see run_iteration_fast)
  PeerView graph_P;
  // sliding window (uba_window_advance): per-track table in caller point order + canonical observation rows on the device
  bool tracks_valid = false;
  std::vector<int32_t> tr_lo, tr_cnt;         // [NP] first keyframe / length of each track (0: no observation)
  std::vector<unsigned char> tr_cid;          // [NP] camID != 0
  DevBuf<double> d_rows[2];                   // canonical rows [NO][M] (caller point order), ping-pong across advances
  DevBuf<int32_t> d_tr_off[2];                // their CSR offsets [NP+1]
  int rows_cur = 0;
  DevBuf<int32_t> d_tr_tmp;                   // per-advance index tables
  DevBuf<unsigned char> d_tr_cid;
  DevBuf<double> d_fresh;                     // new rows / new points of an advance
  // pipelined solve: the band solver of a single large window runs under the lineariser (see DevView::pipe_on)
  bool pipe_on = false;
  cudaStream_t stream2 = nullptr;             // the solver's branch of the iteration
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<int32_t> cam_expect_h;
  DevBuf<int32_t> d_cam_expect, d_pipe;       // d_pipe: [cam_done NC | row_ready n_free], zeroed per iteration
  bool host_obs_valid = true;                 // obs_order / h_obs_internal / h_obs_cam describe the resident window
  bool host_iter_pending = false;             // a device -> host copy into h_pts is still in flight on the stream
  // timing
  bool profiling = false;
  uba_timing timing{};
#ifndef UBA_EMU
  uba_vo_state* vo = nullptr;                 // pose-only mode (uba_vo.cu), created on first use
#endif
};

namespace {

int fail(uba_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}
#define CU(h, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return fail(h, UBA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

uba::Calib make_calib(const uba_calib& in, int M, bool use_bounds) {
  uba::Calib k;
  double b = in.baseline;
  if (M == 2 && b == 0.0) b = 0.5;  // BundleAdjuster.h:389-390
  k.fx0 = in.fx0; k.fy0 = in.fy0; k.cx0 = in.cx0; k.cy0 = in.cy0; k.fx1 = in.fx1; k.cx1 = in.cx1; k.baseline = b;
  k.sigma_inv = 1.0 / std::sqrt(in.feat_var);  // sigma = sqrt(feat_var), :400,:448
  const double zmax = in.fx0 * b / 0.1, zmin = in.fx0 * b / (2.0 * in.cx0);  // :442-443
  k.hi[0] = zmax / in.fx0 * in.cx0; k.lo[0] = -k.hi[0];
  k.hi[1] = zmax / in.fy0 * in.cy0; k.lo[1] = -k.hi[1];
  k.hi[2] = zmax; k.lo[2] = zmin;
  (void)use_bounds;
  return k;
}

void drop_graph(uba_handle* h) {
#ifndef UBA_EMU
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
#else
  (void)h;
#endif
}

void refresh_consumer_view(uba_handle* h) {
  h->Vc = h->V;
  if (h->peer_on) {
    double* r = h->d_acc_red.p;
    h->Vc.Sacc = r + h->off_sacc; h->Vc.Bacc = r + h->off_Bacc; h->Vc.vacc = r + h->off_vacc; h->Vc.zh = r + h->off_zh;
    h->Vc.w_lin = r + h->off_wlin; h->Vc.w_post = r + h->off_wpost; h->Vc.w_max = r + h->off_wmax; h->Vc.w_loc = r + h->off_wloc;
  }
}

void fill_view_static(uba_handle* h) {
  DevView& V = h->V;                          // (a captured graph is re-validated against the view in prepare())
  V.M = h->M; V.nW = h->nW; V.NC = h->NC; V.NP = h->NP; V.NO = h->NO;
  V.w_cam_off = h->d_w_cam_off.p; V.w_pt_off = h->d_w_pt_off.p;
  V.cams[0] = h->d_cams.p; V.cams[1] = h->d_cams.p + (size_t)h->NC * 6;
  V.camR[0] = h->d_camR.p; V.camR[1] = h->d_camR.p + (size_t)h->NC * kCamStride;
  V.cam_win = h->d_cam_win.p; V.cam_s2 = h->d_cam_s2.p; V.cam_lam = h->d_cam_lam.p; V.cam_y = h->d_cam_y.p;
  V.pts[0] = h->d_pts.p; V.pts[1] = h->d_pts.p + (size_t)h->NP * 3;
  V.pt_is2 = h->d_pt_s2.p; V.pt_rec = h->d_pt_rec.p; V.pt_obs_off = h->d_pt_obs_off.p; V.pt_win = h->d_pt_win.p;
  V.feat = h->d_feat.p; V.obs_cam = h->d_obs_cam.p;
  V.Zbuf = h->d_Zbuf.p;
  V.ws = h->d_ws.p; V.recs = h->d_recs.p; V.n_active = h->d_n_active.p;
  V.calib = make_calib(h->calib_in, h->M, h->cfg.use_bounds != 0);
  V.loss.kind = h->cfg.loss_kind; V.loss.a = h->cfg.loss_scale;
  SolverCfg& c = V.cfg;
  c.function_tolerance = h->cfg.function_tolerance; c.gradient_tolerance = h->cfg.gradient_tolerance;
  c.parameter_tolerance = h->cfg.parameter_tolerance; c.max_radius = h->cfg.max_radius; c.min_radius = h->cfg.min_radius;
  c.min_relative_decrease = h->cfg.min_relative_decrease; c.min_lm_diagonal = h->cfg.min_lm_diagonal;
  c.max_lm_diagonal = h->cfg.max_lm_diagonal; c.max_consecutive_invalid_steps = h->cfg.max_consecutive_invalid_steps;
  c.fixed_iterations = h->cfg.fixed_iterations; c.max_iterations = h->cfg.max_iterations;
  c.jacobi_scaling = h->cfg.jacobi_scaling; c.use_bounds = h->cfg.use_bounds;
  refresh_consumer_view(h);
}

int allreduce(uba_handle* h, double* buf, size_t count, int op);
int comm_wait(uba_handle* h);
int allgather_bytes(uba_handle* h, const void* mine, size_t bytes, std::vector<char>& all);
int peer_exchange(uba_handle* h);
int ensure_host_obs_tables(uba_handle* h);
void peer_close(uba_handle* h);

// UBA_TRACE=1: host-side phase times on stderr
#define TT(label) if (getenv("UBA_TRACE")) { auto now_ = std::chrono::steady_clock::now(); fprintf(stderr, "  [trace] %-28s %.2f ms\n", label, std::chrono::duration<double, std::milli>(now_ - tt_).count()); tt_ = now_; }

// Plan of the tiled lineariser: walk the internal point order (sorted by lowest / highest camera)
// and cut it into items whose points share one short ascending camera list; points that do not fit
// (too long a track, a camera seen twice) go to the generic lineariser.
void build_tile_plan(uba_handle* h, int fixed_frames) {
  const double tau = 0.85;  // every point of an item sees at least tau of the item's cameras
  h->parts_h.clear(); h->tile_cams_h.clear(); h->gen_pts_h.clear();
  h->pt_mask_h.assign(h->NP, 0u);
  size_t tile_points = 0;
  struct Item { int w, begin, end, cam_off, nl, nfx, ulo; bool range; };
  std::vector<Item> items;
  std::vector<int> uni, merged, cams_p;
  for (int w = 0; w < h->nW; w++) {
    const int s0 = h->w_pt_off[w], s1 = h->w_pt_off[w + 1];
    int item_begin = -1, min_len = 0;
    bool uni_range = true;   // the union is the camera range [ulo, uhi] (true as long as only contiguous tracks were merged)
    int ulo = 0, uhi = -1;
    uni.clear();
    auto close_item = [&](int end) {
      if (item_begin < 0) return;
      if (uni_range) { uni.clear(); for (int c = ulo; c <= uhi; c++) uni.push_back(c); }
      Item it{w, item_begin, end, (int)h->tile_cams_h.size(), (int)uni.size(), 0, ulo, uni_range};
      for (int c : uni) { h->tile_cams_h.push_back(c); if (c < fixed_frames) it.nfx++; }
      items.push_back(it);
      item_begin = -1;
    };
    for (int s = s0; s < s1; s++) {
      const int o0 = h->pt_obs_off_int[s], k = h->pt_obs_off_int[s + 1] - o0;
      if (k == 0) continue;
      const int lo = h->pt_lo[s], hi = h->pt_hi[s];
      if (h->pt_contig[s]) {
        // a run of identical contiguous tracks (the points are sorted by first / last keyframe) gets one decision
        int e = s + 1;
        while (e < s1 && h->pt_lo[e] == lo && h->pt_hi[e] == hi && h->pt_contig[e] && h->pt_obs_off_int[e + 1] > h->pt_obs_off_int[e]) e++;
        const int nfree = hi - std::max(lo, fixed_frames) + 1;
        if (k > kTileMaxLocal || nfree > kTileMaxFree) {
          for (int q = s; q < e; q++) h->gen_pts_h.push_back(q);
          s = e - 1;
          continue;
        }
        bool ok = item_begin >= 0 && uni_range;
        int nlo = lo, nhi = hi;
        if (ok) {
          nlo = std::min(ulo, lo); nhi = std::max(uhi, hi);
          // a gap between the union and the track would put cameras in the list that nobody sees
          const bool touching = lo <= uhi + 1 && hi >= ulo - 1;
          const int size = nhi - nlo + 1, nf = nhi - std::max(nlo, fixed_frames) + 1;
          ok = touching && size <= kTileMaxLocal && nf <= kTileMaxFree && std::min(min_len, k) >= tau * (double)size;
        }
        if (!ok) { close_item(s); uni_range = true; ulo = lo; uhi = hi; min_len = k; item_begin = s; }
        else { ulo = nlo; uhi = nhi; min_len = std::min(min_len, k); }
        std::fill(h->pt_mask_h.begin() + s, h->pt_mask_h.begin() + e, 1u);   // 1u marks a tile point until the mask pass below
        tile_points += (size_t)(e - s);
        s = e - 1;
        continue;
      }
      // general track: explicit camera list (never reached after uba_window_advance: its tracks are contiguous)
      cams_p.clear();
      bool ascending = true;
      int nfree = 0;
      for (int q = 0; q < k; q++) {
        const int c = h->h_obs_cam.p[o0 + q] & 0x3fffffff;
        if (q && c <= cams_p.back()) ascending = false;
        cams_p.push_back(c);
        if (c >= fixed_frames) nfree++;
      }
      if (!ascending || k > kTileMaxLocal || nfree > kTileMaxFree) { h->gen_pts_h.push_back(s); continue; }
      bool ok = item_begin >= 0;
      if (ok) {
        if (uni_range) { uni.clear(); for (int c = ulo; c <= uhi; c++) uni.push_back(c); }
        merged.clear();
        std::set_union(uni.begin(), uni.end(), cams_p.begin(), cams_p.end(), std::back_inserter(merged));
        int nf = 0;
        for (int c : merged) if (c >= fixed_frames) nf++;
        ok = (int)merged.size() <= kTileMaxLocal && nf <= kTileMaxFree && std::min(min_len, k) >= tau * (double)merged.size();
      }
      if (!ok) { close_item(s); uni = cams_p; min_len = k; item_begin = s; }
      else { uni.swap(merged); min_len = std::min(min_len, k); }
      uni_range = false;
      h->pt_mask_h[s] = 1u; tile_points++;
    }
    close_item(s1);
  }
  // A handful of shorter tracks in front of (or behind) a long run of full ones — c4 has ~5 four-keyframe tracks per home
  // keyframe — would cost a CTA and a flush of their own: an item of less than one chunk whose camera range lies inside its
  // neighbour's joins it (the neighbour's camera list does not change; its masks simply have idle slots for those points).
  if (h->cfg.linearizer != 2) {
    std::vector<Item> kept;
    kept.reserve(items.size());
    auto covers = [](const Item& a, const Item& b) {
      return a.w == b.w && a.range && b.range && a.ulo <= b.ulo && a.ulo + a.nl >= b.ulo + b.nl && 2 * b.nl >= a.nl;
    };
    for (size_t i = 0; i < items.size(); i++) {
      const Item& cur = items[i];
      if (cur.end - cur.begin < 32) {
        if (i + 1 < items.size() && items[i + 1].begin == cur.end && covers(items[i + 1], cur)) { items[i + 1].begin = cur.begin; continue; }
        if (!kept.empty() && kept.back().end == cur.begin && covers(kept.back(), cur)) { kept.back().end = cur.end; continue; }
      }
      kept.push_back(cur);
    }
    items.swap(kept);
  }
  // slot masks of the tile points (bit i = the point sees the item's i-th camera), items in parallel
  const int n_items = (int)items.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int ii = 0; ii < n_items; ii++) {
    const Item& it = items[ii];
    const int* cl = h->tile_cams_h.data() + it.cam_off;
    for (int s = it.begin; s < it.end; s++) {
      if (h->pt_mask_h[s] != 1u) continue;     // generic or unobserved point inside the item's range
      unsigned m = 0;
      if (it.range) {
        const int k = h->pt_hi[s] - h->pt_lo[s] + 1;
        m = (k >= 32 ? 0xffffffffu : ((1u << k) - 1u)) << (unsigned)(h->pt_lo[s] - it.ulo);
      } else {
        for (int o = h->pt_obs_off_int[s]; o < h->pt_obs_off_int[s + 1]; o++) {
          const int c = h->h_obs_cam.p[o] & 0x3fffffff;
          m |= 1u << (unsigned)(std::lower_bound(cl, cl + it.nl, c) - cl);
        }
      }
      h->pt_mask_h[s] = m;
    }
  }
  // Which kernel takes an item: the slot-per-warp lineariser (k_lin_slot, chunks of 32 points) up to kSlotMaxLocal local
  // cameras, k_lin_wide beyond; linearizer = 2 (or UBA_LIN_SLOT=0) sends everything to k_lin_tile2.
  h->use_slot = h->cfg.linearizer != 2;
  if (const char* e = std::getenv("UBA_LIN_SLOT")) h->use_slot = std::atoi(e) != 0 && h->cfg.linearizer != 2;
  // A window that needs k_lin_wide for some of its parts keeps it for all of them: the two kernels' CTAs (352 threads x 168
  // registers, 256 x 255) cannot share an SM, so side by side they take turns with one CTA per SM each and two tails
  // (measured on c2: 0.110 ms split against 0.089 ms all wide).
  std::vector<char> win_wide(h->nW, 0);
  for (const Item& it : items) if (it.nl > kSlotMaxLocal) win_wide[it.w] = 1;
  if (const char* e = std::getenv("UBA_LIN_SPLIT")) { if (std::atoi(e) != 0) std::fill(win_wide.begin(), win_wide.end(), 0); }
  auto slot_item = [&](const Item& it) { return h->use_slot && it.nl <= kSlotMaxLocal && !win_wide[it.w]; };
  // CTA size of the other kernels: k_lin_wide always has 256 threads; k_lin_tile2 128 (two CTAs per SM) when every item is
  // narrow, else 256
  int nt = h->use_slot ? 256 : 128;
  if (const char* e = std::getenv("UBA_TILE_THREADS")) nt = std::atoi(e) == 256 ? 256 : nt;
  for (const Item& it : items) {
    const int nlf = it.nl - it.nfx;
    if (nlf * (nlf + 1) / 2 > 128 || nlf > 10) nt = 256;
  }
  h->tile_threads = nt;
  // parts: aim at a few CTAs per SM, never less than 2 chunks of points per CTA
  // (k_lin_wide, which takes the other parts when the slot kernel is on: four parts per SM measured better than two on c2,
  // 0.089 against 0.100 ms; more than that changes nothing — the two-chunk minimum below takes over)
  size_t target_parts = (h->use_slot ? 4 : 2) * 148 * (256 / nt);
  if (const char* e = std::getenv("UBA_TILE_PARTS")) target_parts = (size_t)std::max(1, std::atoi(e));
  // k_lin_slot parts are sized per class (up to 5 local cameras: two CTAs per SM; up to 10: one) by a model of the pass:
  // a part costs its chunks of 32 points plus a fixed c0 (camera records, pipeline fill, flush: ~1.5 chunk times, from the
  // per-CTA timeline of scripts/part_timing.py and the part-size sweeps of scripts/part_sweep.py), the CTAs are dispatched
  // longest first onto the class's resident slots.
  // Candidates: the largest item cut into 1..8 pieces and the ideal load of a slot, the pieces of an item evened out; the
  // makespan of every candidate is simulated on the histogram of part sizes (list scheduling over groups of equally loaded
  // slots), cheapest wins.  c3 (512 equal windows on 148 slots = 3.46 rounds) becomes 1024 half windows = 6.92 rounds;
  // c4 (196 runs of ~33 chunks on 296 slots) becomes 588 parts of ~11 = 1.99 rounds.  "Full pieces + remainder" splits
  // (e.g. 24 + 9 on c4, one round on paper) measure 15-20 % slower than the model says and are not proposed.
  struct SlotSplit { int cap = 0; bool even = false; };
  SlotSplit slot_split[2];
  if (h->use_slot) {
    static const double c0 = [] { const char* e = getenv("UBA_SLOT_C0"); return e ? atof(e) : 1.5; }();
    static const int n_sm = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return std::max(n, 1); }();
    for (int cls = 0; cls < 2; cls++) {
      std::map<int, int> ihist;                 // item size in chunks -> items
      long total = 0; int lmax = 0, nit = 0;
      for (const Item& it : items) {
        if (!slot_item(it) || (it.nl > 5) != (cls == 1)) continue;
        const int ch = (it.end - it.begin + 31) / 32;
        ihist[ch]++; total += ch; lmax = std::max(lmax, ch); nit++;
      }
      if (!nit) continue;
      const int slots = n_sm * (cls == 0 ? 2 : 1);
      auto pieces = [](int ch, int cap, bool even, auto&& emit) {       // emit(size, count) for one item of ch chunks
        const int k = (ch + cap - 1) / cap;
        if (even) { const int lo = ch / k, hi_n = ch % k; if (hi_n) emit(lo + 1, hi_n); emit(lo, k - hi_n); }
        else { if (ch / cap) emit(cap, ch / cap); if (ch % cap) emit(ch % cap, 1); }
      };
      auto makespan = [&](int cap, bool even) {
        std::map<int, long, std::greater<int>> phist;                   // part size -> parts, longest first
        for (const auto& kv : ihist) pieces(kv.first, cap, even, [&](int sz, int n) { phist[sz] += (long)n * kv.second; });
        std::map<double, long> load{{0.0, slots}};                      // slot load -> slots carrying it
        for (const auto& kv : phist) {
          long left = kv.second;
          const double cost = kv.first + c0;
          while (left > 0) {
            auto lo = load.begin();
            const double at = lo->first; const long n = std::min(left, lo->second);
            if (n == lo->second) load.erase(lo); else lo->second -= n;
            load[at + cost] += n; left -= n;
          }
        }
        return load.rbegin()->first;
      };
      std::vector<int> caps;
      for (int k = 1; k <= 8; k++) caps.push_back(std::max(1, (lmax + k - 1) / k));
      const int ideal = (int)((total + (long)(c0 * nit) + slots - 1) / slots);
      for (int c : {ideal, ideal + 1, (ideal + 1) / 2}) caps.push_back(std::max(1, c));
      double best = 1e300;
      for (int cap : caps) {
        const double m = makespan(cap, true);
        if (m < best - 1e-9) { best = m; slot_split[cls].cap = cap; slot_split[cls].even = true; }
      }
      if (const char* e = std::getenv("UBA_SLOT_CAP")) {       // experiments: "<chunks>" or "<chunks>e" (evened)
        slot_split[cls].cap = std::max(1, std::atoi(e)); slot_split[cls].even = std::strchr(e, 'e') != nullptr;
        best = makespan(slot_split[cls].cap, slot_split[cls].even);
      }
      if (getenv("UBA_TRACE")) fprintf(stderr, "  [trace] slot class %d: %d items, %ld chunks on %d slots -> parts of <= %d chunks (%s), modelled makespan %.1f chunk times\n",
                                       cls, nit, total, slots, slot_split[cls].cap, slot_split[cls].even ? "evened" : "full + remainder", best);
    }
  }
  for (const Item& it : items) {
    if (slot_item(it) && slot_split[it.nl > 5].cap > 0) {
      const SlotSplit sp = slot_split[it.nl > 5];
      const int ch = (it.end - it.begin + 31) / 32, k = (ch + sp.cap - 1) / sp.cap;
      int b = it.begin;
      for (int q = 0; q < k; q++) {
        const int sz = sp.even ? ch / k + (q < ch % k ? 1 : 0) : std::min(sp.cap, ch - q * sp.cap);
        TilePart p{};
        p.window = it.w; p.pt_begin = b; p.pt_end = std::min(it.end, b + 32 * sz); p.cam_list_off = it.cam_off; p.n_local = it.nl; p.n_fixed = it.nfx;
        p.pad_[0] = 1;
        h->parts_h.push_back(p);
        b = p.pt_end;
      }
      continue;
    }
    const int Pc = slot_item(it) ? 32 : nt / it.nl;
    int part_pts = std::max<size_t>(2 * (size_t)Pc, (tile_points + target_parts - 1) / target_parts);
    part_pts = ((part_pts + Pc - 1) / Pc) * Pc;
    for (int b = it.begin; b < it.end; b += part_pts) {
      TilePart p{};
      p.window = it.w; p.pt_begin = b; p.pt_end = std::min(it.end, b + part_pts); p.cam_list_off = it.cam_off; p.n_local = it.nl; p.n_fixed = it.nfx;
      p.pad_[0] = slot_item(it) ? 1 : 0;       // host-side note: this part goes to k_lin_slot
      h->parts_h.push_back(p);
    }
  }
  // one launch per kernel variant: parts grouped by variant; inside a variant the parts of a single window go from BOTH ENDS
  // of the window towards its middle (CTAs are dispatched in this order), which is the order in which the two-sided band
  // solver consumes the rows of the reduced system when it runs under the lineariser
  const bool slot = h->use_slot;
  auto variant = [slot](const TilePart& a) {
    const int v = lin_part_variant(a.n_local, a.n_local - a.n_fixed, slot);
    return slot && v < 2 && !a.pad_[0] ? 7 : v;     // a narrow part of a window that stays with k_lin_wide
  };
  const int ncw = h->nW == 1 ? h->NC : 0;
  auto end_dist = [&](const TilePart& a) {
    if (!ncw) return 0;
    const int first = h->tile_cams_h[a.cam_list_off], last = h->tile_cams_h[a.cam_list_off + a.n_local - 1];
    return std::min(first, ncw - 1 - last);
  };
  static const bool pipe_order = [] { const char* e = getenv("UBA_PIPE_SOLVE"); return e && e[0] == '1'; }();
  static const int part_order = [] { const char* e = getenv("UBA_PART_ORDER"); return e ? atoi(e) : 1; }();   // experiments: 0 = item order, 2 = shortest first
  std::stable_sort(h->parts_h.begin(), h->parts_h.end(), [&](const TilePart& a, const TilePart& b) {
    const int va = variant(a), vb = variant(b);
    if (va != vb) return va < vb;
    if (pipe_order) return end_dist(a) < end_dist(b);
    if (part_order == 0) return false;
    if (part_order == 2) return a.pt_end - a.pt_begin < b.pt_end - b.pt_begin;
    return a.pt_end - a.pt_begin > b.pt_end - b.pt_begin;          // longest first: the hardware hands CTAs out in this order
  });
  h->cam_expect_h.assign(h->NC, 0);
  for (const TilePart& p : h->parts_h)
    for (int q = 0; q < p.n_local; q++) h->cam_expect_h[h->w_cam_off[p.window] + h->tile_cams_h[p.cam_list_off + q]]++;
  if (getenv("UBA_TRACE")) {
    std::map<std::pair<int, int>, int> hist;     // (local cameras, chunks of 32 points) -> parts
    for (const TilePart& p : h->parts_h) hist[{p.n_local, (p.pt_end - p.pt_begin + 31) / 32}]++;
    fprintf(stderr, "  [trace] tile plan: %zu items, %zu parts, %zu generic points;", items.size(), h->parts_h.size(), h->gen_pts_h.size());
    int shown = 0;
    for (const auto& kv : hist) if (shown++ < 40) fprintf(stderr, " nl%d x %dch: %d,", kv.first.first, kv.first.second, kv.second);
    fprintf(stderr, "\n");
  }
  for (int v = 0; v <= kLinVariants; v++) h->variant_off[v] = 0;
  for (const TilePart& p : h->parts_h) h->variant_off[variant(p) + 1]++;
  for (int v = 0; v < kLinVariants; v++) h->variant_off[v + 1] += h->variant_off[v];
}

// Tables that depend on fixed_frames: free cameras, reduced-system layout, accumulators.
int prepare(uba_handle* h, int fixed_frames) {
  if (fixed_frames < 0) fixed_frames = 0;
  if (h->prepared_fixed == fixed_frames) return UBA_OK;
  auto tt_ = std::chrono::steady_clock::now();
  const int nW = h->nW, NC = h->NC;
  if (h->comm) {
    // point-sharded: a camera is in the problem if ANY rank observes it -> max over ranks
    std::vector<double> seen(NC);
    for (int c = 0; c < NC; c++) seen[c] = h->cam_seen[c] ? 1.0 : 0.0;
    CU(h, h->d_cam_lam.reserve((size_t)NC * 6));
    CU(h, cudaMemcpyAsync(h->d_cam_lam.p, seen.data(), sizeof(double) * NC, cudaMemcpyHostToDevice, h->stream));
    int rc = allreduce(h, h->d_cam_lam.p, NC, kNcclMax);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(seen.data(), h->d_cam_lam.p, sizeof(double) * NC, cudaMemcpyDeviceToHost, h->stream));
    rc = comm_wait(h);
    if (rc) return rc;
    for (int c = 0; c < NC; c++) h->cam_seen[c] = seen[c] > 0.0;
    // which cameras each rank's shard touches (keyframe-range shards: a short stretch), and whether any rank's
    // start is infeasible — every rank must take the same decisions in uba_optimise
    struct Info { int32_t lo, hi, infeasible, pad; } mine{h->local_cam_lo, h->local_cam_hi, 0, 0};
    for (int w = 0; w < nW; w++) if (h->win_infeasible[w]) mine.infeasible = 1;
    std::vector<char> all;
    rc = allgather_bytes(h, &mine, sizeof(mine), all);
    if (rc) return rc;
    h->rank_cam_lo.assign(h->n_ranks, 0); h->rank_cam_hi.assign(h->n_ranks, -1);
    h->any_rank_infeasible = false;
    for (int r = 0; r < h->n_ranks; r++) {
      const Info& in = reinterpret_cast<const Info*>(all.data())[r];
      h->rank_cam_lo[r] = in.lo; h->rank_cam_hi[r] = in.hi;
      if (in.infeasible) h->any_rank_infeasible = true;
    }
  }
  h->free_cam_h.assign(NC, -1);
  h->free_list_h.clear();
  h->w_free_off_h.assign(nW + 1, 0);
  h->w_red_off_h.assign(nW + 1, 0);
  h->win_n.assign(nW, 0);
  h->max_n = 0;
  for (int w = 0; w < nW; w++) {
    int nf = 0;
    for (int gc = h->w_cam_off[w]; gc < h->w_cam_off[w + 1]; gc++) {
      const int local = gc - h->w_cam_off[w];
      // constant cameras: camIdx < fixedFrames (BundleAdjuster.h:406-407,:452-454); cameras nobody
      // observes never enter the ceres::Problem
      if (local >= fixed_frames && h->cam_seen[gc]) { h->free_cam_h[gc] = nf++; h->free_list_h.push_back(gc); }
    }
    h->w_free_off_h[w + 1] = h->w_free_off_h[w] + nf;
    h->win_n[w] = 6 * nf;
    h->w_red_off_h[w + 1] = h->w_red_off_h[w] + (int64_t)36 * nf * nf;
    h->max_n = std::max(h->max_n, 6 * nf);
  }
  // block half-bandwidth of each window's reduced system (largest spread of free cameras inside a track);
  // big windows with a narrow band (c4, c5) go to the banded solver
  h->win_beta.assign(nW, 0);
  for (int w = 0; w < nW; w++) {
    if (h->win_n[w] <= solve_small_limit()) continue;
    int bw = 0;
    // free indices grow with the camera index, so the spread of a track is free(hi) - free(first free camera >= lo)
    std::vector<int> next_free(h->w_cam_off[w + 1] - h->w_cam_off[w] + 1, -1);
    for (int c = h->w_cam_off[w + 1] - h->w_cam_off[w] - 1; c >= 0; c--) {
      const int f = h->free_cam_h[h->w_cam_off[w] + c];
      next_free[c] = f >= 0 ? f : next_free[c + 1];
    }
#pragma omp parallel for schedule(static) reduction(max : bw) if (nW < 8)
    for (int s2 = h->w_pt_off[w]; s2 < h->w_pt_off[w + 1]; s2++) {
      if (h->pt_hi[s2] < 0) continue;
      const int fhi = h->free_cam_h[h->w_cam_off[w] + h->pt_hi[s2]];  // the highest camera of a track is observed, hence free unless fixed
      const int flo = next_free[h->pt_lo[s2]];
      if (fhi >= 0 && flo >= 0) bw = std::max(bw, fhi - flo);
    }
    if (h->comm) {  // every rank must take the same path: the band is the max over ranks
      double v = bw;
      CU(h, h->d_cam_lam.reserve(std::max<size_t>((size_t)NC * 6, 1)));
      CU(h, cudaMemcpyAsync(h->d_cam_lam.p, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      int rc = allreduce(h, h->d_cam_lam.p, 1, kNcclMax);
      if (rc) return rc;
      CU(h, cudaMemcpyAsync(&v, h->d_cam_lam.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      rc = comm_wait(h);
      if (rc) return rc;
      bw = (int)v;
    }
    const int beta = 6 * bw + 5;
    // the banded solver keeps {factor rows of both halves, assembled band} = 3 n (beta + 1) doubles in the window's
    // n x n slot of the matrix buffer
    const int64_t nn = h->win_n[w];
    if (beta >= 11 && beta <= kBandMaxBeta && h->cfg.solver != 1 && 3 * nn * (beta + 1) <= nn * nn) h->win_beta[w] = beta;
  }
#ifdef UBA_EMU
  h->win_beta.assign(nW, 0);  // the emulation only has the dense stand-in solver
#endif
  const size_t nfree = h->free_list_h.size();
  const size_t red = (size_t)h->w_red_off_h[nW];
  CU(h, h->d_free_cam.reserve(NC));
  CU(h, h->d_free_list.reserve(nfree));
  CU(h, h->d_w_free_off.reserve(nW + 1));
  CU(h, h->d_w_red_off.reserve(nW + 1));
  CU(h, h->d_w_beta.reserve(nW));
  CU(h, cudaMemcpyAsync(h->d_w_beta.p, h->win_beta.data(), sizeof(int32_t) * nW, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_free_cam.p, h->free_cam_h.data(), sizeof(int32_t) * NC, cudaMemcpyHostToDevice, h->stream));
  if (nfree) CU(h, cudaMemcpyAsync(h->d_free_list.p, h->free_list_h.data(), sizeof(int32_t) * nfree, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_w_free_off.p, h->w_free_off_h.data(), sizeof(int32_t) * (nW + 1), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_w_red_off.p, h->w_red_off_h.data(), sizeof(int64_t) * (nW + 1), cudaMemcpyHostToDevice, h->stream));
  // accumulator block: [Sacc | Bacc | vacc | zh | w_lin] [w_post] [w_max] [w_loc]
  // layout: [B | v | zh | w_lin] [S_acc of every window] [w_post | w_rmax] [w_max] [w_loc].  The per-camera tail comes FIRST so
  // that it is contiguous with the band of the first window: one allreduce per linearisation on a point-sharded c4.
  h->off_Bacc = 0;
  h->off_vacc = h->off_Bacc + (size_t)NC * 36;
  h->off_zh = h->off_vacc + (size_t)NC * 6;
  h->off_wlin = h->off_zh + (size_t)NC * 6;
  h->off_sacc = h->off_wlin + (size_t)nW * WL_COUNT;
  h->acc_sum1 = h->off_sacc + red;
  h->off_wpost = h->acc_sum1;
  h->off_wrmax = h->off_wpost + (size_t)nW * WP_COUNT;            // per-rank gradient max-norms (summed: every slot has one writer)
  h->off_wmax = h->off_wrmax + (h->comm ? (size_t)nW * h->n_ranks : 0);
  h->off_wloc = h->off_wmax + (size_t)nW;
  h->acc_total = h->off_wloc + (size_t)nW * WC_COUNT;
  CU(h, h->d_acc.reserve(h->acc_total));
  // what a point-sharded run has to sum over ranks after each linearisation: the Schur accumulators (only the band
  // of banded windows: 285 KB instead of 11 MB on c4) and the tail {B, v, Z h, cost}; adjacent ranges are merged
  h->sum_ranges.clear();
  auto add_range = [&](size_t off, size_t cnt) {
    if (!cnt) return;
    if (!h->sum_ranges.empty() && h->sum_ranges.back().first + h->sum_ranges.back().second == off) h->sum_ranges.back().second += cnt;
    else h->sum_ranges.emplace_back(off, cnt);
  };
  add_range(0, h->off_sacc);
  for (int w = 0; w < nW; w++) {
    const size_t nn = (size_t)h->win_n[w];
    add_range(h->off_sacc + (size_t)h->w_red_off_h[w], h->win_beta[w] > 0 ? nn * (size_t)(h->win_beta[w] + 1) : nn * nn);
  }
  // ... and what is zeroed before each linearisation: the same pieces plus the per-window scalars at the end
  h->zero_ranges = h->sum_ranges;
  if (!h->zero_ranges.empty() && h->zero_ranges.back().first + h->zero_ranges.back().second == h->off_wpost) h->zero_ranges.back().second += h->acc_total - h->off_wpost;
  else h->zero_ranges.emplace_back(h->off_wpost, h->acc_total - h->off_wpost);
  CU(h, h->d_A.reserve(red));
  CU(h, h->d_rhs.reserve(6 * nfree));
  h->L.off_Bacc = (int64_t)h->off_Bacc; h->L.off_vacc = (int64_t)h->off_vacc; h->L.off_zh = (int64_t)h->off_zh;
  h->L.off_wlin = (int64_t)h->off_wlin; h->L.off_sacc = (int64_t)h->off_sacc; h->L.sum_end = (int64_t)h->acc_sum1;
  h->L.off_wpost = (int64_t)h->off_wpost; h->L.off_wmax = (int64_t)h->off_wmax; h->L.off_wloc = (int64_t)h->off_wloc;
  h->L.total = (int64_t)h->acc_total;
  if (h->comm) {
    const int rc = peer_exchange(h);
    if (rc) return rc;
  }
  if (!h->peer_on) {
    CU(h, h->d_stop_local.reserve(1));
    CU(h, cudaMemsetAsync(h->d_stop_local.p, 0, sizeof(double), h->stream));
    h->d_stop_req = h->d_stop_local.p;
  }
  TT("prepare: free cams, band")
  // lineariser choice: 1 = generic only; otherwise the tiled kernel plus the generic one for leftovers
  h->use_tile = h->cfg.linearizer != 1;
#ifdef UBA_EMU
  if (getenv("UBA_EMU_PLAN")) build_tile_plan(h, fixed_frames);   // host-only look at the planner (UBA_TRACE prints it)
  h->use_tile = false;  // the tiled kernel needs real thread blocks
#endif
  if (h->use_tile) {
    build_tile_plan(h, fixed_frames);
    TT("prepare: tile plan")
    CU(h, h->d_parts.reserve(h->parts_h.size())); CU(h, h->d_tile_cams.reserve(h->tile_cams_h.size()));
    CU(h, h->d_pt_mask.reserve(std::max(h->NP, 1))); CU(h, h->d_gen_pts.reserve(h->gen_pts_h.size()));
    if (!h->parts_h.empty()) CU(h, cudaMemcpyAsync(h->d_parts.p, h->parts_h.data(), sizeof(TilePart) * h->parts_h.size(), cudaMemcpyHostToDevice, h->stream));
    if (!h->tile_cams_h.empty()) CU(h, cudaMemcpyAsync(h->d_tile_cams.p, h->tile_cams_h.data(), sizeof(int32_t) * h->tile_cams_h.size(), cudaMemcpyHostToDevice, h->stream));
    if (h->NP) CU(h, cudaMemcpyAsync(h->d_pt_mask.p, h->pt_mask_h.data(), sizeof(uint32_t) * h->NP, cudaMemcpyHostToDevice, h->stream));
    if (!h->gen_pts_h.empty()) CU(h, cudaMemcpyAsync(h->d_gen_pts.p, h->gen_pts_h.data(), sizeof(int32_t) * h->gen_pts_h.size(), cudaMemcpyHostToDevice, h->stream));
  }
  // pipelined solve (UBA_PIPE_SOLVE=1, experimental): one banded window on this GPU alone, every observed point in a slot part,
  // the two-sided cluster solver.  Correct (the parity suite passes with it) but not a win as the lineariser is scheduled
  // today: its first wave of ~300 parts runs for ~100 us and ends almost at once, so the rows at the two ends of the window —
  // the first ones the solver needs — are complete only 20-30 us before the whole pass is (scripts/pipe_timing.py), while the
  // completion hooks cost the lineariser 20 us: c4 0.42 ms per iteration against 0.32 ms.  It needs parts that complete
  // progressively from the ends (much smaller parts, hence a much cheaper flush) to pay.
  {
    static const bool want = [] { const char* e = getenv("UBA_PIPE_SOLVE"); return e && e[0] == '1'; }();
    static const bool c2_on = [] { const char* e = getenv("UBA_BAND_C2"); const char* b = getenv("UBA_BAND_BCR"); return !(e && e[0] == '0') && !(b && atoi(b) > 0); }();
    bool ok = want && c2_on && nW == 1 && !h->comm && h->use_tile && h->use_slot && h->gen_pts_h.empty() && !h->parts_h.empty() &&
              h->variant_off[2] == h->variant_off[kLinVariants];      // slot parts only
    if (ok) {
      const int beta = h->win_beta[0], n = h->win_n[0];
      ok = beta >= 11 && beta <= 35 && n >= 12 * (beta + 1);
    }
#ifdef UBA_EMU
    ok = false;
#endif
    h->pipe_on = ok;
    if (ok) {
      CU(h, h->d_cam_expect.reserve(NC)); CU(h, h->d_pipe.reserve((size_t)NC + nfree + 1));
      CU(h, cudaMemcpyAsync(h->d_cam_expect.p, h->cam_expect_h.data(), sizeof(int32_t) * NC, cudaMemcpyHostToDevice, h->stream));
    }
  }
  CU(h, cudaStreamSynchronize(h->stream));
  TT("prepare: h2d + sync")
  DevView& V = h->V;
  V.pipe_on = 0;                               // switched on in the copies the pipelined launches get (run_iteration)
  V.cam_expect = h->d_cam_expect.p; V.cam_done = h->d_pipe.p; V.row_ready = h->d_pipe.p ? h->d_pipe.p + NC : nullptr;
  V.tile_threads = h->tile_threads;
  V.parts = h->d_parts.p; V.n_parts = h->use_tile ? (int)h->parts_h.size() : 0; V.tile_cams = h->d_tile_cams.p; V.pt_mask = h->d_pt_mask.p;
  V.gen_pts = h->d_gen_pts.p; V.n_gen = h->use_tile ? (int)h->gen_pts_h.size() : 0;
  V.w_beta = h->d_w_beta.p;
  V.free_cam = h->d_free_cam.p; V.free_list = h->d_free_list.p; V.w_free_off = h->d_w_free_off.p; V.w_red_off = h->d_w_red_off.p;
  V.Sacc = h->d_acc.p + h->off_sacc; V.Bacc = h->d_acc.p + h->off_Bacc; V.vacc = h->d_acc.p + h->off_vacc; V.zh = h->d_acc.p + h->off_zh;
  V.w_lin = h->d_acc.p + h->off_wlin; V.w_post = h->d_acc.p + h->off_wpost; V.w_max = h->d_acc.p + h->off_wmax;
  V.w_loc = h->d_acc.p + h->off_wloc;
  V.A = h->d_A.p; V.rhs = h->d_rhs.p;
  // consumer view: what assemble / solve / epilogue / controller read.  Without peers it IS the producer view (NCCL sums in
  // place); with peers the sums over ranks live in d_acc_red (same layout) and the local partials stay untouched for the
  // other ranks to pull.
  refresh_consumer_view(h);
  // The captured iteration graph stays valid when nothing it bakes in has changed: the host-side launch geometry
  // (checked here) and the device view passed to every kernel by value (checked at launch, run_iteration_fast).
  // A sliding window re-submitted with the same shape (the per-frame case) then skips capture + instantiation.
  {
    std::vector<char> sig;
    auto put = [&](const void* ptr, size_t n) { const char* c = (const char*)ptr; sig.insert(sig.end(), c, c + n); };
    put(h->variant_off, sizeof(h->variant_off)); put(&h->max_n, sizeof(h->max_n));
    put(&h->acc_total, sizeof(h->acc_total)); put(&h->use_tile, sizeof(h->use_tile)); put(&h->use_slot, sizeof(h->use_slot));
    put(&h->peer_on, sizeof(h->peer_on)); put(&h->pipe_on, sizeof(h->pipe_on));
    if (nW) { put(h->win_n.data(), sizeof(int) * nW); put(h->win_beta.data(), sizeof(int) * nW); }
    if (sig != h->graph_sig) { drop_graph(h); h->graph_sig.swap(sig); }
  }
  h->prepared_fixed = fixed_frames;
  return UBA_OK;
}

int allreduce(uba_handle* h, double* buf, size_t count, int op) {
  if (!h->comm || count == 0) return UBA_OK;
  const int rc = h->nccl.AllReduce(buf, buf, count, kNcclDouble, op, h->comm, h->stream);
  if (rc != 0) return fail(h, UBA_ERR_NCCL, "ncclAllReduce failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(rc) : "?");
  return UBA_OK;
}

double env_seconds(const char* name, double dflt) {
  const char* e = std::getenv(name);
  if (!e || !*e) return dflt;
  const double v = std::atof(e);
  return v > 0.0 ? v : dflt;
}

// Waits for the handle's stream with a bound: a rank whose peers never show up (mismatched call sequences, a dead
// process) gets UBA_ERR_NCCL after UBA_COMM_TIMEOUT_S seconds (default 120) instead of spinning inside a collective
// until some watchdog kills the job; asynchronous NCCL errors are polled meanwhile.
int comm_wait(uba_handle* h) {
  if (!h->comm) {
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    return e == cudaSuccess ? UBA_OK : fail(h, UBA_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
  }
  static const double limit = env_seconds("UBA_COMM_TIMEOUT_S", 120.0);
  const auto t0 = std::chrono::steady_clock::now();
  for (int spin = 0;; spin++) {
    const cudaError_t q = cudaStreamQuery(h->stream);
    if (q == cudaSuccess) return UBA_OK;
    if (q != cudaErrorNotReady) return fail(h, UBA_ERR_CUDA, "stream failed while waiting for the other ranks: %s", cudaGetErrorString(q));
    if ((spin & 255) == 255) {
      int async = 0;
      if (h->nccl.CommGetAsyncError && h->nccl.CommGetAsyncError(h->comm, &async) == 0 && async != 0) {
        if (h->nccl.CommAbort) h->nccl.CommAbort(h->comm);
        h->comm = nullptr; h->peer_on = false;
        return fail(h, UBA_ERR_NCCL, "NCCL reported an asynchronous error: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(async) : "?");
      }
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit) {
        if (h->nccl.CommAbort) h->nccl.CommAbort(h->comm);
        h->comm = nullptr; h->peer_on = false;
        return fail(h, UBA_ERR_NCCL, "rank %d waited %.0f s for the other ranks (mismatched call sequence?): communicator aborted", h->rank, limit);
      }
    }
  }
}

// every rank's `bytes` bytes, rank-major, into `all` (set-up exchanges only)
int allgather_bytes(uba_handle* h, const void* mine, size_t bytes, std::vector<char>& all) {
  const size_t n = (size_t)h->n_ranks;
  all.resize(n * bytes);
  if (!h->nccl.AllGather) return fail(h, UBA_ERR_NCCL, "libnccl lacks ncclAllGather");
  CU(h, h->d_xchg.reserve((n + 1) * bytes));
  CU(h, cudaMemcpyAsync(h->d_xchg.p + n * bytes, mine, bytes, cudaMemcpyHostToDevice, h->stream));
  const int rc = h->nccl.AllGather(h->d_xchg.p + n * bytes, h->d_xchg.p, bytes, kNcclChar, h->comm, h->stream);
  if (rc != 0) return fail(h, UBA_ERR_NCCL, "ncclAllGather failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(rc) : "?");
  CU(h, cudaMemcpyAsync(all.data(), h->d_xchg.p, n * bytes, cudaMemcpyDeviceToHost, h->stream));
  return comm_wait(h);
}

void peer_close(uba_handle* h) {
#ifndef UBA_EMU
  for (int r = 0; r < kMaxRanks; r++) {
    if (h->mapped_acc[r]) { cudaIpcCloseMemHandle(h->mapped_acc[r]); h->mapped_acc[r] = nullptr; }
    if (h->mapped_ctl[r]) { cudaIpcCloseMemHandle(h->mapped_ctl[r]); h->mapped_ctl[r] = nullptr; }
  }
#endif
  h->exchanged_acc = h->exchanged_ctl = nullptr;
  h->peer_on = false;
}

// Maps every rank's accumulator block and control block into this process (CUDA IPC) and fills the PeerView.  Called
// from prepare() on every rank of a point-sharded handle; re-done only when some rank's buffers moved.  When IPC is not
// available between two of the GPUs every rank falls back to the NCCL data path together.
int peer_exchange(uba_handle* h) {
#ifdef UBA_EMU
  h->peer_on = false;
  return UBA_OK;
#else
  const int n = h->n_ranks, nW = h->nW;
  if (!h->peer_wanted || n > kMaxRanks) { h->peer_on = false; return UBA_OK; }
  // control block layout
  const size_t off_flags = 0, off_epoch = off_flags + 8 * (size_t)n, off_err = off_epoch + 32, off_stop = off_err + 8,
               off_lo = off_stop + 8, off_hi = off_lo + 4 * (size_t)((n + 1) & ~1), off_inbox = off_hi + 4 * (size_t)((n + 1) & ~1);
  const size_t inbox_doubles = (size_t)2 * n * nW * kInboxSlots;
  const size_t bytes = off_inbox + 8 * inbox_doubles;
  const bool ctl_grows = bytes > h->d_ctl.cap;
  if (ctl_grows) {
    if (h->peer_on) { CU(h, cudaStreamSynchronize(h->stream)); }
    CU(h, h->d_ctl.reserve(bytes + bytes / 2));
  }
  h->ctl_inbox_off = off_inbox; h->ctl_bytes = bytes;
  CU(h, h->d_acc_red.reserve(h->acc_total));
  // did anything move, on any rank?
  double moved = (h->exchanged_acc != (void*)h->d_acc.p || h->exchanged_ctl != (void*)h->d_ctl.p || !h->peer_on) ? 1.0 : 0.0;
  CU(h, cudaMemcpyAsync(h->d_cam_lam.p, &moved, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = allreduce(h, h->d_cam_lam.p, 1, kNcclMax);
  if (rc) return rc;
  CU(h, cudaMemcpyAsync(&moved, h->d_cam_lam.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  rc = comm_wait(h);
  if (rc) return rc;
  if (moved > 0.0) {
    // Every rank must be out of its previous exchanges before mappings are replaced: the allreduce above orders that.
    // The exchange counters restart from zero on EVERY rank (they are compared across ranks); the memset is
    // stream-ordered before the allgather below, so nobody leaves the allgather before every control block is clean.
    peer_close(h);
    CU(h, cudaMemsetAsync(h->d_ctl.p, 0, h->d_ctl.cap, h->stream));
    struct Pack { cudaIpcMemHandle_t acc, ctl; int ok; int pad; } mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.ok = cudaIpcGetMemHandle(&mine.acc, h->d_acc.p) == cudaSuccess && cudaIpcGetMemHandle(&mine.ctl, h->d_ctl.p) == cudaSuccess;
    if (!mine.ok) cudaGetLastError();
    std::vector<char> all;
    rc = allgather_bytes(h, &mine, sizeof(mine), all);
    if (rc) return rc;
    bool ok = true;
    for (int r = 0; r < n; r++) ok = ok && reinterpret_cast<const Pack*>(all.data())[r].ok;
    if (ok) {
      for (int r = 0; r < n && ok; r++) {
        if (r == h->rank) continue;
        const Pack& pk = reinterpret_cast<const Pack*>(all.data())[r];
        ok = cudaIpcOpenMemHandle(&h->mapped_acc[r], pk.acc, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&h->mapped_ctl[r], pk.ctl, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      }
      if (!ok) cudaGetLastError();
    }
    // all or nothing: one rank without mappings sends everybody to the NCCL path
    double bad = ok ? 0.0 : 1.0;
    CU(h, cudaMemcpyAsync(h->d_cam_lam.p, &bad, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = allreduce(h, h->d_cam_lam.p, 1, kNcclMax);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(&bad, h->d_cam_lam.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    rc = comm_wait(h);
    if (rc) return rc;
    if (bad > 0.0) {
      peer_close(h);
      h->peer_wanted = false;
      if (getenv("UBA_TRACE")) fprintf(stderr, "  [trace] rank %d: no CUDA IPC between the GPUs, using the NCCL data path\n", h->rank);
      return UBA_OK;
    }
    h->exchanged_acc = h->d_acc.p; h->exchanged_ctl = h->d_ctl.p;
    h->peer_on = true;
  }
  // camera ranges of the shards, local copy in the control block
  {
    std::vector<int32_t> lohi(2 * (size_t)((n + 1) & ~1), 0);
    for (int r = 0; r < n; r++) { lohi[r] = h->rank_cam_lo[r]; lohi[((n + 1) & ~1) + r] = h->rank_cam_hi[r]; }
    CU(h, cudaMemcpyAsync(h->d_ctl.p + off_lo, lohi.data(), lohi.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  PeerView& P = h->P;
  std::memset(&P, 0, sizeof(P));
  P.rank = h->rank; P.n_ranks = n;
  for (int r = 0; r < n; r++) {
    unsigned char* ctl = r == h->rank ? h->d_ctl.p : (unsigned char*)h->mapped_ctl[r];
    P.acc[r] = r == h->rank ? h->d_acc.p : (double*)h->mapped_acc[r];
    P.flags[r] = (unsigned long long*)(ctl + off_flags);
    P.inbox[r] = (double*)(ctl + off_inbox);
  }
  P.epoch = (unsigned long long*)(h->d_ctl.p + off_epoch);
  P.err = (int32_t*)(h->d_ctl.p + off_err);
  P.stop_req = (const double*)(h->d_ctl.p + off_stop);
  P.cam_lo = (const int32_t*)(h->d_ctl.p + off_lo);
  P.cam_hi = (const int32_t*)(h->d_ctl.p + off_hi);
  P.timeout_ns = (long long)(env_seconds("UBA_PEER_TIMEOUT_S", 20.0) * 1e9);
  h->d_stop_req = (double*)(h->d_ctl.p + off_stop);
  return UBA_OK;
#endif
}

struct PhaseTimer {
  uba_handle* h; int phase; bool on;
  PhaseTimer(uba_handle* h_, int phase_) : h(h_), phase(phase_), on(h_->profiling) { if (on) cudaEventRecord(h->ev[0], h->stream); }
  void stop() {
    if (!on) return;
    cudaEventRecord(h->ev[1], h->stream);
    cudaEventSynchronize(h->ev[1]);
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    double* dst[] = {&h->timing.linearize_ms, &h->timing.solve_ms, &h->timing.backsub_ms, &h->timing.update_ms, &h->timing.comm_ms};
    *dst[phase] += ms;
    on = false;
  }
};

// Debug outputs (residuals, W, C, ...) only exist in the generic kernel, so a parity dump runs it over
// every point and, when the tiled kernel is selected, runs the tiled kernel for the accumulators.
int launch_linearizers(uba_handle* h, const DebugOut& dbg) {
  const bool want_dbg = dbg.residuals || dbg.weights || dbg.C || dbg.W || dbg.grad_pts || dbg.lam_pts;
  if (!h->use_tile) return launch_lin_generic(h->V, dbg, false, h->stream);
  int n = 0;
  if (want_dbg) {
    // generic pass for the per-observation dumps, then discard what it accumulated
    n += launch_lin_generic(h->V, dbg, false, h->stream);
    cudaMemsetAsync(h->d_acc.p, 0, h->acc_total * sizeof(double), h->stream);
  }
  n += launch_lin_tiled(h->V, h->variant_off, h->stream);
  DebugOut none{};
  n += launch_lin_generic(h->V, none, true, h->stream);
  return n;
}

// the linearise + Schur pass (the roofline kernel of the hot path)
int zero_accumulators(uba_handle* h);

int run_linearize(uba_handle* h, const DebugOut& dbg) {
  { const int rcz = zero_accumulators(h); if (rcz) return rcz; }
  PhaseTimer t(h, 0);
  h->timing.kernel_launches += launch_linearizers(h, dbg);
  h->timing.linearize_launches++;
  t.stop();
  if (h->comm) {
    PhaseTimer tc(h, 4);
#ifndef UBA_EMU
    if (h->peer_on) {
      // exchange 1 over NVLink peer memory: pull + sum the other ranks' partials into the consumer copy
      h->timing.kernel_launches += launch_peer_reduce(h->V, h->P, h->L, h->d_acc_red.p, h->dense_override ? 1 : 0, h->stream);
    } else
#endif
    if (h->dense_override) {
      int rc = allreduce(h, h->d_acc.p, h->acc_sum1, kNcclSum);
      if (rc) return rc;
    } else {
      for (const auto& r : h->sum_ranges) {
        int rc = allreduce(h, h->d_acc.p + r.first, r.second, kNcclSum);
        if (rc) return rc;
      }
    }
    tc.stop();
  }
  return UBA_OK;
}

// the zeroing every linearisation starts with
int zero_accumulators(uba_handle* h) {
  // for banded windows only the band of the Schur accumulator is ever touched
  if (!h->dense_override && h->zero_ranges.size() <= 4) {
    for (const auto& r : h->zero_ranges) CU(h, cudaMemsetAsync(h->d_acc.p + r.first, 0, r.second * sizeof(double), h->stream));
  } else {
    CU(h, cudaMemsetAsync(h->d_acc.p, 0, h->acc_total * sizeof(double), h->stream));
  }
  return UBA_OK;
}

int run_iteration(uba_handle* h) {
  DebugOut none{};
  int rc;
#ifndef UBA_EMU
  if (h->pipe_on && !h->profiling && !h->dense_override) {
    // Pipelined: [zero] -> { band solver (launched FIRST, so that its two CTAs are resident) | lineariser, whose CTAs
    // assemble the reduced system camera by camera and raise the row flags the solver waits on } -> epilogue ...
    rc = zero_accumulators(h);
    if (rc) return rc;
    CU(h, cudaMemsetAsync(h->d_pipe.p, 0, sizeof(int32_t) * ((size_t)h->NC + h->free_list_h.size()), h->stream));
    DevView Vp = h->V; Vp.pipe_on = 1;
    DevView Vcp = h->Vc; Vcp.pipe_on = 1;
    CU(h, cudaEventRecord(h->ev_fork, h->stream));
    CU(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    h->timing.kernel_launches += launch_solve(Vcp, h->win_n.data(), h->win_beta.data(), solve_small_limit(), h->stream2, false, 1);
    h->timing.kernel_launches += launch_lin_tiled(Vp, h->variant_off, h->stream);
    h->timing.linearize_launches++;
    CU(h, cudaEventRecord(h->ev_join, h->stream2));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->timing.kernel_launches += launch_solve(h->Vc, h->win_n.data(), h->win_beta.data(), solve_small_limit(), h->stream, false, 2);
  } else
#endif
  {
    rc = run_linearize(h, none);
    if (rc) return rc;
    PhaseTimer t(h, 1);
    h->timing.kernel_launches += launch_assemble(h->Vc, h->max_n, h->stream);
    h->timing.kernel_launches += launch_solve(h->Vc, h->win_n.data(), h->win_beta.data(), solve_small_limit(), h->stream);
    t.stop();
  }
  {
    PhaseTimer t(h, 2);
    h->timing.kernel_launches += launch_backsub(h->V, h->stream);   // producer view: its partial sums go to the local block
    t.stop();
  }
  if (h->comm) {
    PhaseTimer tc(h, 4);
#ifndef UBA_EMU
    if (h->peer_on) {
      // exchange 2: push {candidate cost, model change, norms, gradient max, stop request} into every rank's inbox
      h->timing.kernel_launches += launch_peer_post(h->V, h->Vc, h->P, h->stream);
    } else
#endif
    {
      // one collective for {candidate cost, model change, norms, stop requests} (sums) and the gradient max-norm: every rank
      // writes its max into its own slot of w_rmax, the slots are summed with the rest, the max over slots is taken locally
      h->timing.kernel_launches += launch_rank_max(h->d_acc.p + h->off_wmax, h->d_acc.p + h->off_wrmax, h->d_acc.p + h->off_wpost, h->d_stop_req, h->nW, h->rank, h->n_ranks, 0, h->stream);
      rc = allreduce(h, h->d_acc.p + h->off_wpost, (size_t)h->nW * (WP_COUNT + h->n_ranks), kNcclSum);
      if (rc) return rc;
      h->timing.kernel_launches += launch_rank_max(h->d_acc.p + h->off_wmax, h->d_acc.p + h->off_wrmax, h->d_acc.p + h->off_wpost, h->d_stop_req, h->nW, h->rank, h->n_ranks, 1, h->stream);
    }
    tc.stop();
  }
  {
    PhaseTimer t(h, 3);
    h->timing.kernel_launches += launch_lm_update(h->Vc, h->stream);
    t.stop();
  }
  return UBA_OK;
}

// One LM iteration, replayed from a CUDA graph when nothing needs host attention in between
// (no per-phase profiling, no NCCL on this handle).
int run_iteration_fast(uba_handle* h) {
#ifndef UBA_EMU
  // Point-sharded handles: with the peer-memory exchanges the iteration is kernels only and is captured like the
  // single-GPU one.  On the NCCL fallback path they launch directly: capturing the collectives works (UBA_COMM_GRAPH=1,
  // results identical) but measured SLOWER on 2 B200 (0.44 vs 0.36 ms per iteration)
  static const bool comm_graph = [] { const char* e = getenv("UBA_COMM_GRAPH"); return e && e[0] == '1'; }();
  // A solve of a few fixed iterations (the per-frame sliding window runs K = 4) launches its kernels directly: capturing and
  // instantiating the graph costs ~0.1 ms per problem, more than the launch gaps it removes from four iterations.
  static const int graph_min_iters = [] { const char* e = getenv("UBA_GRAPH_MIN_ITERS"); return e ? atoi(e) : 6; }();
  const bool few_iters = h->in_optimise && h->cfg.fixed_iterations > 0 && h->cfg.fixed_iterations < graph_min_iters && !h->graph_exec &&
                         !h->pipe_on;   // (uba_optimise only: the timing entry points always replay a graph, captured by their
                                        //  warm-up call; the pipelined solve is written for the graph's fork / join)
  if (!h->profiling && !few_iters && (!h->comm || h->peer_on || comm_graph)) {
    // the graph bakes the device views in by value: any change of them (sizes, pointers, solver settings) invalidates it
    if (h->graph_exec && (std::memcmp(&h->V, &h->graph_V, sizeof(DevView)) != 0 || std::memcmp(&h->Vc, &h->graph_Vc, sizeof(DevView)) != 0 ||
                          std::memcmp(&h->P, &h->graph_P, sizeof(PeerView)) != 0)) drop_graph(h);
    if (!h->graph_exec) {
      cudaGraph_t g = nullptr;
      std::memcpy(&h->graph_V, &h->V, sizeof(DevView)); std::memcpy(&h->graph_Vc, &h->Vc, sizeof(DevView));
      std::memcpy(&h->graph_P, &h->P, sizeof(PeerView));
      const int64_t before = h->timing.kernel_launches, lin_before = h->timing.linearize_launches;
      if (getenv("UBA_TRACE")) fprintf(stderr, "  [trace] capturing the iteration graph\n");
      CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = run_iteration(h);
      const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
      h->graph_kernels = h->timing.kernel_launches - before;
      h->timing.kernel_launches = before; h->timing.linearize_launches = lin_before;
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) return fail(h, UBA_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
      const cudaError_t e2 = cudaGraphInstantiate(&h->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (e2 != cudaSuccess) { h->graph_exec = nullptr; return fail(h, UBA_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e2)); }
    }
    CU(h, cudaGraphLaunch(h->graph_exec, h->stream));
    h->timing.kernel_launches += h->graph_kernels;
    h->timing.linearize_launches++;
    return UBA_OK;
  }
#endif
  return run_iteration(h);
}

int start_solve(uba_handle* h, int fixed_frames) {
  int rc = prepare(h, fixed_frames);
  if (rc) return rc;
  h->timing.kernel_launches += launch_init_state(h->V, h->cfg.initial_radius, h->stream);
  h->timing.kernel_launches += launch_cam_prep(h->V, 0, h->stream);
  return UBA_OK;
}

// restore the initial iterate on the device (parity 0) from the pinned staging copy
int upload_state(uba_handle* h) {
  CU(h, cudaMemcpyAsync(h->d_cams.p, h->h_cams.p, sizeof(double) * 6 * h->NC, cudaMemcpyHostToDevice, h->stream));
  if (h->NP) CU(h, cudaMemcpyAsync(h->d_pts.p, h->h_pts.p, sizeof(double) * 3 * h->NP, cudaMemcpyHostToDevice, h->stream));
  return UBA_OK;
}

int build_problem(uba_handle* h, int M, int nW, const int32_t* wc, const int32_t* wp, const int64_t* wo, const double* cams6,
                  const double* pts3, const double* feats, const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                  const uba_calib* calib) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (M != 2 && M != 4) return fail(h, UBA_ERR_INVALID_ARGUMENT, "M must be 2 or 4 (got %d)", M);
  if (nW <= 0 || !wc || !wp || !wo || !calib) return fail(h, UBA_ERR_INVALID_ARGUMENT, "null or empty window tables");
  const int NC = wc[nW], NP = wp[nW];
  const int64_t NO = wo[nW];
  if (NC <= 0 || NP < 0 || NO < 0) return fail(h, UBA_ERR_INVALID_ARGUMENT, "negative or empty sizes");
  if (NO >= (int64_t)INT32_MAX) return fail(h, UBA_ERR_UNSUPPORTED, "more than 2^31-1 observations in one handle");
  if (!cams6 || (NP && !pts3) || (NO && (!feats || !cam_idx || !pt_idx))) return fail(h, UBA_ERR_INVALID_ARGUMENT, "null input array");
  if (!(calib->feat_var > 0.0) || !(calib->fx0 > 0.0) || !(calib->fy0 > 0.0) || !(calib->cx0 > 0.0))
    return fail(h, UBA_ERR_INVALID_ARGUMENT, "calibration must have positive fx, fy, cx and feat_var");
  if (M == 4 && calib->baseline == 0.0) return fail(h, UBA_ERR_INVALID_ARGUMENT, "stereo BA needs a non-zero baseline (BundleAdjuster.h:147)");
  auto tt_ = std::chrono::steady_clock::now();
  h->state = 0; h->prepared_fixed = -1;
  h->M = M; h->nW = nW; h->NC = NC; h->NP = NP; h->NO = NO; h->calib_in = *calib;
  h->w_cam_off.assign(wc, wc + nW + 1); h->w_pt_off.assign(wp, wp + nW + 1); h->w_obs_off.assign(wo, wo + nW + 1);
  // fully overwritten below: resize only (no zero fill of ~10 MB per call)
  h->obs_order.resize(NO); h->pt_obs_off_caller.resize((size_t)NP + 1); h->pt_order.resize(NP);
  h->pt_obs_off_int.resize((size_t)NP + 1);
  h->pt_obs_off_caller[0] = 0; h->pt_obs_off_int[0] = 0;
  h->cam_seen.assign(NC, 0); h->cam_win_h.resize(NC); h->pt_win_h.resize(NP);
  TT("alloc tables")
  // validate + count (parallel inside a window when there are few windows)
  int bad = 0;
  const bool few = nW < 8;
  h->sc_cnt.resize((size_t)NP);
  h->sc_first.resize((size_t)NP + nW + 1);
  std::vector<int32_t>& cnt = h->sc_cnt;
  std::vector<char> win_canonical(nW, 0);
  for (int w = 0; w < nW; w++) {
    const int nc = wc[w + 1] - wc[w], np = wp[w + 1] - wp[w];
    if (nc < 0 || np < 0 || wo[w + 1] < wo[w]) bad++;
    for (int c = wc[w]; c < wc[w + 1]; c++) h->cam_win_h[c] = w;
  }
  if (bad) return fail(h, UBA_ERR_INVALID_ARGUMENT, "malformed window offsets");
  auto count_window = [&](int w, bool parallel) {
    const int nc = wc[w + 1] - wc[w], np = wp[w + 1] - wp[w];
    int32_t* c = cnt.data() + wp[w];
    char* seen = h->cam_seen.data() + wc[w];
    const int64_t o0 = wo[w], o1 = wo[w + 1];
    int badw = 0, unsorted = 0, unsorted_cam = 0;
#pragma omp parallel for schedule(static) reduction(+ : badw, unsorted, unsorted_cam) if (parallel)
    for (int64_t o = o0; o < o1; o++) {
      const int ci = cam_idx[o], pi = pt_idx[o];
      if (ci < 0 || ci >= nc || pi < 0 || pi >= np) { badw++; continue; }
      if (!seen[ci]) seen[ci] = 1;            // test first: 16 threads storing to the same few cache lines is a 10x slowdown
      if (o > o0) {
        if (pi < pt_idx[o - 1]) unsorted++;
        else if (pi == pt_idx[o - 1] && ci < cam_idx[o - 1]) unsorted_cam++;
      }
    }
    if (badw) return badw;
    win_canonical[w] = !unsorted && !unsorted_cam;   // already point-major and camera-ascending inside a point
    if (!unsorted) {
      // point-major input (the reference's order): counts are run lengths, no atomics needed
      int64_t* first = h->sc_first.data() + wp[w] + w;     // np + 1 slots of this window
#pragma omp parallel for schedule(static) if (parallel)
      for (int j = 0; j <= np; j++) first[j] = -1;
      const bool ident = win_canonical[w] != 0;
#pragma omp parallel for schedule(static) if (parallel)
      for (int64_t o = o0; o < o1; o++) {
        if (o == o0 || pt_idx[o] != pt_idx[o - 1]) first[pt_idx[o]] = o;
        if (ident) h->obs_order[o] = (int32_t)o;
      }
      first[np] = o1;
      // points without observations take the offset of the next observed point: a backward fill, done per chunk with the
      // first defined offset at or after each chunk's end handed down from the chunks behind it
      if (parallel && np > 65536) {
        const int T = std::min(omp_get_max_threads(), 64);
        int64_t head[65];
#pragma omp parallel for schedule(static) num_threads(T)
        for (int c = 0; c < T; c++) {
          const int b = (int)((int64_t)np * c / T), e = (int)((int64_t)np * (c + 1) / T);
          int64_t v = -1;
          for (int j = b; j < e; j++) if (first[j] >= 0) { v = first[j]; break; }
          head[c] = v;
        }
        head[T] = o1;
        for (int c = T - 1; c >= 0; c--) if (head[c] < 0) head[c] = head[c + 1];
#pragma omp parallel for schedule(static) num_threads(T)
        for (int c = 0; c < T; c++) {
          const int b = (int)((int64_t)np * c / T), e = (int)((int64_t)np * (c + 1) / T);
          int64_t carry = head[c + 1];
          for (int j = e - 1; j >= b; j--) { if (first[j] < 0) first[j] = carry; else carry = first[j]; }
        }
      } else {
        for (int j = np - 1; j >= 0; j--) if (first[j] < 0) first[j] = first[j + 1];
      }
#pragma omp parallel for schedule(static) if (parallel)
      for (int j = 0; j < np; j++) c[j] = (int32_t)(first[j + 1] - first[j]);
    } else {
#pragma omp parallel for schedule(static) if (parallel)
      for (int j = 0; j < np; j++) c[j] = 0;
#pragma omp parallel for schedule(static) if (parallel)
      for (int64_t o = o0; o < o1; o++) {
#pragma omp atomic
        c[pt_idx[o]]++;
      }
    }
    return 0;
  };
  if (few) { for (int w = 0; w < nW; w++) bad += count_window(w, true); }
  else {
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad)
    for (int w = 0; w < nW; w++) bad += count_window(w, false);
  }
  if (bad) return fail(h, UBA_ERR_INVALID_ARGUMENT, "camIdx / ptIdx out of range (%d offences)", bad);
  h->local_cam_lo = NC; h->local_cam_hi = -1;
  for (int c = 0; c < NC; c++) if (h->cam_seen[c]) { h->local_cam_lo = std::min(h->local_cam_lo, c); h->local_cam_hi = std::max(h->local_cam_hi, c); }
  TT("validate+count: passes")
  if (NP > 65536) {
    // exclusive scan of the counts in chunks: per-chunk totals, their running sum, then every chunk from its own base
    const int T = std::min(omp_get_max_threads(), 64);
    int64_t base[65];
#pragma omp parallel for schedule(static) num_threads(T)
    for (int c = 0; c < T; c++) {
      const int b = (int)((int64_t)NP * c / T), e = (int)((int64_t)NP * (c + 1) / T);
      int64_t sum = 0;
      for (int j = b; j < e; j++) sum += cnt[j];
      base[c + 1] = sum;
    }
    base[0] = 0;
    for (int c = 0; c < T; c++) base[c + 1] += base[c];
#pragma omp parallel for schedule(static) num_threads(T)
    for (int c = 0; c < T; c++) {
      const int b = (int)((int64_t)NP * c / T), e = (int)((int64_t)NP * (c + 1) / T);
      int64_t run = base[c];
      for (int j = b; j < e; j++) { run += cnt[j]; h->pt_obs_off_caller[j + 1] = run; }
    }
  } else {
    for (int j = 0; j < NP; j++) h->pt_obs_off_caller[j + 1] = h->pt_obs_off_caller[j] + cnt[j];
  }
  TT("validate+count")
  // Canonical order: point-major (stable), camera-ascending inside a point.  The reference's own
  // initialiseObservations already produces it (BundleAdjuster.h:364-374): detect that and skip the sort.
  h->sc_lo.resize((size_t)NP); h->sc_hi.resize((size_t)NP); h->sc_contig.resize((size_t)NP);
  std::vector<int32_t>& lo_c = h->sc_lo; std::vector<int32_t>& hi_c = h->sc_hi;   // caller point order; INT_MAX / -1 without observations
  std::vector<char>& contig_c = h->sc_contig;
  int pt_permuted = 0;                        // windows whose internal point order differs from the caller's
  int obs_permuted = 0;                       // windows whose observations had to be reordered
  auto order_window = [&](int w, bool parallel) {
    const int p0 = wp[w], np = wp[w + 1] - wp[w];
    const int64_t o0 = wo[w], o1 = wo[w + 1];
    const int unsorted = win_canonical[w] ? 0 : 1;   // canonical windows got the identity order while counting
    if (unsorted) {
#pragma omp atomic
      obs_permuted++;
      std::vector<int64_t> fill(np);
      for (int j = 0; j < np; j++) fill[j] = h->pt_obs_off_caller[p0 + j];
      for (int64_t o = o0; o < o1; o++) h->obs_order[fill[pt_idx[o]]++] = (int32_t)o;
    }
#pragma omp parallel for schedule(static) if (parallel)
    for (int j = 0; j < np; j++) {
      int32_t* b = &h->obs_order[h->pt_obs_off_caller[p0 + j]];
      int32_t* e = &h->obs_order[h->pt_obs_off_caller[p0 + j + 1]];
      if (unsorted) {
        // insertion sort by camera (stable; tracks are short)
        for (int32_t* i = b + 1; i < e; i++) {
          const int32_t v = *i; int32_t* k = i;
          while (k > b && cam_idx[*(k - 1)] > cam_idx[v]) { *k = *(k - 1); k--; }
          *k = v;
        }
      }
      if (b < e) {
        lo_c[p0 + j] = cam_idx[*b]; hi_c[p0 + j] = cam_idx[*(e - 1)];
        char cg = 1;
        for (int32_t* i = b + 1; i < e; i++) if (cam_idx[*i] != cam_idx[*(i - 1)] + 1) cg = 0;
        contig_c[p0 + j] = cg;
      } else {
        lo_c[p0 + j] = INT_MAX; hi_c[p0 + j] = -1; contig_c[p0 + j] = 1;
      }
    }
    // internal point order: stable by (lowest camera, highest camera), unobserved points last
    const int nc = wc[w + 1] - wc[w];
    const int64_t nkeys = (int64_t)nc * nc + 1;
    auto key_of = [&](int j) -> int64_t { return hi_c[p0 + j] < 0 ? nkeys - 1 : (int64_t)lo_c[p0 + j] * nc + hi_c[p0 + j]; };
    int key_unsorted = 0;
#pragma omp parallel for schedule(static) reduction(+ : key_unsorted) if (parallel)
    for (int j = 1; j < np; j++) if (key_of(j) < key_of(j - 1)) key_unsorted++;
    if (!key_unsorted) {
      // tracks already arrive grouped by (first, last) keyframe: the internal order is the caller's
#pragma omp parallel for schedule(static) if (parallel)
      for (int j = 0; j < np; j++) { h->pt_order[p0 + j] = p0 + j; h->pt_win_h[p0 + j] = w; }
    } else if (nkeys <= 8 * (int64_t)np + 1024) {
      // stable counting sort, parallel over contiguous chunks of points.  Keys are compressed to
      // lo * span + (hi - lo) so that the per-thread histograms stay small when tracks are short.
      int span = 1;
#pragma omp parallel for schedule(static) reduction(max : span) if (parallel)
      for (int j = 0; j < np; j++) if (hi_c[p0 + j] >= 0) span = std::max(span, hi_c[p0 + j] - lo_c[p0 + j] + 1);
      const int64_t nk = (int64_t)nc * span + 1;   // last key: unobserved points
      auto ckey = [&](int j) -> int64_t { return hi_c[p0 + j] < 0 ? nk - 1 : (int64_t)lo_c[p0 + j] * span + (hi_c[p0 + j] - lo_c[p0 + j]); };
      std::vector<int32_t> hist_local;
      std::vector<int32_t>& hist = parallel ? h->sc_hist : hist_local;   // windows sorted concurrently need their own
      int T = 1;
#pragma omp parallel if (parallel)
      {
#pragma omp single
        { T = omp_get_num_threads(); hist.assign((size_t)T * nk, 0); }
        const int t = omp_get_thread_num();
        const int b = (int)((int64_t)np * t / T), e = (int)((int64_t)np * (t + 1) / T);
        int32_t* mine = hist.data() + (size_t)t * nk;
        for (int j = b; j < e; j++) mine[ckey(j)]++;
#pragma omp barrier
#pragma omp single
        {
          int32_t run = 0;
          for (int64_t k2 = 0; k2 < nk; k2++)
            for (int q = 0; q < T; q++) { const int32_t c2 = hist[(size_t)q * nk + k2]; hist[(size_t)q * nk + k2] = run; run += c2; }
        }
        for (int j = b; j < e; j++) h->pt_order[p0 + mine[ckey(j)]++] = p0 + j;
      }
    } else {
      std::vector<int32_t> ids(np);
      std::iota(ids.begin(), ids.end(), 0);
      std::stable_sort(ids.begin(), ids.end(), [&](int a2, int b2) {
        const int la = hi_c[p0 + a2] < 0 ? INT_MAX : lo_c[p0 + a2], lb = hi_c[p0 + b2] < 0 ? INT_MAX : lo_c[p0 + b2];
        if (la != lb) return la < lb;
        return hi_c[p0 + a2] < hi_c[p0 + b2];
      });
      for (int j = 0; j < np; j++) h->pt_order[p0 + j] = p0 + ids[j];
    }
    if (key_unsorted) {
      for (int j = 0; j < np; j++) h->pt_win_h[p0 + j] = w;
#pragma omp atomic
      pt_permuted++;
    }
  };
  if (few) { for (int w = 0; w < nW; w++) order_window(w, true); }
  else {
#pragma omp parallel for schedule(dynamic, 16)
    for (int w = 0; w < nW; w++) order_window(w, false);
  }
  TT("order+sort")
  // internal CSR + observation slots (window ranges coincide with the caller's)
  h->pt_lo.resize(NP); h->pt_hi.resize(NP); h->pt_contig.resize(NP);
  {
    // prefix sum of the permuted track lengths, in parallel (per-thread partial sums, then a second sweep)
    const int nth = std::max(1, omp_get_max_threads());
    std::vector<int64_t> part((size_t)nth + 1, 0);
    const bool ident = pt_permuted == 0;
#pragma omp parallel num_threads(nth)
    {
      const int t = omp_get_thread_num(), T = omp_get_num_threads();
      const int b = (int)((int64_t)NP * t / T), e = (int)((int64_t)NP * (t + 1) / T);
      int64_t sum = 0;
      // first sweep: the gathers through the point order (prefetched: they land at random in the caller's arrays), the
      // track length parked in its slot; second sweep: sequential
      for (int s = b; s < e; s++) {
        if (!ident && s + 16 < e) {
          const int j2 = h->pt_order[s + 16];
          __builtin_prefetch(&cnt[j2]); __builtin_prefetch(&hi_c[j2]); __builtin_prefetch(&lo_c[j2]); __builtin_prefetch(&contig_c[j2]);
        }
        const int j = ident ? s : h->pt_order[s];
        const int32_t len = cnt[j];
        sum += len;
        h->pt_obs_off_int[s + 1] = len;
        h->pt_lo[s] = hi_c[j] < 0 ? -1 : lo_c[j]; h->pt_hi[s] = hi_c[j]; h->pt_contig[s] = contig_c[j];
      }
      part[t + 1] = sum;
#pragma omp barrier
#pragma omp single
      for (int q = 0; q < T; q++) part[q + 1] += part[q];
      int64_t run = part[t];
      for (int s = b; s < e; s++) { run += h->pt_obs_off_int[s + 1]; h->pt_obs_off_int[s + 1] = (int32_t)run; }
    }
  }
  TT("internal csr")
  // staging in internal order (pinned); a previous call's uploads from these buffers must have drained
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, h->h_cams.reserve((size_t)NC * 6));
  CU(h, h->h_pts.reserve((size_t)NP * 3));
  CU(h, h->h_feat.reserve((size_t)NO * M));
  CU(h, h->h_obs_cam.reserve((size_t)NO));
  CU(h, h->h_obs_internal.reserve((size_t)NO));
  std::memcpy(h->h_cams.p, cams6, sizeof(double) * 6 * NC);
  TT("pinned reserve")
  // observed points outside the box make the start infeasible (Ceres refuses it): noted per window while staging
  const uba::Calib kc = make_calib(*calib, M, h->cfg.use_bounds != 0);
  const bool bounds = h->cfg.use_bounds != 0;
  h->win_infeasible.assign(nW, 0);
  // Feature rows are staged AS THEY ARE (one streaming copy into pinned memory) and permuted + transposed into the
  // SoA planes on the device (k_ingest_feats); the host only builds the slot -> caller-observation map.
  const bool canonical = pt_permuted == 0 && obs_permuted == 0;
  const bool obs_ident = obs_permuted == 0;
  // Feature coordinates come from float detections widened to double in the reference (cv::Point2f, BundleAdjuster.h:371):
  // when every value is exactly a float — checked here, in the same pass — the rows are staged and uploaded as float32
  // (half the bytes) and widened on the device, bit for bit the same doubles.  Anything else is staged as doubles.
  bool feat_f32 = true;
  {
    const size_t nd = (size_t)NO * M, chunk = (size_t)1 << 16;
    const int64_t nchunks = (int64_t)((nd + chunk - 1) / chunk);
    float* hf32 = reinterpret_cast<float*>(h->h_feat.p);
    int not_float = 0;
#pragma omp parallel for schedule(static) reduction(+ : not_float)
    for (int64_t c = 0; c < nchunks; c++) {
      const size_t b = (size_t)c * chunk, e = std::min(nd, b + chunk);
      int bad_here = 0;
      for (size_t i = b; i < e; i++) { const float v = (float)feats[i]; hf32[i] = v; bad_here |= ((double)v != feats[i]); }
      not_float += bad_here;
    }
    if (not_float) {
      feat_f32 = false;
#pragma omp parallel for schedule(static)
      for (int64_t c = 0; c < nchunks; c++) {
        const size_t b = (size_t)c * chunk, e = std::min(nd, b + chunk);
        std::memcpy(h->h_feat.p + b, feats + b, sizeof(double) * (e - b));
      }
    }
  }
  // the rows start their way to the device now (scratch: Zbuf is idle until the first iteration) and travel under the rest
  // of the staging; the device-side gather + transpose is enqueued with the other uploads below
  if (NO) {
    CU(h, h->d_Zbuf.reserve((size_t)NO * 18));
    CU(h, cudaMemcpyAsync(h->d_Zbuf.p, h->h_feat.p, (feat_f32 ? sizeof(float) : sizeof(double)) * NO * M, cudaMemcpyHostToDevice, h->stream));
  }
  TT("staging: features")
  if (canonical) {
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < NO; o++) {
      h->h_obs_internal.p[o] = (int32_t)o;
      h->h_obs_cam.p[o] = cam_idx[o] | ((cam_id && cam_id[o] != 0) ? (1 << 30) : 0);
    }
  }
  TT("staging: observation words")
  // (the internal order visits the caller's points, offsets and observation words at random: three prefetch stages, each
  // one dependent load further down the chain  point -> offset -> first observation -> its camera words)
  const int32_t* const pt_order_p = h->pt_order.data();
  const int64_t* const off_caller_p = h->pt_obs_off_caller.data();
  const int32_t* const obs_order_p = h->obs_order.data();
#pragma omp parallel for schedule(static)
  for (int s = 0; s < NP; s++) {
    if (s + 24 < NP) { const int j2 = pt_order_p[s + 24]; __builtin_prefetch(pts3 + (size_t)j2 * 3); __builtin_prefetch(off_caller_p + j2); }
    if (!canonical && !obs_ident && s + 16 < NP) __builtin_prefetch(obs_order_p + off_caller_p[pt_order_p[s + 16]]);
    if (!canonical && s + 8 < NP) {
      const int64_t a = off_caller_p[pt_order_p[s + 8]];
      if (a < NO) { const int64_t o2 = obs_ident ? a : obs_order_p[a]; __builtin_prefetch(cam_idx + o2); if (cam_id) __builtin_prefetch(cam_id + o2); }
    }
    const int j = h->pt_order[s];
    const double px = pts3[(size_t)j * 3], py = pts3[(size_t)j * 3 + 1], pz = pts3[(size_t)j * 3 + 2];
    h->h_pts.p[(size_t)s * 3] = px; h->h_pts.p[(size_t)s * 3 + 1] = py; h->h_pts.p[(size_t)s * 3 + 2] = pz;
    const int64_t src = h->pt_obs_off_caller[j];
    const int32_t dst = h->pt_obs_off_int[s];
    const int k = h->pt_obs_off_int[s + 1] - dst;
    if (bounds && k > 0 && !(px >= kc.lo[0] && px <= kc.hi[0] && py >= kc.lo[1] && py <= kc.hi[1] && pz >= kc.lo[2] && pz <= kc.hi[2])) {
      char& f = h->win_infeasible[h->pt_win_h[j]];
      if (!f) f = 1;
    }
    if (canonical) continue;
    for (int q = 0; q < k; q++) {
      const int32_t o = obs_ident ? (int32_t)(src + q) : h->obs_order[src + q];   // (only the points moved: no lookup)
      h->h_obs_internal.p[dst + q] = o;
      h->h_obs_cam.p[dst + q] = cam_idx[o] | ((cam_id && cam_id[o] != 0) ? (1 << 30) : 0);
    }
  }
  TT("staging: points")
  // device buffers
  CU(h, h->d_cams.reserve((size_t)NC * 12)); CU(h, h->d_camR.reserve((size_t)NC * kCamStride * 2));
  CU(h, h->d_cam_s2.reserve((size_t)NC * 6)); CU(h, h->d_cam_lam.reserve((size_t)NC * 6)); CU(h, h->d_cam_y.reserve((size_t)NC * 6));
  CU(h, h->d_pts.reserve((size_t)NP * 6)); CU(h, h->d_pt_s2.reserve((size_t)NP * 3)); CU(h, h->d_pt_rec.reserve((size_t)NP * kPtRec));
  CU(h, h->d_feat.reserve((size_t)NO * M)); CU(h, h->d_obs_cam.reserve((size_t)NO)); CU(h, h->d_obs_src.reserve((size_t)NO)); CU(h, h->d_Zbuf.reserve((size_t)NO * 18));
  CU(h, h->d_w_cam_off.reserve(nW + 1)); CU(h, h->d_w_pt_off.reserve(nW + 1)); CU(h, h->d_cam_win.reserve(NC));
  CU(h, h->d_pt_obs_off.reserve((size_t)NP + 1)); CU(h, h->d_pt_win.reserve(std::max(NP, 1))); CU(h, h->d_pt_order.reserve(std::max(NP, 1)));
  CU(h, h->d_ws.reserve(nW)); CU(h, h->d_n_active.reserve(1));
  const int rec_stride = std::max(h->cfg.max_iterations, h->cfg.fixed_iterations) + 2;
  CU(h, h->d_recs.reserve((size_t)nW * rec_stride));
  TT("device reserve")
  cudaStream_t st = h->stream;
  CU(h, cudaMemcpyAsync(h->d_w_cam_off.p, h->w_cam_off.data(), sizeof(int32_t) * (nW + 1), cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_w_pt_off.p, h->w_pt_off.data(), sizeof(int32_t) * (nW + 1), cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_cam_win.p, h->cam_win_h.data(), sizeof(int32_t) * NC, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_pt_obs_off.p, h->pt_obs_off_int.data(), sizeof(int32_t) * ((size_t)NP + 1), cudaMemcpyHostToDevice, st));
  if (NP) {
    CU(h, cudaMemcpyAsync(h->d_pt_win.p, h->pt_win_h.data(), sizeof(int32_t) * NP, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->d_pt_order.p, h->pt_order.data(), sizeof(int32_t) * NP, cudaMemcpyHostToDevice, st));
  }
  if (NO) {
    // the raw rows are already on their way (see the staging above): gather + transpose on the device
    CU(h, cudaMemcpyAsync(h->d_obs_src.p, h->h_obs_internal.p, sizeof(int32_t) * NO, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->d_obs_cam.p, h->h_obs_cam.p, sizeof(int32_t) * NO, cudaMemcpyHostToDevice, st));
    h->timing.kernel_launches += launch_ingest_feats(h->d_Zbuf.p, h->d_obs_src.p, h->d_feat.p, NO, M, feat_f32 ? 1 : 0, st);
  }
  int rc = upload_state(h);
  if (rc) return rc;
  CU(h, cudaMemsetAsync(h->d_recs.p, 0, sizeof(IterRec) * (size_t)nW * rec_stride, st));
  // no synchronisation here: the uploads come from pinned staging owned by the handle and are stream-ordered before
  // everything uba_optimise enqueues, so the host-side planning of prepare() overlaps them
  TT("h2d (enqueue)")
  fill_view_static(h);
  h->V.rec_stride = rec_stride;
  h->ws_h.assign(nW, WinState{});
  h->state = 1;
  h->device_dirty = false;
  h->host_obs_valid = true;
  h->tracks_valid = false;
  if (h->cfg.sliding_window && nW == 1 && !h->comm && obs_permuted == 0) {
    // every track a run of consecutive keyframes with one camID: the window can then be slid on the device
    int bad_track = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_track)
    for (int j = 0; j < NP; j++) {
      if (!contig_c[j]) { bad_track++; continue; }
      if (cam_id) for (int64_t o = h->pt_obs_off_caller[j] + 1; o < h->pt_obs_off_caller[j + 1]; o++) if ((cam_id[o] != 0) != (cam_id[o - 1] != 0)) bad_track++;
    }
    if (!bad_track) {
      h->tr_lo.resize(NP); h->tr_cnt.resize(NP); h->tr_cid.resize(NP);
      std::vector<int32_t> off32((size_t)NP + 1);
      for (int j = 0; j < NP; j++) {
        h->tr_cnt[j] = cnt[j]; h->tr_lo[j] = cnt[j] > 0 ? lo_c[j] : 0;
        h->tr_cid[j] = (cnt[j] > 0 && cam_id && cam_id[h->pt_obs_off_caller[j]] != 0) ? 1 : 0;
        off32[j] = (int32_t)h->pt_obs_off_caller[j];
      }
      off32[NP] = (int32_t)h->pt_obs_off_caller[NP];
      h->rows_cur = 0;
      CU(h, h->d_rows[0].reserve((size_t)NO * M + 1)); CU(h, h->d_tr_off[0].reserve((size_t)NP + 1));
      CU(h, cudaMemcpyAsync(h->d_tr_off[0].p, off32.data(), sizeof(int32_t) * ((size_t)NP + 1), cudaMemcpyHostToDevice, st));
      CU(h, cudaStreamSynchronize(st));       // off32 is a local
      // the raw rows are still in the scratch they were uploaded to (nothing has run on it since k_ingest_feats)
      h->timing.kernel_launches += launch_win_rows(h->d_Zbuf.p, feat_f32 ? 1 : 0, h->d_rows[0].p, (int64_t)NO * M, st);
      h->tracks_valid = true;
    }
  }
  return UBA_OK;
}

// obs_order / h_obs_internal / h_obs_cam are O(n_obs) host tables that only the parity dumps, uba_get_tables and the plan of
// non-contiguous tracks read; after uba_window_advance they are rebuilt on demand from the track table.
int ensure_host_obs_tables(uba_handle* h) {
  if (h->host_obs_valid) return UBA_OK;
  const int NP = h->NP; const int64_t NO = h->NO;
  h->obs_order.resize(NO);
  for (int64_t o = 0; o < NO; o++) h->obs_order[o] = (int32_t)o;
  CU(h, h->h_obs_internal.reserve((size_t)NO)); CU(h, h->h_obs_cam.reserve((size_t)NO));
  for (int s = 0; s < NP; s++) {
    const int j = h->pt_order[s];
    const int32_t dst = h->pt_obs_off_int[s];
    const int k = h->pt_obs_off_int[s + 1] - dst;
    for (int q = 0; q < k; q++) {
      h->h_obs_internal.p[dst + q] = (int32_t)(h->pt_obs_off_caller[j] + q);
      h->h_obs_cam.p[dst + q] = (h->tr_lo[j] + q) | (h->tr_cid[j] ? (1 << 30) : 0);
    }
  }
  h->host_obs_valid = true;
  return UBA_OK;
}

// Ceres rejects a problem whose bounded parameter blocks start outside their box
// ([CERES-UPSTREAM] Program::IsFeasible) -> the reference reports Status::FAILED.
bool window_feasible(const uba_handle* h, int w) { return !h->win_infeasible[w]; }   // decided while staging (build_problem)


// Pose covariances at the final iterate: one undamped linearisation, dense factorisation of the
// reduced camera matrix (the banded solver keeps no dense factor, so it is bypassed), S^-1 diagonal blocks.
int compute_covariances(uba_handle* h) {
  const int nW = h->nW, NC = h->NC;
  CU(h, h->d_cov.reserve((size_t)NC * 36));
  CU(h, cudaMemsetAsync(h->d_cov.p, 0, sizeof(double) * 36 * NC, h->stream));
  drop_graph(h);
  h->timing.kernel_launches += launch_cov_state(h->V, 1, h->stream);
  std::vector<int32_t> zeros(nW, 0);
  CU(h, cudaMemcpyAsync(h->d_w_beta.p, zeros.data(), sizeof(int32_t) * nW, cudaMemcpyHostToDevice, h->stream));
  DebugOut none{};
  const bool was_profiling = h->profiling;
  h->profiling = false;
  h->dense_override = true;
  int rc = run_linearize(h, none);
  h->dense_override = false;
  if (!rc) {
    h->timing.kernel_launches += launch_assemble(h->Vc, h->max_n, h->stream);
    h->timing.kernel_launches += launch_solve(h->Vc, h->win_n.data(), zeros.data(), solve_small_limit(), h->stream, true);
    h->timing.kernel_launches += launch_cov_blocks(h->Vc, (int)h->free_list_h.size(), h->max_n, h->d_cov.p, h->stream);
  }
  h->profiling = was_profiling;
  h->timing.kernel_launches += launch_cov_state(h->V, 0, h->stream);
  CU(h, cudaMemcpyAsync(h->d_w_beta.p, h->win_beta.data(), sizeof(int32_t) * nW, cudaMemcpyHostToDevice, h->stream));
  if (rc) return rc;
  h->cov_h.assign((size_t)NC * 36, 0.0);
  CU(h, cudaMemcpyAsync(h->cov_h.data(), h->d_cov.p, sizeof(double) * 36 * NC, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaGetLastError());
  return UBA_OK;
}

}  // namespace

extern "C" {

int uba_version(void) { return UBA_VERSION; }

const char* uba_last_error(const uba_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int uba_create(const uba_config* cfg, uba_handle** out) {
  if (!out) return fail(nullptr, UBA_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  uba_config c;
  if (cfg) c = *cfg; else uba_config_default(&c);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, UBA_ERR_CUDA, "no CUDA device (%s): libuba has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (c.device < 0 || c.device >= ndev) return fail(nullptr, UBA_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", c.device, ndev);
  e = cudaSetDevice(c.device);
  if (e != cudaSuccess) return fail(nullptr, UBA_ERR_CUDA, "cudaSetDevice(%d): %s", c.device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c.device);
  if (prop.major != 10) return fail(nullptr, UBA_ERR_CUDA, "device %d is sm_%d%d; libuba carries sm_100a code only", c.device, prop.major, prop.minor);
  uba_handle* h = new uba_handle();
  h->cfg = c; h->device = c.device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, UBA_ERR_CUDA, "cudaStreamCreate failed"); }
  for (auto& ev : h->ev) cudaEventCreate(&ev);
  {
    // the solver's branch gets the highest priority: its two CTAs must become resident while the lineariser's grid is
    // still queueing for the same SMs
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi) != cudaSuccess) cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  *out = h;
  return UBA_OK;
}

void uba_destroy(uba_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  drop_graph(h);
  peer_close(h);
#ifndef UBA_EMU
  uba_vo_free(h->vo);
#endif
  if (h->comm && h->nccl.CommDestroy) h->nccl.CommDestroy(h->comm);
  h->d_ctl.release(); h->d_acc_red.release(); h->d_xchg.release(); h->d_stop_local.release();
  h->h_cams.release(); h->h_pts.release(); h->h_feat.release(); h->h_out.release(); h->h_obs_cam.release(); h->h_obs_internal.release(); h->h_adv.release(); h->h_cams_prev.release();
  h->d_cams.release(); h->d_camR.release(); h->d_cam_s2.release(); h->d_cam_lam.release(); h->d_cam_y.release(); h->d_pts.release();
  h->d_pt_s2.release(); h->d_pt_rec.release(); h->d_feat.release(); h->d_acc.release(); h->d_A.release(); h->d_rhs.release();
  h->d_Zbuf.release(); h->d_dbg.release(); h->d_export.release(); h->d_flush.release();
  h->d_w_cam_off.release(); h->d_w_pt_off.release(); h->d_w_free_off.release(); h->d_free_list.release(); h->d_free_cam.release();
  h->d_cam_win.release(); h->d_pt_obs_off.release(); h->d_pt_win.release(); h->d_obs_cam.release(); h->d_obs_src.release(); h->d_n_active.release();
  h->d_pt_order.release(); h->d_w_red_off.release(); h->d_ws.release(); h->d_recs.release();
  for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  h->d_cam_expect.release(); h->d_pipe.release();
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int uba_set_batch(uba_handle* h, int M, int n_windows, const int32_t* win_cam_off, const int32_t* win_pt_off,
                  const int64_t* win_obs_off, const double* cams6, const double* pts3, const double* feats,
                  const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id, const uba_calib* calib) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  return build_problem(h, M, n_windows, win_cam_off, win_pt_off, win_obs_off, cams6, pts3, feats, cam_idx, pt_idx, cam_id, calib);
}

int uba_set_problem(uba_handle* h, int M, int n_cams, int n_pts, int n_obs, const double* cams6, const double* pts3,
                    const double* feats, const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                    const uba_calib* calib) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (n_cams <= 0 || n_pts <= 0) {
    // "[Bundle Adjuster] system should be uninitialised and both cameras and points not empty!" (BundleAdjuster.h:314-316)
    return fail(h, UBA_ERR_INVALID_ARGUMENT, "cameras and points must not be empty");
  }
  const int32_t wc[2] = {0, n_cams}, wp[2] = {0, n_pts};
  const int64_t wo[2] = {0, n_obs};
  cudaSetDevice(h->device);
  return build_problem(h, M, 1, wc, wp, wo, cams6, pts3, feats, cam_idx, pt_idx, cam_id, calib);
}

int uba_window_advance(uba_handle* h, int n_drop, int n_new_cams, const double* new_cams6, int n_new_pts, const double* new_pts3,
                       const int32_t* new_pt_cam_id, int n_new_obs, const double* feats, const int32_t* cam_idx, const int32_t* pt_idx,
                       const double* cams6_all, const double* pts3_all, int32_t* pt_id_map) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no window is resident");
  if (!h->tracks_valid)
    return fail(h, UBA_ERR_UNSUPPORTED, "uba_window_advance needs uba_config.sliding_window = 1 and a single window of contiguous, "
                                        "point-major tracks with one camID each: re-submit the window with uba_set_problem");
  const int NCo = h->NC, NPo = h->NP, M = h->M;
  const int NCn = NCo - n_drop + n_new_cams;
  if (n_drop < 0 || n_drop > NCo || n_new_cams < 0 || n_new_pts < 0 || n_new_obs < 0 || NCn <= 0 || (n_new_cams && !new_cams6 && !cams6_all) ||
      (n_new_pts && !new_pts3 && !pts3_all) || (n_new_obs && (!feats || !cam_idx || !pt_idx)))
    return fail(h, UBA_ERR_INVALID_ARGUMENT, "uba_window_advance: bad sizes or null arrays");
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  auto tt_ = std::chrono::steady_clock::now();
  // Every index table of the advance is built in ONE pinned block laid out like the device scratch (three uploads instead
  // of ten from pageable vectors), and the previous window's cameras start their way back before the host plans anything:
  // the call never waits for the stream.
  CU(h, cudaStreamSynchronize(st));            // the staging block may still feed the previous advance's uploads (idle in practice)
  const int par0 = h->state == 2 && h->ws_h[0].done != UBA_TERM_FAILURE && !h->device_dirty ? (h->ws_h[0].cur & 1) : 0;
  const bool fetch_cams = !cams6_all && !h->device_dirty;
  if (fetch_cams) {
    CU(h, h->h_cams_prev.reserve((size_t)NCo * 6));
    CU(h, cudaMemcpyAsync(h->h_cams_prev.p, h->d_cams.p + (size_t)par0 * NCo * 6, sizeof(double) * 6 * NCo, cudaMemcpyDeviceToHost, st));
    CU(h, cudaEventRecord(h->ev[6], st));
  }
  const int NPmax = NPo + n_new_pts;
  const size_t t_drop = 0, t_map = t_drop + NPo, t_pt = t_map + NPo, t_cam = t_pt + n_new_obs;       // known now
  const size_t stage_max = t_cam + n_new_obs + (size_t)5 * NPmax + 2 + (size_t)NPmax + 1 + (size_t)(NPmax + 3) / 4 + 4;
  CU(h, h->h_adv.reserve(stage_max));
  int32_t* const A = h->h_adv.p;
  // ---- 1. survivors ------------------------------------------------------------------------------------------------
  int32_t* const dropped = A + t_drop;
  int32_t* const id_map = A + t_map;
  int n_alive = 0;
  for (int j = 0; j < NPo; j++) {
    const int d = std::min(std::max(n_drop - h->tr_lo[j], 0), h->tr_cnt[j]);
    dropped[j] = d;
    id_map[j] = h->tr_cnt[j] - d > 0 ? n_alive++ : -1;
  }
  const int NPn = n_alive + n_new_pts;
  if (NPn <= 0) return fail(h, UBA_ERR_INVALID_ARGUMENT, "uba_window_advance: the new window has no points");
  const size_t t_src = t_cam + n_new_obs, t_lo = t_src + NPn, t_order = t_lo + NPn, t_offi = t_order + NPn, t_end = t_offi + NPn + 1;
  const size_t t_offn = t_end, t_cid = t_offn + NPn + 1;              // behind the device block: the caller-order CSR, the camID bytes
  int32_t* const src_slot = A + t_src;
  int32_t* const lo_n = A + t_lo;
  int32_t* const order = A + t_order;
  int32_t* const off_i = A + t_offi;
  int32_t* const off_n = A + t_offn;
  unsigned char* const cid_n = reinterpret_cast<unsigned char*>(A + t_cid);
  if (n_new_obs) { std::memcpy(A + t_pt, pt_idx, sizeof(int32_t) * n_new_obs); std::memcpy(A + t_cam, cam_idx, sizeof(int32_t) * n_new_obs); }
  std::vector<int32_t> cnt_n(NPn, 0), old_of(NPn, -1);
  std::fill(lo_n, lo_n + NPn, 0);
  std::memset(cid_n, 0, NPn);
  for (int j = 0; j < NPo; j++) {
    const int nj = id_map[j];
    if (nj < 0) continue;
    lo_n[nj] = std::max(h->tr_lo[j] - n_drop, 0); cnt_n[nj] = h->tr_cnt[j] - dropped[j]; cid_n[nj] = h->tr_cid[j]; old_of[nj] = j;
  }
  for (int k = 0; k < n_new_pts; k++) cid_n[n_alive + k] = (new_pt_cam_id && new_pt_cam_id[k] != 0) ? 1 : 0;
  // ---- 2. the new observations extend their tracks by consecutive keyframes ------------------------------------------
  {
    std::vector<int32_t> add(NPn, 0), minc(NPn, INT_MAX), maxc(NPn, -1);
    std::vector<int64_t> sumc(NPn, 0);
    const int first_new = NCo - n_drop;
    for (int i = 0; i < n_new_obs; i++) {
      const int j = pt_idx[i], c = cam_idx[i];
      if (j < 0 || j >= NPn || c < first_new || c >= NCn) return fail(h, UBA_ERR_INVALID_ARGUMENT, "uba_window_advance: observation %d has ptIdx %d / camIdx %d out of range", i, j, c);
      add[j]++; minc[j] = std::min(minc[j], c); maxc[j] = std::max(maxc[j], c); sumc[j] += c;
    }
    for (int j = 0; j < NPn; j++) {
      if (!add[j]) continue;
      const int64_t a = add[j];
      const bool run = maxc[j] - minc[j] + 1 == add[j] && sumc[j] == a * minc[j] + a * (a - 1) / 2;
      const bool joins = cnt_n[j] == 0 || minc[j] == lo_n[j] + cnt_n[j];
      if (!run || !joins) return fail(h, UBA_ERR_UNSUPPORTED, "uba_window_advance: the new observations of point %d do not extend its track by consecutive keyframes", j);
      if (cnt_n[j] == 0) lo_n[j] = minc[j];
      cnt_n[j] += add[j];
    }
  }
  off_n[0] = 0;
  for (int j = 0; j < NPn; j++) off_n[j + 1] = off_n[j] + cnt_n[j];
  const int64_t NOn = off_n[NPn];
  TT("advance: track tables")
  // ---- 3. internal point order: stable by (first keyframe, last keyframe), unobserved points last --------------------
  {
    int span = 1;
    for (int j = 0; j < NPn; j++) span = std::max(span, cnt_n[j]);
    const int64_t nk = (int64_t)NCn * span + 1;
    auto key = [&](int j) -> int64_t { return cnt_n[j] == 0 ? nk - 1 : (int64_t)lo_n[j] * span + (cnt_n[j] - 1); };
    std::vector<int32_t> hist((size_t)nk + 1, 0);
    for (int j = 0; j < NPn; j++) hist[key(j) + 1]++;
    for (int64_t k = 0; k < nk; k++) hist[k + 1] += hist[k];
    for (int j = 0; j < NPn; j++) order[hist[key(j)]++] = j;
  }
  std::vector<int32_t> old_slot(NPo, -1);
  for (int s = 0; s < NPo; s++) old_slot[h->pt_order[s]] = s;
  h->pt_lo.resize(NPn); h->pt_hi.resize(NPn); h->pt_contig.assign(NPn, 1);
  off_i[0] = 0;
  for (int s = 0; s < NPn; s++) {
    const int j = order[s];
    off_i[s + 1] = off_i[s] + cnt_n[j];
    h->pt_lo[s] = cnt_n[j] ? lo_n[j] : -1; h->pt_hi[s] = cnt_n[j] ? lo_n[j] + cnt_n[j] - 1 : -1;
    src_slot[s] = j < n_alive ? old_slot[old_of[j]] : -(j - n_alive) - 1;
  }
  TT("advance: order")
  // ---- 4. device: shift, append, pack ------------------------------------------------------------------------------
  const int cur = h->rows_cur, nxt = cur ^ 1;
  CU(h, h->d_rows[nxt].reserve((size_t)NOn * M + 1)); CU(h, h->d_tr_off[nxt].reserve((size_t)NPn + 1));
  CU(h, h->d_tr_tmp.reserve(t_end)); CU(h, h->d_tr_cid.reserve(NPn));
  CU(h, h->d_fresh.reserve((size_t)n_new_obs * M + (size_t)n_new_pts * 3 + 1));
  int32_t* T = h->d_tr_tmp.p;
  CU(h, cudaMemcpyAsync(T, A, sizeof(int32_t) * t_end, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_tr_off[nxt].p, off_n, sizeof(int32_t) * ((size_t)NPn + 1), cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_tr_cid.p, cid_n, NPn, cudaMemcpyHostToDevice, st));
  if (n_new_obs) CU(h, cudaMemcpyAsync(h->d_fresh.p, feats, sizeof(double) * (size_t)n_new_obs * M, cudaMemcpyHostToDevice, st));
  h->timing.kernel_launches += launch_win_shift(h->d_rows[cur].p, h->d_tr_off[cur].p, T + t_drop, T + t_map, h->d_tr_off[nxt].p, NPo, M, h->d_rows[nxt].p, st);
  h->timing.kernel_launches += launch_win_append(h->d_fresh.p, T + t_pt, T + t_cam, T + t_lo, h->d_tr_off[nxt].p, n_new_obs, M, h->d_rows[nxt].p, st);
  // points: what the last solve left (its accepted iterate), re-slotted, plus the newcomers — or the caller's values
  const int rec_stride = std::max(h->cfg.max_iterations, h->cfg.fixed_iterations) + 2;
  CU(h, h->h_pts.reserve((size_t)NPn * 3)); CU(h, h->h_cams.reserve((size_t)NCn * 6));
  const uba::Calib kc = make_calib(h->calib_in, M, h->cfg.use_bounds != 0);
  auto inside = [&](const double* p) { return p[0] >= kc.lo[0] && p[0] <= kc.hi[0] && p[1] >= kc.lo[1] && p[1] <= kc.hi[1] && p[2] >= kc.lo[2] && p[2] <= kc.hi[2]; };
  h->win_infeasible.assign(1, 0);
  if (h->device_dirty) { const int rcu = upload_state(h); if (rcu) return rcu; h->device_dirty = false; h->state = 1; }
  const int par = h->state == 2 && h->ws_h[0].done != UBA_TERM_FAILURE ? (h->ws_h[0].cur & 1) : 0;
  if (!cams6_all && !fetch_cams) {             // the benchmark's scratch was on the device: the staged state was just re-uploaded
    CU(h, h->h_cams_prev.reserve((size_t)NCo * 6));
    CU(h, cudaMemcpyAsync(h->h_cams_prev.p, h->d_cams.p + (size_t)par * NCo * 6, sizeof(double) * 6 * NCo, cudaMemcpyDeviceToHost, st));
    CU(h, cudaEventRecord(h->ev[6], st));
  }
  CU(h, h->d_pt_rec.reserve((size_t)std::max(NPn, NPo) * kPtRec));     // scratch for the re-slotted points
  if (pts3_all) {
    CU(h, cudaStreamSynchronize(st));          // h_pts may still feed an earlier upload
    for (int s = 0; s < NPn; s++) {
      const double* p = pts3_all + (size_t)order[s] * 3;
      std::memcpy(h->h_pts.p + (size_t)s * 3, p, sizeof(double) * 3);
      if (h->cfg.use_bounds && cnt_n[order[s]] > 0 && !inside(p)) h->win_infeasible[0] = 1;
    }
  } else {
    if (n_new_pts) {
      CU(h, cudaMemcpyAsync(h->d_fresh.p + (size_t)n_new_obs * M, new_pts3, sizeof(double) * 3 * n_new_pts, cudaMemcpyHostToDevice, st));
      if (h->cfg.use_bounds) for (int k = 0; k < n_new_pts; k++) if (cnt_n[n_alive + k] > 0 && !inside(new_pts3 + (size_t)k * 3)) h->win_infeasible[0] = 1;
    }
    h->timing.kernel_launches += launch_win_points(h->d_pts.p + (size_t)par * NPo * 3, T + t_src, h->d_fresh.p + (size_t)n_new_obs * M, NPn, h->d_pt_rec.p, st);
  }
  // No wait for the stream here: the buffers re-sized below are not the ones the kernels above touch (a buffer that does
  // have to grow is released with cudaFree, which waits for the device), and the uploads above come from pinned staging
  // that is not written again before the next advance.
  TT("advance: shift + append")
  if (cams6_all) std::memcpy(h->h_cams.p, cams6_all, sizeof(double) * 6 * NCn);
  else {
    CU(h, cudaEventSynchronize(h->ev[6]));     // the previous cameras: on their way since the start of the call
    std::memcpy(h->h_cams.p, h->h_cams_prev.p + (size_t)n_drop * 6, sizeof(double) * 6 * (NCo - n_drop));
    if (n_new_cams) std::memcpy(h->h_cams.p + (size_t)(NCo - n_drop) * 6, new_cams6, sizeof(double) * 6 * n_new_cams);
  }
  // ---- 5. commit: the handle now describes the new window ------------------------------------------------------------
  h->NC = NCn; h->NP = NPn; h->NO = NOn; h->nW = 1;
  h->w_cam_off = {0, NCn}; h->w_pt_off = {0, NPn}; h->w_obs_off = {0, NOn};
  h->tr_lo.assign(lo_n, lo_n + NPn); h->tr_cnt.swap(cnt_n); h->tr_cid.assign(cid_n, cid_n + NPn); h->rows_cur = nxt;
  h->pt_order.assign(order, order + NPn);
  h->pt_obs_off_int.assign(off_i, off_i + NPn + 1);
  h->pt_obs_off_caller.resize((size_t)NPn + 1);
  for (int j = 0; j <= NPn; j++) h->pt_obs_off_caller[j] = off_n[j];
  h->cam_seen.assign(NCn, 0);
  {
    std::vector<int32_t> diff((size_t)NCn + 1, 0);
    for (int j = 0; j < NPn; j++) if (h->tr_cnt[j]) { diff[h->tr_lo[j]]++; diff[h->tr_lo[j] + h->tr_cnt[j]]--; }
    int run = 0; h->local_cam_lo = NCn; h->local_cam_hi = -1;
    for (int c = 0; c < NCn; c++) { run += diff[c]; if (run > 0) { h->cam_seen[c] = 1; h->local_cam_lo = std::min(h->local_cam_lo, c); h->local_cam_hi = c; } }
  }
  h->cam_win_h.assign(NCn, 0); h->pt_win_h.assign(NPn, 0);
  h->host_obs_valid = false; h->prepared_fixed = -1; h->cov_h.clear();
  CU(h, h->d_cams.reserve((size_t)NCn * 12)); CU(h, h->d_camR.reserve((size_t)NCn * kCamStride * 2));
  CU(h, h->d_cam_s2.reserve((size_t)NCn * 6)); CU(h, h->d_cam_lam.reserve((size_t)NCn * 6)); CU(h, h->d_cam_y.reserve((size_t)NCn * 6));
  CU(h, h->d_pt_s2.reserve((size_t)NPn * 3));
  CU(h, h->d_feat.reserve((size_t)NOn * M)); CU(h, h->d_obs_cam.reserve((size_t)NOn)); CU(h, h->d_Zbuf.reserve((size_t)NOn * 18));
  CU(h, h->d_cam_win.reserve(NCn)); CU(h, h->d_pt_obs_off.reserve((size_t)NPn + 1)); CU(h, h->d_pt_win.reserve(NPn)); CU(h, h->d_pt_order.reserve(NPn));
  CU(h, h->d_recs.reserve((size_t)rec_stride));
  if (!pts3_all) {
    // the re-slotted points sit in the scratch: move them to parity 0 of the (possibly re-sized) point buffer, keep a host copy
    DevBuf<double> fresh_pts;
    if ((size_t)NPn * 6 > h->d_pts.cap) { CU(h, fresh_pts.reserve((size_t)NPn * 6 + (size_t)NPn * 6 / 4)); std::swap(fresh_pts.p, h->d_pts.p); std::swap(fresh_pts.cap, h->d_pts.cap); fresh_pts.release(); }
    CU(h, cudaMemcpyAsync(h->d_pts.p, h->d_pt_rec.p, sizeof(double) * 3 * NPn, cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(h->h_pts.p, h->d_pts.p, sizeof(double) * 3 * NPn, cudaMemcpyDeviceToHost, st));
    h->host_iter_pending = true;
  } else {
    CU(h, h->d_pts.reserve((size_t)NPn * 6));
    CU(h, cudaMemcpyAsync(h->d_pts.p, h->h_pts.p, sizeof(double) * 3 * NPn, cudaMemcpyHostToDevice, st));
  }
  CU(h, cudaMemcpyAsync(h->d_cams.p, h->h_cams.p, sizeof(double) * 6 * NCn, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_w_cam_off.p, h->w_cam_off.data(), sizeof(int32_t) * 2, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_w_pt_off.p, h->w_pt_off.data(), sizeof(int32_t) * 2, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemsetAsync(h->d_cam_win.p, 0, sizeof(int32_t) * NCn, st)); CU(h, cudaMemsetAsync(h->d_pt_win.p, 0, sizeof(int32_t) * NPn, st));
  CU(h, cudaMemcpyAsync(h->d_pt_obs_off.p, T + t_offi, sizeof(int32_t) * ((size_t)NPn + 1), cudaMemcpyDeviceToDevice, st));
  CU(h, cudaMemcpyAsync(h->d_pt_order.p, T + t_order, sizeof(int32_t) * NPn, cudaMemcpyDeviceToDevice, st));
  h->timing.kernel_launches += launch_win_pack(h->d_rows[nxt].p, h->d_tr_off[nxt].p, T + t_lo, h->d_tr_cid.p, T + t_order, T + t_offi, NPn, NOn, M,
                                               h->d_feat.p, h->d_obs_cam.p, st);
  CU(h, cudaMemsetAsync(h->d_recs.p, 0, sizeof(IterRec) * (size_t)rec_stride, st));
  fill_view_static(h);
  h->V.rec_stride = rec_stride;
  h->ws_h.assign(1, WinState{});
  h->state = 1;
  if (pt_id_map) std::memcpy(pt_id_map, id_map, sizeof(int32_t) * NPo);
  TT("advance: pack (enqueue)")
  return UBA_OK;
}

int uba_get_sizes(const uba_handle* h, int* n_windows, int* n_cams, int* n_pts, int64_t* n_obs) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (n_windows) *n_windows = h->nW;
  if (n_cams) *n_cams = h->NC;
  if (n_pts) *n_pts = h->NP;
  if (n_obs) *n_obs = h->NO;
  return UBA_OK;
}

int uba_get_tables(uba_handle* h, int fixed_frames, int32_t* obs_order, int64_t* pt_obs_off, int32_t* pt_order, int32_t* free_cam) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  { const int rc = ensure_host_obs_tables(h); if (rc) return rc; }
  if (obs_order) std::memcpy(obs_order, h->obs_order.data(), sizeof(int32_t) * h->NO);
  if (pt_obs_off) std::memcpy(pt_obs_off, h->pt_obs_off_caller.data(), sizeof(int64_t) * ((size_t)h->NP + 1));
  if (pt_order) std::memcpy(pt_order, h->pt_order.data(), sizeof(int32_t) * h->NP);
  if (free_cam) {
    cudaSetDevice(h->device);
    int rc = prepare(h, fixed_frames);
    if (rc) return rc;
    std::memcpy(free_cam, h->free_cam_h.data(), sizeof(int32_t) * h->NC);
  }
  return UBA_OK;
}

int uba_linearize(uba_handle* h, int fixed_frames, double radius, uba_linearization_out* out) {
  if (!h || !out) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  cudaSetDevice(h->device);
  { const int rc0 = ensure_host_obs_tables(h); if (rc0) return rc0; }
  const double saved_radius = h->cfg.initial_radius;
  h->cfg.initial_radius = radius;
  // the device copy is about to be used as scratch: results of an earlier uba_optimise are gone, the getters fall back
  // to the staged initial iterate and a later uba_optimise starts from it again
  h->state = 1; h->device_dirty = true; h->cov_h.clear();
  int rc = upload_state(h);
  if (!rc) rc = start_solve(h, fixed_frames);
  h->cfg.initial_radius = saved_radius;
  if (rc) return rc;
  const int M = h->M, NC = h->NC, NP = h->NP;
  const int64_t NO = h->NO;
  // debug scratch: residuals | weights | C | W | grad_pts | lam_pts
  const size_t n_res = (size_t)NO * M, n_w = NO, n_C = (size_t)NP * 9, n_W = (size_t)NO * 18, n_g = (size_t)NP * 3, n_l = (size_t)NP * 3;
  CU(h, h->d_dbg.reserve(n_res + n_w + n_C + n_W + n_g + n_l));
  CU(h, cudaMemsetAsync(h->d_dbg.p, 0, (n_res + n_w + n_C + n_W + n_g + n_l) * sizeof(double), h->stream));
  DebugOut D;
  D.residuals = h->d_dbg.p; D.weights = D.residuals + n_res; D.C = D.weights + n_w; D.W = D.C + n_C; D.grad_pts = D.W + n_W; D.lam_pts = D.grad_pts + n_g;
  // the parity dump wants the dense damped matrix even for windows the banded solver would handle
  std::vector<int32_t> zeros(h->nW, 0);
  CU(h, cudaMemcpyAsync(h->d_w_beta.p, zeros.data(), sizeof(int32_t) * h->nW, cudaMemcpyHostToDevice, h->stream));
  h->dense_override = true;
  rc = run_linearize(h, D);
  h->dense_override = false;
  if (rc) { cudaMemcpyAsync(h->d_w_beta.p, h->win_beta.data(), sizeof(int32_t) * h->nW, cudaMemcpyHostToDevice, h->stream); return rc; }
  h->timing.kernel_launches += launch_assemble(h->Vc, h->max_n, h->stream);
  CU(h, cudaMemcpyAsync(h->d_w_beta.p, h->win_beta.data(), sizeof(int32_t) * h->nW, cudaMemcpyHostToDevice, h->stream));
  rc = comm_wait(h);
  if (rc) return rc;
  CU(h, cudaGetLastError());
  std::vector<double> tmp;
  auto fetch = [&](const double* dev, size_t n) -> const double* { tmp.resize(n); cudaMemcpy(tmp.data(), dev, n * sizeof(double), cudaMemcpyDeviceToHost); return tmp.data(); };
  if (out->residuals) { const double* r = fetch(D.residuals, n_res); for (int64_t s = 0; s < NO; s++) std::memcpy(out->residuals + (size_t)h->h_obs_internal.p[s] * M, r + (size_t)s * M, sizeof(double) * M); }
  if (out->weights) { const double* r = fetch(D.weights, n_w); for (int64_t s = 0; s < NO; s++) out->weights[h->h_obs_internal.p[s]] = r[s]; }
  if (out->W) { const double* r = fetch(D.W, n_W); for (int64_t s = 0; s < NO; s++) std::memcpy(out->W + (size_t)h->h_obs_internal.p[s] * 18, r + (size_t)s * 18, sizeof(double) * 18); }
  if (out->C) { const double* r = fetch(D.C, n_C); for (int s = 0; s < NP; s++) std::memcpy(out->C + (size_t)h->pt_order[s] * 9, r + (size_t)s * 9, sizeof(double) * 9); }
  if (out->grad_pts) { const double* r = fetch(D.grad_pts, n_g); for (int s = 0; s < NP; s++) std::memcpy(out->grad_pts + (size_t)h->pt_order[s] * 3, r + (size_t)s * 3, sizeof(double) * 3); }
  if (out->lm_diag_pts) { const double* r = fetch(D.lam_pts, n_l); for (int s = 0; s < NP; s++) std::memcpy(out->lm_diag_pts + (size_t)h->pt_order[s] * 3, r + (size_t)s * 3, sizeof(double) * 3); }
  if (out->cost) { const double* r = fetch(h->Vc.w_lin, (size_t)h->nW * WL_COUNT); for (int w = 0; w < h->nW; w++) out->cost[w] = r[(size_t)w * WL_COUNT + WL_COST]; }
  if (out->grad_cams) { const double* r = fetch(h->Vc.vacc, (size_t)NC * 6); std::memcpy(out->grad_cams, r, sizeof(double) * 6 * NC); }
  if (out->lm_diag_cams) {
    const double* r = fetch(h->V.cam_lam, (size_t)NC * 6);
    for (int c = 0; c < NC; c++) for (int a = 0; a < 6; a++) out->lm_diag_cams[(size_t)c * 6 + a] = h->free_cam_h[c] >= 0 ? r[(size_t)c * 6 + a] : 0.0;
  }
  if (out->B) {
    const double* r = fetch(h->Vc.Bacc, (size_t)NC * 36);
    for (int c = 0; c < NC; c++) for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
      out->B[(size_t)c * 36 + a * 6 + b] = r[(size_t)c * 36 + std::min(a, b) * 6 + std::max(a, b)];
  }
  if (out->S) { const double* r = fetch(h->V.A, (size_t)h->w_red_off_h[h->nW]); std::memcpy(out->S, r, sizeof(double) * h->w_red_off_h[h->nW]); }
  if (out->rhs) { const double* r = fetch(h->V.rhs, h->free_list_h.size() * 6); std::memcpy(out->rhs, r, sizeof(double) * h->free_list_h.size() * 6); }
  return UBA_OK;
}

int uba_optimise(uba_handle* h, int fixed_frames, uba_summary* summaries) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  // "[Bundle Adjuster] system should be initiliased to perform optimisation!" (BundleAdjuster.h:381-384,:434-437)
  if (h->state != 1) return fail(h, UBA_ERR_STATE, "system should be initialised to perform optimisation (state %d)", h->state);
  h->in_optimise = true;
  cudaSetDevice(h->device);
  const auto t_start = std::chrono::steady_clock::now();
  cudaEventRecord(h->ev[2], h->stream);
  auto tt_ = std::chrono::steady_clock::now();
  int rc = UBA_OK;
  if (h->device_dirty) { rc = upload_state(h); if (rc) return rc; h->device_dirty = false; }
  rc = start_solve(h, fixed_frames);
  if (rc) return rc;
  TT("optimise: start_solve")
  const bool fixedK = h->cfg.fixed_iterations > 0;
  const int max_it = fixedK ? h->cfg.fixed_iterations : h->cfg.max_iterations;
  const bool timed = !fixedK && h->cfg.max_solver_time_s > 0.0;
  // infeasible windows fail before the first iteration; on a point-sharded handle the window fails on EVERY rank when any
  // rank's shard starts outside the box (agreed in prepare), so that all ranks run the same sequence of exchanges
  std::vector<char> infeasible(h->nW, 0);
  bool any_infeasible = false;
  for (int w = 0; w < h->nW; w++) if (!window_feasible(h, w) || (h->comm && h->any_rank_infeasible)) { infeasible[w] = 1; any_infeasible = true; }
  if (any_infeasible) {
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(h->ws_h.data(), h->d_ws.p, sizeof(WinState) * h->nW, cudaMemcpyDeviceToHost));
    int n_act = 0;
    for (int w = 0; w < h->nW; w++) { if (infeasible[w]) h->ws_h[w].done = UBA_TERM_FAILURE; else n_act++; }
    CU(h, cudaMemcpy(h->d_ws.p, h->ws_h.data(), sizeof(WinState) * h->nW, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->d_n_active.p, &n_act, sizeof(int), cudaMemcpyHostToDevice));
  }
  CU(h, cudaMemsetAsync(h->d_stop_req, 0, sizeof(double), h->stream));
  bool stop_sent = false;
  for (int it = 0; it < max_it; it++) {
    rc = run_iteration_fast(h);
    if (rc) return rc;
    if (it == 0) TT("optimise: first launch")
    if (!fixedK) {
      // convergence is decided on the device; the host only needs to know when every window is done
      int n_act = 0;
      CU(h, cudaMemcpyAsync(&n_act, h->d_n_active.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      rc = comm_wait(h);
      if (rc) return rc;
      if (n_act <= 0) break;
      if (timed && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() >= h->cfg.max_solver_time_s) {
        // wall-clock cap (:417,:464).  Alone, this rank just stops.  Point-sharded, every rank has its own clock: the request
        // travels with the next iteration's exchange and the controller stops all ranks after that iteration.
        if (!h->comm) break;
        if (!stop_sent) { const double one = 1.0; CU(h, cudaMemcpyAsync(h->d_stop_req, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream)); stop_sent = true; }
      }
    }
  }
  cudaEventRecord(h->ev[3], h->stream);
  rc = comm_wait(h);
  if (rc) return rc;
  CU(h, cudaGetLastError());
  if (h->peer_on) {
    int32_t perr = 0;
    CU(h, cudaMemcpy(&perr, h->P.err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (perr) {
      CU(h, cudaMemset(h->P.err, 0, sizeof(int32_t)));
      return fail(h, UBA_ERR_NCCL, "rank %d: rank %d did not reach an exchange of the iteration within the time-out (UBA_PEER_TIMEOUT_S)", h->rank, perr - 1);
    }
  }
  TT("optimise: iterations")
  float ms = 0; cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
  h->timing.total_ms += ms;
  h->cov_h.clear();
  if (h->cfg.compute_covariance) {
    rc = compute_covariances(h);
    if (rc) return rc;
  }
  CU(h, cudaMemcpy(h->ws_h.data(), h->d_ws.p, sizeof(WinState) * h->nW, cudaMemcpyDeviceToHost));
  int worst = UBA_OK;
  for (int w = 0; w < h->nW; w++) {
    WinState& s = h->ws_h[w];
    if (s.done == 0) s.done = UBA_TERM_NO_CONVERGENCE;  // iteration or wall-clock cap (:417,:464): still usable
    if (summaries) {
      uba_summary& o = summaries[w];
      o.termination = s.done; o.usable = s.done != UBA_TERM_FAILURE; o.iterations = s.iter; o.successful_steps = s.n_success;
      o.unsuccessful_steps = s.n_unsuccess; o.invalid_steps = s.n_invalid; o.initial_cost = s.initial_cost; o.final_cost = s.cost;
      o.final_radius = s.radius; o.final_gradient_max_norm = s.gmax;
    }
    if (s.done == UBA_TERM_FAILURE) worst = infeasible[w] ? UBA_ERR_INFEASIBLE : UBA_ERR_NUMERICAL;
  }
  h->state = 2;
  if (worst != UBA_OK) fail(h, worst, "at least one window did not produce a usable solution");
  return worst;
}

int uba_get_cameras(uba_handle* h, double* cams6) {
  if (!h || !cams6) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  cudaSetDevice(h->device);
  if (h->host_iter_pending) { CU(h, cudaStreamSynchronize(h->stream)); h->host_iter_pending = false; }
  if (h->state == 1) { std::memcpy(cams6, h->h_cams.p, sizeof(double) * 6 * h->NC); return UBA_OK; }
  std::vector<double> both((size_t)h->NC * 12);
  CU(h, cudaMemcpy(both.data(), h->d_cams.p, sizeof(double) * 12 * h->NC, cudaMemcpyDeviceToHost));
  for (int c = 0; c < h->NC; c++) {
    const WinState& s = h->ws_h[h->cam_win_h[c]];
    const double* src = s.done == UBA_TERM_FAILURE ? h->h_cams.p + (size_t)c * 6 : both.data() + (size_t)s.cur * h->NC * 6 + (size_t)c * 6;
    std::memcpy(cams6 + (size_t)c * 6, src, sizeof(double) * 6);
  }
  return UBA_OK;
}

int uba_get_points(uba_handle* h, double* pts3) {
  if (!h || !pts3) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  cudaSetDevice(h->device);
  const int NP = h->NP;
  if (h->host_iter_pending) { CU(h, cudaStreamSynchronize(h->stream)); h->host_iter_pending = false; }
  if (h->state == 1) {
    for (int s = 0; s < NP; s++) std::memcpy(pts3 + (size_t)h->pt_order[s] * 3, h->h_pts.p + (size_t)s * 3, sizeof(double) * 3);
    return UBA_OK;
  }
  CU(h, h->h_out.reserve((size_t)NP * 6));
  // the accepted iterate lives in one of the two point buffers per window: fetch only the halves in use
  bool need[2] = {false, false};
  for (int w = 0; w < h->nW; w++) if (h->ws_h[w].done != UBA_TERM_FAILURE) need[h->ws_h[w].cur & 1] = true;
  for (int b = 0; b < 2; b++)
    if (need[b]) CU(h, cudaMemcpyAsync(h->h_out.p + (size_t)b * NP * 3, h->d_pts.p + (size_t)b * NP * 3, sizeof(double) * 3 * NP, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
#pragma omp parallel for schedule(static)
  for (int s = 0; s < NP; s++) {
    const WinState& st = h->ws_h[h->pt_win_h[s]];
    const double* src = st.done == UBA_TERM_FAILURE ? h->h_pts.p + (size_t)s * 3 : h->h_out.p + (size_t)st.cur * NP * 3 + (size_t)s * 3;
    std::memcpy(pts3 + (size_t)h->pt_order[s] * 3, src, sizeof(double) * 3);
  }
  return UBA_OK;
}

int uba_get_pose_covariances(uba_handle* h, double* cov36) {
  if (!h || !cov36) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state != 2) return fail(h, UBA_ERR_STATE, "optimise has not run");
  if (h->cov_h.empty())
    return fail(h, UBA_ERR_STATE, "no covariances: set uba_config.compute_covariance before uba_optimise (not available on point-sharded handles)");
  // cameras of windows without a usable solution, fixed cameras and unobserved cameras report zeros
  for (int c = 0; c < h->NC; c++) {
    const bool ok = h->ws_h[h->cam_win_h[c]].done != UBA_TERM_FAILURE && h->free_cam_h[c] >= 0;
    for (int i = 0; i < 36; i++) cov36[(size_t)c * 36 + i] = ok ? h->cov_h[(size_t)c * 36 + i] : 0.0;
  }
  return UBA_OK;
}

int uba_get_iterations(uba_handle* h, int window, uba_iteration* out, int max_records, int* n_records) {
  if (!h || window < 0 || window >= h->nW) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state != 2) return fail(h, UBA_ERR_STATE, "optimise has not run");
  cudaSetDevice(h->device);
  const int n = std::min(h->ws_h[window].iter + 1, h->V.rec_stride);
  if (n_records) *n_records = n;
  if (out && max_records > 0) {
    std::vector<IterRec> r(n);
    CU(h, cudaMemcpy(r.data(), h->d_recs.p + (size_t)window * h->V.rec_stride, sizeof(IterRec) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < max_records; i++) {
      out[i].cost = r[i].cost; out[i].candidate_cost = r[i].candidate_cost; out[i].model_cost_change = r[i].model_cost_change;
      out[i].relative_decrease = r[i].relative_decrease; out[i].radius = r[i].radius; out[i].step_norm = r[i].step_norm;
      out[i].gradient_max_norm = r[i].gradient_max_norm; out[i].accepted = r[i].accepted; out[i].pad_ = 0;
    }
  }
  return UBA_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------
int uba_comm_unique_id(uba_handle* h, char id[UBA_NCCL_UNIQUE_ID_BYTES]) {
  if (!h || !id) return UBA_ERR_INVALID_ARGUMENT;
  std::string err;
  if (!load_nccl(h->nccl, err)) return fail(h, UBA_ERR_NCCL, "%s", err.c_str());
  Id128 u; std::memset(&u, 0, sizeof(u));
  const int rc = h->nccl.GetUniqueId(&u);
  if (rc != 0) return fail(h, UBA_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  std::memcpy(id, u.internal, UBA_NCCL_UNIQUE_ID_BYTES);
  return UBA_OK;
}

int uba_comm_init(uba_handle* h, const char id[UBA_NCCL_UNIQUE_ID_BYTES], int rank, int n_ranks) {
  if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return UBA_ERR_INVALID_ARGUMENT;
  std::string err;
  if (!load_nccl(h->nccl, err)) return fail(h, UBA_ERR_NCCL, "%s", err.c_str());
  // one communicator per handle, set before the problem: the camera tables of a resident problem were built without it
  if (h->comm) return fail(h, UBA_ERR_STATE, "uba_comm_init: this handle already has a communicator");
  if (h->state != 0) return fail(h, UBA_ERR_STATE, "uba_comm_init must come before uba_set_problem");
  cudaSetDevice(h->device);
  if (const char* e = std::getenv("UBA_PEER")) h->peer_wanted = !(e[0] == '0');
  Id128 u; std::memcpy(u.internal, id, UBA_NCCL_UNIQUE_ID_BYTES);
  const int rc = h->nccl.CommInitRank(&h->comm, n_ranks, u, rank);
  if (rc != 0) { h->comm = nullptr; return fail(h, UBA_ERR_NCCL, "ncclCommInitRank failed: %s", h->nccl.GetErrorString ? h->nccl.GetErrorString(rc) : "?"); }
  h->rank = rank; h->n_ranks = n_ranks;
  h->prepared_fixed = -1;                     // the accumulator layout depends on the communicator
  return UBA_OK;
}

// ---- timing ------------------------------------------------------------------------------------
int uba_set_profiling(uba_handle* h, int enabled) { if (!h) return UBA_ERR_INVALID_ARGUMENT; h->profiling = enabled != 0; return UBA_OK; }

int uba_get_timing(uba_handle* h, uba_timing* out, int reset) {
  if (!h) return UBA_ERR_INVALID_ARGUMENT;
  if (out) *out = h->timing;
  if (reset) std::memset(&h->timing, 0, sizeof(h->timing));
  return UBA_OK;
}

static int flush_l2(uba_handle* h) {
  const size_t n = (size_t)48 << 20;  // 384 MB of doubles > 126 MB L2
  CU(h, h->d_flush.reserve(n));
  h->timing.kernel_launches += launch_l2_flush(h->d_flush.p, n, h->stream);
  return UBA_OK;
}

int uba_time_linearize(uba_handle* h, int fixed_frames, double radius, int repeats, int do_flush, double* ms_per_pass) {
  if (!h || repeats <= 0 || !ms_per_pass) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  h->in_optimise = false;
  cudaSetDevice(h->device);
  const double saved = h->cfg.initial_radius;
  h->cfg.initial_radius = radius;
  h->state = 1; h->device_dirty = true; h->cov_h.clear();
  int rc = upload_state(h);
  if (!rc) rc = start_solve(h, fixed_frames);
  h->cfg.initial_radius = saved;
  if (rc) return rc;
  // one full LM iteration first: what is timed is the pass of a RUNNING solve (Jacobi scales captured, the iterate one
  // accepted step in), not the very first linearisation with its extra square roots and stores
  rc = run_iteration(h);
  if (rc) return rc;
  DebugOut none{};
  double total = 0.0;
  for (int r = 0; r < repeats; r++) {
    if (do_flush) { rc = flush_l2(h); if (rc) return rc; }
    CU(h, cudaMemsetAsync(h->d_acc.p, 0, h->acc_total * sizeof(double), h->stream));
    const bool pipe = h->pipe_on && !h->dense_override;    // the pass as the pipelined iteration runs it: with its assembly hooks
    if (pipe) CU(h, cudaMemsetAsync(h->d_pipe.p, 0, sizeof(int32_t) * ((size_t)h->NC + h->free_list_h.size()), h->stream));
    cudaEventRecord(h->ev[4], h->stream);
    if (pipe) { DevView Vp = h->V; Vp.pipe_on = 1; h->timing.kernel_launches += launch_lin_tiled(Vp, h->variant_off, h->stream); }
    else h->timing.kernel_launches += launch_linearizers(h, none);
    h->timing.linearize_launches++;
    cudaEventRecord(h->ev[5], h->stream);
    CU(h, cudaEventSynchronize(h->ev[5]));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
    total += ms;
  }
  CU(h, cudaGetLastError());
  *ms_per_pass = total / repeats;
  return UBA_OK;
}

int uba_time_iteration(uba_handle* h, int fixed_frames, int iterations, int do_flush, double* ms_per_iteration) {
  if (!h || iterations <= 0 || !ms_per_iteration) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  h->in_optimise = false;
  cudaSetDevice(h->device);
  const int saved_fixed = h->cfg.fixed_iterations;
  h->cfg.fixed_iterations = iterations;  // no convergence tests: every iteration runs all phases
  fill_view_static(h);
  h->state = 1; h->device_dirty = true; h->cov_h.clear();
  int rc = upload_state(h);
  if (!rc) rc = start_solve(h, fixed_frames);
  // Each iteration sits between its own pair of events (the L2 flush before it is outside the pair); the host does not
  // wait between iterations, only after every batch of 32, so that launch latency is hidden behind execution exactly as
  // in uba_optimise — it matters on point-sharded handles, which launch kernels and collectives one by one.
  double total = 0.0;
  constexpr int kBatch = 32;
  static thread_local std::vector<cudaEvent_t> evs;
  if (evs.empty()) { evs.resize(2 * kBatch); for (auto& e : evs) cudaEventCreate(&e); }
  CU(h, cudaStreamSynchronize(h->stream));
  for (int it0 = 0; it0 < iterations && !rc; it0 += kBatch) {
    const int nb = std::min(kBatch, iterations - it0);
    for (int i = 0; i < nb && !rc; i++) {
      if (do_flush) rc = flush_l2(h);
      if (rc) break;
#ifndef UBA_EMU
      // point-sharded: line the ranks up after the flush, otherwise the iteration's first exchange waits for whichever
      // peer is still flushing and that wait lands inside this rank's event pair
      if (h->peer_on) h->timing.kernel_launches += launch_peer_barrier(h->P, h->stream);
#endif
      cudaEventRecord(evs[2 * i], h->stream);
      rc = run_iteration_fast(h);
      cudaEventRecord(evs[2 * i + 1], h->stream);
    }
    if (rc) break;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = fail(h, UBA_ERR_CUDA, "stream sync failed"); break; }
    for (int i = 0; i < nb; i++) { float ms = 0; cudaEventElapsedTime(&ms, evs[2 * i], evs[2 * i + 1]); total += ms; }
  }
  h->cfg.fixed_iterations = saved_fixed;
  fill_view_static(h);
  if (rc) return rc;
  CU(h, cudaGetLastError());
  *ms_per_iteration = total / iterations;
  return UBA_OK;
}

#ifdef UBA_EMU
// host-emulation build only (tests/test_host_logic_emu.py): the lineariser's tile plan as the GPU build would make it.
// parts: rows of {window, pt_begin, pt_end, n_local, n_fixed, slot kernel?, first camera}; pt_mask: [n_pts] in internal order
int uba_emu_tile_plan(uba_handle* h, int fixed_frames, int32_t* parts, int max_parts, uint32_t* pt_mask, int32_t* n_parts, int32_t* n_generic) {
  if (!h || !parts || !pt_mask || !n_parts || !n_generic) return UBA_ERR_INVALID_ARGUMENT;
  if (h->state < 1) return fail(h, UBA_ERR_STATE, "no problem set");
  int rc = prepare(h, fixed_frames);
  if (rc) return rc;
  build_tile_plan(h, fixed_frames);
  *n_parts = (int32_t)h->parts_h.size(); *n_generic = (int32_t)h->gen_pts_h.size();
  for (int i = 0; i < std::min<int>(max_parts, *n_parts); i++) {
    const TilePart& p = h->parts_h[i];
    const int32_t row[7] = {p.window, p.pt_begin, p.pt_end, p.n_local, p.n_fixed, p.pad_[0], h->tile_cams_h[p.cam_list_off]};
    std::memcpy(parts + (size_t)i * 7, row, sizeof(row));
  }
  std::memcpy(pt_mask, h->pt_mask_h.data(), sizeof(uint32_t) * h->NP);
  return UBA_OK;
}
#endif

#ifdef UBA_BAND_TIMING
// instrumentation builds only (scripts/band_timing.py): the clocks the band solver leaves in the generic lineariser's scratch
int uba_debug_read_zbuf(uba_handle* h, double* out, int count) {
  if (!h || !out || count < 0 || (size_t)count > h->d_Zbuf.cap) return UBA_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  CU(h, cudaMemcpy(out, h->d_Zbuf.p, sizeof(double) * count, cudaMemcpyDeviceToHost));
  return UBA_OK;
}
#endif

// fp64 FMA peak probe (TFLOP/s), used for the second roofline ceiling of the lineariser
int uba_probe_fp64_tflops(uba_handle* h, double* tflops) {
  if (!h || !tflops) return UBA_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  CU(h, h->d_flush.reserve(1024));
  const int iters = 200000;
  launch_dfma_probe(h->d_flush.p, 1000, h->stream);
  cudaEventRecord(h->ev[4], h->stream);
  launch_dfma_probe(h->d_flush.p, iters, h->stream);
  cudaEventRecord(h->ev[5], h->stream);
  CU(h, cudaEventSynchronize(h->ev[5]));
  float ms = 0; cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
  const double flops = 2.0 * 8.0 * (double)iters * 256.0 * 148.0 * 8.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  h->timing.kernel_launches += 2;
  return UBA_OK;
}

}  // extern "C"

#ifndef UBA_EMU
uba_vo_state* uba_vo_get(uba_handle* h, bool create) {
  cudaSetDevice(h->device);
  if (!h->vo && create) h->vo = uba_vo_new();
  return h->vo;
}
cudaStream_t uba_vo_stream(uba_handle* h) { return h->stream; }
int uba_vo_fail(uba_handle* h, int code, const char* what, const char* detail) { return fail(h, code, "%s: %s", what, detail); }
void uba_vo_count(uba_handle* h, int kernels) { h->timing.kernel_launches += kernels; }
#endif
