// uba_synth.cpp — deterministic synthetic stereo-rig windows.
//
// The reference ships no data set and no calibration; this generator produces the
// KITTI-shaped windows BASELINE.json names.  Counter-based RNG (splitmix64 keyed by
// (seed, stream, index)), so every point / camera / observation is reproducible on its
// own, independent of generation order or thread count.
//
// Rig: K0 = K1 = [718.856 0 607.1928; 0 718.856 185.2157; 0 0 1], baseline 0.537 m,
// image 1241 x 376, sigma = 0.5 px (feat_var 0.25, the reference default,
// include/MotionEstimation/core/file_IO.h:73).  Pose model p_cam = R(r) X + t as in
// BundleAdjuster.h:157-160.  Keyframe i: centre (0.3 sin 0.05 i, 0, 0.8 i), yaw
// 0.02 sin(0.1 i) about +y.  A point is anchored in its "home" keyframe h (the LAST
// keyframe of its track) at a uniform pixel and depth U[4,60] m, and tracked over the
// L keyframes [h-L+1, h]; an observation is dropped if either projection leaves the
// image or z < 1 m.  Features are rounded to float32 and widened, which is what
// cv::Point2f -> double does at BundleAdjuster.h:371.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/uba.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct Rng {
  uint64_t key;
  uint64_t ctr = 0;
  Rng(uint64_t seed, uint64_t stream, uint64_t index) { key = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + index); }
  uint64_t next() { return splitmix64(key + (ctr++) * 0x9E3779B97F4A7C15ull); }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uniform(double a, double b) { return a + (b - a) * uniform(); }
  int uniform_int(int a, int b) { return a + (int)(next() % (uint64_t)(b - a + 1)); }  // inclusive
  double normal() {
    double u1 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    const double u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

constexpr double kImgW = 1241.0, kImgH = 376.0;

void rot_y(double yaw, double R[9]) {
  // exp([0,yaw,0]x)
  const double c = std::cos(yaw), s = std::sin(yaw);
  R[0] = c; R[1] = 0; R[2] = s; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = -s; R[7] = 0; R[8] = c;
}

void keyframe_pose(int i, double cam6[6], double R[9]) {
  const double centre[3] = {0.3 * std::sin(0.05 * i), 0.0, 0.8 * i};
  const double yaw = 0.02 * std::sin(0.1 * i);
  rot_y(yaw, R);
  cam6[3] = 0.0; cam6[4] = yaw; cam6[5] = 0.0;
  for (int a = 0; a < 3; a++) cam6[a] = -(R[a * 3] * centre[0] + R[a * 3 + 1] * centre[1] + R[a * 3 + 2] * centre[2]);
}

}  // namespace

extern "C" {

void uba_synth_default_calib(uba_calib* k) {
  k->fx0 = k->fy0 = k->fx1 = k->fy1 = 718.856;
  k->cx0 = k->cx1 = 607.1928;
  k->cy0 = k->cy1 = 185.2157;
  k->feat_var = 0.25;
  k->baseline = 0.537;
}

int64_t uba_synth_generate(const uba_synth_spec* spec, const uba_calib* calib, int64_t max_obs, double* cams_gt,
                           double* cams_init, double* pts_gt, double* pts_init, double* feats, int32_t* cam_idx,
                           int32_t* pt_idx, int32_t* cam_id) {
  if (!spec || !calib || !cams_init || !pts_init || !feats || !cam_idx || !pt_idx) return UBA_ERR_INVALID_ARGUMENT;
  const int M = spec->M;
  if ((M != 2 && M != 4) || spec->n_cams <= 0 || spec->n_pts < 0) return UBA_ERR_INVALID_ARGUMENT;
  const int nc = spec->n_cams;
  const uba_calib k = *calib;
  std::vector<double> R((size_t)nc * 9), cg((size_t)nc * 6);
  for (int i = 0; i < nc; i++) {
    keyframe_pose(i, &cg[(size_t)i * 6], &R[(size_t)i * 9]);
    Rng rng(spec->seed, 1, (uint64_t)i);
    for (int a = 0; a < 6; a++) {
      const double noise = (a < 3 ? spec->pose_t_sigma : spec->pose_r_sigma) * rng.normal();
      if (cams_gt) cams_gt[(size_t)i * 6 + a] = cg[(size_t)i * 6 + a];
      cams_init[(size_t)i * 6 + a] = cg[(size_t)i * 6 + a] + (i >= spec->fixed_frames ? noise : 0.0);
    }
  }
  int64_t no = 0;
  const int tmin = spec->track_min < 1 ? 1 : spec->track_min;
  const int tmax = spec->track_max < tmin ? tmin : spec->track_max;
  for (int j = 0; j < spec->n_pts; j++) {
    Rng rng(spec->seed, 2, (uint64_t)j);
    int L = spec->full_tracks ? nc : rng.uniform_int(tmin, tmax);
    if (L > nc) L = nc;
    const int h = spec->full_tracks ? nc - 1 : rng.uniform_int(L - 1, nc - 1);
    const double u = rng.uniform(20.0, kImgW - 20.0), v = rng.uniform(20.0, kImgH - 20.0), z = rng.uniform(4.0, 60.0);
    const double pc[3] = {(u - k.cx0) * z / k.fx0, (v - k.cy0) * z / k.fy0, z};
    const double* Rh = &R[(size_t)h * 9];
    const double* th = &cg[(size_t)h * 6];
    const double d[3] = {pc[0] - th[0], pc[1] - th[1], pc[2] - th[2]};
    double X[3];
    for (int a = 0; a < 3; a++) X[a] = Rh[a] * d[0] + Rh[3 + a] * d[1] + Rh[6 + a] * d[2];  // R^T d
    for (int a = 0; a < 3; a++) {
      if (pts_gt) pts_gt[(size_t)j * 3 + a] = X[a];
      pts_init[(size_t)j * 3 + a] = X[a] + spec->point_rel_sigma * z * rng.normal();
    }
    const int right_cam_point = (M == 2) ? (j & 1) : 0;
    for (int i = h - L + 1; i <= h; i++) {
      const double* Ri = &R[(size_t)i * 9];
      const double* ti = &cg[(size_t)i * 6];
      const double p[3] = {Ri[0] * X[0] + Ri[1] * X[1] + Ri[2] * X[2] + ti[0], Ri[3] * X[0] + Ri[4] * X[1] + Ri[5] * X[2] + ti[1],
                           Ri[6] * X[0] + Ri[7] * X[1] + Ri[8] * X[2] + ti[2]};
      Rng orng(spec->seed, 3, (uint64_t)j * (uint64_t)nc + (uint64_t)i);
      double ul = k.fx0 * p[0] / p[2] + k.cx0, vl = k.fy0 * p[1] / p[2] + k.cy0;
      double ur = k.fx1 * (p[0] - k.baseline) / p[2] + k.cx1, vr = vl;
      if (!spec->full_tracks) {
        if (p[2] < 1.0 || ul < 0 || ul > kImgW || ur < 0 || ur > kImgW || vl < 0 || vl > kImgH) continue;
      }
      const bool outlier = orng.uniform() < spec->outlier_fraction;
      double obs4[4];
      if (M == 2 && right_cam_point) {
        // StereoRightError projects the baseline-shifted point with K[0] (BundleAdjuster.h:119-124)
        ul = k.fx0 * (p[0] - k.baseline) / p[2] + k.cx0;
      }
      if (outlier) {
        obs4[0] = orng.uniform(0.0, kImgW); obs4[1] = orng.uniform(0.0, kImgH);
        obs4[2] = orng.uniform(0.0, kImgW); obs4[3] = orng.uniform(0.0, kImgH);
      } else {
        obs4[0] = ul + spec->pixel_sigma * orng.normal(); obs4[1] = vl + spec->pixel_sigma * orng.normal();
        obs4[2] = ur + spec->pixel_sigma * orng.normal(); obs4[3] = vr + spec->pixel_sigma * orng.normal();
      }
      if (no >= max_obs) return UBA_ERR_INVALID_ARGUMENT;
      double* f = feats + (size_t)no * M;
      for (int m = 0; m < M; m++) f[m] = (double)(float)obs4[m];
      cam_idx[no] = i; pt_idx[no] = j;
      if (cam_id) cam_id[no] = right_cam_point;
      no++;
    }
  }
  return no;
}

// Defaults of the pose-only mode (uba_vo.cu); host-only so that libuba_host.so carries them.
void uba_vo_params_default(uba_vo_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  // StereoVisualOdometry::parameters() and VisualOdometry::parameters() (vo/StereoVisualOdometry.h:34, vo/VisualOdometry.h:32)
  p->fu1 = p->fv1 = p->fu2 = p->fv2 = 1.0; p->baseline = 1.0;
  p->method = 0; p->max_iter = 100; p->e1 = 1e-3; p->e2 = 1e-12; p->e3 = 1e-12; p->e4 = 1e-15; p->inlier_threshold = 2.0;
}

// getMotion() (src/vo/StereoVisualOdometry.cpp:331-342): the 4x4 row-major [R(euler)^T | t; 0 0 0 1] of a state
// {roll, pitch, yaw, tx, ty, tz} — the matrix the device code applies to the points (uba_vo.cu: vo_pose).
void uba_vo_pose_matrix(const double state6[6], double T16[16]) {
  if (!state6 || !T16) return;
  const double sr = std::sin(state6[0]), cr = std::cos(state6[0]), sp = std::sin(state6[1]), cp = std::cos(state6[1]);
  const double sy = std::sin(state6[2]), cy = std::cos(state6[2]);
  const double R[3][3] = {{cp * cy, cp * sy, -sp}, {sp * sr * cy - cr * sy, sr * sp * sy + cr * cy, cp * sr}, {cr * sp * cy + sr * sy, cr * sp * sy - sr * cy, cp * cr}};
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T16[i * 4 + j] = R[j][i]; T16[i * 4 + 3] = state6[3 + i]; }
  T16[12] = T16[13] = T16[14] = 0.0; T16[15] = 1.0;
}

// reproject() (:116-141): predicted {left x, left y, right x, right y} of n homogeneous points under `state6`
void uba_vo_predict(const uba_vo_params* P, const double state6[6], int n, const double* pts4, double* pred4) {
  if (!P || !state6 || !pts4 || !pred4) return;
  double T[16];
  uba_vo_pose_matrix(state6, T);
  for (int k = 0; k < n; k++) {
    const double* X = pts4 + (size_t)k * 4;
    double q[4];
    for (int i = 0; i < 4; i++) q[i] = T[i * 4] * X[0] + T[i * 4 + 1] * X[1] + T[i * 4 + 2] * X[2] + T[i * 4 + 3] * X[3];
    const double lz = q[2], rz = q[2];
    pred4[(size_t)k * 4 + 0] = (P->fu1 * q[0] + P->cu1 * q[2]) / lz;
    pred4[(size_t)k * 4 + 1] = (P->fv1 * q[1] + P->cv1 * q[2]) / lz;
    pred4[(size_t)k * 4 + 2] = (P->fu2 * q[0] + P->cu2 * q[2] - P->baseline * P->fu2 * q[3]) / rz;
    pred4[(size_t)k * 4 + 3] = (P->fv2 * q[1] + P->cv2 * q[2]) / rz;
  }
}

// ---- point sharding of one large window ---------------------------------------
// Points go to ranks by KEYFRAME RANGE: order them by (first keyframe of the track, caller index) and cut that order
// into n_ranks pieces with equal observation counts.  A rank's tracks then start inside one contiguous camera range,
// so its Schur products touch only that stretch of the block band of the reduced camera system (plus the track length)
// and the cross-rank sum involves neighbouring ranks only.  Points without observations go to the last rank.
int uba_shard_points(int n_cams, int n_pts, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx, int n_ranks,
                     int32_t* pt_rank, int64_t* rank_obs, int32_t* rank_pts) {
  if (n_cams <= 0 || n_pts < 0 || n_obs < 0 || n_ranks <= 0 || !pt_rank || (n_obs && (!cam_idx || !pt_idx))) return UBA_ERR_INVALID_ARGUMENT;
  std::vector<int32_t> lo((size_t)n_pts, n_cams), cnt((size_t)n_pts, 0);
  for (int64_t o = 0; o < n_obs; o++) {
    const int c = cam_idx[o], p = pt_idx[o];
    if (c < 0 || c >= n_cams || p < 0 || p >= n_pts) return UBA_ERR_INVALID_ARGUMENT;
    if (c < lo[p]) lo[p] = c;
    cnt[p]++;
  }
  std::vector<int64_t> before((size_t)n_cams + 2, 0);      // observations of the tracks starting before keyframe c
  for (int p = 0; p < n_pts; p++) before[(size_t)lo[p] + 1] += cnt[p];
  for (int c = 0; c <= n_cams; c++) before[(size_t)c + 1] += before[c];
  std::vector<int64_t> run(before.begin(), before.end() - 1);
  if (rank_obs) for (int r = 0; r < n_ranks; r++) rank_obs[r] = 0;
  if (rank_pts) for (int r = 0; r < n_ranks; r++) rank_pts[r] = 0;
  for (int p = 0; p < n_pts; p++) {
    int r = n_ranks - 1;
    if (cnt[p] > 0) {
      const int64_t pos = run[lo[p]];                    // observations ahead of this track in (first keyframe, index) order
      run[lo[p]] += cnt[p];
      r = (int)((__int128)pos * n_ranks / (n_obs > 0 ? n_obs : 1));
      if (r >= n_ranks) r = n_ranks - 1;
    }
    pt_rank[p] = r;
    if (rank_obs) rank_obs[r] += cnt[p];
    if (rank_pts) rank_pts[r]++;
  }
  return UBA_OK;
}

// One rank's shard: its points (caller order kept) with window-local point indices renumbered 0..; returns the number of
// observations written.  Output arrays are sized from uba_shard_points' rank_obs / rank_pts.
int64_t uba_shard_extract(int M, int n_pts, int64_t n_obs, const double* pts3, const double* feats, const int32_t* cam_idx,
                          const int32_t* pt_idx, const int32_t* cam_id, const int32_t* pt_rank, int rank, double* pts3_out,
                          double* feats_out, int32_t* cam_idx_out, int32_t* pt_idx_out, int32_t* cam_id_out, int32_t* pt_ids_out) {
  if ((M != 2 && M != 4) || n_pts < 0 || n_obs < 0 || !pt_rank || (n_pts && !pts3) || (n_obs && (!feats || !cam_idx || !pt_idx)))
    return UBA_ERR_INVALID_ARGUMENT;
  std::vector<int32_t> local((size_t)n_pts, -1);
  int np = 0;
  for (int p = 0; p < n_pts; p++) {
    if (pt_rank[p] != rank) continue;
    local[p] = np;
    if (pts3_out) std::memcpy(pts3_out + (size_t)np * 3, pts3 + (size_t)p * 3, sizeof(double) * 3);
    if (pt_ids_out) pt_ids_out[np] = p;
    np++;
  }
  int64_t no = 0;
  for (int64_t o = 0; o < n_obs; o++) {
    const int p = pt_idx[o];
    if (p < 0 || p >= n_pts) return UBA_ERR_INVALID_ARGUMENT;
    if (local[p] < 0) continue;
    if (feats_out) std::memcpy(feats_out + (size_t)no * M, feats + (size_t)o * M, sizeof(double) * M);
    if (cam_idx_out) cam_idx_out[no] = cam_idx[o];
    if (pt_idx_out) pt_idx_out[no] = local[p];
    if (cam_id_out) cam_id_out[no] = cam_id ? cam_id[o] : 0;
    no++;
  }
  return no;
}

// Solver defaults: the reference's hard-coded options (BundleAdjuster.h:416-420,:463-467) over the Ceres defaults it
// inherits.  Lives here (host-only translation unit) so that libuba_host.so carries it without any CUDA dependency.
void uba_config_default(uba_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->loss_kind = UBA_LOSS_HUBER;          // new ceres::HuberLoss(1.0), BundleAdjuster.h:397,:447
  c->loss_scale = 1.0;
  c->max_iterations = 50;                 // Ceres default
  c->function_tolerance = 1e-3;           // :419,:466
  c->gradient_tolerance = 1e-10;
  c->parameter_tolerance = 1e-8;
  c->initial_radius = 1e4;
  c->max_radius = 1e16;
  c->min_radius = 1e-32;
  c->min_relative_decrease = 1e-3;
  c->min_lm_diagonal = 1e-6;
  c->max_lm_diagonal = 1e32;
  c->max_consecutive_invalid_steps = 5;
  c->max_solver_time_s = 1.0;             // :417,:464
  c->fixed_iterations = 0;
  c->jacobi_scaling = 1;
  c->use_bounds = 1;                      // :455-460
  const char* lr = std::getenv("LOCAL_RANK");
  c->device = lr ? std::atoi(lr) : 0;
  c->linearizer = 0;
  c->compute_covariance = 0;              // CalibrationParameters::compute_cov defaults to false (:42-43)
  c->solver = 0;
  c->sliding_window = 0;
}

// log_map_Quat (rotation_utils.h:199-204) with acos clamped to [-1, 1].
void uba_log_map_quat(const double q[4], double r[3]) {
  const double norm = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double theta = norm < 1e-10 ? 1e-10 : norm;
  const double w = q[0] > 1.0 ? 1.0 : (q[0] < -1.0 ? -1.0 : q[0]);
  const double f = std::acos(w) * 2.0 / theta;
  r[0] = f * q[1]; r[1] = f * q[2]; r[2] = f * q[3];
}

// exp_map_Quat (rotation_utils.h:190-197); the Quat constructor normalises.
void uba_exp_map_quat(const double r[3], double q[4]) {
  const double norm = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double theta = norm < 1e-10 ? 1e-10 : norm;
  const double s = std::sin(theta / 2) / theta;
  q[0] = std::cos(theta / 2); q[1] = r[0] * s; q[2] = r[1] * s; q[3] = r[2] * s;
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}

}  // extern "C"
