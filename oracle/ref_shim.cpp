// oracle/ref_shim.cpp — C entry points around the REFERENCE'S OWN SOURCES, compiled from where they lie under
// /root/reference (never copied) into oracle/_ref/libuba_ref.so by oracle/Makefile:
//   include/MotionEstimation/optimisation/BundleAdjuster.h   (functors, problem build, options: the whole class)
//   include/MotionEstimation/core/{feature_types,rotation_utils}.h, src/core/rotation_utils.cpp
//   include/MotionEstimation/vo/StereoVisualOdometry.h, src/vo/StereoVisualOdometry.cpp
// against the minimal Ceres / OpenCV stand-ins of oracle/refstub (neither library is installed here).
//
// TEST INFRASTRUCTURE: the tests use this library to pin oracle/uba_oracle.cpp (and through it the CUDA path) to the
// reference's text: residual rows and autodiff Jacobians of the three functors, log/exp map, the observation ordering of
// initialiseObservations, bounds / fixed cameras / options of optimise(), Status mapping, the stereo VO arithmetic.
// `private` is opened for this translation unit only, to read the tables the class keeps to itself.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>

#define private public
#define protected public
#include "optimisation/BundleAdjuster.h"
#include "vo/StereoVisualOdometry.h"
#undef private
#undef protected

#include "../include/uba.h"

using namespace me;
using namespace me::optimisation;

namespace {

CalibrationParameters make_calib(const uba_calib* k) {
  std::vector<cv::Matx33d> K;
  K.push_back(cv::Matx33d(k->fx0, 0, k->cx0, 0, k->fy0, k->cy0, 0, 0, 1));
  K.push_back(cv::Matx33d(k->fx1, 0, k->cx1, 0, k->fy1, k->cy1, 0, 0, 1));
  return CalibrationParameters(K, k->feat_var, k->baseline);
}

template <typename Functor, int M>
int eval_functor(const Functor* f, const double* cam6, const double* pt3, double* r, double* Jc, double* Jp) {
  // the reference's own route to a Jacobian: AutoDiffCostFunction<Functor, M, 6, 3> (BundleAdjuster.h:97-102,:133-138,:174-179)
  ceres::AutoDiffCostFunction<Functor, M, 6, 3> cost(const_cast<Functor*>(f));
  const double* params[2] = {cam6, pt3};
  double* jac[2] = {Jc, Jp};
  return cost.Evaluate(params, r, (Jc || Jp) ? jac : nullptr) ? 0 : -1;
}

struct Silence {   // the reference prints progress lines to std::cout
  std::streambuf* old; std::ostringstream sink;
  Silence() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Silence() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

// Residual rows (and, when Jc / Jp are non-null, the autodiff Jacobians [M][6], [M][3]) of ONE observation through the
// reference's functor: M = 4 StereoReprojectionError; M = 2 StandardReprojectionError (cam_id == 0) or StereoRightError,
// selected like optimise() does (BundleAdjuster.h:399-402), with sigma = sqrt(feat_var) (:400,:448).
int uba_refsrc_residual(int M, const uba_calib* k, const double* cam6, const double* pt3, const double* obs, int cam_id,
                        double* r, double* Jc, double* Jp) {
  CalibrationParameters calib = make_calib(k);
  const double sigma = sqrt(calib.feat_var);
  if (M == 4) return eval_functor<StereoReprojectionError, 4>(new StereoReprojectionError(obs[0], obs[1], obs[2], obs[3], &calib, sigma), cam6, pt3, r, Jc, Jp);
  if (M != 2) return -1;
  if (calib.baseline == 0) calib.baseline = 0.5;   // BundleAdjuster.h:389-390
  if (cam_id == 0) return eval_functor<StandardReprojectionError, 2>(new StandardReprojectionError(obs[0], obs[1], &calib, sigma), cam6, pt3, r, Jc, Jp);
  return eval_functor<StereoRightError, 2>(new StereoRightError(obs[0], obs[1], &calib, sigma), cam6, pt3, r, Jc, Jp);
}

void uba_refsrc_log_map_quat(const double q[4], double r[3]) {
  // NB: the Quat constructor normalises (rotation_utils.h:120)
  const cv::Vec3d v = log_map_Quat(Quatd(q[0], q[1], q[2], q[3]));
  r[0] = v(0); r[1] = v(1); r[2] = v(2);
}
void uba_refsrc_exp_map_quat(const double r[3], double q[4]) {
  const Quatd Q = exp_map_Quat(cv::Vec3d(r[0], r[1], r[2]));
  q[0] = Q.w(); q[1] = Q.x(); q[2] = Q.y(); q[3] = Q.z();
}

// The whole class, end to end, through the WBA-point constructors (BundleAdjuster.h:206-228):
//   poses     [n_cams][7]  quaternion (w,x,y,z) + position, with their frame IDs cam_ids[n_cams] (cams[0].ID = first frame)
//   tracks    CSR track_off[n_pts+1] over (frame_idx, feats[M]); one WBA point per track, camera ID pt_cam_id[j],
//             homogeneous location pts4[j]
// Outputs (any may be null): the observation table the class built (m_observations, in its order), the packed initial
// camera vectors (m_camera_params after initialiseParameters), and after optimise(fixed_frames): camera vectors, points,
// getCameraPoses() quaternions + IDs, Status (0..3 in the enum's order), 6x6 pose covariances when compute_cov.
int uba_refsrc_ba_run(int M, const uba_calib* k, int compute_cov, int n_cams, const double* poses7, const int32_t* cam_ids,
                      int n_pts, const double* pts4, const int32_t* pt_cam_id, const int64_t* track_off, const int32_t* frame_idx,
                      const double* feats, int fixed_frames, int run_optimise, int64_t max_obs, int64_t* n_obs_out,
                      int32_t* cam_idx_out, int32_t* pt_idx_out, int32_t* cam_id_out, double* feat_out, double* cams6_init_out,
                      double* cams6_out, double* pts3_out, double* quat_id_out /*[n_cams][5] w x y z ID*/, int32_t* status_out,
                      double* cov_out /*[n_cams][36]*/) {
  Silence quiet;
  CalibrationParameters calib = make_calib(k);
  calib.compute_cov = compute_cov != 0;
  std::vector<CamPose_qd> cams;
  for (int c = 0; c < n_cams; c++) {
    const double* p = poses7 + (size_t)c * 7;
    cams.push_back(CamPose_qd(cam_ids[c], Quatd(p[0], p[1], p[2], p[3]), cv::Vec3d(p[4], p[5], p[6])));
  }
  auto finish = [&](auto& ba) -> int {
    const int64_t no = (int64_t)ba.m_observations.size();
    if (n_obs_out) *n_obs_out = no;
    if (no > max_obs) return -2;
    for (int64_t o = 0; o < no; o++) {
      const auto& ob = ba.m_observations[(size_t)o];
      if (cam_idx_out) cam_idx_out[o] = ob.camIdx;
      if (pt_idx_out) pt_idx_out[o] = ob.ptIdx;
      if (cam_id_out) cam_id_out[o] = ob.camID;
      if (feat_out) for (int m = 0; m < M; m++) feat_out[(size_t)o * M + m] = ob.data[m];
    }
    if (cams6_init_out) for (int c = 0; c < n_cams; c++) for (int a = 0; a < 6; a++) cams6_init_out[(size_t)c * 6 + a] = ba.m_camera_params[c](a);
    int status = (int)ba.getStatus();
    if (run_optimise) status = (int)ba.optimise(fixed_frames);
    if (status_out) *status_out = status;
    if (cams6_out) for (int c = 0; c < n_cams; c++) for (int a = 0; a < 6; a++) cams6_out[(size_t)c * 6 + a] = ba.m_camera_params[c](a);
    if (pts3_out) { const auto pts = ba.getPoints(); for (int j = 0; j < (int)pts.size() && j < n_pts; j++) for (int a = 0; a < 3; a++) pts3_out[(size_t)j * 3 + a] = pts[j](a); }
    if (quat_id_out) {
      const auto out = ba.getCameraPoses();
      for (int c = 0; c < (int)out.size() && c < n_cams; c++) {
        double* q = quat_id_out + (size_t)c * 5;
        q[0] = out[c].orientation.w(); q[1] = out[c].orientation.x(); q[2] = out[c].orientation.y(); q[3] = out[c].orientation.z(); q[4] = out[c].ID;
      }
    }
    if (cov_out && run_optimise && compute_cov) {
      auto covs = ba.getPosesCovariance();
      for (int c = 0; c < (int)covs.size() && c < n_cams; c++)
        for (int i = 0; i < 36; i++) cov_out[(size_t)c * 36 + i] = covs[c].empty() ? 0.0 : covs[c].template at<double>(i / 6, i % 6);
    }
    return 0;
  };
  if (M == 4) {
    typedef std::pair<cv::Point2f, cv::Point2f> Feat;
    std::vector<WBA_Point<Feat>> tracks;
    for (int j = 0; j < n_pts; j++) {
      const int64_t b = track_off[j], e = track_off[j + 1];
      if (e <= b) return -3;   // a WBA point is born with its first match
      const double* X = pts4 + (size_t)j * 4;
      auto feat = [&](int64_t i) { const double* f = feats + (size_t)i * 4; return Feat(cv::Point2f((float)f[0], (float)f[1]), cv::Point2f((float)f[2], (float)f[3])); };
      WBA_Point<Feat> wp(feat(b), frame_idx[b], pt_cam_id ? pt_cam_id[j] : 0, cv::Matx22d::zeros(), ptH3D(X[0], X[1], X[2], X[3]));
      for (int64_t i = b + 1; i < e; i++) wp.addMatch(feat(i), frame_idx[i]);
      tracks.push_back(wp);
    }
    BundleAdjuster<4> ba(calib, cams, tracks);
    return finish(ba);
  }
  if (M == 2) {
    std::vector<WBA_Point<cv::Point2f>> tracks;
    for (int j = 0; j < n_pts; j++) {
      const int64_t b = track_off[j], e = track_off[j + 1];
      if (e <= b) return -3;
      const double* X = pts4 + (size_t)j * 4;
      auto feat = [&](int64_t i) { const double* f = feats + (size_t)i * 2; return cv::Point2f((float)f[0], (float)f[1]); };
      WBA_Point<cv::Point2f> wp(feat(b), frame_idx[b], pt_cam_id ? pt_cam_id[j] : 0, cv::Matx22d::zeros(), ptH3D(X[0], X[1], X[2], X[3]));
      for (int64_t i = b + 1; i < e; i++) wp.addMatch(feat(i), frame_idx[i]);
      tracks.push_back(wp);
    }
    BundleAdjuster<2> ba(calib, cams, tracks);
    return finish(ba);
  }
  return -1;
}

// ---- stereo visual odometry (src/vo/StereoVisualOdometry.cpp) --------------------------------------------------------
// params10: fu1 fv1 cu1 cv1 fu2 fv2 cu2 cv2 baseline inlier_threshold; opt6: method (0 GN, 1 LM), max_iter, e1, e2, e3, e4
static StereoVisualOdometry::parameters vo_params(const double* p10, const double* opt6, int ransac, int n_ransac) {
  StereoVisualOdometry::parameters P;
  P.fu1 = p10[0]; P.fv1 = p10[1]; P.cu1 = p10[2]; P.cv1 = p10[3]; P.fu2 = p10[4]; P.fv2 = p10[5]; P.cu2 = p10[6]; P.cv2 = p10[7];
  P.baseline = p10[8]; P.inlier_threshold = p10[9];
  if (opt6) { P.method = opt6[0] != 0 ? VisualOdometry::Method::LM : VisualOdometry::Method::GN; P.max_iter = (int)opt6[1]; P.e1 = opt6[2]; P.e2 = opt6[3]; P.e3 = opt6[4]; P.e4 = opt6[5]; }
  P.ransac = ransac != 0; P.n_ransac = n_ransac;
  return P;
}
static std::vector<StereoOdoMatchesf> vo_matches(int n, const double* quads8) {
  std::vector<StereoOdoMatchesf> m;
  for (int i = 0; i < n; i++) {
    const double* q = quads8 + (size_t)i * 8;   // previous left, previous right, current left, current right (x, y each)
    m.push_back(StereoOdoMatchesf(cv::Point2f((float)q[0], (float)q[1]), cv::Point2f((float)q[2], (float)q[3]), cv::Point2f((float)q[4], (float)q[5]), cv::Point2f((float)q[6], (float)q[7])));
  }
  return m;
}

// project3D (:22-32): triangulation from disparity; out pts4 [n][4] (normalised homogeneous)
int uba_refsrc_vo_project3d(const double* p10, int n, const double* quads8, double* pts4) {
  StereoVisualOdometry vo(vo_params(p10, nullptr, 0, 0));
  vo.project3D(vo_matches(n, quads8));
  for (int i = 0; i < n; i++) for (int a = 0; a < 4; a++) pts4[(size_t)i * 4 + a] = vo.m_pts3D[i](a);
  return 0;
}

// One evaluation at `state6` (3 Euler angles, translation) over `selection`: predictions (reproject :116-143), residuals
// as optimize() forms them (:183-189, observed - predicted), the 6 x 4n Jacobian of updateJacobian (:291-329), and the
// normal equations A = J J^T, B = J r (:199-203).  Any output may be null.
int uba_refsrc_vo_linearize(const double* p10, int n, const double* quads8, const double* state6, int n_sel, const int32_t* selection,
                            double* pred4, double* res4, double* J, double* A36, double* B6) {
  Silence quiet;
  StereoVisualOdometry vo(vo_params(p10, nullptr, 0, 0));
  auto matches = vo_matches(n, quads8);
  vo.project3D(matches);
  vo.updateObservations(matches);
  for (int a = 0; a < 6; a++) vo.m_state(a) = state6[a];
  std::vector<int> sel(selection, selection + n_sel);
  auto pred = vo.reproject(vo.m_state, sel);
  cv::Mat residuals(4 * n_sel, 1, CV_64F);
  for (int i = 0; i < n_sel; i++) {
    const double p[4] = {pred[i].first(0), pred[i].first(1), pred[i].second(0), pred[i].second(1)};
    const double o[4] = {vo.m_obs[sel[i]].first(0), vo.m_obs[sel[i]].first(1), vo.m_obs[sel[i]].second(0), vo.m_obs[sel[i]].second(1)};
    for (int q = 0; q < 4; q++) { if (pred4) pred4[(size_t)i * 4 + q] = p[q]; residuals.at<double>(i * 4 + q) = o[q] - p[q]; if (res4) res4[(size_t)i * 4 + q] = o[q] - p[q]; }
  }
  vo.updateJacobian(sel);
  if (J) for (int a = 0; a < 6; a++) for (int c = 0; c < 4 * n_sel; c++) J[(size_t)a * 4 * n_sel + c] = vo.m_J.at<double>(a, c);
  cv::Mat A = vo.m_J * vo.m_J.t(), B = vo.m_J * residuals;
  if (A36) for (int i = 0; i < 36; i++) A36[i] = A.at<double>(i / 6, i % 6);
  if (B6) for (int i = 0; i < 6; i++) B6[i] = B.at<double>(i, 0);
  return 0;
}

// optimize() (:165-283) from `state6` over `selection`; returns its bool, the final state in state_out; then
// computeInliers() (:94-114) at that state: inliers_out [n] (capacity), n_inliers_out.
int uba_refsrc_vo_optimize(const double* p10, const double* opt6, int n, const double* quads8, const double* state6, int n_sel,
                           const int32_t* selection, double* state_out, int32_t* inliers_out, int32_t* n_inliers_out) {
  Silence quiet;
  StereoVisualOdometry vo(vo_params(p10, opt6, 0, 0));
  auto matches = vo_matches(n, quads8);
  vo.project3D(matches);
  vo.updateObservations(matches);
  for (int a = 0; a < 6; a++) vo.m_state(a) = state6[a];
  const bool ok = vo.optimize(std::vector<int>(selection, selection + n_sel), false);
  for (int a = 0; a < 6; a++) state_out[a] = vo.m_state(a);
  if (inliers_out || n_inliers_out) {
    const std::vector<int> in = vo.computeInliers();
    if (n_inliers_out) *n_inliers_out = (int)in.size();
    if (inliers_out) for (size_t i = 0; i < in.size(); i++) inliers_out[i] = in[i];
  }
  return ok ? 1 : 0;
}

// computeInliers() (:94-114) at a given state
int uba_refsrc_vo_inliers(const double* p10, int n, const double* quads8, const double* state6, int32_t* inliers_out) {
  Silence quiet;
  StereoVisualOdometry vo(vo_params(p10, nullptr, 0, 0));
  auto matches = vo_matches(n, quads8);
  vo.project3D(matches);
  vo.updateObservations(matches);
  for (int a = 0; a < 6; a++) vo.m_state(a) = state6[a];
  const std::vector<int> in = vo.computeInliers();
  if (inliers_out) for (size_t i = 0; i < in.size(); i++) inliers_out[i] = in[i];
  return (int)in.size();
}

// process() (:34-92) with RANSAC driven by the C library's rand() after srand(seed); returns its bool; the motion as
// state6 (Euler angles + translation), the inlier indices.
int uba_refsrc_vo_process(const double* p10, const double* opt6, int n_ransac, unsigned seed, int n, const double* quads8,
                          double* state_out, int32_t* inliers_out, int32_t* n_inliers_out) {
  Silence quiet;
  StereoVisualOdometry vo(vo_params(p10, opt6, 1, n_ransac));
  srand(seed);
  const bool ok = vo.process(vo_matches(n, quads8));
  for (int a = 0; a < 6; a++) state_out[a] = vo.m_state(a);
  const std::vector<int> in = vo.getInliers_idx();
  if (n_inliers_out) *n_inliers_out = (int)in.size();
  if (inliers_out) for (size_t i = 0; i < in.size(); i++) inliers_out[i] = in[i];
  return ok ? 1 : 0;
}

}  // extern "C"
