#ifndef UBA_DROPIN_BUNDLEADJUSTER_H_INCLUDED
#define UBA_DROPIN_BUNDLEADJUSTER_H_INCLUDED
/** \file BundleAdjuster.h
 *  \brief Drop-in replacement for the reference's Ceres-backed windowed bundle adjuster.
 *
 *  Same namespace, class template, public types and member signatures as
 *  include/MotionEstimation/optimisation/BundleAdjuster.h of abeauvisage/uasl_motion_estimation
 *  (class :182-278, Observation :23-33, CalibrationParameters :35-45), so callers compile unchanged.
 *  There is no <ceres/...> include: optimise() packs the window into flat arrays and hands it to
 *  libuba (include/uba.h), whose sm_100a CUDA kernels do everything ceres::Solve did (:422, :469).
 *
 *  Deliberate differences from the reference header:
 *   - the raw constructor takes `const VecCams& cams` (the reference declares `const VecObs& cams`
 *     and initialises a VecCams with it, :196, which cannot compile once instantiated);
 *   - the stereo constructor has no default for `obs` (the reference's default has the mono type, :220);
 *   - camIdx / ptIdx are range-checked (the reference reads out of bounds when a track outlives the
 *     window, :370-371); an offending window ends in Status::FAILED;
 *   - log_map_Quat's acos argument is clamped to [-1, 1] (rotation_utils.h:203 yields NaN for w > 1);
 *   - glog is not initialised (nothing logs through it any more);
 *   - every point carries bounds (:455-460), so Ceres runs a projected Armijo line search after each trust-region step;
 *     libuba clamps the full step to the box and accepts or rejects it: results differ from a Ceres run only on steps
 *     whose clamped full step fails sufficient decrease (none on the benchmark windows);
 *   - getPosesCovariance() (CalibrationParameters::compute_cov) returns the 6x6 blocks of the inverse
 *     undamped reduced camera matrix, i.e. the camera blocks of (J^T J)^-1 that ceres::Covariance
 *     computes (:502-512); fixed cameras get a zero 6x6 block instead of Ceres' refusal; on failure the
 *     reference's own message "[Bundle Adjuster] error computing the covariance matrix" is printed (:525).
 *  Like the reference, the object is single use (:7): UNINITIALISED -> INITIALISED -> SUCCESSFUL | FAILED.
 */
#include <array>
#include <cmath>
#include <iostream>
#include <utility>
#include <vector>

#include "core/feature_types.h"
#include "uba.h"

namespace me {
namespace optimisation {

//! One 2-D (N = 2) or stereo (N = 4) measurement of a 3-D point in a frame of the window.
template <int N>
struct Observation {
  std::array<double, N> data;
  int camIdx;
  int ptIdx;
  int camID;
  Observation(const std::array<double, N>& d, const int cam, const int pt, const int id = 0) : data(d), camIdx(cam), ptIdx(pt), camID(id) {}
  const double& operator[](int i) const { return data[i]; }
  double& operator[](int i) { return data[i]; }
};

struct CalibrationParameters {
  std::vector<cv::Matx33d> K;
  double feat_var;
  double baseline;
  bool compute_cov;
  explicit CalibrationParameters(const cv::Matx33d& K_, double var, double baseline_ = 0.0) : K(1, K_), feat_var(var), baseline(baseline_), compute_cov(false) {}
  CalibrationParameters(const std::vector<cv::Matx33d>& K_, double var, double baseline_ = 0.0) : K(K_), feat_var(var), baseline(baseline_), compute_cov(false) {}
};

namespace detail {
inline void unpack(const cv::Point2f& p, double* out) { out[0] = p.x; out[1] = p.y; }
inline void unpack(const cv::Point2d& p, double* out) { out[0] = p.x; out[1] = p.y; }
template <typename T>
inline void unpack(const std::pair<cv::Point_<T>, cv::Point_<T>>& p, double* out) {
  out[0] = p.first.x; out[1] = p.first.y; out[2] = p.second.x; out[3] = p.second.y;
}
}  // namespace detail

template <int M>
class BundleAdjuster {
 public:
  enum class Status { UNINITIALISED, INITIALISED, SUCCESSFUL, FAILED };
  using VecPts = std::vector<pt3D>;
  using VecCams = std::vector<cv::Matx61d>;
  using VecObs = std::vector<Observation<M>>;

  explicit BundleAdjuster(const CalibrationParameters& params, const VecCams& cams, const VecPts& pts, const VecObs& obs = VecObs())
      : calib_params(params), m_status(Status::UNINITIALISED), m_camera_params(cams), m_point_params(pts) {
    initialiseObservations(obs);
  }
  //! monocular windowed BA
  template <typename T1, typename T2>
  BundleAdjuster(const CalibrationParameters& params, const std::vector<CamPose<me::Quat<T1>, T1>>& cams,
                 const std::vector<WBA_Point<cv::Point_<T2>>>& obs = std::vector<WBA_Point<cv::Point_<T2>>>())
      : calib_params(params), m_status(Status::UNINITIALISED) {
    initialiseParameters(cams);
    if (!cams.empty()) initialiseObservations(obs, cams[0].ID);
  }
  //! stereo windowed BA
  template <typename T1, typename T2>
  BundleAdjuster(const CalibrationParameters& params, const std::vector<CamPose<me::Quat<T1>, T1>>& cams,
                 const std::vector<WBA_Point<std::pair<cv::Point_<T2>, cv::Point_<T2>>>>& obs)
      : calib_params(params), m_status(Status::UNINITIALISED) {
    initialiseParameters(cams);
    if (!cams.empty()) initialiseObservations(obs, cams[0].ID);
  }
  ~BundleAdjuster() { if (m_handle) uba_destroy(m_handle); }
  BundleAdjuster(const BundleAdjuster&) = delete;
  BundleAdjuster& operator=(const BundleAdjuster&) = delete;

  std::vector<pt3D> getPoints() const { return m_point_params; }
  std::vector<CamPose_qd> getCameraPoses() const {
    std::vector<CamPose_qd> out;
    int id = 0;  // frame IDs are renumbered from 0, as in the reference (:233-235)
    for (const auto& c : m_camera_params) {
      const double r[3] = {c(3), c(4), c(5)};
      double q[4];
      uba_exp_map_quat(r, q);
      out.push_back(CamPose_qd{id++, Quatd{q[0], q[1], q[2], q[3]}, cv::Vec3d{c(0), c(1), c(2)}});
    }
    return out;
  }
  std::vector<cv::Mat> getPosesCovariance() { return m_camera_covs; }
  std::vector<cv::Mat> getPointsCovariance() { return m_point_covs; }
  int getNbPoints() const { return (int)m_point_params.size(); }
  int getNbCameras() const { return (int)m_camera_params.size(); }
  int getNbObservations() const { return (int)m_observations.size(); }
  Status getStatus() { return m_status; }
  //! last libuba diagnostic (not in the reference)
  const char* lastError() const { return uba_last_error(m_handle); }

  void initialiseParameters(const VecCams& cams, const VecPts& pts) {
    if (!uninitialised("system should be uninitialised!")) return;
    m_camera_params = cams;
    m_point_params = pts;
  }
  template <typename T>
  void initialiseParameters(const std::vector<CamPose<me::Quat<T>, T>>& cams, const std::vector<pt3D>& pts = std::vector<pt3D>()) {
    if (!uninitialised("system should be uninitialised!")) return;
    m_point_params = pts;
    m_camera_params.clear();
    for (const auto& pose : cams) {
      const double q[4] = {pose.orientation.w(), pose.orientation.x(), pose.orientation.y(), pose.orientation.z()};
      double r[3];
      uba_log_map_quat(q, r);  // [t, angle-axis], p_cam = R(r) X + t
      m_camera_params.push_back(cv::Matx61d{(double)pose.position(0), (double)pose.position(1), (double)pose.position(2), r[0], r[1], r[2]});
    }
  }
  void initialiseObservations(const VecObs& observations) {
    if (m_status != Status::UNINITIALISED || m_camera_params.empty() || m_point_params.empty()) {
      std::cerr << "[Bundle Adjuster] system should be uninitialised and both cameras and points not empty!" << std::endl;
      return;
    }
    m_observations = observations;
    m_status = Status::INITIALISED;
  }
  //! tracks from feature_types.h; first_frame is the frame index of the window's first pose
  template <typename F>
  void initialiseObservations(const std::vector<WBA_Point<F>>& tracks, const int first_frame) {
    if (m_status != Status::UNINITIALISED || m_camera_params.empty()) {
      std::cerr << "[Bundle Adjuster] system should be uninitialised and cameras not empty!" << std::endl;
      return;
    }
    const bool init_points = m_point_params.empty();
    m_observations.clear();
    int pt_idx = 0;
    for (const auto& track : tracks) {  // point-major, frame-ascending: the order libuba keeps
      if (init_points) m_point_params.push_back(to_euclidean(track.get3DLocation()));
      for (unsigned int i = 0; i < track.getNbFeatures(); i++) {
        const int cam = (int)track.getFrameIdx(i) - first_frame;
        if (cam < 0) continue;
        std::array<double, M> d;
        detail::unpack(track.getFeat(i), d.data());
        m_observations.push_back(Observation<M>(d, cam, pt_idx, track.getCameraID()));
      }
      pt_idx++;  // counts points that end up without observations too
    }
    m_status = Status::INITIALISED;
  }

  //! runs the optimisation; cameras with camIdx < fixedFrames stay constant
  Status optimise(int fixedFrames) {
    if (m_status != Status::INITIALISED) {
      std::cerr << "[Bundle Adjuster] system should be initiliased to perform optimisation!" << std::endl;
      return m_status;
    }
    std::cout << "[Bundle Adjuster] optimising (" << m_camera_params.size() << " cam poses and " << m_point_params.size() << " pts with "
              << m_observations.size() << " observations." << std::endl;
    const int nc = getNbCameras(), np = getNbPoints(), no = getNbObservations();
    std::vector<double> cams((size_t)nc * 6), pts((size_t)np * 3), feats((size_t)no * M);
    std::vector<int32_t> ci(no), pi(no), cid(no);
    for (int i = 0; i < nc; i++) for (int a = 0; a < 6; a++) cams[(size_t)i * 6 + a] = m_camera_params[i](a);
    for (int j = 0; j < np; j++) for (int a = 0; a < 3; a++) pts[(size_t)j * 3 + a] = m_point_params[j](a);
    for (int o = 0; o < no; o++) {
      for (int m = 0; m < M; m++) feats[(size_t)o * M + m] = m_observations[o].data[m];
      ci[o] = m_observations[o].camIdx; pi[o] = m_observations[o].ptIdx; cid[o] = m_observations[o].camID;
    }
    uba_calib k;
    const cv::Matx33d& K0 = calib_params.K.at(0);
    const cv::Matx33d& K1 = calib_params.K.size() > 1 ? calib_params.K[1] : calib_params.K[0];
    k.fx0 = K0(0, 0); k.fy0 = K0(1, 1); k.cx0 = K0(0, 2); k.cy0 = K0(1, 2);
    k.fx1 = K1(0, 0); k.fy1 = K1(1, 1); k.cx1 = K1(0, 2); k.cy1 = K1(1, 2);
    k.feat_var = calib_params.feat_var; k.baseline = calib_params.baseline;
    uba_config cfg;
    uba_config_default(&cfg);  // Huber(1.0), SPARSE_SCHUR-equivalent LM, function_tolerance 1e-3, 1 s cap
    cfg.compute_covariance = calib_params.compute_cov ? 1 : 0;
    m_status = Status::FAILED;
    if (!m_handle && uba_create(&cfg, &m_handle) != UBA_OK) {
      std::cerr << "[Bundle Adjuster] " << uba_last_error(nullptr) << std::endl;
      return m_status;
    }
    if (uba_set_problem(m_handle, M, nc, np, no, cams.data(), pts.data(), feats.data(), ci.data(), pi.data(), cid.data(), &k) != UBA_OK) {
      std::cerr << "[Bundle Adjuster] " << uba_last_error(m_handle) << std::endl;
      return m_status;
    }
    uba_summary summary;
    const int rc = uba_optimise(m_handle, fixedFrames, &summary);
    if (rc != UBA_OK && rc != UBA_ERR_NUMERICAL && rc != UBA_ERR_INFEASIBLE) {
      std::cerr << "[Bundle Adjuster] " << uba_last_error(m_handle) << std::endl;
      return m_status;
    }
    // Ceres optimises in place through the raw double* of the parameter vectors (:404, :449)
    if (uba_get_cameras(m_handle, cams.data()) == UBA_OK && uba_get_points(m_handle, pts.data()) == UBA_OK) {
      for (int i = 0; i < nc; i++) for (int a = 0; a < 6; a++) m_camera_params[i](a) = cams[(size_t)i * 6 + a];
      for (int j = 0; j < np; j++) for (int a = 0; a < 3; a++) m_point_params[j](a) = pts[(size_t)j * 3 + a];
    }
    if (calib_params.compute_cov) extract_covariance();
    m_status = summary.usable ? Status::SUCCESSFUL : Status::FAILED;  // summary.IsSolutionUsable() (:427, :474)
    return m_status;
  }

 private:
  bool uninitialised(const char* msg) const {
    if (m_status == Status::UNINITIALISED) return true;
    std::cerr << "[Bundle Adjuster] " << msg << std::endl;
    return false;
  }
  void extract_covariance() {
    std::vector<double> cov((size_t)getNbCameras() * 36);
    if (uba_get_pose_covariances(m_handle, cov.data()) != UBA_OK) {
      std::cerr << "[Bundle Adjuster] error computing the covariance matrix" << std::endl;
      return;
    }
    m_camera_covs.assign(getNbCameras(), cv::Mat());
    for (int i = 0; i < getNbCameras(); i++) cv::Mat(6, 6, CV_64F, cov.data() + (size_t)i * 36).copyTo(m_camera_covs[i]);
  }

  CalibrationParameters calib_params;
  Status m_status;
  VecCams m_camera_params;
  VecPts m_point_params;
  VecObs m_observations;
  std::vector<cv::Mat> m_camera_covs;
  std::vector<cv::Mat> m_point_covs;
  uba_handle* m_handle = nullptr;
};

}  // namespace optimisation
}  // namespace me

#endif
