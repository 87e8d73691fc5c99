// MotionEstimation/vo/StereoVisualOdometry.h — drop-in for the reference's me::StereoVisualOdometry
// (include/MotionEstimation/vo/StereoVisualOdometry.h:18-90, src/vo/StereoVisualOdometry.cpp) on top of libuba's pose-only
// mode (include/uba.h: uba_vo_*).  Header-only, C++11, no Ceres, no OpenCV beyond the types of the public interface.
//
// Same public interface and behaviour: parameters and defaults; process() needs at least 6 matches, resets a malformed
// `init` to zero, draws n_ransac triples with rand() exactly as selectRandomIndices does (:143-163 — one rand() % n per
// draw, duplicates redrawn; every iteration consumes its draws whether or not the triple passes the area test, so a
// program that seeds srand() sees the same hypotheses), keeps the EARLIEST hypothesis with the most inliers (:68-71),
// prints the reference's "[Motion Estimation] N inliers" line, refines over the inliers from `init` and returns whether
// that optimisation converged; getMotion() is [R(euler)^T | t]; getPts3D / getInliers_idx / getPredictions / getParams
// as in the reference.
//
// What runs where: triangulation, all RANSAC hypotheses (3-point fit + inlier scoring, one CTA each) and the final
// refinement run on the GPU behind three calls; the class keeps only the parameters, the state, the inlier list.
//
// Deliberate differences: optimize()'s loop ends after max_iter iterations (the reference compares the
// counter with an enum value, :277, and does not terminate on ordinary data); `weighting` is carried but unused, as in
// the reference (process passes weight = false); the library's error text goes to std::cerr and process() returns false.
#ifndef UBA_DROPIN_STEREO_VISUAL_ODOMETRY_H
#define UBA_DROPIN_STEREO_VISUAL_ODOMETRY_H

#include <cstdlib>
#include <iostream>
#include <utility>
#include <vector>

#include "core/feature_types.h"
#include "uba.h"
#include "vo/VisualOdometry.h"

namespace me {

class StereoVisualOdometry : public VisualOdometry {
 public:
  struct parameters : public VisualOdometry::parameters {   // (:24-35)
    double baseline;
    bool weighting;
    double fu1, fv1, fu2, fv2;
    double cu1, cu2;
    double cv1, cv2;
    parameters() : baseline(1.0), weighting(false), fu1(1.0), fv1(1.0), fu2(1.0), fv2(1.0), cu1(0.0), cu2(0.0), cv1(0.0), cv2(0.0) {}
  };

  StereoVisualOdometry(parameters param = parameters()) : m_param(param), m_h(nullptr) { for (int i = 0; i < 6; i++) m_state(i) = 0.0; }
  ~StereoVisualOdometry() { if (m_h) uba_destroy(m_h); }
  StereoVisualOdometry(const StereoVisualOdometry& o) : VisualOdometry(), m_pts3D(o.m_pts3D), m_inliers_idx(o.m_inliers_idx), m_param(o.m_param), m_state(o.m_state), m_h(nullptr) {}
  StereoVisualOdometry& operator=(const StereoVisualOdometry& o) {
    if (this != &o) { m_pts3D = o.m_pts3D; m_inliers_idx = o.m_inliers_idx; m_param = o.m_param; m_state = o.m_state; }
    return *this;   // the device handle is per object and created on first use
  }

  bool process(const std::vector<StereoOdoMatchesf>& matches, cv::Mat init = cv::Mat::zeros(6, 1, CV_64F)) {
    if (init.rows != 6 || init.cols != 1 || init.type() != CV_64F) init = cv::Mat::zeros(6, 1, CV_64F);   // (:36-38)
    if (matches.size() < 6) return false;                                                                  // (:40-42)
    double init6[6];
    for (int i = 0; i < 6; i++) { init6[i] = init.at<double>(i, 0); m_state(i) = init6[i]; }
    m_inliers_idx.clear();
    if (!m_h) {
      uba_config cfg;
      uba_config_default(&cfg);
      if (uba_create(&cfg, &m_h) != UBA_OK) { std::cerr << "[Motion Estimation] libuba: " << uba_last_error(nullptr) << std::endl; m_h = nullptr; return false; }
    }
    // project3D + updateObservations (:23-32, :145-163 of the .cpp): the four features of every match, packed
    const int n = (int)matches.size();
    std::vector<float> quads((size_t)n * 8);
    for (int i = 0; i < n; i++) {
      const StereoOdoMatchesf& m = matches[i];
      float* q = &quads[(size_t)i * 8];
      q[0] = m.f1.x; q[1] = m.f1.y; q[2] = m.f2.x; q[3] = m.f2.y; q[4] = m.f3.x; q[5] = m.f3.y; q[6] = m.f4.x; q[7] = m.f4.y;
    }
    const uba_vo_params p = vo_params();
    if (!ok(uba_vo_set_matches(m_h, &p, n, quads.data()))) return false;
    std::vector<double> pts((size_t)n * 4);
    if (!ok(uba_vo_get_points(m_h, pts.data()))) return false;
    m_pts3D.resize(n);
    for (int i = 0; i < n; i++) m_pts3D[i] = ptH3D(pts[(size_t)i * 4], pts[(size_t)i * 4 + 1], pts[(size_t)i * 4 + 2], pts[(size_t)i * 4 + 3]);

    std::vector<int32_t> inliers;
    if (m_param.ransac) {
      // selectRandomIndices(3, n) for every iteration, in the reference's order (:55-58)
      std::vector<int32_t> triples((size_t)std::max(m_param.n_ransac, 0) * 3);
      for (int it = 0; it < m_param.n_ransac; it++) {
        int got = 0;
        int32_t* t = &triples[(size_t)it * 3];
        while (got < 3) {
          const int idx = std::rand() % n;
          bool exists = false;
          for (int j = 0; j < got; j++) if (t[j] == idx) exists = true;
          if (!exists) t[got++] = idx;
        }
      }
      if (m_param.n_ransac > 0) {
        int32_t best = -1;
        if (!ok(uba_vo_ransac(m_h, init6, m_param.n_ransac, triples.data(), &best, nullptr, nullptr, nullptr))) return false;
        inliers.resize(n);
        int32_t n_in = 0;
        if (!ok(uba_vo_get_inliers(m_h, inliers.data(), &n_in))) return false;
        inliers.resize(n_in);
      }
    } else {
      inliers.resize(n);
      for (int i = 0; i < n; i++) inliers[i] = i;   // (:73-77)
    }
    m_inliers_idx.assign(inliers.begin(), inliers.end());
    std::cout << "[Motion Estimation] " << m_inliers_idx.size() << " inliers" << std::endl;   // (:83)
    if (m_inliers_idx.size() < 6) return false;                                               // (:85,:91-92)
    double state6[6];
    int32_t converged = 0, iterations = 0;
    if (!ok(uba_vo_refine(m_h, init6, (int)inliers.size(), inliers.data(), state6, &converged, &iterations))) return false;
    for (int i = 0; i < 6; i++) m_state(i) = state6[i];   // optimize() leaves its last iterate in m_state either way
    return converged != 0;
  }

  virtual cv::Mat getMotion() {   // (:331-342)
    double T[16];
    uba_vo_pose_matrix(&m_state(0), T);
    return cv::Mat(4, 4, CV_64F, T).clone();
  }

  std::vector<ptH3D> getPts3D() { return m_pts3D; }
  std::vector<int> getInliers_idx() { return m_inliers_idx; }
  std::vector<std::pair<ptH2D, ptH2D> > getPredictions() {   // reproject(m_state, m_inliers_idx) (:116-141)
    const size_t k = m_inliers_idx.size();
    std::vector<double> pts(k * 4), pred(k * 4);
    for (size_t i = 0; i < k; i++) for (int c = 0; c < 4; c++) pts[i * 4 + c] = m_pts3D[m_inliers_idx[i]](c);
    const uba_vo_params p = vo_params();
    uba_vo_predict(&p, &m_state(0), (int)k, pts.data(), pred.data());
    std::vector<std::pair<ptH2D, ptH2D> > out(k);
    for (size_t i = 0; i < k; i++) out[i] = std::make_pair(ptH2D(pred[i * 4], pred[i * 4 + 1], 1.0), ptH2D(pred[i * 4 + 2], pred[i * 4 + 3], 1.0));
    return out;
  }
  parameters getParams() { return m_param; }

 private:
  std::vector<ptH3D> m_pts3D;       // 3D features in the previous frame (normalised homogeneous)
  std::vector<int> m_inliers_idx;   // indices of the inliers
  parameters m_param;
  cv::Matx61d m_state;              // three Euler angles and the translation
  uba_handle* m_h;

  uba_vo_params vo_params() const {
    uba_vo_params p;
    uba_vo_params_default(&p);
    p.fu1 = m_param.fu1; p.fv1 = m_param.fv1; p.cu1 = m_param.cu1; p.cv1 = m_param.cv1;
    p.fu2 = m_param.fu2; p.fv2 = m_param.fv2; p.cu2 = m_param.cu2; p.cv2 = m_param.cv2; p.baseline = m_param.baseline;
    p.method = m_param.method == Method::LM ? 1 : 0; p.max_iter = m_param.max_iter;
    p.e1 = m_param.e1; p.e2 = m_param.e2; p.e3 = m_param.e3; p.e4 = m_param.e4; p.inlier_threshold = m_param.inlier_threshold;
    return p;
  }
  bool ok(int rc) const {
    if (rc == UBA_OK) return true;
    std::cerr << "[Motion Estimation] libuba: " << uba_last_error(m_h) << std::endl;
    return false;
  }
};

}  // namespace me
#endif
