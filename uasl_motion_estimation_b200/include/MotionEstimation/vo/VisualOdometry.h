// MotionEstimation/vo/VisualOdometry.h — the base class of the reference's visual odometry (include/MotionEstimation/vo/
// VisualOdometry.h:7-49), restated for the B200 drop-in: same names, same parameters and defaults, nothing else.
#ifndef UBA_DROPIN_VISUAL_ODOMETRY_H
#define UBA_DROPIN_VISUAL_ODOMETRY_H

#include "core/feature_types.h"   // cv::Mat (the reference includes opencv2/core/core.hpp here)

namespace me {

class VisualOdometry {
 public:
  enum class Method { GN, LM };   // Gauss-Newton, Levenberg-Marquardt (:16)

  struct parameters {   // (:20-36)
    Method method;
    double step_size;
    double eps, e1, e2, e3, e4;
    int max_iter;
    int nb_fixed_frames;
    bool ransac;
    int n_ransac;
    double inlier_threshold;
    parameters()
        : method(Method::GN), step_size(1.0), eps(1e-9), e1(1e-3), e2(1e-12), e3(1e-12), e4(1e-15), max_iter(100),
          nb_fixed_frames(2), ransac(true), n_ransac(200), inlier_threshold(2.0) {}
  };

  virtual cv::Mat getMotion() = 0;
  VisualOdometry() {}
  virtual ~VisualOdometry() {}
};

}  // namespace me
#endif
