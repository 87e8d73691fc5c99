// oracle/refstub/opencv2/calib3d/calib3d.hpp — cv::Rodrigues only (see ../core/core.hpp for what this directory is).
#ifndef UBA_REFSTUB_OPENCV_CALIB3D_HPP
#define UBA_REFSTUB_OPENCV_CALIB3D_HPP
#include "../core/core.hpp"

namespace cv {
namespace refstub {
inline void rodrigues_vec_to_mat(const double r[3], double R[9]) {
  const double th = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 2.220446049250313e-16) { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0); return; }
  const double c = std::cos(th), s = std::sin(th), c1 = 1 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
  R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
inline void rodrigues_mat_to_vec(const double R[9], double r[3]) {
  double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  const double s = 0.5 * std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  c = c > 1 ? 1 : (c < -1 ? -1 : c);
  const double th = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
    double t = (R[0] + 1) * 0.5; r[0] = std::sqrt(std::max(t, 0.0));
    t = (R[4] + 1) * 0.5; r[1] = std::sqrt(std::max(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5; r[2] = std::sqrt(std::max(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
    if (std::fabs(r[0]) < std::fabs(r[1]) && std::fabs(r[0]) < std::fabs(r[2]) && (R[5] > 0) != (r[1] * r[2] > 0)) r[2] = -r[2];
    const double nn = th / std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    for (int i = 0; i < 3; i++) r[i] *= nn;
    return;
  }
  const double k = 0.5 * th / s;
  for (int i = 0; i < 3; i++) r[i] = ax[i] * k;
}
inline int count(const Mat& a) { return (int)a.total(); }
template <typename T, int m, int n> int count(const Matx<T, m, n>&) { return m * n; }
inline double get(const Mat& a, int i) { return a.at<double>(i / a.cols, i % a.cols); }
template <typename T, int m, int n> double get(const Matx<T, m, n>& a, int i) { return (double)a.val[i]; }
inline void put(Mat& a, int rows, int cols, const double* v) { a.create(rows, cols, CV_64F); for (int i = 0; i < rows * cols; i++) a.at<double>(i / cols, i % cols) = v[i]; }
template <typename T, int m, int n> void put(Matx<T, m, n>& a, int rows, int cols, const double* v) { assert(rows * cols == m * n); for (int i = 0; i < m * n; i++) a.val[i] = static_cast<T>(v[i]); }
}  // namespace refstub

// rotation vector <-> rotation matrix, either direction, chosen by the number of input elements
template <typename A, typename B>
void Rodrigues(const A& src, B& dst) {
  double in[9], out[9];
  const int k = refstub::count(src);
  for (int i = 0; i < k && i < 9; i++) in[i] = refstub::get(src, i);
  if (k == 3) { refstub::rodrigues_vec_to_mat(in, out); refstub::put(dst, 3, 3, out); }
  else { assert(k == 9); refstub::rodrigues_mat_to_vec(in, out); refstub::put(dst, 3, 1, out); }
}
}  // namespace cv
#endif
