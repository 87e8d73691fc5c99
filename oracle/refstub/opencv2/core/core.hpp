// oracle/refstub/opencv2/core/core.hpp — a MINIMAL stand-in for the slice of the OpenCV C++ API that the reference's
// core/rotation_utils.{h,cpp}, core/feature_types.h, optimisation/BundleAdjuster.h and vo/StereoVisualOdometry.cpp use.
//
// TEST INFRASTRUCTURE.  OpenCV's C++ headers/libraries are not installed in this image (only the Python cv2 module), so
// the reference's own sources cannot be compiled against the real thing.  This header exists so that those sources can be
// compiled VERBATIM, from where they lie under /root/reference, into oracle/_ref/libuba_ref.so (recipe: oracle/Makefile),
// which the tests use to pin the oracle.  Written from the public OpenCV API documentation; fixed-size matrices
// (cv::Matx / cv::Vec), a small dense double matrix (cv::Mat, CV_64F only) and the handful of free functions used.
// Nothing under uasl_motion_estimation_b200/ includes it.
#ifndef UBA_REFSTUB_OPENCV_CORE_HPP
#define UBA_REFSTUB_OPENCV_CORE_HPP

#include <sys/types.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 5
#define CV_64FC1 6

namespace cv {

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4 };

class Mat;

// ---- fixed-size matrices --------------------------------------------------------------------
template <typename T, int m, int n>
class Matx {
 public:
  enum { rows = m, cols = n, channels = m * n };
  T val[m * n];

  Matx() { for (int i = 0; i < m * n; i++) val[i] = T(0); }
  template <typename... A, typename = typename std::enable_if<(sizeof...(A) >= 1 && sizeof...(A) <= m * n) && !(sizeof...(A) == 1 && m * n != 1)>::type>
  Matx(A... a) {
    const T tmp[] = {static_cast<T>(a)...};
    for (int i = 0; i < m * n; i++) val[i] = i < (int)sizeof...(A) ? tmp[i] : T(0);
  }
  explicit Matx(T v0) { for (int i = 0; i < m * n; i++) val[i] = T(0); val[0] = v0; }
  template <typename T2>
  Matx(const Matx<T2, m, n>& o) { for (int i = 0; i < m * n; i++) val[i] = static_cast<T>(o.val[i]); }

  static Matx zeros() { return Matx(); }
  static Matx all(T v) { Matx r; for (int i = 0; i < m * n; i++) r.val[i] = v; return r; }
  static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); i++) r.val[i * n + i] = T(1); return r; }

  const T& operator()(int i, int j) const { return val[i * n + j]; }
  T& operator()(int i, int j) { return val[i * n + j]; }
  const T& operator()(int i) const { return val[i]; }
  T& operator()(int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }

  Matx<T, n, m> t() const { Matx<T, n, m> r; for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) r.val[j * m + i] = val[i * n + j]; return r; }
};

template <typename T, int m, int n, int l>
Matx<T, m, n> operator*(const Matx<T, m, l>& a, const Matx<T, l, n>& b) {
  Matx<T, m, n> r;
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) { T s = T(0); for (int k = 0; k < l; k++) s += a.val[i * l + k] * b.val[k * n + j]; r.val[i * n + j] = s; }
  return r;
}
template <typename T, int m, int n> Matx<T, m, n> operator+(const Matx<T, m, n>& a, const Matx<T, m, n>& b) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = a.val[i] + b.val[i]; return r; }
template <typename T, int m, int n> Matx<T, m, n> operator-(const Matx<T, m, n>& a, const Matx<T, m, n>& b) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = a.val[i] - b.val[i]; return r; }
template <typename T, int m, int n> Matx<T, m, n> operator-(const Matx<T, m, n>& a) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = -a.val[i]; return r; }
template <typename T, int m, int n> Matx<T, m, n>& operator+=(Matx<T, m, n>& a, const Matx<T, m, n>& b) { for (int i = 0; i < m * n; i++) a.val[i] += b.val[i]; return a; }
template <typename T, int m, int n> Matx<T, m, n>& operator-=(Matx<T, m, n>& a, const Matx<T, m, n>& b) { for (int i = 0; i < m * n; i++) a.val[i] -= b.val[i]; return a; }
#define UBA_MATX_SCALAR(S)                                                                                                              \
  template <typename T, int m, int n> Matx<T, m, n> operator*(const Matx<T, m, n>& a, S s) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = static_cast<T>(a.val[i] * s); return r; } \
  template <typename T, int m, int n> Matx<T, m, n> operator*(S s, const Matx<T, m, n>& a) { return a * s; }                            \
  template <typename T, int m, int n> Matx<T, m, n> operator/(const Matx<T, m, n>& a, S s) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = static_cast<T>(a.val[i] / s); return r; }
UBA_MATX_SCALAR(double)
UBA_MATX_SCALAR(float)
UBA_MATX_SCALAR(int)
#undef UBA_MATX_SCALAR
template <typename T, int m, int n>
std::ostream& operator<<(std::ostream& os, const Matx<T, m, n>& a) {
  os << "[";
  for (int i = 0; i < m; i++) { for (int j = 0; j < n; j++) os << a.val[i * n + j] << (j + 1 < n ? ", " : ""); os << (i + 1 < m ? ";\n " : ""); }
  return os << "]";
}

template <typename T, int n>
class Vec : public Matx<T, n, 1> {
 public:
  Vec() {}
  template <typename... A, typename = typename std::enable_if<(sizeof...(A) >= 2 && sizeof...(A) <= n)>::type>
  Vec(A... a) : Matx<T, n, 1>(a...) {}
  explicit Vec(T v0) : Matx<T, n, 1>(v0) {}
  template <typename T2>
  Vec(const Matx<T2, n, 1>& o) : Matx<T, n, 1>(o) {}
};

typedef Matx<double, 2, 1> Matx21d; typedef Matx<double, 3, 1> Matx31d; typedef Matx<double, 4, 1> Matx41d; typedef Matx<double, 6, 1> Matx61d;
typedef Matx<double, 2, 2> Matx22d; typedef Matx<double, 3, 3> Matx33d; typedef Matx<float, 3, 3> Matx33f;
typedef Matx<double, 3, 4> Matx34d; typedef Matx<double, 4, 3> Matx43d; typedef Matx<float, 4, 3> Matx43f;
typedef Matx<double, 4, 4> Matx44d; typedef Matx<float, 4, 4> Matx44f; typedef Matx<double, 6, 6> Matx66d;
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<float, 3> Vec3f; typedef Vec<double, 4> Vec4d; typedef Vec<float, 4> Vec4f;
typedef Vec<double, 6> Vec6d;

template <typename T, int n> double norm(const Matx<T, n, 1>& v) { double s = 0; for (int i = 0; i < n; i++) s += (double)v.val[i] * v.val[i]; return std::sqrt(s); }

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x_, T y_) : x(x_), y(y_) {} };
typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point_<int> Point2i; typedef Point2i Point;
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } double operator[](int i) const { return val[i]; } };

// ---- dense matrix (doubles only; a header over shared storage, so that views alias like in OpenCV) ----------------
class Mat {
 public:
  int rows = 0, cols = 0;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* data) { create(r, c, type); assert(type == CV_64F); std::memcpy(ptr<double>(), data, sizeof(double) * r * c); }
  template <typename T, int m, int n>
  Mat(const Matx<T, m, n>& a) { create(m, n, CV_64F); for (int i = 0; i < m * n; i++) (*buf_)[i] = (double)a.val[i]; }

  void create(int r, int c, int type) {
    type_ = type; rows = r; cols = c; step_ = c; off_ = 0;
    buf_ = std::make_shared<std::vector<double>>((size_t)r * c, 0.0);
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.at<double>(i, i) = 1.0; return m; }
  static Mat eye(Size s, int type) { return eye(s.height, s.width, type); }

  int type() const { return type_; }
  int channels() const { return 1; }
  void convertTo(Mat& dst, int type) const { copyTo(dst); dst.type_ = type; }   // storage is double whatever the nominal type
  Size size() const { return Size(cols, rows); }
  bool empty() const { return rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }

  template <typename T> T& at(int i, int j) { static_assert(sizeof(T) == sizeof(double), "refstub cv::Mat holds doubles"); return (*buf_)[off_ + (size_t)i * step_ + j]; }
  template <typename T> const T& at(int i, int j) const { static_assert(sizeof(T) == sizeof(double), "refstub cv::Mat holds doubles"); return (*buf_)[off_ + (size_t)i * step_ + j]; }
  template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
  template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
  // typed row pointer; only the double case is backed by storage (the reference reads float quaternions through it in a
  // constructor the tests never reach)
  template <typename T> T* ptr(int row = 0) { assert(sizeof(T) == sizeof(double)); return reinterpret_cast<T*>(buf_->data() + off_ + (size_t)row * step_); }
  template <typename T> const T* ptr(int row = 0) const { assert(sizeof(T) == sizeof(double)); return reinterpret_cast<const T*>(buf_->data() + off_ + (size_t)row * step_); }

  Mat clone() const { Mat r(rows, cols, type_); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) r.at<double>(i, j) = at<double>(i, j); return r; }
  void copyTo(Mat& dst) const {
    if (empty()) { dst = Mat(); return; }
    if (dst.rows != rows || dst.cols != cols || !dst.buf_) dst.create(rows, cols, type_);
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) dst.at<double>(i, j) = at<double>(i, j);
  }
  void copyTo(Mat&& view) const { assert(view.rows == rows && view.cols == cols); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) view.at<double>(i, j) = at<double>(i, j); }
  Mat colRange(int a, int b) const { Mat v = *this; v.off_ = off_ + a; v.cols = b - a; return v; }
  Mat rowRange(int a, int b) const { Mat v = *this; v.off_ = off_ + (size_t)a * step_; v.rows = b - a; return v; }
  Mat t() const { Mat r(cols, rows, type_); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) r.at<double>(j, i) = at<double>(i, j); return r; }
  Mat diag() const { const int k = std::min(rows, cols); Mat r(k, 1, type_); for (int i = 0; i < k; i++) r.at<double>(i, 0) = at<double>(i, i); return r; }

  template <typename T, int m, int n>
  operator Matx<T, m, n>() const { assert(rows == m && cols == n); Matx<T, m, n> r; for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) r.val[i * n + j] = static_cast<T>(at<double>(i, j)); return r; }
  template <typename T, int n>
  operator Vec<T, n>() const { assert(rows * cols == n); Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = static_cast<T>(at<double>(i)); return r; }

  Mat& operator+=(const Mat& b) { assert(rows == b.rows && cols == b.cols); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) at<double>(i, j) += b.at<double>(i, j); return *this; }
  Mat& operator-=(const Mat& b) { assert(rows == b.rows && cols == b.cols); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) at<double>(i, j) -= b.at<double>(i, j); return *this; }

 private:
  int type_ = CV_64F;
  size_t step_ = 0, off_ = 0;
  std::shared_ptr<std::vector<double>> buf_;
};

inline Mat operator*(const Mat& a, const Mat& b) {
  assert(a.cols == b.rows);
  Mat r(a.rows, b.cols, CV_64F);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < b.cols; j++) { double s = 0; for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j); r.at<double>(i, j) = s; }
  return r;
}
inline Mat operator+(const Mat& a, const Mat& b) { Mat r = a.clone(); r += b; return r; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat r = a.clone(); r -= b; return r; }
inline Mat operator/(const Mat& a, const Mat& b) { assert(a.rows == b.rows && a.cols == b.cols); Mat r = a.clone(); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<double>(i, j) /= b.at<double>(i, j); return r; }
inline Mat operator*(double s, const Mat& a) { Mat r = a.clone(); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<double>(i, j) *= s; return r; }
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { return (1.0 / s) * a; }
inline std::ostream& operator<<(std::ostream& os, const Mat& a) {
  os << "[";
  for (int i = 0; i < a.rows; i++) { for (int j = 0; j < a.cols; j++) os << a.at<double>(i, j) << (j + 1 < a.cols ? ", " : ""); os << (i + 1 < a.rows ? ";\n " : ""); }
  return os << "]";
}

// cv::Mat_<T>(r, c) << a, b, c ... (comma initialiser)
template <typename T> class Mat_;
template <typename T>
class MatCommaInitializer_ {
 public:
  MatCommaInitializer_(Mat_<T>* m, T first) : m_(m), k_(0) { put(first); }
  MatCommaInitializer_& operator,(T v) { put(v); return *this; }
  operator Mat_<T>() const { return *m_; }
  operator Mat() const { return *m_; }
 private:
  void put(T v) { m_->template at<double>(k_ / m_->cols, k_ % m_->cols) = (double)v; k_++; }
  Mat_<T>* m_;
  int k_;
};
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_64F) {}
  Mat_(const Mat& m) : Mat(m) {}
  Mat_& operator=(const Mat& m) { Mat::operator=(m); return *this; }
  T operator()(int i, int j) const { return static_cast<T>(this->template at<double>(i, j)); }
};
template <typename T, typename T2>
MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) { return MatCommaInitializer_<T>(const_cast<Mat_<T>*>(&m), static_cast<T>(v)); }

// ---- free functions -------------------------------------------------------------------------
inline Scalar sum(const Mat& a) { double s = 0; for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) s += a.at<double>(i, j); return Scalar(s); }
inline double norm(const Mat& a, int type = NORM_L2) {
  double s = 0;
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) {
    const double v = std::fabs(a.at<double>(i, j));
    if (type == NORM_INF) s = std::max(s, v); else if (type == NORM_L1) s += v; else s += v * v;
  }
  return type == NORM_L2 ? std::sqrt(s) : s;
}
inline void minMaxLoc(const Mat& a, double* mn, double* mx) {
  double lo = a.at<double>(0, 0), hi = lo;
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) { lo = std::min(lo, a.at<double>(i, j)); hi = std::max(hi, a.at<double>(i, j)); }
  if (mn) *mn = lo;
  if (mx) *mx = hi;
}
// Least-squares / linear solve by Householder QR (the only method the reference asks for: DECOMP_QR).
inline bool solve(const Mat& A, const Mat& B, Mat& X, int /*flags*/ = DECOMP_LU) {
  const int m = A.rows, n = A.cols, nb = B.cols;
  if (m < n || B.rows != m) return false;
  std::vector<double> a((size_t)m * n), b((size_t)m * nb);
  for (int i = 0; i < m; i++) { for (int j = 0; j < n; j++) a[(size_t)i * n + j] = A.at<double>(i, j); for (int j = 0; j < nb; j++) b[(size_t)i * nb + j] = B.at<double>(i, j); }
  for (int k = 0; k < n; k++) {
    double nrm = 0; for (int i = k; i < m; i++) nrm += a[(size_t)i * n + k] * a[(size_t)i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = a[(size_t)k * n + k] > 0 ? -nrm : nrm;
    std::vector<double> v(m - k);
    for (int i = k; i < m; i++) v[i - k] = a[(size_t)i * n + k];
    v[0] -= alpha;
    double vv = 0; for (double x : v) vv += x * x;
    if (vv > 0) {
      for (int j = k; j < n; j++) { double d = 0; for (int i = k; i < m; i++) d += v[i - k] * a[(size_t)i * n + j]; d = 2 * d / vv; for (int i = k; i < m; i++) a[(size_t)i * n + j] -= d * v[i - k]; }
      for (int j = 0; j < nb; j++) { double d = 0; for (int i = k; i < m; i++) d += v[i - k] * b[(size_t)i * nb + j]; d = 2 * d / vv; for (int i = k; i < m; i++) b[(size_t)i * nb + j] -= d * v[i - k]; }
    }
  }
  double dmax = 0; for (int k = 0; k < n; k++) dmax = std::max(dmax, std::fabs(a[(size_t)k * n + k]));
  for (int k = 0; k < n; k++) if (std::fabs(a[(size_t)k * n + k]) <= 1e-15 * dmax) return false;
  X.create(n, nb, CV_64F);
  for (int j = 0; j < nb; j++)
    for (int i = n - 1; i >= 0; i--) { double s = b[(size_t)i * nb + j]; for (int k = i + 1; k < n; k++) s -= a[(size_t)i * n + k] * X.at<double>(k, j); X.at<double>(i, j) = s / a[(size_t)i * n + i]; }
  return true;
}

}  // namespace cv
#endif
