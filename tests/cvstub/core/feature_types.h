// tests/cvstub/core/feature_types.h — a MINIMAL stand-in for the parts of OpenCV and of the
// reference's core/feature_types.h + core/rotation_utils.h that the drop-in BundleAdjuster.h and StereoVisualOdometry.h touch,
// so the adapter can be compiled and exercised in a container without OpenCV C++ headers.  Written
// from the interface, not from the reference sources; it is test scaffolding only —
// a real build includes the reference's own header instead.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <deque>
#include <utility>
#include <vector>

#define CV_64F 6
namespace cv {
template <typename T, int m, int n>
struct Matx {
  T val[m * n];
  Matx() { for (T& v : val) v = T(0); }
  Matx(std::initializer_list<T> l) { int i = 0; for (T v : l) val[i++] = v; for (; i < m * n; i++) val[i] = T(0); }
  template <typename... A, typename = typename std::enable_if<sizeof...(A) == m * n && (m * n > 1)>::type>
  Matx(A... a) : val{T(a)...} {}
  T& operator()(int i, int j) { return val[i * n + j]; }
  const T& operator()(int i, int j) const { return val[i * n + j]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
};
template <typename T, int n>
struct Vec : Matx<T, n, 1> {
  Vec() {}
  Vec(std::initializer_list<T> l) : Matx<T, n, 1>(l) {}
  Vec(T a, T b, T c) { this->val[0] = a; this->val[1] = b; this->val[2] = c; }
};
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<double, 6, 1> Matx61d;
typedef Matx<double, 3, 1> Matx31d;
typedef Matx<double, 4, 1> Matx41d;
typedef Vec<double, 3> Vec3d;
template <typename T>
struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int r, int c, int, double* p) : rows(r), cols(c), d(p, p + r * c) {}
  void copyTo(Mat& o) const { o = *this; }
  bool empty() const { return d.empty(); }
  // what the drop-in StereoVisualOdometry.h touches
  int type() const { return CV_64F; }
  Mat clone() const { return *this; }
  static Mat zeros(int r, int c, int) { Mat m; m.rows = r; m.cols = c; m.d.assign((size_t)r * c, 0.0); return m; }
  template <typename T> T& at(int i, int j) { return d[(size_t)i * cols + j]; }
  template <typename T> const T& at(int i, int j) const { return d[(size_t)i * cols + j]; }
};
}  // namespace cv

namespace me {
typedef cv::Matx31d pt3D;
typedef cv::Matx41d ptH3D;
typedef cv::Matx31d ptH2D;
template <typename T>
struct StereoMatch {
  T f1, f2;
  float m_score;
  StereoMatch() : m_score(-1) {}
  StereoMatch(const T& a, const T& b, float score = -1) : f1(a), f2(b), m_score(score) {}
};
template <typename T>
struct StereoOdoMatches : public StereoMatch<T> {
  T f3, f4;
  StereoOdoMatches() {}
  StereoOdoMatches(const T& a, const T& b, const T& c, const T& d, float score = -1) : StereoMatch<T>(a, b, score), f3(c), f4(d) {}
};
typedef StereoMatch<cv::Point2f> StereoMatchf;
typedef StereoOdoMatches<cv::Point2f> StereoOdoMatchesf;
inline pt3D to_euclidean(const ptH3D& p) { return pt3D{p(0) / p(3), p(1) / p(3), p(2) / p(3)}; }

template <typename T>
class Quat {
  T m_w, m_x, m_y, m_z;
 public:
  Quat(T w = 1, T x = 0, T y = 0, T z = 0) : m_w(w), m_x(x), m_y(y), m_z(z) {
    const T n = std::sqrt(w * w + x * x + y * y + z * z);
    if (n > 0) { m_w /= n; m_x /= n; m_y /= n; m_z /= n; }
  }
  double w() const { return m_w; }
  double x() const { return m_x; }
  double y() const { return m_y; }
  double z() const { return m_z; }
};
typedef Quat<double> Quatd;

template <class O, class T>
class CamPose {
 public:
  O orientation;
  cv::Vec<T, 3> position;
  cv::Mat Cov;
  int ID;
  CamPose(int id = 0, const O& e = O(), const cv::Vec<T, 3>& v = cv::Vec<T, 3>(), const cv::Mat& c = cv::Mat()) : orientation(e), position(v), Cov(c), ID(id) {}
};
typedef CamPose<Quatd, double> CamPose_qd;

template <typename T>
struct WBA_Point {
  WBA_Point(const T match, const int frame_nb, const int cam = 0, const ptH3D pt_ = ptH3D{0, 0, 0, 1}) : pt(pt_), camID(cam) { addMatch(match, frame_nb); }
  void addMatch(const T match, const int frame_nb) { features.push_back(match); indices.push_back((unsigned)frame_nb); }
  T getFeat(unsigned int i) const { return features[i]; }
  unsigned int getFrameIdx(unsigned int i) const { return indices[i]; }
  unsigned int getNbFeatures() const { return (unsigned)features.size(); }
  int getCameraID() const { return camID; }
  ptH3D get3DLocation() const { return pt; }
 private:
  std::deque<T> features;
  std::deque<unsigned int> indices;
  ptH3D pt;
  int camID;
};
}  // namespace me
