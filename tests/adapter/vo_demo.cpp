// tests/adapter/vo_demo.cpp — an "application" of the drop-in me::StereoVisualOdometry, written the way a user of the
// reference drives the class: fill parameters, seed rand(), process() a vector of StereoOdoMatchesf, read getMotion(),
// getInliers_idx(), getPts3D(), getPredictions().  Input and output go through small binary files so that
// tests/test_adapter.py can compare with the C-ABI path and the oracle.
//   vo_demo --compile-only
//   vo_demo <in.bin> <out.bin>
//     in:  int32 n, int32 seed, int32 ransac, int32 n_ransac, int32 method, double fu, fv, cu, cv, baseline, thr, init[6], float quads[n][8]
//     out: int32 ok, int32 n_inliers, double state-as-motion[16], int32 inliers[n_inliers], double pts[n][4], double pred[n_inliers][4]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vo/StereoVisualOdometry.h"

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "--compile-only") return 0;
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[5];
  double cal[6], init6[6];
  if (std::fread(hdr, 4, 5, f) != 5 || std::fread(cal, 8, 6, f) != 6 || std::fread(init6, 8, 6, f) != 6) return 4;
  const int n = hdr[0];
  std::vector<float> q((size_t)n * 8);
  if (std::fread(q.data(), 4, q.size(), f) != q.size()) return 4;
  std::fclose(f);

  me::StereoVisualOdometry::parameters p;
  p.fu1 = p.fu2 = cal[0]; p.fv1 = p.fv2 = cal[1]; p.cu1 = p.cu2 = cal[2]; p.cv1 = p.cv2 = cal[3]; p.baseline = cal[4]; p.inlier_threshold = cal[5];
  p.ransac = hdr[2] != 0; p.n_ransac = hdr[3];
  p.method = hdr[4] ? me::VisualOdometry::Method::LM : me::VisualOdometry::Method::GN;
  me::StereoVisualOdometry vo(p);
  std::vector<me::StereoOdoMatchesf> matches;
  for (int i = 0; i < n; i++) {
    const float* m = &q[(size_t)i * 8];
    matches.push_back(me::StereoOdoMatchesf(cv::Point2f(m[0], m[1]), cv::Point2f(m[2], m[3]), cv::Point2f(m[4], m[5]), cv::Point2f(m[6], m[7])));
  }
  cv::Mat init = cv::Mat::zeros(6, 1, CV_64F);
  for (int i = 0; i < 6; i++) init.at<double>(i, 0) = init6[i];
  std::srand((unsigned)hdr[1]);
  const bool ok = vo.process(matches, init);
  cv::Mat T = vo.getMotion();
  std::vector<int> inl = vo.getInliers_idx();
  std::vector<me::ptH3D> pts = vo.getPts3D();
  std::vector<std::pair<me::ptH2D, me::ptH2D> > pred = vo.getPredictions();

  f = std::fopen(argv[2], "wb");
  if (!f) return 5;
  const int32_t o[2] = {ok ? 1 : 0, (int32_t)inl.size()};
  std::fwrite(o, 4, 2, f);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { const double v = T.at<double>(i, j); std::fwrite(&v, 8, 1, f); }
  for (size_t i = 0; i < inl.size(); i++) { const int32_t v = inl[i]; std::fwrite(&v, 4, 1, f); }
  for (size_t i = 0; i < pts.size(); i++) for (int c = 0; c < 4; c++) { const double v = pts[i](c); std::fwrite(&v, 8, 1, f); }
  for (size_t i = 0; i < pred.size(); i++) {
    const double v[4] = {pred[i].first(0), pred[i].first(1), pred[i].second(0), pred[i].second(1)};
    std::fwrite(v, 8, 4, f);
  }
  std::fclose(f);
  return 0;
}
