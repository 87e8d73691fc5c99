// tests/adapter/adapter_demo.cpp — drives the drop-in me::optimisation::BundleAdjuster<M> exactly like an
// application of the reference would (tracks + poses in, poses + points out) and dumps the result.
//   usage: adapter_demo <M> <n_cams> <n_pts> <seed> <fixed_frames> <out.bin>     (needs a B200)
//   usage: adapter_demo --compile-only
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "MotionEstimation/optimisation/BundleAdjuster.h"

using namespace me;
using namespace me::optimisation;

template <int M, typename Feat>
static Feat make_feat(const double* f);
template <>
cv::Point2f make_feat<2, cv::Point2f>(const double* f) { return cv::Point2f((float)f[0], (float)f[1]); }
template <>
std::pair<cv::Point2f, cv::Point2f> make_feat<4, std::pair<cv::Point2f, cv::Point2f>>(const double* f) {
  return std::make_pair(cv::Point2f((float)f[0], (float)f[1]), cv::Point2f((float)f[2], (float)f[3]));
}

template <int M, typename Feat>
static int run(int n_cams, int n_pts, unsigned long long seed, int fixed, const char* out_path) {
  uba_calib k; uba_synth_default_calib(&k);
  uba_synth_spec spec; std::memset(&spec, 0, sizeof(spec));
  spec.M = M; spec.n_cams = n_cams; spec.n_pts = n_pts; spec.track_min = 3; spec.track_max = n_cams; spec.pixel_sigma = 0.5;
  spec.pose_t_sigma = 0.05; spec.pose_r_sigma = 0.005; spec.point_rel_sigma = 0.01; spec.fixed_frames = fixed; spec.seed = seed;
  const long long max_obs = (long long)n_cams * n_pts;
  std::vector<double> cams((size_t)n_cams * 6), pts((size_t)n_pts * 3), feats((size_t)max_obs * M);
  std::vector<int32_t> ci(max_obs), pi(max_obs), cid(max_obs);
  const long long no = uba_synth_generate(&spec, &k, max_obs, nullptr, cams.data(), nullptr, pts.data(), feats.data(), ci.data(), pi.data(), cid.data());
  if (no < 0) return 2;
  const int first_frame = 100;  // frame numbering of the "sequence": the window starts at frame 100
  std::vector<CamPose_qd> poses;
  for (int i = 0; i < n_cams; i++) {
    double q[4]; uba_exp_map_quat(&cams[(size_t)i * 6 + 3], q);
    poses.push_back(CamPose_qd(first_frame + i, Quatd(q[0], q[1], q[2], q[3]), cv::Vec3d(cams[(size_t)i * 6], cams[(size_t)i * 6 + 1], cams[(size_t)i * 6 + 2])));
  }
  std::vector<WBA_Point<Feat>> tracks;
  long long o = 0;
  for (int j = 0; j < n_pts; j++) {
    const ptH3D X{pts[(size_t)j * 3], pts[(size_t)j * 3 + 1], pts[(size_t)j * 3 + 2], 1.0};
    bool first = true;
    for (; o < no && pi[o] == j; o++) {
      const Feat f = make_feat<M, Feat>(&feats[(size_t)o * M]);
      if (first) { tracks.push_back(WBA_Point<Feat>(f, first_frame + ci[o], cid[o], X)); first = false; }
      else tracks.back().addMatch(f, first_frame + ci[o]);
    }
    if (first) return 3;  // the generator gives every point at least one observation
  }
  std::vector<cv::Matx33d> K(2, cv::Matx33d{k.fx0, 0, k.cx0, 0, k.fy0, k.cy0, 0, 0, 1});
  CalibrationParameters calib(K, k.feat_var, k.baseline);
  calib.compute_cov = (M == 4);  // extract_covariance path (BundleAdjuster.h:471-472)
  BundleAdjuster<M> ba(calib, poses, tracks);
  if (ba.getStatus() != BundleAdjuster<M>::Status::INITIALISED) return 4;
  const typename BundleAdjuster<M>::Status st = ba.optimise(fixed);
  std::vector<CamPose_qd> out_poses = ba.getCameraPoses();
  std::vector<pt3D> out_pts = ba.getPoints();
  FILE* f = std::fopen(out_path, "wb");
  if (!f) return 5;
  const int hdr[4] = {(int)st, ba.getNbCameras(), ba.getNbPoints(), ba.getNbObservations()};
  std::fwrite(hdr, sizeof(int), 4, f);
  for (const auto& p : out_poses) {
    const double v[8] = {p.orientation.w(), p.orientation.x(), p.orientation.y(), p.orientation.z(), p.position(0), p.position(1), p.position(2), (double)p.ID};
    std::fwrite(v, sizeof(double), 8, f);
  }
  for (const auto& p : out_pts) std::fwrite(p.val, sizeof(double), 3, f);
  std::vector<cv::Mat> covs = ba.getPosesCovariance();
  const int ncov = (int)covs.size();
  std::fwrite(&ncov, sizeof(int), 1, f);
  for (const auto& c : covs) { std::vector<double> z(36, 0.0); std::fwrite(c.empty() ? z.data() : c.d.data(), sizeof(double), 36, f); }
  std::fclose(f);
  // single use: a second optimise() must refuse and keep the status (BundleAdjuster.h:381-384)
  if (ba.optimise(fixed) != st) return 6;
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 2 && !std::strcmp(argv[1], "--compile-only")) return 0;
  if (argc != 7) { std::fprintf(stderr, "usage: adapter_demo M n_cams n_pts seed fixed out.bin\n"); return 1; }
  const int M = std::atoi(argv[1]);
  if (M == 4) return run<4, std::pair<cv::Point2f, cv::Point2f>>(std::atoi(argv[2]), std::atoi(argv[3]), std::strtoull(argv[4], nullptr, 10), std::atoi(argv[5]), argv[6]);
  return run<2, cv::Point2f>(std::atoi(argv[2]), std::atoi(argv[3]), std::strtoull(argv[4], nullptr, 10), std::atoi(argv[5]), argv[6]);
}
