// uba_oracle.cpp — CPU ORACLE for the windowed bundle-adjustment hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (libuba) never links, loads or calls anything in this directory.
//
// PINNING.  The reference (abeauvisage/uasl_motion_estimation) ships no tests, fixtures or golden vectors for this
// path, and its solver arithmetic lives in Ceres Solver, an un-vendored and un-versioned third-party dependency
// ("Ceres 1.12 minimum", reference README.md:8; find_package(Ceres QUIET ...), CMakeLists.txt:12) that is not installed
// here (nor are OpenCV-C++ and glog).  What pins this file all the same:
//   * PINNED BY THE REFERENCE'S OWN TEXT: oracle/_ref/libuba_ref.so is the reference's BundleAdjuster.h,
//     rotation_utils.cpp and StereoVisualOdometry.cpp compiled where they lie (oracle/Makefile, oracle/ref_shim.cpp)
//     against the minimal Ceres / OpenCV stand-ins of oracle/refstub.  tests/test_ref_pin.py holds this file to it:
//     residual rows and autodiff Jacobians of the three functors bit for bit, log / exp map, the observation table of
//     initialiseObservations, parameter packing, bounds / fixed cameras / options / Status of optimise() with the reference
//     class run end to end (its ceres::Solve being refstub's independent dense restatement) to 1e-12; the committed
//     fixture tests/golden/ref_golden.json carries the same outputs to the GPU box.
//   * STILL RECALLED, NOT PINNED ([CERES-UPSTREAM]): the published Ceres algorithms behind ceres::Solve — the Huber /
//     Cauchy corrector, Jacobi column scaling, the Levenberg-Marquardt trust-region strategy and its accept / reject /
//     termination rules, bound projection.  Real Ceres cannot run here; two independent restatements (this file: per
//     point through the Schur complement; refstub: dense normal equations) agreeing to 1e-12 guard against slips, not
//     against a shared misreading of Ceres.  DESIGN.md section 5 says so.
// This file RESTATES
//   * the reference's residual functors, evaluated like the reference does — through forward-mode dual numbers (the role
//     ceres::Jet<double,9> plays for AutoDiffCostFunction<.,M,6,3>, BundleAdjuster.h:97-102,:133-138,:174-179);
//   * AngleAxisRotatePoint, the loss functions + corrector, Jacobi scaling, the LM strategy, Schur elimination of the
//     point blocks, the monotonic trust-region minimiser's rules, bound projection;
// anchored on the reference's own call sites (cited per function below).
//
// Plain C++17, fp64 only, no dependencies.  Optional OpenMP (compile with -fopenmp).

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/uba.h"

namespace {

// ------------------------------------------------------------------------------------
// Forward-mode dual number with 9 partials: d/d(cam[0..5]), d/d(point[0..2]).
// Plays the role of ceres::Jet<double,9> in AutoDiffCostFunction<F,M,6,3>.
// ------------------------------------------------------------------------------------
struct Dual {
  double a;
  double v[9];
  Dual() : a(0) { for (double& x : v) x = 0; }
  explicit Dual(double s) : a(s) { for (double& x : v) x = 0; }
  Dual(double s, int k) : a(s) { for (double& x : v) x = 0; v[k] = 1.0; }
};
inline Dual operator+(const Dual& f, const Dual& g) { Dual h; h.a = f.a + g.a; for (int i = 0; i < 9; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Dual operator-(const Dual& f, const Dual& g) { Dual h; h.a = f.a - g.a; for (int i = 0; i < 9; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Dual operator-(const Dual& f) { Dual h; h.a = -f.a; for (int i = 0; i < 9; i++) h.v[i] = -f.v[i]; return h; }
inline Dual operator*(const Dual& f, const Dual& g) { Dual h; h.a = f.a * g.a; for (int i = 0; i < 9; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Dual operator/(const Dual& f, const Dual& g) {
  Dual h; const double ginv = 1.0 / g.a; const double q = f.a * ginv; h.a = q;
  for (int i = 0; i < 9; i++) h.v[i] = (f.v[i] - q * g.v[i]) * ginv;
  return h;
}
inline Dual operator+(const Dual& f, double s) { Dual h = f; h.a += s; return h; }
inline Dual operator-(const Dual& f, double s) { Dual h = f; h.a -= s; return h; }
inline Dual operator*(double s, const Dual& f) { Dual h; h.a = s * f.a; for (int i = 0; i < 9; i++) h.v[i] = s * f.v[i]; return h; }
inline Dual operator*(const Dual& f, double s) { return s * f; }
inline Dual operator+(double s, const Dual& f) { return f + s; }
inline Dual operator-(double s, const Dual& f) { return (-f) + s; }
inline Dual sqrt(const Dual& f) { Dual h; h.a = std::sqrt(f.a); const double d = 1.0 / (2.0 * h.a); for (int i = 0; i < 9; i++) h.v[i] = f.v[i] * d; return h; }
inline Dual sin(const Dual& f) { Dual h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < 9; i++) h.v[i] = c * f.v[i]; return h; }
inline Dual cos(const Dual& f) { Dual h; h.a = std::cos(f.a); const double s = -std::sin(f.a); for (int i = 0; i < 9; i++) h.v[i] = s * f.v[i]; return h; }
inline double value_of(const Dual& f) { return f.a; }
inline double value_of(double f) { return f; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }

// [CERES-UPSTREAM] ceres::AngleAxisRotatePoint (ceres/rotation.h), called at
// BundleAdjuster.h:82,:117,:157.  Rodrigues' formula for theta^2 > DBL_EPSILON, the
// first-order expansion p + r x p otherwise (so derivatives stay finite at r = 0,
// which is where log_map_Quat puts the identity pose, rotation_utils.h:199-204).
template <typename T>
void rotate_by_angle_axis(const T r[3], const T p[3], T out[3]) {
  const T theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (value_of(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt(theta2);
    const T c = cos(theta);
    const T s = sin(theta);
    const T ti = T(1.0) / theta;
    const T w[3] = {r[0] * ti, r[1] * ti, r[2] * ti};
    const T wxp[3] = {w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]};
    const T tmp = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (T(1.0) - c);
    out[0] = p[0] * c + wxp[0] * s + w[0] * tmp;
    out[1] = p[1] * c + wxp[1] * s + w[1] * tmp;
    out[2] = p[2] * c + wxp[2] * s + w[2] * tmp;
  } else {
    const T rxp[3] = {r[1] * p[2] - r[2] * p[1], r[2] * p[0] - r[0] * p[2], r[0] * p[1] - r[1] * p[0]};
    out[0] = p[0] + rxp[0];
    out[1] = p[1] + rxp[1];
    out[2] = p[2] + rxp[2];
  }
}

// StereoReprojectionError::operator() — BundleAdjuster.h:153-171.  Rows 1 and 3 share
// the same predicted y (K[0]); the right x uses K[1](0,0), K[1](0,2) (:163).
template <typename T>
void stereo_rows(const uba_calib& k, double sigma_inv, const double o[4], const T cam[6], const T X[3], T r[4]) {
  T p[3];
  rotate_by_angle_axis(cam + 3, X, p);
  p[0] = p[0] + cam[0];
  p[1] = p[1] + cam[1];
  p[2] = p[2] + cam[2];
  const T x1 = k.fx0 * (p[0] / p[2]) + k.cx0;
  const T x2 = k.fx1 * ((p[0] - k.baseline) / p[2]) + k.cx1;
  const T y = k.fy0 * (p[1] / p[2]) + k.cy0;
  r[0] = sigma_inv * (x1 - o[0]);
  r[1] = sigma_inv * (y - o[1]);
  r[2] = sigma_inv * (x2 - o[2]);
  r[3] = sigma_inv * (y - o[3]);
}

// StandardReprojectionError::operator() (:78-94) when cam_id == 0, StereoRightError::operator()
// (:113-130) otherwise: the latter shifts p.x by -baseline (:119) and still projects with K[0]
// (:123-124).  Selection by camID: :399-402.
template <typename T>
void mono_rows(const uba_calib& k, double sigma_inv, const double o[2], int cam_id, const T cam[6], const T X[3], T r[2]) {
  T p[3];
  rotate_by_angle_axis(cam + 3, X, p);
  if (cam_id == 0) p[0] = p[0] + cam[0];
  else p[0] = p[0] + (cam[0] - k.baseline);
  p[1] = p[1] + cam[1];
  p[2] = p[2] + cam[2];
  const T x = k.fx0 * (p[0] / p[2]) + k.cx0;
  const T y = k.fy0 * (p[1] / p[2]) + k.cy0;
  r[0] = sigma_inv * (x - o[0]);
  r[1] = sigma_inv * (y - o[1]);
}

// [CERES-UPSTREAM] ceres::HuberLoss / ceres::CauchyLoss Evaluate(s, rho[3]); the reference
// builds HuberLoss(1.0) per observation (BundleAdjuster.h:397,:447).
void loss_eval(int kind, double a, double s, double rho[3]) {
  const double b = a * a;
  if (kind == UBA_LOSS_HUBER) {
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (kind == UBA_LOSS_CAUCHY) {
    const double c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Problem {
  int M = 4;
  int n_cams = 0, n_pts = 0;
  int64_t n_obs = 0;
  const double* feats = nullptr;  // [n_obs][M]
  const int32_t* cam_idx = nullptr;
  const int32_t* pt_idx = nullptr;
  const int32_t* cam_id = nullptr;  // may be null
  uba_calib calib{};
  double sigma_inv = 1.0;
};

struct Bounds { double lo[3], hi[3]; };

// Point box of BundleAdjuster.h:442-443 (:391-392) and :455-460 (:408-413).
Bounds point_bounds(const uba_calib& k) {
  const double zmax = k.fx0 * k.baseline / 0.1;
  const double zmin = k.fx0 * k.baseline / (2.0 * k.cx0);
  Bounds b;
  b.hi[0] = zmax / k.fx0 * k.cx0; b.lo[0] = -zmax / k.fx0 * k.cx0;
  b.hi[1] = zmax / k.fy0 * k.cy0; b.lo[1] = -zmax / k.fy0 * k.cy0;
  b.hi[2] = zmax; b.lo[2] = zmin;
  return b;
}

uba_calib effective_calib(const uba_calib& in, int M) {
  uba_calib k = in;
  if (M == 2 && k.baseline == 0.0) k.baseline = 0.5;  // BundleAdjuster.h:389-390
  return k;
}

// One residual block: residuals + (optionally) autodiff Jacobians, row-major.
void eval_block(const Problem& P, int64_t o, const double* cam, const double* X, double* r, double* Jc, double* Jp) {
  const int M = P.M;
  const double* f = P.feats + o * M;
  const int cid = P.cam_id ? P.cam_id[o] : 0;
  if (!Jc) {
    if (M == 4) stereo_rows<double>(P.calib, P.sigma_inv, f, cam, X, r);
    else mono_rows<double>(P.calib, P.sigma_inv, f, cid, cam, X, r);
    return;
  }
  Dual c[6], x[3], rr[4];
  for (int i = 0; i < 6; i++) c[i] = Dual(cam[i], i);
  for (int i = 0; i < 3; i++) x[i] = Dual(X[i], 6 + i);
  if (M == 4) stereo_rows<Dual>(P.calib, P.sigma_inv, f, c, x, rr);
  else mono_rows<Dual>(P.calib, P.sigma_inv, f, cid, c, x, rr);
  for (int m = 0; m < M; m++) {
    r[m] = rr[m].a;
    for (int i = 0; i < 6; i++) Jc[m * 6 + i] = rr[m].v[i];
    for (int i = 0; i < 3; i++) Jp[m * 3 + i] = rr[m].v[6 + i];
  }
}

// Dense Cholesky A = L L^T in place (lower); returns false on a non-positive pivot.
bool cholesky_lower(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    const double dinv = 1.0 / d;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n - j > 256)
#endif
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      const double* ai = A + (size_t)i * n;
      const double* aj = A + (size_t)j * n;
      for (int k = 0; k < j; k++) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s * dinv;
    }
  }
  return true;
}
void cholesky_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

// Everything one evaluation of the (robustified) linear least-squares model holds.
struct Linearization {
  std::vector<double> r_raw;    // [n_obs][M]
  std::vector<double> w;        // [n_obs] sqrt(rho')
  std::vector<double> r;        // [n_obs][M] corrected
  std::vector<double> Jc;       // [n_obs][M][6] corrected (unscaled)
  std::vector<double> Jp;       // [n_obs][M][3] corrected (unscaled)
  double cost = 0;
  std::vector<double> g_c;      // [n_cams][6]
  std::vector<double> g_p;      // [n_pts][3]
  std::vector<double> colsq_c;  // [n_cams][6] squared column norms of J~
  std::vector<double> colsq_p;  // [n_pts][3]
};

struct Structure {
  std::vector<int> free_cam;        // [n_cams] -1 or compact index
  std::vector<char> pt_active;      // [n_pts]
  int n_free = 0;
  std::vector<int64_t> pt_off;      // CSR over observation ids grouped by point (stable)
  std::vector<int64_t> pt_obs;      // observation ids, point-major, camera-ascending
};

Structure build_structure(const Problem& P, int fixed_frames) {
  Structure S;
  S.free_cam.assign(P.n_cams, -1);
  S.pt_active.assign(P.n_pts, 0);
  std::vector<char> seen(P.n_cams, 0);
  for (int64_t o = 0; o < P.n_obs; o++) { seen[P.cam_idx[o]] = 1; S.pt_active[P.pt_idx[o]] = 1; }
  // constant cameras: camIdx < fixedFrames (BundleAdjuster.h:406-407,:452-454); cameras that no
  // observation references never enter the ceres::Problem at all.
  for (int i = 0; i < P.n_cams; i++) if (i >= fixed_frames && seen[i]) S.free_cam[i] = S.n_free++;
  S.pt_off.assign(P.n_pts + 1, 0);
  for (int64_t o = 0; o < P.n_obs; o++) S.pt_off[P.pt_idx[o] + 1]++;
  for (int j = 0; j < P.n_pts; j++) S.pt_off[j + 1] += S.pt_off[j];
  std::vector<int64_t> ids(P.n_obs);
  std::iota(ids.begin(), ids.end(), (int64_t)0);
  std::stable_sort(ids.begin(), ids.end(), [&](int64_t a, int64_t b) {
    if (P.pt_idx[a] != P.pt_idx[b]) return P.pt_idx[a] < P.pt_idx[b];
    return P.cam_idx[a] < P.cam_idx[b];
  });
  S.pt_obs = ids;
  return S;
}

void evaluate(const Problem& P, const uba_config& cfg, const double* cams, const double* pts, bool with_jac,
              Linearization& L, double* cost_only) {
  const int M = P.M;
  if (with_jac) {
    L.r_raw.assign((size_t)P.n_obs * M, 0); L.w.assign(P.n_obs, 0); L.r.assign((size_t)P.n_obs * M, 0);
    L.Jc.assign((size_t)P.n_obs * M * 6, 0); L.Jp.assign((size_t)P.n_obs * M * 3, 0);
    L.g_c.assign((size_t)P.n_cams * 6, 0); L.g_p.assign((size_t)P.n_pts * 3, 0);
    L.colsq_c.assign((size_t)P.n_cams * 6, 0); L.colsq_p.assign((size_t)P.n_pts * 3, 0);
  }
  double cost = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : cost)
#endif
  for (int64_t o = 0; o < P.n_obs; o++) {
    double r[4], Jc[24], Jp[12];
    const double* cam = cams + (size_t)P.cam_idx[o] * 6;
    const double* X = pts + (size_t)P.pt_idx[o] * 3;
    eval_block(P, o, cam, X, r, with_jac ? Jc : nullptr, with_jac ? Jp : nullptr);
    double s = 0;
    for (int m = 0; m < M; m++) s += r[m] * r[m];
    double rho[3];
    loss_eval(cfg.loss_kind, cfg.loss_scale, s, rho);
    cost += 0.5 * rho[0];  // [CERES-UPSTREAM] ResidualBlock::Evaluate: cost = 0.5 * rho(s)
    if (with_jac) {
      // [CERES-UPSTREAM] Corrector: rho'' <= 0 (Huber, Cauchy) or s == 0 -> plain sqrt(rho') scaling.
      const double ws = std::sqrt(rho[1]);
      L.w[o] = ws;
      for (int m = 0; m < M; m++) {
        L.r_raw[(size_t)o * M + m] = r[m];
        L.r[(size_t)o * M + m] = ws * r[m];
        for (int i = 0; i < 6; i++) L.Jc[((size_t)o * M + m) * 6 + i] = ws * Jc[m * 6 + i];
        for (int i = 0; i < 3; i++) L.Jp[((size_t)o * M + m) * 3 + i] = ws * Jp[m * 3 + i];
      }
    }
  }
  if (cost_only) *cost_only = cost;
  if (!with_jac) return;
  L.cost = cost;
  for (int64_t o = 0; o < P.n_obs; o++) {
    const int ci = P.cam_idx[o], pj = P.pt_idx[o];
    for (int m = 0; m < M; m++) {
      const double rm = L.r[(size_t)o * M + m];
      const double* jc = &L.Jc[((size_t)o * M + m) * 6];
      const double* jp = &L.Jp[((size_t)o * M + m) * 3];
      for (int i = 0; i < 6; i++) { L.g_c[(size_t)ci * 6 + i] += jc[i] * rm; L.colsq_c[(size_t)ci * 6 + i] += jc[i] * jc[i]; }
      for (int i = 0; i < 3; i++) { L.g_p[(size_t)pj * 3 + i] += jp[i] * rm; L.colsq_p[(size_t)pj * 3 + i] += jp[i] * jp[i]; }
    }
  }
}

// [CERES-UPSTREAM] TrustRegionMinimizer::IterationZero -> EstimateScale: 1 / (1 + sqrt(||col||^2)).
void jacobi_scale_from(const Linearization& L, const uba_config& cfg, std::vector<double>& sc, std::vector<double>& sp) {
  sc.resize(L.colsq_c.size()); sp.resize(L.colsq_p.size());
  for (size_t i = 0; i < sc.size(); i++) sc[i] = cfg.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(L.colsq_c[i])) : 1.0;
  for (size_t i = 0; i < sp.size(); i++) sp[i] = cfg.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(L.colsq_p[i])) : 1.0;
}

struct StepResult {
  bool solver_ok = false;
  std::vector<double> d_c;   // [n_cams][6] step (unscaled space), zeros for constant cameras
  std::vector<double> d_p;   // [n_pts][3]
  double model_cost_change = 0;
  // diagnostics in UNSCALED space for parity against the GPU path
  std::vector<double> S, rhs;         // [n][n], [n]
  std::vector<double> B;              // [n_cams][36]
  std::vector<double> C;              // [n_pts][9]
  std::vector<double> W;              // [n_obs][18]
  std::vector<double> lam_c, lam_p;   // damping in unscaled space
};

// [CERES-UPSTREAM] LevenbergMarquardtStrategy::ComputeStep + SchurComplementSolver
// (linear_solver_type = SPARSE_SCHUR, BundleAdjuster.h:418,:465), worked in the
// Jacobi-scaled space exactly as Ceres does:
//   D^2 = clamp(||J_s col||^2, min_lm_diagonal, max_lm_diagonal) / radius
//   (J_s^T J_s + D^2) y = J_s^T r,  step_s = -y,  step = scale .* step_s
//   model_cost_change = -(J_s step_s) . (r + J_s step_s / 2)
// Point blocks are the e-blocks of the Schur elimination; constant cameras have been
// removed from the program but their residuals still shape the point blocks.
void compute_step(const Problem& P, const uba_config& cfg, const Structure& ST, const Linearization& L,
                  const std::vector<double>& sc, const std::vector<double>& sp, double radius, bool want_diag,
                  StepResult& R) {
  const int M = P.M;
  const int nf = ST.n_free, n = 6 * nf;
  R.d_c.assign((size_t)P.n_cams * 6, 0); R.d_p.assign((size_t)P.n_pts * 3, 0);
  const bool damped = radius > 0;
  // scaled LM diagonal
  std::vector<double> D2c((size_t)P.n_cams * 6, 0), D2p((size_t)P.n_pts * 3, 0);
  if (damped) {
    for (size_t i = 0; i < D2c.size(); i++) {
      const double d = std::min(std::max(L.colsq_c[i] * sc[i] * sc[i], cfg.min_lm_diagonal), cfg.max_lm_diagonal);
      D2c[i] = d / radius;
    }
    for (size_t i = 0; i < D2p.size(); i++) {
      const double d = std::min(std::max(L.colsq_p[i] * sp[i] * sp[i], cfg.min_lm_diagonal), cfg.max_lm_diagonal);
      D2p[i] = d / radius;
    }
  }
  std::vector<double> Sm((size_t)n * n, 0), rhs(n, 0);
  std::vector<double> Cinv((size_t)P.n_pts * 9, 0);
  if (want_diag) {
    R.B.assign((size_t)P.n_cams * 36, 0); R.C.assign((size_t)P.n_pts * 9, 0); R.W.assign((size_t)P.n_obs * 18, 0);
    R.lam_c.assign((size_t)P.n_cams * 6, 0); R.lam_p.assign((size_t)P.n_pts * 3, 0);
    for (size_t i = 0; i < R.lam_c.size(); i++) R.lam_c[i] = D2c[i] / (sc[i] * sc[i]);
    for (size_t i = 0; i < R.lam_p.size(); i++) R.lam_p[i] = D2p[i] / (sp[i] * sp[i]);
  }
  // camera diagonal blocks and gradient (scaled)
  for (int64_t o = 0; o < P.n_obs; o++) {
    const int ci = P.cam_idx[o];
    const int fi = ST.free_cam[ci];
    if (want_diag && ci >= 0) {
      for (int m = 0; m < M; m++) {
        const double* jc = &L.Jc[((size_t)o * M + m) * 6];
        if (fi >= 0) for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) R.B[(size_t)ci * 36 + a * 6 + b] += jc[a] * jc[b];
      }
    }
    if (fi < 0) continue;
    const double* s = &sc[(size_t)ci * 6];
    for (int m = 0; m < M; m++) {
      const double* jc = &L.Jc[((size_t)o * M + m) * 6];
      const double rm = L.r[(size_t)o * M + m];
      for (int a = 0; a < 6; a++) {
        rhs[fi * 6 + a] += s[a] * jc[a] * rm;
        for (int b = 0; b < 6; b++) Sm[(size_t)(fi * 6 + a) * n + fi * 6 + b] += s[a] * jc[a] * s[b] * jc[b];
      }
    }
  }
  for (int i = 0; i < P.n_cams; i++) {
    const int fi = ST.free_cam[i];
    if (fi < 0) continue;
    for (int a = 0; a < 6; a++) Sm[(size_t)(fi * 6 + a) * n + fi * 6 + a] += D2c[(size_t)i * 6 + a];
  }
  // per point: eliminate
  const int nthreads =
#ifdef _OPENMP
      omp_get_max_threads();
#else
      1;
#endif
  std::vector<std::vector<double>> Sloc(nthreads > 1 ? nthreads : 0), rloc(nthreads > 1 ? nthreads : 0);
  bool all_ok = true;
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double* Sacc = Sm.data(); double* racc = rhs.data();
    if (nthreads > 1) { Sloc[tid].assign((size_t)n * n, 0); rloc[tid].assign(n, 0); Sacc = Sloc[tid].data(); racc = rloc[tid].data(); }
    std::vector<double> Wj; std::vector<int> Wf;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (int j = 0; j < P.n_pts; j++) {
      const int64_t b0 = ST.pt_off[j], b1 = ST.pt_off[j + 1];
      if (b0 == b1) continue;
      const double* s3 = &sp[(size_t)j * 3];
      double C[9] = {0}, g[3] = {0}, Cu[9] = {0};
      for (int64_t q = b0; q < b1; q++) {
        const int64_t o = ST.pt_obs[q];
        for (int m = 0; m < M; m++) {
          const double* jp = &L.Jp[((size_t)o * M + m) * 3];
          const double rm = L.r[(size_t)o * M + m];
          for (int a = 0; a < 3; a++) {
            g[a] += s3[a] * jp[a] * rm;
            for (int b = 0; b < 3; b++) { C[a * 3 + b] += s3[a] * jp[a] * s3[b] * jp[b]; Cu[a * 3 + b] += jp[a] * jp[b]; }
          }
        }
      }
      if (want_diag) for (int a = 0; a < 9; a++) R.C[(size_t)j * 9 + a] = Cu[a];
      for (int a = 0; a < 3; a++) C[a * 3 + a] += D2p[(size_t)j * 3 + a];
      // inverse of the 3x3 SPD block through its Cholesky factor
      double Lc[9]; std::memcpy(Lc, C, sizeof(Lc));
      if (!cholesky_lower(Lc, 3)) {
#ifdef _OPENMP
#pragma omp critical
#endif
        all_ok = false;
        continue;
      }
      double Ci[9];
      for (int col = 0; col < 3; col++) {
        double e[3] = {0, 0, 0}; e[col] = 1.0;
        cholesky_solve(Lc, 3, e);
        for (int a = 0; a < 3; a++) Ci[a * 3 + col] = e[a];
      }
      std::memcpy(&Cinv[(size_t)j * 9], Ci, sizeof(Ci));
      // W_ij = F^T E (scaled), for free cameras
      const int k = (int)(b1 - b0);
      Wj.assign((size_t)k * 18, 0); Wf.assign(k, -1);
      for (int q = 0; q < k; q++) {
        const int64_t o = ST.pt_obs[b0 + q];
        const int ci = P.cam_idx[o];
        const int fi = ST.free_cam[ci];
        Wf[q] = fi;
        if (fi < 0) continue;
        const double* s6 = &sc[(size_t)ci * 6];
        double* Wq = &Wj[(size_t)q * 18];
        for (int m = 0; m < M; m++) {
          const double* jc = &L.Jc[((size_t)o * M + m) * 6];
          const double* jp = &L.Jp[((size_t)o * M + m) * 3];
          for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) {
            Wq[a * 3 + b] += s6[a] * jc[a] * s3[b] * jp[b];
            if (want_diag) R.W[(size_t)o * 18 + a * 3 + b] += jc[a] * jp[b];
          }
        }
      }
      // S -= W C^-1 W^T, rhs -= W C^-1 g
      double Cig[3];
      for (int a = 0; a < 3; a++) Cig[a] = Ci[a * 3 + 0] * g[0] + Ci[a * 3 + 1] * g[1] + Ci[a * 3 + 2] * g[2];
      for (int q = 0; q < k; q++) {
        if (Wf[q] < 0) continue;
        const double* Wq = &Wj[(size_t)q * 18];
        double Y[18];
        for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++)
          Y[a * 3 + b] = Wq[a * 3 + 0] * Ci[0 * 3 + b] + Wq[a * 3 + 1] * Ci[1 * 3 + b] + Wq[a * 3 + 2] * Ci[2 * 3 + b];
        for (int a = 0; a < 6; a++) racc[Wf[q] * 6 + a] -= Wq[a * 3 + 0] * Cig[0] + Wq[a * 3 + 1] * Cig[1] + Wq[a * 3 + 2] * Cig[2];
        for (int q2 = 0; q2 < k; q2++) {
          if (Wf[q2] < 0) continue;
          const double* W2 = &Wj[(size_t)q2 * 18];
          for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
            Sacc[(size_t)(Wf[q] * 6 + a) * n + Wf[q2] * 6 + b] -= Y[a * 3 + 0] * W2[b * 3 + 0] + Y[a * 3 + 1] * W2[b * 3 + 1] + Y[a * 3 + 2] * W2[b * 3 + 2];
        }
      }
    }
  }
  for (int t = 0; t < (int)Sloc.size(); t++) {
    for (size_t i = 0; i < Sm.size(); i++) Sm[i] += Sloc[t][i];
    for (int i = 0; i < n; i++) rhs[i] += rloc[t][i];
  }
  if (want_diag) {
    // report in unscaled space: S_u = Sc^-1 S_s Sc^-1, rhs_u = Sc^-1 rhs_s
    std::vector<double> sfree(n, 1.0);
    for (int i = 0; i < P.n_cams; i++) if (ST.free_cam[i] >= 0) for (int a = 0; a < 6; a++) sfree[ST.free_cam[i] * 6 + a] = sc[(size_t)i * 6 + a];
    R.S.assign((size_t)n * n, 0); R.rhs.assign(n, 0);
    for (int a = 0; a < n; a++) { R.rhs[a] = rhs[a] / sfree[a]; for (int b = 0; b < n; b++) R.S[(size_t)a * n + b] = Sm[(size_t)a * n + b] / (sfree[a] * sfree[b]); }
  }
  if (!all_ok) { R.solver_ok = false; return; }
  // reduced solve
  std::vector<double> y(rhs);
  if (n > 0) {
    if (!cholesky_lower(Sm.data(), n)) { R.solver_ok = false; return; }
    cholesky_solve(Sm.data(), n, y.data());
  }
  // back-substitute: y_p = C^-1 (g - sum_i W_ij^T y_i), all in scaled space
  std::vector<double> yp((size_t)P.n_pts * 3, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int j = 0; j < P.n_pts; j++) {
    const int64_t b0 = ST.pt_off[j], b1 = ST.pt_off[j + 1];
    if (b0 == b1) continue;
    const double* s3 = &sp[(size_t)j * 3];
    double t[3] = {0, 0, 0};
    for (int64_t q = b0; q < b1; q++) {
      const int64_t o = ST.pt_obs[q];
      const int ci = P.cam_idx[o];
      const int fi = ST.free_cam[ci];
      const double* s6 = &sc[(size_t)ci * 6];
      for (int m = 0; m < M; m++) {
        const double* jc = &L.Jc[((size_t)o * M + m) * 6];
        const double* jp = &L.Jp[((size_t)o * M + m) * 3];
        double fy = 0;
        if (fi >= 0) for (int a = 0; a < 6; a++) fy += s6[a] * jc[a] * y[fi * 6 + a];
        const double rm = L.r[(size_t)o * M + m];
        for (int b = 0; b < 3; b++) t[b] += s3[b] * jp[b] * (rm - fy);
      }
    }
    const double* Ci = &Cinv[(size_t)j * 9];
    for (int a = 0; a < 3; a++) yp[(size_t)j * 3 + a] = Ci[a * 3 + 0] * t[0] + Ci[a * 3 + 1] * t[1] + Ci[a * 3 + 2] * t[2];
  }
  // model cost change, evaluated on the residual rows like Ceres does
  double mcc = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : mcc)
#endif
  for (int64_t o = 0; o < P.n_obs; o++) {
    const int ci = P.cam_idx[o], pj = P.pt_idx[o];
    const int fi = ST.free_cam[ci];
    const double* s6 = &sc[(size_t)ci * 6];
    const double* s3 = &sp[(size_t)pj * 3];
    for (int m = 0; m < M; m++) {
      const double* jc = &L.Jc[((size_t)o * M + m) * 6];
      const double* jp = &L.Jp[((size_t)o * M + m) * 3];
      double jd = 0;  // (J_s * step_s) row, step_s = -y
      if (fi >= 0) for (int a = 0; a < 6; a++) jd -= s6[a] * jc[a] * y[fi * 6 + a];
      for (int b = 0; b < 3; b++) jd -= s3[b] * jp[b] * yp[(size_t)pj * 3 + b];
      mcc -= jd * (L.r[(size_t)o * M + m] + 0.5 * jd);
    }
  }
  R.model_cost_change = mcc;
  for (int i = 0; i < P.n_cams; i++) {
    const int fi = ST.free_cam[i];
    if (fi < 0) continue;
    for (int a = 0; a < 6; a++) R.d_c[(size_t)i * 6 + a] = -y[fi * 6 + a] * sc[(size_t)i * 6 + a];
  }
  for (int j = 0; j < P.n_pts; j++) for (int a = 0; a < 3; a++) R.d_p[(size_t)j * 3 + a] = -yp[(size_t)j * 3 + a] * sp[(size_t)j * 3 + a];
  R.solver_ok = true;
}

double clampd(double x, double lo, double hi) { return std::min(std::max(x, lo), hi); }

// [CERES-UPSTREAM] bounds-projected gradient max-norm: || x - Pi(x - g) ||_inf over the reduced program.
double gradient_max_norm(const Problem& P, const uba_config& cfg, const Structure& ST, const Linearization& L,
                         const double* pts, const Bounds& bd) {
  double gmax = 0;
  for (int i = 0; i < P.n_cams; i++) if (ST.free_cam[i] >= 0) for (int a = 0; a < 6; a++) gmax = std::max(gmax, std::fabs(L.g_c[(size_t)i * 6 + a]));
  for (int j = 0; j < P.n_pts; j++) {
    if (!ST.pt_active[j]) continue;
    for (int a = 0; a < 3; a++) {
      const double x = pts[(size_t)j * 3 + a], g = L.g_p[(size_t)j * 3 + a];
      const double proj = cfg.use_bounds ? clampd(x - g, bd.lo[a], bd.hi[a]) : x - g;
      gmax = std::max(gmax, std::fabs(x - proj));
    }
  }
  return gmax;
}

void fill_problem(Problem& P, int M, int n_cams, int n_pts, int64_t n_obs, const double* feats, const int32_t* cam_idx,
                  const int32_t* pt_idx, const int32_t* cam_id, const uba_calib* calib) {
  P.M = M; P.n_cams = n_cams; P.n_pts = n_pts; P.n_obs = n_obs; P.feats = feats; P.cam_idx = cam_idx; P.pt_idx = pt_idx;
  P.cam_id = cam_id; P.calib = effective_calib(*calib, M);
  P.sigma_inv = 1.0 / std::sqrt(P.calib.feat_var);  // sigma = sqrt(feat_var), BundleAdjuster.h:400,:448
}

bool valid_problem(int M, int n_cams, int n_pts, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx) {
  if (M != 2 && M != 4) return false;
  if (n_cams < 0 || n_pts < 0 || n_obs < 0) return false;
  for (int64_t o = 0; o < n_obs; o++) if (cam_idx[o] < 0 || cam_idx[o] >= n_cams || pt_idx[o] < 0 || pt_idx[o] >= n_pts) return false;
  return true;
}

}  // namespace

extern "C" {

int uba_ref_version(void) { return UBA_VERSION; }

int uba_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void uba_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// One residual block with autodiff Jacobians (unit-test entry).  Jc [M][6], Jp [M][3] may be NULL.
int uba_ref_residual(int M, const uba_calib* calib, const double* cam6, const double* pt3, const double* obs, int cam_id,
                     double* r, double* Jc, double* Jp) {
  if ((M != 2 && M != 4) || !calib || !cam6 || !pt3 || !obs || !r) return UBA_ERR_INVALID_ARGUMENT;
  Problem P; int32_t ci = 0, pi = 0, cid = cam_id;
  fill_problem(P, M, 1, 1, 1, obs, &ci, &pi, &cid, calib);
  double jc[24], jp[12];
  eval_block(P, 0, cam6, pt3, r, (Jc || Jp) ? jc : nullptr, (Jc || Jp) ? jp : nullptr);
  if (Jc) std::memcpy(Jc, jc, sizeof(double) * M * 6);
  if (Jp) std::memcpy(Jp, jp, sizeof(double) * M * 3);
  return UBA_OK;
}

void uba_ref_rotate(const double r[3], const double p[3], double out[3]) { rotate_by_angle_axis<double>(r, p, out); }

void uba_ref_loss(int kind, double a, double s, double rho[3]) { loss_eval(kind, a, s, rho); }

void uba_ref_point_bounds(const uba_calib* calib, int M, double lo[3], double hi[3]) {
  const Bounds b = point_bounds(effective_calib(*calib, M));
  for (int a = 0; a < 3; a++) { lo[a] = b.lo[a]; hi[a] = b.hi[a]; }
}

// Pose packing / unpacking at the boundary: log_map_Quat (rotation_utils.h:199-204) used by
// initialiseParameters (BundleAdjuster.h:306-309) and exp_map_Quat (rotation_utils.h:190-197)
// used by getCameraPoses (:232-237).  acos is clamped to [-1,1] (the reference does not; a w a
// hair above 1 after normalisation gives NaN there).
void uba_ref_log_map_quat(const double q[4], double r[3]) {
  const double norm = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double theta = norm < 1e-10 ? 1e-10 : norm;
  const double w = std::min(1.0, std::max(-1.0, q[0]));
  const double f = std::acos(w) * 2.0 / theta;
  r[0] = f * q[1]; r[1] = f * q[2]; r[2] = f * q[3];
}
void uba_ref_exp_map_quat(const double r[3], double q[4]) {
  const double norm = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double theta = norm < 1e-10 ? 1e-10 : norm;
  const double s = std::sin(theta / 2) / theta;
  q[0] = std::cos(theta / 2); q[1] = r[0] * s; q[2] = r[1] * s; q[3] = r[2] * s;
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // Quat ctor normalises
  for (int i = 0; i < 4; i++) q[i] /= n;
}

// Ordering / index tables (SURVEY.md §8(a) A3 + derived tables).  All outputs optional.
//  obs_order : observation ids point-major, camera-ascending, stable (identity for inputs produced by
//              initialiseObservations, BundleAdjuster.h:364-374)
//  pt_obs_off: CSR offsets, caller point order
//  pt_order  : points stably sorted by (lowest observed camera, highest observed camera); points
//              without observations last
//  free_cam  : -1 for camIdx < fixed_frames (:452-454) or unobserved cameras, else compact index
int uba_ref_tables(int n_cams, int n_pts, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx, int fixed_frames,
                   int32_t* obs_order, int64_t* pt_obs_off, int32_t* pt_order, int32_t* free_cam) {
  if (!valid_problem(4, n_cams, n_pts, n_obs, cam_idx, pt_idx)) return UBA_ERR_INVALID_ARGUMENT;
  Problem P; P.n_cams = n_cams; P.n_pts = n_pts; P.n_obs = n_obs; P.cam_idx = cam_idx; P.pt_idx = pt_idx;
  Structure S = build_structure(P, fixed_frames);
  if (obs_order) for (int64_t i = 0; i < n_obs; i++) obs_order[i] = (int32_t)S.pt_obs[i];
  if (pt_obs_off) for (int j = 0; j <= n_pts; j++) pt_obs_off[j] = S.pt_off[j];
  if (free_cam) for (int i = 0; i < n_cams; i++) free_cam[i] = S.free_cam[i];
  if (pt_order) {
    std::vector<int> lo(n_pts, INT32_MAX), hi(n_pts, -1);
    for (int64_t o = 0; o < n_obs; o++) { lo[pt_idx[o]] = std::min(lo[pt_idx[o]], (int)cam_idx[o]); hi[pt_idx[o]] = std::max(hi[pt_idx[o]], (int)cam_idx[o]); }
    std::vector<int32_t> ids(n_pts);
    std::iota(ids.begin(), ids.end(), 0);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { if (lo[a] != lo[b]) return lo[a] < lo[b]; return hi[a] < hi[b]; });
    for (int j = 0; j < n_pts; j++) pt_order[j] = ids[j];
  }
  return UBA_OK;
}

// One linearisation + Schur elimination at (cams6, pts3).  jacobi_scale: [6*n_cams + 3*n_pts] or
// NULL (then computed at this very point, i.e. this is "iteration 0").  Arrays in caller order.
int uba_ref_linearize(int M, int n_cams, int n_pts, int64_t n_obs, const double* cams6, const double* pts3,
                      const double* feats, const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                      const uba_calib* calib, const uba_config* cfg, int fixed_frames, double radius,
                      const double* jacobi_scale, uba_linearization_out* out, double* jacobi_scale_out,
                      double* step_cams, double* step_pts, double* model_cost_change) {
  if (!calib || !cfg || !out || !valid_problem(M, n_cams, n_pts, n_obs, cam_idx, pt_idx)) return UBA_ERR_INVALID_ARGUMENT;
  Problem P; fill_problem(P, M, n_cams, n_pts, n_obs, feats, cam_idx, pt_idx, cam_id, calib);
  Structure ST = build_structure(P, fixed_frames);
  Linearization L;
  evaluate(P, *cfg, cams6, pts3, true, L, nullptr);
  std::vector<double> sc, sp;
  if (jacobi_scale) { sc.assign(jacobi_scale, jacobi_scale + (size_t)6 * n_cams); sp.assign(jacobi_scale + (size_t)6 * n_cams, jacobi_scale + (size_t)6 * n_cams + (size_t)3 * n_pts); }
  else jacobi_scale_from(L, *cfg, sc, sp);
  if (jacobi_scale_out) { std::copy(sc.begin(), sc.end(), jacobi_scale_out); std::copy(sp.begin(), sp.end(), jacobi_scale_out + sc.size()); }
  StepResult R;
  compute_step(P, *cfg, ST, L, sc, sp, radius, true, R);
  const int n = 6 * ST.n_free;
  if (out->residuals) std::copy(L.r_raw.begin(), L.r_raw.end(), out->residuals);
  if (out->weights) std::copy(L.w.begin(), L.w.end(), out->weights);
  if (out->cost) out->cost[0] = L.cost;
  if (out->grad_cams) for (int i = 0; i < n_cams; i++) for (int a = 0; a < 6; a++) out->grad_cams[(size_t)i * 6 + a] = ST.free_cam[i] >= 0 ? L.g_c[(size_t)i * 6 + a] : 0.0;
  if (out->grad_pts) std::copy(L.g_p.begin(), L.g_p.end(), out->grad_pts);
  if (out->B) std::copy(R.B.begin(), R.B.end(), out->B);
  if (out->C) std::copy(R.C.begin(), R.C.end(), out->C);
  if (out->W) std::copy(R.W.begin(), R.W.end(), out->W);
  if (out->S) std::copy(R.S.begin(), R.S.end(), out->S);
  if (out->rhs) std::copy(R.rhs.begin(), R.rhs.end(), out->rhs);
  if (out->lm_diag_cams) for (int i = 0; i < n_cams; i++) for (int a = 0; a < 6; a++) out->lm_diag_cams[(size_t)i * 6 + a] = ST.free_cam[i] >= 0 ? R.lam_c[(size_t)i * 6 + a] : 0.0;
  if (out->lm_diag_pts) for (int j = 0; j < n_pts; j++) for (int a = 0; a < 3; a++) out->lm_diag_pts[(size_t)j * 3 + a] = ST.pt_active[j] ? R.lam_p[(size_t)j * 3 + a] : 0.0;
  (void)n;
  if (step_cams) std::copy(R.d_c.begin(), R.d_c.end(), step_cams);
  if (step_pts) std::copy(R.d_p.begin(), R.d_p.end(), step_pts);
  if (model_cost_change) *model_cost_change = R.model_cost_change;
  return R.solver_ok ? UBA_OK : UBA_ERR_NUMERICAL;
}

// Cost only: 0.5 * sum rho(||r||^2).
int uba_ref_cost(int M, int n_cams, int n_pts, int64_t n_obs, const double* cams6, const double* pts3, const double* feats,
                 const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id, const uba_calib* calib,
                 const uba_config* cfg, double* cost) {
  if (!calib || !cfg || !cost || !valid_problem(M, n_cams, n_pts, n_obs, cam_idx, pt_idx)) return UBA_ERR_INVALID_ARGUMENT;
  Problem P; fill_problem(P, M, n_cams, n_pts, n_obs, feats, cam_idx, pt_idx, cam_id, calib);
  Linearization L;
  evaluate(P, *cfg, cams6, pts3, false, L, cost);
  return UBA_OK;
}

// BundleAdjuster<M>::optimise (BundleAdjuster.h:378-476) with ceres::Solve restated:
// [CERES-UPSTREAM] TrustRegionMinimizer (monotonic steps, LEVENBERG_MARQUARDT strategy),
// options as set at :416-420 / :463-467, everything else the Ceres defaults mirrored in
// uba_config_default.  cams6 / pts3 are optimised in place (as Ceres does through the raw
// double* of :404,:449); they are restored when the solution is not usable.
// Deviations, stated: (1) the wall-clock cap (:417,:464) is honoured only when
// cfg->max_solver_time_s > 0; (2) the bounds-induced Armijo line search of the constrained
// trust-region loop is NOT run; *armijo_violations counts the steps on which it would have
// shortened the step (f(x+d) > f(x) + 1e-4 g.d); (3) fixed_iterations > 0 disables all
// convergence tests.
int uba_ref_optimise(int M, int n_cams, int n_pts, int64_t n_obs, double* cams6, double* pts3, const double* feats,
                     const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id, const uba_calib* calib,
                     const uba_config* cfg_in, int fixed_frames, uba_summary* summary, uba_iteration* iters, int max_records,
                     int* n_records, int* armijo_violations) {
  if (!calib || !cfg_in || !cams6 || !pts3 || !valid_problem(M, n_cams, n_pts, n_obs, cam_idx, pt_idx)) return UBA_ERR_INVALID_ARGUMENT;
  const uba_config cfg = *cfg_in;
  Problem P; fill_problem(P, M, n_cams, n_pts, n_obs, feats, cam_idx, pt_idx, cam_id, calib);
  Structure ST = build_structure(P, fixed_frames);
  const Bounds bd = point_bounds(P.calib);
  uba_summary sum; std::memset(&sum, 0, sizeof(sum));
  int nrec = 0, armijo = 0;
  auto record = [&](const uba_iteration& it) { if (iters && nrec < max_records) iters[nrec] = it; nrec++; };
  auto finish = [&](int term, double cost, double radius, double gmax) {
    sum.termination = term; sum.usable = term != UBA_TERM_FAILURE; sum.final_cost = cost; sum.final_radius = radius;
    sum.final_gradient_max_norm = gmax;
    if (summary) *summary = sum;
    if (n_records) *n_records = nrec;
    if (armijo_violations) *armijo_violations = armijo;
    return sum.usable ? UBA_OK : UBA_ERR_NUMERICAL;
  };
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<double> cams0(cams6, cams6 + (size_t)6 * n_cams), pts0(pts3, pts3 + (size_t)3 * n_pts);
  // [CERES-UPSTREAM] Program::IsFeasible: a bounded, non-constant block outside its box -> FAILURE.
  if (cfg.use_bounds) {
    for (int j = 0; j < n_pts; j++) if (ST.pt_active[j]) for (int a = 0; a < 3; a++) {
      const double x = pts3[(size_t)j * 3 + a];
      if (!(x >= bd.lo[a] && x <= bd.hi[a])) { finish(UBA_TERM_FAILURE, 0, 0, 0); return UBA_ERR_INFEASIBLE; }
    }
  }
  std::vector<double> x_c(cams0), x_p(pts0), cand_c(cams0), cand_p(pts0);
  Linearization L;
  evaluate(P, cfg, x_c.data(), x_p.data(), true, L, nullptr);
  if (!std::isfinite(L.cost)) return finish(UBA_TERM_FAILURE, L.cost, 0, 0);
  double cost = L.cost;
  sum.initial_cost = cost;
  std::vector<double> sc, sp;
  jacobi_scale_from(L, cfg, sc, sp);
  double radius = cfg.initial_radius, decrease_factor = 2.0;
  double gmax = gradient_max_norm(P, cfg, ST, L, x_p.data(), bd);
  { uba_iteration it; std::memset(&it, 0, sizeof(it)); it.cost = cost; it.candidate_cost = cost; it.radius = radius; it.gradient_max_norm = gmax; it.accepted = 1; record(it); }
  const bool fixedK = cfg.fixed_iterations > 0;
  const int max_it = fixedK ? cfg.fixed_iterations : cfg.max_iterations;
  if (!fixedK && gmax <= cfg.gradient_tolerance) return finish(UBA_TERM_CONVERGENCE_GRADIENT, cost, radius, gmax);
  int consecutive_invalid = 0;
  int term = UBA_TERM_NO_CONVERGENCE;
  for (int iter = 1; iter <= max_it; iter++) {
    if (!fixedK && cfg.max_solver_time_s > 0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el >= cfg.max_solver_time_s) break;
    }
    sum.iterations = iter;
    uba_iteration it; std::memset(&it, 0, sizeof(it)); it.radius = radius; it.gradient_max_norm = gmax; it.cost = cost;
    StepResult R;
    compute_step(P, cfg, ST, L, sc, sp, radius, false, R);
    if (!R.solver_ok || !(R.model_cost_change > 0.0)) {
      // invalid step: [CERES-UPSTREAM] HandleInvalidStep / LevenbergMarquardtStrategy::StepIsInvalid
      it.accepted = -1; it.model_cost_change = R.model_cost_change; it.candidate_cost = cost; record(it);
      sum.invalid_steps++;
      if (++consecutive_invalid >= cfg.max_consecutive_invalid_steps && !fixedK) { term = UBA_TERM_FAILURE; break; }
      radius *= 0.5;
      continue;
    }
    consecutive_invalid = 0;
    // candidate = Plus(x, delta), then projection onto the point box
    double step2 = 0, x2 = 0;
    for (int i = 0; i < n_cams; i++) for (int a = 0; a < 6; a++) {
      const size_t q = (size_t)i * 6 + a;
      cand_c[q] = x_c[q] + R.d_c[q];
      if (ST.free_cam[i] >= 0) { step2 += (x_c[q] - cand_c[q]) * (x_c[q] - cand_c[q]); x2 += x_c[q] * x_c[q]; }
    }
    for (int j = 0; j < n_pts; j++) for (int a = 0; a < 3; a++) {
      const size_t q = (size_t)j * 3 + a;
      double v = x_p[q] + R.d_p[q];
      if (cfg.use_bounds && ST.pt_active[j]) v = clampd(v, bd.lo[a], bd.hi[a]);
      cand_p[q] = v;
      if (ST.pt_active[j]) { step2 += (x_p[q] - v) * (x_p[q] - v); x2 += x_p[q] * x_p[q]; }
    }
    double cand_cost = 0;
    evaluate(P, cfg, cand_c.data(), cand_p.data(), false, L, &cand_cost);
    {  // Armijo diagnostic (see header comment)
      double gd = 0;
      for (size_t q = 0; q < R.d_c.size(); q++) gd += L.g_c[q] * R.d_c[q];
      for (size_t q = 0; q < R.d_p.size(); q++) gd += L.g_p[q] * R.d_p[q];
      if (cand_cost > cost + 1e-4 * gd) armijo++;
    }
    it.candidate_cost = cand_cost; it.model_cost_change = R.model_cost_change; it.step_norm = std::sqrt(step2);
    const double cost_change = cost - cand_cost;
    it.relative_decrease = cost_change / R.model_cost_change;
    if (!fixedK) {
      if (it.step_norm <= cfg.parameter_tolerance * (std::sqrt(x2) + cfg.parameter_tolerance)) { it.accepted = 0; record(it); term = UBA_TERM_CONVERGENCE_PARAMETER; break; }
      if (std::fabs(cost_change) <= cfg.function_tolerance * cost) { it.accepted = 0; record(it); term = UBA_TERM_CONVERGENCE_FUNCTION; break; }
    }
    if (std::isfinite(cand_cost) && it.relative_decrease > cfg.min_relative_decrease) {
      x_c = cand_c; x_p = cand_p; cost = cand_cost;
      evaluate(P, cfg, x_c.data(), x_p.data(), true, L, nullptr);
      cost = L.cost;
      gmax = gradient_max_norm(P, cfg, ST, L, x_p.data(), bd);
      it.accepted = 1; it.cost = cost; it.gradient_max_norm = gmax; record(it);
      sum.successful_steps++;
      // LevenbergMarquardtStrategy::StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(cfg.max_radius, radius);
      decrease_factor = 2.0;
      if (!fixedK && gmax <= cfg.gradient_tolerance) { term = UBA_TERM_CONVERGENCE_GRADIENT; break; }
    } else {
      it.accepted = 0; record(it);
      sum.unsuccessful_steps++;
      // LevenbergMarquardtStrategy::StepRejected
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      if (!fixedK && radius < cfg.min_radius) { term = UBA_TERM_CONVERGENCE_RADIUS; break; }
    }
  }
  if (term != UBA_TERM_FAILURE) {
    std::copy(x_c.begin(), x_c.end(), cams6);
    std::copy(x_p.begin(), x_p.end(), pts3);
  } else {
    std::copy(cams0.begin(), cams0.end(), cams6);
    std::copy(pts0.begin(), pts0.end(), pts3);
  }
  return finish(term, cost, radius, gmax);
}

// Times `repeats` linearise+Schur+solve passes (the per-LM-iteration work) and returns seconds per pass.
double uba_ref_time_iteration(int M, int n_cams, int n_pts, int64_t n_obs, const double* cams6, const double* pts3,
                              const double* feats, const int32_t* cam_idx, const int32_t* pt_idx, const int32_t* cam_id,
                              const uba_calib* calib, const uba_config* cfg, int fixed_frames, int repeats, double* linearize_s) {
  Problem P; fill_problem(P, M, n_cams, n_pts, n_obs, feats, cam_idx, pt_idx, cam_id, calib);
  Structure ST = build_structure(P, fixed_frames);
  Linearization L;
  std::vector<double> sc, sp;
  double tl = 0, tt = 0;
  for (int r = 0; r < repeats; r++) {
    const auto a = std::chrono::steady_clock::now();
    evaluate(P, *cfg, cams6, pts3, true, L, nullptr);
    const auto b = std::chrono::steady_clock::now();
    if (r == 0) jacobi_scale_from(L, *cfg, sc, sp);
    StepResult R;
    compute_step(P, *cfg, ST, L, sc, sp, cfg->initial_radius, false, R);
    double c2 = 0;
    evaluate(P, *cfg, cams6, pts3, false, L, &c2);
    const auto c = std::chrono::steady_clock::now();
    tl += std::chrono::duration<double>(b - a).count();
    tt += std::chrono::duration<double>(c - a).count();
  }
  if (linearize_s) *linearize_s = tl / repeats;
  return tt / repeats;
}

}  // extern "C"
