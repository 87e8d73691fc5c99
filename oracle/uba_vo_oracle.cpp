// uba_vo_oracle.cpp — CPU ORACLE of the pose-only (stereo visual odometry) path.  TEST INFRASTRUCTURE, like
// uba_oracle.cpp: only tests/, smoke() and bench.py's CPU legs may load it.
//
// Restates, function by function, src/vo/StereoVisualOdometry.cpp of the reference (plain loops over small fixed-size
// matrices instead of cv::Matx / cv::Mat):
//   project3D :22-32, updateObservations :285-289, reproject :116-143, updateJacobian :291-329, optimize :165-283,
//   computeInliers :94-114, the RANSAC loop of process :59-75 (triples supplied by the caller instead of rand()).
// PINNED by oracle/_ref (the reference's own file compiled in this container): project3D, reproject, the residuals, the
// Jacobian, A = J J^T and B = J r agree with it to rounding (tests/test_vo.py), and so does computeInliers.
// ONE DELIBERATE DIFFERENCE: the reference's do/while condition (:277) compares the iteration counter with the
// StopCondition enum value; as written optimize() does not return on ordinary data (the compiled reference hangs).
// Here — and in libuba — the loop ends when a stop condition is set or after max_iter iterations.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/uba.h"

namespace {

enum Stop { NO_STOP = 0, SMALL_GRADIENT, SMALL_INCREMENT, MAX_ITERATIONS, SMALL_DECREASE_FUNCTION, SMALL_REPROJ_ERROR, NO_CONVERGENCE };  // rotation_utils.h:19

struct Vo {
  uba_vo_params P;
  int n;
  std::vector<double> pts;   // [n][4]
  std::vector<double> obs;   // [n][4]
  std::vector<float> f3;     // [n][2] current-left features as floats (the area test of process() is float arithmetic)
};

void build(Vo& V, const uba_vo_params* p, int n, const float* quads) {
  V.P = *p; V.n = n; V.pts.resize((size_t)n * 4); V.obs.resize((size_t)n * 4); V.f3.resize((size_t)n * 2);
  for (int i = 0; i < n; i++) {
    const float* q = quads + (size_t)i * 8;
    // project3D: d = (f1.x - cu1) - (f2.x - cu2); (x b, y b, fu1 b, d > 0 ? d : 1e-5), then normalize()
    const double d0 = (q[0] - p->cu1) - (q[2] - p->cu2);
    double h[4] = {(q[0] - p->cu1) * p->baseline, (q[1] - p->cv1) * p->baseline, p->fu1 * p->baseline, d0 > 0 ? d0 : 0.00001};
    h[0] /= h[3]; h[1] /= h[3]; h[2] /= h[3]; h[3] /= h[3];
    for (int a = 0; a < 4; a++) V.pts[(size_t)i * 4 + a] = h[a];
    for (int a = 0; a < 4; a++) V.obs[(size_t)i * 4 + a] = q[4 + a];
    V.f3[(size_t)i * 2] = q[4]; V.f3[(size_t)i * 2 + 1] = q[5];
  }
}

// Euler::getR4 (src/core/rotation_utils.cpp:36-46) and its three derivatives (:58-90)
void euler_mats(const double* s, double R[3][3], double Dr[3][3], double Dp[3][3], double Dy[3][3]) {
  const double cr = cos(s[0]), sr = sin(s[0]), cp = cos(s[1]), sp = sin(s[1]), cy = cos(s[2]), sy = sin(s[2]);
  const double r[3][3] = {{cp * cy, cp * sy, -sp}, {sp * sr * cy - cr * sy, sr * sp * sy + cr * cy, cp * sr}, {cr * sp * cy + sr * sy, cr * sp * sy - sr * cy, cp * cr}};
  const double dr[3][3] = {{0, 0, 0}, {cr * sp * cy + sr * sy, cr * sp * sy - sr * cy, cr * cp}, {-sr * sp * cy + cr * sy, -sr * sp * sy - cr * cy, -sr * cp}};
  const double dp[3][3] = {{-cy * sp, -sy * sp, -cp}, {sr * cp * cy, sr * cp * sy, -sr * sp}, {cr * cp * cy, cr * cp * sy, -cr * sp}};
  const double dy[3][3] = {{-cp * sy, cp * cy, 0}, {-sr * sp * sy - cr * cy, sr * sp * cy - cr * sy, 0}, {-cr * sp * sy + sr * cy, cr * sp * cy + sr * sy, 0}};
  memcpy(R, r, sizeof(r)); memcpy(Dr, dr, sizeof(dr)); memcpy(Dp, dp, sizeof(dp)); memcpy(Dy, dy, sizeof(dy));
}

// reproject :116-143 for one point: Tr = R^T with the translation column; P1, P2 applied to Tr * X, then normalize()
void reproject_one(const Vo& V, const double* state, const double* X4, double pt[4], double pred[4]) {
  double R[3][3], a[3][3], b[3][3], c[3][3];
  euler_mats(state, R, a, b, c);
  for (int i = 0; i < 3; i++) pt[i] = R[0][i] * X4[0] + R[1][i] * X4[1] + R[2][i] * X4[2] + state[3 + i] * X4[3];
  pt[3] = X4[3];
  const uba_vo_params& P = V.P;
  double l[3] = {P.fu1 * pt[0] + P.cu1 * pt[2], P.fv1 * pt[1] + P.cv1 * pt[2], pt[2]};
  double r[3] = {P.fu2 * pt[0] + P.cu2 * pt[2] - P.baseline * P.fu2 * pt[3], P.fv2 * pt[1] + P.cv2 * pt[2], pt[2]};
  pred[0] = l[0] / l[2]; pred[1] = l[1] / l[2]; pred[2] = r[0] / r[2]; pred[3] = r[1] / r[2];
}

// residuals (observed - predicted, :183-189), J [6][4 n_sel] (:291-329), A = J J^T, B = J r (:199-203)
void linearize(const Vo& V, const double* state, int n_sel, const int32_t* sel, double* res, double* J, double* A, double* B, double* rr) {
  double R[3][3], Dr[3][3], Dp[3][3], Dy[3][3];
  euler_mats(state, R, Dr, Dp, Dy);
  const uba_vo_params& P = V.P;
  const int cols = 4 * n_sel;
  std::vector<double> Jl((size_t)6 * cols), rl(cols);
  *rr = 0;
  for (int i = 0; i < n_sel; i++) {
    const double* X = &V.pts[(size_t)sel[i] * 4];
    double pt[4], pred[4];
    reproject_one(V, state, X, pt, pred);
    for (int q = 0; q < 4; q++) { rl[i * 4 + q] = V.obs[(size_t)sel[i] * 4 + q] - pred[q]; *rr += rl[i * 4 + q] * rl[i * 4 + q]; }
    // pt_next = Tr * pt, normalize() divides by the homogeneous coordinate (1)
    const double pn[3] = {pt[0] / pt[3], pt[1] / pt[3], pt[2] / pt[3]};
    for (int j = 0; j < 6; j++) {
      double d[3];
      const double (*D)[3] = j == 0 ? Dr : (j == 1 ? Dp : Dy);
      if (j < 3) for (int a = 0; a < 3; a++) d[a] = D[0][a] * X[0] + D[1][a] * X[1] + D[2][a] * X[2];   // (dRdx).t() * pt_
      else { d[0] = j == 3; d[1] = j == 4; d[2] = j == 5; }
      Jl[(size_t)j * cols + i * 4 + 0] = P.fu1 * (d[0] * pn[2] - pn[0] * d[2]) / (pn[2] * pn[2]);
      Jl[(size_t)j * cols + i * 4 + 1] = P.fv1 * (d[1] * pn[2] - pn[1] * d[2]) / (pn[2] * pn[2]);
      Jl[(size_t)j * cols + i * 4 + 2] = P.fu2 * (d[0] * pn[2] - (pn[0] - P.baseline) * d[2]) / (pn[2] * pn[2]);
      Jl[(size_t)j * cols + i * 4 + 3] = P.fv2 * (d[1] * pn[2] - pn[1] * d[2]) / (pn[2] * pn[2]);
    }
  }
  for (int a = 0; a < 6; a++) {
    for (int b = 0; b < 6; b++) { double s = 0; for (int c = 0; c < cols; c++) s += Jl[(size_t)a * cols + c] * Jl[(size_t)b * cols + c]; A[a * 6 + b] = s; }
    double s = 0; for (int c = 0; c < cols; c++) s += Jl[(size_t)a * cols + c] * rl[c];
    B[a] = s;
  }
  if (res) memcpy(res, rl.data(), sizeof(double) * cols);
  if (J) memcpy(J, Jl.data(), sizeof(double) * 6 * cols);
}

// cv::solve(A, B, X, DECOMP_QR) for a 6x6 system: Householder QR; false when rank-deficient
bool solve6(const double* A, const double* B, double* X) {
  double a[6][6], b[6];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) a[i][j] = A[i * 6 + j]; b[i] = B[i]; }
  double dmax = 0;
  for (int k = 0; k < 6; k++) {
    double nrm = 0; for (int i = k; i < 6; i++) nrm += a[i][k] * a[i][k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) return false;
    const double alpha = a[k][k] > 0 ? -nrm : nrm;
    double v[6], vv = 0;
    for (int i = k; i < 6; i++) v[i] = a[i][k];
    v[k] -= alpha;
    for (int i = k; i < 6; i++) vv += v[i] * v[i];
    if (vv > 0) {
      for (int j = k; j < 6; j++) { double d = 0; for (int i = k; i < 6; i++) d += v[i] * a[i][j]; d = 2 * d / vv; for (int i = k; i < 6; i++) a[i][j] -= d * v[i]; }
      double d = 0; for (int i = k; i < 6; i++) d += v[i] * b[i]; d = 2 * d / vv; for (int i = k; i < 6; i++) b[i] -= d * v[i];
    }
    dmax = fmax(dmax, fabs(a[k][k]));
  }
  for (int k = 0; k < 6; k++) if (!(fabs(a[k][k]) > 1e-15 * dmax)) return false;
  for (int i = 5; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < 6; k++) s -= a[i][k] * X[k]; X[i] = s / a[i][i]; }
  for (int i = 0; i < 6; i++) if (!std::isfinite(X[i])) return false;
  return true;
}

double res_sq(const Vo& V, const double* state, int n_sel, const int32_t* sel) {
  double rr = 0;
  for (int i = 0; i < n_sel; i++) {
    double pt[4], pred[4];
    reproject_one(V, state, &V.pts[(size_t)sel[i] * 4], pt, pred);
    for (int q = 0; q < 4; q++) { const double r = V.obs[(size_t)sel[i] * 4 + q] - pred[q]; rr += r * r; }
  }
  return rr;
}

// optimize :165-283 (bounded loop, see the header); returns the reference's bool
bool optimize(const Vo& V, double* state, int n_sel, const int32_t* sel, int* iters_out, int* stop_out) {
  if (n_sel < 3) { if (iters_out) *iters_out = 0; if (stop_out) *stop_out = NO_STOP; return false; }
  const uba_vo_params& P = V.P;
  int k = 0;
  double v = 2, tau = 1e-5, mu = 1e-20;
  int stop = NO_STOP;
  for (;; k++) {
    double A[36], B[6], rr;
    linearize(V, state, n_sel, sel, nullptr, nullptr, A, B, &rr);
    if (rr / (4 * n_sel) < P.e1) stop = SMALL_REPROJ_ERROR;
    double binf = 0; for (int i = 0; i < 6; i++) binf = fmax(binf, fabs(B[i]));
    if (binf < P.e2) stop = SMALL_GRADIENT;
    if (P.method == 1 && k == 0) { double mx = A[0]; for (int i = 1; i < 6; i++) mx = fmax(mx, A[i * 7]); mu = fmax(mu, mx); mu = tau * mu; }
    for (;;) {
      if (P.method == 1) for (int i = 0; i < 6; i++) A[i * 7] += mu;
      double X[6];
      if (solve6(A, B, X)) {
        double nx = 0, ns = 0; for (int i = 0; i < 6; i++) { nx += X[i] * X[i]; ns += state[i] * state[i]; }
        if (sqrt(nx) <= P.e3 * sqrt(ns)) { stop = SMALL_INCREMENT; break; }
        if (P.method == 0) { for (int i = 0; i < 6; i++) state[i] += X[i]; break; }
        double xt[6]; for (int i = 0; i < 6; i++) xt[i] = state[i] + X[i];
        const double rt = res_sq(V, xt, n_sel, sel);
        double den = 0; for (int i = 0; i < 6; i++) den += X[i] * (mu * X[i] + B[i]);
        const double rho = (rr - rt) / den;
        if (rho > 0) {
          mu *= fmax(0.333, 1 - pow(2 * rho - 1, 3));
          v = 2;
          if (pow(rr - rt, 2) < P.e4 * rr) stop = SMALL_DECREASE_FUNCTION;
          for (int i = 0; i < 6; i++) state[i] = xt[i];
          break;
        }
        mu *= v;
        const double v2 = 2 * v;
        if (v2 <= v) { stop = NO_CONVERGENCE; break; }
        v = v2;
      } else { stop = NO_CONVERGENCE; break; }
    }
    if (stop != NO_STOP) break;
    if (k + 1 >= P.max_iter) { stop = MAX_ITERATIONS; break; }
  }
  if (iters_out) *iters_out = k + 1;
  if (stop_out) *stop_out = stop;
  return !(stop == NO_CONVERGENCE || stop == MAX_ITERATIONS);
}

int inliers(const Vo& V, const double* state, int32_t* idx) {
  int k = 0;
  const double thr2 = pow(V.P.inlier_threshold, 2);
  for (int i = 0; i < V.n; i++) {
    double pt[4], pred[4];
    reproject_one(V, state, &V.pts[(size_t)i * 4], pt, pred);
    const double score = pow(pred[0] - V.obs[(size_t)i * 4], 2) + pow(pred[1] - V.obs[(size_t)i * 4 + 1], 2) + pow(pred[2] - V.obs[(size_t)i * 4 + 2], 2) + pow(pred[3] - V.obs[(size_t)i * 4 + 3], 2);
    if (score < thr2) { if (idx) idx[k] = i; k++; }
  }
  return k;
}

}  // namespace

extern "C" {

int uba_ref_vo_project3d(const uba_vo_params* p, int n, const float* quads, double* pts4) {
  Vo V; build(V, p, n, quads);
  memcpy(pts4, V.pts.data(), sizeof(double) * 4 * n);
  return 0;
}

int uba_ref_vo_linearize(const uba_vo_params* p, int n, const float* quads, const double* state, int n_sel, const int32_t* sel,
                         double* A36, double* B6, double* res, double* J) {
  Vo V; build(V, p, n, quads);
  double A[36], B[6], rr;
  linearize(V, state, n_sel, sel, res, J, A, B, &rr);
  if (A36) memcpy(A36, A, sizeof(A));
  if (B6) memcpy(B6, B, sizeof(B));
  return 0;
}

int uba_ref_vo_optimize(const uba_vo_params* p, int n, const float* quads, const double* init, int n_sel, const int32_t* sel,
                        double* state_out, int32_t* iters, int32_t* stop) {
  Vo V; build(V, p, n, quads);
  double s[6]; memcpy(s, init, sizeof(s));
  int it = 0, st = 0;
  const bool ok = optimize(V, s, n_sel, sel, &it, &st);
  memcpy(state_out, s, sizeof(s));
  if (iters) *iters = it;
  if (stop) *stop = st;
  return ok ? 1 : 0;
}

int uba_ref_vo_inliers(const uba_vo_params* p, int n, const float* quads, const double* state, int32_t* idx) {
  Vo V; build(V, p, n, quads);
  return inliers(V, state, idx);
}

// the RANSAC loop of process() (:59-75) over caller-supplied triples; returns the index of the winning hypothesis
int uba_ref_vo_ransac(const uba_vo_params* p, int n, const float* quads, const double* init, int n_hyp, const int32_t* triples,
                      int32_t* counts, int32_t* ok_out, double* states) {
  Vo V; build(V, p, n, quads);
  int best = -1, best_cnt = 0;
  for (int h = 0; h < n_hyp; h++) {
    const int32_t* s = triples + (size_t)h * 3;
    double st[6]; memcpy(st, init, sizeof(st));
    int ok = 0, cnt = 0;
    const float* a = &V.f3[(size_t)s[0] * 2]; const float* b = &V.f3[(size_t)s[1] * 2]; const float* c = &V.f3[(size_t)s[2] * 2];
    if ((a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1])) / 2 > 1000) {
      if (optimize(V, st, 3, s, nullptr, nullptr)) { ok = 1; cnt = inliers(V, st, nullptr); if (cnt > best_cnt) { best = h; best_cnt = cnt; } }
    }
    if (counts) counts[h] = cnt;
    if (ok_out) ok_out[h] = ok;
    if (states) memcpy(states + (size_t)h * 6, st, sizeof(st));
  }
  return best;
}

}  // extern "C"
