// uba_vo.cu — POSE-ONLY mode: the numerical core of the reference's frame-to-frame stereo visual odometry
// (src/vo/StereoVisualOdometry.cpp), SURVEY.md §8(f) ranks 1 and 4.  Points are fixed (triangulated from the previous
// stereo pair), the 6 unknowns are the Euler angles and the translation of the motion, so there is no Schur
// complement: each evaluation is a 6x6 normal-equation system accumulated over the selected matches.
//
//   project3D            (:22-32)    -> k_vo_project3d   thread per quad match
//   reproject            (:116-143)  -> vo_project (device)
//   updateJacobian       (:291-329)  -> vo_point_terms (device): closed form of the same 4x6 block
//   optimize             (:165-283)  -> vo_optimize_cta: Gauss-Newton / Levenberg-Marquardt, one CTA per problem
//   computeInliers       (:94-114)   -> scoring loop of k_vo_ransac / k_vo_flags
//   process (RANSAC)     (:34-92)    -> k_vo_ransac: ALL hypotheses at once, one CTA each (the reference runs its 200
//                                       3-point fits + 200 scoring passes one after the other on one core)
//
// Deliberate difference, documented in INTEGRATION.md: the reference's loop condition
//   while(!(k++ < (m_param.max_iter?stop:stop=StopCondition::MAX_ITERATIONS)))            (:277)
// compares the iteration counter with the StopCondition ENUM VALUE, so optimize() only returns when a stop condition
// fires at an iteration index smaller than its enum value and otherwise never returns (it hangs on ordinary data:
// oracle/_ref, the reference compiled as is, does).  Here the loop ends when a stop condition is set or after
// max_iter iterations (-> MAX_ITERATIONS -> "false"), which is what the parameter documents.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/uba.h"
#include "uba_vo.h"

namespace uba {

namespace {

struct VoDev {
  uba_vo_params P;
  int n;
  const float* quads;   // [n][8]
  double* pts;          // [n][3]  triangulated points (the homogeneous w is 1 after normalisation)
  double* obs;          // [n][4]  current left x, y, current right x, y
};

// rows of Tr = [R(euler)^T | t]  (reproject :122-127; Euler::getR4, src/core/rotation_utils.cpp:36-46)
struct VoPose { double T[3][4]; double dR[3][3][3]; };   // dR[a] = d(R^T)/d(angle a) (getdRdr/p/y transposed, :294-296)

__device__ void vo_pose(const double* s, VoPose& o, bool with_derivs) {
  double sr, cr, sp, cp, sy, cy;
  sincos(s[0], &sr, &cr); sincos(s[1], &sp, &cp); sincos(s[2], &sy, &cy);
  const double R[3][3] = {{cp * cy, cp * sy, -sp}, {sp * sr * cy - cr * sy, sr * sp * sy + cr * cy, cp * sr}, {cr * sp * cy + sr * sy, cr * sp * sy - sr * cy, cp * cr}};
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) o.T[i][j] = R[j][i]; o.T[i][3] = s[3 + i]; }
  if (!with_derivs) return;
  const double Dr[3][3] = {{0, 0, 0}, {cr * sp * cy + sr * sy, cr * sp * sy - sr * cy, cr * cp}, {-sr * sp * cy + cr * sy, -sr * sp * sy - cr * cy, -sr * cp}};
  const double Dp[3][3] = {{-cy * sp, -sy * sp, -cp}, {sr * cp * cy, sr * cp * sy, -sr * sp}, {cr * cp * cy, cr * cp * sy, -cr * sp}};
  const double Dy[3][3] = {{-cp * sy, cp * cy, 0}, {-sr * sp * sy - cr * cy, sr * sp * cy - cr * sy, 0}, {-cr * sp * sy + sr * cy, cr * sp * cy + sr * sy, 0}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { o.dR[0][i][j] = Dr[j][i]; o.dR[1][i][j] = Dp[j][i]; o.dR[2][i][j] = Dy[j][i]; }
}

// predictions of one point: left (x, y), right (x, y)   (:129-140)
__device__ __forceinline__ void vo_project(const uba_vo_params& P, const VoPose& T, const double* X, double* q, double* pred) {
  for (int i = 0; i < 3; i++) q[i] = T.T[i][0] * X[0] + T.T[i][1] * X[1] + T.T[i][2] * X[2] + T.T[i][3];
  const double iz = 1.0 / q[2];
  pred[0] = (P.fu1 * q[0] + P.cu1 * q[2]) * iz;
  pred[1] = (P.fv1 * q[1] + P.cv1 * q[2]) * iz;
  pred[2] = (P.fu2 * q[0] + P.cu2 * q[2] - P.baseline * P.fu2) * iz;
  pred[3] = (P.fv2 * q[1] + P.cv2 * q[2]) * iz;
}

// 4x6 Jacobian block of one point, J[a][row]   (:312-326)
__device__ __forceinline__ void vo_jac(const uba_vo_params& P, const VoPose& T, const double* X, const double* q, double J[6][4]) {
  const double iz2 = 1.0 / (q[2] * q[2]);
  for (int a = 0; a < 6; a++) {
    double d[3];
    if (a < 3) for (int i = 0; i < 3; i++) d[i] = T.dR[a][i][0] * X[0] + T.dR[a][i][1] * X[1] + T.dR[a][i][2] * X[2];
    else { d[0] = a == 3; d[1] = a == 4; d[2] = a == 5; }
    J[a][0] = P.fu1 * (d[0] * q[2] - q[0] * d[2]) * iz2;
    J[a][1] = P.fv1 * (d[1] * q[2] - q[1] * d[2]) * iz2;
    J[a][2] = P.fu2 * (d[0] * q[2] - (q[0] - P.baseline) * d[2]) * iz2;
    J[a][3] = P.fv2 * (d[1] * q[2] - q[1] * d[2]) * iz2;
  }
}

constexpr int kAcc = 28;   // A upper triangle (21), B (6), r^T r (1)

// Sum of acc[kAcc] over the CTA; result valid in every thread.  sm: [32 * kAcc] scratch.
__device__ void cta_sum(double* acc, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < kAcc; i++) { double v = acc[i]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); acc[i] = v; }
  __syncthreads();
  if (lane == 0) for (int i = 0; i < kAcc; i++) sm[warp * kAcc + i] = acc[i];
  __syncthreads();
  for (int i = 0; i < kAcc; i++) { double v = 0.0; for (int w = 0; w < nw; w++) v += sm[w * kAcc + i]; acc[i] = v; }
}

// normal equations over the selection at `state`: acc = {A upper, B = J r, r^T r}; r = observed - predicted (:183-189)
__device__ void vo_accumulate(const VoDev& D, const double* state, const int32_t* sel, const unsigned char* flags, int n_sel, bool with_jac, double* acc, double* sm) {
  VoPose T;
  vo_pose(state, T, with_jac);
  for (int i = 0; i < kAcc; i++) acc[i] = 0.0;
  const int count = sel ? n_sel : D.n;
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    const int m = sel ? sel[k] : k;
    if (!sel && flags && !flags[m]) continue;
    const double* X = D.pts + (size_t)m * 3;
    double q[3], pred[4], r[4];
    vo_project(D.P, T, X, q, pred);
    for (int c = 0; c < 4; c++) { r[c] = D.obs[(size_t)m * 4 + c] - pred[c]; acc[27] += r[c] * r[c]; }
    if (with_jac) {
      double J[6][4];
      vo_jac(D.P, T, X, q, J);
      int e = 0;
      for (int a = 0; a < 6; a++) {
        for (int b = a; b < 6; b++) acc[e++] += J[a][0] * J[b][0] + J[a][1] * J[b][1] + J[a][2] * J[b][2] + J[a][3] * J[b][3];
        acc[21 + a] += J[a][0] * r[0] + J[a][1] * r[1] + J[a][2] * r[2] + J[a][3] * r[3];
      }
    }
  }
  cta_sum(acc, sm);
}

// 6x6 solve by Householder QR, like cv::solve(A, B, X, DECOMP_QR) (:219); false when a column has no pivot left
__device__ bool vo_solve6(const double* Aup, double mu, const double* B, double* X) {
  double a[6][6], b[6];
  int e = 0;
  for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++) { a[i][j] = a[j][i] = Aup[e++]; }
  double dmax = 0.0;
  for (int i = 0; i < 6; i++) { a[i][i] += mu; b[i] = B[i]; }
  for (int k = 0; k < 6; k++) {
    double nrm = 0.0;
    for (int i = k; i < 6; i++) nrm += a[i][k] * a[i][k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) return false;
    const double alpha = a[k][k] > 0.0 ? -nrm : nrm;
    double v[6], vv = 0.0;
    for (int i = k; i < 6; i++) v[i] = a[i][k];
    v[k] -= alpha;
    for (int i = k; i < 6; i++) vv += v[i] * v[i];
    if (vv > 0.0) {
      for (int j = k; j < 6; j++) { double d = 0.0; for (int i = k; i < 6; i++) d += v[i] * a[i][j]; d = 2.0 * d / vv; for (int i = k; i < 6; i++) a[i][j] -= d * v[i]; }
      double d = 0.0; for (int i = k; i < 6; i++) d += v[i] * b[i]; d = 2.0 * d / vv; for (int i = k; i < 6; i++) b[i] -= d * v[i];
    }
    dmax = fmax(dmax, fabs(a[k][k]));
  }
  for (int k = 0; k < 6; k++) if (!(fabs(a[k][k]) > 1e-15 * dmax)) return false;
  for (int i = 5; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < 6; k++) s -= a[i][k] * X[k]; X[i] = s / a[i][i]; }
  for (int i = 0; i < 6; i++) if (!isfinite(X[i])) return false;
  return true;
}

enum VoStop { VO_NO_STOP = 0, VO_SMALL_GRADIENT = 1, VO_SMALL_INCREMENT = 2, VO_MAX_ITERATIONS = 3, VO_SMALL_DECREASE = 4, VO_SMALL_REPROJ = 5, VO_NO_CONVERGENCE = 6 };

// optimize() (:165-283) by the whole CTA.  state: in/out, shared memory [6].  Returns the stop condition (same in every
// thread) and the number of outer iterations in *iters.
__device__ int vo_optimize_cta(const VoDev& D, double* state, const int32_t* sel, const unsigned char* flags, int n_sel, double* sm, int* iters) {
  const uba_vo_params& P = D.P;
  __shared__ double s_x[6], s_test[6], s_mu, s_v;
  __shared__ int s_stop, s_inner;
  double acc[kAcc];
  if (threadIdx.x == 0) { s_mu = 1e-20; s_v = 2.0; s_stop = VO_NO_STOP; }
  __syncthreads();
  int k = 0;
  const int rows = 4 * n_sel;
  for (;; k++) {
    vo_accumulate(D, state, sel, flags, n_sel, true, acc, sm);
    if (threadIdx.x == 0) {
      if (acc[27] / rows < P.e1) s_stop = VO_SMALL_REPROJ;                 // :192-196
      double binf = 0.0; for (int i = 0; i < 6; i++) binf = fmax(binf, fabs(acc[21 + i]));
      if (binf < P.e2) s_stop = VO_SMALL_GRADIENT;                          // :205-207
      if (P.method == 1 && k == 0) {                                        // :209-215
        const double dg[6] = {acc[0], acc[6], acc[11], acc[15], acc[18], acc[20]};
        double mx = dg[0]; for (int i = 1; i < 6; i++) mx = fmax(mx, dg[i]);
        s_mu = 1e-5 * fmax(s_mu, mx);
      }
      s_inner = 1;
    }
    __syncthreads();
    double damp = 0.0;                                                      // LM: A += mu I on EVERY pass of the inner loop (:219-220)
    while (s_inner) {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (P.method == 1) damp += s_mu;
        double X[6];
        if (vo_solve6(acc, damp, acc + 21, X)) {
          double nx = 0.0, ns = 0.0;
          for (int i = 0; i < 6; i++) { nx += X[i] * X[i]; ns += state[i] * state[i]; s_x[i] = X[i]; }
          if (sqrt(nx) <= P.e3 * sqrt(ns)) { s_stop = VO_SMALL_INCREMENT; s_inner = 0; }
          else if (P.method == 0) { for (int i = 0; i < 6; i++) state[i] += X[i]; s_inner = 0; }
          else { for (int i = 0; i < 6; i++) s_test[i] = state[i] + X[i]; s_inner = 2; }
        } else { s_stop = VO_NO_CONVERGENCE; s_inner = 0; }
      }
      __syncthreads();
      if (s_inner == 2) {                                                   // Levenberg-Marquardt trial point (:235-268)
        double acct[kAcc];
        vo_accumulate(D, s_test, sel, flags, n_sel, false, acct, sm);
        if (threadIdx.x == 0) {
          double den = 0.0;
          for (int i = 0; i < 6; i++) den += s_x[i] * (s_mu * s_x[i] + acc[21 + i]);
          const double diff = acc[27] - acct[27];
          const double rho = diff / den;
          if (rho > 0.0) {
            const double q = 2.0 * rho - 1.0;
            s_mu *= fmax(0.333, 1.0 - q * q * q);
            s_v = 2.0;
            if (diff * diff < P.e4 * acc[27]) s_stop = VO_SMALL_DECREASE;
            for (int i = 0; i < 6; i++) state[i] = s_test[i];
            s_inner = 0;
          } else {
            s_mu *= s_v;
            const double v2 = 2.0 * s_v;
            if (v2 <= s_v) { s_stop = VO_NO_CONVERGENCE; s_inner = 0; } else { s_v = v2; s_inner = 1; }
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    if (s_stop != VO_NO_STOP) break;
    if (k + 1 >= P.max_iter) { if (threadIdx.x == 0) s_stop = VO_MAX_ITERATIONS; __syncthreads(); break; }
  }
  if (iters) *iters = k + 1;
  return s_stop;
}

__device__ __forceinline__ bool vo_is_inlier(const VoDev& D, const VoPose& T, int m, double thr2) {
  double q[3], pred[4];
  vo_project(D.P, T, D.pts + (size_t)m * 3, q, pred);
  double s = 0.0;
  for (int c = 0; c < 4; c++) { const double d = pred[c] - D.obs[(size_t)m * 4 + c]; s += d * d; }
  return s < thr2;                                                          // :108-110
}

}  // namespace

// project3D (:22-32) + updateObservations (:285-289)
__global__ void k_vo_project3d(VoDev D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.n) return;
  const float* q = D.quads + (size_t)i * 8;
  const uba_vo_params& P = D.P;
  const double d0 = ((double)q[0] - P.cu1) - ((double)q[2] - P.cu2);
  const double d = d0 > 0 ? d0 : 0.00001;
  D.pts[(size_t)i * 3 + 0] = ((double)q[0] - P.cu1) * P.baseline / d;
  D.pts[(size_t)i * 3 + 1] = ((double)q[1] - P.cv1) * P.baseline / d;
  D.pts[(size_t)i * 3 + 2] = P.fu1 * P.baseline / d;
  for (int c = 0; c < 4; c++) D.obs[(size_t)i * 4 + c] = (double)q[4 + c];
}

// One CTA per RANSAC hypothesis: the 3-point fit from `init` (process :64-73), then the inlier count over all matches.
__global__ void __launch_bounds__(128) k_vo_ransac(VoDev D, const int32_t* __restrict__ triples, const double* __restrict__ init,
                                                    int32_t* __restrict__ ok, double* __restrict__ states, int32_t* __restrict__ counts) {
  __shared__ double sm[32 * kAcc];
  __shared__ double s_state[6];
  __shared__ int s_cnt;
  const int h = blockIdx.x;
  const int32_t* sel = triples + (size_t)h * 3;
  // "selecting random matches scattered in the image": twice the triangle area of the current-left features (:66)
  const float* a = D.quads + (size_t)sel[0] * 8 + 4; const float* b = D.quads + (size_t)sel[1] * 8 + 4; const float* c = D.quads + (size_t)sel[2] * 8 + 4;
  const float area = (a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1])) / 2;
  const bool distinct = sel[0] != sel[1] && sel[0] != sel[2] && sel[1] != sel[2];
  if (!(area > 1000.f) || !distinct) { if (threadIdx.x == 0) { ok[h] = 0; counts[h] = 0; for (int i = 0; i < 6; i++) states[h * 6 + i] = init[i]; } return; }
  if (threadIdx.x < 6) s_state[threadIdx.x] = init[threadIdx.x];
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int stop = vo_optimize_cta(D, s_state, sel, nullptr, 3, sm, nullptr);
  const bool good = !(stop == VO_NO_CONVERGENCE || stop == VO_MAX_ITERATIONS);
  int cnt = 0;
  if (good) {
    VoPose T;
    vo_pose(s_state, T, false);
    const double thr2 = D.P.inlier_threshold * D.P.inlier_threshold;
    for (int m = threadIdx.x; m < D.n; m += blockDim.x) cnt += vo_is_inlier(D, T, m, thr2) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) { ok[h] = good ? 1 : 0; counts[h] = good ? s_cnt : 0; }
  if (threadIdx.x < 6) states[h * 6 + threadIdx.x] = s_state[threadIdx.x];
}

// inlier flags of one state (computeInliers :94-114)
__global__ void k_vo_flags(VoDev D, const double* __restrict__ state, unsigned char* __restrict__ flags) {
  __shared__ VoPose T;
  if (threadIdx.x == 0) vo_pose(state, T, false);
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= D.n) return;
  flags[m] = vo_is_inlier(D, T, m, D.P.inlier_threshold * D.P.inlier_threshold) ? 1 : 0;
}

// optimize() over an index list (sel != null) or over the flagged matches; out: state6, stop, iterations
__global__ void __launch_bounds__(256) k_vo_refine(VoDev D, const int32_t* __restrict__ sel, const unsigned char* __restrict__ flags, int n_sel,
                                                    const double* __restrict__ init, double* __restrict__ out_state, int32_t* __restrict__ out_info) {
  __shared__ double sm[32 * kAcc];
  __shared__ double s_state[6];
  if (threadIdx.x < 6) s_state[threadIdx.x] = init[threadIdx.x];
  __syncthreads();
  int iters = 0;
  const int stop = vo_optimize_cta(D, s_state, sel, flags, n_sel, sm, &iters);
  if (threadIdx.x < 6) out_state[threadIdx.x] = s_state[threadIdx.x];
  if (threadIdx.x == 0) { out_info[0] = stop; out_info[1] = iters; }
}

// parity dump: per selected match predictions, residuals, the 6 x 4 Jacobian block; A and B summed by one CTA
__global__ void __launch_bounds__(256) k_vo_linearize(VoDev D, const int32_t* __restrict__ sel, int n_sel, const double* __restrict__ state,
                                                       double* __restrict__ res, double* __restrict__ J, double* __restrict__ AB) {
  __shared__ double sm[32 * kAcc];
  __shared__ double s_state[6];
  if (threadIdx.x < 6) s_state[threadIdx.x] = state[threadIdx.x];
  __syncthreads();
  VoPose T;
  vo_pose(s_state, T, true);
  for (int k = threadIdx.x; k < n_sel; k += blockDim.x) {
    const int m = sel[k];
    double q[3], pred[4], Jb[6][4];
    vo_project(D.P, T, D.pts + (size_t)m * 3, q, pred);
    vo_jac(D.P, T, D.pts + (size_t)m * 3, q, Jb);
    for (int c = 0; c < 4; c++) {
      if (res) res[(size_t)k * 4 + c] = D.obs[(size_t)m * 4 + c] - pred[c];
      if (J) for (int a = 0; a < 6; a++) J[(size_t)a * 4 * n_sel + (size_t)k * 4 + c] = Jb[a][c];
    }
  }
  double acc[kAcc];
  vo_accumulate(D, s_state, sel, nullptr, n_sel, true, acc, sm);
  if (threadIdx.x == 0) for (int i = 0; i < kAcc; i++) AB[i] = acc[i];
}

}  // namespace uba

// ------------------------------------------------------------------------------------------------
// host side: the uba_vo_* entry points of include/uba.h
// ------------------------------------------------------------------------------------------------
using namespace uba;

struct uba_vo_state {
  uba_vo_params P{};
  int n = 0, n_hyp = 0, best = -1;
  bool have_matches = false, have_flags = false;
  float* d_quads = nullptr; double* d_pts = nullptr; double* d_obs = nullptr; unsigned char* d_flags = nullptr;
  int32_t* d_sel = nullptr; int32_t* d_ok = nullptr; int32_t* d_counts = nullptr; int32_t* d_info = nullptr;
  double* d_states = nullptr; double* d_state = nullptr; double* d_scratch = nullptr;
  size_t cap_n = 0, cap_sel = 0, cap_hyp = 0, cap_scratch = 0;
  std::vector<unsigned char> flags_h;
};

namespace {
template <typename T> cudaError_t grow(T*& p, size_t& cap, size_t need, size_t elems_per) {
  if (need <= cap && p) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  const cudaError_t e = cudaMalloc((void**)&p, (need ? need : 1) * elems_per * sizeof(T));
  return e;
}
VoDev view(const uba_vo_state* s) { VoDev D; D.P = s->P; D.n = s->n; D.quads = s->d_quads; D.pts = s->d_pts; D.obs = s->d_obs; return D; }
}  // namespace

#define VO_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return uba_vo_fail(h, UBA_ERR_CUDA, #call, cudaGetErrorString(e_)); } while (0)

extern "C" {

int uba_vo_set_matches(uba_handle* h, const uba_vo_params* params, int n, const float* quads8) {
  if (!h || !params || n < 0 || (n && !quads8)) return UBA_ERR_INVALID_ARGUMENT;
  uba_vo_state* s = uba_vo_get(h, true);
  cudaStream_t st = uba_vo_stream(h);
  s->P = *params; s->n = n; s->have_flags = false; s->best = -1; s->n_hyp = 0;
  if ((size_t)n > s->cap_n || !s->d_quads) {
    size_t c0 = s->cap_n, c1 = s->cap_n, c2 = s->cap_n, c3 = s->cap_n;
    VO_CU(grow(s->d_quads, c0, (size_t)n, 8)); VO_CU(grow(s->d_pts, c1, (size_t)n, 3)); VO_CU(grow(s->d_obs, c2, (size_t)n, 4)); VO_CU(grow(s->d_flags, c3, (size_t)n, 1));
    s->cap_n = (size_t)n;
  }
  if (!s->d_state) { VO_CU(cudaMalloc((void**)&s->d_state, 12 * sizeof(double))); VO_CU(cudaMalloc((void**)&s->d_info, 4 * sizeof(int32_t))); }
  if (n) {
    VO_CU(cudaMemcpyAsync(s->d_quads, quads8, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, st));
    k_vo_project3d<<<(n + 127) / 128, 128, 0, st>>>(view(s));
    uba_vo_count(h, 1);
  }
  VO_CU(cudaStreamSynchronize(st));   // the caller's quads buffer is free again on return
  s->have_matches = true;
  return UBA_OK;
}

int uba_vo_get_points(uba_handle* h, double* pts4) {
  uba_vo_state* s = h ? uba_vo_get(h, false) : nullptr;
  if (!s || !pts4) return UBA_ERR_INVALID_ARGUMENT;
  if (!s->have_matches) return uba_vo_fail(h, UBA_ERR_STATE, "uba_vo_get_points", "no matches set");
  std::vector<double> p((size_t)s->n * 3);
  VO_CU(cudaMemcpy(p.data(), s->d_pts, sizeof(double) * 3 * s->n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < s->n; i++) { for (int a = 0; a < 3; a++) pts4[(size_t)i * 4 + a] = p[(size_t)i * 3 + a]; pts4[(size_t)i * 4 + 3] = 1.0; }
  return UBA_OK;
}

static int vo_upload_sel(uba_handle* h, uba_vo_state* s, const int32_t* sel, size_t count, cudaStream_t st) {
  for (size_t i = 0; i < count; i++) if (sel[i] < 0 || sel[i] >= s->n) return uba_vo_fail(h, UBA_ERR_INVALID_ARGUMENT, "selection", "match index out of range");
  if (count > s->cap_sel || !s->d_sel) { size_t c = s->cap_sel; VO_CU(grow(s->d_sel, c, count, 1)); s->cap_sel = count; }
  if (count) VO_CU(cudaMemcpyAsync(s->d_sel, sel, sizeof(int32_t) * count, cudaMemcpyHostToDevice, st));
  return UBA_OK;
}

int uba_vo_linearize(uba_handle* h, const double state6[6], int n_sel, const int32_t* selection, double* A36, double* B6,
                     double* residuals, double* J) {
  uba_vo_state* s = h ? uba_vo_get(h, false) : nullptr;
  if (!s || !state6 || n_sel <= 0 || !selection) return UBA_ERR_INVALID_ARGUMENT;
  if (!s->have_matches) return uba_vo_fail(h, UBA_ERR_STATE, "uba_vo_linearize", "no matches set");
  cudaStream_t st = uba_vo_stream(h);
  int rc = vo_upload_sel(h, s, selection, (size_t)n_sel, st);
  if (rc) return rc;
  const size_t need = (size_t)n_sel * 4 + (size_t)n_sel * 24 + 32;
  if (need > s->cap_scratch || !s->d_scratch) { size_t c = s->cap_scratch; VO_CU(grow(s->d_scratch, c, need, 1)); s->cap_scratch = need; }
  double* d_res = s->d_scratch; double* d_J = d_res + (size_t)n_sel * 4; double* d_AB = d_J + (size_t)n_sel * 24;
  VO_CU(cudaMemcpyAsync(s->d_state, state6, 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  k_vo_linearize<<<1, 256, 0, st>>>(view(s), s->d_sel, n_sel, s->d_state, d_res, d_J, d_AB);
  uba_vo_count(h, 1);
  double AB[28];
  VO_CU(cudaMemcpyAsync(AB, d_AB, sizeof(AB), cudaMemcpyDeviceToHost, st));
  if (residuals) VO_CU(cudaMemcpyAsync(residuals, d_res, sizeof(double) * 4 * n_sel, cudaMemcpyDeviceToHost, st));
  if (J) VO_CU(cudaMemcpyAsync(J, d_J, sizeof(double) * 24 * n_sel, cudaMemcpyDeviceToHost, st));
  VO_CU(cudaStreamSynchronize(st));
  if (A36) { int e = 0; for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++) { A36[i * 6 + j] = A36[j * 6 + i] = AB[e++]; } }
  if (B6) for (int i = 0; i < 6; i++) B6[i] = AB[21 + i];
  return UBA_OK;
}

int uba_vo_ransac(uba_handle* h, const double init6[6], int n_hyp, const int32_t* triples, int32_t* best_hyp, int32_t* inlier_counts,
                  int32_t* hyp_ok, double* hyp_states) {
  uba_vo_state* s = h ? uba_vo_get(h, false) : nullptr;
  if (!s || !init6 || n_hyp <= 0 || !triples) return UBA_ERR_INVALID_ARGUMENT;
  if (!s->have_matches) return uba_vo_fail(h, UBA_ERR_STATE, "uba_vo_ransac", "no matches set");
  if (s->n < 3) return uba_vo_fail(h, UBA_ERR_INVALID_ARGUMENT, "uba_vo_ransac", "fewer than 3 matches");
  cudaStream_t st = uba_vo_stream(h);
  int rc = vo_upload_sel(h, s, triples, (size_t)n_hyp * 3, st);
  if (rc) return rc;
  if ((size_t)n_hyp > s->cap_hyp || !s->d_ok) {
    size_t c0 = s->cap_hyp, c1 = s->cap_hyp, c2 = s->cap_hyp;
    VO_CU(grow(s->d_ok, c0, (size_t)n_hyp, 1)); VO_CU(grow(s->d_counts, c1, (size_t)n_hyp, 1)); VO_CU(grow(s->d_states, c2, (size_t)n_hyp, 6));
    s->cap_hyp = (size_t)n_hyp;
  }
  VO_CU(cudaMemcpyAsync(s->d_state, init6, 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  k_vo_ransac<<<n_hyp, 128, 0, st>>>(view(s), s->d_sel, s->d_state, s->d_ok, s->d_states, s->d_counts);
  uba_vo_count(h, 1);
  std::vector<int32_t> ok(n_hyp), cnt(n_hyp);
  VO_CU(cudaMemcpyAsync(ok.data(), s->d_ok, sizeof(int32_t) * n_hyp, cudaMemcpyDeviceToHost, st));
  VO_CU(cudaMemcpyAsync(cnt.data(), s->d_counts, sizeof(int32_t) * n_hyp, cudaMemcpyDeviceToHost, st));
  if (hyp_states) VO_CU(cudaMemcpyAsync(hyp_states, s->d_states, sizeof(double) * 6 * n_hyp, cudaMemcpyDeviceToHost, st));
  VO_CU(cudaStreamSynchronize(st));
  // "if more inliers obtained, inliers are saved" (:70-72): strictly more, so the EARLIEST best hypothesis wins
  int best = -1, best_cnt = 0;
  for (int i = 0; i < n_hyp; i++) if (ok[i] && cnt[i] > best_cnt) { best = i; best_cnt = cnt[i]; }
  s->best = best; s->n_hyp = n_hyp; s->have_flags = false;
  if (best >= 0) {
    k_vo_flags<<<(s->n + 127) / 128, 128, 0, st>>>(view(s), s->d_states + (size_t)best * 6, s->d_flags);
    uba_vo_count(h, 1);
    s->flags_h.resize(s->n);
    VO_CU(cudaMemcpyAsync(s->flags_h.data(), s->d_flags, s->n, cudaMemcpyDeviceToHost, st));
    VO_CU(cudaStreamSynchronize(st));
    s->have_flags = true;
  }
  if (best_hyp) *best_hyp = best;
  if (inlier_counts) std::memcpy(inlier_counts, cnt.data(), sizeof(int32_t) * n_hyp);
  if (hyp_ok) std::memcpy(hyp_ok, ok.data(), sizeof(int32_t) * n_hyp);
  return UBA_OK;
}

int uba_vo_get_inliers(uba_handle* h, int32_t* idx, int32_t* n_inliers) {
  uba_vo_state* s = h ? uba_vo_get(h, false) : nullptr;
  if (!s || !n_inliers) return UBA_ERR_INVALID_ARGUMENT;
  int k = 0;
  if (s->have_flags) for (int i = 0; i < s->n; i++) if (s->flags_h[i]) { if (idx) idx[k] = i; k++; }
  *n_inliers = k;
  return UBA_OK;
}

int uba_vo_refine(uba_handle* h, const double init6[6], int n_sel, const int32_t* selection, double state6[6], int32_t* converged,
                  int32_t* iterations) {
  uba_vo_state* s = h ? uba_vo_get(h, false) : nullptr;
  if (!s || !init6 || !state6) return UBA_ERR_INVALID_ARGUMENT;
  if (!s->have_matches) return uba_vo_fail(h, UBA_ERR_STATE, "uba_vo_refine", "no matches set");
  cudaStream_t st = uba_vo_stream(h);
  const unsigned char* flags = nullptr; const int32_t* dsel = nullptr;
  int count = n_sel;
  if (selection) {
    if (n_sel <= 0) return UBA_ERR_INVALID_ARGUMENT;
    int rc = vo_upload_sel(h, s, selection, (size_t)n_sel, st);
    if (rc) return rc;
    dsel = s->d_sel;
  } else {
    if (!s->have_flags) return uba_vo_fail(h, UBA_ERR_STATE, "uba_vo_refine", "no inlier set: run uba_vo_ransac first or pass a selection");
    count = 0; for (int i = 0; i < s->n; i++) count += s->flags_h[i] ? 1 : 0;
    flags = s->d_flags;
  }
  if (count < 3) { for (int i = 0; i < 6; i++) state6[i] = init6[i]; if (converged) *converged = 0; if (iterations) *iterations = 0; return UBA_OK; }   // :167-168
  VO_CU(cudaMemcpyAsync(s->d_state, init6, 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  k_vo_refine<<<1, 256, 0, st>>>(view(s), dsel, flags, count, s->d_state, s->d_state + 6, s->d_info);
  uba_vo_count(h, 1);
  int32_t info[2];
  VO_CU(cudaMemcpyAsync(state6, s->d_state + 6, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  VO_CU(cudaMemcpyAsync(info, s->d_info, sizeof(info), cudaMemcpyDeviceToHost, st));
  VO_CU(cudaStreamSynchronize(st));
  if (converged) *converged = !(info[0] == 6 || info[0] == 3);   // :279-282
  if (iterations) *iterations = info[1];
  return UBA_OK;
}

}  // extern "C"

void uba_vo_free(uba_vo_state* s) {
  if (!s) return;
  cudaFree(s->d_quads); cudaFree(s->d_pts); cudaFree(s->d_obs); cudaFree(s->d_flags); cudaFree(s->d_sel); cudaFree(s->d_ok);
  cudaFree(s->d_counts); cudaFree(s->d_info); cudaFree(s->d_states); cudaFree(s->d_state); cudaFree(s->d_scratch);
  delete s;
}
uba_vo_state* uba_vo_new() { return new uba_vo_state(); }
