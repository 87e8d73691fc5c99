// oracle/refstub/ceres/ceres.h — a MINIMAL stand-in for the slice of the Ceres Solver API that the reference's
// optimisation/BundleAdjuster.h uses, so that header can be compiled VERBATIM from /root/reference into
// oracle/_ref/libuba_ref.so (recipe: oracle/Makefile).
//
// TEST INFRASTRUCTURE.  Ceres is an un-vendored, un-versioned dependency of the reference (CMakeLists.txt:12,
// README.md:8) and is not installed here.  What this header provides, written from the published Ceres API and
// algorithms ([CERES-UPSTREAM], recalled):
//   * Jet<T,N> forward-mode dual numbers and AutoDiffCostFunction<F,M,N0,N1> — so the reference's functor templates
//     (BundleAdjuster.h:78-94,:113-130,:153-171) produce residuals AND Jacobians through their own text;
//   * LossFunction / HuberLoss / CauchyLoss, Problem (residual blocks, constant blocks, bounds), Solver::Options /
//     Summary, Solve(): a dense-algebra restatement of the trust-region Levenberg-Marquardt minimiser (Jacobi scaling,
//     LM diagonal clamps, model cost change, step acceptance, radius update, bounds projection, the termination tests)
//     — an implementation INDEPENDENT of oracle/uba_oracle.cpp (which works per point through the Schur complement);
//   * Covariance: blocks of (J^T J)^-1 over the non-constant parameter blocks.
// Small problems only (dense normal equations).  Nothing under uasl_motion_estimation_b200/ includes it.
#ifndef UBA_REFSTUB_CERES_H
#define UBA_REFSTUB_CERES_H

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace google { inline void InitGoogleLogging(const char*) {} }

namespace ceres {

// ---- forward-mode dual numbers --------------------------------------------------------------
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; i++) v[i] = T(0); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; i++) v[i] = T(0); }   // NOLINT: implicit, like ceres::Jet
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = T(0); v[k] = T(1); }
};
#define UBA_JET template <typename T, int N> inline Jet<T, N>
UBA_JET operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
UBA_JET operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
UBA_JET operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
UBA_JET operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
UBA_JET operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; const T gi = T(1) / g.a; const T q = f.a * gi; h.a = q;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - q * g.v[i]) * gi;
  return h;
}
UBA_JET operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a += s; return h; }
UBA_JET operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a += s; return h; }
UBA_JET operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a -= s; return h; }
UBA_JET operator-(T s, const Jet<T, N>& f) { Jet<T, N> h = -f; h.a += s; return h; }
UBA_JET operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * s; return h; }
UBA_JET operator*(T s, const Jet<T, N>& f) { return f * s; }
UBA_JET operator/(const Jet<T, N>& f, T s) { const T si = T(1) / s; return f * si; }
UBA_JET operator/(T s, const Jet<T, N>& g) { const T m = -s / (g.a * g.a); Jet<T, N> h; h.a = s / g.a; for (int i = 0; i < N; i++) h.v[i] = g.v[i] * m; return h; }
UBA_JET& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
UBA_JET& operator-=(Jet<T, N>& f, const Jet<T, N>& g) { f = f - g; return f; }
UBA_JET& operator+=(Jet<T, N>& f, T s) { f.a += s; return f; }
UBA_JET& operator-=(Jet<T, N>& f, T s) { f.a -= s; return f; }
UBA_JET sqrt(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T d = T(1) / (T(2) * h.a); for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h; }
UBA_JET sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; i++) h.v[i] = c * f.v[i]; return h; }
UBA_JET cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T s = -std::sin(f.a); for (int i = 0; i < N; i++) h.v[i] = s * f.v[i]; return h; }
#undef UBA_JET
template <typename T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <typename T, int N> inline bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }

// ---- cost and loss functions ----------------------------------------------------------------
class CostFunction {
 public:
  virtual ~CostFunction() {}
  // parameters: one pointer per parameter block; jacobians[i] (row-major num_residuals x block_size) may be null
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  int num_residuals() const { return num_residuals_; }
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
 protected:
  int num_residuals_ = 0;
  std::vector<int> sizes_;
};

template <typename Functor, int M, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) { num_residuals_ = M; sizes_ = {N0, N1}; }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    if (!jacobians) return (*functor_)(p[0], p[1], residuals);
    typedef Jet<double, N0 + N1> J;
    J x0[N0], x1[N1], r[M];
    for (int i = 0; i < N0; i++) x0[i] = J(p[0][i], i);
    for (int i = 0; i < N1; i++) x1[i] = J(p[1][i], N0 + i);
    if (!(*functor_)(x0, x1, r)) return false;
    for (int k = 0; k < M; k++) {
      residuals[k] = r[k].a;
      if (jacobians[0]) for (int i = 0; i < N0; i++) jacobians[0][k * N0 + i] = r[k].v[i];
      if (jacobians[1]) for (int i = 0; i < N1; i++) jacobians[1][k * N1 + i] = r[k].v[N0 + i];
    }
    return true;
  }
 private:
  std::unique_ptr<Functor> functor_;
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double s, double out[3]) const = 0;   // rho(s), rho'(s), rho''(s)
};
class TrivialLoss : public LossFunction { public: void Evaluate(double s, double o[3]) const override { o[0] = s; o[1] = 1; o[2] = 0; } };
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double o[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); o[0] = 2 * a_ * r - b_; o[1] = std::max(std::numeric_limits<double>::min(), a_ / r); o[2] = -o[1] / (2 * s); }
    else { o[0] = s; o[1] = 1; o[2] = 0; }
  }
 private:
  double a_, b_;
};
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
  void Evaluate(double s, double o[3]) const override {
    const double sum = 1 + s * c_, inv = 1 / sum;
    o[0] = b_ * std::log(sum); o[1] = std::max(std::numeric_limits<double>::min(), inv); o[2] = -c_ * (inv * inv);
  }
 private:
  double b_, c_;
};

// ---- problem --------------------------------------------------------------------------------
class Problem {
 public:
  struct Block { double* x; int size; bool constant = false; std::vector<double> lo, hi; int offset = -1; };
  struct Residual { std::unique_ptr<CostFunction> cost; std::unique_ptr<LossFunction> loss; int b0, b1; };
  Problem() {}
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
    Residual r; r.cost.reset(cost); r.loss.reset(loss);
    r.b0 = block_of(x0, cost->parameter_block_sizes()[0]); r.b1 = block_of(x1, cost->parameter_block_sizes()[1]);
    residuals_.push_back(std::move(r));
  }
  void SetParameterBlockConstant(double* x) { blocks_[index_.at(x)].constant = true; }
  void SetParameterUpperBound(double* x, int i, double v) { blocks_[index_.at(x)].hi[i] = v; }
  void SetParameterLowerBound(double* x, int i, double v) { blocks_[index_.at(x)].lo[i] = v; }
  void GetParameterBlocks(std::vector<double*>* out) const { out->clear(); for (const Block& b : blocks_) out->push_back(b.x); }
  int NumParameterBlocks() const { return (int)blocks_.size(); }
  int NumResidualBlocks() const { return (int)residuals_.size(); }
  std::vector<Block>& blocks() { return blocks_; }
  std::vector<Residual>& residual_blocks() { return residuals_; }
  int index_of(const double* x) const { auto it = index_.find(const_cast<double*>(x)); return it == index_.end() ? -1 : it->second; }
 private:
  int block_of(double* x, int size) {
    auto it = index_.find(x);
    if (it != index_.end()) return it->second;
    Block b; b.x = x; b.size = size;
    b.lo.assign(size, -std::numeric_limits<double>::max()); b.hi.assign(size, std::numeric_limits<double>::max());
    blocks_.push_back(b);
    index_[x] = (int)blocks_.size() - 1;
    return (int)blocks_.size() - 1;
  }
  std::vector<Block> blocks_;            // in order of first appearance, like ceres::Program
  std::vector<Residual> residuals_;
  std::map<double*, int> index_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

struct IterationSummary {
  int iteration = 0; bool step_is_valid = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, step_norm = 0, relative_decrease = 0, trust_region_radius = 0, model_cost_change = 0;
};

class Solver {
 public:
  struct Options {
    int max_num_iterations = 50;
    double max_solver_time_in_seconds = 1e9;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
  };
  struct Summary {
    TerminationType termination_type = FAILURE;
    std::string message;
    double initial_cost = 0, final_cost = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    std::vector<IterationSummary> iterations;
    bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE || termination_type == USER_SUCCESS; }
    std::string BriefReport() const { return message; }
    std::string FullReport() const { return message; }
  };
};

namespace internal {
// dense lower Cholesky solve of A x = b (A symmetric n x n, row-major); false when not positive definite
inline bool chol_solve(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}

struct Linearisation {          // robustified residuals and Jacobian blocks at one point
  std::vector<double> r;        // [sum M]
  std::vector<double> J0, J1;   // per residual block: M x size(b0), M x size(b1) (zeros for constant blocks)
  double cost = 0;
};

// Evaluate every residual block at the values currently stored in the parameter blocks; with_jac also fills the
// corrected Jacobians.  Corrector for rho'' <= 0 (Huber, Cauchy): r~ = sqrt(rho') r, J~ = sqrt(rho') J.
inline bool evaluate(Problem& pb, bool with_jac, Linearisation* L) {
  auto& rb = pb.residual_blocks();
  auto& bl = pb.blocks();
  size_t nr = 0, n0 = 0, n1 = 0;
  for (auto& r : rb) { const int m = r.cost->num_residuals(); nr += m; n0 += (size_t)m * bl[r.b0].size; n1 += (size_t)m * bl[r.b1].size; }
  L->r.assign(nr, 0.0);
  if (with_jac) { L->J0.assign(n0, 0.0); L->J1.assign(n1, 0.0); }
  L->cost = 0;
  size_t ro = 0, o0 = 0, o1 = 0;
  for (auto& r : rb) {
    const int m = r.cost->num_residuals(), s0 = bl[r.b0].size, s1 = bl[r.b1].size;
    const double* params[2] = {bl[r.b0].x, bl[r.b1].x};
    double* res = L->r.data() + ro;
    double* jac[2] = {with_jac ? L->J0.data() + o0 : nullptr, with_jac ? L->J1.data() + o1 : nullptr};
    if (!r.cost->Evaluate(params, res, with_jac ? jac : nullptr)) return false;
    double s = 0; for (int k = 0; k < m; k++) s += res[k] * res[k];
    double rho[3] = {s, 1, 0};
    if (r.loss) r.loss->Evaluate(s, rho);
    L->cost += 0.5 * rho[0];
    const double w = std::sqrt(rho[1]);
    for (int k = 0; k < m; k++) res[k] *= w;
    if (with_jac) {
      for (int k = 0; k < m * s0; k++) jac[0][k] = bl[r.b0].constant ? 0.0 : jac[0][k] * w;
      for (int k = 0; k < m * s1; k++) jac[1][k] = bl[r.b1].constant ? 0.0 : jac[1][k] * w;
    }
    ro += m; o0 += (size_t)m * s0; o1 += (size_t)m * s1;
  }
  return std::isfinite(L->cost);
}
}  // namespace internal

// [CERES-UPSTREAM] TrustRegionMinimizer + LevenbergMarquardtStrategy, dense algebra.  Not implemented: the
// bounds-induced projected line search (the step is projected onto the box and taken whole), inner iterations,
// non-monotonic steps.
inline void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* sum) {
  using internal::Linearisation;
  const auto t_start = std::chrono::steady_clock::now();
  Problem& pb = *problem;
  auto& bl = pb.blocks();
  auto& rb = pb.residual_blocks();
  *sum = Solver::Summary();
  int n = 0;
  for (auto& b : bl) { b.offset = b.constant ? -1 : n; if (!b.constant) n += b.size; }
  // feasibility of the bounded, non-constant blocks (Program::IsFeasible)
  for (auto& b : bl) if (!b.constant) for (int i = 0; i < b.size; i++) if (b.x[i] < b.lo[i] || b.x[i] > b.hi[i]) { sum->termination_type = FAILURE; sum->message = "infeasible start"; return; }
  std::vector<double> lo(n), hi(n), x(n), x_new(n), scale(n, 1.0), g(n), D(n), delta(n);
  for (auto& b : bl) if (!b.constant) for (int i = 0; i < b.size; i++) { lo[b.offset + i] = b.lo[i]; hi[b.offset + i] = b.hi[i]; x[b.offset + i] = b.x[i]; }
  auto store = [&](const std::vector<double>& v) { for (auto& b : bl) if (!b.constant) for (int i = 0; i < b.size; i++) b.x[i] = v[b.offset + i]; };
  Linearisation L;
  if (!internal::evaluate(pb, true, &L)) { sum->termination_type = FAILURE; sum->message = "initial evaluation failed"; return; }
  // column norms -> Jacobi scaling, once, at the initial point
  auto col_sq = [&](std::vector<double>& out) {
    std::fill(out.begin(), out.end(), 0.0);
    size_t o0 = 0, o1 = 0;
    for (auto& r : rb) {
      const int m = r.cost->num_residuals(), s0 = bl[r.b0].size, s1 = bl[r.b1].size;
      if (!bl[r.b0].constant) for (int k = 0; k < m; k++) for (int i = 0; i < s0; i++) { const double v = L.J0[o0 + k * s0 + i]; out[bl[r.b0].offset + i] += v * v; }
      if (!bl[r.b1].constant) for (int k = 0; k < m; k++) for (int i = 0; i < s1; i++) { const double v = L.J1[o1 + k * s1 + i]; out[bl[r.b1].offset + i] += v * v; }
      o0 += (size_t)m * s0; o1 += (size_t)m * s1;
    }
  };
  std::vector<double> csq(n);
  col_sq(csq);
  if (opt.jacobi_scaling) for (int j = 0; j < n; j++) scale[j] = 1.0 / (1.0 + std::sqrt(csq[j]));
  auto gradient_and_norm = [&](double* gmax) {
    std::fill(g.begin(), g.end(), 0.0);
    size_t ro = 0, o0 = 0, o1 = 0;
    for (auto& r : rb) {
      const int m = r.cost->num_residuals(), s0 = bl[r.b0].size, s1 = bl[r.b1].size;
      if (!bl[r.b0].constant) for (int k = 0; k < m; k++) for (int i = 0; i < s0; i++) g[bl[r.b0].offset + i] += L.J0[o0 + k * s0 + i] * L.r[ro + k];
      if (!bl[r.b1].constant) for (int k = 0; k < m; k++) for (int i = 0; i < s1; i++) g[bl[r.b1].offset + i] += L.J1[o1 + k * s1 + i] * L.r[ro + k];
      ro += m; o0 += (size_t)m * s0; o1 += (size_t)m * s1;
    }
    double mx = 0;   // projected gradient: || x - P(x - g) ||_inf
    for (int j = 0; j < n; j++) { const double p = std::min(hi[j], std::max(lo[j], x[j] - g[j])); mx = std::max(mx, std::fabs(x[j] - p)); }
    *gmax = mx;
  };
  double cost = L.cost, gmax = 0;
  gradient_and_norm(&gmax);
  sum->initial_cost = cost;
  { IterationSummary it; it.cost = cost; it.gradient_max_norm = gmax; it.trust_region_radius = opt.initial_trust_region_radius; it.step_is_valid = it.step_is_successful = true; sum->iterations.push_back(it); }
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  int invalid = 0;
  auto finish = [&](TerminationType t, const char* msg) { sum->termination_type = t; sum->message = msg; sum->final_cost = cost; store(x); };
  if (gmax <= opt.gradient_tolerance) { finish(CONVERGENCE, "gradient tolerance"); return; }
  std::vector<double> A, rhs;
  for (int iter = 1;; iter++) {
    if (iter > opt.max_num_iterations) { finish(NO_CONVERGENCE, "maximum number of iterations"); return; }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() >= opt.max_solver_time_in_seconds) { finish(NO_CONVERGENCE, "maximum solver time"); return; }
    if (iter > 1 && gmax <= opt.gradient_tolerance) { finish(CONVERGENCE, "gradient tolerance"); return; }
    if (radius < opt.min_trust_region_radius) { finish(CONVERGENCE, "minimum trust region radius"); return; }
    IterationSummary it; it.iteration = iter; it.trust_region_radius = radius; it.cost = cost; it.gradient_max_norm = gmax;
    // scaled normal equations  (S J^T J S + D^2) y = S J^T r,  D^2 = clamp(diag(S J^T J S)) / radius
    A.assign((size_t)n * n, 0.0); rhs.assign(n, 0.0);
    {
      size_t ro = 0, o0 = 0, o1 = 0;
      for (auto& r : rb) {
        const int m = r.cost->num_residuals(), s0 = bl[r.b0].size, s1 = bl[r.b1].size;
        const int f0 = bl[r.b0].offset, f1 = bl[r.b1].offset;
        const double* J0 = L.J0.data() + o0; const double* J1 = L.J1.data() + o1;
        for (int k = 0; k < m; k++) {
          if (f0 >= 0) for (int i = 0; i < s0; i++) {
            const double a = J0[k * s0 + i] * scale[f0 + i];
            rhs[f0 + i] += a * L.r[ro + k];
            for (int j = 0; j < s0; j++) A[(size_t)(f0 + i) * n + f0 + j] += a * J0[k * s0 + j] * scale[f0 + j];
            if (f1 >= 0) for (int j = 0; j < s1; j++) { const double v = a * J1[k * s1 + j] * scale[f1 + j]; A[(size_t)(f0 + i) * n + f1 + j] += v; A[(size_t)(f1 + j) * n + f0 + i] += v; }
          }
          if (f1 >= 0) for (int i = 0; i < s1; i++) {
            const double a = J1[k * s1 + i] * scale[f1 + i];
            rhs[f1 + i] += a * L.r[ro + k];
            for (int j = 0; j < s1; j++) A[(size_t)(f1 + i) * n + f1 + j] += a * J1[k * s1 + j] * scale[f1 + j];
          }
        }
        ro += m; o0 += (size_t)m * s0; o1 += (size_t)m * s1;
      }
    }
    std::vector<double> JtJ_diag(n);
    for (int j = 0; j < n; j++) {
      JtJ_diag[j] = A[(size_t)j * n + j];
      D[j] = std::min(std::max(JtJ_diag[j], opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
      A[(size_t)j * n + j] += D[j];
    }
    std::vector<double> Aun;   // undamped copy for the model cost change
    Aun = A; for (int j = 0; j < n; j++) Aun[(size_t)j * n + j] -= D[j];
    std::vector<double> y = rhs;
    bool ok = internal::chol_solve(A, y, n);
    double model_cost_change = 0;
    if (ok) {
      // step = -y (scaled space); model_cost_change = -step^T (g_s + 1/2 H_s step) = y^T rhs - 1/2 y^T H y
      double yr = 0, yHy = 0;
      for (int i = 0; i < n; i++) { yr += y[i] * rhs[i]; double s = 0; for (int j = 0; j < n; j++) s += Aun[(size_t)i * n + j] * y[j]; yHy += y[i] * s; }
      model_cost_change = yr - 0.5 * yHy;
      for (int i = 0; i < n; i++) if (!std::isfinite(y[i])) ok = false;
    }
    it.model_cost_change = model_cost_change;
    if (!ok || !(model_cost_change > 0.0)) {
      invalid++;
      it.step_is_valid = false;
      sum->iterations.push_back(it);
      if (invalid >= opt.max_num_consecutive_invalid_steps) { finish(FAILURE, "too many consecutive invalid steps"); return; }
      radius *= 0.5;
      continue;
    }
    invalid = 0; it.step_is_valid = true;
    double step2 = 0, x2 = 0;
    for (int j = 0; j < n; j++) {
      delta[j] = -y[j] * scale[j];
      x_new[j] = std::min(hi[j], std::max(lo[j], x[j] + delta[j]));
      step2 += (x_new[j] - x[j]) * (x_new[j] - x[j]); x2 += x[j] * x[j];
    }
    store(x_new);
    Linearisation Lc;
    const bool eval_ok = internal::evaluate(pb, false, &Lc);
    const double cand = eval_ok ? Lc.cost : std::numeric_limits<double>::max();
    it.step_norm = std::sqrt(step2);
    if (it.step_norm <= opt.parameter_tolerance * (std::sqrt(x2) + opt.parameter_tolerance)) { sum->iterations.push_back(it); finish(CONVERGENCE, "parameter tolerance"); return; }
    it.cost_change = cost - cand;
    if (std::fabs(it.cost_change) <= opt.function_tolerance * cost) { sum->iterations.push_back(it); finish(CONVERGENCE, "function tolerance"); return; }
    it.relative_decrease = it.cost_change / model_cost_change;
    if (eval_ok && it.relative_decrease > opt.min_relative_decrease) {
      it.step_is_successful = true;
      x = x_new; cost = cand;
      const double q = 2.0 * it.relative_decrease - 1.0;
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - q * q * q));
      decrease_factor = 2.0;
      sum->num_successful_steps++;
      store(x);
      internal::evaluate(pb, true, &L);
      gradient_and_norm(&gmax);
      it.cost = cost; it.gradient_max_norm = gmax;
    } else {
      store(x);
      radius /= decrease_factor; decrease_factor *= 2.0;
      sum->num_unsuccessful_steps++;
    }
    sum->iterations.push_back(it);
  }
}

// ---- covariance: blocks of (J^T J)^-1 over the non-constant blocks (unscaled, robustified Jacobian) ----
class Covariance {
 public:
  struct Options {};
  explicit Covariance(const Options&) {}
  bool Compute(const std::vector<std::pair<const double*, const double*>>& /*blocks*/, Problem* problem) {
    Problem& pb = *problem;
    auto& bl = pb.blocks(); auto& rb = pb.residual_blocks();
    n_ = 0;
    for (auto& b : bl) { b.offset = b.constant ? -1 : n_; if (!b.constant) n_ += b.size; }
    internal::Linearisation L;
    if (!internal::evaluate(pb, true, &L)) return false;
    std::vector<double> H((size_t)n_ * n_, 0.0);
    size_t o0 = 0, o1 = 0;
    for (auto& r : rb) {
      const int m = r.cost->num_residuals(), s0 = bl[r.b0].size, s1 = bl[r.b1].size, f0 = bl[r.b0].offset, f1 = bl[r.b1].offset;
      const double* J0 = L.J0.data() + o0; const double* J1 = L.J1.data() + o1;
      for (int k = 0; k < m; k++) {
        if (f0 >= 0) for (int i = 0; i < s0; i++) {
          for (int j = 0; j < s0; j++) H[(size_t)(f0 + i) * n_ + f0 + j] += J0[k * s0 + i] * J0[k * s0 + j];
          if (f1 >= 0) for (int j = 0; j < s1; j++) { const double v = J0[k * s0 + i] * J1[k * s1 + j]; H[(size_t)(f0 + i) * n_ + f1 + j] += v; H[(size_t)(f1 + j) * n_ + f0 + i] += v; }
        }
        if (f1 >= 0) for (int i = 0; i < s1; i++) for (int j = 0; j < s1; j++) H[(size_t)(f1 + i) * n_ + f1 + j] += J1[k * s1 + i] * J1[k * s1 + j];
      }
      o0 += (size_t)m * s0; o1 += (size_t)m * s1;
    }
    // invert through Cholesky, one unit vector at a time
    inv_.assign((size_t)n_ * n_, 0.0);
    std::vector<double> Lc = H, e(n_);
    {
      std::vector<double> b0(n_, 0.0); b0[0] = 1.0;
      if (n_ == 0 || !internal::chol_solve(Lc, b0, n_)) return false;
      for (int i = 0; i < n_; i++) inv_[(size_t)i * n_] = b0[i];
    }
    for (int c = 1; c < n_; c++) {
      std::fill(e.begin(), e.end(), 0.0); e[c] = 1.0;
      for (int i = 0; i < n_; i++) { double s = e[i]; for (int k = 0; k < i; k++) s -= Lc[(size_t)i * n_ + k] * e[k]; e[i] = s / Lc[(size_t)i * n_ + i]; }
      for (int i = n_ - 1; i >= 0; i--) { double s = e[i]; for (int k = i + 1; k < n_; k++) s -= Lc[(size_t)k * n_ + i] * e[k]; e[i] = s / Lc[(size_t)i * n_ + i]; }
      for (int i = 0; i < n_; i++) inv_[(size_t)i * n_ + c] = e[i];
    }
    pb_ = problem;
    return true;
  }
  bool GetCovarianceBlock(const double* p0, const double* p1, double* out) const {
    if (!pb_) return false;
    const int i0 = pb_->index_of(p0), i1 = pb_->index_of(p1);
    if (i0 < 0 || i1 < 0) return false;
    const Problem::Block& b0 = pb_->blocks()[i0]; const Problem::Block& b1 = pb_->blocks()[i1];
    for (int i = 0; i < b0.size; i++) for (int j = 0; j < b1.size; j++)
      out[i * b1.size + j] = (b0.constant || b1.constant) ? 0.0 : inv_[(size_t)(b0.offset + i) * n_ + b1.offset + j];
    return true;
  }
 private:
  Problem* pb_ = nullptr;
  int n_ = 0;
  std::vector<double> inv_;
};

}  // namespace ceres
#endif
