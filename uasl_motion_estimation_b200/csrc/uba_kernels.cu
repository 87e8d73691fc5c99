// uba_kernels.cu — sm_100a CUDA kernels of the windowed bundle-adjustment inner loop.
//
// One LM iteration is the fixed launch sequence
//   zero accumulators -> linearise (+ per-point Schur elimination) -> assemble reduced system
//   -> dense Cholesky solve (+ candidate cameras) -> back-substitution + candidate cost -> LM controller
// over ALL windows of the handle at once; the controller state lives on the device
// (WinState), so a whole uba_optimise call needs no host round trip between iterations.
//
// What the reference does inside ceres::Solve (BundleAdjuster.h:422,:469) maps to:
//   residual blocks + autodiff Jacobians + loss corrector  -> obs_linearize (uba_math.h)
//   SchurEliminator (points are the e-blocks)              -> k_lin_generic / k_lin_tile
//   LevenbergMarquardtStrategy diagonal                    -> lm_lambda, k_assemble
//   sparse Cholesky of the reduced camera matrix           -> k_chol_small / k_chol_* (dense fp64)
//   back substitution, candidate evaluation                -> k_backsub
//   TrustRegionMinimizer accept / reject / radius          -> k_lm_update
#include <cuda_runtime.h>
#ifndef UBA_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "uba_device.h"

// UBA_EMU is defined ONLY by tests/emu (a serial host emulation of the thread-independent kernels,
// used to debug host logic on a box without a GPU).  libuba.so is never built with it.
#ifdef UBA_EMU
#include "emu_launch.h"
#else
#define UBA_LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#endif

namespace uba {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// Entry (row, col), col >= row, of a window's Schur accumulator.  Windows that go to the banded solver keep only the
// band (row-major, beta + 1 entries per row: col - row <= beta by construction of beta), which is also all that a
// point-sharded run has to all-reduce; the others keep the dense n x n upper triangle.
__device__ __forceinline__ size_t sacc_index(int n, int beta, int row, int col) {
  return beta > 0 ? (size_t)row * (beta + 1) + (col - row) : (size_t)row * n + col;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

#ifndef UBA_EMU
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ bool warp_leader() { return (threadIdx.x & 31) == 0; }
#endif

// Adds v into acc[w][slot]; aggregates over the warp when every lane targets the same window.
__device__ __forceinline__ void win_add(double* acc, int stride, int w, int slot, double v, bool active) {
  const unsigned full = 0xffffffffu;
  const int w0 = __shfl_sync(full, w, 0);
  const bool uniform = __all_sync(full, w == w0);
  if (uniform) {
    const double s = warp_sum(active ? v : 0.0);
    if (warp_leader() && s != 0.0) atomicAdd(&acc[(size_t)w0 * stride + slot], s);
  } else if (active && v != 0.0) {
    atomicAdd(&acc[(size_t)w * stride + slot], v);
  }
}
__device__ __forceinline__ void win_max(double* acc, int w, double v, bool active) {
  const unsigned full = 0xffffffffu;
  const int w0 = __shfl_sync(full, w, 0);
  const bool uniform = __all_sync(full, w == w0);
  if (uniform) {
    const double s = warp_max(active ? v : 0.0);
    if (warp_leader() && s > 0.0) atomic_max_nonneg(&acc[w0], s);
  } else if (active && v > 0.0) {
    atomic_max_nonneg(&acc[w], v);
  }
}

// ---------------------------------------------------------------------------------------------
// camera prologue: R, t, G = J_l(r) for every camera of the given parity buffer
// ---------------------------------------------------------------------------------------------
__global__ void k_cam_prep(DevView V, int parity) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= V.NC) return;
  double out[kCamStride];
  cam_derive(V.cams[parity] + (size_t)c * 6, out);
  double* dst = V.camR[parity] + (size_t)c * kCamStride;
#pragma unroll
  for (int i = 0; i < kCamStride; i++) dst[i] = out[i];
}

__global__ void k_init_state(DevView V, double initial_radius) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= V.nW) return;
  WinState s;
  s.radius = initial_radius; s.decrease_factor = 2.0; s.cost = 0.0; s.gmax = 0.0; s.initial_cost = 0.0;
  s.cur = 0; s.done = 0; s.iter = 0; s.n_success = 0; s.n_unsuccess = 0; s.n_invalid = 0; s.consecutive_invalid = 0;
  s.scale_ready = 0; s.pad_[0] = 0; s.pad_[1] = 0;
  V.ws[w] = s;
  if (w == 0) *V.n_active = V.nW;
}

// ---------------------------------------------------------------------------------------------
// generic lineariser: one thread per point, fp64 global atomics for the camera-indexed sums.
// Works for any track length and any visibility pattern; the tiled kernel (k_lin_tile) is the
// fast path and this one is its fallback and the parity/debug path (DebugOut).
// ---------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(128) k_lin_generic(DevView V, DebugOut D, int only_listed) {
  constexpr int NR = (M == 4) ? 3 : 2;
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  const int count = only_listed ? V.n_gen : V.NP;
  const bool in_range = tix < count;
  const int p = only_listed ? V.gen_pts[in_range ? tix : count - 1] : (in_range ? tix : count - 1);
  int w = V.pt_win[p];
  const WinState* st = &V.ws[w];
  const bool live = in_range && st->done == 0;
  int o0 = 0, o1 = 0;
  if (live) { o0 = V.pt_obs_off[p]; o1 = V.pt_obs_off[p + 1]; }
  const bool active = live && o1 > o0;
  double cost = 0.0, gmax = 0.0, fail = 0.0;
  if (active) {
    const int cur = st->cur;
    const double radius = st->radius;
    const int cbase = V.w_cam_off[w];
    const double* camR = V.camR[cur];
    const double X[3] = {V.pts[cur][(size_t)p * 3], V.pts[cur][(size_t)p * 3 + 1], V.pts[cur][(size_t)p * 3 + 2]};
    double C6[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int o = o0; o < o1; o++) {
      const int oc = V.obs_cam[o];
      const int gc = cbase + (oc & 0x3fffffff);
      double f[M];
#pragma unroll
      for (int m = 0; m < M; m++) f[m] = V.feat[(size_t)m * V.NO + o];
      double rraw[M], wgt, F[NR][6], E[NR][3], rh[NR];
      const double rho0 = obs_linearize<M>(camR + (size_t)gc * kCamStride, X, f, (oc >> 30) & 1, V.calib, V.loss, rraw, wgt, F, E, rh);
      cost += 0.5 * rho0;
#pragma unroll
      for (int a = 0; a < NR; a++) {
        C6[0] += E[a][0] * E[a][0]; C6[1] += E[a][0] * E[a][1]; C6[2] += E[a][0] * E[a][2];
        C6[3] += E[a][1] * E[a][1]; C6[4] += E[a][1] * E[a][2]; C6[5] += E[a][2] * E[a][2];
        g[0] += E[a][0] * rh[a]; g[1] += E[a][1] * rh[a]; g[2] += E[a][2] * rh[a];
      }
      if (V.free_cam[gc] >= 0) {
        double* B = V.Bacc + (size_t)gc * 36;
        double* v = V.vacc + (size_t)gc * 6;
#pragma unroll
        for (int r = 0; r < 6; r++) {
          double vr = 0.0;
#pragma unroll
          for (int a = 0; a < NR; a++) vr += F[a][r] * rh[a];
          atomicAdd(&v[r], vr);
#pragma unroll
          for (int c = r; c < 6; c++) {
            double b = 0.0;
#pragma unroll
            for (int a = 0; a < NR; a++) b += F[a][r] * F[a][c];
            atomicAdd(&B[r * 6 + c], b);
          }
        }
      }
      if (D.residuals) {
#pragma unroll
        for (int m = 0; m < M; m++) D.residuals[(size_t)o * M + m] = rraw[m];
      }
      if (D.weights) D.weights[o] = wgt;
    }
    // Jacobi scale (captured at the first linearisation) and LM damping of the point block
    double s2[3], lam[3];
    const double Cd[3] = {C6[0], C6[3], C6[5]};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (!st->scale_ready) { s2[c] = jacobi_is2(Cd[c], V.cfg.jacobi_scaling); V.pt_is2[(size_t)p * 3 + c] = s2[c]; }
      else s2[c] = V.pt_is2[(size_t)p * 3 + c];
      lam[c] = lm_lambda_inv(Cd[c], s2[c], radius > 0.0 ? 1.0 / radius : 0.0, V.cfg.min_lm_diagonal, V.cfg.max_lm_diagonal);
    }
    if (D.C) {
      double* Cp = D.C + (size_t)p * 9;
      Cp[0] = C6[0]; Cp[1] = C6[1]; Cp[2] = C6[2]; Cp[3] = C6[1]; Cp[4] = C6[3]; Cp[5] = C6[4]; Cp[6] = C6[2]; Cp[7] = C6[4]; Cp[8] = C6[5];
    }
    if (D.grad_pts) { D.grad_pts[(size_t)p * 3] = g[0]; D.grad_pts[(size_t)p * 3 + 1] = g[1]; D.grad_pts[(size_t)p * 3 + 2] = g[2]; }
    if (D.lam_pts) { D.lam_pts[(size_t)p * 3] = lam[0]; D.lam_pts[(size_t)p * 3 + 1] = lam[1]; D.lam_pts[(size_t)p * 3 + 2] = lam[2]; }
    const double Cdamp[6] = {C6[0] + lam[0], C6[1], C6[2], C6[3] + lam[1], C6[4], C6[5] + lam[2]};
    double Li[6], h[3];
    double* rec = V.pt_rec + (size_t)p * kPtRec;
    if (!point_factor(Cdamp, Li)) {
      fail = 1.0;
#pragma unroll
      for (int i = 0; i < kPtRec; i++) rec[i] = 0.0;
    } else {
      linv_mul(Li, g, h);
#pragma unroll
      for (int i = 0; i < 6; i++) rec[i] = Li[i];
#pragma unroll
      for (int i = 0; i < 3; i++) { rec[6 + i] = h[i]; rec[9 + i] = g[i]; rec[12 + i] = lam[i]; }
      rec[15] = 0.0;
      // projected-gradient max norm, point part
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double proj = V.cfg.use_bounds ? clampd(X[c] - g[c], V.calib.lo[c], V.calib.hi[c]) : X[c] - g[c];
        gmax = fmax(gmax, fabs(X[c] - proj));
      }
      // second pass: Z = W Linv^T per free-camera observation, Schur outer products
      const int64_t red = V.w_red_off[w];
      const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
      const int sbeta = V.w_beta[w];
      double* S = V.Sacc + red;
      for (int o = o0; o < o1; o++) {
        const int oc = V.obs_cam[o];
        const int gc = cbase + (oc & 0x3fffffff);
        const int fb = V.free_cam[gc];
        if (fb < 0 && !D.W) continue;
        double f[M];
#pragma unroll
        for (int m = 0; m < M; m++) f[m] = V.feat[(size_t)m * V.NO + o];
        double rraw[M], wgt, F[NR][6], E[NR][3], rh[NR];
        obs_linearize<M>(camR + (size_t)gc * kCamStride, X, f, (oc >> 30) & 1, V.calib, V.loss, rraw, wgt, F, E, rh);
        if (D.W) {
          double* Wp = D.W + (size_t)o * 18;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
              double s = 0.0;
#pragma unroll
              for (int a = 0; a < NR; a++) s += F[a][r] * E[a][c];
              Wp[r * 3 + c] = fb >= 0 ? s : 0.0;
            }
        }
        if (fb < 0) continue;
        double Zb[18];
        obs_schur_factor<NR>(F, E, Li, Zb);
        double* zo = V.Zbuf + (size_t)o * 18;
#pragma unroll
        for (int i = 0; i < 18; i++) zo[i] = Zb[i];
        double* zh = V.zh + (size_t)gc * 6;
#pragma unroll
        for (int r = 0; r < 6; r++) atomicAdd(&zh[r], Zb[r * 3] * h[0] + Zb[r * 3 + 1] * h[1] + Zb[r * 3 + 2] * h[2]);
        for (int oa = o0; oa <= o; oa++) {
          const int gca = cbase + (V.obs_cam[oa] & 0x3fffffff);
          const int fa = V.free_cam[gca];
          if (fa < 0) continue;
          double Za[18];
          if (oa == o) {
#pragma unroll
            for (int i = 0; i < 18; i++) Za[i] = Zb[i];
          } else {
            const double* za = V.Zbuf + (size_t)oa * 18;
#pragma unroll
            for (int i = 0; i < 18; i++) Za[i] = za[i];
          }
          // block (fa, fb) with fa <= fb: observations of a point are camera-ascending
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++)
            {
              if (fa == fb && c < r) continue;   // only the upper triangle is kept
              double v = Za[r * 3] * Zb[c * 3] + Za[r * 3 + 1] * Zb[c * 3 + 1] + Za[r * 3 + 2] * Zb[c * 3 + 2];
              if (fa == fb && oa != o)  // one camera seen twice by this point: add the transposed product too
                v += Zb[r * 3] * Za[c * 3] + Zb[r * 3 + 1] * Za[c * 3 + 1] + Zb[r * 3 + 2] * Za[c * 3 + 2];
              atomicAdd(&S[sacc_index(n, sbeta, 6 * fa + r, 6 * fb + c)], v);
            }
        }
      }
    }
  }
  win_add(V.w_lin, WL_COUNT, w, WL_COST, cost, active);
  win_add(V.w_lin, WL_COUNT, w, WL_FAIL, fail, active);
  win_max(V.w_max, w, gmax, active);
}


// ---------------------------------------------------------------------------------------------
// tiled lineariser for WIDE parts (k_lin_wide: more than kSlotMaxLocal local cameras, e.g. the long tracks of c2):
// one 256-thread CTA per part (TilePart), the first generation of the tiled design.  No global atomics in the loop:
//   * every thread owns output tiles in REGISTERS for the whole part —
//       phase 1 role: thread (point-in-chunk, camera slot) owns the 6x6 camera block B and the
//                     gradient v of its slot;
//       phase 2 role: thread (camera-pair block, K-group) owns one 6x6 block of sum_j Z_a Z_b^T
//                     (and sum_j Z_a h_j on the diagonal);
//   * per chunk of points, the Schur factors Z = W L^-T (6x3 per observation) are staged in
//     shared memory and consumed by the phase-2 threads (a small SYRK over the chunk's points);
//   * one flush per part: tiles are summed across threads through shared memory and only the sums
//     go to global memory (fp64 red.global.add).
// Observations of a chunk are read with coalesced 8-byte loads from the SoA feature arrays (point-
// sorted order makes the (point, slot) thread grid hit consecutive observations).
// ---------------------------------------------------------------------------------------------
#ifndef UBA_EMU
constexpr int kZStride = 19;   // doubles per staged Z (18 + 1 pad: conflict-free 64-bit stores)
constexpr int kFlushStride = 43;

template <int M, int NT>
__global__ void __launch_bounds__(NT, kTileThreads / NT) k_lin_wide(DevView V, int first) {
  constexpr int NR = (M == 4) ? 3 : 2;
  extern __shared__ double sm[];
  const TilePart part = V.parts[first + blockIdx.x];
  const int w = part.window;
  const WinState* st = &V.ws[w];
  if (st->done) return;
  const int t = threadIdx.x;
  const int nl = part.n_local, nfx = part.n_fixed, nlf = nl - nfx;
  const int cur = st->cur;
  const double radius = st->radius;
  const bool scale_ready = st->scale_ready != 0;
  const int cbase = V.w_cam_off[w];
  // shared memory carve-up
  double* camS = sm;                                  // [kTileMaxLocal][kCamStride]
  double* CgS = camS + kTileMaxLocal * kCamStride;    // [NT][9]   per point: C (6, upper) and g (3)
  double* hS = CgS + NT * 9;                          // [NT][3]
  double* LiS = hS + NT * 3;                          // [NT][6] inverse Cholesky factor of each point block
  unsigned* maskS = reinterpret_cast<unsigned*>(LiS + NT * 6);  // [NT]
  double* big = reinterpret_cast<double*>(maskS + NT);        // union { Es [NT][9] | Zs [NT][kZStride] | flush [NT][kFlushStride] }
  __shared__ int s_free[kTileMaxLocal];               // compact index of each local slot (-1 fixed)
  __shared__ int s_gc[kTileMaxLocal];                 // global camera index of each local slot

  for (int i = t; i < nl * kCamStride; i += NT) {
    const int sl = i / kCamStride, k = i % kCamStride;
    const int gc = cbase + V.tile_cams[part.cam_list_off + sl];
    camS[i] = V.camR[cur][(size_t)gc * kCamStride + k];
    if (k == 0) { s_gc[sl] = gc; s_free[sl] = V.free_cam[gc]; }
  }
  // phase-1 role
  const int Pc = NT / nl;
  const int pl = t / nl, sl = t - pl * nl;
  const bool p1_thread = pl < Pc;
  const bool my_free = sl >= nfx;
  // phase-2 role: block (a, b), a <= b, among the nlf free slots; K-group kg
  const int nblk = nlf * (nlf + 1) / 2;
  const int G = nblk ? NT / nblk : 0;
  const int kg = nblk ? t / nblk : 0;
  const bool p2_thread = nblk && kg < G;
  int ba = 0, bb = 0;
  if (p2_thread) {
    int r = t - kg * nblk;
    while (r >= nlf - ba) { r -= nlf - ba; ba++; }
    bb = ba + r;
  }
  const bool diag = ba == bb;
  double acc[36], zacc[6], Bq[21], vq[6];
#pragma unroll
  for (int i = 0; i < 36; i++) acc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) { zacc[i] = 0.0; vq[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < 21; i++) Bq[i] = 0.0;
  double cost = 0.0, gmax = 0.0, fail = 0.0;
  __syncthreads();

  for (int c0 = part.pt_begin; c0 < part.pt_end; c0 += Pc) {
    const int np = min(Pc, part.pt_end - c0);
    // ---- phase 1a: linearise my observation -------------------------------------------------
    const bool have_pt = p1_thread && pl < np;
    const int p = c0 + pl;
    unsigned mask = 0;
    double X[3] = {0, 0, 0};
    double Wm[18];
    bool seen = false;
    if (have_pt) {
      mask = V.pt_mask[p];
      if (sl == 0) maskS[pl] = mask;
      seen = (mask >> sl) & 1u;
      if (mask) { X[0] = V.pts[cur][(size_t)p * 3]; X[1] = V.pts[cur][(size_t)p * 3 + 1]; X[2] = V.pts[cur][(size_t)p * 3 + 2]; }
    }
    if (seen && (0xff & 1)) {
      const int o = V.pt_obs_off[p] + __popc(mask & ((1u << sl) - 1u));
      double f[M];
#pragma unroll
      for (int m = 0; m < M; m++) f[m] = V.feat[(size_t)m * V.NO + o];
      const int cid = (M == 2) ? ((V.obs_cam[o] >> 30) & 1) : 0;
      double rraw[M], wgt, F[NR][6], E[NR][3], rh[NR];
      const double rho0 = obs_linearize<M>(camS + sl * kCamStride, X, f, cid, V.calib, V.loss, rraw, wgt, F, E, rh);
      cost += 0.5 * rho0;
      double* es = big + t * 9;
      double c6[6] = {0, 0, 0, 0, 0, 0}, g3[3] = {0, 0, 0};
#pragma unroll
      for (int a = 0; a < NR; a++) {
        c6[0] = fma(E[a][0], E[a][0], c6[0]); c6[1] = fma(E[a][0], E[a][1], c6[1]); c6[2] = fma(E[a][0], E[a][2], c6[2]);
        c6[3] = fma(E[a][1], E[a][1], c6[3]); c6[4] = fma(E[a][1], E[a][2], c6[4]); c6[5] = fma(E[a][2], E[a][2], c6[5]);
        g3[0] = fma(E[a][0], rh[a], g3[0]); g3[1] = fma(E[a][1], rh[a], g3[1]); g3[2] = fma(E[a][2], rh[a], g3[2]);
      }
#pragma unroll
      for (int i = 0; i < 6; i++) es[i] = c6[i];
#pragma unroll
      for (int i = 0; i < 3; i++) es[6 + i] = g3[i];
      if (my_free) {
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
          for (int a = 0; a < NR; a++) vq[r] = fma(F[a][r], rh[a], vq[r]);
#pragma unroll
          for (int c = r; c < 6; c++) {
#pragma unroll
            for (int a = 0; a < NR; a++) Bq[q] = fma(F[a][r], F[a][c], Bq[q]);
            q++;
          }
        }
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double sacc = 0.0;
#pragma unroll
            for (int a = 0; a < NR; a++) sacc = fma(F[a][r], E[a][c], sacc);
            Wm[r * 3 + c] = sacc;
          }
      }
    }
    __syncthreads();
    // ---- phase 1b: per-point sums of E^T E and E^T r over the point's observations ---------------
    if (0xff & 2)
    for (int idx = t; idx < np * 9; idx += NT) {
      const int q = idx / 9, e = idx - q * 9;
      unsigned m = maskS[q];
      double sacc = 0.0;
      while (m) {
        const int s2i = __ffs(m) - 1;
        m &= m - 1;
        sacc += big[(q * nl + s2i) * 9 + e];
      }
      CgS[idx] = sacc;
    }
    __syncthreads();
    // ---- phase 1c: one thread per point: damping, 3x3 factor, h = L^-1 g, point record -----------
    if (t < np && (0xff & 4)) {
      const unsigned pm = maskS[t];
      if (pm) {
        const int pp = c0 + t;
        const double* cg = CgS + t * 9;
        const double Cd[3] = {cg[0], cg[3], cg[5]};
        const double Xp[3] = {V.pts[cur][(size_t)pp * 3], V.pts[cur][(size_t)pp * 3 + 1], V.pts[cur][(size_t)pp * 3 + 2]};
        double s2[3], lam[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          s2[c] = scale_ready ? V.pt_is2[(size_t)pp * 3 + c] : jacobi_is2(Cd[c], V.cfg.jacobi_scaling);
          lam[c] = lm_lambda_inv(Cd[c], s2[c], radius > 0.0 ? 1.0 / radius : 0.0, V.cfg.min_lm_diagonal, V.cfg.max_lm_diagonal);
        }
        const double Cdamp[6] = {cg[0] + lam[0], cg[1], cg[2], cg[3] + lam[1], cg[4], cg[5] + lam[2]};
        const double g[3] = {cg[6], cg[7], cg[8]};
        double Li[6] = {0, 0, 0, 0, 0, 0}, h[3] = {0, 0, 0};
        const bool ok = point_factor(Cdamp, Li);
        double* rec = V.pt_rec + (size_t)pp * kPtRec;
        if (!ok) {
          fail += 1.0;
#pragma unroll
          for (int i = 0; i < 6; i++) Li[i] = 0.0;
#pragma unroll
          for (int i = 0; i < kPtRec; i++) rec[i] = 0.0;
        } else {
          linv_mul(Li, g, h);
#pragma unroll
          for (int i = 0; i < 6; i++) rec[i] = Li[i];
#pragma unroll
          for (int i = 0; i < 3; i++) { rec[6 + i] = h[i]; rec[9 + i] = g[i]; rec[12 + i] = lam[i]; }
          rec[15] = 0.0;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double proj = V.cfg.use_bounds ? clampd(Xp[c] - g[c], V.calib.lo[c], V.calib.hi[c]) : Xp[c] - g[c];
            gmax = fmax(gmax, fabs(Xp[c] - proj));
          }
        }
        if (!scale_ready) { V.pt_is2[(size_t)pp * 3] = s2[0]; V.pt_is2[(size_t)pp * 3 + 1] = s2[1]; V.pt_is2[(size_t)pp * 3 + 2] = s2[2]; }
#pragma unroll
        for (int i = 0; i < 6; i++) LiS[t * 6 + i] = Li[i];
        hS[t * 3] = h[0]; hS[t * 3 + 1] = h[1]; hS[t * 3 + 2] = h[2];
      }
    }
    __syncthreads();
    // ---- phase 1d: Z = W L^-T for my observation (zero if the point block was not positive definite)
    if (seen && my_free && (0xff & 8)) {
      const double* Li = LiS + pl * 6;
      const double l0 = Li[0], l1 = Li[1], l2 = Li[2], l3 = Li[3], l4 = Li[4], l5 = Li[5];
      double* z = big + (size_t)(pl * nlf + (sl - nfx)) * kZStride;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        // Z[r][m] = sum_{c<=m} W[r][c] Linv[m][c]
        z[r * 3 + 0] = Wm[r * 3] * l0;
        z[r * 3 + 1] = fma(Wm[r * 3], l1, Wm[r * 3 + 1] * l2);
        z[r * 3 + 2] = fma(Wm[r * 3], l3, fma(Wm[r * 3 + 1], l4, Wm[r * 3 + 2] * l5));
      }
    }
    __syncthreads();
    // ---- phase 2: camera-pair blocks, acc += Z_a Z_b^T over my K-group's points ------------------
    if (p2_thread && (0xff & 16)) {
      for (int q = kg; q < np; q += G) {
        const unsigned m = maskS[q] >> nfx;
        if (!((m >> ba) & 1u) || !((m >> bb) & 1u)) continue;
        const double* za = big + (size_t)(q * nlf + ba) * kZStride;
        const double* zb = big + (size_t)(q * nlf + bb) * kZStride;
        double A[18];
#pragma unroll
        for (int i = 0; i < 18; i++) A[i] = za[i];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          const double b0 = zb[c * 3], b1 = zb[c * 3 + 1], b2 = zb[c * 3 + 2];
#pragma unroll
          for (int r = 0; r < 6; r++) acc[r * 6 + c] = fma(A[r * 3], b0, fma(A[r * 3 + 1], b1, fma(A[r * 3 + 2], b2, acc[r * 6 + c])));
        }
        if (diag) {
          const double h0 = hS[q * 3], h1 = hS[q * 3 + 1], h2 = hS[q * 3 + 2];
#pragma unroll
          for (int r = 0; r < 6; r++) zacc[r] = fma(A[r * 3], h0, fma(A[r * 3 + 1], h1, fma(A[r * 3 + 2], h2, zacc[r])));
        }
      }
    }
    __syncthreads();
  }

  // ---- flush: Schur tiles --------------------------------------------------------------------
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* S = V.Sacc + V.w_red_off[w];
  if (p2_thread) {
    double* o = big + (size_t)t * kFlushStride;
#pragma unroll
    for (int i = 0; i < 36; i++) o[i] = acc[i];
#pragma unroll
    for (int i = 0; i < 6; i++) o[36 + i] = zacc[i];
  }
  __syncthreads();
  for (int idx = t; idx < nblk * 42; idx += NT) {
    const int blk = idx / 42, e = idx - blk * 42;
    double sacc = 0.0;
    for (int k2 = 0; k2 < G; k2++) sacc += big[(size_t)(k2 * nblk + blk) * kFlushStride + e];
    int a = 0, r = blk;
    while (r >= nlf - a) { r -= nlf - a; a++; }
    const int b = a + r;
    const int fa = s_free[nfx + a], fb = s_free[nfx + b];
    if (e < 36) {
      const int row = 6 * fa + e / 6, colx = 6 * fb + e % 6;
      if (sacc != 0.0 && colx >= row) atomicAdd(&S[sacc_index(n, V.w_beta[w], row, colx)], sacc);
    } else if (a == b) {
      if (sacc != 0.0) atomicAdd(&V.zh[(size_t)s_gc[nfx + a] * 6 + (e - 36)], sacc);
    }
  }
  __syncthreads();
  // ---- flush: camera blocks B and gradients v ----------------------------------------------------
  if (p1_thread && my_free) {
    double* o = big + (size_t)t * 27;
#pragma unroll
    for (int i = 0; i < 21; i++) o[i] = Bq[i];
#pragma unroll
    for (int i = 0; i < 6; i++) o[21 + i] = vq[i];
  }
  __syncthreads();
  for (int idx = t; idx < nlf * 27; idx += NT) {
    const int s2i = nfx + idx / 27, e = idx % 27;
    double sacc = 0.0;
    for (int q = 0; q < Pc; q++) sacc += big[(size_t)(q * nl + s2i) * 27 + e];
    if (sacc == 0.0) continue;
    const int gc = s_gc[s2i];
    if (e < 21) {
      int r = 0, k2 = e;
      while (k2 >= 6 - r) { k2 -= 6 - r; r++; }
      atomicAdd(&V.Bacc[(size_t)gc * 36 + r * 6 + r + k2], sacc);
    } else {
      atomicAdd(&V.vacc[(size_t)gc * 6 + (e - 21)], sacc);
    }
  }
  // ---- per-window scalars ------------------------------------------------------------------------
  cost = warp_sum(cost); fail = warp_sum(fail); gmax = warp_max(gmax);
  if (warp_leader()) {
    if (cost != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_COST], cost);
    if (fail != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_FAIL], fail);
    if (gmax > 0.0) atomic_max_nonneg(&V.w_max[w], gmax);
  }
}
#endif  // !UBA_EMU


// ---------------------------------------------------------------------------------------------
// tiled lineariser, second generation (k_lin_tile2): one CTA per part (TilePart), lane = one observation
// (a point's slot lanes sit in one warp), no global atomics in the loop, one flush per part.
//   * the per-point sums are reduced with a segmented shuffle tree and the 3x3 block is factored redundantly by
//     every lane of the point;
//   * the Schur products  sum_j Z_j Z_j^T  run on the FP64 tensor-core path
//     (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4): the chunk's Schur factors are staged as ONE
//     matrix Zm[6 nlf (padded to 8 T)][3 Pc] in shared memory and every warp accumulates its
//     8x8 output tiles (2 doubles per lane per tile) over its share of the K = 3 Pc columns.
//     One DMMA replaces 8 warp-level DFMAs and ~3 LDS, so phase 2 needs 1/7 of the issue slots and
//     a fraction of the registers and shared-memory traffic of the DFMA version.
//   * Z h (the rhs term) is accumulated by the phase-1 threads next to B and v.
// Peak fp64 throughput of DMMA equals DFMA's on B200 (37 TFLOP/s, tests/cuda/dmma_bench.cu); the
// win is latency hiding, not a higher ceiling.
// ---------------------------------------------------------------------------------------------
#ifndef UBA_EMU
// experiment hook (scripts/tile_phases.py): phases of k_lin_tile2 can be compiled out one by one
#ifndef UBA_TILE_PHASES
#define UBA_TILE_PHASES 0xff
#endif
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// asynchronous global -> shared copies (LDGSTS): the prefetch of the next chunk holds no registers
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(__cvta_generic_to_global(gsrc)) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(__cvta_generic_to_global(gsrc)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kT2MaxPc = 64;          // points per chunk are capped so that K = 3 Pc <= 192
constexpr int kT2StageD = 8;          // per lane and buffer: point (3) + features (<= 4) doubles, 1 spare
constexpr int kT2StageI = 4;          // ... and ints: next chunk's mask, its first-observation offset, obs_cam word, spare
constexpr int kCamSm = 13;            // per-slot camera record in shared memory: R[9] t[3] small-angle flag; odd stride = slots in distinct banks
// Zm capacity in doubles: the worst case over (T, nl) of 8 T * t2_ldz(Pc) is 96 x 84 (T = 12, Pc = 23) for
// 256-thread CTAs and 32 x 196 (T = 4, Pc = 64) for 128-thread CTAs
#ifndef UBA_T2_ZM128
#define UBA_T2_ZM128 6400
#endif
__host__ __device__ constexpr int t2_zm_doubles(int nt) { return nt == 256 ? 8192 : UBA_T2_ZM128; }
// doubles of the region {Zm | flush scratch} (the scratch aliases Zm): [points x slots][33] + [slots][33] <= (nt + 21) * 33, or the
// (8T) x (8T+1) reduction / transform tile of the Schur blocks (T <= 16 for 256-thread CTAs, <= 8 for 128-thread ones)
__host__ __device__ constexpr int t2_main_doubles(int nt) {
  return nt == 256 ? 128 * 129 : (UBA_T2_ZM128 > 128 * 33 + 21 * 33 ? (UBA_T2_ZM128 > 64 * 65 ? UBA_T2_ZM128 : 64 * 65) : 128 * 33 + 21 * 33);
}
// ... followed by the two prefetch staging buffers
__host__ __device__ constexpr int t2_stage_doubles(int nt) { return 2 * nt * (kT2StageD + kT2StageI / 2); }

__host__ __device__ constexpr int t2_ldz(int pc) {
  int k = ((3 * pc + 3) / 4) * 4;     // K padded to the DMMA k = 4
  while ((k & 15) != 4) k += 4;       // row stride == 4 (mod 16) doubles: conflict-free fragment loads
  return k;
}

// index of entry (r, c), r <= c, in the packed upper triangle of a 6x6 block
__host__ __device__ constexpr int ut6(int r, int c) { return r * 6 - r * (r - 1) / 2 + (c - r); }

template <int M, int NT, int T, int TG>
__device__ __forceinline__ void tile2_part(const DevView& V, const TilePart& part, double* sm) {
  constexpr int W = NT / 32;
  constexpr int KG = W / TG;                       // K-groups
  constexpr int NTILES = T * (T + 1) / 2;
  constexpr int TPW = (NTILES + TG - 1) / TG;      // tiles per warp
  const unsigned FULL = 0xffffffffu;
  const int w = part.window;
  const WinState* st = &V.ws[w];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nl = part.n_local, nfx = part.n_fixed, nlf = nl - nfx;
  const int cur = st->cur;
  const double radius = st->radius;
  const double inv_radius = radius > 0.0 ? 1.0 / radius : 0.0;
  const bool scale_ready = st->scale_ready != 0;
  const int cbase = V.w_cam_off[w];
  // phase-1 role, warp aligned: a point's nl slot lanes sit in ONE warp, so its 3x3 block is reduced with
  // shuffles and factored redundantly by every lane of the point (no shared staging, no idle threads)
  const int ppw = min(32 / nl, kT2MaxPc / W);      // points per warp
  const int Pc = W * ppw;                          // points per chunk (one observation per lane)
  const int plw = lane / nl, sl = lane - plw * nl;
  const bool p1_thread = plw < ppw;
  const int pl = warp * ppw + plw;                 // point slot in the chunk
  const int seg_base = plw * nl;                   // first lane of my point
  const bool my_free = sl >= nfx;
  const int ldz = t2_ldz(Pc);
  const int ksteps = (3 * Pc + 3) / 4;
  // shared memory carve-up (doubles)
  double* camS = sm;                                 // [kTileMaxLocal][kCamSm]
  double* Zm = camS + kTileMaxLocal * kCamSm;        // [8 T][ldz]
  double* scratch = Zm;                              // flush scratch aliases Zm
  __shared__ int s_free[kTileMaxLocal];
  __shared__ int s_gc[kTileMaxLocal];

  for (int i = t; i < nl * kCamSm; i += NT) {
    const int sl2 = i / kCamSm, k = i - sl2 * kCamSm;
    const int gc = cbase + V.tile_cams[part.cam_list_off + sl2];
    camS[i] = V.camR[cur][(size_t)gc * kCamStride + (k < 12 ? k : 21)];   // R (9), t (3), small-angle flag
    if (k == 0) { s_gc[sl2] = gc; s_free[sl2] = V.free_cam[gc]; }
  }
  for (int i = t; i < 8 * T * ldz; i += NT) Zm[i] = 0.0;   // padding rows / columns stay zero for the whole part
  // phase-2 role: warp -> (tile group g, K-group kq); my tiles t = g + i*TG, enumerated (I <= J) row by row
  const int g = warp % TG, kq = warp / TG;
  int tI[TPW], tJ[TPW];
#pragma unroll
  for (int i = 0; i < TPW; i++) {
    int idx = g + i * TG, I = 0;
    if (idx >= NTILES) { tI[i] = -1; tJ[i] = 0; continue; }
    while (idx >= T - I) { idx -= T - I; I++; }
    tI[i] = I; tJ[i] = I + idx;
  }
  double acc[TPW][2];
#pragma unroll
  for (int i = 0; i < TPW; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  // Camera-side sums of my slot in the frame where the angle-axis part has NOT yet been multiplied by G = J_l(r):
  //   F = d [I | -[u]x G]  =>  B = D B' D^T, v = D v', Z h = D (Z' h'),  D = diag(I, G^T)   (applied once per part, at the flush)
  //   B'tt = sum Q,  B'tr = -sum Q [u]x =: -sum U,  B'rr = sum [u]x^T U     (Q = d^T d, Q01 = 0)
  double Bt[5], Bu[9], Br[6], vq[6], zq[6];
#pragma unroll
  for (int i = 0; i < 5; i++) Bt[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) Bu[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) { Br[i] = 0.0; vq[i] = 0.0; zq[i] = 0.0; }
  double cost = 0.0, gmax = 0.0, fail = 0.0;
  __syncthreads();

  // Prefetch pipeline over chunks, two implementations (chosen per tile variant by measurement on B200):
  //  * kAsyncPipe (the T = 4 variant: c4, c5): through shared memory with cp.async, no registers held across a chunk.  While
  //    chunk k is processed, the point and features of this lane's observation in chunk k+1 and the mask / first-observation
  //    offset of its point in chunk k+2 are in flight into the lane's staging slot (two buffers, alternating).
  //  * otherwise (wider variants: c1, c2, c3) in registers: stage B holds the mask / offset of the chunk after next, stage A
  //    the mask, point and features of the next chunk.
  constexpr bool kAsyncPipe = (T == 4);
  double* stageD = Zm + t2_main_doubles(NT);                       // [2][NT][kT2StageD]
  int* stageI = reinterpret_cast<int*>(stageD + 2 * NT * kT2StageD);   // [2][NT][kT2StageI]
  unsigned mask_cur = 0;
  // issue the copies for the chunk whose point is pa (mask m, offset off known) into buffer b, plus the mask / offset
  // of the chunk after it
  auto prefetch = [&](int b, int pa, unsigned m, int off) {
    double* sd = stageD + ((size_t)b * NT + t) * kT2StageD;
    int* si = stageI + ((size_t)b * NT + t) * kT2StageI;
    if (m) {
      const double* px = V.pts[cur] + (size_t)pa * 3;
      cp_async8(sd, px); cp_async8(sd + 1, px + 1); cp_async8(sd + 2, px + 2);
      if ((m >> sl) & 1u) {
        const int o = off + __popc(m & ((1u << sl) - 1u));
#pragma unroll
        for (int q = 0; q < M; q++) cp_async8(sd + 3 + q, V.feat + (size_t)q * V.NO + o);
        if (M == 2) cp_async4(si + 2, V.obs_cam + o);
      }
    }
    const int pb = pa + Pc;
    if (pb < part.pt_end) { cp_async4(si, V.pt_mask + pb); cp_async4(si + 1, V.pt_obs_off + pb); }
    else si[0] = 0;
    cp_async_commit();
  };
  unsigned maskA = 0, maskB = 0;
  int oA = 0, offB = 0;
  double XA[3] = {0, 0, 0}, fA[M];
  int cidA = 0;
#pragma unroll
  for (int m = 0; m < M; m++) fA[m] = 0.0;
  if constexpr (kAsyncPipe) {
    if (p1_thread) {
      const int pa = part.pt_begin + pl;
      int off0 = 0;
      if (pa < part.pt_end) { mask_cur = V.pt_mask[pa]; off0 = V.pt_obs_off[pa]; }
      prefetch(0, pa, mask_cur, off0);
    }
  } else {
    const int pa = part.pt_begin + pl, pb = pa + Pc;
    if (p1_thread && pa < part.pt_end) {
      maskA = V.pt_mask[pa];
      if (maskA) {
        XA[0] = V.pts[cur][(size_t)pa * 3]; XA[1] = V.pts[cur][(size_t)pa * 3 + 1]; XA[2] = V.pts[cur][(size_t)pa * 3 + 2];
        if ((maskA >> sl) & 1u) {
          oA = V.pt_obs_off[pa] + __popc(maskA & ((1u << sl) - 1u));
#pragma unroll
          for (int m = 0; m < M; m++) fA[m] = V.feat[(size_t)m * V.NO + oA];
          if (M == 2) cidA = (V.obs_cam[oA] >> 30) & 1;
        }
      }
    }
    if (p1_thread && pb < part.pt_end) { maskB = V.pt_mask[pb]; offB = V.pt_obs_off[pb]; }
  }
  int buf = 0;
  for (int c0 = part.pt_begin; c0 < part.pt_end; c0 += Pc, buf ^= 1) {
    const int np = min(Pc, part.pt_end - c0);
    // ---- phase 1a: linearise my observation -------------------------------------------------
    const bool have_pt = p1_thread && pl < np;
    const int p = c0 + pl;
    const unsigned mask = kAsyncPipe ? mask_cur : maskA;
    const bool seen = have_pt && ((mask >> sl) & 1u);
    double X[3] = {0, 0, 0}, f[M];
#pragma unroll
    for (int q = 0; q < M; q++) f[q] = 0.0;
    int cid = 0;
    if constexpr (kAsyncPipe) {
      if (p1_thread) {
        cp_async_wait_all();
        const double* sd = stageD + ((size_t)buf * NT + t) * kT2StageD;
        const int* si = stageI + ((size_t)buf * NT + t) * kT2StageI;
        if (mask) { X[0] = sd[0]; X[1] = sd[1]; X[2] = sd[2]; }
        if (seen) {
#pragma unroll
          for (int q = 0; q < M; q++) f[q] = sd[3 + q];
          if (M == 2) cid = (si[2] >> 30) & 1;
        }
        const unsigned mask_next = (unsigned)si[0];
        const int off_next = si[1];
        prefetch(buf ^ 1, p + Pc, mask_next, off_next);
        mask_cur = mask_next;
      }
    } else {
      X[0] = XA[0]; X[1] = XA[1]; X[2] = XA[2];
#pragma unroll
      for (int m = 0; m < M; m++) f[m] = fA[m];
      cid = cidA;
      // advance the pipeline: A <- next chunk (using stage B's mask / offset), B <- the chunk after next
      const int pa = p + Pc, pb = pa + Pc;
      maskA = maskB;
      const int offA = offB;
      maskB = 0;
      if (p1_thread && pb < part.pt_end) { maskB = V.pt_mask[pb]; offB = V.pt_obs_off[pb]; }
      if (maskA) {
        XA[0] = V.pts[cur][(size_t)pa * 3]; XA[1] = V.pts[cur][(size_t)pa * 3 + 1]; XA[2] = V.pts[cur][(size_t)pa * 3 + 2];
        if ((maskA >> sl) & 1u) {
          oA = offA + __popc(maskA & ((1u << sl) - 1u));
#pragma unroll
          for (int m = 0; m < M; m++) fA[m] = V.feat[(size_t)m * V.NO + oA];
          if (M == 2) cidA = (V.obs_cam[oA] >> 30) & 1;
        }
      }
    }
    double Pm[9], u[3] = {0, 0, 0};             // P = Q R: the translation rows of W' = F'^T E; W'rot = [u]x P
    double cg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; i++) Pm[i] = 0.0;
    if (seen && (UBA_TILE_PHASES & 1)) {
      const double* R = camS + sl * kCamSm;
      double Q[5], m3[3];
      const double rho0 = obs_linearize_q<M>(R, R + 9, R[12] != 0.0, X, f, cid, V.calib, V.loss, Q, m3, u);
      cost += 0.5 * rho0;
      const double Q00 = Q[0], Q02 = Q[1], Q11 = Q[2], Q12 = Q[3], Q22 = Q[4];
      // P = Q R (Q01 = 0)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        Pm[c] = fma(Q00, R[c], Q02 * R[6 + c]);
        Pm[3 + c] = fma(Q11, R[3 + c], Q12 * R[6 + c]);
        Pm[6 + c] = fma(Q02, R[c], fma(Q12, R[3 + c], Q22 * R[6 + c]));
      }
      // E^T E = R^T P (upper), E^T r = R^T m
      cg[0] = fma(R[0], Pm[0], fma(R[3], Pm[3], R[6] * Pm[6]));
      cg[1] = fma(R[0], Pm[1], fma(R[3], Pm[4], R[6] * Pm[7]));
      cg[2] = fma(R[0], Pm[2], fma(R[3], Pm[5], R[6] * Pm[8]));
      cg[3] = fma(R[1], Pm[1], fma(R[4], Pm[4], R[7] * Pm[7]));
      cg[4] = fma(R[1], Pm[2], fma(R[4], Pm[5], R[7] * Pm[8]));
      cg[5] = fma(R[2], Pm[2], fma(R[5], Pm[5], R[8] * Pm[8]));
      cg[6] = fma(R[0], m3[0], fma(R[3], m3[1], R[6] * m3[2]));
      cg[7] = fma(R[1], m3[0], fma(R[4], m3[1], R[7] * m3[2]));
      cg[8] = fma(R[2], m3[0], fma(R[5], m3[1], R[8] * m3[2]));
      if (my_free) {
        const double ux = u[0], uy = u[1], uz = u[2];
        Bt[0] += Q00; Bt[1] += Q02; Bt[2] += Q11; Bt[3] += Q12; Bt[4] += Q22;
        // U = Q [u]x
        const double U00 = -Q02 * uy, U01 = fma(Q02, ux, -Q00 * uz), U02 = Q00 * uy;
        const double U10 = fma(Q11, uz, -Q12 * uy), U11 = Q12 * ux, U12 = -Q11 * ux;
        const double U20 = fma(Q12, uz, -Q22 * uy), U21 = fma(Q22, ux, -Q02 * uz), U22 = fma(Q02, uy, -Q12 * ux);
        Bu[0] += U00; Bu[1] += U01; Bu[2] += U02; Bu[3] += U10; Bu[4] += U11; Bu[5] += U12; Bu[6] += U20; Bu[7] += U21; Bu[8] += U22;
        // [u]x^T U, upper triangle
        Br[0] = fma(uz, U10, fma(-uy, U20, Br[0]));
        Br[1] = fma(uz, U11, fma(-uy, U21, Br[1]));
        Br[2] = fma(uz, U12, fma(-uy, U22, Br[2]));
        Br[3] = fma(ux, U21, fma(-uz, U01, Br[3]));
        Br[4] = fma(ux, U22, fma(-uz, U02, Br[4]));
        Br[5] = fma(uy, U02, fma(-ux, U12, Br[5]));
        vq[0] += m3[0]; vq[1] += m3[1]; vq[2] += m3[2];
        vq[3] = fma(uy, m3[2], fma(-uz, m3[1], vq[3]));
        vq[4] = fma(uz, m3[0], fma(-ux, m3[2], vq[4]));
        vq[5] = fma(ux, m3[1], fma(-uy, m3[0], vq[5]));
      }
    }
    // ---- phase 1b: sum E^T E / E^T r over the point's slot lanes (segmented shuffle tree + broadcast) ----
    if (UBA_TILE_PHASES & 4) {
      for (int off = 1; off < nl; off <<= 1) {
        const bool take = sl + off < nl;
#pragma unroll
        for (int e = 0; e < 9; e++) {
          const double o = __shfl_down_sync(FULL, cg[e], off);
          if (take) cg[e] += o;
        }
      }
#pragma unroll
      for (int e = 0; e < 9; e++) cg[e] = __shfl_sync(FULL, cg[e], seg_base);
    }
    // ---- phase 1c: damping + 3x3 factor, redundantly in every lane of the point ------------------------
    double Li[6] = {0, 0, 0, 0, 0, 0}, h[3] = {0, 0, 0};
    if (have_pt && mask && (UBA_TILE_PHASES & 4)) {
      const double Cd[3] = {cg[0], cg[3], cg[5]};
      double s2[3], lam[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        s2[c] = scale_ready ? V.pt_is2[(size_t)p * 3 + c] : jacobi_is2(Cd[c], V.cfg.jacobi_scaling);
        lam[c] = lm_lambda_inv(Cd[c], s2[c], inv_radius, V.cfg.min_lm_diagonal, V.cfg.max_lm_diagonal);
      }
      const double Cdamp[6] = {cg[0] + lam[0], cg[1], cg[2], cg[3] + lam[1], cg[4], cg[5] + lam[2]};
      const double gg[3] = {cg[6], cg[7], cg[8]};
      const bool ok = point_factor(Cdamp, Li);
      if (ok) linv_mul(Li, gg, h);
      else {
#pragma unroll
        for (int i = 0; i < 6; i++) Li[i] = 0.0;
      }
      if (sl == (int)__ffs(mask) - 1) {        // the point's first observing lane writes its record
        double* rec = V.pt_rec + (size_t)p * kPtRec;
        if (!ok) {
          fail += 1.0;
#pragma unroll
          for (int i = 0; i < kPtRec; i++) rec[i] = 0.0;
        } else {
#pragma unroll
          for (int i = 0; i < 6; i++) rec[i] = Li[i];
#pragma unroll
          for (int i = 0; i < 3; i++) { rec[6 + i] = h[i]; rec[9 + i] = gg[i]; rec[12 + i] = lam[i]; }
          rec[15] = 0.0;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double proj = V.cfg.use_bounds ? clampd(X[c] - gg[c], V.calib.lo[c], V.calib.hi[c]) : X[c] - gg[c];
            gmax = fmax(gmax, fabs(X[c] - proj));
          }
        }
        if (!scale_ready) { V.pt_is2[(size_t)p * 3] = s2[0]; V.pt_is2[(size_t)p * 3 + 1] = s2[1]; V.pt_is2[(size_t)p * 3 + 2] = s2[2]; }
      }
    }
    // ---- phase 1d: Z' = W' L^-T into the chunk matrix Zm (zeros where the point does not see my slot) ----
    //      rows 0..2: P L^-T, rows 3..5: [u]x (P L^-T)
    if (p1_thread && my_free && (UBA_TILE_PHASES & 8)) {
      double* zc = Zm + (size_t)(6 * (sl - nfx)) * ldz + 3 * pl;
      double zt[3][3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        zt[r][0] = Pm[r * 3] * Li[0];
        zt[r][1] = fma(Pm[r * 3], Li[1], Pm[r * 3 + 1] * Li[2]);
        zt[r][2] = fma(Pm[r * 3], Li[3], fma(Pm[r * 3 + 1], Li[4], Pm[r * 3 + 2] * Li[5]));
      }
      // (not seen: Pm = 0, u = 0 -> zeros, which is what the slot must hold)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        zc[c] = zt[0][c]; zc[(size_t)ldz + c] = zt[1][c]; zc[(size_t)2 * ldz + c] = zt[2][c];
        zc[(size_t)3 * ldz + c] = fma(u[1], zt[2][c], -u[2] * zt[1][c]);
        zc[(size_t)4 * ldz + c] = fma(u[2], zt[0][c], -u[0] * zt[2][c]);
        zc[(size_t)5 * ldz + c] = fma(u[0], zt[1][c], -u[1] * zt[0][c]);
      }
      // Z' h: translation rows y = (P L^-T) h, rotation rows u x y
      const double y0 = fma(zt[0][0], h[0], fma(zt[0][1], h[1], zt[0][2] * h[2]));
      const double y1 = fma(zt[1][0], h[0], fma(zt[1][1], h[1], zt[1][2] * h[2]));
      const double y2 = fma(zt[2][0], h[0], fma(zt[2][1], h[1], zt[2][2] * h[2]));
      zq[0] += y0; zq[1] += y1; zq[2] += y2;
      zq[3] = fma(u[1], y2, fma(-u[2], y1, zq[3]));
      zq[4] = fma(u[2], y0, fma(-u[0], y2, zq[4]));
      zq[5] = fma(u[0], y1, fma(-u[1], y0, zq[5]));
    }
    __syncthreads();
    // ---- phase 2: tensor-core SYRK over the chunk: acc(I,J) += Zm[8I.., k] Zm[8J.., k]^T ------------
    if (nlf > 0 && (UBA_TILE_PHASES & 16)) {
      const int frow = lane >> 2, fk = lane & 3;
      for (int ks = kq; ks < ksteps; ks += KG) {
        const double* col = Zm + 4 * ks + fk;
        if (TG == 1) {
          double fr[T];
#pragma unroll
          for (int I = 0; I < T; I++) fr[I] = col[(size_t)(8 * I + frow) * ldz];
          int i = 0;
#pragma unroll
          for (int I = 0; I < T; I++)
#pragma unroll
            for (int J = I; J < T; J++) { dmma884(acc[i][0], acc[i][1], fr[I], fr[J]); i++; }
        } else {
#pragma unroll
          for (int i = 0; i < TPW; i++) {
            if (tI[i] >= 0) {
              const double a = col[(size_t)(8 * tI[i] + frow) * ldz];
              const double b = col[(size_t)(8 * tJ[i] + frow) * ldz];
              dmma884(acc[i][0], acc[i][1], a, b);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- flush: Schur tiles.  K-groups are summed in shared memory, every 6x6 camera-pair block is brought from the
  //      primed frame to the real one, S_ab = D_a S'_ab D_b^T, then one red.add per entry ------------------------------
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* S = V.Sacc + V.w_red_off[w];
  const int sbeta = V.w_beta[w];
  const int nloc = 6 * nlf;
  const double* camG = V.camR[cur];             // G of camera gc: camG[gc * kCamStride + 12 .. 20], row-major
  if (nlf > 0) {
    const int frow = lane >> 2, fc = (lane & 3) * 2;
    constexpr int LDS2 = 8 * T + 1;
    double* Sl = scratch;                         // [8 T][LDS2]
    for (int r = 0; r < KG; r++) {
      if (kq == r) {
#pragma unroll
        for (int i = 0; i < TPW; i++) {
          if (tI[i] >= 0) {
            double* d = Sl + (size_t)(8 * tI[i] + frow) * LDS2 + 8 * tJ[i] + fc;
            if (r == 0) { d[0] = acc[i][0]; d[1] = acc[i][1]; } else { d[0] += acc[i][0]; d[1] += acc[i][1]; }
          }
        }
      }
      __syncthreads();
    }
    // thread (block pair a <= b, row r): row r of D_a X D_b^T; all rows are read before any is written back
    const int nblk = nlf * (nlf + 1) / 2;
    constexpr int PER = (NT / 6) * 6;             // a block's six rows never straddle two passes
    for (int base = 0; base < nblk * 6; base += PER) {
      const int idx = base + t;
      const bool mine = t < PER && idx < nblk * 6;
      double out[6] = {0, 0, 0, 0, 0, 0};
      int a = 0, b = 0, r = 0;
      if (mine) {
        int blk = idx / 6; r = idx - blk * 6;
        while (blk >= nlf - a) { blk -= nlf - a; a++; }
        b = a + blk;
        const double* Ga = camG + (size_t)s_gc[nfx + a] * kCamStride + 12;
        const double* Gb = camG + (size_t)s_gc[nfx + b] * kCamStride + 12;
        // X(i, j) of the block; diagonal blocks hold their upper triangle only
        auto X = [&](int i, int j) -> double {
          const int ri = 6 * a + i, cj = 6 * b + j;
          return (a == b && j < i) ? Sl[(size_t)(6 * a + j) * LDS2 + 6 * a + i] : Sl[(size_t)ri * LDS2 + cj];
        };
        double row[6];
        if (r < 3) {
#pragma unroll
          for (int j = 0; j < 6; j++) row[j] = X(r, j);
        } else {
          const double g0 = Ga[r - 3], g1 = Ga[3 + r - 3], g2 = Ga[6 + r - 3];     // column r-3 of G_a = row of G_a^T
#pragma unroll
          for (int j = 0; j < 6; j++) row[j] = fma(g0, X(3, j), fma(g1, X(4, j), g2 * X(5, j)));
        }
        out[0] = row[0]; out[1] = row[1]; out[2] = row[2];
#pragma unroll
        for (int j = 0; j < 3; j++) out[3 + j] = fma(row[3], Gb[j], fma(row[4], Gb[3 + j], row[5] * Gb[6 + j]));
      }
      __syncthreads();
      if (mine) {
#pragma unroll
        for (int j = 0; j < 6; j++) if (a != b || j >= r) Sl[(size_t)(6 * a + r) * LDS2 + 6 * b + j] = out[j];
      }
      __syncthreads();
    }
    for (int idx = t; idx < nloc * nloc; idx += NT) {
      const int row = idx / nloc, colx = idx - row * nloc;
      if (colx < row) continue;
      const double v = Sl[(size_t)row * LDS2 + colx];
      if (v == 0.0) continue;
      const int a = row / 6, b = colx / 6;
      const int fa = s_free[nfx + a], fb = s_free[nfx + b];
      atomicAdd(&S[sacc_index(n, sbeta, 6 * fa + row - 6 * a, 6 * fb + (colx - 6 * b))], v);
    }
    __syncthreads();
  }
  // ---- flush: camera blocks B', gradients v', rhs terms Z' h: summed over a slot's lanes through shared memory, brought
  //      to the real frame (B = D B' D^T, v = D v', Z h = D Z' h) and added to the global sums -------------------------------
  if (p1_thread && my_free) {
    double* o = scratch + (size_t)(pl * nl + sl) * 33;
    // packed upper triangle of B': (0,0) (0,1)=0 (0,2) | -U row 0 ; (1,1) (1,2) | -U row 1 ; (2,2) | -U row 2 ; B'rr
    o[ut6(0, 0)] = Bt[0]; o[ut6(0, 1)] = 0.0; o[ut6(0, 2)] = Bt[1]; o[ut6(1, 1)] = Bt[2]; o[ut6(1, 2)] = Bt[3]; o[ut6(2, 2)] = Bt[4];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) o[ut6(r, 3 + c)] = -Bu[r * 3 + c];
    o[ut6(3, 3)] = Br[0]; o[ut6(3, 4)] = Br[1]; o[ut6(3, 5)] = Br[2]; o[ut6(4, 4)] = Br[3]; o[ut6(4, 5)] = Br[4]; o[ut6(5, 5)] = Br[5];
#pragma unroll
    for (int i = 0; i < 6; i++) { o[21 + i] = vq[i]; o[27 + i] = zq[i]; }
  }
  __syncthreads();
  double* Bs = scratch + (size_t)NT * 33;       // [nlf][33] sums over the chunk's point slots
  for (int idx = t; idx < nlf * 33; idx += NT) {
    const int s2i = nfx + idx / 33, e = idx % 33;
    double sacc = 0.0;
    for (int q = 0; q < Pc; q++) sacc += scratch[(size_t)(q * nl + s2i) * 33 + e];
    Bs[idx] = sacc;
  }
  __syncthreads();
  for (int idx = t; idx < nlf * 6; idx += NT) {
    const int a = idx / 6, r = idx - a * 6;
    const int gc = s_gc[nfx + a];
    const double* Bp = Bs + (size_t)a * 33;
    const double* G = camG + (size_t)gc * kCamStride + 12;
    auto X = [&](int i, int j) -> double { return i <= j ? Bp[ut6(i, j)] : Bp[ut6(j, i)]; };
    double row[6], vr, zr;
    if (r < 3) {
#pragma unroll
      for (int j = 0; j < 6; j++) row[j] = X(r, j);
      vr = Bp[21 + r]; zr = Bp[27 + r];
    } else {
      const double g0 = G[r - 3], g1 = G[3 + r - 3], g2 = G[6 + r - 3];
#pragma unroll
      for (int j = 0; j < 6; j++) row[j] = fma(g0, X(3, j), fma(g1, X(4, j), g2 * X(5, j)));
      vr = fma(g0, Bp[24], fma(g1, Bp[25], g2 * Bp[26]));
      zr = fma(g0, Bp[30], fma(g1, Bp[31], g2 * Bp[32]));
    }
    double out[6];
    out[0] = row[0]; out[1] = row[1]; out[2] = row[2];
#pragma unroll
    for (int j = 0; j < 3; j++) out[3 + j] = fma(row[3], G[j], fma(row[4], G[3 + j], row[5] * G[6 + j]));
#pragma unroll
    for (int j = 0; j < 6; j++) if (j >= r && out[j] != 0.0) atomicAdd(&V.Bacc[(size_t)gc * 36 + r * 6 + j], out[j]);
    if (vr != 0.0) atomicAdd(&V.vacc[(size_t)gc * 6 + r], vr);
    if (zr != 0.0) atomicAdd(&V.zh[(size_t)gc * 6 + r], zr);
  }
  cost = warp_sum(cost); fail = warp_sum(fail); gmax = warp_max(gmax);
  if (warp_leader()) {
    if (cost != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_COST], cost);
    if (fail != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_FAIL], fail);
    if (gmax > 0.0) atomic_max_nonneg(&V.w_max[w], gmax);
  }
}

#ifndef UBA_T2_MINBLOCKS
#define UBA_T2_MINBLOCKS 2
#endif
// One kernel per (CTA size, tile variant) so that each variant gets its own register allocation.
// Parts are grouped by variant on the host: this launch covers parts [first, first + gridDim.x).
template <int M, int NT, int T, int TG>
__global__ void __launch_bounds__(NT, NT == 128 ? UBA_T2_MINBLOCKS : 1) k_lin_tile2(DevView V, int first) {
  extern __shared__ double sm[];
  const TilePart part = V.parts[first + blockIdx.x];
  if (V.ws[part.window].done) return;
  tile2_part<M, NT, T, TG>(V, part, sm);
}
#endif  // !UBA_EMU

// ---------------------------------------------------------------------------------------------
// tiled lineariser, third generation (k_lin_slot): WARP = CAMERA SLOT, LANE = POINT, plus one POINT WARP.
//
// Same plan (TilePart), same factored algebra and the same tensor-core Schur products as k_lin_tile2, but the
// (point, slot) grid is laid out the other way round: warp s of the CTA owns local camera slot s and its 32 lanes
// take 32 consecutive points of the part; one more warp (the point warp, lane = point) does everything that is per
// POINT.  What that buys (ncu on c4: k_lin_tile2 issues ~170 warp instructions per point, two thirds of them overhead):
//   * the camera record is warp-uniform: broadcast shared-memory loads, no bank conflicts;
//   * the per-point sums  C = sum E^T E,  g = sum E^T r  cross WARPS through shared memory ([slot][value][lane],
//     conflict-free) instead of a segmented shuffle tree per warp;
//   * damping, 3x3 factor, the point record and the gradient norm are done ONCE per point, by the point warp — which
//     holds none of the slot state, so its 36+ loads are all in flight at once — instead of redundantly by every slot
//     lane of the point; meanwhile the slot warps run the tensor-core products of the PREVIOUS chunk;
//   * every lane always holds an observation of its warp's slot: no lane padding (32 / nl), and Z' lands in the chunk
//     matrix with a lane stride of 3 doubles (conflict-free);
//   * each 8x8 output tile of the SYRK has ONE owner warp (tiles dealt round-robin over the slot warps): no K-group
//     reduction at the flush and 4-8 accumulator registers instead of 40.
// Per chunk of 32 points, slot warps:  linearise -> [A] -> SYRK of the previous chunk -> [B] -> stage Z';
//                        point warp:   prefetch  -> [A] -> sum, damp, factor, record  -> [B].          Two barriers.
// Features and points of the next chunk arrive through cp.async into per-lane staging slots ([value][thread]).
// Handles parts with up to kSlotMaxLocal local cameras (c4, c5: 5; c1, c3: 10); wider parts go to k_lin_wide.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cam_damping(const DevView& V, const WinState* st, int gc, int r, double d, bool store);
#ifndef UBA_EMU
constexpr int kSlotLdz = t2_ldz(32);              // 100: K = 96 columns per chunk of 32 points
constexpr int kSlotStageD = 7;                    // staged doubles per lane and buffer: point (3) + features (<= 4)
constexpr int kSlotStageI = 4;                    // ... and ints: next chunk's mask, its first-observation offset, obs_cam word, spare
__host__ __device__ constexpr int slot_rows(int nl_max) { return ((6 * nl_max + 7) / 8) * 8; }
// shared memory (doubles): camS | Zm [rows][kSlotLdz] | Cs [nl][9][32] | Ls [10][32] | staging; the flush scratch aliases Zm and Cs
__host__ __device__ constexpr int slot_main_doubles(int nl_max) {
  const int rows = slot_rows(nl_max), zc = rows * kSlotLdz + nl_max * 9 * 32, sl = rows * (rows + 1) + nl_max * 33;
  return zc > sl ? zc : sl;
}
__host__ __device__ constexpr int slot_stage_doubles(int nl_max) { return 2 * 32 * nl_max * (kSlotStageD + kSlotStageI / 2); }
__host__ __device__ constexpr size_t slot_smem_bytes(int nl_max) {
  return sizeof(double) * ((size_t)nl_max * kCamSm + slot_main_doubles(nl_max) + 10 * 32 + slot_stage_doubles(nl_max));
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <int ID>
__device__ __forceinline__ void named_bar(int nthreads) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory"); }

template <int M, int NLMAX, bool PIPE>
__device__ __forceinline__ void slot_part(const DevView& V, const TilePart& part, double* sm) {
  const unsigned FULL = 0xffffffffu;
  const int nl = part.n_local, nfx = part.n_fixed, nlf = nl - nfx;
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;        // my point lane; my camera slot (sl == nl: the point warp)
  if (sl > nl) return;                                            // the CTA is as wide as the widest part of its class
  const bool pt_warp = sl == nl;
  const int nslot = 32 * nl, nall = nslot + 32;                    // threads of the slot warps / of all live warps
  const int t = threadIdx.x;
  const int w = part.window;
  const WinState* st = &V.ws[w];
  const int cur = st->cur;
  const int cbase = V.w_cam_off[w];
  const int T = (6 * nlf + 7) / 8;                                 // 8-row tiles of the chunk matrix
  constexpr int ROWS = slot_rows(NLMAX);
  double* camS = sm;                                               // [nl][kCamSm]
  double* Zm = camS + NLMAX * kCamSm;                              // [ROWS][kSlotLdz]
  double* Cs = Zm + ROWS * kSlotLdz;                               // [nl][9][32]
  double* Ls = Zm + slot_main_doubles(NLMAX);                      // [10][32]  L^-1 (6), h (3) of the chunk's points
  double* stageD = Ls + 10 * 32;                                   // [2][kSlotStageD][32 NLMAX]
  int* stageI = reinterpret_cast<int*>(stageD + 2 * kSlotStageD * 32 * NLMAX);   // [2][kSlotStageI][32 NLMAX]
  double* scratch = Zm;
  __shared__ int s_free[NLMAX];
  __shared__ int s_gc[NLMAX];
  // the first chunk's mask / observation offset: issued ahead of the camera records so that the two chains of dependent
  // misses (part -> camera list -> camera record, part -> mask / offset -> point / features) run side by side
  unsigned mask_first = 0;
  int off_first = 0;
  if (!pt_warp && part.pt_begin + lane < part.pt_end) { mask_first = V.pt_mask[part.pt_begin + lane]; off_first = V.pt_obs_off[part.pt_begin + lane]; }
  __shared__ double s_G[NLMAX * 9];                                // G = J_l(r) of every slot's camera, for the flush
  for (int i = t; i < nl * 22; i += nall) {
    const int sl2 = i / 22, k = i - sl2 * 22;
    const int gc = cbase + V.tile_cams[part.cam_list_off + sl2];
    const double v = V.camR[cur][(size_t)gc * kCamStride + k];             // R (9), t (3), G (9), small-angle flag
    if (k < 12) camS[sl2 * kCamSm + k] = v; else if (k < 21) s_G[sl2 * 9 + k - 12] = v; else camS[sl2 * kCamSm + 12] = v;
    if (k == 0) { s_gc[sl2] = gc; s_free[sl2] = V.free_cam[gc]; }
  }
  for (int i = t; i < 8 * T * kSlotLdz; i += nall) Zm[i] = 0.0;     // padding rows / columns stay zero for the whole part
  // prefetch pipeline (cp.async, no registers held): while chunk k is processed, the point and features of my observation in
  // chunk k+1 and the mask / first-observation offset of my point in chunk k+2 are in flight into my staging column
  constexpr int NS = 32 * NLMAX;
  auto prefetch = [&](int b, int pa_, unsigned m, int off) {
    double* sd = stageD + (size_t)b * kSlotStageD * NS + t;
    int* si = stageI + (size_t)b * kSlotStageI * NS + t;
    if ((m >> sl) & 1u) {
      const double* px = V.pts[cur] + (size_t)pa_ * 3;
      cp_async8(sd, px); cp_async8(sd + NS, px + 1); cp_async8(sd + 2 * NS, px + 2);
      const int o = off + __popc(m & ((1u << sl) - 1u));
#pragma unroll
      for (int q = 0; q < M; q++) cp_async8(sd + (3 + q) * NS, V.feat + (size_t)q * V.NO + o);
      if (M == 2) cp_async4(si + 2 * NS, V.obs_cam + o);
    }
    const int pb_ = pa_ + 32;
    if (pb_ < part.pt_end) { cp_async4(si, V.pt_mask + pb_); cp_async4(si + NS, V.pt_obs_off + pb_); }
    else si[0] = 0;
    cp_async_commit();
  };
  if (!pt_warp) prefetch(0, part.pt_begin + lane, mask_first, off_first);
  named_bar<1>(nall);

  if (pt_warp) {
    // ================= the point warp: lane = point ===============================================================
    const double radius = st->radius;
    const double inv_radius = radius > 0.0 ? 1.0 / radius : 0.0;
    const bool scale_ready = st->scale_ready != 0;
    double gmax = 0.0, fail = 0.0;
    for (int c0 = part.pt_begin; c0 < part.pt_end; c0 += 32) {
      const int p = c0 + lane;
      // this chunk's per-point inputs: issued here, long before they are needed after [A]
      unsigned mask = 0;
      double X[3] = {0, 0, 0}, is2[3] = {1, 1, 1};
      if (p < part.pt_end) mask = V.pt_mask[p];
      if (mask) {
        X[0] = V.pts[cur][(size_t)p * 3]; X[1] = V.pts[cur][(size_t)p * 3 + 1]; X[2] = V.pts[cur][(size_t)p * 3 + 2];
        if (scale_ready) { is2[0] = V.pt_is2[(size_t)p * 3]; is2[1] = V.pt_is2[(size_t)p * 3 + 1]; is2[2] = V.pt_is2[(size_t)p * 3 + 2]; }
      }
      named_bar<1>(nall);                                           // [A] the slots' partial sums are in Cs
      double Li[6] = {0, 0, 0, 0, 0, 0}, h[3] = {0, 0, 0}, lam[3] = {0, 0, 0}, gg[3] = {0, 0, 0};
      bool ok = false;
      if (mask) {
        double c9[9];
#pragma unroll
        for (int e = 0; e < 9; e++) c9[e] = Cs[e * 32 + lane];
        for (int s2 = 1; s2 < nl; s2++) {
          const double* cs = Cs + (size_t)s2 * 9 * 32 + lane;
#pragma unroll
          for (int e = 0; e < 9; e++) c9[e] += cs[e * 32];
        }
        const double Cd[3] = {c9[0], c9[3], c9[5]};
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (!scale_ready) is2[c] = jacobi_is2(Cd[c], V.cfg.jacobi_scaling);
          lam[c] = lm_lambda_inv(Cd[c], is2[c], inv_radius, V.cfg.min_lm_diagonal, V.cfg.max_lm_diagonal);
        }
        const double Cdamp[6] = {c9[0] + lam[0], c9[1], c9[2], c9[3] + lam[1], c9[4], c9[5] + lam[2]};
        gg[0] = c9[6]; gg[1] = c9[7]; gg[2] = c9[8];
        ok = point_factor(Cdamp, Li);
        if (ok) linv_mul(Li, gg, h);
        else {
#pragma unroll
          for (int i = 0; i < 6; i++) Li[i] = 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) Ls[i * 32 + lane] = Li[i];
#pragma unroll
      for (int i = 0; i < 3; i++) Ls[(6 + i) * 32 + lane] = h[i];
      named_bar<2>(nall);                                           // [B] (the record and the norms are off the slots' path)
      if (mask) {
        double2* rec = reinterpret_cast<double2*>(V.pt_rec + (size_t)p * kPtRec);   // 128-byte records: eight 16-byte stores
        if (!ok) {
          fail += 1.0;
#pragma unroll
          for (int i = 0; i < kPtRec / 2; i++) rec[i] = make_double2(0.0, 0.0);
        } else {
          rec[0] = make_double2(Li[0], Li[1]); rec[1] = make_double2(Li[2], Li[3]); rec[2] = make_double2(Li[4], Li[5]);
          rec[3] = make_double2(h[0], h[1]); rec[4] = make_double2(h[2], gg[0]); rec[5] = make_double2(gg[1], gg[2]);
          rec[6] = make_double2(lam[0], lam[1]); rec[7] = make_double2(lam[2], 0.0);
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double proj = V.cfg.use_bounds ? clampd(X[c] - gg[c], V.calib.lo[c], V.calib.hi[c]) : X[c] - gg[c];
            gmax = fmax(gmax, fabs(X[c] - proj));
          }
        }
        if (!scale_ready) { V.pt_is2[(size_t)p * 3] = is2[0]; V.pt_is2[(size_t)p * 3 + 1] = is2[1]; V.pt_is2[(size_t)p * 3 + 2] = is2[2]; }
      }
    }
    named_bar<1>(nall);                                             // [A'] of the epilogue
    fail = warp_sum(fail); gmax = warp_max(gmax);
    if (lane == 0) {
      if (fail != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_FAIL], fail);
      if (gmax > 0.0) atomic_max_nonneg(&V.w_max[w], gmax);
    }
    return;
  }

  // ================= slot warps: warp = camera slot, lane = point ====================================================
  const bool my_free = sl >= nfx;
  const int ntiles = T * (T + 1) / 2;
  // my SYRK tiles: idx = sl + i nl, enumerated (I <= J) row by row
  constexpr int TPW = NLMAX <= 5 ? 2 : 4;                          // T (T + 1) / 2 <= TPW nl for every part of the class
  int tI[TPW], tJ[TPW], nmine = 0;
  unsigned pa[TPW], pb[TPW];                                       // shared-window byte addresses of my lane's fragment rows of tile i
#pragma unroll
  for (int i = 0; i < TPW; i++) {
    int idx = sl + i * nl, I = 0;
    const unsigned zm0 = (unsigned)__cvta_generic_to_shared(Zm);
    if (idx >= ntiles) { tI[i] = -1; tJ[i] = 0; pa[i] = zm0; pb[i] = zm0; continue; }
    while (idx >= T - I) { idx -= T - I; I++; }
    tI[i] = I; tJ[i] = I + idx; nmine = i + 1;
    pa[i] = zm0 + 8u * (unsigned)((8 * I + (lane >> 2)) * kSlotLdz + (lane & 3));
    pb[i] = zm0 + 8u * (unsigned)((8 * tJ[i] + (lane >> 2)) * kSlotLdz + (lane & 3));
  }
  double acc[TPW][2];
#pragma unroll
  for (int i = 0; i < TPW; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  // tensor-core SYRK over the chunk in Zm, every tile on its owner warp (K = 96: 24 steps of k = 4)
  auto syrk = [&]() {
    if (nmine > 0) {
      auto syrk2 = [&](int i, int j) {        // two tiles side by side: independent accumulator chains
#pragma unroll
        for (int ks = 0; ks < 24; ks++) {
          dmma884(acc[i][0], acc[i][1], lds_f64(pa[i] + 32 * ks), lds_f64(pb[i] + 32 * ks));
          dmma884(acc[j][0], acc[j][1], lds_f64(pa[j] + 32 * ks), lds_f64(pb[j] + 32 * ks));
        }
      };
      auto syrk1 = [&](int i) {
#pragma unroll
        for (int ks = 0; ks < 24; ks++) dmma884(acc[i][0], acc[i][1], lds_f64(pa[i] + 32 * ks), lds_f64(pb[i] + 32 * ks));
      };
      if (nmine >= 2) syrk2(0, 1); else syrk1(0);
      if (TPW > 2) {
        if (nmine >= 4) syrk2(TPW > 2 ? 2 : 0, TPW > 2 ? 3 : 0); else if (nmine == 3) syrk1(TPW > 2 ? 2 : 0);
      }
    }
  };
  // camera-side sums of my slot over my lane's points, in the frame where G = J_l(r) has not been applied (see k_lin_tile2)
  double Bt[5], Bu[9], Br[6], vq[6], zq[6];
#pragma unroll
  for (int i = 0; i < 5; i++) Bt[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) Bu[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) { Br[i] = 0.0; vq[i] = 0.0; zq[i] = 0.0; }
  double cost = 0.0;

  unsigned mask_cur = mask_first;
  const double* R = camS + sl * kCamSm;
  int buf = 0;
  bool have_prev = false;
  for (int c0 = part.pt_begin; c0 < part.pt_end; c0 += 32, buf ^= 1) {
    const int p = c0 + lane;
    const unsigned mask = mask_cur;
    const bool seen = (mask >> sl) & 1u;
    double X[3] = {0, 0, 0}, f[M];
#pragma unroll
    for (int q = 0; q < M; q++) f[q] = 0.0;
    int cid = 0;
    {
      cp_async_wait_all();
#ifdef UBA_BAND_TIMING
      if (t == 0 && c0 == part.pt_begin && blockIdx.x < 1000) { long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); V.Zbuf[2100 + blockIdx.x] = (double)(g_ % 1000000000ll); }
#endif
      const double* sd = stageD + (size_t)buf * kSlotStageD * NS + t;
      const int* si = stageI + (size_t)buf * kSlotStageI * NS + t;
      if (seen) {
        X[0] = sd[0]; X[1] = sd[NS]; X[2] = sd[2 * NS];
#pragma unroll
        for (int q = 0; q < M; q++) f[q] = sd[(3 + q) * NS];
        if (M == 2) cid = (si[2 * NS] >> 30) & 1;
      }
      const unsigned mask_next = (unsigned)si[0];
      const int off_next = si[NS];
      prefetch(buf ^ 1, p + 32, mask_next, off_next);
      mask_cur = mask_next;
    }
    // ---- linearise my observation; partial sums of the point block to shared memory ---------------------------------
    double Pm[9], u[3] = {0, 0, 0}, cg[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { Pm[i] = 0.0; cg[i] = 0.0; }
    if (seen) {
      double Q[5], m3[3];
      const double rho0 = obs_linearize_q<M>(R, R + 9, R[12] != 0.0, X, f, cid, V.calib, V.loss, Q, m3, u);
      cost += 0.5 * rho0;
      const double Q00 = Q[0], Q02 = Q[1], Q11 = Q[2], Q12 = Q[3], Q22 = Q[4];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        Pm[c] = fma(Q00, R[c], Q02 * R[6 + c]);
        Pm[3 + c] = fma(Q11, R[3 + c], Q12 * R[6 + c]);
        Pm[6 + c] = fma(Q02, R[c], fma(Q12, R[3 + c], Q22 * R[6 + c]));
      }
      cg[0] = fma(R[0], Pm[0], fma(R[3], Pm[3], R[6] * Pm[6]));
      cg[1] = fma(R[0], Pm[1], fma(R[3], Pm[4], R[6] * Pm[7]));
      cg[2] = fma(R[0], Pm[2], fma(R[3], Pm[5], R[6] * Pm[8]));
      cg[3] = fma(R[1], Pm[1], fma(R[4], Pm[4], R[7] * Pm[7]));
      cg[4] = fma(R[1], Pm[2], fma(R[4], Pm[5], R[7] * Pm[8]));
      cg[5] = fma(R[2], Pm[2], fma(R[5], Pm[5], R[8] * Pm[8]));
      cg[6] = fma(R[0], m3[0], fma(R[3], m3[1], R[6] * m3[2]));
      cg[7] = fma(R[1], m3[0], fma(R[4], m3[1], R[7] * m3[2]));
      cg[8] = fma(R[2], m3[0], fma(R[5], m3[1], R[8] * m3[2]));
      if (my_free) {
        const double ux = u[0], uy = u[1], uz = u[2];
        Bt[0] += Q00; Bt[1] += Q02; Bt[2] += Q11; Bt[3] += Q12; Bt[4] += Q22;
        const double U00 = -Q02 * uy, U01 = fma(Q02, ux, -Q00 * uz), U02 = Q00 * uy;
        const double U10 = fma(Q11, uz, -Q12 * uy), U11 = Q12 * ux, U12 = -Q11 * ux;
        const double U20 = fma(Q12, uz, -Q22 * uy), U21 = fma(Q22, ux, -Q02 * uz), U22 = fma(Q02, uy, -Q12 * ux);
        Bu[0] += U00; Bu[1] += U01; Bu[2] += U02; Bu[3] += U10; Bu[4] += U11; Bu[5] += U12; Bu[6] += U20; Bu[7] += U21; Bu[8] += U22;
        Br[0] = fma(uz, U10, fma(-uy, U20, Br[0]));
        Br[1] = fma(uz, U11, fma(-uy, U21, Br[1]));
        Br[2] = fma(uz, U12, fma(-uy, U22, Br[2]));
        Br[3] = fma(ux, U21, fma(-uz, U01, Br[3]));
        Br[4] = fma(ux, U22, fma(-uz, U02, Br[4]));
        Br[5] = fma(uy, U02, fma(-ux, U12, Br[5]));
        vq[0] += m3[0]; vq[1] += m3[1]; vq[2] += m3[2];
        vq[3] = fma(uy, m3[2], fma(-uz, m3[1], vq[3]));
        vq[4] = fma(uz, m3[0], fma(-ux, m3[2], vq[4]));
        vq[5] = fma(ux, m3[1], fma(-uy, m3[0], vq[5]));
      }
    }
    {
      double* cs = Cs + (size_t)sl * 9 * 32 + lane;
#pragma unroll
      for (int e = 0; e < 9; e++) cs[e * 32] = cg[e];
    }
    named_bar<1>(nall);                       // [A] partial sums visible to the point warp; Z' of the previous chunk complete
    // ---- while the point warp factors this chunk's points: Schur products of the PREVIOUS chunk ---------------------
    if (have_prev) syrk();
    have_prev = true;
    named_bar<2>(nall);                       // [B] L^-1, h of this chunk in Ls; every warp is done reading the old Zm
    // ---- (free slots) Z' = W' L^-T into the chunk matrix: rows 0..2 P L^-T, rows 3..5 [u]x (P L^-T) ------------------
    if (my_free) {
      double Li[6], h[3];
#pragma unroll
      for (int i = 0; i < 6; i++) Li[i] = Ls[i * 32 + lane];
#pragma unroll
      for (int i = 0; i < 3; i++) h[i] = Ls[(6 + i) * 32 + lane];
      double zt[3][3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        zt[r][0] = Pm[r * 3] * Li[0];
        zt[r][1] = fma(Pm[r * 3], Li[1], Pm[r * 3 + 1] * Li[2]);
        zt[r][2] = fma(Pm[r * 3], Li[3], fma(Pm[r * 3 + 1], Li[4], Pm[r * 3 + 2] * Li[5]));
      }
      double* zc = Zm + (size_t)(6 * (sl - nfx)) * kSlotLdz + 3 * lane;     // (not seen: P = 0, u = 0 -> zeros)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        zc[c] = zt[0][c]; zc[kSlotLdz + c] = zt[1][c]; zc[2 * kSlotLdz + c] = zt[2][c];
        zc[3 * kSlotLdz + c] = fma(u[1], zt[2][c], -u[2] * zt[1][c]);
        zc[4 * kSlotLdz + c] = fma(u[2], zt[0][c], -u[0] * zt[2][c]);
        zc[5 * kSlotLdz + c] = fma(u[0], zt[1][c], -u[1] * zt[0][c]);
      }
      const double y0 = fma(zt[0][0], h[0], fma(zt[0][1], h[1], zt[0][2] * h[2]));
      const double y1 = fma(zt[1][0], h[0], fma(zt[1][1], h[1], zt[1][2] * h[2]));
      const double y2 = fma(zt[2][0], h[0], fma(zt[2][1], h[1], zt[2][2] * h[2]));
      zq[0] += y0; zq[1] += y1; zq[2] += y2;
      zq[3] = fma(u[1], y2, fma(-u[2], y1, zq[3]));
      zq[4] = fma(u[2], y0, fma(-u[0], y2, zq[4]));
      zq[5] = fma(u[0], y1, fma(-u[1], y0, zq[5]));
    }
  }
  named_bar<1>(nall);                         // [A'] the last chunk's Z' is staged
  syrk();
  named_bar<3>(nslot);                        // from here on only the slot warps

  // ---- flush: Schur tiles -> shared memory, S_ab = D_a S'_ab D_b^T per camera-pair block, one red.add per entry ------
#ifdef UBA_BAND_TIMING
  if (t == 0 && blockIdx.x < 1000) { long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); V.Zbuf[3100 + blockIdx.x] = (double)(g_ % 1000000000ll); }
#endif
  //      One barrier: every slot warp stores its tiles and (free slots) the lane sums of its B', v', Z' h; then one thread
  //      per block row brings it to the real frame (G from shared memory) and adds it to the accumulators directly.
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* S = V.Sacc + V.w_red_off[w];
  const int sbeta = V.w_beta[w];
  const int LDS2 = 8 * T + 1;
  double* Sl = scratch;                           // [8 T][LDS2]
  double* Bs = scratch + ROWS * (ROWS + 1);       // [nlf][32]: packed upper triangle of B' (without the zero at (0,1)), v', Z' h
  if (nlf > 0) {
    const int frow = lane >> 2, fc = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < TPW; i++) {
      if (tI[i] >= 0) {
        double* d = Sl + (size_t)(8 * tI[i] + frow) * LDS2 + 8 * tJ[i] + fc;
        d[0] = acc[i][0]; d[1] = acc[i][1];
      }
    }
  }
  if (my_free) {
    // 32 values summed over 32 lanes by a transposing reduction (31 exchanges instead of 5 x 32): round h keeps, on the
    // lanes with bit h set, the values whose index has bit h set; lane l ends up with the total of value l
    double val[32];
    auto E = [](int e) { return e ? e - 1 : 0; };                     // packed index without ut6(0, 1)
    val[E(ut6(0, 0))] = Bt[0]; val[E(ut6(0, 2))] = Bt[1]; val[E(ut6(1, 1))] = Bt[2]; val[E(ut6(1, 2))] = Bt[3]; val[E(ut6(2, 2))] = Bt[4];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) val[E(ut6(r, 3 + c))] = -Bu[r * 3 + c];
    val[E(ut6(3, 3))] = Br[0]; val[E(ut6(3, 4))] = Br[1]; val[E(ut6(3, 5))] = Br[2]; val[E(ut6(4, 4))] = Br[3]; val[E(ut6(4, 5))] = Br[4]; val[E(ut6(5, 5))] = Br[5];
#pragma unroll
    for (int i = 0; i < 6; i++) { val[E(21 + i)] = vq[i]; val[E(27 + i)] = zq[i]; }
#pragma unroll
    for (int hh = 16; hh >= 1; hh >>= 1) {
      const bool up = (lane & hh) != 0;
#pragma unroll
      for (int e = 0; e < hh; e++) {
        const double keep = up ? val[e + hh] : val[e];
        const double give = up ? val[e] : val[e + hh];
        val[e] = keep + __shfl_xor_sync(FULL, give, hh);
      }
    }
    Bs[(size_t)(sl - nfx) * 32 + lane] = val[0];
  }
  named_bar<3>(nslot);
  const int nS = 3 * nlf * (nlf + 1), nB = 6 * nlf;                   // block rows of S' (upper block triangle), rows of B'
  for (int idx = t; idx < nS + nB; idx += nslot) {
    if (idx < nS) {
      int blk = idx / 6, a = 0;
      const int r = idx - blk * 6;
      while (blk >= nlf - a) { blk -= nlf - a; a++; }
      const int b = a + blk;
      const double* Ga = s_G + (nfx + a) * 9;
      const double* Gb = s_G + (nfx + b) * 9;
      auto X = [&](int i, int j) -> double {
        return (a == b && j < i) ? Sl[(size_t)(6 * a + j) * LDS2 + 6 * a + i] : Sl[(size_t)(6 * a + i) * LDS2 + 6 * b + j];
      };
      double row[6];
      if (r < 3) {
#pragma unroll
        for (int j = 0; j < 6; j++) row[j] = X(r, j);
      } else {
        const double g0 = Ga[r - 3], g1 = Ga[3 + r - 3], g2 = Ga[6 + r - 3];
#pragma unroll
        for (int j = 0; j < 6; j++) row[j] = fma(g0, X(3, j), fma(g1, X(4, j), g2 * X(5, j)));
      }
      double out[6];
      out[0] = row[0]; out[1] = row[1]; out[2] = row[2];
#pragma unroll
      for (int j = 0; j < 3; j++) out[3 + j] = fma(row[3], Gb[j], fma(row[4], Gb[3 + j], row[5] * Gb[6 + j]));
      const int fa = s_free[nfx + a], fb = s_free[nfx + b];
#pragma unroll
      for (int j = 0; j < 6; j++)
        if ((a != b || j >= r) && out[j] != 0.0) atomicAdd(&S[sacc_index(n, sbeta, 6 * fa + r, 6 * fb + j)], out[j]);
    } else {
      const int a = (idx - nS) / 6, r = (idx - nS) - a * 6;
      const int gc = s_gc[nfx + a];
      const double* Bp = Bs + (size_t)a * 32;
      const double* G = s_G + (nfx + a) * 9;
      auto P = [&](int e) -> double { return Bp[e ? e - 1 : 0]; };
      auto X = [&](int i, int j) -> double { const int e = i <= j ? ut6(i, j) : ut6(j, i); return e == 1 ? 0.0 : P(e); };
      double row[6], vr, zr;
      if (r < 3) {
#pragma unroll
        for (int j = 0; j < 6; j++) row[j] = X(r, j);
        vr = P(21 + r); zr = P(27 + r);
      } else {
        const double g0 = G[r - 3], g1 = G[3 + r - 3], g2 = G[6 + r - 3];
#pragma unroll
        for (int j = 0; j < 6; j++) row[j] = fma(g0, X(3, j), fma(g1, X(4, j), g2 * X(5, j)));
        vr = fma(g0, P(24), fma(g1, P(25), g2 * P(26)));
        zr = fma(g0, P(30), fma(g1, P(31), g2 * P(32)));
      }
      double out[6];
      out[0] = row[0]; out[1] = row[1]; out[2] = row[2];
#pragma unroll
      for (int j = 0; j < 3; j++) out[3 + j] = fma(row[3], G[j], fma(row[4], G[3 + j], row[5] * G[6 + j]));
#pragma unroll
      for (int j = 0; j < 6; j++) if (j >= r && out[j] != 0.0) atomicAdd(&V.Bacc[(size_t)gc * 36 + r * 6 + j], out[j]);
      if (vr != 0.0) atomicAdd(&V.vacc[(size_t)gc * 6 + r], vr);
      if (zr != 0.0) atomicAdd(&V.zh[(size_t)gc * 6 + r], zr);
    }
  }
  cost = warp_sum(cost);
  if (lane == 0 && cost != 0.0) atomicAdd(&V.w_lin[(size_t)w * WL_COUNT + WL_COST], cost);
  // ---- pipelined solve: count this part done for each of its cameras; the part that completes a camera assembles the
  //      camera's six rows of the damped band  A = B + Lambda - S  and of  rhs = v - Z h  (what k_assemble does for the whole
  //      matrix when the solve is not pipelined) and raises the rows' flag for the band solver, which is already running ------
  if constexpr (PIPE) {                         // (its own instantiation: the hooks cost registers the plain pass needs)
    __shared__ int s_last[NLMAX];
    __threadfence();                            // my atomics are performed before the counters move
    named_bar<3>(nslot);
    if (t < nl) {
      const int gc = s_gc[t];
      const int old = atomicAdd(&V.cam_done[gc], 1);
      s_last[t] = (old + 1 == V.cam_expect[gc] && s_free[t] >= 0) ? 1 : 0;
    }
    named_bar<3>(nslot);
    const int bw1 = sbeta + 1;
    double* Ab = V.A + V.w_red_off[w] + (size_t)2 * n * bw1;
    double* rhs = V.rhs + (size_t)6 * V.w_free_off[w];
    for (int a = 0; a < nl; a++) {
      if (!s_last[a]) continue;                 // (uniform over the CTA)
      __threadfence();                          // acquire side: everybody else's sums for this camera are in L2
      const int gc = s_gc[a], f = s_free[a];
      double gm = 0.0;
      for (int e = t; e < 6 * bw1; e += nslot) {
        const int r = e / bw1, c = e - r * bw1;
        const int i = 6 * f + r, k = i - sbeta + c;
        double val = 0.0;
        if (k >= 0) {
          const int fk = k / 6, ck = k - 6 * fk;
          if (fk == f) {
            const int rr = ck < r ? ck : r, cc = ck < r ? r : ck;
            const double b = __ldcg(&V.Bacc[(size_t)gc * 36 + rr * 6 + cc]);
            val = b - __ldcg(&S[sacc_index(n, sbeta, 6 * f + rr, 6 * f + cc)]);
            if (rr == cc) val += cam_damping(V, st, gc, r, b, true);
          } else {
            val = -__ldcg(&S[sacc_index(n, sbeta, k, i)]);
          }
        }
        Ab[(size_t)i * bw1 + c] = val;
      }
      if (t < 6) {
        const double v = __ldcg(&V.vacc[(size_t)gc * 6 + t]);
        rhs[6 * f + t] = v - __ldcg(&V.zh[(size_t)gc * 6 + t]);
        gm = fabs(v);
      }
      if (t < 32) { gm = warp_max(gm); if (t == 0 && gm > 0.0) atomic_max_nonneg(&V.w_max[w], gm); }
      __threadfence();
      named_bar<3>(nslot);
      if (t == 0) *(volatile int32_t*)&V.row_ready[f] = 1;
    }
  }
}

// One kernel per slot-count class (blockDim = 32 (NLMAX + 1)); parts are grouped by class on the host.  168 registers:
// twelve warps per SM (two 192-thread CTAs) put three warps on each of the four register-file partitions of 16 K
// registers, so 170 per thread is the most that can launch.
#ifndef UBA_SLOT_MAXNREG
#define UBA_SLOT_MAXNREG 168
#endif
template <int M, int NLMAX, bool PIPE>
__global__ void __maxnreg__(UBA_SLOT_MAXNREG) k_lin_slot(DevView V, int first) {
  extern __shared__ double sm[];
  const TilePart part = V.parts[first + blockIdx.x];
  if (V.ws[part.window].done) return;
#ifdef UBA_BAND_TIMING
  if (threadIdx.x == 0 && blockIdx.x < 1000) { long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); V.Zbuf[100 + blockIdx.x] = (double)(g_ % 1000000000ll); }
#endif
  slot_part<M, NLMAX, PIPE>(V, part, sm);
#ifdef UBA_BAND_TIMING
  if (threadIdx.x == 0 && blockIdx.x < 1000) { long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); V.Zbuf[1100 + blockIdx.x] = (double)(g_ % 1000000000ll); }
#endif
}
#endif  // !UBA_EMU

// LM damping of camera column r of camera gc (B_rr = d); the Jacobi scale is captured at iteration 0.
__device__ __forceinline__ double cam_damping(const DevView& V, const WinState* st, int gc, int r, double d, bool store) {
  double s2;
  if (!st->scale_ready) { s2 = jacobi_s2(d, V.cfg.jacobi_scaling); if (store) V.cam_s2[(size_t)gc * 6 + r] = s2; }
  else s2 = V.cam_s2[(size_t)gc * 6 + r];
  const double lam = lm_lambda(d, s2, st->radius, V.cfg.min_lm_diagonal, V.cfg.max_lm_diagonal);
  if (store) V.cam_lam[(size_t)gc * 6 + r] = lam;
  return lam;
}

// symmetric entry (i, j), i >= j, of the damped reduced camera matrix, read from the upper block
// triangle of the Schur accumulator, the camera blocks B and the LM damping
__device__ __forceinline__ double reduced_entry(const DevView& V, const WinState* st, int f0, int n, int sbeta, const double* S, int i, int j) {
  const int fr = j / 6, r = j - fr * 6, fc = i / 6, c = i - fc * 6;
  if (fr == fc) {
    const int gc = V.free_list[f0 + fr];
    const double b = V.Bacc[(size_t)gc * 36 + r * 6 + c];
    double val = b - S[sacc_index(n, sbeta, 6 * fr + r, 6 * fr + c)];
    if (r == c) val += cam_damping(V, st, gc, r, b, false);
    return val;
  }
  return -S[sacc_index(n, sbeta, j, i)];
}

// ---------------------------------------------------------------------------------------------
// assemble the damped reduced camera system  A = B + Lambda_c - sum_j Z Z^T,  rhs = v - sum_j Z h
// grid: (blocks, nW)
// ---------------------------------------------------------------------------------------------
__global__ void k_assemble(DevView V) {
  const int w = blockIdx.y;
  WinState* st = &V.ws[w];
  if (st->done) return;
  const int f0 = V.w_free_off[w];
  const int nf = V.w_free_off[w + 1] - f0;
  const int n = 6 * nf;
  if (n == 0) return;
  const int64_t red = V.w_red_off[w];
  const double* S = V.Sacc + red;
  double* A = V.A + red;
  double* rhs = V.rhs + (size_t)6 * f0;
  // banded windows: pass 1 (this loop) builds lambda on the diagonal + rhs; pass 2 (below) writes the band
  const int sbeta = V.w_beta[w];
  const bool banded = sbeta > 0;
  const int64_t nmat = banded ? (int64_t)n : (int64_t)n * n;
  const int64_t total = nmat + n;
  double gmax = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nmat) {
      const int i = banded ? (int)e : (int)(e / n), j = banded ? (int)e : (int)(e % n);
      const int fa = i / 6, r = i % 6, fb = j / 6, c = j % 6;
      double val;
      if (fa == fb) {
        const int gc = V.free_list[f0 + fa];
        const double* B = V.Bacc + (size_t)gc * 36;
        const int rr = r < c ? r : c, cc = r < c ? c : r;
        val = B[rr * 6 + cc] - S[sacc_index(n, sbeta, 6 * fa + rr, 6 * fa + cc)];
        if (r == c) val += cam_damping(V, st, gc, r, B[r * 6 + r], true);
      } else if (fa < fb) {
        val = -S[sacc_index(n, sbeta, i, j)];
      } else {
        val = -S[sacc_index(n, sbeta, j, i)];
      }
      if (!banded) A[(size_t)i * n + j] = val;
    } else {
      const int i = (int)(e - nmat);
      const int gc = V.free_list[f0 + i / 6];
      const double v = V.vacc[(size_t)gc * 6 + i % 6];
      rhs[i] = v - V.zh[(size_t)gc * 6 + i % 6];
      gmax = fmax(gmax, fabs(v));
    }
  }
  gmax = warp_max(gmax);
  if (warp_leader() && gmax > 0.0) atomic_max_nonneg(&V.w_max[w], gmax);
  if (banded) {
    // compact band copy for k_chol_banded: Ab[i][c] = A[i][i - beta + c], stored after the factor rows
    const int beta = V.w_beta[w], bw1 = beta + 1;
    double* Ab = A + (size_t)2 * n * bw1;   // after the factor rows of both halves of the two-sided solver
    const int64_t nb = (int64_t)n * bw1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nb; e += (int64_t)gridDim.x * blockDim.x) {
      const int i = (int)(e / bw1), c = (int)(e - (int64_t)i * bw1);
      const int k = i - beta + c;
      Ab[e] = k >= 0 ? reduced_entry(V, st, f0, n, V.w_beta[w], S, i, k) : 0.0;
    }
  }
}

#ifndef UBA_EMU
// ---------------------------------------------------------------------------------------------
// dense Cholesky solve, one CTA per window, matrix resident in shared memory (n <= max_n)
// ---------------------------------------------------------------------------------------------
// 6x6 Cholesky of a diagonal block held in registers (lower triangle of L[6][6]), one thread,
// right-looking so that only rsqrt -> mul -> fma sits on the dependent chain of each pivot.
// inv[c] = 1 / L_cc.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol6(double (&L)[6][6], double (&inv)[6]) {
  // The pivot test is kept OFF the dependent chain (measured: 1045 -> 669 cycles per block on B200,
  // tests/cuda/panel_bench.cu): a bad pivot poisons the block with NaN and the caller's flag discards the solve.
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const double d = L[c][c];
    ok = ok && (d >= 2.2250738585072014e-308) && (d <= 1.7976931348623157e308);
    const double iv = uba_rsqrt(d);
    inv[c] = iv;
    L[c][c] = d * iv;
#pragma unroll
    for (int r = 0; r < 6; r++) if (r > c) L[r][c] *= iv;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int k = 0; k < 6; k++) if (r > c && k > c && k <= r) L[r][k] = fma(-L[r][c], L[k][c], L[r][k]);
  }
  return ok;
}

// Dense Cholesky solve, one CTA per window, matrix resident in shared memory (n = 6 nf <= max_n).
// Left-looking by 6-wide block columns (a camera block): per block column one dot-product update of
// the column (all threads), one serial 6x6 factor (one thread), one triangular solve of the rows
// below (thread per row).  The right-hand side rides along as row n, so the forward substitution is
// free; the backward substitution is blocked the same way.  Barriers: 3 + 2 per BLOCK column.
__global__ void __launch_bounds__(256) k_chol_small(DevView V, int max_n, int keep_factor) {
  extern __shared__ double sm[];
  const int w = blockIdx.x;
  WinState* st = &V.ws[w];
  if (st->done) return;
  const int f0 = V.w_free_off[w];
  const int nf = V.w_free_off[w + 1] - f0;
  const int n = 6 * nf;
  if (n == 0 || n > max_n || V.w_beta[w] > 0) return;
  const int ld = n + 1;                 // odd: conflict-free column walks
  double* a = sm;                       // [n + 1][ld], row n = right-hand side
  double* invd = sm + (size_t)(n + 1) * ld;  // [n] 1 / L_cc
  __shared__ int s_fail;
  double* A = V.A + V.w_red_off[w];
  double* rhs = V.rhs + (size_t)6 * f0;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_fail = 0;
  // lower triangle in, one row per warp pass (no index divisions)
  for (int i = tid >> 5; i < n; i += nt >> 5)
    for (int j = tid & 31; j <= i; j += 32) a[i * ld + j] = A[(size_t)i * n + j];
  for (int i = tid; i < n; i += nt) a[n * ld + i] = rhs[i];
  __syncthreads();
  for (int kb = 0; kb < nf; kb++) {
    const int c0 = 6 * kb;
    // (1) left-looking update of block column kb (rows c0 .. n, the rhs row included)
    if (kb > 0) {
      for (int e = tid; e < (n + 1 - c0) * 6; e += nt) {
        const int i = c0 + e / 6, c = c0 + e % 6;
        if (c <= i) {
          const double* ri = a + i * ld;
          const double* rc = a + c * ld;
          double s0 = 0.0, s1 = 0.0;
          int m = 0;
          for (; m + 1 < c0; m += 2) { s0 = fma(ri[m], rc[m], s0); s1 = fma(ri[m + 1], rc[m + 1], s1); }
          if (m < c0) s0 = fma(ri[m], rc[m], s0);
          a[i * ld + c] -= s0 + s1;
        }
      }
      __syncthreads();
    }
    // (2) serial 6x6 factor of the diagonal block
    if (tid == 0) {
      double L[6][6], iv[6];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) L[r][c] = c <= r ? a[(c0 + r) * ld + c0 + c] : 0.0;
      if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        invd[c0 + r] = iv[r];
#pragma unroll
        for (int c = 0; c < 6; c++) if (c <= r) a[(c0 + r) * ld + c0 + c] = L[r][c];
      }
    }
    __syncthreads();
    // (3) rows below (and the rhs row): X L_kk^T = A
    for (int i = c0 + 6 + tid; i <= n; i += nt) {
      double x[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        double v = a[i * ld + c0 + c];
#pragma unroll
        for (int m = 0; m < 6; m++) if (m < c) v = fma(-x[m], a[(c0 + c) * ld + c0 + m], v);
        x[c] = v * invd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; c++) a[i * ld + c0 + c] = x[c];
    }
    __syncthreads();
  }
  // backward substitution L^T x = z (z sits in row n), blocked
  double* z = a + n * ld;
  for (int kb = nf - 1; kb >= 0; kb--) {
    const int c0 = 6 * kb;
    if (tid == 0) {
      double xb[6];
#pragma unroll
      for (int c = 5; c >= 0; c--) {
        double v = z[c0 + c];
#pragma unroll
        for (int m = 0; m < 6; m++) if (m > c) v = fma(-a[(c0 + m) * ld + c0 + c], xb[m], v);
        xb[c] = v * invd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; c++) z[c0 + c] = xb[c];
    }
    __syncthreads();
    for (int i = tid; i < c0; i += nt) {
      double v = z[i];
#pragma unroll
      for (int c = 0; c < 6; c++) v = fma(-a[(c0 + c) * ld + i], z[c0 + c], v);
      z[i] = v;
    }
    __syncthreads();
  }
  const bool failed = s_fail != 0;
  for (int i = tid; i < n; i += nt) rhs[i] = failed ? 0.0 : z[i];
  // the factor (lower triangle) goes back only when the covariance extraction asks for it
  if (keep_factor) {
    for (int i = tid >> 5; i < n; i += nt >> 5)
      for (int j = tid & 31; j <= i; j += 32) A[(size_t)i * n + j] = a[i * ld + j];
  }
  if (tid == 0 && failed) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
}

// ---- large windows: blocked right-looking Cholesky in global memory (lower triangle of A) -----
constexpr int NB = 32;

#ifndef UBA_EMU
// Dense Cholesky solve for small windows (n <= max_n <= 160), RIGHT-looking with lookahead — the dense sibling of
// k_chol_banded_la.  Per 6-wide block step: warp 7 (the panel warp) solves the next block's six rows against L_kk,
// updates that block's 6x6 corner and factors it in registers (pivot tests off the dependent chain), while the 224
// workers solve all rows below (thread per row, the rhs rides along as one more row), then apply the rank-6 trailing
// update on the FP64 tensor-core path (DMMA.8x8x4 tiles over the remaining lower triangle, dealt to the 7 worker warps).
// One named barrier (the panel warp only ARRIVES: it has finished reading the entries the workers then overwrite in
// place) and one CTA barrier per step, against five in k_chol_small, whose serial 6x6 factor also stalls the whole CTA.
__global__ void __launch_bounds__(256) k_chol_small_la(DevView V, int max_n, int keep_factor) {
  extern __shared__ double sm[];
  const int w = blockIdx.x;
  WinState* st = &V.ws[w];
  if (st->done) return;
  constexpr int NWORK = 224, XS = 9;
  const int f0 = V.w_free_off[w];
  const int nf = V.w_free_off[w + 1] - f0;
  const int n = 6 * nf;
  if (n == 0 || n > max_n || V.w_beta[w] > 0) return;
  const int ld = n + 1;                 // odd: conflict-free column walks
  double* a = sm;                       // [n + 1][ld], row n = right-hand side
  double* invd = a + (size_t)(n + 1) * ld;   // [n] 1 / L_cc
  double* Xbuf = invd + n + (n & 1);    // [n + 8][XS] row solves of the current step; columns 6..8 stay zero (DMMA padding)
  __shared__ int s_fail;
  __shared__ double s_Lkk[2][36], s_invk[2][6], s_xp[36], s_corner[21];
  double* A = V.A + V.w_red_off[w];
  double* rhs = V.rhs + (size_t)6 * f0;
  const int t = threadIdx.x, nt = blockDim.x;
  const bool panel = t >= NWORK;
  const int pl = t - NWORK, lane = t & 31, warp = t >> 5;
  if (t == 0) s_fail = 0;
  // lower triangle in: four rows x five 32-column pieces per warp pass, all loads issued before the first store (a plain
  // load-store loop costs one memory round trip per element)
  for (int i0 = warp; i0 < n; i0 += 4 * (nt >> 5)) {
    double v[4][5];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * (nt >> 5);
#pragma unroll
      for (int c = 0; c < 5; c++) { const int j = lane + 32 * c; v[u][c] = (i < n && j <= i) ? A[(size_t)i * n + j] : 0.0; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * (nt >> 5);
#pragma unroll
      for (int c = 0; c < 5; c++) { const int j = lane + 32 * c; if (i < n && j <= i) a[i * ld + j] = v[u][c]; }
    }
  }
  for (int i = t; i < n; i += nt) a[n * ld + i] = rhs[i];
  for (int i = t; i < (n + 8) * XS; i += nt) Xbuf[i] = 0.0;
  int cr = 0, ce = pl;                        // corner entry of panel lane pl: (cr, ce), ce <= cr
  while (ce > cr) { ce -= cr + 1; cr++; }
  __syncthreads();
  if (t == NWORK) {                           // prologue: factor of block 0
    double L[6][6], iv[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) L[r][c] = c <= r ? a[r * ld + c] : 0.0;
    if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      s_invk[0][r] = iv[r]; invd[r] = iv[r];
#pragma unroll
      for (int c = 0; c < 6; c++) { s_Lkk[0][r * 6 + c] = L[r][c]; if (c <= r) a[r * ld + c] = L[r][c]; }
    }
  }
  __syncthreads();
  for (int kb = 0; kb < nf; kb++) {
    const int c0 = 6 * kb, par = kb & 1;
    const int mrows = n - c0 - 6;             // matrix rows below the block; row index mrows of Xbuf is the rhs
    const double* Lk = s_Lkk[par];
    const double* ivk = s_invk[par];
    if (panel) {
      const bool more = kb + 1 < nf;
      if (more && pl < 6) {
        const double* row = a + (c0 + 6 + pl) * ld + c0;
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = row[c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) s_xp[pl * 6 + c] = x[c];
      }
      __syncwarp();
      asm volatile("bar.arrive 1, 256;" ::: "memory");   // the workers may now overwrite these rows' entries in place
      if (more) {
        if (pl < 21) {
          double v = a[(c0 + 6 + cr) * ld + c0 + 6 + ce];
#pragma unroll
          for (int m = 0; m < 6; m++) v = fma(-s_xp[cr * 6 + m], s_xp[ce * 6 + m], v);
          s_corner[pl] = v;
        }
        __syncwarp();
        if (pl == 0) {
          double L[6][6], iv[6];
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) L[r][c] = c <= r ? s_corner[r * (r + 1) / 2 + c] : 0.0;
          if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            s_invk[par ^ 1][r] = iv[r]; invd[c0 + 6 + r] = iv[r];
#pragma unroll
            for (int c = 0; c < 6; c++) { s_Lkk[par ^ 1][r * 6 + c] = L[r][c]; if (c <= r) a[(c0 + 6 + r) * ld + c0 + 6 + c] = L[r][c]; }
          }
        }
      }
    } else {
      // rows below the block and the rhs: X L_kk^T = A (at most one row per thread: n + 1 - c0 - 6 <= 155 < 224)
      const int r = t;
      const bool have = r <= mrows;
      const int i = r < mrows ? c0 + 6 + r : n;
      double x[6];
      if (have) {
        const double* row = a + i * ld + c0;
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = row[c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) Xbuf[r * XS + c] = x[c];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (have) {
        double* row = a + i * ld + c0;
#pragma unroll
        for (int c = 0; c < 6; c++) row[c] = x[c];
      }
      // trailing update of rows c0+6 .. n-1 and the rhs row: 8x8 tiles (I >= J) over mrows + 1 rows x mrows columns
      if (mrows > 0) {
        const int NBt = (mrows + 1 + 7) >> 3, ntile = NBt * (NBt + 1) / 2;
        const int fr = lane >> 2, fk = lane & 3;
        // four tiles in flight per warp: a tile is a latency chain (fragment loads -> two dependent DMMAs -> read-modify-
        // write of the matrix), so independent tiles are interleaved by hand
        constexpr int UN = 4;
        int I = 0, J = warp;
        while (J > I) { J -= I + 1; I++; }
        for (int idx = warp; idx < ntile; idx += 7 * UN) {
          int tI[UN], tJ[UN];
          double d0[UN], d1[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            tI[u] = idx + 7 * u < ntile ? I : -1; tJ[u] = J;
            J += 7;
            while (J > I) { J -= I + 1; I++; }
          }
#pragma unroll
          for (int u = 0; u < UN; u++) {
            d0[u] = 0.0; d1[u] = 0.0;
            if (tI[u] >= 0) {                     // warp-uniform
              const double* xa = Xbuf + (8 * tI[u] + fr) * XS + fk;
              const double* xb = Xbuf + (8 * tJ[u] + fr) * XS + fk;
              dmma884(d0[u], d1[u], xa[0], xb[0]);
              dmma884(d0[u], d1[u], xa[4], xb[4]);
            }
          }
#pragma unroll
          for (int u = 0; u < UN; u++) {
            if (tI[u] < 0) continue;
            const int ti = 8 * tI[u] + fr, tk = 8 * tJ[u] + 2 * fk;
            if (ti >= 6 && ti <= mrows) {       // ti < 6: the next block's rows, whose corner belongs to the panel warp
              double* dst = a + (ti < mrows ? c0 + 6 + ti : n) * ld + c0 + 6 + tk;
              const int kmax = ti < mrows ? ti : mrows - 1;
              if (tk <= kmax) dst[0] -= d0[u];
              if (tk + 1 <= kmax) dst[1] -= d1[u];
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // backward substitution L^T x = z (z sits in row n), blocked
  double* z = a + n * ld;
  for (int kb = nf - 1; kb >= 0; kb--) {
    const int c0 = 6 * kb;
    if (t == 0) {
      double xb[6];
#pragma unroll
      for (int c = 5; c >= 0; c--) {
        double v = z[c0 + c];
#pragma unroll
        for (int m = 0; m < 6; m++) if (m > c) v = fma(-a[(c0 + m) * ld + c0 + c], xb[m], v);
        xb[c] = v * invd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; c++) z[c0 + c] = xb[c];
    }
    __syncthreads();
    for (int i = t; i < c0; i += nt) {
      double v = z[i];
#pragma unroll
      for (int c = 0; c < 6; c++) v = fma(-a[(c0 + c) * ld + i], z[c0 + c], v);
      z[i] = v;
    }
    __syncthreads();
  }
  const bool failed = s_fail != 0;
  for (int i = t; i < n; i += nt) rhs[i] = failed ? 0.0 : z[i];
  if (keep_factor) {
    for (int i = warp; i < n; i += nt >> 5)
      for (int j = lane; j <= i; j += 32) A[(size_t)i * n + j] = a[i * ld + j];
  }
  if (t == 0 && failed) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
}
#endif  // UBA_EMU

__global__ void __launch_bounds__(256) k_chol_diag(DevView V, int w, int j0) {
  __shared__ double a[NB][NB + 1];
  __shared__ int s_fail;
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* A = V.A + V.w_red_off[w];
  const int nb = min(NB, n - j0);
  const int tid = threadIdx.x;
  if (tid == 0) s_fail = 0;
  for (int e = tid; e < nb * nb; e += blockDim.x) { const int i = e / nb, j = e % nb; a[i][j] = A[(size_t)(j0 + i) * n + j0 + j]; }
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    if (tid == 0) {
      const double d = a[j][j];
      if (!(d > 0.0) || !isfinite(d)) { s_fail = 1; a[j][j] = 1.0; } else a[j][j] = sqrt(d);
    }
    __syncthreads();
    const double dinv = 1.0 / a[j][j];
    for (int i = j + 1 + tid; i < nb; i += blockDim.x) a[i][j] *= dinv;
    __syncthreads();
    const int m = nb - j - 1;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int ii = e / m, kk = e % m;
      if (kk <= ii) a[j + 1 + ii][j + 1 + kk] -= a[j + 1 + ii][j] * a[j + 1 + kk][j];
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += blockDim.x) { const int i = e / nb, j = e % nb; if (j <= i) A[(size_t)(j0 + i) * n + j0 + j] = a[i][j]; }
  if (tid == 0 && s_fail) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
}

// rows below the panel: X L^T = A  ->  X   (each CTA: NB rows)
__global__ void __launch_bounds__(256) k_chol_trsm(DevView V, int w, int j0) {
  __shared__ double l[NB][NB + 1];
  __shared__ double x[NB][NB + 1];
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* A = V.A + V.w_red_off[w];
  const int nb = min(NB, n - j0);
  const int i0 = j0 + nb + blockIdx.x * NB;
  if (i0 >= n) return;
  const int nr = min(NB, n - i0);
  const int tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += blockDim.x) { const int i = e / nb, j = e % nb; l[i][j] = A[(size_t)(j0 + i) * n + j0 + j]; }
  for (int e = tid; e < nr * nb; e += blockDim.x) { const int i = e / nb, j = e % nb; x[i][j] = A[(size_t)(i0 + i) * n + j0 + j]; }
  __syncthreads();
  // each thread (row i < nr) solves its own row sequentially: x[i][:] L^T = a[i][:]
  if (tid < nr) {
    for (int j = 0; j < nb; j++) {
      double s = x[tid][j];
      for (int k = 0; k < j; k++) s -= x[tid][k] * l[j][k];
      x[tid][j] = s / l[j][j];
    }
  }
  __syncthreads();
  for (int e = tid; e < nr * nb; e += blockDim.x) { const int i = e / nb, j = e % nb; A[(size_t)(i0 + i) * n + j0 + j] = x[i][j]; }
}

// trailing update: A[bi][bj] -= L[bi][panel] L[bj][panel]^T for tiles bi >= bj below the panel
__global__ void __launch_bounds__(256) k_chol_update(DevView V, int w, int j0) {
  __shared__ double li[NB][NB + 1];
  __shared__ double lj[NB][NB + 1];
  const int n = 6 * (V.w_free_off[w + 1] - V.w_free_off[w]);
  double* A = V.A + V.w_red_off[w];
  const int nb = min(NB, n - j0);
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int i0 = j0 + nb + bi * NB, c0 = j0 + nb + bj * NB;
  if (i0 >= n) return;
  const int nr = min(NB, n - i0), nc = min(NB, n - c0);
  const int tid = threadIdx.x;
  for (int e = tid; e < nr * nb; e += blockDim.x) { const int i = e / nb, k = e % nb; li[i][k] = A[(size_t)(i0 + i) * n + j0 + k]; }
  for (int e = tid; e < nc * nb; e += blockDim.x) { const int i = e / nb, k = e % nb; lj[i][k] = A[(size_t)(c0 + i) * n + j0 + k]; }
  __syncthreads();
  for (int e = tid; e < nr * nc; e += blockDim.x) {
    const int i = e / nc, j = e % nc;
    if (c0 + j > i0 + i) continue;
    double s = 0.0;
    for (int k = 0; k < nb; k++) s += li[i][k] * lj[j][k];
    A[(size_t)(i0 + i) * n + c0 + j] -= s;
  }
}


// ---- banded reduced systems (large windows whose camera co-visibility is a band: c4, c5) ------
// The matrix is never materialised densely: k_assemble writes the damped band (Ab) from the Schur accumulator, the
// camera blocks B and the LM damping; rows stream through a ring of kBandRing rows in shared memory.
constexpr int kBandRing = 126;  // multiple of 6: a block of rows never straddles the wrap

// Band Cholesky by 6-wide block columns (right-looking) with LOOKAHEAD, one CTA, beta >= 11.  Per block column:
// serial 6x6 factor of the diagonal block; thread per row below the block inside the band (+ one for the rhs "row"):
// X L_kk^T = A; trailing update of the band window, rhs update, factor rows out to global memory
// (Lt[i] = {1/L_ii, L_{i,i-1}, ..., L_{i,i-beta}}).  The band was assembled by k_assemble (Ab); rows stream through a ring
// of kBandRing rows in shared memory.  The serial part of a block step does not stall
// the CTA: warp 7 is the "panel" warp: during step k it computes, for the rows of block k+1 only, their triangular solve against L_kk(k), the update of the next diagonal block, and its
// 6x6 factor L_kk(k+1) — while warps 0..6 do step k's triangular solves for all rows (into Xbuf, not in place,
// so the panel warp still sees the untouched entries) and the trailing update of the band window (minus the
// corner the panel warp owns).  One CTA-wide barrier per block step instead of three.
template <int PER>
__global__ void __launch_bounds__(256) k_chol_banded_la(DevView V, int w, int beta) {
  extern __shared__ double sm[];
  const WinState* st = &V.ws[w];
  if (st->done) return;
  constexpr int NWORK = 224;                  // worker threads (warps 0..6); warp 7 is the panel warp
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  const int bw1 = beta + 1;
  const int ring_size = kBandRing * bw1;
  double* ring = sm;                          // [kBandRing][bw1]: row i holds A[i][i-beta .. i]
  double* y = ring + ring_size;               // [n + beta + 7] (tail zero-padded)
  double* Xbuf = y + (n + beta + 8);          // [beta][6] triangular-solve results of the current step
  __shared__ int s_fail;
  __shared__ double s_Lkk[2][36], s_invk[2][6], s_z[6], s_xp[36], s_corner[21];
  double* rhs = V.rhs + (size_t)6 * f0;
  double* Lt = V.A + V.w_red_off[w];
  const double* Ab = Lt + (size_t)2 * n * bw1;
  const int t = threadIdx.x, nt = blockDim.x;
  const bool panel = t >= NWORK;
  const int pl = t - NWORK;                   // lane of the panel warp
  if (t == 0) s_fail = 0;
  for (int i = t; i < n + beta + 7; i += nt) y[i] = i < n ? rhs[i] : 0.0;
  for (int e = t; e < kBandRing * bw1; e += nt) { const int i = e / bw1; ring[e] = i < n ? Ab[e] : 0.0; }
  // trailing-update pairs (ti >= tk) over the beta rows below the block; the corner (both rows inside the
  // next block) belongs to the panel warp
  const int npairs = beta * (beta + 1) / 2;
  int pti[PER], ptk[PER];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int e = t + q * NWORK;
    int ti = -1, tk = 0;
    if (!panel && e < npairs) {
      int d0 = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((d0 + 1) * (d0 + 2) / 2 <= e) d0++;
      while (d0 * (d0 + 1) / 2 > e) d0--;
      ti = d0; tk = e - d0 * (d0 + 1) / 2;
      if (ti < 6) ti = -1;                    // corner pair
    }
    pti[q] = ti; ptk[q] = tk;
  }
  const int nblk = n / 6;
#ifdef UBA_BAND_TIMING
  long long tmx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tqx;
#define LCLK(v) asm volatile("mov.u64 %0, %%clock64;" : "=l"(v) :: "memory")
#define LT0() LCLK(tqx);
#define LT(k) { long long now_; LCLK(now_); tmx[k] += now_ - tqx; tqx = now_; }
#else
#define LT0()
#define LT(k)
#endif
  __syncthreads();
  // prologue: factor of block 0
  if (t == NWORK) {
    double L[6][6], iv[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) L[r][c] = c <= r ? ring[r * bw1 + beta - r + c] : 0.0;
    if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      s_invk[0][r] = iv[r];
#pragma unroll
      for (int c = 0; c < 6; c++) s_Lkk[0][r * 6 + c] = L[r][c];
    }
  }
  __syncthreads();
  int o0 = 0;                                 // ring offset of row c0 (blocks never straddle the wrap)
  for (int kb = 0; kb < nblk; kb++) {
    const int c0 = 6 * kb, par = kb & 1;
    const double* Lk = s_Lkk[par];
    const double* ivk = s_invk[par];
    LT0()
    if (panel) {
      // ---- panel warp: next diagonal block (k+1) --------------------------------------------------
      if (kb + 1 < nblk) {
        int on = o0 + 6 * bw1; if (on >= ring_size) on -= ring_size;   // ring offset of row c0 + 6
        if (pl < 6) {
          // row c0+6+pl against L_kk(k): entries (i, c0+c) at column offset beta - 6 - pl + c (>= 0: beta >= 11)
          const double* row = ring + on + pl * bw1 + (beta - 6 - pl);
          double x[6];
#pragma unroll
          for (int c = 0; c < 6; c++) x[c] = row[c];
#pragma unroll
          for (int c = 0; c < 6; c++) {
            x[c] *= ivk[c];
#pragma unroll
            for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
          }
#pragma unroll
          for (int c = 0; c < 6; c++) s_xp[pl * 6 + c] = x[c];
        }
        __syncwarp();
        LT(0)
        if (pl < 21) {
          int r = 0, e = pl;
          while (e > r) { e -= r + 1; r++; }   // pl -> (r, c = e), c <= r
          double v = ring[on + r * bw1 + beta - r + e];
#pragma unroll
          for (int m = 0; m < 6; m++) v = fma(-s_xp[r * 6 + m], s_xp[e * 6 + m], v);
          s_corner[pl] = v;
        }
        __syncwarp();
        LT(1)
        if (pl == 0) {
          double L[6][6], iv[6];
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) L[r][c] = c <= r ? s_corner[r * (r + 1) / 2 + c] : 0.0;
          if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            s_invk[par ^ 1][r] = iv[r];
#pragma unroll
            for (int c = 0; c < 6; c++) s_Lkk[par ^ 1][r * 6 + c] = L[r][c];
          }
        }
        LT(2)
      }
    } else {
      // ---- workers: ring reload prefetch -------------------------------------------------------------
      constexpr int kPre = (30 * (kBandMaxBeta + 1) + NWORK - 1) / NWORK;
      double pre[kPre];
      const bool reload = kb > 0 && (kb % 5) == 0;
      if (reload) {
        const int r0 = c0 + kBandRing - 30;
#pragma unroll
        for (int q = 0; q < kPre; q++) {
          const int e = t + q * NWORK;
          const int i = r0 + e / bw1;
          pre[q] = (e < 30 * bw1 && i < n) ? Ab[(size_t)r0 * bw1 + e] : 0.0;
        }
      }
      // ---- (3) rows below the block inside the band, and the rhs: X L_kk^T = A --------------------------
      // all lanes of warp 0 on one path (clamped loads + select); the rhs rides along as one more row in another warp
      if (t < beta) {
        int orow = o0 + (6 + t) * bw1; if (orow >= ring_size) orow -= ring_size;
        const double* row = ring + orow;
        const int base = beta - 6 - t;        // column offset of (i, c0); entries with base + c < 0 lie outside the band
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) { const int ix = base + c; const double v = row[ix < 0 ? 0 : ix]; x[c] = ix >= 0 ? v : 0.0; }
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) Xbuf[t * 6 + c] = x[c];
      } else if (t == 192) {
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = y[c0 + c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) { s_z[c] = x[c]; y[c0 + c] = x[c]; }
      }
      LT(3)
      asm volatile("bar.sync 1, %0;" ::"n"(NWORK));
      LT(4)
      // ---- (4) trailing update of the band window (corner excluded), rhs update, factor rows out ------
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int ti = pti[q], tk = ptk[q];
        if (ti >= 0) {
          int oi = o0 + (6 + ti) * bw1; if (oi >= ring_size) oi -= ring_size;
          const double* xi = Xbuf + ti * 6;
          const double* xk = Xbuf + tk * 6;
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < 6; c++) acc = fma(xi[c], xk[c], acc);
          ring[oi + (beta - ti + tk)] -= acc;
        }
      }
      if (t >= 64 && t < 64 + beta) {
        const int tt = t - 64, i = c0 + 6 + tt;
        const int base = beta - 6 - tt;
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) {
          const double l = Xbuf[tt * 6 + c];
          acc = fma(l, s_z[c], acc);
          if (base + c >= 0 && i < n) Lt[(size_t)i * bw1 + (6 + tt - c)] = l;
        }
        y[i] -= acc;
      } else if (t >= 160 && t < 166) {
        const int r = t - 160;
        Lt[(size_t)(c0 + r) * bw1] = ivk[r];
        for (int c = 0; c < r; c++) Lt[(size_t)(c0 + r) * bw1 + (r - c)] = Lk[r * 6 + c];
      }
      if (reload) {
        const int r0 = c0 + kBandRing - 30;
#pragma unroll
        for (int q = 0; q < kPre; q++) {
          const int e = t + q * NWORK;
          if (e < 30 * bw1) ring[((r0 + e / bw1) % kBandRing) * bw1 + e % bw1] = pre[q];
        }
      }
      LT(5)
    }
    __syncthreads();
    LT(6)
    o0 += 6 * bw1; if (o0 >= ring_size) o0 -= ring_size;
  }
#ifdef UBA_BAND_TIMING
  if (t == 0 || t == 64 || t == 100 || t == 224) { for (int k = 0; k < 8; k++) V.Zbuf[(t == 0 ? 0 : t == 64 ? 8 : t == 100 ? 16 : 24) + k] = (double)tmx[k]; }
#endif
  // backward substitution L^T x = z, blocked, with lookahead: warp 7 solves the 6x6 diagonal block and applies its
  // update to the rows of the NEXT block itself; the other threads apply the update to the remaining band rows
  // one block behind (one CTA barrier per block).  Factor rows are staged through shared memory in chunks.
  constexpr int kChunk = 126;
  __shared__ double s_xb[2][6];
  for (int i1 = n; i1 > 0; i1 -= kChunk) {
    const int i0 = max(0, i1 - kChunk);
    __syncthreads();
    for (int e = t; e < (i1 - i0) * bw1; e += nt) {
      const int i = i0 + e / bw1, c = e % bw1;
      ring[e] = (c <= i) ? Lt[(size_t)i0 * bw1 + e] : 0.0;
    }
    __syncthreads();
    int it = 0;
    for (int c0 = i1 - 6; c0 >= i0; c0 -= 6, it++) {
      const double* blk = ring + (c0 - i0) * bw1;   // row c0 + m at blk + m*bw1: [0] = 1/L, [d] = L[c0+m][c0+m-d]
      const int par = it & 1;
      if (panel) {
        if (pl == 0) {
          double xb[6];
#pragma unroll
          for (int c = 0; c < 6; c++) xb[c] = y[c0 + c];
#pragma unroll
          for (int c = 5; c >= 0; c--) {
            xb[c] *= blk[c * bw1];
#pragma unroll
            for (int m = 0; m < 6; m++) if (m < c) xb[m] = fma(-blk[c * bw1 + (c - m)], xb[c], xb[m]);
          }
#pragma unroll
          for (int c = 0; c < 6; c++) { y[c0 + c] = xb[c]; s_xb[par][c] = xb[c]; }
        }
        __syncwarp();
        if (pl < 6) {
          // rows of the next block (c0 - 6 + pl): this block's update (inside the band since beta >= 11) and the
          // previous block's (cp = c0 + 6), which the workers leave to this warp so that no row has two writers
          const int j = c0 - 6 + pl;
          if (j >= 0) {
            double v = y[j], v2 = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) v = fma(-blk[c * bw1 + (c0 + c - j)], s_xb[par][c], v);
            if (it > 0) {
              const double* blkp = blk + 6 * bw1;
#pragma unroll
              for (int c = 0; c < 6; c++) { const int d = c0 + 6 + c - j; if (d <= beta) v2 = fma(blkp[c * bw1 + d], s_xb[par ^ 1][c], v2); }
            }
            y[j] = v - v2;
          }
        }
        __syncwarp();
      } else if (it > 0 && t < beta - 12) {
        // workers, one block behind: rows [cp - beta, cp - 12) of the previous block cp = c0 + 6
        const int cp = c0 + 6;
        const double* blkp = blk + 6 * bw1;
        const int j = cp - 13 - t;
        if (j >= 0) {
          double v = y[j];
#pragma unroll
          for (int c = 0; c < 6; c++) { const int d = cp + c - j; if (d <= beta) v = fma(-blkp[c * bw1 + d], s_xb[par ^ 1][c], v); }
          y[j] = v;
        }
      }
      __syncthreads();
    }
    // drain: the last block of the chunk still owes its update to the rows beyond the next block
    if (!panel && t < beta - 6) {
      const int cp = i0;
      const double* blkp = ring;
      const int j = cp - 7 - t;
      if (j >= 0) {
        double v = y[j];
#pragma unroll
        for (int c = 0; c < 6; c++) { const int d = cp + c - j; if (d <= beta) v = fma(-blkp[c * bw1 + d], s_xb[(it - 1) & 1][c], v); }
        y[j] = v;
      }
    }
  }
  __syncthreads();
  const bool failed = s_fail != 0;
  for (int i = t; i < n; i += nt) rhs[i] = failed ? 0.0 : y[i];
  if (t == 0 && failed) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
}

#ifndef UBA_EMU
// Two-sided band Cholesky on a CLUSTER OF TWO CTAs (two SMs), beta in [11, 35]: CTA 0 eliminates block columns
// [0, m) top-down, CTA 1 eliminates the rows below the separator bottom-up (top-down on the reversed matrix), each
// with all 256 threads like k_chol_banded_la (panel warp 7 + 224 workers; trailing update on DMMA tiles, ring refilled
// every step).  After the forward loops CTA 0 reads CTA 1's separator corner and rhs through distributed shared memory,
// adds them to its own ring rows, continues its elimination over the separator (sw / 6 more steps), back-substitutes the
// separator and writes its solution into both CTAs' y; both then back-substitute their half concurrently with a sweep
// whose triangular solves are precomputed per chunk (see the backward section).  Sequential depth: n/12 + sw/6 block
// steps on each SM instead of n/6 on one.
struct BandHalf {
  int dir;        // 0: natural order, 1: reversed
  int nh;         // rows of the half's local system (eliminated rows + separator)
  int ne;         // rows eliminated by this half (multiple of 6)
};
constexpr int kC2ChunkBlocks = 34;      // blocks per chunk of the backward sweep of k_chol_banded_c2
__host__ __device__ constexpr size_t c2_backward_doubles(int beta) {
  return (size_t)(kC2ChunkBlocks + (beta + 6) / 6 + 1) * 6 * (beta + 1) + (size_t)kC2ChunkBlocks * ((beta + 6) / 6 + 1) * 36;
}

template <int PER, bool PIPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256) k_chol_banded_c2(DevView V, int w, int beta) {
  extern __shared__ double sm[];
  const WinState* st = &V.ws[w];
  if (st->done) return;
  constexpr int NH = 256, NWORKH = 224;
  constexpr int kPollThread = 200;              // a worker without side duties (pipelined solve: polls the row flags)
  const int half = (int)(blockIdx.x & 1);
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  const int bw1 = beta + 1;
  const int sw = ((beta + 1 + 5) / 6) * 6;    // separator rows
  const int m = (((n - sw) / 2) / 6) * 6;     // rows eliminated by the top half
  const int ring_size = kBandRing * bw1;
  const int t = threadIdx.x;
  BandHalf H;
  H.dir = half; H.ne = half == 0 ? m : n - m - sw; H.nh = H.ne + sw;
  const int ylen = (n - m) + beta + 8;        // >= nh of either half + tail
  double* ring = sm;
  double* y = ring + ring_size;
  double* Xbuf = y + ylen;
  constexpr int XS = 9;                       // row stride of Xbuf [40][XS]: columns 6..8 and rows >= beta stay zero (DMMA padding)
  __shared__ int s_fail;
  __shared__ double s_Lkk[2][36], s_invk[2][6], s_z[6], s_xp[36], s_corner[21];
  double* rhs = V.rhs + (size_t)6 * f0;
  double* A0 = V.A + V.w_red_off[w];
  double* Lt = A0 + (size_t)half * n * bw1;   // this half's factor rows (local row numbering)
  const double* Ab = A0 + (size_t)2 * n * bw1;
  const bool panel = t >= NWORKH;
  const int pl = t - NWORKH;
#ifdef UBA_BAND_TIMING
  long long tph[8]; long long tgl[8]; int nph = 0; long long busy = 0; long long bseg[3] = {0, 0, 0}; long long brole = 0; long long seg[6] = {0, 0, 0, 0, 0, 0}; long long tq = 0;
#define SG(i) { const long long now_ = clock64(); seg[i] += now_ - tq; tq = now_; }
#define PH() { if (t == 0) { tph[nph] = clock64(); long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); tgl[nph] = g_; nph++; } }
#else
#define PH()
#endif
  PH()
  if (t == 0) s_fail = 0;
  // Pipelined solve (V.pipe_on): the lineariser is still running; every row of the band (and of the rhs) is assembled by the
  // lineariser CTA that completed its camera, which then raises row_ready[camera].  Whoever loads a row waits for its flag
  // (bounded: a flag that never comes fails the solve instead of hanging the device) and reads through L2.
  constexpr bool pipe = PIPE;                  // (a separate instantiation: the plain solve keeps its cached loads and up-front rhs)
  auto wait_row = [&](int io) {
    if (!pipe) return;
    const volatile int32_t* fl = &V.row_ready[io / 6];
    if (*fl == 0) {
      long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while (*fl == 0) {
        __nanosleep(100);
        long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ll) { s_fail = 1; break; }
      }
    }
    __threadfence();
  };
  auto ab_load = [&](int io, int c) -> double { wait_row(io); return pipe ? __ldcg(&Ab[(size_t)io * bw1 + c]) : Ab[(size_t)io * bw1 + c]; };
  auto rhs_load = [&](int io) -> double { wait_row(io); return pipe ? __ldcg(&rhs[io]) : rhs[io]; };
  auto band_entry = [&](int i, int c) -> double {
    if (i >= H.nh || i - beta + c < 0) return 0.0;
    return H.dir == 0 ? ab_load(i, c) : ab_load(n - 1 - i + beta - c, c);
  };
  // the rhs rows travel with the band rows: the first kBandRing now, six more with every ring refill
  for (int i = t; i < ylen; i += NH) y[i] = (i < H.nh && (!pipe || i < kBandRing)) ? rhs_load(H.dir == 0 ? i : n - 1 - i) : 0.0;
  // global -> shared staging in batches of 8 independent loads per thread: a plain `dst[e] = src[e]` loop is compiled as one
  // load-store pair after the other (shared stores may alias generic loads), i.e. one memory round trip per element
  {
    constexpr int U = 8;
    for (int base = 0; base < kBandRing * bw1; base += NH * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const int e = base + t + u * NH; v[u] = e < kBandRing * bw1 ? band_entry(e / bw1, e % bw1) : 0.0; }
#pragma unroll
      for (int u = 0; u < U; u++) { const int e = base + t + u * NH; if (e < kBandRing * bw1) ring[e] = v[u]; }
    }
  }
  for (int i = t; i < 40 * XS; i += NH) Xbuf[i] = 0.0;
  if (pipe && t == kPollThread && kBandRing < H.nh) wait_row(H.dir == 0 ? kBandRing : n - 1 - (kBandRing + 5));   // rows entering at step 0
  // trailing update W -= X X^T on the FP64 tensor-core path: 8x8 tiles (I >= J) of the beta x beta window, dealt to
  // the 7 worker warps; the X fragments of a tile are 4 shared-memory loads per lane instead of 12 per scalar pair
  constexpr int kTW = 3;                      // tiles per worker warp (15 tiles for beta = 35, 10 for beta = 29)
  const int NBt = (beta + 7) / 8, ntile = NBt * (NBt + 1) / 2;
  int tI[kTW], tJ[kTW];
#pragma unroll
  for (int q = 0; q < kTW; q++) {
    int idx = (t >> 5) + 7 * q, I = 0;
    if (panel || idx >= ntile) { tI[q] = -1; tJ[q] = 0; continue; }
    while (idx > I) { idx -= I + 1; I++; }    // row-major over the lower triangle: (I, J = idx), J <= I
    tI[q] = I; tJ[q] = idx;
  }
  int cr = 0, ce = pl;                        // corner entry of panel lane pl: (cr, ce), ce <= cr
  while (ce > cr) { ce -= cr + 1; cr++; }
  const bool rf_on = !panel && t < 6 * bw1;   // ring refill role: entry (rf_row, rf_col) of the 6 incoming rows
  const int rf_row = t / bw1, rf_col = t - rf_row * bw1;
  int dg_r = 0, dg_c = t - (NWORKH - 21);     // the last 21 workers: entry (dg_r, dg_c) of the diagonal block's factor
  if (dg_c >= 0 && dg_c < 21) { while (dg_c > dg_r) { dg_c -= dg_r + 1; dg_r++; } } else dg_c = 0;
  const int nblk = H.ne / 6;
  __syncthreads();
  if (t == NWORKH) {                          // prologue: factor of block 0
    double L[6][6], iv[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) L[r][c] = c <= r ? ring[r * bw1 + beta - r + c] : 0.0;
    if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      s_invk[0][r] = iv[r];
#pragma unroll
      for (int c = 0; c < 6; c++) s_Lkk[0][r * 6 + c] = L[r][c];
    }
  }
  __syncthreads();
  int o0 = 0;
  PH()
  // block steps [kb0, kb1).  At the end of a half (tail == false) the block after the last eliminated one is the
  // first separator block: it still gets its corner update (written back to the ring) but is not factored here.
  // At the end of the separator (tail == true) there is no next block.
  auto forward = [&](int kb0, int kb1, bool tail) {
  for (int kb = kb0; kb < kb1; kb++) {
#ifdef UBA_BAND_TIMING
    const long long tb0 = clock64(); tq = tb0;
#endif
    const int c0 = 6 * kb, par = kb & 1;
    const double* Lk = s_Lkk[par];
    const double* ivk = s_invk[par];
    if (panel && !(tail && kb + 1 == kb1)) {
      const bool last = kb + 1 == kb1;
      int on = o0 + 6 * bw1; if (on >= ring_size) on -= ring_size;
      if (pl < 6) {
        const double* row = ring + on + pl * bw1 + (beta - 6 - pl);
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = row[c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int mm = 0; mm < 6; mm++) if (mm > c) x[mm] = fma(-x[c], Lk[mm * 6 + c], x[mm]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) s_xp[pl * 6 + c] = x[c];
      }
      __syncwarp();
      if (pl < 21) {
        double v = ring[on + cr * bw1 + beta - cr + ce];
#pragma unroll
        for (int mm = 0; mm < 6; mm++) v = fma(-s_xp[cr * 6 + mm], s_xp[ce * 6 + mm], v);
        if (last) ring[on + cr * bw1 + beta - cr + ce] = v;
        s_corner[pl] = v;
      }
      __syncwarp();
      if (pl == 0 && !last) {
        double L[6][6], iv[6];
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) L[r][c] = c <= r ? s_corner[r * (r + 1) / 2 + c] : 0.0;
        if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
        for (int r = 0; r < 6; r++) {
          s_invk[par ^ 1][r] = iv[r];
#pragma unroll
          for (int c = 0; c < 6; c++) s_Lkk[par ^ 1][r * 6 + c] = L[r][c];
        }
      }
    } else if (!panel) {
      // Ring refill, every step: the 6 rows of the block that was eliminated in the previous step are dead; the rows 126
      // further down take their slots.  One entry per thread (6 (beta+1) <= 216), row / column of the entry fixed per
      // thread (no index arithmetic in the loop), loaded here and stored at the end of the step.
      double pre = 0.0, prey = 0.0;
      if (rf_on) {
        const int i = c0 + kBandRing + rf_row;
        if (i < H.nh) {                          // (pipelined: the previous step's poll made sure these rows are assembled)
          const double* src = &Ab[(size_t)(H.dir == 0 ? i : n - 1 - i + beta - rf_col) * bw1 + rf_col];
          pre = pipe ? __ldcg(src) : *src;
          if (pipe && rf_col == 0) prey = __ldcg(&rhs[H.dir == 0 ? i : n - 1 - i]);
        }
      }
      // pipelined: one worker polls the flag of the camera whose rows enter the ring at the NEXT step; the load is issued here
      // and looked at only at the end of the step, so a flag that is already up costs nothing
      int nxt_flag = 1;
      const volatile int32_t* nxt_ptr = nullptr;
      if (pipe && t == kPollThread) {
        const int i1 = c0 + 6 + kBandRing + 5;   // last row entering at the next step
        if (i1 - 5 < H.nh) {
          const int cam = H.dir == 0 ? (i1 - 5) / 6 : (n - 1 - i1) / 6;
          if (cam >= 0 && cam < n / 6) { nxt_ptr = &V.row_ready[cam]; nxt_flag = *nxt_ptr; }
        }
      }
#ifdef UBA_BAND_TIMING
      SG(0)
#endif
      // Row solves X L_kk^T = A in warp 0, all lanes on one path (clamped loads + select: divergent loads cost
      // 2x here, tests/cuda/panel_bench.cu); the rhs rides along as one more row in another warp.
      if (t < beta) {
        int orow = o0 + (6 + t) * bw1; if (orow >= ring_size) orow -= ring_size;
        const double* row = ring + orow;
        const int basec = beta - 6 - t;
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) { const int ix = basec + c; const double v = row[ix < 0 ? 0 : ix]; x[c] = ix >= 0 ? v : 0.0; }
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int mm = 0; mm < 6; mm++) if (mm > c) x[mm] = fma(-x[c], Lk[mm * 6 + c], x[mm]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) Xbuf[t * XS + c] = x[c];
      } else if (t == 192) {
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = y[c0 + c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int mm = 0; mm < 6; mm++) if (mm > c) x[mm] = fma(-x[c], Lk[mm * 6 + c], x[mm]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) { s_z[c] = x[c]; y[c0 + c] = x[c]; }
      }
#ifdef UBA_BAND_TIMING
      SG(1)
#endif
      asm volatile("bar.sync 1, %0;" ::"n"(NWORKH));
#ifdef UBA_BAND_TIMING
      SG(2)
#endif
      {
        const int fr = (t & 31) >> 2, fk = t & 3;
#pragma unroll
        for (int q = 0; q < kTW; q++) {
          if (tI[q] < 0) continue;
          const double* xa = Xbuf + (8 * tI[q] + fr) * XS + fk;
          const double* xb = Xbuf + (8 * tJ[q] + fr) * XS + fk;
          double d0 = 0.0, d1 = 0.0;
          dmma884(d0, d1, xa[0], xb[0]);
          dmma884(d0, d1, xa[4], xb[4]);
          const int ti = 8 * tI[q] + fr, tk = 8 * tJ[q] + 2 * fk;
          if (ti >= 6 && ti < beta) {           // rows inside the window; the corner (ti < 6) belongs to the panel warp
            int oi = o0 + (6 + ti) * bw1; if (oi >= ring_size) oi -= ring_size;
            double* dst = ring + oi + (beta - ti + tk);
            if (tk <= ti) dst[0] -= d0;
            if (tk + 1 <= ti) dst[1] -= d1;
          }
        }
      }
#ifdef UBA_BAND_TIMING
      SG(3)
#endif
      // side duties.  Factor rows go out to global memory one value per thread, six consecutive lanes per row (the six
      // values of a row are contiguous there): a thread-per-row store walks 29 sectors per instruction and made its
      // warp the step's straggler.  The diagonal block's factor goes out the same way, the rhs update sits in warp 3.
      if (t < 6 * beta) {
        const int tt = t / 6, c = t - 6 * tt, i = c0 + 6 + tt;
        if (i < H.nh && beta - 6 - tt + c >= 0) Lt[(size_t)i * bw1 + (6 + tt - c)] = Xbuf[tt * XS + c];
      }
      if (t >= NWORKH - 21) Lt[(size_t)(c0 + dg_r) * bw1 + (dg_r - dg_c)] = (dg_r == dg_c) ? ivk[dg_r] : Lk[dg_r * 6 + dg_c];
      if (t >= 96 && t < 96 + beta) {
        const int tt = t - 96, i = c0 + 6 + tt;
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fma(Xbuf[tt * XS + c], s_z[c], acc);
        y[i] -= acc;
      }
#ifdef UBA_BAND_TIMING
      SG(4)
#endif
      if (rf_on) {
        ring[o0 + rf_row * bw1 + rf_col] = pre;   // slot of row c0 + rf_row, dead since the previous step
        if (pipe && rf_col == 0 && c0 + kBandRing + rf_row < H.nh) y[c0 + kBandRing + rf_row] = prey;
      }
      if (nxt_ptr && nxt_flag == 0) {           // the lineariser has not finished that camera yet: wait here (bounded)
        long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*nxt_ptr == 0) {
          __nanosleep(200);
          long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > 2000000000ll) { s_fail = 1; break; }
        }
      }
      if (nxt_ptr) __threadfence();
    }
#ifdef UBA_BAND_TIMING
    if (!panel) SG(5)
    busy += clock64() - tb0;
#endif
    __syncthreads();
    o0 += 6 * bw1; if (o0 >= ring_size) o0 -= ring_size;
  }
  };
  forward(0, nblk, false);
  PH()
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  PH()
  // ---- separator (CTA 0): S = T + B - A_ss, rhs = y_top + y_bot - b_s.  The merged rows simply continue CTA 0's
  // block elimination (sw / 6 more steps), followed by a short backward substitution inside the separator ----
  if (half == 0) {
    const double* ring1 = cluster.map_shared_rank(ring, 1);
    double* y1 = cluster.map_shared_rank(y, 1);
    int* fail1 = cluster.map_shared_rank(&s_fail, 1);
    const int ne0 = m, ne1 = n - m - sw;
    for (int e = t; e < sw * sw; e += NH) {
      const int a = e / sw, b = e % sw;
      if (b > a || a - b > beta) continue;
      const int it = ne0 + a, kt = ne0 + b;                          // top: local = original numbering
      const int ib = ne1 + (sw - 1 - b), kbm = ne1 + (sw - 1 - a);   // bottom (reversed): row >= col
      ring[(it % kBandRing) * bw1 + (kt - it + beta)] += ring1[(ib % kBandRing) * bw1 + (kbm - ib + beta)] - ab_load(m + a, b - a + beta);
    }
    for (int a = t; a < sw; a += NH) y[ne0 + a] += y1[ne1 + (sw - 1 - a)] - rhs_load(m + a);
    __syncthreads();
    if (t == NWORKH) {                        // factor of the first separator block
      const int ob = (ne0 % kBandRing) * bw1;
      double L[6][6], iv[6];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) L[r][c] = c <= r ? ring[ob + r * bw1 + beta - r + c] : 0.0;
      if (!chol6(L, iv)) s_fail = 1;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        s_invk[nblk & 1][r] = iv[r];
#pragma unroll
        for (int c = 0; c < 6; c++) s_Lkk[nblk & 1][r * 6 + c] = L[r][c];
      }
    }
    __syncthreads();
    forward(nblk, nblk + sw / 6, true);
    // backward substitution inside the separator (factor rows staged from global memory)
    {
      constexpr int U = 4;
      const double* srcp = Lt + (size_t)ne0 * bw1;
      for (int base = 0; base < sw * bw1; base += NH * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int e = base + t + u * NH; v[u] = e < sw * bw1 ? srcp[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < U; u++) { const int e = base + t + u * NH; if (e < sw * bw1) ring[e] = v[u]; }
      }
    }
    __syncthreads();
    for (int c0 = sw - 6; c0 >= 0; c0 -= 6) {
      const double* blk = ring + c0 * bw1;
      if (t == 0) {
        double xb[6];
#pragma unroll
        for (int c = 0; c < 6; c++) xb[c] = y[ne0 + c0 + c];
#pragma unroll
        for (int c = 5; c >= 0; c--) {
          xb[c] *= blk[c * bw1];
#pragma unroll
          for (int mm = 0; mm < 6; mm++) if (mm < c) xb[mm] = fma(-blk[c * bw1 + (c - mm)], xb[c], xb[mm]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) { y[ne0 + c0 + c] = xb[c]; s_z[c] = xb[c]; }
      }
      __syncthreads();
      if (t < c0) {                           // separator rows above the block
        const int j = c0 - 1 - t;
        double v = y[ne0 + j];
#pragma unroll
        for (int c = 0; c < 6; c++) { const int d = c0 + c - j; if (d <= beta) v = fma(-blk[c * bw1 + d], s_z[c], v); }
        y[ne0 + j] = v;
      }
      __syncthreads();
    }
    // the separator solution becomes known boundary values of both halves
    for (int a = t; a < sw; a += NH) y1[ne1 + (sw - 1 - a)] = y[ne0 + a];
    if (t == 0) { const int f = s_fail | *fail1; s_fail = f; *fail1 = f; }
  }
  PH()
  cluster.sync();
  PH()
  // ---- backward substitution of this half (local numbering); separator rows are known ----
  // L^T x = z by 6-row blocks, x_b = L_bb^-T (z_b - sum_{j=1..NJ} L_{b+j,b}^T x_{b+j}).  The solve is taken OFF the sweep's dependent
  // chain by precomputing, per chunk of blocks and in parallel over all threads, M_b = L_bb^-T and P^j_b = M_b L_{b+j,b}^T:
  //   x_b = [M_b z_b - sum_{j>=2} P^j_b x_{b+j}] - P^1_b x_{b+1} = u_b - P^1_b x_{b+1}.
  // One warp prepares u for the NEXT block (its inputs are known one step ahead: 6 x NJ lanes, a 6-term dot product each, shuffle
  // reduction) while six lanes of the panel warp finish the current one (one 6-term dot product per lane).  Per block the chain is
  // a dot product and a barrier instead of a 12-deep triangular solve, a dependent 12-term update and a barrier.
  {
    const int NJ = (beta + 6) / 6;              // block sub-diagonals inside the band (5 for beta = 29, 6 for 35)
    const int NG = NJ + 1;                      // groups per block: M, P^1 .. P^NJ
    double* Lb = Xbuf + 40 * XS + 8;            // [(kC2ChunkBlocks + NJ + 1) * 6][bw1] staged factor rows
    double* MP = Lb + (size_t)(kC2ChunkBlocks + NJ + 1) * 6 * bw1;   // [kC2ChunkBlocks][NG][6][6]
    __shared__ double s_part[2][6][6];          // [block parity][group][row]
    const int nbu = H.ne / 6;                   // unknown blocks 0 .. nbu-1
    for (int bH = nbu - 1; bH >= 0; bH -= kC2ChunkBlocks) {
      const int bL = max(0, bH - kC2ChunkBlocks + 1);
      const int i0 = 6 * bL, itop = min(H.nh, 6 * (bH + 1 + NJ));
      __syncthreads();
#ifdef UBA_BAND_TIMING
      long long tb_ = clock64();
#endif
      {
        // factor rows [i0, itop) of the chunk and of the NJ blocks above it (batched loads); rows beyond the half are zero
        constexpr int U = 8;
        const int cnt = (6 * (bH + 1 + NJ) - i0) * bw1, have = (itop - i0) * bw1;
        const double* srcp = Lt + (size_t)i0 * bw1;
        for (int base = 0; base < cnt; base += NH * U) {
          double v[U];
#pragma unroll
          for (int u = 0; u < U; u++) { const int e = base + t + u * NH; v[u] = e < have ? srcp[e] : 0.0; }
#pragma unroll
          for (int u = 0; u < U; u++) { const int e = base + t + u * NH; if (e < cnt) Lb[e] = v[u]; }
        }
      }
      __syncthreads();
#ifdef UBA_BAND_TIMING
      { const long long n_ = clock64(); bseg[0] += n_ - tb_; tb_ = n_; }
#endif
      // M_b and P^j_b: one triangular solve L_bb^T m = v per (block, group, column a)
      const int nblk_c = bH - bL + 1;
      for (int e = t; e < nblk_c * NG * 6; e += NH) {
        const int bl = e / (NG * 6), g = (e / 6) % NG, a = e % 6;
        const double* Lrow0 = Lb + (size_t)(6 * bl) * bw1;       // row 6 (bL + bl) of the staged rows
        double v[6];
        if (g == 0) {
#pragma unroll
          for (int r = 0; r < 6; r++) v[r] = r == a ? 1.0 : 0.0;
        } else {
          const double* Lk = Lb + (size_t)(6 * (bl + g) + a) * bw1;  // row a of block b + g
#pragma unroll
          for (int r = 0; r < 6; r++) { const int d = 6 * g + a - r; const double l = Lk[d <= beta ? d : 0]; v[r] = d <= beta ? l : 0.0; }
        }
#pragma unroll
        for (int r = 5; r >= 0; r--) {
          double acc = v[r];
#pragma unroll
          for (int c2 = 0; c2 < 6; c2++) if (c2 > r) acc = fma(-Lrow0[c2 * bw1 + (c2 - r)], v[c2], acc);
          v[r] = acc * Lrow0[r * bw1];
        }
        double* dst = MP + ((size_t)bl * NG + g) * 36 + a;
#pragma unroll
        for (int r = 0; r < 6; r++) dst[r * 6] = v[r];
      }
      __syncthreads();
#ifdef UBA_BAND_TIMING
      { const long long n_ = clock64(); bseg[1] += n_ - tb_; tb_ = n_; }
#endif
      // sweep: step `b` finishes block b (panel warp) while the helper warp prepares u of block b-1; the step before the first
      // one (b == bH + 1) only prepares u of block bH
      for (int b = bH + 1; b >= bL; b--) {
#ifdef UBA_BAND_TIMING
        const long long ts_ = clock64();
#endif
        const int hw = t >> 5, hl = t & 31;
        if (hw >= 1 && hw <= NJ && hl < 6) {
          // helper warps 1 .. NJ, one group each (a single warp issuing all ~90 instructions of the five dot products takes
          // ~550 cycles: dependent-issue latency, not arithmetic): warp q+1 computes row hl of M z_hb (q = 0) or of
          // P^{q+1} x_{hb+q+1} (q >= 1) for the NEXT block hb = b - 1
          const int hb = b - 1, q = hw - 1;
          if (hb >= bL) {
            const int g = q == 0 ? 0 : q + 1;
            const double* row = MP + ((size_t)(hb - bL) * NG + g) * 36 + hl * 6;
            const double* src = y + 6 * (hb + g);
            double acc = 0.0;
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) acc = fma(row[c2], src[c2], acc);
            s_part[hb & 1][q][hl] = q == 0 ? acc : -acc;
          }
        } else if (panel && pl < 6 && b <= bH) {         // panel warp: x_b = u_b - P^1_b x_{b+1}, u_b = sum of the helpers' parts
          const double* row = MP + ((size_t)(b - bL) * NG + 1) * 36 + pl * 6;
          const double* xn = y + 6 * (b + 1);
          double acc = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < 6; c2++) acc = fma(row[c2], xn[c2], acc);
          double u0 = 0.0, u1 = 0.0;
#pragma unroll
          for (int q = 0; q < 6; q += 2) {
            if (q < NJ) u0 += s_part[b & 1][q][pl];
            if (q + 1 < NJ) u1 += s_part[b & 1][q + 1][pl];
          }
          y[6 * b + pl] = (u0 + u1) - acc;
        }
#ifdef UBA_BAND_TIMING
        brole += clock64() - ts_;
#endif
        __syncthreads();
      }
#ifdef UBA_BAND_TIMING
      { const long long n_ = clock64(); bseg[2] += n_ - tb_; tb_ = n_; }
#endif
    }
  }
  __syncthreads();
  PH()
#ifdef UBA_BAND_TIMING
  if (t == 0) for (int q = 0; q < nph; q++) { V.Zbuf[half * 8 + q] = (double)(tph[q] - tph[0]); V.Zbuf[64 + half * 8 + q] = (double)(tgl[q] % 1000000000ll); }
  if (half == 0 && (t & 31) == 0) V.Zbuf[16 + (t >> 5)] = (double)busy;
  if (half == 0 && t == 0) for (int q = 0; q < 3; q++) V.Zbuf[56 + q] = (double)bseg[q];
  if (half == 0 && (t == 0 || t == NWORKH - 32 || t == NWORKH)) V.Zbuf[59 + (t == 0 ? 0 : t == NWORKH ? 2 : 1)] = (double)brole;
  if (half == 0 && (t == 0 || t == 100 || t == 192)) for (int q = 0; q < 6; q++) V.Zbuf[32 + (t == 0 ? 0 : t == 100 ? 8 : 16) + q] = (double)seg[q];
#endif
#undef PH
  const bool failed = s_fail != 0;
  for (int i = t; i < H.nh; i += NH) {
    if (H.dir == 1 && i >= H.ne) continue;    // the separator is written once, by the top half
    rhs[H.dir == 0 ? i : n - 1 - i] = failed ? 0.0 : y[i];
  }
  if (t == 0 && half == 0 && failed) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
}
// ---------------------------------------------------------------------------------------------
// Block cyclic reduction of a banded reduced camera system on ONE CLUSTER of CTAs (k_chol_bcr).
//
// With camera spread bw (half-bandwidth beta = 6 bw + 5) the matrix is block TRIDIAGONAL in super-blocks of m = 6 bw rows
// (c4, c5: m = 24, N = 50 / 25 super-blocks).  Cholesky in the nested-dissection order of that path graph: level l
// eliminates every super-block j = s (2 q + 1), s = 2^l, whose two neighbours j - s and j + s stay — all of them
// INDEPENDENT, so a level is one round of parallel eliminations over the cluster and the sequential depth is
// log2(N) + 1 eliminations of a 24-row panel instead of N / 2 block steps of the two-sided sweep (k_chol_banded_c2).
// One elimination (one CTA, panel in shared memory; P = [D_j | A_ja | A_jb | r_j], m x (3 m + 1)):
//   LDL^T by rows with the pivot reciprocal computed one step ahead, rows scaled to the Cholesky factor at the end:
//       R = L^T,  W_a = L^-1 A_ja,  W_b = L^-1 A_jb,  z = L^-1 r_j;
//   Schur updates on the FP64 MMA path:  D_a -= W_a^T W_a,  D_b -= W_b^T W_b  (fp64 red.add: a surviving block has two
//   eliminated neighbours),  new coupling A_ba = -W_b^T W_a  (single writer),  r_a -= W_a^T z,  r_b -= W_b^T z.
// The panel STAYS in the CTA's shared memory (a CTA eliminates ~N / cluster size blocks), so the backward pass
//       x_j = R^-1 (z - W_a x_a - W_b x_b)
// reads nothing but the neighbours' solutions from global memory.  One cluster barrier per level in each direction.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int bcr_ldp(int mp) { return ((3 * mp + 1 + 15) / 16) * 16 + 8; }   // == 8 (mod 16): conflict-free MMA fragments
// doubles of scratch per super-block in global memory: D, C (MP x MP each), r, x (MP each)
__host__ __device__ constexpr size_t bcr_block_doubles(int mp) { return (size_t)2 * mp * mp + 2 * mp; }
// how many eliminations CTA `rank` of `ncta` performs (eliminations are dealt round-robin in level order); also the
// number of panels it keeps
__host__ __device__ inline int bcr_slots(int N, int ncta) { return (N + ncta - 1) / ncta; }

template <int MP>
__global__ void __launch_bounds__(256) k_chol_bcr(DevView V, int w, int beta) {
  extern __shared__ double sm[];
  const WinState* st = &V.ws[w];
  if (st->done) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
  constexpr int LDP = bcr_ldp(MP);
  constexpr int TT = MP / 8;
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  const int bw1 = beta + 1;
  const int m = beta - 5;                       // 6 * camera spread
  const int N = (n + m - 1) / m;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double* rhs = V.rhs + (size_t)6 * f0;
  double* Lt = V.A + V.w_red_off[w];
  const double* Ab = Lt + (size_t)2 * n * bw1;  // assembled band: Ab[i][c] = A[i][i - beta + c]
  double* G = Lt + (size_t)3 * n * bw1;         // scratch: per super-block D | C | r | x, then the failure flag
  const size_t BS = bcr_block_doubles(MP);
  auto Dg = [&](int j) { return G + (size_t)j * BS; };
  auto Cg = [&](int j) { return G + (size_t)j * BS + MP * MP; };
  auto rg = [&](int j) { return G + (size_t)j * BS + 2 * MP * MP; };
  auto xg = [&](int j) { return G + (size_t)j * BS + 2 * MP * MP + MP; };
  double* gfail = G + (size_t)N * BS;
  __shared__ double s_rd[MP + 1];               // pivot reciprocals of the panel being eliminated
  __shared__ double s_x[2 * MP];                // neighbours' solutions (backward)
  __shared__ int s_fail;
#ifdef UBA_BAND_TIMING
  long long tmark[24]; int nmark = 0;
#define BCR_MARK() { if (t == 0 && nmark < 24) { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); tmark[nmark++] = c_; } }
#else
#define BCR_MARK()
#endif
  BCR_MARK()
  if (t == 0) s_fail = 0;
  auto Aband = [&](int i, int k) -> double {     // A[i][k] of the assembled matrix, any (i, k) inside [0, n)
    if (k > i) { const int q = i; i = k; k = q; }
    return i - k <= beta ? Ab[(size_t)i * bw1 + (k - i + beta)] : 0.0;
  };

  // ---- phase 0: band -> block tridiagonal form in global scratch (identity on the padded tail).  Loads in batches of
  //      independent registers: a plain load / store loop is compiled as one memory round trip per element ----------------------
  for (int j = rank; j < N; j += ncta) {
    const int i0 = j * m;
    constexpr int U = (MP * MP + 255) / 256;
    double dv[U], cv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = t + u * 256;
      const int a = e / MP, b = e - a * MP;
      const int ia = i0 + a, ib = i0 + b;
      double d = 0.0, c = 0.0;
      if (e < MP * MP && a < m && b < m) {
        if (ia < n && ib < n) d = Aband(ia, ib); else d = a == b ? 1.0 : 0.0;
        const int ic = i0 + m + a;              // coupling with the next super-block: A[(j+1) m + a][j m + b]
        if (ic < n && ib < n) c = Aband(ic, ib);
      }
      dv[u] = d; cv[u] = c;
    }
#pragma unroll
    for (int u = 0; u < U; u++) { const int e = t + u * 256; if (e < MP * MP) { Dg(j)[e] = dv[u]; Cg(j)[e] = cv[u]; } }
    for (int a = t; a < MP; a += 256) { rg(j)[a] = (a < m && i0 + a < n) ? rhs[i0 + a] : 0.0; xg(j)[a] = 0.0; }
  }
  if (rank == 0 && t == 0) *gfail = 0.0;
  BCR_MARK()
  __threadfence();
  cluster.sync();
  BCR_MARK()

  // my panels, in the order I eliminate them
  constexpr int kMaxSlots = 16;
  int sj[kMaxSlots], ss[kMaxSlots];             // super-block and stride of each of my eliminations
  int nslot = 0;

  // ---- one elimination -----------------------------------------------------------------------------------------------------
  // The LDL^T runs on a REGISTER-resident panel, one thread per COLUMN of [D_j | A_ja | A_jb | r_j] (3 MP + 1 threads, three
  // warps; the other warps wait at the CTA barrier behind the loop): the thread keeps its column's MP rows.  Step k needs
  // its own row-k entry (a register), the pivot reciprocal and the row-k entries of the columns below the pivot — the
  // multipliers — which every column thread published in shared memory when its row k became final:
  //     v[i] -= row_k[i] * (v[k] / d_k)        for the rows i > k,
  // one load and one FMA per entry, all indices static.  One named barrier (three warps) per step; the owner of the next
  // pivot publishes its reciprocal together with its row entry.
  constexpr int NCOL = 3 * MP + 1;
  constexpr int NLD = ((NCOL + 31) / 32) * 32;  // threads in the LDL^T loop (whole warps)
  static_assert(NLD <= 256, "k_chol_bcr: panel does not fit 256 threads");
  __shared__ double s_row[2][NCOL + 3];         // published rows (double-buffered by step parity); NCOL >= 2 MP: reads past a row's end stay inside
  __shared__ double s_d[MP];                    // pivots
  const int pc = t;                             // my column (t < NCOL)
  auto eliminate = [&](double* P, int j, int s) {
    const int a = j - s, b = j + s;
    const bool has_a = a >= 0 && s > 0, has_b = b < N && s > 0;
    if (t < NLD) {
      double v[MP];
#pragma unroll
      for (int i = 0; i < MP; i++) {
        double x = 0.0;
        if (pc < MP) x = Dg(j)[i * MP + pc];
        else if (pc < 2 * MP) x = has_a ? Cg(a)[i * MP + (pc - MP)] : 0.0;               // A_ja = C_a
        else if (pc < 3 * MP) x = has_b ? Cg(j)[(pc - 2 * MP) * MP + i] : 0.0;           // A_jb = C_j^T
        else if (pc == 3 * MP) x = rg(j)[i];
        v[i] = x;
      }
      if (pc < NCOL) s_row[0][pc] = v[0];
      if (pc == 0) {
        const double d0 = v[0];
        if (!(d0 >= 2.2250738585072014e-308 && d0 <= 1.7976931348623157e308)) s_fail = 1;
        s_d[0] = d0; s_rd[0] = uba_rcp(d0);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NLD) : "memory");
      BCR_MARK()
      // A ROLLED loop over the pivots with a compact body: straight-line code that runs once per elimination is fetched
      // at ~6.5 cycles per instruction (measured: 450 cycles per step for 70 instructions, barriers and the reciprocal
      // chain not counting), a loop body that stays in the instruction cache is not.  The register column SHIFTS by one row
      // per step (the shift rides on the FMA's destination), so that the pivot row is always v[0] and every index is static;
      // a row's final value goes to the shared-memory panel when it leaves.
      const unsigned srow = (unsigned)__cvta_generic_to_shared(&s_row[0][0]);
#pragma unroll 1
      for (int kk = 0; kk < m; kk++) {
        const unsigned row = srow + (unsigned)(((kk & 1) * (NCOL + 3) + kk) * 8);   // &s_row[kk & 1][kk]
        const double f = v[0] * s_rd[kk];
        if (pc < NCOL) P[kk * LDP + pc] = v[0];                                     // row kk, unscaled
#pragma unroll
        for (int j = 1; j < MP; j++) v[j - 1] = fma(-lds_f64(row + 8 * j), f, v[j]);   // entries past the last row: unused
        v[MP - 1] = 0.0;
        if (kk + 1 < m) {
          if (pc < NCOL) s_row[(kk + 1) & 1][pc] = v[0];
          if (pc == kk + 1) {
            const double nv = v[0];
            if (!(nv >= 2.2250738585072014e-308 && nv <= 1.7976931348623157e308)) s_fail = 1;
            s_d[kk + 1] = nv; s_rd[kk + 1] = uba_rcp(nv);
          }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(NLD) : "memory");
      }
      BCR_MARK()
      // rows -> Cholesky factor in the shared-memory panel: R[i][c] = P[i][c] / sqrt(d_i); the strict lower triangle of the D
      // part receives R^T (row k = column k of R, contiguous: the backward substitution reads it without bank conflicts),
      // the diagonal slot 1 / R[i][i]; rows >= m stay zero
      if (t < MP) s_rd[t] = t < m ? uba_rsqrt(s_d[t]) : 0.0;
      asm volatile("bar.sync 2, %0;" ::"n"(NLD) : "memory");
      if (pc < NCOL) {
#pragma unroll 1
        for (int i = 0; i < MP; i++) {
          const double isd = s_rd[i];
          const double r = i < m ? P[i * LDP + pc] * isd : 0.0;
          if (pc >= MP) P[i * LDP + pc] = r;
          else if (pc > i) { P[i * LDP + pc] = pc < m ? r : 0.0; if (pc < m) P[pc * LDP + i] = r; }
          else if (pc == i) P[i * LDP + i] = isd;
        }
      }
    }
    __syncthreads();
    BCR_MARK()
    // Schur updates of the neighbours (FP64 MMA): tiles of W_a^T W_a, W_b^T W_b, W_b^T W_a over K = rows of the panel
    if (has_a || has_b) {
      const int frow = lane >> 2, fk = lane & 3;
      constexpr int NTILE = 3 * TT * TT, TPWB = (NTILE + 7) / 8;
      int cx[TPWB], cy[TPWB], wh[TPWB];
      double c0[TPWB], c1[TPWB];
#pragma unroll
      for (int q = 0; q < TPWB; q++) {
        const int tile = warp + 8 * q;
        const int which = tile / (TT * TT), ij = tile - which * TT * TT, I = ij / TT, J = ij - I * TT;
        const bool on = tile < NTILE && !((which == 0 && !has_a) || (which == 1 && !has_b) || (which == 2 && !(has_a && has_b)));
        wh[q] = on ? which : -1;
        cx[q] = (which == 0 ? MP : 2 * MP) + 8 * I + frow; cy[q] = (which == 1 ? 2 * MP : MP) + 8 * J + frow;
        if (!on) { cx[q] = MP + frow; cy[q] = MP + frow; }
        c0[q] = 0.0; c1[q] = 0.0;
      }
#pragma unroll 1
      for (int ks = 0; ks < MP / 4; ks++) {
        const double* Pk = P + (4 * ks + fk) * LDP;
#pragma unroll
        for (int q = 0; q < TPWB; q++) dmma884(c0[q], c1[q], Pk[cx[q]], Pk[cy[q]]);
      }
#pragma unroll
      for (int q = 0; q < TPWB; q++) {
        if (wh[q] < 0) continue;
        const int oi = cx[q] - (wh[q] == 0 ? MP : 2 * MP), oj = cy[q] - frow - (wh[q] == 1 ? 2 * MP : MP) + 2 * fk;
        if (wh[q] == 0) { atomicAdd(&Dg(a)[oi * MP + oj], -c0[q]); atomicAdd(&Dg(a)[oi * MP + oj + 1], -c1[q]); }
        else if (wh[q] == 1) { atomicAdd(&Dg(b)[oi * MP + oj], -c0[q]); atomicAdd(&Dg(b)[oi * MP + oj + 1], -c1[q]); }
        else { Cg(a)[oi * MP + oj] = -c0[q]; Cg(a)[oi * MP + oj + 1] = -c1[q]; }   // A_ba: rows of b, columns of a
      }
      // r_a -= W_a^T z, r_b -= W_b^T z: four partial sums per entry
      {
        const int e = t >> 2, part = t & 3;       // entry e of [r_a | r_b], quarter of the rows
        const int side = e / MP, i = e - side * MP;
        double acc = 0.0;
        if (e < 2 * MP) {
#pragma unroll
          for (int kk = 0; kk < MP / 4; kk++) { const int kr = part + 4 * kk; acc = fma(P[kr * LDP + (1 + side) * MP + i], P[kr * LDP + 3 * MP], acc); }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0 && e < 2 * MP && ((side == 0 && has_a) || (side == 1 && has_b))) atomicAdd(&rg(side == 0 ? a : b)[i], -acc);
      }
    }
    __syncthreads();
    BCR_MARK()
  };

  // ---- forward: levels of independent eliminations -------------------------------------------------------------------------
  int g = 0;                                    // running index of the elimination (level order)
  for (int s = 1; s < N; s <<= 1) {
    for (int j = s; j < N; j += 2 * s, g++) {
      if (g % ncta != rank) continue;
      if (nslot < kMaxSlots) { sj[nslot] = j; ss[nslot] = s; eliminate(sm + (size_t)nslot * MP * LDP, j, s); nslot++; }
    }
    BCR_MARK()
    __threadfence();
    cluster.sync();
    BCR_MARK()
  }
  // the root (super-block 0, no neighbours left) and its solution
  const bool root_mine = g % ncta == rank;
  auto solve_block = [&](const double* P, int j, int s) {
    // y = z - W_a x_a - W_b x_b (eight partial sums per row, shuffle-reduced), then R x = y on warp 0
    const int a = j - s, b = j + s;
    const bool has_a = a >= 0 && s > 0, has_b = b < N && s > 0;
    if (t < 2 * MP) { const int side = t / MP, i = t - side * MP; s_x[t] = (side == 0 ? has_a : has_b) ? xg(side == 0 ? a : b)[i] : 0.0; }
    __syncthreads();
    {
      const int i = t >> 3, q = t & 7;
      double acc = 0.0;
      if (i < m) {
        for (int c = q; c < 2 * MP; c += 8) acc = fma(P[i * LDP + MP + c], s_x[c], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (q == 0 && i < MP) s_rd[i] = i < m ? P[i * LDP + 3 * MP] - acc : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double y = lane < MP ? s_rd[lane] : 0.0;
#pragma unroll 1
      for (int k = m - 1; k >= 0; k--) {
        const double l = lane < k ? P[k * LDP + lane] : 0.0;                      // row k of the lower triangle = column k of R
        const double xk = __shfl_sync(0xffffffffu, y, k) * P[k * LDP + k];        // diagonal slot = 1 / R[k][k]
        y = lane == k ? xk : fma(-l, xk, y);
      }
      if (lane < MP) {
        xg(j)[lane] = y;
        const int i = j * m + lane;
        if (lane < m && i < n) rhs[i] = y;
      }
    }
    __syncthreads();
  };
  if (root_mine) {
    double* P = sm + (size_t)nslot * MP * LDP;
    eliminate(P, 0, 0);
    solve_block(P, 0, 0);
  }
  __threadfence();
  cluster.sync();
  // ---- backward: the levels in reverse -------------------------------------------------------------------------------------
  int top = 1;
  while (top < N) top <<= 1;
  int slot = nslot - 1;
  for (int s = top >> 1; s >= 1; s >>= 1) {
    while (slot >= 0 && ss[slot] == s) {
      solve_block(sm + (size_t)slot * MP * LDP, sj[slot], s);
      slot--;
    }
    __threadfence();
    cluster.sync();
  }
  BCR_MARK()
#ifdef UBA_BAND_TIMING
  if (t == 0 && rank < 2) { for (int i = 0; i < 24; i++) V.Zbuf[rank * 24 + i] = i < nmark ? (double)(tmark[i] - tmark[0]) : 0.0; }
#endif
  // ---- failure: a bad pivot anywhere poisons the solve: report zeros and count the failure ---------------------------------
  if (t == 0 && s_fail) atomicAdd(gfail, 1.0);
  __threadfence();
  cluster.sync();
  if (*((volatile double*)gfail) != 0.0) {
    for (int i = rank * 256 + t; i < n; i += ncta * 256) rhs[i] = 0.0;
    if (rank == 0 && t == 0) atomicAdd(&V.w_loc[(size_t)w * WC_COUNT + WC_FAIL], 1.0);
  }
}

#endif  // UBA_EMU

// blocked forward + backward substitution with the factor in global memory; one CTA
__global__ void __launch_bounds__(1024) k_trsv_large(DevView V, int w) {
  extern __shared__ double y[];  // [n]
  __shared__ double l[NB][NB + 1];
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  const double* A = V.A + V.w_red_off[w];
  double* rhs = V.rhs + (size_t)6 * f0;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) y[i] = rhs[i];
  __syncthreads();
  // forward: L z = b
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = min(NB, n - j0);
    for (int e = tid; e < nb * nb; e += nt) { const int i = e / nb, j = e % nb; l[i][j] = A[(size_t)(j0 + i) * n + j0 + j]; }
    __syncthreads();
    if (tid < 32) {
      for (int i = 0; i < nb; i++) {
        double s = (tid < i) ? l[i][tid] * y[j0 + tid] : 0.0;
        s = warp_sum(s);
        if (tid == 0) y[j0 + i] = (y[j0 + i] - s) / l[i][i];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = j0 + nb + tid; i < n; i += nt) {
      double s = 0.0;
      const double* row = A + (size_t)i * n + j0;
      for (int k = 0; k < nb; k++) s += row[k] * y[j0 + k];
      y[i] -= s;
    }
    __syncthreads();
  }
  // backward: L^T x = z
  for (int j0 = ((n - 1) / NB) * NB; j0 >= 0; j0 -= NB) {
    const int nb = min(NB, n - j0);
    for (int e = tid; e < nb * nb; e += nt) { const int i = e / nb, j = e % nb; l[i][j] = A[(size_t)(j0 + i) * n + j0 + j]; }
    __syncthreads();
    if (tid < 32) {
      for (int i = nb - 1; i >= 0; i--) {
        double s = (tid > i && tid < nb) ? l[tid][i] * y[j0 + tid] : 0.0;
        s = warp_sum(s);
        if (tid == 0) y[j0 + i] = (y[j0 + i] - s) / l[i][i];
        __syncwarp();
      }
    }
    __syncthreads();
    // y[i] -= sum_k L[j0+k][i] * y[j0+k] for i < j0
    for (int i = tid; i < j0; i += nt) {
      double s = 0.0;
      for (int k = 0; k < nb; k++) s += A[(size_t)(j0 + k) * n + i] * y[j0 + k];
      y[i] -= s;
    }
    __syncthreads();
  }
  const bool failed = V.w_loc[(size_t)w * WC_COUNT + WC_FAIL] > 0.0;
  for (int i = tid; i < n; i += nt) rhs[i] = failed ? 0.0 : y[i];
}

#else
constexpr int NB = 32;
#endif  // !UBA_EMU

// ---------------------------------------------------------------------------------------------
// solve epilogue: camera step, candidate cameras (+ their R, t, G), camera part of the model change
// grid: nW blocks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_solve_epilogue(DevView V) {
  const int w = blockIdx.x;
  WinState* st = &V.ws[w];
  if (st->done) return;
  const int cur = st->cur, nxt = cur ^ 1;
  const int c0 = V.w_cam_off[w], c1 = V.w_cam_off[w + 1];
  const int f0 = V.w_free_off[w];
  const double* y = V.rhs + (size_t)6 * f0;
  double mc = 0.0, step2 = 0.0, x2 = 0.0;
  // grid (window, camera slices): a 200-keyframe window is spread over several SMs instead of looping in one CTA
  for (int gc = c0 + blockIdx.y * blockDim.x + threadIdx.x; gc < c1; gc += gridDim.y * blockDim.x) {
    const int f = V.free_cam[gc];
    double cand[6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
      const double x = V.cams[cur][(size_t)gc * 6 + a];
      double yv = 0.0;
      if (f >= 0) {
        yv = y[f * 6 + a];
        mc += yv * V.vacc[(size_t)gc * 6 + a] + V.cam_lam[(size_t)gc * 6 + a] * yv * yv;
        x2 += x * x;
      }
      cand[a] = x - yv;
      if (f >= 0) step2 += (x - cand[a]) * (x - cand[a]);
      V.cam_y[(size_t)gc * 6 + a] = yv;
      V.cams[nxt][(size_t)gc * 6 + a] = cand[a];
    }
    double out[kCamStride];
    cam_derive(cand, out);
    double* dst = V.camR[nxt] + (size_t)gc * kCamStride;
#pragma unroll
    for (int i = 0; i < kCamStride; i++) dst[i] = out[i];
  }
  mc = warp_sum(mc); step2 = warp_sum(step2); x2 = warp_sum(x2);
  if (warp_leader()) {
    double* acc = V.w_loc + (size_t)w * WC_COUNT;
    if (mc != 0.0) atomicAdd(&acc[WC_MCCAM], mc);
    if (step2 != 0.0) atomicAdd(&acc[WC_STEP2], step2);
    if (x2 != 0.0) atomicAdd(&acc[WC_X2], x2);
  }
}

// ---------------------------------------------------------------------------------------------
// back-substitution + candidate point + candidate cost; G adjacent lanes per point (G = 1, 2, 4 or 8, chosen by the mean
// track length).  Lane g of a group takes observations g, g + G, ... of the point: G consecutive feature words per load
// instead of one word every track-length words, 1 / G of the loads and of the register state per thread, the partial
// E^T (F y_c) summed over the group by a butterfly (every lane ends with the same bits), the 3x3 solve repeated on every
// lane, each lane's share of the candidate cost added up by the window reduction.
// ---------------------------------------------------------------------------------------------
#ifndef UBA_BACKSUB_MINBLOCKS
#define UBA_BACKSUB_MINBLOCKS 3
#endif
template <int M, int G>
__global__ void __launch_bounds__(128, G == 1 ? UBA_BACKSUB_MINBLOCKS : 4) k_backsub(DevView V) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = tid / G, g = tid % G;
  const bool in_range = p < V.NP;
  int w = in_range ? V.pt_win[p] : V.pt_win[V.NP - 1];
  const WinState* st = &V.ws[w];
  const bool live = in_range && st->done == 0;
  double mc = 0.0, step2 = 0.0, x2 = 0.0, cnew = 0.0;
  // observations held in registers per lane (the rest of a longer track takes the one-at-a-time path)
  constexpr int KC = G == 1 ? 6 : (G == 2 ? 4 : 3);
  int o0 = 0, o1 = 0, cur = 0, cbase = 0, kk = 0;
  double X[3] = {0, 0, 0}, t3[3] = {0, 0, 0};
  int occ[KC];
  double fcc[KC][M];
  double2 recv[kPtRec / 2];
  if (live) {
    cur = st->cur;
    o0 = V.pt_obs_off[p]; o1 = V.pt_obs_off[p + 1];
    X[0] = V.pts[cur][(size_t)p * 3]; X[1] = V.pts[cur][(size_t)p * 3 + 1]; X[2] = V.pts[cur][(size_t)p * 3 + 2];
  }
  const bool active = live && o1 > o0;
  if (active) {
    cbase = V.w_cam_off[w];
    const double* camR = V.camR[cur];
    // the point's record is needed only after the observation loop: fetch it now, eight 16-byte loads in flight
    {
      const double2* rec2 = reinterpret_cast<const double2*>(V.pt_rec + (size_t)p * kPtRec);
#pragma unroll
      for (int i = 0; i < kPtRec / 2; i++) recv[i] = __ldcs(rec2 + i);
    }
    // The kernel is bound by the latency of its global loads (ncu: long scoreboard), so the camera words and features of
    // my first KC observations are fetched up front (KC x (M + 1) independent loads in flight) and kept in registers
    // for the candidate-cost pass below.
    const int mine = o1 - o0 > g ? (o1 - o0 - g + G - 1) / G : 0;
    kk = min(mine, KC);
#pragma unroll
    for (int q = 0; q < KC; q++) {
      if (q < kk) {
        const int o = o0 + g + q * G;
        occ[q] = V.obs_cam[o];
#pragma unroll
        for (int m = 0; m < M; m++) fcc[q][m] = V.feat[(size_t)m * V.NO + o];
      }
    }
#pragma unroll
    for (int q = 0; q < KC; q++) {
      if (q < kk) {
        const int gc = cbase + (occ[q] & 0x3fffffff);
        if (V.free_cam[gc] >= 0) {
          const double* ycp = V.cam_y + (size_t)gc * 6;
          const double yc[6] = {ycp[0], ycp[1], ycp[2], ycp[3], ycp[4], ycp[5]};
          // t3 += E^T (F y_c), matrix-free (uba_math.h: obs_apply)
          obs_apply<M>(camR + (size_t)gc * kCamStride, X, fcc[q], (occ[q] >> 30) & 1, V.calib, V.loss, yc, t3);
        }
      }
    }
    for (int o = o0 + g + KC * G; o < o1; o += G) {
      const int oc = V.obs_cam[o];
      const int gc = cbase + (oc & 0x3fffffff);
      if (V.free_cam[gc] < 0) continue;
      double f[M];
#pragma unroll
      for (int m = 0; m < M; m++) f[m] = V.feat[(size_t)m * V.NO + o];
      const double* ycp = V.cam_y + (size_t)gc * 6;
      const double yc[6] = {ycp[0], ycp[1], ycp[2], ycp[3], ycp[4], ycp[5]};
      obs_apply<M>(camR + (size_t)gc * kCamStride, X, f, (oc >> 30) & 1, V.calib, V.loss, yc, t3);
    }
  }
  if constexpr (G > 1) {                          // (groups are aligned inside a warp; idle groups exchange zeros)
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
#pragma unroll
      for (int c = 0; c < 3; c++) t3[c] += __shfl_xor_sync(0xffffffffu, t3[c], off);
    }
  }
  if (live) {
    double Xn[3] = {X[0], X[1], X[2]};
    if (active) {
      const int nxt = cur ^ 1;
      const double Li[6] = {recv[0].x, recv[0].y, recv[1].x, recv[1].y, recv[2].x, recv[2].y};
      const double h[3] = {recv[3].x, recv[3].y, recv[4].x}, gr[3] = {recv[4].y, recv[5].x, recv[5].y}, lam[3] = {recv[6].x, recv[6].y, recv[7].x};
      double u[3], yp[3];
      linv_mul(Li, t3, u);
      u[0] = h[0] - u[0]; u[1] = h[1] - u[1]; u[2] = h[2] - u[2];
      linvT_mul(Li, u, yp);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double v = X[c] - yp[c];
        if (V.cfg.use_bounds) v = clampd(v, V.calib.lo[c], V.calib.hi[c]);
        Xn[c] = v;
        if (g == 0) {
          mc += yp[c] * gr[c] + lam[c] * yp[c] * yp[c];
          step2 += (X[c] - v) * (X[c] - v);
          x2 += X[c] * X[c];
        }
      }
      // candidate cost at (candidate cameras, candidate point): my observations
      const double* camRn = V.camR[nxt];
#pragma unroll
      for (int q = 0; q < KC; q++) {
        if (q < kk) {
          const int gc = cbase + (occ[q] & 0x3fffffff);
          double rraw[M];
          const double s = obs_residual<M>(camRn + (size_t)gc * kCamStride, Xn, fcc[q], (occ[q] >> 30) & 1, V.calib, rraw);
          cnew += 0.5 * loss_rho(V.loss, s);
        }
      }
      for (int o = o0 + g + KC * G; o < o1; o += G) {
        const int oc = V.obs_cam[o];
        const int gc = cbase + (oc & 0x3fffffff);
        double f[M];
#pragma unroll
        for (int m = 0; m < M; m++) f[m] = V.feat[(size_t)m * V.NO + o];
        double rraw[M];
        const double s = obs_residual<M>(camRn + (size_t)gc * kCamStride, Xn, f, (oc >> 30) & 1, V.calib, rraw);
        cnew += 0.5 * loss_rho(V.loss, s);
      }
    }
    if (g == 0) { const int nxt = cur ^ 1; V.pts[nxt][(size_t)p * 3] = Xn[0]; V.pts[nxt][(size_t)p * 3 + 1] = Xn[1]; V.pts[nxt][(size_t)p * 3 + 2] = Xn[2]; }
  }
  win_add(V.w_post, WP_COUNT, w, WP_MCPT, mc, active);
  win_add(V.w_post, WP_COUNT, w, WP_STEP2, step2, active);
  win_add(V.w_post, WP_COUNT, w, WP_X2, x2, active);
  win_add(V.w_post, WP_COUNT, w, WP_COSTNEW, cnew, active);
}

// ---------------------------------------------------------------------------------------------
// LM controller: [CERES-UPSTREAM] TrustRegionMinimizer + LevenbergMarquardtStrategy rules
// ---------------------------------------------------------------------------------------------
__global__ void k_lm_update(DevView V) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= V.nW) return;
  WinState st = V.ws[w];
  if (st.done) return;
  const double* alin = V.w_lin + (size_t)w * WL_COUNT;
  const double* apost = V.w_post + (size_t)w * WP_COUNT;
  const double* aloc = V.w_loc + (size_t)w * WC_COUNT;
  const SolverCfg& C = V.cfg;
  const double cost = alin[WL_COST];
  const double gmax = V.w_max[w];
  IterRec* recs = V.recs + (size_t)w * V.rec_stride;
  const bool fixedK = C.fixed_iterations > 0;
  const int max_it = fixedK ? C.fixed_iterations : C.max_iterations;
  st.cost = cost; st.gmax = gmax;
  if (st.iter == 0) {
    st.initial_cost = cost;
    IterRec r0; r0.cost = cost; r0.candidate_cost = cost; r0.model_cost_change = 0; r0.relative_decrease = 0; r0.radius = st.radius;
    r0.step_norm = 0; r0.gradient_max_norm = gmax; r0.accepted = 1; r0.pad_ = 0;
    recs[0] = r0;
  } else if (st.iter < V.rec_stride && recs[st.iter].accepted == 1) {
    recs[st.iter].gradient_max_norm = gmax;  // gradient at the iterate accepted last iteration
  }
  int done = 0;
  if (!isfinite(cost)) done = 5;                                      // UBA_TERM_FAILURE
  else if (!fixedK && gmax <= C.gradient_tolerance) done = 2;         // CONVERGENCE_GRADIENT
  if (!done) {
    const int it = st.iter + 1;
    st.iter = it;
    IterRec r; r.cost = cost; r.candidate_cost = cost; r.model_cost_change = 0; r.relative_decrease = 0; r.radius = st.radius;
    r.step_norm = 0; r.gradient_max_norm = gmax; r.accepted = 0; r.pad_ = 0;
    const double mc = 0.5 * (apost[WP_MCPT] + aloc[WC_MCCAM]);
    const bool fail = alin[WL_FAIL] > 0.0 || aloc[WC_FAIL] > 0.0;
    r.model_cost_change = mc;
    if (fail || !(mc > 0.0)) {
      r.accepted = -1;
      st.n_invalid++;
      st.consecutive_invalid++;
      if (st.consecutive_invalid >= C.max_consecutive_invalid_steps && !fixedK) done = 5;
      else st.radius *= 0.5;
    } else {
      st.consecutive_invalid = 0;
      const double cand = apost[WP_COSTNEW];
      const double step_norm = sqrt(apost[WP_STEP2] + aloc[WC_STEP2]);
      const double x_norm = sqrt(apost[WP_X2] + aloc[WC_X2]);
      const double cost_change = cost - cand;
      const double rel = cost_change / mc;
      r.candidate_cost = cand; r.step_norm = step_norm; r.relative_decrease = rel;
      if (!fixedK && step_norm <= C.parameter_tolerance * (x_norm + C.parameter_tolerance)) done = 3;
      else if (!fixedK && fabs(cost_change) <= C.function_tolerance * cost) done = 1;
      else if (isfinite(cand) && rel > C.min_relative_decrease) {
        st.cur ^= 1;
        st.cost = cand;
        st.n_success++;
        r.accepted = 1; r.cost = cand;
        const double q = 2.0 * rel - 1.0;
        st.radius = st.radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
        st.radius = fmin(C.max_radius, st.radius);
        st.decrease_factor = 2.0;
      } else {
        st.n_unsuccess++;
        st.radius = st.radius / st.decrease_factor;
        st.decrease_factor *= 2.0;
        if (!fixedK && st.radius < C.min_radius) done = 6;
      }
    }
    if (it < V.rec_stride) recs[it] = r;
    if (!done && it >= max_it) done = 4;  // NO_CONVERGENCE: iteration cap
    // point-sharded handles: some rank ran into max_solver_time_s — every rank sees the same sum and stops here
    if (!done && !fixedK && apost[WP_STOP] > 0.0) done = 4;
  }
  st.scale_ready = 1;
  st.done = done;
  V.ws[w] = st;
  if (done) atomicSub(V.n_active, 1);
}


// ---------------------------------------------------------------------------------------------
// pose covariance blocks (extract_covariance, BundleAdjuster.h:478-528): with the points
// marginalised the covariance of free camera f is the (f, f) block of S^-1, S the UNDAMPED reduced
// camera matrix at the final iterate.  S = L L^T is already factorised (dense, lower, in V.A):
// solve L Y = E_f (6 right-hand sides) column by column and form Y^T Y.   grid: one CTA per free camera.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cov_blocks(DevView V, double* cov36 /*[NC][36]*/) {
#ifndef UBA_EMU
  extern __shared__ double sm[];
  const int gf = blockIdx.x;                 // index into free_list
  const int gc = V.free_list[gf];
  const int w = V.cam_win[gc];
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  const int c0 = 6 * (gf - f0);
  const double* L = V.A + V.w_red_off[w];
  double* Y = sm;                            // [n - c0][6]
  const int t = threadIdx.x, nt = blockDim.x;
  const int m = n - c0;
  for (int e = t; e < m * 6; e += nt) Y[e] = (e / 6 == e % 6) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 0; k < m; k++) {
    // row k is final once divided by the pivot; then eliminate it from the rows below
    const double inv = 1.0 / L[(size_t)(c0 + k) * n + c0 + k];
    if (t < 6) Y[k * 6 + t] *= inv;
    __syncthreads();
    for (int e = t; e < (m - k - 1) * 6; e += nt) {
      const int r = k + 1 + e / 6, c = e % 6;
      Y[r * 6 + c] = fma(-L[(size_t)(c0 + r) * n + c0 + k], Y[k * 6 + c], Y[r * 6 + c]);
    }
    __syncthreads();
  }
  // cov = Y^T Y
  __shared__ double red[36];
  if (t < 36) red[t] = 0.0;
  __syncthreads();
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; i++) acc[i] = 0.0;
  for (int r = t; r < m; r += nt) {
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int b = 0; b < 6; b++) acc[a * 6 + b] = fma(Y[r * 6 + a], Y[r * 6 + b], acc[a * 6 + b]);
  }
#pragma unroll
  for (int i = 0; i < 36; i++) {
    const double v = warp_sum(acc[i]);
    if ((t & 31) == 0) atomicAdd(&red[i], v);
  }
  __syncthreads();
  if (t < 36) cov36[(size_t)gc * 36 + t] = red[t];
#endif
}

// forces a window state for the covariance pass: running, undamped
__global__ void k_cov_state(DevView V, int enter) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= V.nW) return;
  WinState* st = &V.ws[w];
  if (enter) { st->pad_[0] = st->done; st->pad_[1] = 0; if (st->done != 5) st->done = 0; st->decrease_factor = st->radius; st->radius = -1.0; }
  else { st->done = st->pad_[0]; st->radius = st->decrease_factor; st->decrease_factor = 2.0; }
}

int launch_cov_state(const DevView& V, int enter, cudaStream_t st) {
  UBA_LAUNCH(k_cov_state, (V.nW + 127) / 128, 128, 0, st, V, enter);
  return 1;
}

int launch_cov_blocks(const DevView& V, int n_free_total, int max_n, double* cov36, cudaStream_t st) {
  if (n_free_total == 0) return 0;
  const size_t smem = (size_t)max_n * 6 * sizeof(double);
#ifndef UBA_EMU
  cudaFuncSetAttribute(k_cov_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  UBA_LAUNCH(k_cov_blocks, n_free_total, 256, smem, st, V, cov36);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------------------------
// Ingest: feature rows arrive as the caller stores them (AoS [NO][M], caller observation order; the reference's
// Observation<N>::data, BundleAdjuster.h:23-33).  Internal slot i holds caller observation src[i]; the SoA planes
// feat[m][i] are what the linearisers read coalesced.  The rows travel as float32 when every value is exactly a float.
template <typename TRaw>
__global__ void k_ingest_feats(const TRaw* __restrict__ raw, const int32_t* __restrict__ src, double* __restrict__ feat, int64_t NO, int M) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < NO; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = src[i];
    for (int m = 0; m < M; m++) feat[(size_t)m * NO + i] = (double)raw[(size_t)o * M + m];
  }
}

// ---- sliding window (uba_window_advance): the canonical observation rows live on the device, [NO][M] doubles in the
// caller's point order, tracks contiguous in the keyframe index, so that observation (point j, camera c) sits at row
// off[j] + (c - lo[j]).  Advancing the window is then three small kernels instead of a host re-ingest + upload.
// (1) surviving rows of the surviving points move to their new offsets
__global__ void k_win_shift(const double* __restrict__ old_rows, const int32_t* __restrict__ old_off, const int32_t* __restrict__ dropped,
                            const int32_t* __restrict__ id_map, const int32_t* __restrict__ new_off, int old_np, int M, double* __restrict__ new_rows) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= old_np) return;
  const int nj = id_map[j];
  if (nj < 0) return;
  const int o0 = old_off[j] + dropped[j], o1 = old_off[j + 1];
  double* dst = new_rows + (size_t)new_off[nj] * M;
  const double* src = old_rows + (size_t)o0 * M;
  for (int e = 0; e < (o1 - o0) * M; e++) dst[e] = src[e];
}
// (2) the new keyframes' observations drop into place
__global__ void k_win_append(const double* __restrict__ feats, const int32_t* __restrict__ pt_idx, const int32_t* __restrict__ cam_idx,
                             const int32_t* __restrict__ lo, const int32_t* __restrict__ off, int n, int M, double* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = pt_idx[i];
  double* dst = rows + ((size_t)off[j] + (cam_idx[i] - lo[j])) * M;
  for (int m = 0; m < M; m++) dst[m] = feats[(size_t)i * M + m];
}
// (3) canonical rows -> the lineariser's point-sorted SoA planes + camera words; thread per internal point slot
__global__ void k_win_pack(const double* __restrict__ rows, const int32_t* __restrict__ off_c, const int32_t* __restrict__ lo_c,
                           const unsigned char* __restrict__ cid_c, const int32_t* __restrict__ pt_order, const int32_t* __restrict__ off_i,
                           int NP, int64_t NO, int M, double* __restrict__ feat, int32_t* __restrict__ obs_cam) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= NP) return;
  const int j = pt_order[s];
  const int o0 = off_i[s], k = off_i[s + 1] - o0;
  const double* src = rows + (size_t)off_c[j] * M;
  const int bit = cid_c[j] ? (1 << 30) : 0;
  for (int q = 0; q < k; q++) {
    for (int m = 0; m < M; m++) feat[(size_t)m * NO + o0 + q] = src[(size_t)q * M + m];
    obs_cam[o0 + q] = (lo_c[j] + q) | bit;
  }
}
// refined points of the previous window -> their slots in the new internal order; new points from the upload
__global__ void k_win_points(const double* __restrict__ old_pts, const int32_t* __restrict__ src_slot, const double* __restrict__ fresh,
                             int NP, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= NP) return;
  const int src = src_slot[s];
  const double* p = src >= 0 ? old_pts + (size_t)src * 3 : fresh + (size_t)(-src - 1) * 3;
  out[(size_t)s * 3] = p[0]; out[(size_t)s * 3 + 1] = p[1]; out[(size_t)s * 3 + 2] = p[2];
}
// raw rows as uploaded by uba_set_problem (float32 or double, caller order) -> canonical double rows
template <typename TRaw>
__global__ void k_win_rows(const TRaw* __restrict__ raw, double* __restrict__ rows, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) rows[i] = (double)raw[i];
}

// point-sharded runs: per-window gradient max-norm through a SUM allreduce — scatter (gather = 0): this rank's max into its
// slot; gather (gather = 1, after the allreduce): max over the ranks' slots
__global__ void k_rank_max(double* w_max, double* w_rmax, double* w_post, const double* stop_req, int nW, int rank, int n_ranks, int gather) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nW) return;
  if (!gather) {
    w_rmax[(size_t)w * n_ranks + rank] = w_max[w];
    w_post[(size_t)w * WP_COUNT + WP_STOP] = *stop_req;   // summed with the rest: any rank over its wall-clock cap stops all
    return;
  }
  double m = 0.0;
  for (int r = 0; r < n_ranks; r++) m = fmax(m, w_rmax[(size_t)w * n_ranks + r]);
  w_max[w] = m;
}

__global__ void k_l2_flush(double* buf, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = (double)i;
}

// fp64 FMA throughput probe: 8 independent chains per thread.  out[0] keeps the compiler honest.
__global__ void __launch_bounds__(256) k_dfma_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_cam_prep(const DevView& V, int parity, cudaStream_t st) {
  if (V.NC == 0) return 0;
  UBA_LAUNCH(k_cam_prep, (V.NC + 127) / 128, 128, 0, st, V, parity);
  return 1;
}

int launch_init_state(const DevView& V, double initial_radius, cudaStream_t st) {
  UBA_LAUNCH(k_init_state, (V.nW + 127) / 128, 128, 0, st, V, initial_radius);
  return 1;
}

int launch_lin_generic(const DevView& V, const DebugOut& dbg, bool only_listed, cudaStream_t st) {
  const int count = only_listed ? V.n_gen : V.NP;
  if (count == 0) return 0;
  const int grid = (count + 127) / 128;
  const int listed = only_listed ? 1 : 0;
  if (V.M == 4) UBA_LAUNCH(k_lin_generic<4>, grid, 128, 0, st, V, dbg, listed);
  else UBA_LAUNCH(k_lin_generic<2>, grid, 128, 0, st, V, dbg, listed);
  return 1;
}


size_t lin_tile2_smem_bytes(int nt) {
#ifdef UBA_EMU
  (void)nt;
  return 0;
#else
  // camS | Zm (rows 8T <= 128, row stride t2_ldz(Pc)); the flush scratch ([points x slots][33] <= nt*33 doubles, or
  // the (8T) x (8T+1) K-group reduction tile, T <= 12 there) aliases Zm
  return sizeof(double) * ((size_t)kTileMaxLocal * kCamSm + t2_main_doubles(nt) + t2_stage_doubles(nt));
#endif
}

#ifndef UBA_EMU
template <int M, int NT, int T, int TG>
static int launch_t2_variant(const DevView& V, int first, int count, cudaStream_t st) {
  if (count == 0) return 0;
  const size_t smem = lin_tile2_smem_bytes(NT);
  static bool configured = false;
  if (!configured) { cudaFuncSetAttribute(k_lin_tile2<M, NT, T, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = true; }
  UBA_LAUNCH((k_lin_tile2<M, NT, T, TG>), count, NT, smem, st, V, first);
  return 1;
}

#endif

// variant of a part: the slot kernel's two width classes and k_lin_wide beyond them (default), or, with
// uba_config.linearizer = 2, k_lin_tile2's row-tile classes (8-row tiles over 6 * (free local cameras) rows)
int lin_part_variant(int n_local, int n_free_local, bool slot_kernel) {
  if (slot_kernel) return n_local <= 5 ? 0 : n_local <= kSlotMaxLocal ? 1 : 7;   // 7: k_lin_wide
  const int rows = 6 * n_free_local;
  return 2 + (rows <= 32 ? 0 : rows <= 48 ? 1 : rows <= 64 ? 2 : rows <= 96 ? 3 : 4);
}

#ifndef UBA_EMU
template <int M, int NLMAX>
static int launch_slot_variant(const DevView& V, int first, int count, cudaStream_t st) {
  if (count == 0) return 0;
  const size_t smem = slot_smem_bytes(NLMAX);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_lin_slot<M, NLMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_lin_slot<M, NLMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  if (V.pipe_on) UBA_LAUNCH((k_lin_slot<M, NLMAX, true>), count, 32 * (NLMAX + 1), smem, st, V, first);
  else UBA_LAUNCH((k_lin_slot<M, NLMAX, false>), count, 32 * (NLMAX + 1), smem, st, V, first);
  return 1;
}
#endif

size_t lin_wide_smem_bytes(int nt) {
#ifdef UBA_EMU
  (void)nt;
  return 0;
#else
  return sizeof(double) * (kTileMaxLocal * kCamStride + nt * 9 + nt * 3 + nt * 6 + nt / 2 + nt * kFlushStride);
#endif
}

// variant_off[kLinVariants + 1]: parts are sorted by variant; variant v covers [variant_off[v], variant_off[v+1]).
// When parts of several variants exist (ragged windows such as c2: short tracks on k_lin_slot, long ones on k_lin_wide),
// their kernels run SIDE BY SIDE: every variant after the first is launched on a side stream forked from `st` and joined
// back (inside a stream capture this becomes a fork / join in the graph).  They only meet in the fp64 atomics.
int launch_lin_tiled(const DevView& V, const int* variant_off, cudaStream_t st) {
  if (V.n_parts == 0) return 0;
#ifdef UBA_EMU
  (void)st; (void)variant_off;
  return 0;
#else
  constexpr int kSide = 3;
  static thread_local cudaStream_t side[kSide] = {};
  static thread_local cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {};
  if (!ev_fork) {
    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < kSide; i++) { cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming); }
  }
  int n = 0, used = 0;
  bool forked = false;
  // the stream for the next non-empty variant: the caller's for the first, a side stream for the following ones
  auto next_stream = [&](int count) -> cudaStream_t {
    if (count <= 0 || n == 0) return st;
    if (!forked) { cudaEventRecord(ev_fork, st); forked = true; }
    cudaStream_t s2 = side[used % kSide];
    if (used < kSide) cudaStreamWaitEvent(s2, ev_fork, 0);
    used++;
    return s2;
  };
  const int* o = variant_off;
  { const int c0 = o[1] - o[0]; cudaStream_t s2 = next_stream(c0); n += V.M == 4 ? launch_slot_variant<4, 5>(V, o[0], c0, s2) : launch_slot_variant<2, 5>(V, o[0], c0, s2); }
  { const int c1 = o[2] - o[1]; cudaStream_t s2 = next_stream(c1); n += V.M == 4 ? launch_slot_variant<4, 10>(V, o[1], c1, s2) : launch_slot_variant<2, 10>(V, o[1], c1, s2); }
  o += 2;
#define T2_ONE(MM, NN, TT, GG, k) { const int cc = o[k + 1] - o[k]; cudaStream_t s2 = next_stream(cc); n += launch_t2_variant<MM, NN, TT, GG>(V, o[k], cc, s2); }
#define T2_ALL(MM, NN) T2_ONE(MM, NN, 4, 1, 0) T2_ONE(MM, NN, 6, 2, 1) T2_ONE(MM, NN, 8, 2, 2) T2_ONE(MM, NN, 12, 4, 3)
  if (V.tile_threads == 128) {
    if (V.M == 4) { T2_ALL(4, 128) } else { T2_ALL(2, 128) }
  } else {
    if (V.M == 4) { T2_ALL(4, 256) T2_ONE(4, 256, 16, 8, 4) }
    else { T2_ALL(2, 256) T2_ONE(2, 256, 16, 8, 4) }
  }
#undef T2_ALL
#undef T2_ONE
  o += 5;
  if (o[1] > o[0]) {     // wide parts
    const size_t smem = lin_wide_smem_bytes(256);
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(k_lin_wide<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_lin_wide<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = true;
    }
    cudaStream_t s2 = next_stream(o[1] - o[0]);
    if (V.M == 4) UBA_LAUNCH((k_lin_wide<4, 256>), o[1] - o[0], 256, smem, s2, V, o[0]);
    else UBA_LAUNCH((k_lin_wide<2, 256>), o[1] - o[0], 256, smem, s2, V, o[0]);
    n++;
  }
  // join the side streams back
  for (int i = 0; i < used && i < kSide; i++) { cudaEventRecord(ev_join[i], side[i]); cudaStreamWaitEvent(st, ev_join[i], 0); }
  return n;
#endif
}

int launch_assemble(const DevView& V, int max_n, cudaStream_t st) {
  if (max_n == 0) return 0;
  const int64_t total = (int64_t)max_n * max_n + max_n;
  int bx = (int)((total + 255) / 256);
  if (bx > 1184) bx = 1184;
  dim3 grid(bx, V.nW);
  UBA_LAUNCH(k_assemble, grid, 256, 0, st, V);
  return 1;
}

int solve_small_limit() { return 160; }

int launch_solve(const DevView& V, const int* h_win_n, const int* h_win_beta, int max_small_n, cudaStream_t st, bool keep_factor, int parts) {
  int launches = 0;
  if (parts & 1) {
  int small_max = 0, n_large = 0;
  for (int w = 0; w < V.nW; w++) {
    if (h_win_beta[w] > 0) n_large++;
    else if (h_win_n[w] <= max_small_n) small_max = h_win_n[w] > small_max ? h_win_n[w] : small_max;
    else n_large++;
  }
#ifdef UBA_EMU
  for (int w = 0; w < V.nW; w++) uba_emu::dense_solve(V, w);
  (void)small_max; (void)n_large;
#else
  if (small_max > 0) {
    const size_t smem = ((size_t)(small_max + 1) * (small_max + 1) + small_max) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
      cudaFuncSetAttribute(k_chol_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = smem;
    }
    static const bool use_small_la = [] { const char* e = getenv("UBA_SMALL_LA"); return !(e && e[0] == '0'); }();
    if (use_small_la) {
      const size_t smem_la = smem + ((size_t)(small_max + 8) * 9 + 2) * sizeof(double);
      static size_t configured_la = 0;
      if (smem_la > configured_la) {
        cudaFuncSetAttribute(k_chol_small_la, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_la);
        configured_la = smem_la;
      }
      UBA_LAUNCH(k_chol_small_la, V.nW, 256, smem_la, st, V, max_small_n, keep_factor ? 1 : 0);
    } else {
      UBA_LAUNCH(k_chol_small, V.nW, 256, smem, st, V, max_small_n, keep_factor ? 1 : 0);
    }
    launches++;
  }
  if (n_large) {
    for (int w = 0; w < V.nW; w++) {
      const int n = h_win_n[w];
      if (h_win_beta[w] > 0) {
        const int beta = h_win_beta[w];
#ifndef UBA_EMU
        // UBA_BAND_BCR=<cluster size>: block cyclic reduction on one cluster of CTAs.  Correct (tested), log-depth, but its
        // per-elimination constant is not there yet: c4 0.149 ms on 16 CTAs against 0.116 ms for the two-sided sweep below
        // (scripts/bcr_timing.py gives the phase split), so it is off by default.
        static const int bcr_ctas = [] { const char* e = getenv("UBA_BAND_BCR"); return e ? atoi(e) : 0; }();
        if (bcr_ctas > 0 && beta >= 11 && beta <= 35 && (beta - 5) % 6 == 0 && n >= 12 * (beta + 1)) {
          const int m = beta - 5, mp = ((m + 7) / 8) * 8, N = (n + m - 1) / m;
          int ncta = bcr_ctas > 16 ? 16 : bcr_ctas;
          if (ncta > N) ncta = N;
          const size_t smem = (size_t)bcr_slots(N, ncta) * mp * bcr_ldp(mp) * sizeof(double);
          const size_t scratch = (size_t)3 * n * (beta + 1) + (size_t)N * bcr_block_doubles(mp) + 1;
          if (mp <= 24 && smem <= 200 * 1024 && bcr_slots(N, ncta) <= 16 && scratch <= (size_t)n * n) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(ncta); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem; lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
#define UBA_BCR_LAUNCH(MPV) { cudaFuncSetAttribute(k_chol_bcr<MPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (ncta > 8) cudaFuncSetAttribute(k_chol_bcr<MPV>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); \
            cudaLaunchKernelEx(&lc, k_chol_bcr<MPV>, V, w, beta); }
            if (mp == 8) UBA_BCR_LAUNCH(8) else if (mp == 16) UBA_BCR_LAUNCH(16) else UBA_BCR_LAUNCH(24)
#undef UBA_BCR_LAUNCH
            launches++;
            continue;
          }
        }
        // default for long bands: the two halves of the band on a cluster of two CTAs
        static const bool use_c2 = [] { const char* e = getenv("UBA_BAND_C2"); return !(e && e[0] == '0'); }();
        if (use_c2 && beta >= 11 && beta <= 35 && n >= 12 * (beta + 1)) {
          const int sw = ((beta + 1 + 5) / 6) * 6, mm = (((n - sw) / 2) / 6) * 6;
          const size_t smem = ((size_t)kBandRing * (beta + 1) + (n - mm) + beta + 8 + (size_t)40 * 9 + 8 + c2_backward_doubles(beta) + 8) * sizeof(double);
          const int per = (beta * (beta + 1) / 2 + 223) / 224;
#define UBA_C2_LAUNCH(PP, PI) { cudaFuncSetAttribute(k_chol_banded_c2<PP, PI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); UBA_LAUNCH((k_chol_banded_c2<PP, PI>), 2, 256, smem, st, V, w, beta); }
          if (V.pipe_on) { if (per <= 2) UBA_C2_LAUNCH(2, true) else UBA_C2_LAUNCH(3, true) }
          else { if (per <= 2) UBA_C2_LAUNCH(2, false) else UBA_C2_LAUNCH(3, false) }
#undef UBA_C2_LAUNCH
          launches++;
          continue;
        }
#endif
        if (beta >= 11) {
          const size_t smem = ((size_t)kBandRing * (beta + 1) + n + beta + 8 + (size_t)beta * 6 + 8) * sizeof(double);
          const int per = (beta * (beta + 1) / 2 + 223) / 224;
#define UBA_LA_LAUNCH(PP) { cudaFuncSetAttribute(k_chol_banded_la<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); UBA_LAUNCH(k_chol_banded_la<PP>, 1, 256, smem, st, V, w, beta); }
          if (per <= 1) UBA_LA_LAUNCH(1) else if (per <= 2) UBA_LA_LAUNCH(2) else if (per <= 4) UBA_LA_LAUNCH(4) else UBA_LA_LAUNCH(9)
#undef UBA_LA_LAUNCH
          launches++;
          continue;
        }
        continue;   // unreachable: prepare() only marks windows with beta >= 11 as banded
      }
      if (n <= max_small_n) continue;
      for (int j0 = 0; j0 < n; j0 += NB) {
        const int nb = n - j0 < NB ? n - j0 : NB;
        UBA_LAUNCH(k_chol_diag, 1, 256, 0, st, V, w, j0);
        launches++;
        const int rem = n - j0 - nb;
        if (rem > 0) {
          const int tiles = (rem + NB - 1) / NB;
          UBA_LAUNCH(k_chol_trsm, tiles, 256, 0, st, V, w, j0);
          UBA_LAUNCH(k_chol_update, dim3(tiles, tiles), 256, 0, st, V, w, j0);
          launches += 2;
        }
      }
      const size_t smem = (size_t)n * sizeof(double);
      if (smem > 48 * 1024) cudaFuncSetAttribute(k_trsv_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      UBA_LAUNCH(k_trsv_large, 1, 1024, smem, st, V, w);
      launches++;
    }
  }
#endif
  }
  if (parts & 2) {
    const int avg = (V.NC + V.nW - 1) / V.nW;          // any split is correct (grid-stride loop); this one suits uniform windows
    const int bs = avg <= 32 ? 32 : 64;
    int slices = (avg + bs - 1) / bs; if (slices > 16) slices = 16;
    UBA_LAUNCH(k_solve_epilogue, dim3(V.nW, slices), bs, 0, st, V);
    launches++;
  }
  return launches;
}

int launch_backsub(const DevView& V, cudaStream_t st) {
  if (V.NP == 0) return 0;
#ifdef UBA_EMU
  const int G = 1;                               // the emulation runs the threads one after the other: no lane groups
#else
  // lanes per point (UBA_BACKSUB_GROUP overrides: 1, 2, 4, 8).  Measured on B200: with enough points to fill the GPU one
  // thread per point wins (c4: 0.050 ms against 0.074 / 0.126 with 2 / 4 lanes — the per-point work every lane repeats
  // outweighs the shorter observation loops); small windows gain from the extra parallelism (c2: 0.032 -> 0.021 ms).
  static const int forced = [] { const char* e = getenv("UBA_BACKSUB_GROUP"); return e ? atoi(e) : 0; }();
  int G = forced ? forced : (V.NP < 40000 ? 2 : 1);
  if (G != 1 && G != 2 && G != 4 && G != 8) G = 1;
#endif
  const int grid = (int)(((int64_t)V.NP * G + 127) / 128);
#define UBA_BS(MM, GG) UBA_LAUNCH((k_backsub<MM, GG>), grid, 128, 0, st, V)
  if (V.M == 4) { if (G == 1) UBA_BS(4, 1);
#ifndef UBA_EMU
    else if (G == 2) UBA_BS(4, 2); else if (G == 4) UBA_BS(4, 4); else UBA_BS(4, 8);
#endif
  } else { if (G == 1) UBA_BS(2, 1);
#ifndef UBA_EMU
    else if (G == 2) UBA_BS(2, 2); else if (G == 4) UBA_BS(2, 4); else UBA_BS(2, 8);
#endif
  }
#undef UBA_BS
  return 1;
}

int launch_lm_update(const DevView& V, cudaStream_t st) {
  UBA_LAUNCH(k_lm_update, (V.nW + 127) / 128, 128, 0, st, V);
  return 1;
}

int launch_ingest_feats(const void* raw, const int32_t* src, double* feat, int64_t NO, int M, int raw_is_f32, cudaStream_t st) {
  if (NO == 0) return 0;
  const int64_t blocks = (NO + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  if (raw_is_f32) UBA_LAUNCH(k_ingest_feats<float>, grid, 256, 0, st, (const float*)raw, src, feat, NO, M);
  else UBA_LAUNCH(k_ingest_feats<double>, grid, 256, 0, st, (const double*)raw, src, feat, NO, M);
  return 1;
}

int launch_win_shift(const double* old_rows, const int32_t* old_off, const int32_t* dropped, const int32_t* id_map, const int32_t* new_off,
                     int old_np, int M, double* new_rows, cudaStream_t st) {
  if (old_np == 0) return 0;
  UBA_LAUNCH(k_win_shift, (old_np + 127) / 128, 128, 0, st, old_rows, old_off, dropped, id_map, new_off, old_np, M, new_rows);
  return 1;
}
int launch_win_append(const double* feats, const int32_t* pt_idx, const int32_t* cam_idx, const int32_t* lo, const int32_t* off, int n, int M,
                      double* rows, cudaStream_t st) {
  if (n == 0) return 0;
  UBA_LAUNCH(k_win_append, (n + 127) / 128, 128, 0, st, feats, pt_idx, cam_idx, lo, off, n, M, rows);
  return 1;
}
int launch_win_pack(const double* rows, const int32_t* off_c, const int32_t* lo_c, const unsigned char* cid_c, const int32_t* pt_order,
                    const int32_t* off_i, int NP, int64_t NO, int M, double* feat, int32_t* obs_cam, cudaStream_t st) {
  if (NP == 0) return 0;
  UBA_LAUNCH(k_win_pack, (NP + 127) / 128, 128, 0, st, rows, off_c, lo_c, cid_c, pt_order, off_i, NP, NO, M, feat, obs_cam);
  return 1;
}
int launch_win_points(const double* old_pts, const int32_t* src_slot, const double* fresh, int NP, double* out, cudaStream_t st) {
  if (NP == 0) return 0;
  UBA_LAUNCH(k_win_points, (NP + 127) / 128, 128, 0, st, old_pts, src_slot, fresh, NP, out);
  return 1;
}
int launch_win_rows(const void* raw, int raw_is_f32, double* rows, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  const int64_t blocks = (n + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  if (raw_is_f32) UBA_LAUNCH(k_win_rows<float>, grid, 256, 0, st, (const float*)raw, rows, n);
  else UBA_LAUNCH(k_win_rows<double>, grid, 256, 0, st, (const double*)raw, rows, n);
  return 1;
}

int launch_rank_max(double* w_max, double* w_rmax, double* w_post, const double* stop_req, int nW, int rank, int n_ranks, int gather, cudaStream_t st) {
  UBA_LAUNCH(k_rank_max, (nW + 127) / 128, 128, 0, st, w_max, w_rmax, w_post, stop_req, nW, rank, n_ranks, gather);
  return 1;
}

int launch_l2_flush(double* buf, size_t n, cudaStream_t st) {
  UBA_LAUNCH(k_l2_flush, 148 * 4, 256, 0, st, buf, n);
  return 1;
}

int launch_dfma_probe(double* out, int iters, cudaStream_t st) {
  UBA_LAUNCH(k_dfma_probe, 148 * 8, 256, 0, st, out, iters);
  return 1;
}

}  // namespace uba
