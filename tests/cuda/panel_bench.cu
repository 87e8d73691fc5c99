// tests/cuda/panel_bench.cu — where do the cycles of one band-Cholesky panel step go?  (not part of libuba)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_bench panel_bench.cu && ./panel_bench
// One warp repeats the panel warp's work of k_chol_banded_la on shared-memory data and times each piece with clock64.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ bool chol6(double (&L)[6][6], double (&inv)[6]) {
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double d = L[c][c];
    if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
    const double iv = rsqrt(d);
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

// rsqrt without the library's special-case handling: MUFU.RSQ64H seed + 2 Newton steps (inputs are positive, normal)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ bool chol6_fast(double (&L)[6][6], double (&inv)[6]) {
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double d = L[c][c];
    if (!(d > 1e-300) || !(d < 1e300)) { ok = false; d = 1.0; }
    const double iv = fast_rsqrt(d);
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

// one third-order step from the MUFU seed (the formula of the library's fast path), no special cases, and the
// validity test kept OFF the dependent chain (speculative: a bad pivot poisons the block with NaN and raises the flag)
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(y * y), x, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y * e, y);
}
__device__ __forceinline__ bool chol6_spec(double (&L)[6][6], double (&inv)[6]) {
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const double d = L[c][c];
    ok = ok && (d >= 2.2250738585072014e-308) && (d <= 1.7976931348623157e308);
    const double iv = rsqrt_pos(d);
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

template <int MODE>
__global__ void panel(double* out, long long* cyc, int iters) {
  __shared__ double ring[12 * 30], s_Lkk[2][36], s_invk[2][6], s_xp[36], s_corner[21];
  __shared__ int s_fail;
  const int pl = threadIdx.x;
  for (int i = pl; i < 12 * 30; i += 32) ring[i] = 0.01 * ((i * 7) % 13);
  for (int i = pl; i < 36; i += 32) { s_Lkk[0][i] = (i % 7 == 0) ? 2.0 : 0.1; s_Lkk[1][i] = s_Lkk[0][i]; }
  if (pl < 6) { s_invk[0][pl] = 0.5; s_invk[1][pl] = 0.5; }
  if (pl == 0) s_fail = 0;
  __syncwarp();
  int cr = 0, ce = pl;
  while (ce > cr) { ce -= cr + 1; cr++; }
  long long t[4] = {0, 0, 0, 0};
  const int bw1 = 30, beta = 29;
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    const double* Lk = s_Lkk[par];
    const double* ivk = s_invk[par];
    long long c0 = clock64();
    if (pl < 6) {
      const double* row = ring + pl * bw1 + (beta - 6 - pl);
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
    long long c1 = clock64();
    if (pl < 21) {
      double v = 40.0 * (cr == ce) + ring[6 * bw1 + cr * bw1 + beta - cr + ce];
#pragma unroll
      for (int m = 0; m < 6; m++) v = fma(-s_xp[cr * 6 + m], s_xp[ce * 6 + m], v);
      s_corner[pl] = v;
    }
    __syncwarp();
    long long c2 = clock64();
    if (MODE == 0 || MODE == 1 || MODE == 3) {
      if (pl == 0) {
        double L[6][6], iv[6];
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) L[r][c] = c <= r ? s_corner[r * (r + 1) / 2 + c] : 0.0;
        const bool ok = MODE == 0 ? chol6(L, iv) : MODE == 1 ? chol6_fast(L, iv) : chol6_spec(L, iv);
        if (!ok) s_fail = 1;
#pragma unroll
        for (int r = 0; r < 6; r++) {
          s_invk[par ^ 1][r] = iv[r];
#pragma unroll
          for (int c = 0; c < 6; c++) s_Lkk[par ^ 1][r * 6 + c] = L[r][c];
        }
      }
    } else {
      // MODE 2: lane-parallel Cholesky: lane pl < 21 owns entry (cr, ce); shuffles broadcast the pivot column
      double v = pl < 21 ? s_corner[pl] : 1.0;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double d = __shfl_sync(0xffffffffu, v, c * (c + 1) / 2 + c);
        const double iv = fast_rsqrt(d);
        if (ce == c) v = (cr == c) ? d * iv : v * iv;
        const double a = __shfl_sync(0xffffffffu, v, cr * (cr + 1) / 2 + c);   // L[cr][c]
        const double b = __shfl_sync(0xffffffffu, v, ce * (ce + 1) / 2 + c);   // L[ce][c]
        if (ce > c) v = fma(-a, b, v);
        if (pl == c) s_invk[par ^ 1][c] = iv;
      }
      if (pl < 21) s_Lkk[par ^ 1][cr * 6 + ce] = v;
    }
    __syncwarp();
    long long c3 = clock64();
    t[0] += c1 - c0; t[1] += c2 - c1; t[2] += c3 - c2;
  }
  if (pl == 0) { cyc[0] = t[0]; cyc[1] = t[1]; cyc[2] = t[2]; out[0] = s_Lkk[0][7] + s_invk[1][3] + s_fail; }
}

// ---- the worker side of one block step: 224 threads, beta = 29 ------------------------------------------------
// PARTS bit 0: triangular solves (30 threads), bit 1: trailing update of the band window, bit 2: rhs/y update
template <int PARTS, int XS>
__global__ void workers(double* out, long long* cyc, int iters) {
  constexpr int NWORK = 224, beta = 29, bw1 = 30, kRing = 126, PER = 2;
  __shared__ double ring[kRing * bw1], y[256], Xbuf[beta * 8 + 8], s_Lkk[36], s_invk[6], s_z[6];
  const int t = threadIdx.x;
  const int ring_size = kRing * bw1;
  for (int i = t; i < ring_size; i += NWORK) ring[i] = 1e-3 * ((i * 7) % 13);
  for (int i = t; i < 256; i += NWORK) y[i] = 0.5;
  for (int i = t; i < 36; i += NWORK) s_Lkk[i] = (i % 7 == 0) ? 2.0 : 0.1;
  if (t < 6) { s_invk[t] = 0.5; s_z[t] = 0.25; }
  for (int i = t; i < beta * 8 + 8; i += NWORK) Xbuf[i] = 1e-3 * i;
  const int npairs = beta * (beta + 1) / 2;
  int pti[PER], ptk[PER];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int e = t + q * NWORK;
    int ti = -1, tk = 0;
    if (e < npairs) {
      int d0 = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((d0 + 1) * (d0 + 2) / 2 <= e) d0++;
      while (d0 * (d0 + 1) / 2 > e) d0--;
      ti = d0; tk = e - d0 * (d0 + 1) / 2;
      if (ti < 6) ti = -1;
    }
    pti[q] = ti; ptk[q] = tk;
  }
  __syncthreads();
  int o0 = 0;
  const long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double* Lk = s_Lkk; const double* ivk = s_invk;
    if ((PARTS & 8) == 0) {
    if ((PARTS & 1) && t <= beta) {
        const bool is_rhs = t == beta;
        int orow = o0 + (6 + t) * bw1; if (orow >= ring_size) orow -= ring_size;
        const double* row = ring + orow;
        const int base = beta - 6 - t;
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = is_rhs ? y[(it & 15) + c] : ((base + c >= 0) ? row[base + c] : 0.0);
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
        if (is_rhs) {
#pragma unroll
          for (int c = 0; c < 6; c++) { s_z[c] = x[c]; y[(it & 15) + c] = x[c]; }
        } else {
#pragma unroll
          for (int c = 0; c < 6; c++) Xbuf[t * XS + c] = x[c];
        }
      }
    } else if (PARTS & 1) {
      // uniform row solves in warp 0 (clamped loads + select instead of divergent loads), rhs in another warp
      if (t < beta) {
        int orow = o0 + (6 + t) * bw1; if (orow >= ring_size) orow -= ring_size;
        const double* row = ring + orow;
        const int base = beta - 6 - t;
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
        for (int c = 0; c < 6; c++) Xbuf[t * XS + c] = x[c];
      } else if (t == 192) {
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; c++) x[c] = y[(it & 15) + c];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          x[c] *= ivk[c];
#pragma unroll
          for (int m = 0; m < 6; m++) if (m > c) x[m] = fma(-x[c], Lk[m * 6 + c], x[m]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) { s_z[c] = x[c]; y[(it & 15) + c] = x[c]; }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NWORK));
    if (PARTS & 2) {
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int ti = pti[q], tk = ptk[q];
        if (ti >= 0) {
          int oi = o0 + (6 + ti) * bw1; if (oi >= ring_size) oi -= ring_size;
          const double* xi = Xbuf + ti * XS;
          const double* xk = Xbuf + tk * XS;
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < 6; c++) acc = fma(xi[c], xk[c], acc);
          ring[oi + (beta - ti + tk)] -= acc * 1e-9;
        }
      }
    }
    if ((PARTS & 4) && t >= 64 && t < 64 + beta) {
      const int tt = t - 64, i = (it & 15) + 6 + tt;
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 6; c++) acc = fma(Xbuf[tt * XS + c], s_z[c], acc);
      y[i] -= acc * 1e-9;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(NWORK));
    o0 += 6 * bw1; if (o0 >= ring_size) o0 -= ring_size;
  }
  const long long c1 = clock64();
  if (t == 0) { cyc[0] = c1 - c0; out[0] = ring[5] + y[3] + Xbuf[9]; }
}

// ---- one block of the backward substitution (k_chol_banded_c2), 256 threads, beta = 29 ---------------------------------
// PARTS bit 0: diagonal solve by lane 0 of warp 7, bit 1: next block's rows by 6 lanes of warp 7, bit 2: workers one block behind
template <int PARTS>
__global__ void backward(double* out, long long* cyc, int iters) {
  constexpr int beta = 29, bw1 = 30, NWORK = 224;
  __shared__ double ring[126 * bw1], y[256], s_xb[2][6];
  const int t = threadIdx.x;
  const bool panel = t >= NWORK; const int pl = t - NWORK;
  for (int i = t; i < 126 * bw1; i += 256) ring[i] = (i % bw1 == 0) ? 0.5 : 1e-3 * ((i * 7) % 13);
  for (int i = t; i < 256; i += 256) y[i] = 0.5 + 1e-3 * i;
  if (t < 12) s_xb[t / 6][t % 6] = 0.1;
  __syncthreads();
  const long long c00 = clock64();
  for (int rep = 0; rep < iters; rep++) {
    int it = 1;
    for (int c0 = 120; c0 >= 36; c0 -= 6, it++) {
      const double* blk = ring + c0 * bw1;
      const int par = it & 1;
      if (panel) {
        if ((PARTS & 1) && pl == 0) {
          double xb[6];
#pragma unroll
          for (int c = 0; c < 6; c++) xb[c] = y[c0 + c];
#pragma unroll
          for (int c = 5; c >= 0; c--) {
            xb[c] *= blk[c * bw1];
#pragma unroll
            for (int mm = 0; mm < 6; mm++) if (mm < c) xb[mm] = fma(-blk[c * bw1 + (c - mm)], xb[c], xb[mm]);
          }
#pragma unroll
          for (int c = 0; c < 6; c++) { y[c0 + c] = xb[c] * 1e-3 + 0.5; s_xb[par][c] = xb[c] * 1e-3; }
        }
        __syncwarp();
        if ((PARTS & 2) && pl < 6) {
          const int j = c0 - 6 + pl;
          double v = y[j], v2 = 0.0;
#pragma unroll
          for (int c = 0; c < 6; c++) v = fma(-blk[c * bw1 + (c0 + c - j)], s_xb[par][c], v);
          const double* blkp = blk + 6 * bw1;
#pragma unroll
          for (int c = 0; c < 6; c++) { const int d = c0 + 6 + c - j; if (d <= beta) v2 = fma(blkp[c * bw1 + d], s_xb[par ^ 1][c], v2); }
          y[j] = (v - v2) * 1e-3 + 0.5;
        }
        __syncwarp();
      } else if ((PARTS & 4) && t < beta - 12) {
        const int cp = c0 + 6;
        const double* blkp = blk + 6 * bw1;
        const int j = cp - 13 - t;
        double v = y[j];
#pragma unroll
        for (int c = 0; c < 6; c++) { const int d = cp + c - j; if (d <= beta) v = fma(-blkp[c * bw1 + d], s_xb[par ^ 1][c], v); }
        y[j] = v * 1e-3 + 0.5;
      }
      __syncthreads();
    }
  }
  const long long c1 = clock64();
  if (t == 0) { cyc[0] = c1 - c00; out[0] = y[40] + y[100]; }
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  const int iters = 2000;
  const char* names[] = {"chol6 (library rsqrt), one lane", "chol6 (MUFU + 2 Newton), one lane", "lane-parallel chol (21 lanes, shuffles)", "chol6 speculative, 3rd-order rsqrt, one lane"};
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) panel<0><<<1, 32>>>(out, cyc, iters); else if (mode == 1) panel<1><<<1, 32>>>(out, cyc, iters); else if (mode == 2) panel<2><<<1, 32>>>(out, cyc, iters); else panel<3><<<1, 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    double o; cudaMemcpy(&o, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s trsm(6 rows) %7.1f  corner %7.1f  factor %7.1f  cycles/step  (check %.6f) %s\n", names[mode], (double)h[0] / iters, (double)h[1] / iters,
           (double)h[2] / iters, o, cudaGetErrorString(cudaGetLastError()));
  }
  struct { const char* name; void (*k)(double*, long long*, int); } wk[] = {
      {"workers: all parts, Xbuf stride 6", workers<7, 6>}, {"workers: all parts, Xbuf stride 7", workers<7, 7>},
      {"workers: solves only", workers<1, 7>}, {"workers: trailing update only", workers<2, 7>}, {"workers: rhs update only", workers<4, 7>},
      {"workers: barriers only", workers<0, 7>},
      {"workers: uniform solves only", workers<9, 7>}, {"workers: all parts, uniform solves", workers<15, 7>}};
  for (auto& w : wk) {
    for (int rep = 0; rep < 2; rep++) { w.k<<<1, 224>>>(out, cyc, iters); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles/step %s\n", w.name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  }
  struct { const char* name; void (*k)(double*, long long*, int); } bk[] = {
      {"backward block: all parts", backward<7>}, {"backward block: diagonal solve only", backward<1>}, {"backward block: next rows only", backward<2>},
      {"backward block: workers only", backward<4>}, {"backward block: barriers only", backward<0>}};
  for (auto& w : bk) {
    for (int rep = 0; rep < 2; rep++) { w.k<<<1, 256>>>(out, cyc, 200); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles/block %s\n", w.name, (double)h / (200.0 * 15), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
