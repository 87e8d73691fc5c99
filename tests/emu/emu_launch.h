// tests/emu/emu_launch.h — serial "launch" of a thread-independent kernel + host stand-ins for the
// shared-memory Cholesky kernels.  Test infrastructure only (see cuda_runtime.h in this directory).
#pragma once
#include <vector>
#include "uba_device.h"

namespace uba_emu {
template <typename F>
inline void launch(dim3 grid, dim3 block, F&& body) {
  t_gridDim = grid; t_blockDim = block;
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++)
      for (unsigned tx = 0; tx < block.x; tx++) {
        t_blockIdx = dim3(bx, by, 0); t_threadIdx = dim3(tx, 0, 0);
        body();
      }
}
// dense Cholesky solve of window w (stand-in for k_chol_small / k_chol_* / k_trsv_large)
inline void dense_solve(const uba::DevView& V, int w) {
  if (V.ws[w].done) return;
  const int f0 = V.w_free_off[w];
  const int n = 6 * (V.w_free_off[w + 1] - f0);
  if (n == 0) return;
  double* A = V.A + V.w_red_off[w];
  double* y = V.rhs + (size_t)6 * f0;
  bool failed = false;
  for (int j = 0; j < n && !failed; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) { failed = true; break; }
    d = std::sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  if (failed) { V.w_loc[(size_t)w * uba::WC_COUNT + uba::WC_FAIL] += 1.0; for (int i = 0; i < n; i++) y[i] = 0.0; return; }
  for (int i = 0; i < n; i++) { double s = y[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k]; y[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * y[k]; y[i] = s / A[(size_t)i * n + i]; }
}
}  // namespace uba_emu

namespace uba {
inline double warp_sum(double v) { return v; }
inline double warp_max(double v) { return v; }
inline bool warp_leader() { return true; }
}
#define UBA_LAUNCH(kern, grid, block, smem, st, ...) uba_emu::launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); })
