// Host memory bandwidth of the box the ingest path runs on: OpenMP copy / float-narrowing passes over 256 MB.
// gcc -O3 -fopenmp scripts/host_bw.c -o /tmp/host_bw && /tmp/host_bw
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static double now(void) { return omp_get_wtime(); }
int main(void) {
  const size_t n = (size_t)32 << 20;  // doubles
  double* a = malloc(n * 8); double* b = malloc(n * 8); float* f = malloc(n * 4);
  printf("omp_get_max_threads %d, procs %d\n", omp_get_max_threads(), omp_get_num_procs());
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) { a[i] = (double)(float)(i * 0.37); b[i] = 0; f[i] = 0; }
  const int tmax = omp_get_max_threads();
  for (int threads = 1; threads <= tmax; threads *= 2) {
    omp_set_num_threads(threads);
    double t0 = now();
    for (int r = 0; r < 4; r++) {
#pragma omp parallel for schedule(static)
      for (size_t c = 0; c < n / 65536; c++) memcpy(b + c * 65536, a + c * 65536, 65536 * 8);
    }
    double t1 = now();
    int bad = 0;
    for (int r = 0; r < 4; r++) {
#pragma omp parallel for schedule(static) reduction(+ : bad)
      for (size_t c = 0; c < n / 65536; c++) { int bh = 0; for (size_t i = c * 65536; i < (c + 1) * 65536; i++) { const float v = (float)a[i]; f[i] = v; bh |= ((double)v != a[i]); } bad += bh; }
    }
    double t2 = now();
    double t3s = now();
    for (int r = 0; r < 200; r++) {
#pragma omp parallel
      { volatile int x = 0; (void)x; }
    }
    double t3 = now();
    printf("threads %2d: copy %.1f GB/s (r+w), narrow-to-float %.1f GB/s (r+w) bad %d, parallel region %.1f us\n", threads,
           4 * 2 * n * 8 / (t1 - t0) / 1e9, 4 * (n * 12.0) / (t2 - t1) / 1e9, bad, (t3 - t3s) / 200 * 1e6);
  }
  return 0;
}
