// tests/cuda/hostbw.c — host memory bandwidth floor for the ingest path (not part of libuba).
//   gcc -O3 -fopenmp -o hostbw hostbw.c && ./hostbw
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main() {
  const size_t n = 6u << 20;  // 48 MB of doubles
  double* a = malloc(n * 8); double* b = malloc(n * 8);
  for (size_t i = 0; i < n; i++) { a[i] = i; b[i] = 0; }
  printf("max threads %d\n", omp_get_max_threads());
  const int maxth = omp_get_max_threads();
  for (int th = 1; th <= maxth; th *= 2) {
    omp_set_num_threads(th);
    double best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
      double t0 = now();
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < n; i++) b[i] = a[i];
      double t1 = now(); if (t1 - t0 < best) best = t1 - t0;
    }
    // AoS[n/4][4] -> SoA[4][n/4]
    double bestT = 1e9; const size_t q = n / 4;
    for (int rep = 0; rep < 6; rep++) {
      double t0 = now();
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < q; i++) { b[i] = a[4 * i]; b[q + i] = a[4 * i + 1]; b[2 * q + i] = a[4 * i + 2]; b[3 * q + i] = a[4 * i + 3]; }
      double t1 = now(); if (t1 - t0 < bestT) bestT = t1 - t0;
    }
    double t0 = now();
#pragma omp parallel
    { volatile int x = 0; (void)x; }
    double tr = now() - t0;
    printf("threads %2d: copy 48MB %.2f ms (%.1f GB/s r+w), transpose %.2f ms, empty region %.1f us\n", th, best * 1e3, 2 * n * 8 / best / 1e9, bestT * 1e3, tr * 1e6);
  }
  printf("%f\n", b[123]);
  return 0;
}
