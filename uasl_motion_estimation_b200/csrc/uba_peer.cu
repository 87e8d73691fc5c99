// uba_peer.cu — the two cross-rank exchanges of a point-sharded LM iteration, done by plain kernels over NVLink
// PEER MEMORY instead of library collectives ("allreduce of the reduced camera system only where
// points span shards").  One process per GPU; every rank has mapped the other ranks' accumulator block, arrival words
// and inbox through CUDA IPC (uba_host.cu: peer_exchange).  See PeerView in uba_device.h for the protocol.
//
// Why not ncclAllReduce: the message is ~0.3 MB (the block band of the reduced camera system of a 200-keyframe
// window) twice per 0.3 ms iteration; a library collective costs 20-40 us of launch + protocol latency each and cannot
// sit inside the iteration's CUDA graph without slowing it further (measured in round 1).  A pull over NVLink of the
// stretches of the band the neighbouring ranks produced is a few microseconds, and the arrive/wait words make it a
// kernel like any other in the captured graph.
#include <cuda_runtime.h>
#include <stdint.h>

#include "uba_device.h"

namespace uba {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ long long now_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Warp 0 of a CTA (all 32 lanes call this): announce exchange `e` to every rank (only when `announce`), then wait until
// every rank has announced it.  Lane p talks to rank p, so the n arrival stores travel over NVLink side by side and the n
// local words are polled side by side: one round trip instead of n of them (the serial version cost ~18 us per rank and
// made 8 GPUs slower than 4).  Returns false when a peer stays silent for longer than the time-out (or an earlier exchange
// failed).
__device__ bool arrive_and_wait(const PeerView& P, unsigned long long e, bool announce) {
  const int lane = threadIdx.x & 31;
  bool ok = *(volatile int32_t*)P.err == 0;
  if (ok && announce) {
    __threadfence_system();                   // everything this rank produced before the exchange is visible system-wide
    for (int p = lane; p < P.n_ranks; p += 32) st_relaxed_sys(&P.flags[p][P.rank], e);
  }
  if (ok) {
    const long long t0 = now_ns();
    const unsigned long long* mine = P.flags[P.rank];
    for (int p = lane; p < P.n_ranks; p += 32) {
      while (ld_acquire_sys(&mine[p]) < e) {
        if (now_ns() - t0 > P.timeout_ns) { atomicExch(P.err, 1 + p); ok = false; break; }
        __nanosleep(32);
      }
    }
  }
  return __all_sync(0xffffffffu, ok);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Exchange 1 (after the lineariser): acc_red[0, sum_end) = sum over ranks of their partial accumulators, in rank
// order, reading only from the ranks whose shard observes the element's camera.  Banded windows: only the band of the
// Schur accumulator exists.  Also zeroes the consumer-side scalars behind sum_end.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peer_reduce(DevView V, PeerView P, AccLayout L, double* __restrict__ acc_red, int dense) {
  __shared__ unsigned long long s_e;
  __shared__ int s_ok;
  const int t = threadIdx.x;
  if (t < 32) {
    const unsigned long long e = *(volatile unsigned long long*)&P.epoch[0] + 1;
    const bool ok = arrive_and_wait(P, e, blockIdx.x == 0);
    if (t == 0) { s_e = e; s_ok = ok ? 1 : 0; }
  }
  __syncthreads();
  if (s_ok) {
    const int nr = P.n_ranks;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool single = V.nW == 1;
    const int n = single ? 6 * (V.w_free_off[1] - V.w_free_off[0]) : 0;
    const int beta = single && !dense ? V.w_beta[0] : 0;
    // elements the lineariser can have produced: the per-camera tail, the cost words, and per window the band (or the
    // dense upper triangle) of the Schur accumulator
    const int64_t sacc_cnt = single ? (beta > 0 ? (int64_t)n * (beta + 1) : (int64_t)n * n) : L.sum_end - L.off_sacc;
    const int64_t total = L.off_sacc + sacc_cnt;
    for (int64_t off = (int64_t)blockIdx.x * blockDim.x + t; off < total; off += stride) {
      int cam = -1;                                  // -1: every rank contributes
      if (off < L.off_vacc) cam = (int)(off / 36);
      else if (off < L.off_zh) cam = (int)((off - L.off_vacc) / 6);
      else if (off < L.off_wlin) cam = (int)((off - L.off_zh) / 6);
      else if (off >= L.off_sacc && single) {
        const int64_t e = off - L.off_sacc;
        const int row = beta > 0 ? (int)(e / (beta + 1)) : (int)(e / n);
        cam = V.free_list[row / 6];
      }
      // all the remote loads first (they travel side by side), then the sum in rank order: identical on every rank
      double v[kMaxRanks];
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) {
        const bool on = r < nr && !(cam >= 0 && (cam < P.cam_lo[r] || cam > P.cam_hi[r]));   // off: that rank's shard never touches this camera
        v[r] = on ? ld_peer(P.acc[r] + off) : 0.0;
      }
      double sum = 0.0;
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) sum += v[r];
      acc_red[off] = sum;
    }
    if (blockIdx.x == 0)
      for (int64_t off = L.off_wpost + t; off < L.total; off += blockDim.x) acc_red[off] = 0.0;
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(&P.epoch[2], 1ull) == (unsigned long long)gridDim.x - 1) {   // last CTA out closes the exchange
      P.epoch[2] = 0;
      *(volatile unsigned long long*)&P.epoch[0] = s_e;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Exchange 2 (after the back-substitution): every rank pushes its per-window partial scalars into every rank's inbox
// (double-buffered by the post count, so a fast rank's next iteration cannot overwrite what a slow rank still reads),
// then sums the inbox in rank order into the consumer view.  One CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_peer_post(DevView V, DevView Vc, PeerView P) {
  __shared__ int s_ok;
  const int t = threadIdx.x;
  const int nr = P.n_ranks, nW = V.nW;
  const unsigned long long e = *(volatile unsigned long long*)&P.epoch[0] + 1;
  const int par = (int)(*(volatile unsigned long long*)&P.epoch[1] & 1ull);
  const int items = nW * kInboxSlots;
  const double stop = *P.stop_req;
  for (int i = t; i < items; i += blockDim.x) {
    const int w = i / kInboxSlots, s = i - w * kInboxSlots;
    const double val = s < WP_STOP ? V.w_post[(size_t)w * WP_COUNT + s] : (s == WP_STOP ? stop : V.w_max[w]);
    const size_t slot = ((size_t)(par * nr + P.rank) * nW + w) * kInboxSlots + s;
    for (int p = 0; p < nr; p++) st_peer(P.inbox[p] + slot, val);
  }
  __threadfence_system();
  __syncthreads();
  if (t < 32) { const bool ok = arrive_and_wait(P, e, true); if (t == 0) s_ok = ok ? 1 : 0; }
  __syncthreads();
  if (s_ok) {
    const double* in = P.inbox[P.rank];
    for (int i = t; i < items; i += blockDim.x) {
      const int w = i / kInboxSlots, s = i - w * kInboxSlots;
      double v[kMaxRanks];
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) v[r] = r < nr ? ld_peer(in + ((size_t)(par * nr + r) * nW + w) * kInboxSlots + s) : 0.0;
      double sum = 0.0, mx = 0.0;
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) { sum += v[r]; mx = fmax(mx, v[r]); }
      if (s < WP_COUNT) Vc.w_post[(size_t)w * WP_COUNT + s] = sum;
      else Vc.w_max[w] = fmax(Vc.w_max[w], mx);       // the camera part is already there (k_assemble, identical on every rank)
    }
  }
  __syncthreads();
  if (t == 0) {
    *(volatile unsigned long long*)&P.epoch[0] = e;
    *(volatile unsigned long long*)&P.epoch[1] = *(volatile unsigned long long*)&P.epoch[1] + 1;
  }
}

// Plain rendezvous of the ranks on the device (no payload).  The benchmark puts one between the L2 flush and the start
// event of every timed iteration, so that an iteration's exchanges do not wait for a peer that is still flushing.
__global__ void k_peer_barrier(PeerView P) {
  const unsigned long long e = *(volatile unsigned long long*)&P.epoch[0] + 1;
  arrive_and_wait(P, e, true);
  __syncwarp();
  if (threadIdx.x == 0) *(volatile unsigned long long*)&P.epoch[0] = e;
}

int launch_peer_barrier(const PeerView& P, cudaStream_t st) {
  k_peer_barrier<<<1, 32, 0, st>>>(P);
  return 1;
}

int launch_peer_reduce(const DevView& V, const PeerView& P, const AccLayout& L, double* acc_red, int dense, cudaStream_t st) {
  const int64_t blocks = (L.sum_end + 255) / 256;
  const int grid = (int)(blocks < 296 ? (blocks > 0 ? blocks : 1) : 296);
  k_peer_reduce<<<grid, 256, 0, st>>>(V, P, L, acc_red, dense);
  return 1;
}

int launch_peer_post(const DevView& V, const DevView& Vc, const PeerView& P, cudaStream_t st) {
  k_peer_post<<<1, 128, 0, st>>>(V, Vc, P);
  return 1;
}

}  // namespace uba
