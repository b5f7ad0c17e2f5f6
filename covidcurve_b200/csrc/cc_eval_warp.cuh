// cc_eval_warp.cuh - warp-level building blocks of the hot path: the Toeplitz products on the FP64 tensor cores and the
// shared-memory layout they read.  (The per-vector evaluation built from them lives in cc_eval_tile.cuh.)
//
// Stage map (line numbers: /root/reference/src/natcubicspline_loglike_binomial.cpp):
//   S0 role map (R/Agecpp2strings.R:255-368)  S1 attack-rate weights (:60-73)   S3 knot check (:88-96)
//   S4 spline coefficients (:102-139)         S2 gamma cdf table (:77-80)       S5 spline + exp (:142-164)
//   S6 population check (:170-184)            S7 death-delay convolution (:187-193)
//   S8 Poisson L1 (:199-221)                  S9 chained-binomial L2 (:227-251) S10 sero hazard (:257-284)
//   S11/S12 sero convolution + window mean (:288-311)   S13 L3 binomial (:316-340) | logit-normal (logit.cpp:314-339)
//   S14 total + clamp (:342-347)
//
// How the work is laid out on a B200 SM:
//   * one warp = one evaluation; 4 warps per CTA share nothing but the model constants, so the only synchronisation
//     is __syncwarp().  ~10-15 kB of shared memory per warp keeps 12-16 evaluations in flight per SM.
//   * S7 (and the seroreversion hazard convolution of S10) is a lower-triangular Toeplitz matrix-vector product.  It is
//     recast as a sum over lag blocks delta of (8x8 Toeplitz block G_delta) x (8 x T/8 matrix X of the series cut in
//     blocks of 8 days), which is a real matrix-matrix product and runs on the FP64 tensor cores:
//         mma.sync.m8n8k4.f64   D[r][p] += sum_s g[8 delta + r - s] * x[8 (p - delta) + s]
//     The series lives in shared memory with a 12-double pitch per 8-day block (bank-conflict-free B fragments) behind
//     7 zero blocks (so partial tiles need no predication); the kernel lives behind 7 zeros (negative lags).  On B200 the
//     DMMA and DFMA instructions share the FP64 pipe (measured: profiles/r1_fp64_probes.txt), so this does not raise
//     the peak - it removes the lane imbalance of the triangle and 7/8 of the issue slots.
//   * S2: the T+1 regularised incomplete gamma values share their shape and sit on an equispaced grid: series below the
//     median, exactly integrated interval masses above it, one coefficient table per evaluation (cc_eval_tile.cuh,
//     tile_gamma); shapes outside the validated box fall back to the general algorithm (nmath_dev.cuh pgamma_lower).
//   * S8: dpois(x, lambda, log) = x log(lambda) - lambda - lgamma(x+1) with the data-only part summed on the host;
//     S13: dbinom(pos, n, p, log) = lchoose(n, pos) + pos log p + (n - pos) log(1 - p) with lchoose on the host.  Both
//     are the same functions as R's saddle-point forms to ~1e-13 absolute, far inside the 1e-10 relative budget
//     (the log-likelihood is O(1e3) or larger); every special case of R's dpois/dbinom ends in the same value.
#pragma once
#include "cc_eval.cuh"

namespace ccdev {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ inline int warp_tiles(int T) { return (T + 63) / 64; }  // 64-day output tiles of the Toeplitz product

constexpr int kXsLead = 84;  // 7 zero blocks x 12 doubles in front of the padded series
__device__ __forceinline__ int xs_index(int i) { return i + ((i >> 3) << 2) + kXsLead; }

// ------------------------------------------------------------------------------------------------ Toeplitz DMMA
// acc[P][e] (+)= sum over lags of g * x for output day j = 64 P + 16 (lane%4) + 8 e + lane/4;  nb = ceil(len/8)
template <int NT>
__device__ __forceinline__ void toeplitz_dmma(const double* __restrict__ xs, const double* __restrict__ gs, int nb, double (&acc)[NT][2]) {
  const int lane = threadIdx.x & 31, n = lane >> 2, t = lane & 3;
  const double* ap = gs + 7 + n - t;           // A[r = n][s = 4h + t] = g[8 delta + n - 4h - t]
  const double* bp = xs + kXsLead + 12 * n + t;  // B[s = 4h + t][col n] = x[8 (8P + n - delta) + 4h + t]
  // two accumulator sets (one per k-half) keep twice as many independent MMA chains in flight; summed at the end
  double acc1[NT][2];
#pragma unroll
  for (int P = 0; P < NT; P++) acc[P][0] = acc[P][1] = acc1[P][0] = acc1[P][1] = 0.0;
  // lag dl = 8 D + e needs series block row P - D at offset e: with e outermost the NT B-fragment pairs are loaded once
  // per e and shared by all D (half the shared-memory loads of the D-outer order)
#pragma unroll 2
  for (int e = 0; e < 8; e++) {
    const double* b = bp - 12 * e;
    double b0[NT], b1[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) {
      b0[q] = b[96 * q];
      b1[q] = b[96 * q + 4];
    }
#pragma unroll
    for (int D = 0; D < NT; D++) {
      const int dl = 8 * D + e;
      if (dl >= nb) break;
      const double a0 = ap[8 * dl], a1 = ap[8 * dl - 4];
#pragma unroll
      for (int P = D; P < NT; P++) {
        dmma884(acc[P][0], acc[P][1], a0, b0[P - D]);
        dmma884(acc1[P][0], acc1[P][1], a1, b1[P - D]);
      }
    }
  }
#pragma unroll
  for (int P = 0; P < NT; P++) {
    acc[P][0] += acc1[P][0];
    acc[P][1] += acc1[P][1];
  }
}

}  // namespace ccdev
