// cc_model.cuh - kernel-side view of one model (passed BY VALUE as a kernel parameter, ~1.5 kB) and
// the host-side handle behind the opaque `cc_model` of include/covidcurve_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/covidcurve_b200.h"
#include "cc_fastmath.cuh"

namespace ccdev {

// Everything a kernel needs that is shared by the whole batch: misc_list + data_list + role map + prior table.
// Small fixed-size arrays travel in the parameter (constant) bank; per-day tables live in global memory and
// are read through the read-only path (they stay L1/L2 resident: <= 4 x T doubles).
struct KModel {
  int d, A, K, T, S, M;
  int rcensor_day, serorev, binomial;
  int reparam_ifr, reparam_knots, reparam_infxn;
  int idx_sens, idx_spec, idx_rate, idx_shape, idx_scale, idx_mod, idx_sod;
  int idx_rel_knot, idx_rel_infxn, idx_max_ma;
  short idx_knots[CC_MAX_KNOTS];   // rows of node_x[1..K-1]
  short idx_infxn[CC_MAX_KNOTS];   // rows of node_y[0..K-1]
  short idx_ifr[CC_MAX_STRATA];    // rows of ma[]
  short idx_noise[CC_MAX_STRATA];  // rows of ne[]
  int demog[CC_MAX_STRATA];        // popN truncated to int (binomial.cpp:10)
  double l2_prob[CC_MAX_STRATA];   // paobsd[a] / (running paobsd_sum), binomial.cpp:248-250 (data only)
  int sero_start[CC_MAX_SEROSURVEYS];
  int sero_end[CC_MAX_SEROSURVEYS];
  double* debug_dump;              // developer aid: when non-null, vector 0 of a batch dumps dg[T], infxn[T], auc[T], CH[T] here
  int debug_stop;                  // developer switch (CC_DEBUG_STOP): stop the evaluation after stage n (0 = run everything)
  int gamma_mode;                  // 0: fast shared-shape evaluation with fallback, 1: always the general (R region-split) algorithm
  int jac_rows[3];
  double jac_counts[3];
  // device tables
  const int* obsd;        // [T] observed deaths as int, -1 = missing (binomial.cpp:204)
  const double* pois_c;   // [T] -0.5*log(2*pi*x) - stirlerr(x) for x > 0, else 0   (x-only part of dpois)
  const double* lgx1;     // [T] lgamma(x+1)
  const double* logk;     // [T+1] log(k), k >= 1 (entry 0 unused): log of the integer days / lags
  const LogTab* logtab;   // [128] table of fast_log (cc_fastmath.cuh)
  const double* rk;       // [max(T,640)+2] 1/k, k >= 1
  const int* dep;         // [d] stages a proposal of parameter i invalidates (kDep* bits, Metropolis sweep)
  double l1_sumx;         // sum of the observed counts x_j > 0 over the days that enter L1
  double l1_const;        // sum of lgamma(x+1) over the days that enter L1 (uncensored, observed, x >= 0)
  const int* l3_kind;     // [S*A] 0: skipped (pos == n == -1), 1: regular counts, 2: irregular -> R's full dbinom case analysis
  const double* l3_lchoose;  // [S*A] lchoose(n, pos) for regular counts
  const double* sero_a;   // [S*A]
  const double* sero_b;   // [S*A]
  const double* pmin;     // [d]
  const double* pmax;     // [d]
  const int* prior_family;  // [d]
  const double* prior_p1;   // [d]
  const double* prior_p2;   // [d]
};

}  // namespace ccdev

struct cc_model {
  int device = 0;
  ccdev::KModel k{};
  std::vector<void*> dev_allocs;   // everything cudaMalloc'ed for the tables
  cudaStream_t stream = nullptr;   // library-owned stream
  cudaStream_t stream2 = nullptr;  // second and third stream of the triple-buffered host path
  cudaStream_t stream3 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_slot0 = nullptr;  // end of the last device-pointer call (serialises calls that arrive on different streams)
  cudaStream_t slot0_stream = nullptr;
  bool slot0_used = false;
  double last_ms = 0.0;
  int last_launches = 0;
  // host staging (pinned) + device staging for CC_MEM_HOST calls
  double* h_pin = nullptr; size_t h_pin_bytes = 0;
  double* d_stage = nullptr; size_t d_stage_bytes = 0;
  double* d_out = nullptr; size_t d_out_bytes = 0;
  std::vector<double> pinit;
  int sm_count = 0;
  int grid_p2 = 1;                 // resident CTAs of pass C of the batched likelihood
  int p2_warps = 4;                // warps per CTA of pass C (chosen at creation, covidcurve_b200.cu: p2_candidates)
  int grid_max = 1;                // resident CTAs of the evaluation kernel on this device
  // work buffers of the batched likelihood, one set per launch slot ([0] device-pointer calls, [1]..[3] the three host-path
  // streams): records rec[F][cap], regime keys, permutation, histogram + cursors; grown on demand
  struct Work { double* rec = nullptr; int* key = nullptr; int* perm = nullptr; unsigned int* counts = nullptr; long long cap = 0; };
  Work work[4];
  int nt = 0;                      // 64-day output tiles per evaluation (template parameter of the warp kernels)
};
