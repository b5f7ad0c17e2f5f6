// cc_mcmc_dev.cuh - device pieces of the GPU-resident Metropolis-coupled MCMC that replaces drjacoby::run_mcmc
// (external engine, mrc-ide/drjacoby@e83884e; call site /root/reference/R/main.R:179-193; algorithm as restated in
// SURVEY.md 3.1): parameter transforms, single-site Metropolis-Hastings with Robbins-Monro bandwidth adaptation,
// and the adjacent-rung Metropolis-coupling pass.
#pragma once
#include "cc_eval.cuh"
#include "cc_rng.cuh"

namespace ccdev {

// drjacoby transform type from the finiteness of (min, max): 0 none, 1 max only, 2 min only, 3 both
__host__ __device__ inline int trans_type(double lo, double hi) {
  const bool flo = isfinite(lo), fhi = isfinite(hi);
  return flo ? (fhi ? 3 : 2) : (fhi ? 1 : 0);
}

__host__ __device__ inline double theta_to_phi(int tt, double th, double lo, double hi) {
  switch (tt) {
    case 1: return log(hi - th);
    case 2: return log(th - lo);
    case 3: return log(th - lo) - log(hi - th);
    default: return th;
  }
}

__host__ __device__ inline double phi_to_theta(int tt, double ph, double lo, double hi) {
  switch (tt) {
    case 1: return hi - exp(ph);
    case 2: return lo + exp(ph);
    case 3: { const double e = exp(ph); return (hi * e + lo) / (1 + e); }
    default: return ph;
  }
}

// log Jacobian ratio of the proposal in phi-space
__host__ __device__ inline double mh_adjust(int tt, double th, double thp, double lo, double hi) {
  switch (tt) {
    case 1: return log((hi - thp) / (hi - th));
    case 2: return log((thp - lo) / (th - lo));
    case 3: return log(((hi - thp) * (thp - lo)) / ((hi - th) * (th - lo)));  // = the four-log form of drjacoby, one log
    default: return 0.0;
  }
}

// generated logprior, evaluated cooperatively: one term per thread, block reduction, clamp
// (Agecpp2strings.R:66-219).  Valid in every thread.
__device__ inline double block_logprior(const KModel& m, const double* p, double* red) {
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < m.d; i += blockDim.x) {
    const int fam = m.prior_family[i];
    if (fam != CC_PRIOR_NONE) {
      const double x = p[i], a = m.prior_p1[i], b = m.prior_p2[i];
      v[0] += (fam == CC_PRIOR_DUNIF) ? dunif_log(x, a, b) : (fam == CC_PRIOR_DNORM) ? dnorm_log(x, a, b) : dbeta_log(x, a, b);
    }
  }
  if (threadIdx.x < 3 && m.jac_rows[threadIdx.x] >= 0) v[0] += m.jac_counts[threadIdx.x] * log(p[m.jac_rows[threadIdx.x]]);
  block_sum(v, red);
  double ret = v[0];
  if (!isfinite(ret)) ret = -CC_OVERFLO;
  return ret;
}

// device-resident run state (pointers into one allocation owned by the host-side cc_mcmc)
struct McmcState {
  int d, rungs, n_chains, chain_offset;
  int burnin, samples, coupling_on, record_all, thin, bw_per_particle, adapt_in_sampling;
  unsigned long long seed;
  double* theta;      // [P][d] current state of particle p = chain*rungs + id
  double* phi;        // [P][d]
  double* ll;         // [P]
  double* lp;         // [P]
  double* bw;         // [P][d] proposal bandwidths (slot = rung position, or particle when bw_per_particle)
  int* bw_index;      // [P][d] Robbins-Monro counters
  unsigned int* move_acc;  // [P][d] accepted single-site moves (slot = rung position)
  int* at_rung;       // [n_chains][rungs] particle id (within chain) currently at rung position r
  double* beta;       // [rungs] thermodynamic powers (already raised to GTI_pow), ascending when cold_rung_last
  unsigned int* swap_acc;  // [n_chains][rungs-1] accepted swaps during sampling
  unsigned int* swap_try;  // [n_chains][rungs-1]
  // recording
  int n_rec_rungs;    // rungs if record_all else 1
  int cold_pos;       // rung position of the cold chain (beta = 1)
  long long rows_cap; // rows allocated per (chain, recorded rung)
  double* rec_theta;  // [n_chains][n_rec_rungs][rows_cap][d]
  double* rec_ll;     // [n_chains][n_rec_rungs][rows_cap]
  double* rec_lp;
  // cached evaluation state of every particle's current parameter vector (layout: sweep_cache_layout), zero = not built yet
  double* cache;
  int* ticket;        // particle queue of the sweep kernel (reset before every launch)
};

}  // namespace ccdev
