// cc_sim.cuh - device-side forward simulator, the aggregate form of
// /root/reference/R/simulators.R:114-263 (Agesim_infxn_2_death) and :1-92 (sim_seroprev).
//
// The reference draws I_t ~ Poisson(infections[t]) infections per day, splits them over the strata with a multinomial
// (probability P_a = popN_a Rho_a / sum), materialises ONE ROW PER INFECTION, and then draws per row: death (Bernoulli IFR_a),
// onset-to-death time (gamma, shape 1/s_od^2, scale m_od s_od^2), time to seroconversion (exponential, mean
// sero_delay_rate), inclusion in the serosurvey (Bernoulli smplfrac); events are binned to the day they fall in,
// (j-1, j], and events after curr_day are dropped.  By Poisson splitting / thinning every (day, stratum) cell of the
// aggregated tables is an INDEPENDENT Poisson count whose mean is the discrete convolution of the infection curve with the
// delay's interval masses - the line list is never needed:
//     Deaths[j][a]       ~ Poisson( P_a IFR_a      sum_{t<=j} infections[t] (G(j-t+1) - G(j-t)) )       G = gamma cdf
//     SeroConverts[j][a] ~ Poisson( P_a smplfrac   sum_{t<=j} infections[t] (H(j-t+1) - H(j-t)) )       H = exponential cdf
// (the same convolutions as the likelihood, binomial.cpp:187-193,288-298).
// Seroreversion (sim_seroprev, simulators.R:26-33: time from seroconversion to reversion ~ Weibull) couples the conversion
// day c and the reversion day r of one person, so the independent cells are the PAIRS (c, r):
//     N[c][r][a] ~ Poisson( P_a smplfrac sum_{t<=c} infections[t] pi(c - t, r - t) ),   r = c .. curr_day,
//     pi(u, v) = int_u^{u+1} f_exp(s) [F_weib(v + 1 - s) - F_weib(max(v - s, 0))] ds       (16-point Gauss-Legendre),
// plus the people who converted on day c and do not revert within the study.  Daily conversions are the row sums, daily
// reversions the column sums.
// Random numbers: Philox4x32-10 keyed by (seed, "SIMU", replicate, cell, draw counter) - reproducible for a seed and
// independent of the launch shape; NOT R's Mersenne-Twister stream (parity is distributional).
#pragma once
#include "cc_rng.cuh"
#include "nmath_dev.cuh"

namespace ccdev {

constexpr uint32_t kStreamSim = 0x53494d55u;  // "SIMU"

struct SimRng {
  uint64_t seed;
  uint32_t rep, cell, kind, ctr;
  ccrng::u4 buf;
  int have;
  __device__ double next() {  // uniform (0, 1), two per Philox block
    if (have == 0) {
      ccrng::u4 c;
      c.x = ctr++; c.y = kind; c.z = cell; c.w = rep;
      buf = ccrng::philox4x32_10(c, (uint32_t)seed ^ kStreamSim, (uint32_t)(seed >> 32));
      have = 2;
    }
    have--;
    return have == 1 ? ccrng::u01(buf.x, buf.y) : ccrng::u01(buf.z, buf.w);
  }
};

// Poisson(lam): sequential-search inversion below 10, Hoermann's transformed rejection (PTRS, 1993) from 10 up
__device__ inline double rpois_dev(double lam, SimRng& g) {
  if (!(lam > 0.0)) return (lam == 0.0) ? 0.0 : d_nan();
  if (!isfinite(lam)) return d_nan();
  if (lam < 10.0) {
    const double u = g.next();
    double p = exp(-lam), s = p;
    int k = 0;
    while (u > s && k < 200) {
      k++;
      p *= lam / (double)k;
      s += p;
    }
    return (double)k;
  }
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; it++) {
    const double U = g.next() - 0.5, V = g.next();
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
  }
  return floor(lam);  // unreachable in practice (acceptance > 0.9 per round)
}

// means of the per-day counts: mu_d[j] = sum_t inf[t] dG(j - t), mu_s[j] = sum_t inf[t] dH(j - t); one block
__global__ void sim_means_kernel(const double* __restrict__ inf, int T, double m_od, double s_od, double sero_rate,
                                 double* __restrict__ dG, double* __restrict__ dH, double* __restrict__ mu_d, double* __restrict__ mu_s) {
  const double alph = 1.0 / (s_od * s_od), scale = m_od * (s_od * s_od);
  const GammaShape gs = gamma_shape(alph);
  for (int k = threadIdx.x; k < T; k += blockDim.x) {
    dG[k] = pgamma_lower(gs, (double)(k + 1) / scale) - pgamma_lower(gs, (double)k / scale);
    dH[k] = exp(-(double)k / sero_rate) * (-expm1(-1.0 / sero_rate));  // H(k+1) - H(k)
  }
  __syncthreads();
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int t = 0; t <= j; t++) {
      a = fma(inf[t], dG[j - t], a);
      b = fma(inf[t], dH[j - t], b);
    }
    mu_d[j] = a;
    mu_s[j] = b;
  }
}

// one thread per (replicate, day, stratum): the two Poisson counts of the cell
__global__ void sim_draw_kernel(const double* __restrict__ mu_d, const double* __restrict__ mu_s, const double* __restrict__ w_death,
                                const double* __restrict__ w_sero, int T, int A, long long n_rep, unsigned long long seed,
                                double* __restrict__ deaths, double* __restrict__ serocon) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cells = (long long)T * A;
  if (idx >= n_rep * cells) return;
  const long long rep = idx / cells;
  const int cell = (int)(idx - rep * cells), j = cell / A, a = cell - j * A;
  SimRng g{seed, (uint32_t)rep, (uint32_t)cell, 0u, 0u, {0, 0, 0, 0}, 0};
  deaths[idx] = rpois_dev(mu_d[j] * w_death[a], g);
  g.kind = 1u; g.ctr = 0u; g.have = 0;
  serocon[idx] = rpois_dev(mu_s[j] * w_sero[a], g);
}


// ------------------------------------------------------------------------------------------------ seroreversion
// pi[u * T + v], 0 <= u <= v < T: probability that a person infected at time 0 seroconverts in (u, u+1] and reverts in
// (v, v+1].  One thread per pair.
__global__ void sim_serorev_pi_kernel(int T, double sero_rate, double shape, double scale, double* __restrict__ pi) {
  const double gx[8] = {0.0950125098376374402, 0.2816035507792589132, 0.4580167776572273863, 0.6178762444026437484,
                        0.7554044083550030339, 0.8656312023878317439, 0.9445750230732325761, 0.9894009349916499326};
  const double gw[8] = {0.1894506104550684963, 0.1826034150449235888, 0.1691565193950025382, 0.1495959888165767321,
                        0.1246289712555338720, 0.0951585116824927848, 0.0622535239386478929, 0.0271524594117540949};
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * T) return;
  const int u = (int)(idx / T), v = (int)(idx - (long long)u * T);
  double acc = 0.0;
  if (v >= u) {
    for (int q = 0; q < 16; q++) {
      const double xq = (q < 8) ? -gx[7 - q] : gx[q - 8], wq = (q < 8) ? gw[7 - q] : gw[q - 8];
      const double sq = (double)u + 0.5 + 0.5 * xq;
      const double f = exp(-sq / sero_rate) / sero_rate;
      const double hi = pweibull_lower((double)(v + 1) - sq, shape, scale), lo = pweibull_lower(fmax((double)v - sq, 0.0), shape, scale);
      acc = fma(0.5 * wq * f, hi - lo, acc);
    }
  }
  pi[idx] = acc;
}

// lam[c * T + r] (r >= c) = sum_{t <= c} inf[t] pi(c - t, r - t); lam_never[c] = mu_s[c] - sum_r lam[c][r].  One thread per (c, r).
__global__ void sim_serorev_means_kernel(const double* __restrict__ inf, const double* __restrict__ pi, int T, double* __restrict__ lam) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * T) return;
  const int c = (int)(idx / T), r = (int)(idx - (long long)c * T);
  double a = 0.0;
  if (r >= c)
    for (int t = 0; t <= c; t++) a = fma(inf[t], pi[(long long)(c - t) * T + (r - t)], a);
  lam[idx] = a;
}

// one thread per (replicate, c, stratum): everybody who converts on day c, split by reversion day (r = c .. T-1, or not
// within the study).  The daily conversions are the sum of the cells (overwriting the count drawn without reversion), the
// daily reversions are accumulated with atomics (exact: the counts are integers far below 2^53).
__global__ void sim_serorev_draw_kernel(const double* __restrict__ lam, const double* __restrict__ mu_s, const double* __restrict__ w_sero,
                                        int T, int A, long long n_rep, unsigned long long seed, double* __restrict__ serocon,
                                        double* __restrict__ serorev) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cells = (long long)T * A;
  if (idx >= n_rep * cells) return;
  const long long rep = idx / cells;
  const int cell = (int)(idx - rep * cells), c = cell / A, a = cell - c * A;
  SimRng g{seed, (uint32_t)rep, (uint32_t)cell, 2u, 0u, {0, 0, 0, 0}, 0};
  const double wa = w_sero[a];
  double tot = 0.0, lsum = 0.0;
  for (int r = c; r < T; r++) {
    const double l = lam[(long long)c * T + r];
    lsum += l;
    const double n = rpois_dev(l * wa, g);
    if (n > 0.0) {
      tot += n;
      atomicAdd(&serorev[(rep * T + r) * A + a], n);
    }
  }
  tot += rpois_dev(fmax(mu_s[c] - lsum, 0.0) * wa, g);
  serocon[idx] = tot;
}

}  // namespace ccdev
