// cc_eval.cuh - block-cooperative evaluation of ONE parameter vector (the body of the reference's loglike()).
//
// Stage numbering follows SURVEY.md 3.2 / the reference file src/natcubicspline_loglike_binomial.cpp:
//   S0 role map (Agecpp2strings.R:255-368) -> S1 attack-rate weights (:60-73) -> S2 gamma cdf table (:77-80)
//   -> S3 knot check (:88-96) -> S4 spline coefficients (:102-139) -> S5 spline + exp (:142-164)
//   -> S6 population check (:170-184) -> S7 death-delay convolution (:187-193) -> S8 Poisson L1 (:199-221)
//   -> S9 chained-binomial L2 (:227-251) -> S10 sero hazard (:257-284) -> S11/S12 sero convolution + window mean
//   (:288-311) -> S13 L3 binomial (:316-340) or logit-normal (logit.cpp:314-339) -> S14 total + clamp (:342-347).
//
// Exact algebraic restructurings relative to the reference (each changes results only at the 1e-15 level,
// SURVEY.md 8c "numerical headroom"):  ne[a] is factored out of the sero convolution; the survey-window sum uses
// the prefix-summed hazard; expd[i] = auc[i] * sum_a ne[a] ma[a]; the population check uses ne[a] * sum_t infxn[t].
#pragma once
#include "cc_model.cuh"
#include "nmath_dev.cuh"

namespace ccdev {

constexpr int kMaxK = CC_MAX_KNOTS;
constexpr int kMaxA = CC_MAX_STRATA;
constexpr int kMaxS = CC_MAX_SEROSURVEYS;

// per-vector scalars, produced by one thread and broadcast through shared memory
struct Scalars {
  double sens, spec, rate, rev_shape, rev_scale, mod, sod;
  double node_x[kMaxK], node_y[kMaxK], sp1[kMaxK], sp2[kMaxK], sp3[kMaxK];
  double cday[kMaxK];   // cday[m], m = 1..K-2: last day index evaluated with interval m-1 (interval-lag rule)
  double ne[kMaxA], ma[kMaxA];
  double snm;           // sum_a ne[a]*ma[a]
  double base[kMaxS];   // survey-window sums of the (un-weighted) sero convolution
  double strata_expd[kMaxA];
  double texpd;
  double gscale_inv;    // 1/(mod*sod^2) is NOT used for x (we divide like R); kept for the fast path
  GammaShape gs;
  int nodex_pass;
  int popn_pass;
};

// shared-memory carve-up for one evaluation
struct EvalSmem {
  double* p;      // [d]    parameter vector
  double* pg;     // [T+1]  gamma cdf table, later reused
  double* dg;     // [T]    pg[i+1]-pg[i]
  double* infxn;  // [T]
  double* auc;    // [T]
  double* haz;    // [M]    cumulative seroconversion hazard
  double* ch;     // [M+1]  prefix sums of haz
  double* tmp1;   // [M]    serorev: seroconversion density
  double* tmp2;   // [M]    serorev: 1 - weibull cdf
  double* red;    // [128]  reduction scratch (block_sum<4> needs 4 x 32)
  Scalars* sc;
};

__host__ __device__ inline size_t eval_smem_bytes(int d, int T, int M) {
  size_t n = (size_t)d + (T + 1) + 3 * (size_t)T + 4 * (size_t)M + 1 + 128 + 8;
  return n * sizeof(double) + sizeof(Scalars) + 16;
}

__device__ inline EvalSmem carve_smem(unsigned char* base, int d, int T, int M) {
  EvalSmem s;
  double* q = reinterpret_cast<double*>(base);
  s.p = q; q += d;
  s.pg = q; q += T + 1;
  s.dg = q; q += T;
  s.infxn = q; q += T;
  s.auc = q; q += T;
  s.haz = q; q += M;
  s.ch = q; q += M + 1;
  s.tmp1 = q; q += M;
  s.tmp2 = q; q += M;
  s.red = q; q += 128;
  uintptr_t a = (reinterpret_cast<uintptr_t>(q) + 15) & ~uintptr_t(15);
  s.sc = reinterpret_cast<Scalars*>(a);
  return s;
}

// ---------------------------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums NV values per thread over the block; result valid in every thread. `red` needs NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red[k * 32 + w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double t = (lane < nw) ? red[k * 32 + lane] : 0.0;
    v[k] = warp_sum(t);
  }
}

// ---------------------------------------------------------------------------------------------- S0..S4 scalars
// Role map + re-parameterisation lift (Agecpp2strings.R:277-355), attack-rate weights (binomial.cpp:60-73),
// knot check (:88-96), spline coefficients (:102-139).  Runs on ONE thread.
__device__ inline void setup_scalars(const KModel& m, const double* p, Scalars& s) {
  const int K = m.K, A = m.A;
  s.sens = p[m.idx_sens];
  s.spec = p[m.idx_spec];
  s.rate = p[m.idx_rate];
  s.rev_shape = m.serorev ? p[m.idx_shape] : -1.0;
  s.rev_scale = m.serorev ? p[m.idx_scale] : -1.0;
  s.mod = p[m.idx_mod];
  s.sod = p[m.idx_sod];
  s.node_x[0] = 1.0;
  for (int i = 0; i < K - 1; i++) {
    int r = m.idx_knots[i];
    double v = p[r];
    if (m.reparam_knots && r != m.idx_rel_knot) v = v * p[m.idx_rel_knot];
    s.node_x[i + 1] = v;
  }
  for (int i = 0; i < K; i++) {
    int r = m.idx_infxn[i];
    double v = p[r];
    if (m.reparam_infxn && r != m.idx_rel_infxn) v = v * p[m.idx_rel_infxn];
    s.node_y[i] = v;
  }
  for (int a = 0; a < A; a++) {
    int r = m.idx_ifr[a];
    double v = p[r];
    if (m.reparam_ifr && r != m.idx_max_ma) v = v * p[m.idx_max_ma];
    s.ma[a] = v;
  }
  if (A > 1) {
    double nedom = 0.0;
    for (int a = 0; a < A; a++) {
      double v = p[m.idx_noise[a]] * m.demog[a];
      s.ne[a] = v;
      nedom += v;
    }
    for (int a = 0; a < A; a++) s.ne[a] = s.ne[a] / nedom;
  } else {
    s.ne[0] = m.binomial ? 1.0 : 0.0;  // binomial.cpp:72 vs logit.cpp:71
  }
  double snm = 0.0;
  for (int a = 0; a < A; a++) snm += s.ne[a] * s.ma[a];
  s.snm = snm;

  int pass = 1;
  for (int i = 1; i < K; i++)
    if (s.node_x[i] <= s.node_x[i - 1]) pass = 0;
  s.nodex_pass = pass;
  s.popn_pass = 1;
  if (!pass) return;

  // natural cubic spline: tridiagonal sweep (local arrays are tiny: K <= 16)
  double denom[kMaxK], z[kMaxK], g[kMaxK], k[kMaxK], mm[kMaxK];
  for (int i = 0; i < K - 1; i++) denom[i] = s.node_x[i + 1] - s.node_x[i];
  z[0] = 0;
  for (int i = 1; i <= K - 2; i++)
    mm[i - 1] = (3 / denom[i]) * (s.node_y[i + 1] - s.node_y[i]) - (3 / denom[i - 1]) * (s.node_y[i] - s.node_y[i - 1]);
  g[0] = 1;
  k[0] = 0;
  for (int i = 1; i < K - 1; i++) {
    g[i] = 2 * (s.node_x[i + 1] - s.node_x[i - 1]) - denom[i - 1] * k[i - 1];
    k[i] = denom[i] / g[i];
    z[i] = (mm[i - 1] - denom[i - 1] * z[i - 1]) / g[i];
  }
  s.sp2[K - 1] = 0;
  for (int i = K - 2; i >= 0; i--) {
    s.sp2[i] = z[i] - k[i] * s.sp2[i + 1];
    s.sp1[i] = (s.node_y[i + 1] - s.node_y[i]) / denom[i] - (denom[i] * (s.sp2[i + 1] + 2 * s.sp2[i])) / 3;
    s.sp3[i] = (s.sp2[i + 1] - s.sp2[i]) / (3 * denom[i]);
  }
  // interval-lag rule in closed form: the spline index advances by at most one per day,
  // after evaluating day index i, iff 1+i >= node_x[j+1] and j < K-2 (binomial.cpp:153-159).
  double c = 0.0;
  s.cday[0] = 0.0;
  for (int mI = 1; mI <= K - 2; mI++) {
    double need = ceil(s.node_x[mI]) - 1.0;
    c = fmax(c + 1.0, need);
    s.cday[mI] = c;
  }
}

// log-incidence at day index i >= 1 (calendar day i+1), binomial.cpp:145-151
__device__ __forceinline__ double spline_day(const KModel& m, const Scalars& s, int i) {
  int j = 0;
  const double di = (double)i;
  for (int mI = 1; mI <= m.K - 2; mI++) j += (s.cday[mI] < di) ? 1 : 0;
  double dx = (double)(i + 1) - s.node_x[j];
  double dx2 = dx * dx;
  return s.node_y[j] + s.sp1[j] * dx + s.sp2[j] * dx2 + s.sp3[j] * (dx2 * dx);
}

// ---------------------------------------------------------------------------------------------- causal convolution
// out[j] = sum_{i<=j} x[i] * k[j-i], j < n, accumulated in ascending i like the reference loops
// (binomial.cpp:188-193 and :272-278).  v1: one output row per thread.
__device__ __forceinline__ void causal_conv(const double* __restrict__ x, const double* __restrict__ k, int n,
                                            double* __restrict__ out) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i <= j; i++) acc = fma(x[i], k[j - i], acc);
    out[j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- the evaluation
// All threads of the block call this with the parameter vector already in sm.p[0..d) and a __syncthreads()
// behind it.  Returns the log-likelihood in every thread; sm.infxn / sm.auc / sm.haz / sm.sc->ne are left valid for the
// caller.  `popn_exit` = false (curve extraction, R/summarize_plot.R:729 drops the population check) records a failed
// population check in sc.popn_pass and carries on instead of returning the failure value.
__device__ inline double block_loglike(const KModel& m, EvalSmem& sm, bool popn_exit = true) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int T = m.T, M = m.M, A = m.A, S = m.S;
  Scalars& sc = *sm.sc;

  // ---- S0,S1,S3,S4 on one thread; the gamma-shape constants on another warp
  if (tid == 0) setup_scalars(m, sm.p, sc);
  if (tid == 32 % nt) {
    double sod = sm.p[m.idx_sod];
    sc.gs = gamma_shape(1.0 / (sod * sod));
  }
  __syncthreads();
  const int pass = sc.nodex_pass;

  // ---- S2 gamma cdf table (always built in the reference; its value is irrelevant after an early exit)
  if (pass) {
    const double scale = sc.mod * (sc.sod * sc.sod);
    const GammaShape gs = sc.gs;
    const bool bad = isnan(scale) || isnan(gs.alph) || !(scale > 0.0) || gs.alph < 0.0;
    for (int i = tid; i <= T; i += nt) {
      double v;
      if (bad) v = (isnan(scale) || isnan(gs.alph)) ? (scale + gs.alph) : d_nan();
      else v = pgamma_lower(gs, (double)i / scale);
      sm.pg[i] = v;
    }
  }

  // ---- S5 spline + exp, S6 population check
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  if (pass) {
    for (int i = tid; i < T; i += nt) {
      double y = (i == 0) ? sc.node_y[0] : spline_day(m, sc, i);
      double v = exp(y);
      sm.infxn[i] = v;
      part[0] += v;
    }
  }
  block_sum(part, sm.red);  // also orders pg/infxn writes before the reads below
  if (!pass) return -CC_OVERFLO;
  {
    const double tot = part[0];
    int ok = 1;
    for (int a = 0; a < A; a++)
      if (sc.ne[a] * tot > (double)m.demog[a]) ok = 0;
    if (!ok && popn_exit) return -CC_OVERFLO;  // block-uniform: every thread sees the same `tot`
    if (!ok && tid == 0) sc.popn_pass = 0;
  }

  // ---- S7 death-delay convolution
  for (int i = tid; i < T; i += nt) sm.dg[i] = sm.pg[i + 1] - sm.pg[i];
  __syncthreads();
  causal_conv(sm.infxn, sm.dg, T, sm.auc);

  // ---- S10 sero hazard (independent of S7; uses other buffers)
  if (m.serorev) {
    for (int dd = tid; dd < M; dd += nt) {
      sm.tmp1[dd] = (1 / sc.rate) * exp(-(double)(dd + 1) / sc.rate);
      sm.tmp2[dd] = 1 - pweibull_lower((double)dd, sc.rev_shape, sc.rev_scale);
    }
    __syncthreads();
    causal_conv(sm.tmp1, sm.tmp2, M, sm.haz);
  } else {
    for (int dd = tid; dd < M; dd += nt) sm.haz[dd] = 1 - exp(-(double)(dd + 1) / sc.rate);
  }
  __syncthreads();

  // prefix sums ch[k] = sum_{u<k} haz[u] (serial per chunk + block scan of chunk totals)
  {
    const int chunk = (M + nt - 1) / nt;
    const int lo = min(tid * chunk, M), hi = min(lo + chunk, M);
    double loc = 0.0;
    for (int u = lo; u < hi; u++) loc += sm.haz[u];
    // inclusive warp scan
    const int lane = tid & 31, w = tid >> 5, nw = (nt + 31) >> 5;
    double incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sm.red[w] = incl;
    __syncthreads();
    double woff = 0.0;
    for (int k = 0; k < w; k++) woff += sm.red[k];
    (void)nw;
    double run = woff + incl - loc;  // exclusive prefix of this thread's chunk
    for (int u = lo; u < hi; u++) {
      sm.ch[u] = run;
      run += sm.haz[u];
    }
    if (hi == M && lo < M) sm.ch[M] = run;
    if (M == 0 && tid == 0) sm.ch[0] = 0.0;
  }
  __syncthreads();

  // ---- S8 L1 (Poisson) + sums for S9; S11/S12 survey sums; one fused block reduction
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // L1, sum expd (all days), sum auc over uncensored days, unused
  for (int i = tid; i < T; i += nt) {
    const double auc = sm.auc[i];
    const double expd = auc * sc.snm;
    acc[1] += expd;
    if ((i + 1) < m.rcensor_day) {
      acc[2] += auc;
      const int x = m.obsd[i];
      if (x != -1) {
        double lp;
        const double xd = (double)x;
        if (isnan(expd)) lp = expd;
        else if (expd < 0) lp = d_nan();
        else if (x < 0) lp = -d_inf();
        else if (expd == 0) lp = (x == 0) ? 0.0 : -d_inf();
        else if (!isfinite(expd)) lp = -d_inf();
        else if (xd <= expd * DBL_MIN) lp = -expd;
        else if (expd < xd * DBL_MIN) lp = -expd + xd * log(expd) - m.lgx1[i];
        else lp = m.pois_c[i] - bd0(xd, expd);
        acc[0] += lp;
      }
    }
  }
  block_sum(acc, sm.red);

  double sv[kMaxS];
  // survey sums: base[s] = sum_{t<e} infxn[t] * (ch[e-t] - ch[max(s-1-t,0)])
  for (int s = 0; s < S; s += 4) {
    double b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (s + q < S) {
        const int st = m.sero_start[s + q], en = m.sero_end[s + q];
        for (int t = tid; t < en; t += nt) {
          int lo = st - 1 - t;
          lo = lo < 0 ? 0 : lo;
          b[q] = fma(sm.infxn[t], sm.ch[en - t] - sm.ch[lo], b[q]);
        }
      }
    }
    block_sum(b, sm.red);
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (s + q < S) sv[s + q] = b[q] / (double)(m.sero_end[s + q] - m.sero_start[s + q] + 1);
  }

  // ---- S9 L2 + S13 L3, one (stratum) or (survey, stratum) term per thread
  double tail[2] = {0.0, 0.0};
  const double texpd = acc[1];
  for (int a = tid; a < A - 1; a += nt) {
    // running subtraction texpd -= strata_expd[b] for b < a (binomial.cpp:249), same order as the reference
    double te = texpd;
    for (int b = 0; b < a; b++) te -= (sc.ne[b] * sc.ma[b]) * acc[2];
    const double se = (sc.ne[a] * sc.ma[a]) * acc[2];
    tail[0] += dbinom_log(round(se), round(te), m.l2_prob[a]);
  }
  for (int k = tid; k < S * A; k += nt) {
    const int s = k / A, a = k - s * A;
    const double num = sc.ne[a] * sv[s];
    const double frac = num / (double)m.demog[a];
    const double prev = sc.sens * frac + (1 - sc.spec) * (1 - frac);
    const double da = m.sero_a[k], db = m.sero_b[k];
    if (m.binomial) {
      if (da != -1 || db != -1) tail[1] += dbinom_log(da, db, prev);
    } else {
      const double eps = DBL_MIN / 100.0;
      const double lg = log((prev + eps) / (1 - (prev + eps)));
      tail[1] += dnorm_log(lg, da, db);
    }
  }
  block_sum(tail, sm.red);

  double loglik = acc[0] + tail[0] + tail[1];
  if (!isfinite(loglik)) loglik = -CC_OVERFLO;
  return loglik;
}

// ---------------------------------------------------------------------------------------------- prior
// Generated logprior (Agecpp2strings.R:66-219): table-driven sum + Jacobians + clamp.  One thread.
__device__ inline double thread_logprior(const KModel& m, const double* p, int stride) {
  double ret = 0.0;
  for (int i = 0; i < m.d; i++) {
    const int fam = m.prior_family[i];
    if (fam == CC_PRIOR_NONE) continue;
    const double x = p[(size_t)i * stride], a = m.prior_p1[i], b = m.prior_p2[i];
    double v;
    if (fam == CC_PRIOR_DUNIF) v = dunif_log(x, a, b);
    else if (fam == CC_PRIOR_DNORM) v = dnorm_log(x, a, b);
    else v = dbeta_log(x, a, b);
    ret += v;
  }
  for (int j = 0; j < 3; j++)
    if (m.jac_rows[j] >= 0) ret += m.jac_counts[j] * log(p[(size_t)m.jac_rows[j] * stride]);
  if (!isfinite(ret)) ret = -CC_OVERFLO;
  return ret;
}

}  // namespace ccdev
