// cc_eval_warp.cuh - one WARP evaluates one parameter vector (the hot-path form of the reference's loglike()).
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
//   * S2: the T+1 regularised incomplete gamma values share their shape, so the ascending series
//         P(a, x) = x^a e^-x / Gamma(a+1) * sum_n x^n / ((a+1)...(a+n))
//     is evaluated for all days by Horner's rule against ONE table of coefficients per evaluation (one FMA per term and
//     day, no divisions); days whose upper tail is below 2^-54 are exactly 1.  Parameter regions where the series would
//     need more terms than the table holds fall back to the general algorithm (nmath_dev.cuh pgamma_lower).
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

constexpr int kXsLead = 84;  // 7 zero blocks x 12 doubles in front of the padded series
__device__ __forceinline__ int xs_index(int i) { return i + ((i >> 3) << 2) + kXsLead; }

// per-warp scalars (shared memory)
struct WScal {
  double node_x[kMaxK], node_y[kMaxK], sp1[kMaxK], sp2[kMaxK], sp3[kMaxK], cday[kMaxK];
  double ne[kMaxA], nm[kMaxA];   // normalised attack-rate weights, ne[a]*ma[a]
  double base[kMaxS];            // survey-window means of the un-weighted sero convolution
  double rden[kMaxK];            // 1/(node_x[i+1]-node_x[i])
  double ctab[32];               // interval-mass coefficients of the gamma table
};

struct WarpSmem {
  double* p;   // [dpad] parameter vector
  double* xs;  // [12*(8NT+7)] padded series (infection curve / seroconversion density); also the series table of S2
  double* gs;  // [64NT+16]    7 zeros + kernel (gamma cdf differences / Weibull survival)
  double* sk;  // [64NT+2]     cumulative seroconversion hazard prefix sums CH[k]
  WScal* sc;
};

__host__ __device__ inline int warp_tiles(int T) { return (T + 63) / 64; }
__host__ __device__ inline size_t warp_smem_doubles(int d, int NT) {
  return (size_t)((d + 3) & ~3) + 12 * (8 * NT + 7) + (64 * NT + 16) + (64 * NT + 4) + (sizeof(WScal) + 7) / 8;
}
__device__ inline WarpSmem carve_warp(double* q, int d, int NT) {
  WarpSmem w;
  w.p = q; q += (d + 3) & ~3;
  w.xs = q; q += 12 * (8 * NT + 7);
  w.gs = q; q += 64 * NT + 16;
  w.sk = q; q += 64 * NT + 4;
  w.sc = reinterpret_cast<WScal*>(q);
  return w;
}

// zero the regions that are only ever read (leading zero blocks / negative lags); once per warp
__device__ inline void warp_smem_init(const WarpSmem& w) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < kXsLead; i += 32) w.xs[i] = 0.0;
  if (lane < 7) w.gs[lane] = 0.0;
  __syncwarp();
}

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
#pragma unroll
  for (int D = 0; D < NT; D++) {
#pragma unroll 2
    for (int e = 0; e < 8; e++) {
      const int dl = 8 * D + e;
      if (dl >= nb) break;
      const double a0 = ap[8 * dl], a1 = ap[8 * dl - 4];
      const double* b = bp - 12 * dl;
      double b0[NT], b1[NT];
#pragma unroll
      for (int P = D; P < NT; P++) {
        b0[P] = b[96 * P];
        b1[P] = b[96 * P + 4];
      }
#pragma unroll
      for (int P = D; P < NT; P++) {
        dmma884(acc[P][0], acc[P][1], a0, b0[P]);
        dmma884(acc1[P][0], acc1[P][1], a1, b1[P]);
      }
    }
  }
#pragma unroll
  for (int P = 0; P < NT; P++) {
    acc[P][0] += acc1[P][0];
    acc[P][1] += acc1[P][1];
  }
}

// ------------------------------------------------------------------------------------------------ gamma cdf table
// gs[7 + k] <- pgamma(k+1; shape, scale) - pgamma(k; shape, scale), k = 0..T-1, and zeros up to the end of the region.
//
// Fast evaluation (all T+1 values share the shape a = 1/sod^2 and sit on the grid x_k = k h, h = 1/scale):
//   * k <= km (km ~ the median, >= 16): lower tail by the ascending series
//         P(a, x) = x^a e^-x / Gamma(a+1) * sum_n x^n / ((a+1)...(a+n))            (relative accuracy)
//   * k >  km: the interval masses are integrated exactly term by term,
//         P(x_{k+1}) - P(x_k) = f(x_k) * sum_m binom(a-1, m) gamma(m+1, h) x_k^-m,   f = gamma density,
//     (Taylor expansion of (1 + tau/x_k)^(a-1) on [0, h]; the expansion parameter is 1/k) against ONE 32-entry table per
//     evaluation, which gives every mass to relative accuracy even where P -> 1; the upper tail Q(x_k) is their suffix
//     sum on top of Q(x_T) (Legendre continued fraction, one point).
//   * finally P(x_k) = 1 - Q(x_k) is ROUNDED like the reference's lower-tail value and differenced, so that the
//     far-tail masses carry the same 2^-53 rounding pattern as differences of R::pgamma values (they are reproduced
//     bit for bit on ~90 % of the days and to < 1e-11 relative on the rest; tests/test_gpu_parity.py).
// Shapes/scales outside the validated box of this scheme (h > 3.6, x_T < 3, table not converged, non-finite input) take
// the general algorithm of nmath_dev.cuh (R's pgamma_raw region split) - a warp-uniform decision.
template <int NT>
__device__ inline void gamma_delay_kernel(const KModel& m, const WarpSmem& w, double mod, double sod) {
  const int lane = threadIdx.x & 31, T = m.T;
  const double alph = 1.0 / (sod * sod), scale = mod * (sod * sod);
  double* P = w.gs + 7;          // P[k], k = 0..T
  double* tab = w.xs + kXsLead;  // series coefficients 1/(a+n) (the series region is free at this point)
  double* ctab = w.sc->ctab;     // interval-mass coefficients
  constexpr int kTabCap = 12 * 8 * NT - 4;
  const double kLogTol = 41.5;   // terms below e^-41.5 of the leading one are dropped (< 2^-59)

  bool fast = (m.gamma_mode == 0) && (alph > 0.0) && (scale > 0.0) && isfinite(alph) && isfinite(scale);
  double h = 0.0, xT = 0.0;
  int km = 0, nser = 0;
  if (fast) {
    h = 1.0 / scale;
    xT = (double)T * h;
    if (!(h <= 3.6) || !(xT >= 3.0)) fast = false;
  }
  if (fast) {
    const double kmed = fmax(alph - 1.0 / 3.0, 0.0) * scale;   // ~ median in days
    km = (kmed >= (double)T) ? T : max((int)ceil(kmed), 16);
    km = min(km, T);
    // series length for x = km h
    const double x = (double)km * h;
    const double nstar = fmax(floor(x - alph), 0.0);
    const double y = kLogTol / x;
    const double v = y * (1.0 / 3.0) + sqrt(y * y * (1.0 / 9.0) + 2.0 * y);
    const double nn = nstar + v * x + 4.0;
    if (!(nn < (double)kTabCap)) fast = false;
    else nser = (int)nn;
  }
  if (fast && km < T) {
    // ---- interval-mass table: lane mm holds c_mm = binom(a-1, mm) * gamma(mm+1, h)
    const int mm = lane;
    // gamma(mm+1, h) = e^-h h^(mm+1) sum_j h^j / ((mm+1)...(mm+1+j))
    double term = m.rk[mm + 1], sum = term;
    int j = 1;
    while (__any_sync(0xffffffffu, term > 1e-18 * sum) && j < 120) {
      term *= h * m.rk[mm + 1 + j];
      sum += term;
      j++;
    }
    double hp = h, bn = (mm >= 1) ? (alph - (double)mm) * m.rk[mm] : 1.0;  // inclusive prefix products -> h^(mm+1), binom(a-1, mm)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t1 = __shfl_up_sync(0xffffffffu, hp, o), t2 = __shfl_up_sync(0xffffffffu, bn, o);
      if (lane >= o) { hp *= t1; bn *= t2; }
    }
    double cm = bn * (exp(-h) * hp * sum);
    if (mm == 0) cm = -expm1(-h);  // gamma(1, h) exactly
    ctab[mm] = cm;
    // convergence of the 32-term table at the first upper day: |c_31| u^31 must be negligible against c_0
    const double u1 = scale * m.rk[km + 1];
    double pw = u1;
#pragma unroll
    for (int q = 0; q < 5; q++) pw *= pw;  // u1^32 >= u1^31 when u1 <= 1; otherwise the test fails anyway
    const double c31 = __shfl_sync(0xffffffffu, cm, 31), c0 = __shfl_sync(0xffffffffu, cm, 0);
    if (!(fabs(c31) * pw <= 1e-18 * c0 * fmin(u1, 1.0)) || !(u1 < 4.0)) fast = false;
  }

  if (!fast) {
    // general algorithm (R's pgamma_raw region split), one day per lane and round
    const bool bad = isnan(scale) || isnan(alph) || !(scale > 0.0) || alph < 0.0;
    GammaShape gsh;
    if (!bad) gsh = gamma_shape(alph);
    for (int k = lane; k <= T; k += 32) {
      double v;
      if (bad) v = (isnan(scale) || isnan(alph)) ? (scale + alph) : d_nan();
      else v = pgamma_lower(gsh, (double)k / scale);
      P[k] = v;
    }
  } else {
    const double lgam_a = lgamma(alph), lsc = log(scale);
    for (int n = lane; n <= nser; n += 32) tab[n] = 1.0 / (alph + (double)n);
    __syncwarp();
    // ---- upper part: interval masses dgt[k], k = km+1 .. T-1, written to P[k]
    const double A = fmax(1.0, fabs(alph - 1.0));
    for (int k0 = km + 1; k0 < T; k0 += 32) {
      const int k = k0 + lane;
      int M = 32;
      if ((double)k0 > 1.5 * A) M = min(32, 2 + (int)(kLogTol / log((double)k0 / A)));
      const double u = scale * m.rk[min(k, T)];
      double I = ctab[M - 1];
      for (int q = M - 2; q >= 0; q--) I = fma(I, u, ctab[q]);
      if (k < T) P[k] = exp((alph - 1.0) * (m.logk[k] - lsc) - (double)k * h - lgam_a) * I;
    }
    // ---- Q(x_T): Legendre continued fraction, forward recurrence (all lanes redundantly)
    double QT = 0.0;
    if (km < T && xT < alph + 9.0 * sqrt(alph) + 45.0) {
      double Am1 = 1.0, Ac = xT + 1.0 - alph, Bm1 = 0.0, Bc = 1.0, f = Ac, prev = 0.0;
      for (int n = 1; n < 600; n++) {
        const double an = -(double)n * ((double)n - alph), bnn = xT + (double)(2 * n + 1) - alph;
        const double An = fma(bnn, Ac, an * Am1), Bn = fma(bnn, Bc, an * Bm1);
        Am1 = Ac; Ac = An; Bm1 = Bc; Bc = Bn;
        if (fabs(Ac) > 1e200) { Ac *= 1e-200; Am1 *= 1e-200; Bc *= 1e-200; Bm1 *= 1e-200; }
        if ((n & 1) == 0) {
          f = Ac / Bc;
          if (fabs(f - prev) <= 1e-16 * fabs(f)) break;
          prev = f;
        }
      }
      QT = exp(alph * (m.logk[T] - lsc) - xT - lgam_a) / f;
    }
    // ---- lower part: series, days 1..km (km <= 32 in practice; loop for generality)
    double plow[2] = {0.0, 0.0};
    {
      const double lgam_a1 = lgam_a + log(alph);
      for (int k0 = 1, q = 0; k0 <= km; k0 += 32, q++) {
        const int k = k0 + lane;
        const double x = (double)k * h;
        double ssum = 1.0;
        for (int n = nser; n >= 1; n--) ssum = fma(x * tab[n], ssum, 1.0);
        const double pv = exp(alph * (m.logk[min(k, T)] - lsc) - x - lgam_a1) * ssum;
        if (q < 2) plow[q] = pv;
        else if (k <= km) P[k] = pv;  // beyond 64 days the suffix pass below never touches these
      }
    }
    __syncwarp();
    // ---- suffix sums of the masses -> Q[k], then P[k] = 1 - Q[k] rounded like the reference
    if (km < T) {
      const int per = (T + 31) >> 5;
      const int lo = lane * per, hi = min(lo + per, T);  // this lane owns k in [lo, hi)
      double loc = 0.0;
      for (int k = hi - 1; k >= lo; k--)
        if (k > km) loc += P[k];
      double incl = loc;  // inclusive suffix over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double tnext = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += tnext;
      }
      double run = QT + (incl - loc);  // Q at the end of this lane's chunk
      for (int k = hi - 1; k >= lo; k--)
        if (k > km) {
          run += P[k];
          P[k] = 1.0 - run;
        }
      if (lane == 0) P[T] = 1.0 - QT;
    }
    __syncwarp();
    if (1 + lane <= km) P[1 + lane] = plow[0];
    if (33 + lane <= km) P[33 + lane] = plow[1];
    if (lane == 0) P[0] = 0.0;
  }
  __syncwarp();
  // ---- differences in place
  constexpr int kPer = 2 * NT;  // days per lane
  double dgv[kPer];
#pragma unroll
  for (int u = 0; u < kPer; u++) {
    const int k = lane + 32 * u;
    dgv[u] = (k < T) ? (P[k + 1] - P[k]) : 0.0;
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < kPer; u++) P[lane + 32 * u] = dgv[u];
  if (lane < 9) P[64 * NT + lane] = 0.0;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ scalars
// S0 role map + S1 weights + S3 knot check + S4 spline coefficients.  Returns false (warp-uniform) on knot disorder.
__device__ inline bool warp_setup(const KModel& m, const WarpSmem& w, double& snm_out) {
  const int lane = threadIdx.x & 31, K = m.K, A = m.A;
  const double* p = w.p;
  WScal& s = *w.sc;
  // knots / heights with the re-parameterisation lift (Agecpp2strings.R:277-355)
  if (lane < K) {
    if (lane == 0) s.node_x[0] = 1.0;
    else {
      const int r = m.idx_knots[lane - 1];
      double v = p[r];
      if (m.reparam_knots && r != m.idx_rel_knot) v = v * p[m.idx_rel_knot];
      s.node_x[lane] = v;
    }
    const int r = m.idx_infxn[lane];
    double v = p[r];
    if (m.reparam_infxn && r != m.idx_rel_infxn) v = v * p[m.idx_rel_infxn];
    s.node_y[lane] = v;
  }
  // attack-rate weights and ne*ma (binomial.cpp:60-73, 201-203)
  double snm = 0.0;
  for (int a0 = 0; a0 < A; a0 += 32) {
    const int a = a0 + lane;
    double v = 0.0, ma = 0.0;
    if (a < A) {
      const int r = m.idx_ifr[a];
      ma = p[r];
      if (m.reparam_ifr && r != m.idx_max_ma) ma = ma * p[m.idx_max_ma];
      v = (A > 1) ? p[m.idx_noise[a]] * (double)m.demog[a] : 0.0;
    }
    double nedom = warp_sum(v);  // A <= 32 always (CC_MAX_STRATA)
    double ne = (A > 1) ? v / nedom : (m.binomial ? 1.0 : 0.0);
    if (a < A) {
      s.ne[a] = ne;
      s.nm[a] = ne * ma;
      snm += ne * ma;
    }
  }
  snm_out = warp_sum(snm);
  __syncwarp();
  // knot order (binomial.cpp:88-96)
  bool ok = true;
  if (lane >= 1 && lane < K) ok = s.node_x[lane] > s.node_x[lane - 1];
  if (!__all_sync(0xffffffffu, ok)) return false;
  if (lane < K - 1) s.rden[lane] = 1.0 / (s.node_x[lane + 1] - s.node_x[lane]);
  __syncwarp();
  if (lane == 0) {
    // natural cubic spline, tridiagonal sweep (binomial.cpp:102-139); divisions by the knot spacing use its reciprocal
    double z[kMaxK], kk[kMaxK];
    z[0] = 0.0;
    kk[0] = 0.0;
    for (int i = 1; i < K - 1; i++) {
      const double mm = 3.0 * s.rden[i] * (s.node_y[i + 1] - s.node_y[i]) - 3.0 * s.rden[i - 1] * (s.node_y[i] - s.node_y[i - 1]);
      const double den_im1 = s.node_x[i] - s.node_x[i - 1];
      const double rg = 1.0 / (2.0 * (s.node_x[i + 1] - s.node_x[i - 1]) - den_im1 * kk[i - 1]);
      kk[i] = (s.node_x[i + 1] - s.node_x[i]) * rg;
      z[i] = (mm - den_im1 * z[i - 1]) * rg;
    }
    double sp2n = 0.0;  // sp2[K-1]
    for (int i = K - 2; i >= 0; i--) {
      const double den = s.node_x[i + 1] - s.node_x[i];
      const double sp2i = z[i] - kk[i] * sp2n;
      s.sp2[i] = sp2i;
      s.sp1[i] = (s.node_y[i + 1] - s.node_y[i]) * s.rden[i] - (den * (sp2n + 2.0 * sp2i)) * (1.0 / 3.0);
      s.sp3[i] = (sp2n - sp2i) * (1.0 / 3.0) * s.rden[i];
      sp2n = sp2i;
    }
    // interval-lag rule in closed form (binomial.cpp:153-159)
    double c = 0.0;
    s.cday[0] = 0.0;
    for (int mI = 1; mI <= K - 2; mI++) {
      c = fmax(c + 1.0, ceil(s.node_x[mI]) - 1.0);
      s.cday[mI] = c;
    }
  }
  __syncwarp();
  return true;
}

__device__ __forceinline__ double wspline_day(const KModel& m, const WScal& s, int i) {
  int j = 0;
  const double di = (double)i;
  for (int mI = 1; mI <= m.K - 2; mI++) j += (s.cday[mI] < di) ? 1 : 0;
  const double dx = (double)(i + 1) - s.node_x[j];
  return fma(fma(fma(s.sp3[j], dx, s.sp2[j]), dx, s.sp1[j]), dx, s.node_y[j]);
}

// ------------------------------------------------------------------------------------------------ sero hazard
// sk[k] <- CH[k] = sum_{u<k} H[u], k = 0..M, H = cumulative seroconversion hazard (binomial.cpp:257-284)
template <int NT>
__device__ inline void sero_hazard_prefix(const KModel& m, const WarpSmem& w) {
  const int lane = threadIdx.x & 31, M = m.M;
  const double rate = w.p[m.idx_rate];
  if (!m.serorev) {
    // H[u] = 1 - exp(-(u+1)/rate)  =>  CH[k] = k - q G_k / G_1,  G_k = 1 - q^k = -expm1(-k/rate), q = 1 - G_1
    const double g1 = -expm1(-1.0 / rate), g32 = -expm1(-32.0 / rate);
    const double fac = (g1 == 0.0) ? 0.0 : (1.0 - g1) / g1;
    double gk = -expm1(-(double)lane / rate);
    for (int k = lane; k <= M; k += 32) {
      w.sk[k] = (g1 == 0.0) ? 0.0 : ((double)k - fac * gk);
      gk = fma(-gk, g32, gk + g32);  // G_{k+32} = G_k + G_32 - G_k G_32
    }
    __syncwarp();
    return;
  }
  // seroreversion: H = (exponential seroconversion density) (*) (Weibull survival), a Toeplitz product on the tensor cores
  const double shape = w.p[m.idx_shape], wscale = w.p[m.idx_scale];
  const bool badw = !(shape > 0.0) || !(wscale > 0.0);  // R::pweibull -> NaN
  const double lws = log(wscale), rr = 1.0 / rate;
  for (int k = lane; k < 64 * NT; k += 32) {
    double c = 0.0, sv = 0.0;
    if (k < M) {
      c = rr * exp(-(double)(k + 1) / rate);
      sv = badw ? d_nan() : ((k == 0) ? 1.0 : exp(-exp(shape * (m.logk[k] - lws))));
    }
    w.xs[xs_index(k)] = c;
    w.gs[7 + k] = sv;
  }
  if (lane < 9) w.gs[7 + 64 * NT + lane] = 0.0;
  __syncwarp();
  double acc[NT][2];
  toeplitz_dmma<NT>(w.xs, w.gs, (M + 7) >> 3, acc);
  __syncwarp();
  double* haz = w.gs + 8;  // reuse (the kernel region is free again)
  {
    const int n = lane >> 2, t = lane & 3;
#pragma unroll
    for (int P = 0; P < NT; P++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int j = 64 * P + 16 * t + 8 * e + n;
        haz[j] = (j < M) ? acc[P][e] : 0.0;
      }
  }
  __syncwarp();
  // exclusive prefix sums: contiguous chunk per lane + warp scan
  const int per = (M + 31) >> 5;
  const int lo = min(lane * per, M), hi = min(lo + per, M);
  double loc = 0.0;
  for (int u = lo; u < hi; u++) loc += haz[u];
  double incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double tprev = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += tprev;
  }
  double run = incl - loc;
  for (int u = lo; u < hi; u++) {
    w.sk[u] = run;
    run += haz[u];
  }
  if (hi == M && lo < M) w.sk[M] = run;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ the evaluation
// All 32 lanes call this with the parameter vector in w.p (and a __syncwarp() behind it).  Result in every lane.
template <int NT>
__device__ inline double warp_loglike(const KModel& m, const WarpSmem& w) {
  const int lane = threadIdx.x & 31;
  const int T = m.T, M = m.M, A = m.A, S = m.S;
  WScal& sc = *w.sc;
  double snm;
  if (!warp_setup(m, w, snm)) return -CC_OVERFLO;

  // ---- S10: cumulative hazard prefix sums (uses the series/kernel regions when seroreversion is on, so it goes first)
  sero_hazard_prefix<NT>(m, w);

  // ---- S2: onset-to-death kernel
  gamma_delay_kernel<NT>(m, w, w.p[m.idx_mod], w.p[m.idx_sod]);

  // ---- S5: infection curve, S6: population check
  double tot = 0.0;
  for (int i = lane; i < 64 * NT; i += 32) {
    double v = 0.0;
    if (i < T) {
      const double y = (i == 0) ? sc.node_y[0] : wspline_day(m, sc, i);
      v = exp(y);
      tot += v;
    }
    w.xs[xs_index(i)] = v;
  }
  tot = warp_sum(tot);
  {
    bool ok = true;
    for (int a = lane; a < A; a += 32) ok = ok && !(sc.ne[a] * tot > (double)m.demog[a]);
    if (!__all_sync(0xffffffffu, ok)) return -CC_OVERFLO;
  }
  __syncwarp();

  // ---- S7: death-delay convolution on the tensor cores
  double auc[NT][2];
  toeplitz_dmma<NT>(w.xs, w.gs, (T + 7) >> 3, auc);

  // ---- S8: Poisson term straight from the accumulator fragments
  double l1 = 0.0, s_all = 0.0, s_unc = 0.0;
  {
    const int n = lane >> 2, t = lane & 3;
#pragma unroll
    for (int P = 0; P < NT; P++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int j = 64 * P + 16 * t + 8 * e + n;
        if (j < T) {
          const double a = auc[P][e];
          const double expd = a * snm;
          s_all += expd;
          if ((j + 1) < m.rcensor_day) {
            s_unc += a;
            const int x = m.obsd[j];
            if (x != -1) {
              double lp;
              if (!(expd >= 0.0)) lp = d_nan();            // NaN or negative mean -> NaN (R::dpois)
              else if (x < 0) lp = -d_inf();
              else if (x == 0) lp = -expd;
              else lp = (double)x * log(expd) - expd;      // - lgamma(x+1) is data only: m.l1_const
              l1 += lp;
            }
          }
        }
      }
  }
  l1 = warp_sum(l1) - m.l1_const;
  const double texpd = warp_sum(s_all);
  const double sum_unc = warp_sum(s_unc);

  // ---- S11/S12: survey-window means of infxn (*) H through the prefix-summed hazard
  for (int s = 0; s < S; s++) {
    const int st = m.sero_start[s], en = m.sero_end[s];
    double b = 0.0;
    for (int t = lane; t < en; t += 32) {
      int lo = st - 1 - t;
      lo = lo < 0 ? 0 : lo;
      b = fma(w.xs[xs_index(t)], w.sk[en - t] - w.sk[lo], b);
    }
    b = warp_sum(b);
    if (lane == 0) sc.base[s] = b / (double)(en - st + 1);
  }
  __syncwarp();

  // ---- S9: chained binomial on the age split, S13: serology
  double tail = 0.0;
  for (int a = lane; a < A - 1; a += 32) {
    double te = texpd;
    for (int b = 0; b < a; b++) te -= sc.nm[b] * sum_unc;
    tail += dbinom_log(round(sc.nm[a] * sum_unc), round(te), m.l2_prob[a]);
  }
  const double sens = w.p[m.idx_sens], spec = w.p[m.idx_spec];
  for (int k = lane; k < S * A; k += 32) {
    const int s = k / A, a = k - s * A;
    const double frac = (sc.ne[a] * sc.base[s]) / (double)m.demog[a];
    const double prev = sens * frac + (1 - spec) * (1 - frac);
    if (m.binomial) {
      const int kind = m.l3_kind[k];
      if (kind == 1) {
        const double x = m.sero_a[k], nn = m.sero_b[k];
        double v;
        if (!(prev >= 0.0 && prev <= 1.0)) v = d_nan();
        else if (x == 0.0) v = (nn == 0.0) ? 0.0 : nn * log(1 - prev);
        else if (x == nn) v = nn * log(prev);
        else v = m.l3_lchoose[k] + x * log(prev) + (nn - x) * log(1 - prev);
        tail += v;
      } else if (kind == 2) {
        tail += dbinom_log(m.sero_a[k], m.sero_b[k], prev);  // irregular data (non-integer, x > n, ...): R's own case analysis
      }
    } else {
      const double eps = DBL_MIN / 100.0;
      const double lg = log((prev + eps) / (1 - (prev + eps)));
      tail += dnorm_log(lg, m.sero_a[k], m.sero_b[k]);
    }
  }
  tail = warp_sum(tail);
  double loglik = l1 + tail;
  if (!isfinite(loglik)) loglik = -CC_OVERFLO;
  return loglik;
}

// generated logprior on one warp: one term per lane, warp reduction, clamp (Agecpp2strings.R:66-219)
__device__ inline double warp_logprior(const KModel& m, const double* p) {
  const int lane = threadIdx.x & 31;
  double v = 0.0;
  for (int i = lane; i < m.d; i += 32) {
    const int fam = m.prior_family[i];
    if (fam != CC_PRIOR_NONE) {
      const double x = p[i], a = m.prior_p1[i], b = m.prior_p2[i];
      v += (fam == CC_PRIOR_DUNIF) ? dunif_log(x, a, b) : (fam == CC_PRIOR_DNORM) ? dnorm_log(x, a, b) : dbeta_log(x, a, b);
    }
  }
  if (lane < 3 && m.jac_rows[lane] >= 0) v += m.jac_counts[lane] * log(p[m.jac_rows[lane]]);
  v = warp_sum(v);
  if (!isfinite(v)) v = -CC_OVERFLO;
  return v;
}

}  // namespace ccdev
