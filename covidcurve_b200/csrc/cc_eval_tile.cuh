// cc_eval_tile.cuh - tile-phased evaluation: one warp owns a tile of E <= 32 parameter vectors and alternates between
//   phase 1  (one LANE per vector)   everything that is scalar per vector: role map, attack-rate weights, knot check, spline
//                                    coefficients, gamma-table constants (lgamma, logs, continued fraction of the tail), sero
//                                    constants  -> a record per vector in an L2-resident scratch (structure-of-arrays per tile)
//   phase 2  (the WARP per vector)   everything that is per day: delay-kernel masses, spline + exp, the two Toeplitz products on
//                                    the FP64 tensor cores, the Poisson term, the survey-window sums
//   phase 3  (one LANE per vector)   everything that is per stratum / survey: chained binomial, serology term, total + clamp.
// Serial scalar code therefore runs 32 vectors wide instead of on one lane of a warp, every transcendental call site exists
// once (the instruction footprint of the per-day loop is a few thousand instructions), and global loads/stores of theta/out
// are perfectly coalesced (lane <-> consecutive vector).  For small batches (the Metropolis step) the tile is narrowed to
// E = 1, 2, 4 ... vectors so that all warps of the GPU stay busy; phases 1 and 3 then run E lanes wide.
//
// Stage map and reference line numbers: see cc_eval_warp.cuh / SURVEY.md 3.2.
#pragma once
#include "cc_eval_warp.cuh"
#ifndef CC_P2_EXPERIMENT
#define CC_P2_EXPERIMENT 0
#endif

namespace ccdev {

// ---- record layout (doubles); phase-2 fields first
struct RecLayout {
  int flags, gam, sero, snm, ne, spl, cday, n2;  // [0, n2) is what phase 2 loads
  int nm, sens, spec, o_status, o_l1, o_texpd, o_sunc, o_base, F;
};
__host__ __device__ inline RecLayout rec_layout(int K, int A, int S) {
  RecLayout L;
  L.flags = 0;            // 1 = knots ordered
  L.gam = 1;              // alph, scale, h, lgam_a, lsc, km, nser, QT, fast, lgam_a1, logA, exp(-h), k_hi, sigma_top
  L.sero = 15;            // rate, g1, g32, fac | rate, shape, lws, badw
  L.snm = 19;
  L.ne = 20;              // [A]
  L.spl = L.ne + A;       // per interval j < K-1: x_j, y_j, sp1_j, sp2_j, sp3_j
  L.cday = L.spl + 5 * (K - 1);  // cday[1..K-2]
  L.n2 = L.cday + (K - 2);
  L.nm = L.n2;            // [A]
  L.sens = L.nm + A;
  L.spec = L.sens + 1;
  L.o_status = L.spec + 1;  // 0 ok, 1 population check failed
  L.o_l1 = L.o_status + 1;
  L.o_texpd = L.o_l1 + 1;
  L.o_sunc = L.o_texpd + 1;
  L.o_base = L.o_sunc + 1;  // [S]
  L.F = L.o_base + S;
  return L;
}

// read-only per-day tables.  Staging them in shared memory per CTA was measured SLOWER than leaving them in global memory
// (they are a few kB and stay L1 resident; the extra 6 kB of shared memory per CTA shrinks the L1): 58 vs 65 M evals/s.
struct TileTables {
  const double* logk;  // [T+1]
  const double* rk;    // [nrk]
  const int* obsd;     // [T]
  int nrk;
};
__host__ __device__ inline size_t tile_tables_doubles(int) { return 0; }
__device__ inline TileTables tile_tables_load(const KModel& m, double*) {
  TileTables t;
  t.logk = m.logk; t.rk = m.rk; t.obsd = m.obsd; t.nrk = 0x7fffffff;
  return t;
}

struct TileSmem {
  double* rec;  // [n2 rounded] record of the vector being processed
  double* xs;   // padded series
  double* gs;   // 7 zeros + delay kernel
  double* auc;  // convolution output in day order (the batch kernels alias it onto sk: the survey sums run before the convolution)
  double* sk;   // CH prefix sums
  double* ctab; // [128] interval-mass coefficients
};
__host__ __device__ inline size_t tile_smem_doubles(int K, int A, int S, int NT) {
  const RecLayout L = rec_layout(K, A, S);
  return (size_t)((L.n2 + 3) & ~3) + 12 * (8 * NT + 7) + (64 * NT + 16) + (64 * NT + 4) + 128;
}
__device__ inline TileSmem carve_tile(double* q, const RecLayout& L, int NT) {
  TileSmem w;
  w.rec = q; q += (L.n2 + 3) & ~3;
  w.xs = q; q += 12 * (8 * NT + 7);
  w.gs = q; q += 64 * NT + 16;
  w.auc = q;  // = sk
  w.sk = q; q += 64 * NT + 4;
  w.ctab = q;
  return w;
}
__device__ inline void tile_smem_init(const TileSmem& w) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < kXsLead; i += 32) w.xs[i] = 0.0;
  if (lane < 7) w.gs[lane] = 0.0;
  __syncwarp();
}

// single out-of-line copies of the heavy scalar routines (instruction footprint)
__device__ __noinline__ double dbinom_log_shared(double x, double n, double p) { return dbinom_log(x, n, p); }
__device__ __forceinline__ double lgamma_shared(double x) { return lgamma(x); }
// the general incomplete-gamma algorithm is ~8 700 instructions and runs only for shapes outside the validated box
__device__ __noinline__ double pgamma_lower_shared(const GammaShape& g, double x) { return pgamma_lower(g, x); }

// ------------------------------------------------------------------------------------------------ phase 1
// Scalar-per-vector pieces.  `par(i)` returns parameter i of the vector, `put(f, v)` stores record field f.
// The pieces are independent so that the Metropolis step can redo only what a single-parameter proposal touches.

// IFR of stratum a with the re-parameterisation lift (the value p1_weights multiplies into ne*ma)
template <typename Par>
__device__ __forceinline__ double ifr_of(const KModel& m, Par par, int a) {
  const int r = m.idx_ifr[a];
  double ma = par(r);
  if (m.reparam_ifr && r != m.idx_max_ma) ma = ma * par(m.idx_max_ma);
  return ma;
}

// attack-rate weights (binomial.cpp:60-73), ne*ma, test characteristics
template <typename Par, typename Put>
__device__ inline void p1_weights(const KModel& m, const RecLayout& L, Par par, Put put) {
  const int A = m.A;
  double snm = 0.0, nedom = 0.0;
  if (A > 1)
    for (int a = 0; a < A; a++) nedom += par(m.idx_noise[a]) * (double)m.demog[a];
  for (int a = 0; a < A; a++) {
    const int r = m.idx_ifr[a];
    double ma = par(r);
    if (m.reparam_ifr && r != m.idx_max_ma) ma = ma * par(m.idx_max_ma);
    const double ne = (A > 1) ? (par(m.idx_noise[a]) * (double)m.demog[a]) / nedom : (m.binomial ? 1.0 : 0.0);
    put(L.ne + a, ne);
    put(L.nm + a, ne * ma);
    snm += ne * ma;
  }
  put(L.snm, snm);
  put(L.sens, par(m.idx_sens));
  put(L.spec, par(m.idx_spec));
}

// the same, by a whole warp on a record in shared memory: one stratum per lane for the divisions, the two sums in the
// serial order of the scalar version (identical values)
// `keep_ne`: rec[L.ne + a] already holds the weights of this parameter vector (the proposal moved none of the noise scalars)
template <typename Par>
__device__ inline void p1_weights_warp(const KModel& m, const RecLayout& L, Par par, double* rec, bool keep_ne = false) {
  const int A = m.A, lane = threadIdx.x & 31;
  double nedom = 0.0;
  if (A > 1 && !keep_ne)
    for (int a = 0; a < A; a++) nedom += par(m.idx_noise[a]) * (double)m.demog[a];
  for (int a = lane; a < A; a += 32) {
    const int r = m.idx_ifr[a];
    double ma = par(r);
    if (m.reparam_ifr && r != m.idx_max_ma) ma = ma * par(m.idx_max_ma);
    const double ne = keep_ne ? rec[L.ne + a] : (A > 1) ? (par(m.idx_noise[a]) * (double)m.demog[a]) / nedom : (m.binomial ? 1.0 : 0.0);
    rec[L.ne + a] = ne;
    rec[L.nm + a] = ne * ma;
  }
  __syncwarp();
  if (lane == 0) {
    double snm = 0.0;
    for (int a = 0; a < A; a++) snm += rec[L.nm + a];
    rec[L.snm] = snm;
    rec[L.sens] = par(m.idx_sens);
    rec[L.spec] = par(m.idx_spec);
  }
  __syncwarp();
}

// role map with the re-parameterisation lift (Agecpp2strings.R:277-355), knot order (binomial.cpp:88-96), natural cubic
// spline by the tridiagonal sweep (binomial.cpp:102-139), interval-lag thresholds (binomial.cpp:153-159)
template <typename Par, typename Put>
__device__ inline bool p1_spline(const KModel& m, const RecLayout& L, Par par, Put put) {
  const int K = m.K;
  double nx[kMaxK], ny[kMaxK];
  nx[0] = 1.0;
  for (int i = 0; i < K - 1; i++) {
    const int r = m.idx_knots[i];
    double v = par(r);
    if (m.reparam_knots && r != m.idx_rel_knot) v = v * par(m.idx_rel_knot);
    nx[i + 1] = v;
  }
  for (int i = 0; i < K; i++) {
    const int r = m.idx_infxn[i];
    double v = par(r);
    if (m.reparam_infxn && r != m.idx_rel_infxn) v = v * par(m.idx_rel_infxn);
    ny[i] = v;
  }
  bool ok = true;
  for (int i = 1; i < K; i++) ok = ok && (nx[i] > nx[i - 1]);
  put(L.flags, ok ? 1.0 : 0.0);
  if (!ok) return false;
  double z[kMaxK], kk[kMaxK], rden[kMaxK];
  for (int i = 0; i < K - 1; i++) rden[i] = 1.0 / (nx[i + 1] - nx[i]);
  z[0] = 0.0;
  kk[0] = 0.0;
  for (int i = 1; i < K - 1; i++) {
    const double mm = 3.0 * rden[i] * (ny[i + 1] - ny[i]) - 3.0 * rden[i - 1] * (ny[i] - ny[i - 1]);
    const double den_im1 = nx[i] - nx[i - 1];
    const double rg = 1.0 / (2.0 * (nx[i + 1] - nx[i - 1]) - den_im1 * kk[i - 1]);
    kk[i] = (nx[i + 1] - nx[i]) * rg;
    z[i] = (mm - den_im1 * z[i - 1]) * rg;
  }
  double sp2n = 0.0;
  for (int i = K - 2; i >= 0; i--) {
    const double den = nx[i + 1] - nx[i];
    const double sp2i = z[i] - kk[i] * sp2n;
    const int o = L.spl + 5 * i;
    put(o, nx[i]);
    put(o + 1, ny[i]);
    put(o + 2, (ny[i + 1] - ny[i]) * rden[i] - (den * (sp2n + 2.0 * sp2i)) * (1.0 / 3.0));
    put(o + 3, sp2i);
    put(o + 4, (sp2n - sp2i) * (1.0 / 3.0) * rden[i]);
    sp2n = sp2i;
  }
  double c = 0.0;
  for (int mI = 1; mI <= K - 2; mI++) {
    c = fmax(c + 1.0, ceil(nx[mI]) - 1.0);
    put(L.cday + mI - 1, c);
  }
  return true;
}

// number of interval-mass coefficients: the scaled coefficients c'_m = c_m / x_{km+1}^m decay like P(m+1, h)
__device__ __forceinline__ int gamma_ncoef(double h) { return (h <= 3.6) ? 32 : (h <= 16.0) ? 64 : (h <= 32.0) ? 96 : 128; }

// delay-kernel constants (scheme: tile_gamma below)
template <int NT, typename Par, typename Put>
__device__ inline void p1_gamma(const KModel& m, const RecLayout& L, Par par, Put put) {
  const int T = m.T;
  const double sod = par(m.idx_sod), mod = par(m.idx_mod);
  const double alph = 1.0 / (sod * sod), scale = mod * (sod * sod);
  constexpr int kTabCap = 12 * 8 * NT - 4;
  bool fast = (m.gamma_mode == 0) && (alph > 0.0) && (scale > 0.0) && isfinite(alph) && isfinite(scale);
  double h = 0.0, xT = 0.0, lgam_a = 0.0, lsc = 0.0, QT = 0.0;
  int km = 0, nser = 0;
  if (fast) {
    h = 1.0 / scale;
    xT = (double)T * h;
    if (!(h <= 48.0) || !(xT >= 3.0)) fast = false;
  }
  if (fast) {
    const double kmed = fmax(alph - 1.0 / 3.0, 0.0) * scale;
    km = (kmed >= (double)T) ? T : max((int)ceil(kmed), 16);
    km = min(km, T);
    const double x = (double)km * h;
    const double nstar = fmax(floor(x - alph), 0.0);
    const double y = 41.5 / x;
    const double v = y * (1.0 / 3.0) + sqrt(y * y * (1.0 / 9.0) + 2.0 * y);
    const double nn = nstar + v * x + 4.0;
    if (!(nn < (double)kTabCap)) fast = false;
    else nser = (int)nn;
  }
  if (fast) {
    lgam_a = lgamma_shared(alph);
    lsc = log(scale);
    if (km < T && xT < alph + 9.0 * sqrt(alph) + 45.0) {
      // Q(x_T): Legendre continued fraction, forward recurrence
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
  }
  put(L.gam + 0, alph); put(L.gam + 1, scale); put(L.gam + 2, h); put(L.gam + 3, lgam_a); put(L.gam + 4, lsc);
  put(L.gam + 5, (double)km); put(L.gam + 6, (double)nser); put(L.gam + 7, QT); put(L.gam + 8, fast ? 1.0 : 0.0);
  // per-vector scalars of the per-day pass, computed once here instead of by every lane of the warp there
  double lgam_a1 = 0.0, logA = 0.0, emh = 0.0, k_hi = d_inf();
  if (fast) {
    lgam_a1 = lgam_a + log(alph);
    logA = log(fmax(1.0, fabs(alph - 1.0)));
    emh = exp(-h);
  }
  // Beyond x = alph + 9 sqrt(alph) + 45 the upper tail is below 2^-54 (Chernoff: exp(-(t - alph log(1 + t / alph))) <= e^-40.5
  // for every shape), so the lower-tail value rounds to exactly 1 - in R's pgamma as here - and the masses out there are exactly
  // zero.  Holds for every valid shape, so the general algorithm skips those days too (they were most of its 301 evaluations for
  // the sharply peaked delays that need it).
  if ((alph > 0.0) && (scale > 0.0) && isfinite(alph) && isfinite(scale)) k_hi = (alph + 9.0 * sqrt(alph) + 45.0) * scale;
  put(L.gam + 9, lgam_a1); put(L.gam + 10, logA); put(L.gam + 11, emh); put(L.gam + 12, k_hi);
  // top of the sigma recurrence of the interval-mass coefficients (tile_gamma): ONE short series at m = nct (ratio
  // h / (nct + j) <= 3/8) - serial work, so it belongs to the thread-per-vector pass, not to every lane of a warp
  double sig_top = 0.0;
  if (fast && km < T) {
    const int nct = gamma_ncoef(h);
    double term = m.rk[nct + 1], sum = term;
    for (int j = 1; j < 200 && term > 1e-18 * sum; j++) {
      term *= h * m.rk[nct + 1 + j];
      sum += term;
    }
    sig_top = sum;
  }
  put(L.gam + 13, sig_top);
}

// sero constants (binomial.cpp:257-284)
template <typename Par, typename Put>
__device__ inline void p1_sero(const KModel& m, const RecLayout& L, Par par, Put put) {
  const double rate = par(m.idx_rate);
  put(L.sero, rate);
  if (!m.serorev) {
    const double g1 = -expm1(-1.0 / rate), g32 = -expm1(-32.0 / rate);
    put(L.sero + 1, g1);
    put(L.sero + 2, g32);
    put(L.sero + 3, (g1 == 0.0) ? 0.0 : (1.0 - g1) / g1);
  } else {
    const double shape = par(m.idx_shape), wscale = par(m.idx_scale);
    put(L.sero + 1, shape);
    put(L.sero + 2, log(wscale));
    put(L.sero + 3, (!(shape > 0.0) || !(wscale > 0.0)) ? 1.0 : 0.0);
  }
}

// all of phase 1 for lane e of a tile; scr = tile scratch, field f of lane e at scr[f*32 + e]
template <int NT, typename Par>
__device__ inline void tile_phase1(const KModel& m, const RecLayout& L, double* __restrict__ scr, int e, Par par) {
  auto put = [&](int f, double v) { scr[f * 32 + e] = v; };
  p1_weights(m, L, par, put);
  if (!p1_spline(m, L, par, put)) return;
  p1_gamma<NT>(m, L, par, put);
  p1_sero(m, L, par, put);
}

// ------------------------------------------------------------------------------------------------ phase 2 pieces
// delay kernel: w.gs[7 + k] <- P(x_{k+1}) - P(x_k)
template <int NT>
__device__ inline void tile_gamma(const KModel& m, const RecLayout& L, const TileSmem& w, const TileTables& tb) {
  const int lane = threadIdx.x & 31, T = m.T;
  const double alph = w.rec[L.gam + 0], scale = w.rec[L.gam + 1], h = w.rec[L.gam + 2], lgam_a = w.rec[L.gam + 3],
               lsc = w.rec[L.gam + 4], QT = w.rec[L.gam + 7];
  const int km = (int)w.rec[L.gam + 5], nser = (int)w.rec[L.gam + 6];
  bool fast = w.rec[L.gam + 8] != 0.0;
  double* P = w.gs + 7;
  double* tab = w.xs + kXsLead;
  double* ctab = w.ctab;
  const double kLogTol = 36.5;  // interval-mass series truncated below e^-36.5 = 1.4e-16 of the leading term

  const int nct = gamma_ncoef(h);
  if (fast && km < T) {
    // c'_m = binom(a-1, m) (km+1)^-m * h e^-h sigma_m,  sigma_m = sum_j h^j / ((m+1)...(m+1+j));  lane holds m = lane + 32 g.
    // sigma obeys sigma_m = (1 + h sigma_{m+1}) / (m+1): ONE short series at m = nct (p1_gamma: record field gam + 13), then the
    // recurrence downwards as a warp scan over affine maps, 32 values per round (errors shrink going down).  A series per
    // lane - up to 100 rounds of a dependent multiply-add for every group at large h - was 8 % of the box workload.
    const double emh = w.rec[L.gam + 11], rk1 = tb.rk[km + 1];
    const int ng = nct >> 5;
    double sig_next = w.rec[L.gam + 13];
    double sig[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int g = 3; g >= 0; g--) {
      if (g < ng) {
        const int mm = lane + 32 * g;
        double a = tb.rk[mm + 1], b = h * a;  // sigma_m = a + b sigma_{m+1}
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double a2 = __shfl_down_sync(0xffffffffu, a, o), b2 = __shfl_down_sync(0xffffffffu, b, o);
          if (lane + o < 32) {
            a = fma(b, a2, a);
            b *= b2;
          }
        }
        sig[g] = fma(b, sig_next, a);
        sig_next = __shfl_sync(0xffffffffu, sig[g], 0);
      }
    }
    double carry = 1.0, c0 = 0.0, clast = 0.0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      if (g < ng) {
        const int mm = lane + 32 * g;
        double bn = (mm >= 1) ? (alph - (double)mm) * tb.rk[mm] * rk1 : 1.0;  // inclusive prefix product -> binom(a-1, m) (km+1)^-m
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t1 = __shfl_up_sync(0xffffffffu, bn, o);
          if (lane >= o) bn *= t1;
        }
        bn *= carry;
        carry = __shfl_sync(0xffffffffu, bn, 31);
        const double cm = bn * (h * emh * sig[g]);
        ctab[mm] = cm;
        if (g == 0) c0 = __shfl_sync(0xffffffffu, cm, 0);
        clast = __shfl_sync(0xffffffffu, cm, 31);
      }
    }
    if (!(fabs(clast) <= 1e-18 * c0)) fast = false;  // table not converged at the first upper day
  }
  if (!fast) {
    // general algorithm (R's pgamma_raw region split), one day per lane and round
    const bool bad = isnan(scale) || isnan(alph) || !(scale > 0.0) || alph < 0.0;
    GammaShape gsh;
    if (!bad) gsh = gamma_shape(alph);
    const double k_hi = w.rec[L.gam + 12];
    const int kone = (bad || k_hi >= (double)T) ? T + 1 : min(T + 1, (int)k_hi + 2);  // first day whose value is exactly 1
    for (int k = lane; k <= T; k += 32) {
      double v;
      if (bad) v = (isnan(scale) || isnan(alph)) ? (scale + alph) : d_nan();
      else if (k >= kone) v = 1.0;
      else v = pgamma_lower_shared(gsh, (double)k / scale);
      P[k] = v;
    }
  } else {
    for (int n = lane; n <= nser; n += 32) tab[n] = 1.0 / (alph + (double)n);
    // upper part: interval masses for k = km+1 .. T-1, written to P[k]; two days per lane and round.
    // I(x_k) = sum_m c'_m v^m with v = (km+1)/k; terms beyond M are below e^-41.5 of the leading one ((A/k)^m decay)
    const double A = fmax(1.0, fabs(alph - 1.0));
    const double k_hi = w.rec[L.gam + 12];                                    // masses beyond are < 2^-54 of the total: 0
    const int kend = (k_hi >= (double)T) ? T : min(T, (int)k_hi + 2);
    const double kp1 = (double)(km + 1), logA = w.rec[L.gam + 10];
    int Mr = nct;  // series length of round `lane` (64 days per round)
    {
      const int k0 = km + 1 + 64 * lane;
      if (k0 < T && (double)k0 > 1.5 * A) Mr = min(nct, 2 + (int)(kLogTol / (tb.logk[k0] - logA)));
    }
    __syncwarp();
    for (int k0 = km + 1; k0 < T; k0 += 64) {
      const int ka = k0 + lane, kb = ka + 32;
      if (k0 >= kend) {
        if (ka < T) P[ka] = 0.0;
        if (kb < T) P[kb] = 0.0;
        continue;
      }
      const int M = __shfl_sync(0xffffffffu, Mr, (k0 - km - 1) >> 6);
      const int kac = min(ka, T), kbc = min(kb, T);
      const double va = kp1 * tb.rk[kac], vb = kp1 * tb.rk[kbc];
      double Ia = ctab[M - 1], Ib = Ia;
      for (int q = M - 2; q >= 0; q--) {
        const double cq = ctab[q];
        Ia = fma(Ia, va, cq);
        Ib = fma(Ib, vb, cq);
      }
      const double am1 = alph - 1.0;
      double fa, fb;
      cc_exp2(am1 * (tb.logk[kac] - lsc) - (double)ka * h - lgam_a, am1 * (tb.logk[kbc] - lsc) - (double)kb * h - lgam_a, fa, fb);
      if (ka < T) P[ka] = fa * Ia;
      if (kb < T) P[kb] = fb * Ib;
    }
    // lower part: series, days 1..km
    const double lgam_a1 = w.rec[L.gam + 9];
    for (int k0 = 1; k0 <= km; k0 += 32) {
      const int k = k0 + lane;
      const double x = (double)k * h;
      double ssum = 1.0;
      for (int n = nser; n >= 1; n--) ssum = fma(x * tab[n], ssum, 1.0);
      const double lx = tb.logk[min(k, T)] - lsc;
      const double E = alph * lx - x - lgam_a1;
      double pv;
      if (E > -690.0 || !(alph > 1.0)) {
        pv = cc_exp(E) * ssum;
      } else {
        // Below DBL_MIN / DBL_EPSILON R's pgamma_raw recomputes the value in log space and applies ONE exp, so a subnormal
        // result is rounded once, from the exact value: do the same (log P = E + log(series))
        pv = exp_gradual(E + log(ssum));
      }
      if (k <= km) P[k] = pv;
    }
    __syncwarp();
    // suffix sums of the masses -> Q[k]; P[k] = 1 - Q[k] rounded like the reference's lower-tail value
    if (km < T) {
      const int per = (T + 31) >> 5;
      const int lo = lane * per, hi = min(lo + per, T);
      double loc = 0.0;
      for (int k = hi - 1; k >= lo; k--)
        if (k > km) loc += P[k];
      double incl = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double tnext = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += tnext;
      }
      // masses of all later lanes = the next lane's inclusive sum.  (NOT incl - loc: where this lane holds the bulk of the
      // distribution and the later lanes only the far tail, that difference cancels - it cost six digits of the upper tail
      // at the lane boundary and, through the rounding of 1 - Q, one ulp of the reference's lower-tail value.)
      double excl = __shfl_down_sync(0xffffffffu, incl, 1);
      if (lane == 31) excl = 0.0;
      double run = QT + excl;
      for (int k = hi - 1; k >= lo; k--)
        if (k > km) {
          run += P[k];
#ifdef CC_DEBUG_RAWQ  // developer build: the unrounded upper tail of vector 0 instead of the hazard prefix sums
          if (m.debug_dump && blockIdx.x == 0 && (threadIdx.x >> 5) == 0) m.debug_dump[3 * T + k] = run;
#endif
          P[k] = 1.0 - run;
        }
      if (lane == 0) P[T] = 1.0 - QT;
    }
    if (lane == 0) P[0] = 0.0;
  }
  __syncwarp();
  // differences in place
  constexpr int kPer = 2 * NT;
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

// w.sk[k] <- CH[k] = sum_{u<k} H[u], k = 0..M  (binomial.cpp:257-284)
template <int NT>
__device__ inline void tile_sero_prefix(const KModel& m, const RecLayout& L, const TileSmem& w, const TileTables& tb) {
  const int lane = threadIdx.x & 31, M = m.M;
  const double rate = w.rec[L.sero];
  if (!m.serorev) {
    const double g1 = w.rec[L.sero + 1], g32 = w.rec[L.sero + 2], fac = w.rec[L.sero + 3];
    double gk = -expm1(-(double)lane / rate);
    for (int k = lane; k <= M; k += 32) {
      w.sk[k] = (g1 == 0.0) ? 0.0 : ((double)k - fac * gk);
      gk = fma(-gk, g32, gk + g32);
    }
    __syncwarp();
    return;
  }
  const double shape = w.rec[L.sero + 1], lws = w.rec[L.sero + 2];
  const bool badw = w.rec[L.sero + 3] != 0.0;
  const double rr = 1.0 / rate;
  for (int k = lane; k < 64 * NT; k += 32) {
    double c = 0.0, sv = 0.0;
    if (k < M) {
      c = rr * cc_exp(-(double)(k + 1) / rate);
      sv = badw ? d_nan() : ((k == 0) ? 1.0 : cc_exp(-cc_exp(shape * (tb.logk[k] - lws))));
    }
    w.xs[xs_index(k)] = c;
    w.gs[7 + k] = sv;
  }
  if (lane < 9) w.gs[7 + 64 * NT + lane] = 0.0;
  __syncwarp();
  double acc[NT][2];
  toeplitz_dmma<NT>(w.xs, w.gs, (M + 7) >> 3, acc);
  __syncwarp();
  double* haz = w.gs + 8;
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
  double run = __shfl_up_sync(0xffffffffu, incl, 1);  // sum of the earlier lanes, without the cancellation of incl - loc
  if (lane == 0) run = 0.0;
  for (int u = lo; u < hi; u++) {
    w.sk[u] = run;
    run += haz[u];
  }
  if (hi == M && lo < M) w.sk[M] = run;
  __syncwarp();
}

// reduce four values at once (one shuffle loop in the instruction stream)
__device__ __forceinline__ void warp_sum4(double& a, double& b, double& c, double& d) {
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
}

// knot vectors longer than K = 5: interval thresholds beyond the three kept in registers (out of line: never runs at K = 5)
__device__ __noinline__ int spline_interval_beyond3(const double* cday, int K, double di) {
  int j = 0;
  for (int mI = 3; mI < K - 2; mI++) j += (cday[mI] < di) ? 1 : 0;
  return j;
}

// ---- S5 infection curve: w.xs <- exp(spline) (zero beyond T); returns the warp-reduced total
template <int NT>
__device__ inline double stage_spline(const KModel& m, const RecLayout& L, const double* __restrict__ rec, double* __restrict__ xs) {
  const int lane = threadIdx.x & 31, T = m.T, K = m.K;
  double tot = 0.0;
  // interval thresholds in registers (the common K = 5 has three); longer knot vectors fall back to shared memory
  // (integer-valued by construction, p1_spline; compared as integers: a double compare costs a slot of the FP64 pipe)
  auto as_day = [](double c) { return (c < 2.0e9) ? (int)c : 0x7fffffff; };
  const int cd0 = (K > 2) ? as_day(rec[L.cday]) : 0x7fffffff, cd1 = (K > 3) ? as_day(rec[L.cday + 1]) : 0x7fffffff,
            cd2 = (K > 4) ? as_day(rec[L.cday + 2]) : 0x7fffffff;
  // two days per lane and round, branch-free (days beyond T are evaluated at T - 1 and zeroed): the two exp() chains are
  // adjacent in the instruction stream and the compiler interleaves them
#pragma unroll 1
  for (int i0 = lane; i0 < 64 * NT; i0 += 64) {
    double v[2], y[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = i0 + 32 * u, ic = min(i, T - 1);
      const double di = (double)ic;
      int j = ((cd0 < ic) ? 1 : 0) + ((cd1 < ic) ? 1 : 0) + ((cd2 < ic) ? 1 : 0);
      if (K > 5) j += spline_interval_beyond3(rec + L.cday, K, di);
      const double* c = rec + L.spl + 5 * j;
      const double dx = (double)(ic + 1) - c[0];
      // day 1 (ic == 0) lies in interval 0 whose left knot is exactly 1, so dx = 0 and the cubic returns its constant term
      // node_y[0] - binomial.cpp:144 - without a special case (with non-finite coefficients the evaluation fails either way)
      y[u] = fma(fma(fma(c[4], dx, c[3]), dx, c[2]), dx, c[1]);
    }
    cc_exp2(y[0], y[1], v[0], v[1]);
    v[0] = (i0 < T) ? v[0] : 0.0;
    v[1] = (i0 + 32 < T) ? v[1] : 0.0;
    tot += v[0];
    tot += v[1];
    xs[xs_index(i0)] = v[0];
    xs[xs_index(i0 + 32)] = v[1];
  }
  return warp_sum(tot);
}

// ---- S6 population check (binomial.cpp:170-184), warp-uniform result
__device__ inline bool stage_popn(const KModel& m, const RecLayout& L, const double* __restrict__ rec, double tot) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  for (int a = lane; a < m.A; a += 32) ok = ok && !(rec[L.ne + a] * tot > (double)m.demog[a]);
  return __all_sync(0xffffffffu, ok);
}

// Where a convolution output is (near) subnormal the reference's value is decided by gradual underflow: it rounds every
// product infxn[i] * dg[j - i] to the subnormal grid and adds them in the order i = 0, 1, ... (binomial.cpp:187-193), while
// the tensor-core accumulation keeps the products exact.  Redo exactly those outputs the reference's way (early days of very
// peaked delay distributions; a handful of terms each).  An exact zero needs no redo (every product is zero in both), and
// above 1e-300 a rounding of 2.5e-324 per term is below 1e-23 of the sum.  Needs the series and the kernel still in place.
__device__ __noinline__ void conv_fixup_tiny(int T, const double* __restrict__ xs, const double* __restrict__ gs, double* aucs) {
  const int lane = threadIdx.x & 31;
  for (int j = lane; j < T; j += 32) {
    const double a = aucs[j];
    if (a != 0.0 && fabs(a) < 1e-300) {
      double sref = 0.0;
      for (int i = 0; i <= j; i++) sref = __dadd_rn(sref, __dmul_rn(xs[xs_index(i)], gs[7 + j - i]));
      aucs[j] = sref;
    }
  }
  __syncwarp();
}

// ---- S7 death-delay convolution on the tensor cores, then back to day order through shared memory
// Number of 8-day lag blocks of the delay kernel that can be non-zero.  Beyond x = alph + 9 sqrt(alph) + 45 the upper tail
// is below 2^-54, every P value rounds to exactly 1 (in R as here) and the kernel masses are exactly zero, so the lag blocks
// out there are skipped without changing a bit (short delays: most of the triangle for sod < 0.5).
__device__ inline int conv_lag_blocks(const KModel& m, const RecLayout& L, const double* rec) {
  const int nb = (m.T + 7) >> 3;
  const double k_hi = rec[L.gam + 12];   // +inf for an invalid shape: keep everything
  if (!(k_hi < (double)m.T)) return nb;
  const int kend = (int)k_hi + 2;                     // first day whose mass is exactly zero (tile_gamma)
  return min(nb, ((kend + 14) >> 3) + 1);            // block dl holds lags 8 dl - 7 .. 8 dl + 7
}

template <int NT>
__device__ inline void stage_conv(const KModel& m, const double* __restrict__ xs, const double* __restrict__ gs, double* aucs, int nb) {
  const int lane = threadIdx.x & 31;
  double auc[NT][2];
  nb = __shfl_sync(0xffffffffu, nb, 0);  // warp-uniform by construction; tell the compiler so (uniform loop exits around the MMAs)
  toeplitz_dmma<NT>(xs, gs, nb, auc);
  const int n = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int P = 0; P < NT; P++) {
    aucs[64 * P + 16 * t + n] = auc[P][0];
    aucs[64 * P + 16 * t + 8 + n] = auc[P][1];
  }
  __syncwarp();
}

// ---- S8 Poisson term, evaluated day by day (binomial.cpp:199-221): l1 (without the data-only constant), sum of expd over
// all days, sum of auc over the uncensored days
// `ne` (attack-rate weights) and `par` (the parameter vector) are read only for days whose expected count is zero or
// subnormal: there the reference's own sum  sum_a (auc * ne[a]) * ma[a]  (binomial.cpp:209-212, every product rounded to the
// subnormal grid) decides the value, not auc * sum_a ne[a] ma[a].
// `fix_xs` / `fix_gs`: series and delay kernel of this evaluation if they are still in shared memory (else nullptr): the
// rare path then first redoes the (near) subnormal convolution outputs the reference's way (conv_fixup_tiny).
template <typename Par>
__device__ inline void stage_l1_direct(const KModel& m, const TileTables& tb, double* aucs, double snm, const double* __restrict__ ne,
                                       Par par, double& l1, double& s_all, double& s_unc, const double* fix_xs = nullptr,
                                       const double* fix_gs = nullptr) {
  const int lane = threadIdx.x & 31, T = m.T;
  double dummy = 0.0;
  l1 = s_all = s_unc = 0.0;
  // common case without branches: every observed day has a positive normal mean and a count >= 0
  int odd = 0;
#pragma unroll 2
  for (int j = lane; j < T; j += 32) {
    const double a = aucs[j];
    const double expd = a * snm;
    const int x = tb.obsd[j];
    const bool unc = (j + 1) < m.rcensor_day, use = unc && (x != -1);
    const double lg = fast_log_normal(expd, m.logtab);
    const double lp = (x > 0) ? fma((double)x, lg, -expd) : -expd;  // - lgamma(x+1) is data only: m.l1_const
    // not a positive normal mean (one unsigned compare on the high word), or a negative count
    odd |= (use && (((unsigned)(__double2hiint(expd) - 0x00100000) >= 0x7fe00000u) || x < 0)) ? 1 : 0;
    s_all += expd;
    s_unc += unc ? a : 0.0;
    l1 += use ? lp : 0.0;
  }
  if (__any_sync(0xffffffffu, odd)) {
    // some observed day has a zero / subnormal / negative / non-finite mean or a negative count: R's case analysis
    if (fix_xs) conv_fixup_tiny(m.T, fix_xs, fix_gs, aucs);
    l1 = 0.0;
    for (int j = lane; j < T; j += 32) {
      double expd = aucs[j] * snm;
      const int x = tb.obsd[j];
      if ((j + 1) < m.rcensor_day && x != -1) {
        if (expd >= 0.0 && expd < DBL_MIN && aucs[j] != 0.0) {  // (would-be) subnormal: the reference's operation order
          const double a = aucs[j];
          expd = 0.0;
          for (int q = 0; q < m.A; q++) expd = __dadd_rn(expd, __dmul_rn(__dmul_rn(a, ne[q]), ifr_of(m, par, q)));
        }
        double lp;
        if (!(expd >= 0.0)) lp = d_nan();        // NaN or negative mean -> NaN (R::dpois)
        else if (x < 0) lp = -d_inf();
        else if (x == 0) lp = -expd;
        else lp = (double)x * fast_log(expd, m.logtab) - expd;
        l1 += lp;
      }
    }
  }
  warp_sum4(l1, s_all, s_unc, dummy);
}

// ---- the same term factored for the Metropolis step: with expd_j = auc_j * snm,
//   L1 = sum_{x_j>0} x_j log auc_j + (sum_{x_j>0} x_j) log snm - snm * sum_{observed} auc_j - const,
// so proposals that only move snm (IFR and attack-rate parameters) cost O(1).  `irregular` (any observed day with a
// non-positive / NaN auc, or a negative count) sends the caller back to stage_l1_direct.
struct L1Sums { double s1, sa_obs, s_all_auc, s_unc; int irregular; };
__device__ inline L1Sums stage_l1_sums(const KModel& m, const TileTables& tb, const double* __restrict__ aucs) {
  const int lane = threadIdx.x & 31, T = m.T;
  double s1 = 0.0, sa = 0.0, sall = 0.0, sunc = 0.0;
  int irr = 0;
#pragma unroll 2
  for (int j = lane; j < T; j += 32) {
    const double a = aucs[j];
    sall += a;
    if ((j + 1) < m.rcensor_day) {
      sunc += a;
      const int x = tb.obsd[j];
      if (x != -1) {
        if (!(a > 0.0) || x < 0) irr = 1;
        sa += a;
        if (x > 0) s1 = fma((double)x, fast_log(a, m.logtab), s1);
      }
    }
  }
  warp_sum4(s1, sa, sall, sunc);
  L1Sums r;
  r.s1 = s1; r.sa_obs = sa; r.s_all_auc = sall; r.s_unc = sunc;
  r.irregular = __any_sync(0xffffffffu, irr) ? 1 : 0;
  return r;
}

// ---- S11/S12 survey-window means of infxn (*) H through the prefix-summed hazard (binomial.cpp:288-311); put(s, mean)
template <typename PutBase>
__device__ inline void stage_surveys(const KModel& m, const double* __restrict__ xs, const double* __restrict__ sk, PutBase put) {
  const int lane = threadIdx.x & 31, S = m.S;
  // surveys three at a time in ONE pass over the series: each incidence value is loaded once and feeds three independent
  // accumulators (the order of additions within a survey is unchanged: lane-strided partial sums, then the butterfly)
  for (int s0 = 0; s0 < S; s0 += 3) {
    const int n = min(3, S - s0);
    const int st0 = m.sero_start[s0], en0 = m.sero_end[s0];
    const int st1 = (n > 1) ? m.sero_start[s0 + 1] : 1, en1 = (n > 1) ? m.sero_end[s0 + 1] : 0;
    const int st2 = (n > 2) ? m.sero_start[s0 + 2] : 1, en2 = (n > 2) ? m.sero_end[s0 + 2] : 0;
    const int emax = max(en0, max(en1, en2));
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, dummy = 0.0;
#pragma unroll 2
    for (int t = lane; t < emax; t += 32) {
      const double x = xs[xs_index(t)];
      // past a survey's last day the weight is CH[0] - CH[0] = 0 (indices clamped to 0): the term adds nothing
      const int h0 = max(en0 - t, 0), l0 = max(st0 - 1 - t, 0);
      const int h1 = max(en1 - t, 0), l1 = max(st1 - 1 - t, 0);
      const int h2 = max(en2 - t, 0), l2 = max(st2 - 1 - t, 0);
      if (t < en0) b0 = fma(x, sk[h0] - sk[l0], b0);
      if (t < en1) b1 = fma(x, sk[h1] - sk[l1], b1);
      if (t < en2) b2 = fma(x, sk[h2] - sk[l2], b2);
    }
    warp_sum4(b0, b1, b2, dummy);
    if (lane == 0) {
      put(s0, b0 / (double)(en0 - st0 + 1));
      if (n > 1) put(s0 + 1, b1 / (double)(en1 - st1 + 1));
      if (n > 2) put(s0 + 2, b2 / (double)(en2 - st2 + 1));
    }
  }
}

// phase 2 for vector e of the tile: per-day stages; writes o_status, o_l1, o_texpd, o_sunc, o_base[S] to the scratch
template <int NT, typename Par>
__device__ inline void tile_phase2(const KModel& m, const RecLayout& L, const TileSmem& w, const TileTables& tb, double* __restrict__ scr, int e,
                                   Par par, long long stride = 32) {
  const int lane = threadIdx.x & 31, T = m.T;
  __syncwarp();
  for (int f = lane; f < L.n2; f += 32) w.rec[f] = scr[f * stride + e];
  __syncwarp();
  if (w.rec[L.flags] == 0.0) return;  // knots not increasing: phase 3 returns the failure value

#if CC_P2_EXPERIMENT == 2  // timing experiment (results invalid): the delay kernel alone
  tile_gamma<NT>(m, L, w, tb);
  if (lane == 0) scr[L.o_l1 * stride + e] = w.gs[7 + 5] + w.gs[7 + T - 1];
  return;
#endif
  tile_sero_prefix<NT>(m, L, w, tb);   // S10 (first: with seroreversion it borrows the series/kernel regions)
#if CC_P2_EXPERIMENT == 1  // timing experiment (results invalid): everything but the delay kernel
  for (int k = lane; k < 64 * NT + 9; k += 32) w.gs[7 + k] = (k < T) ? 1.0 / (double)T : 0.0;
  __syncwarp();
#else
  tile_gamma<NT>(m, L, w, tb);         // S2
#endif
  const double tot = stage_spline<NT>(m, L, w.rec, w.xs);
  const bool ok = stage_popn(m, L, w.rec, tot);
  if (lane == 0) scr[L.o_status * stride + e] = ok ? 0.0 : 1.0;
  if (!ok) return;
  __syncwarp();
  // the survey sums come first: they are the last readers of the hazard prefix sums, whose buffer then takes the convolution
  // output - so the series and the delay kernel are still in place when the Poisson term needs them (rare subnormal path)
  stage_surveys(m, w.xs, w.sk, [&](int s, double v) { scr[(L.o_base + s) * stride + e] = v; });
  if (m.debug_dump && e == 0 && blockIdx.x == 0 && (threadIdx.x >> 5) == 0)
    for (int j = lane; j < T; j += 32) {
      m.debug_dump[j] = w.gs[7 + j];
#ifndef CC_DEBUG_RAWQ
      m.debug_dump[3 * T + j] = w.sk[min(j, m.M)];
#endif
    }
  __syncwarp();
  stage_conv<NT>(m, w.xs, w.gs, w.auc, conv_lag_blocks(m, L, w.rec));
  double l1, s_all, s_unc;
  stage_l1_direct(m, tb, w.auc, w.rec[L.snm], w.rec + L.ne, par, l1, s_all, s_unc, w.xs, w.gs);
  if (m.debug_dump && e == 0 && blockIdx.x == 0 && (threadIdx.x >> 5) == 0) {
    for (int j = lane; j < T; j += 32) {
      m.debug_dump[T + j] = w.xs[xs_index(j)];
      m.debug_dump[2 * T + j] = w.auc[j];
    }
  }
  if (lane == 0) {
    scr[L.o_l1 * stride + e] = l1 - m.l1_const;
    scr[L.o_texpd * stride + e] = s_all;
    scr[L.o_sunc * stride + e] = s_unc;
  }
}

// ------------------------------------------------------------------------------------------------ phase 3
// G = 32 / E lanes cooperate on one vector: lane = sub * E + e; terms are strided over sub and xor-reduced over the high bits
// one term of the chained binomial on the age split (binomial.cpp:227-251): stratum a, running subtraction in the reference's order
template <typename Get>
__device__ __forceinline__ double p3_l2_term(const KModel& m, const RecLayout& L, Get get, int a, double texpd, double sum_unc) {
  double te = texpd;
  for (int b = 0; b < a; b++) te -= get(L.nm + b) * sum_unc;
  return dbinom_log_shared(round(get(L.nm + a) * sum_unc), round(te), m.l2_prob[a]);
}

// one term of the serology likelihood (binomial.cpp:316-340; logit.cpp:314-339): survey-stratum cell k = s * A + a
template <typename Get>
__device__ __forceinline__ double p3_l3_term(const KModel& m, const RecLayout& L, Get get, int k, double sens, double spec) {
  const int A = m.A, s = k / A, a = k - s * A;
  const double frac = (get(L.ne + a) * get(L.o_base + s)) / (double)m.demog[a];
  const double prev = sens * frac + (1 - spec) * (1 - frac);
  if (m.binomial) {
    const int kind = m.l3_kind[k];
    if (kind == 1) {
      const double x = m.sero_a[k], nn = m.sero_b[k];
      if (!(prev >= 0.0 && prev <= 1.0)) return d_nan();
      if (x == 0.0) return (nn == 0.0) ? 0.0 : nn * log(1 - prev);
      if (x == nn) return nn * log(prev);
      return m.l3_lchoose[k] + x * log(prev) + (nn - x) * log(1 - prev);
    }
    return (kind == 2) ? dbinom_log_shared(m.sero_a[k], m.sero_b[k], prev) : 0.0;
  }
  const double eps = DBL_MIN / 100.0;
  const double lg = log((prev + eps) / (1 - (prev + eps)));
  return dnorm_log(lg, m.sero_a[k], m.sero_b[k]);
}

__device__ inline double tile_phase3(const KModel& m, const RecLayout& L, const double* __restrict__ scr, int e, int sub, int G, int E,
                                   long long stride = 32) {
  const int A = m.A, S = m.S;
  auto get = [&](int f) { return scr[f * stride + e]; };
  const bool pass = get(L.flags) != 0.0 && get(L.o_status) == 0.0;
  double tail = 0.0;
  if (pass) {
    const double texpd = get(L.o_texpd), sum_unc = get(L.o_sunc);
    for (int a = sub; a < A - 1; a += G) tail += p3_l2_term(m, L, get, a, texpd, sum_unc);   // S9
    const double sens = get(L.sens), spec = get(L.spec);
    for (int k = sub; k < S * A; k += G) tail += p3_l3_term(m, L, get, k, sens, spec);        // S13
  }
  for (int o = E; o < 32; o <<= 1) tail += __shfl_xor_sync(0xffffffffu, tail, o);
  if (!pass) return -CC_OVERFLO;
  double loglik = get(L.o_l1) + tail;
  if (!isfinite(loglik)) loglik = -CC_OVERFLO;
  return loglik;
}

// The same for the Metropolis sweep (one warp, record in shared memory), with the two sums kept apart: a single-site proposal
// that cannot change the age split (test characteristics, seroconversion) or the serology term (IFRs, onset-to-death delay)
// reuses the cached sum.  Returns false when the evaluation has failed a validity check.
__device__ inline bool sweep_phase3(const KModel& m, const RecLayout& L, const double* rec, bool do_l2, bool do_l3, double& l2, double& l3) {
  const int lane = threadIdx.x & 31, A = m.A, S = m.S;
  auto get = [&](int f) { return rec[f]; };
  const bool pass = get(L.flags) != 0.0 && get(L.o_status) == 0.0;
  if (!pass) return false;
  if (do_l2) {
    const double texpd = get(L.o_texpd), sum_unc = get(L.o_sunc);
    double v = 0.0;
    for (int a = lane; a < A - 1; a += 32) v += p3_l2_term(m, L, get, a, texpd, sum_unc);
    l2 = warp_sum(v);
  }
  if (do_l3) {
    const double sens = get(L.sens), spec = get(L.spec);
    double v = 0.0;
    for (int k = lane; k < S * A; k += 32) v += p3_l3_term(m, L, get, k, sens, spec);
    l3 = warp_sum(v);
  }
  return true;
}

// one term of the generated logprior
__device__ __forceinline__ double prior_term(const KModel& m, int i, double x) {
  const int fam = m.prior_family[i];
  if (fam == CC_PRIOR_NONE) return 0.0;
  const double a = m.prior_p1[i], b = m.prior_p2[i];
  return (fam == CC_PRIOR_DUNIF) ? dunif_log(x, a, b) : (fam == CC_PRIOR_DNORM) ? dnorm_log(x, a, b) : dbeta_log(x, a, b);
}

// out-of-line copy for the Metropolis sweep (the beta density alone is ~2 000 instructions)
__device__ __noinline__ double prior_density_shared(int fam, double x, double a, double b) {
  return (fam == CC_PRIOR_DUNIF) ? dunif_log(x, a, b) : (fam == CC_PRIOR_DNORM) ? dnorm_log(x, a, b) : dbeta_log(x, a, b);
}
__device__ __forceinline__ double prior_term_shared(const KModel& m, int i, double x) {
  const int fam = m.prior_family[i];
  return (fam == CC_PRIOR_NONE) ? 0.0 : prior_density_shared(fam, x, m.prior_p1[i], m.prior_p2[i]);
}

// generated logprior for E vectors per warp: G lanes per vector, terms strided over sub (Agecpp2strings.R:66-219)
template <typename Par>
__device__ inline double tile_logprior(const KModel& m, int sub, int G, int E, Par par) {
  double v = 0.0;
  for (int i = sub; i < m.d; i += G) {
    const int fam = m.prior_family[i];
    if (fam != CC_PRIOR_NONE) {
      const double x = par(i), a = m.prior_p1[i], b = m.prior_p2[i];
      v += (fam == CC_PRIOR_DUNIF) ? dunif_log(x, a, b) : (fam == CC_PRIOR_DNORM) ? dnorm_log(x, a, b) : dbeta_log(x, a, b);
    }
  }
  for (int j = sub; j < 3; j += G)
    if (m.jac_rows[j] >= 0) v += m.jac_counts[j] * log(par(max(m.jac_rows[j], 0)));
  for (int o = E; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (!isfinite(v)) v = -CC_OVERFLO;
  return v;
}

}  // namespace ccdev
