// nmath_dev.cuh - FP64 device versions of the R nmath routines on the COVIDCurve hot path.
//
// Reference call sites (under /root/reference): src/natcubicspline_loglike_binomial.cpp:79 (pgamma), :216 (dpois),
// :248,:337 (dbinom), :269 (pweibull); src/natcubicspline_loglike_logit.cpp:337 (dnorm); generated prior
// R/Agecpp2strings.R:73-148 (dunif, dnorm, dbeta).  The algorithms are R 4.0.x nmath's (Loader saddle-point
// dbinom/dpois via stirlerr/bd0; Welinder's pgamma_raw region split), written for the GPU: no recursion, bounded
// loops, reciprocal tables instead of divides where the value is a constant.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "nmath_tables_dev.cuh"

namespace ccdev {

#define CC_LN_SQRT_2PI 0.918938533204672741780329736406
#define CC_LN_2PI 1.837877066409345483560659472811
#define CC_2PI 6.283185307179586476925286766559
#define CC_1_SQRT_2PI 0.398942280401432677939946059934
#define CC_LN2 0.693147180559945309417232121458
#define CC_OVERFLO (DBL_MAX / 100.0)

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------------------------------------- stirlerr / bd0
__device__ __forceinline__ double stirlerr(double n) {
  const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778, S2 = 0.00079365079365079365079365,
               S3 = 0.000595238095238095238095238, S4 = 0.0008417508417508417508417508;
  if (n <= 15.0) {
    double nn = n + n;
    if (nn == (double)(int)nn) return c_sferr_halves[(int)nn];
    return lgamma(n + 1.) - (n + 0.5) * log(n) + n - CC_LN_SQRT_2PI;
  }
  const double rn = 1.0 / n;  // one division; the series in 1/n^2 is evaluated with multiplications
  const double r = rn * rn;
  if (n > 500) return (S0 - S1 * r) * rn;
  if (n > 80) return (S0 - (S1 - S2 * r) * r) * rn;
  if (n > 35) return (S0 - (S1 - (S2 - S3 * r) * r) * r) * rn;
  return (S0 - (S1 - (S2 - (S3 - S4 * r) * r) * r) * r) * rn;
}

// reciprocals of the odd numbers 3, 5, ..., 65 for the Taylor loop of bd0
__device__ __constant__ double c_inv_odd[32] = {
    1.0 / 3,  1.0 / 5,  1.0 / 7,  1.0 / 9,  1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19, 1.0 / 21, 1.0 / 23,
    1.0 / 25, 1.0 / 27, 1.0 / 29, 1.0 / 31, 1.0 / 33, 1.0 / 35, 1.0 / 37, 1.0 / 39, 1.0 / 41, 1.0 / 43, 1.0 / 45,
    1.0 / 47, 1.0 / 49, 1.0 / 51, 1.0 / 53, 1.0 / 55, 1.0 / 57, 1.0 / 59, 1.0 / 61, 1.0 / 63, 1.0 / 65};

// bd0(x, np) = x log(x/np) + np - x, evaluated stably (Loader).
__device__ __forceinline__ double bd0(double x, double np) {
  if (!isfinite(x) || !isfinite(np) || np == 0.0) return d_nan();
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    // |v| < 0.01: ej shrinks by >= 100x per term; exit as soon as a term is absorbed (R's loop, with 1/(2j+1) tabulated)
#pragma unroll 1
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      const double s1 = (j <= 32) ? fma(ej, c_inv_odd[j - 1], s) : s + ej / (double)((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

// ---------------------------------------------------------------------------------------------- dpois
__device__ __forceinline__ double dpois_raw_log(double x, double lambda) {
  if (lambda == 0) return (x == 0) ? 0.0 : -d_inf();
  if (!isfinite(lambda)) return -d_inf();
  if (x < 0) return -d_inf();
  if (x <= lambda * DBL_MIN) return -lambda;
  if (lambda < x * DBL_MIN) {
    if (!isfinite(x)) return -d_inf();
    return -lambda + x * log(lambda) - lgamma(x + 1);
  }
  return -0.5 * log(CC_2PI * x) + (-stirlerr(x) - bd0(x, lambda));
}

// exp(dpois_raw_log) but following R's non-log formula (exp(x)/sqrt(f)) for the mid branch
__device__ __forceinline__ double dpois_raw(double x, double lambda) {
  if (lambda == 0) return (x == 0) ? 1.0 : 0.0;
  if (!isfinite(lambda)) return 0.0;
  if (x < 0) return 0.0;
  if (x <= lambda * DBL_MIN) return exp(-lambda);
  if (lambda < x * DBL_MIN) {
    if (!isfinite(x)) return 0.0;
    return exp(-lambda + x * log(lambda) - lgamma(x + 1));
  }
  return exp(-stirlerr(x) - bd0(x, lambda)) * rsqrt(CC_2PI * x);
}

// R::dgamma(x, shape, scale, log = FALSE) (nmath/dgamma.c): the density kernel of posterior_check_infxns_to_death,
// R/summarize_plot.R:609
__device__ __forceinline__ double dgamma_density(double x, double shape, double scale) {
  if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
  if (shape < 0 || scale <= 0) return d_nan();
  if (x < 0) return 0.0;
  if (shape == 0) return (x == 0) ? d_inf() : 0.0;
  if (x == 0) {
    if (shape < 1) return d_inf();
    if (shape > 1) return 0.0;
    return 1 / scale;
  }
  if (shape < 1) return dpois_raw(shape, x / scale) * shape / x;
  return dpois_raw(shape - 1, x / scale) / scale;
}

// R::dpois(x, lambda, log = TRUE) for integer-valued x (obs deaths are ints in the reference, binomial.cpp:204)
__device__ __forceinline__ double dpois_log_int(double x, double lambda) {
  if (isnan(lambda)) return lambda;
  if (lambda < 0) return d_nan();
  if (x < 0) return -d_inf();
  return dpois_raw_log(x, lambda);
}

// ---------------------------------------------------------------------------------------------- dbinom
__device__ __forceinline__ double dbinom_raw_log(double x, double n, double p, double q) {
  if (p == 0) return (x == 0) ? 0.0 : -d_inf();
  if (q == 0) return (x == n) ? 0.0 : -d_inf();
  if (x == 0) {
    if (n == 0) return 0.0;
    return (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
  }
  if (x == n) return (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
  if (x < 0 || x > n) return -d_inf();
  double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
  double lf = CC_LN_2PI + log(x) + log1p(-x / n);
  return lc - 0.5 * lf;
}

__device__ __forceinline__ bool r_nonint(double x) { return fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x)); }

// R::dbinom(x, n, p, log = TRUE)
__device__ __forceinline__ double dbinom_log(double x, double n, double p) {
  if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
  if (p < 0 || p > 1 || n < 0 || r_nonint(n)) return d_nan();
  if (r_nonint(x)) return -d_inf();
  if (x < 0 || !isfinite(x)) return -d_inf();
  n = nearbyint(n);
  x = nearbyint(x);
  return dbinom_raw_log(x, n, p, 1 - p);
}

// ---------------------------------------------------------------------------------------------- simple densities
__device__ __forceinline__ double dnorm_log(double x, double mu, double sigma) {
  if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
  if (!isfinite(sigma)) return -d_inf();
  if (!isfinite(x) && mu == x) return d_nan();
  if (sigma <= 0) {
    if (sigma < 0) return d_nan();
    return (x == mu) ? d_inf() : -d_inf();
  }
  x = (x - mu) / sigma;
  if (!isfinite(x)) return -d_inf();
  x = fabs(x);
  if (x >= 2 * sqrt(DBL_MAX)) return -d_inf();
  return -(CC_LN_SQRT_2PI + 0.5 * x * x + log(sigma));
}

__device__ __forceinline__ double dunif_log(double x, double a, double b) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (b <= a) return d_nan();
  if (a <= x && x <= b) return -log(b - a);
  return -d_inf();
}

__device__ __forceinline__ double dbeta_log(double x, double a, double b) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (a < 0 || b < 0) return d_nan();
  if (x < 0 || x > 1) return -d_inf();
  if (a == 0 || b == 0 || !isfinite(a) || !isfinite(b)) {
    if (a == 0 && b == 0) return (x == 0 || x == 1) ? d_inf() : -d_inf();
    if (a == 0 || a / b == 0) return (x == 0) ? d_inf() : -d_inf();
    if (b == 0 || b / a == 0) return (x == 1) ? d_inf() : -d_inf();
    return (x == 0.5) ? d_inf() : -d_inf();
  }
  if (x == 0) {
    if (a > 1) return -d_inf();
    if (a < 1) return d_inf();
    return log(b);
  }
  if (x == 1) {
    if (b > 1) return -d_inf();
    if (b < 1) return d_inf();
    return log(a);
  }
  if (a <= 2 || b <= 2) return (a - 1) * log(x) + (b - 1) * log1p(-x) - (lgamma(a) + lgamma(b) - lgamma(a + b));
  return log(a + b - 1) + dbinom_raw_log(a - 1, a + b - 2, x, 1 - x);
}

__device__ __forceinline__ double pweibull_lower(double x, double shape, double scale) {
  if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
  if (shape <= 0 || scale <= 0) return d_nan();
  if (x <= 0) return 0.0;
  return -expm1(-pow(x / scale, shape));
}

// ---------------------------------------------------------------------------------------------- pgamma (lower tail)
// Regularised lower incomplete gamma P(alph, x), x already divided by the scale.  Region split of R's pgamma_raw:
//   x < 1                                   : A&S 6.5.29 alternating series
//   x <= alph-1 and x < 0.8(alph+50)        : d * sum_{n>=1} x^n / (alph (alph+1) ... (alph+n-1))
//   alph-1 < x and alph < 0.8(x+50)         : 1 - d * (1 + descending series [+ continued-fraction tail])
//   otherwise (both large, x near alph)     : Temme-type asymptotic expansion of the Poisson cdf
// with d = dpois_wrap(alph, x) = x^(alph-1) e^-x / Gamma(alph).
struct GammaShape {  // everything that depends on alph only (one per parameter vector)
  double alph;
  double am1;          // alph - 1
  double lgam_a1;      // lgamma(alph + 1)   (small-x prefactor)
  double stir_am1;     // stirlerr(alph - 1) when alph > 1
  double stir_a;       // stirlerr(alph)
  double rs2pi_am1;    // 1/sqrt(2 pi (alph-1))
  double rs2pi_a;      // 1/sqrt(2 pi alph)
  double mhl2pi_am1;   // -0.5 log(2 pi (alph-1)) when alph > 1 (log form of the Poisson density)
};

__device__ __forceinline__ GammaShape gamma_shape(double alph) {
  GammaShape g;
  g.alph = alph;
  g.am1 = alph - 1;
  g.lgam_a1 = lgamma(alph + 1.0);
  g.stir_a = (alph > 0) ? stirlerr(alph) : 0.0;
  g.rs2pi_a = (alph > 0) ? rsqrt(CC_2PI * alph) : 0.0;
  g.stir_am1 = (alph > 1) ? stirlerr(alph - 1) : 0.0;
  g.rs2pi_am1 = (alph > 1) ? rsqrt(CC_2PI * (alph - 1)) : 0.0;
  g.mhl2pi_am1 = (alph > 1) ? -0.5 * log(CC_2PI * (alph - 1)) : 0.0;
  return g;
}

// exp(x) whose SUBNORMAL results are rounded once, to the subnormal grid (like a correctly rounded libm exp): the value is
// formed in the normal range and scaled by an exact power of two.  Matters only below ~1e-308, where the reference's
// gamma-cdf values carry a handful of bits and every rounding shows in the log-likelihood.
__device__ __forceinline__ double exp_gradual(double x) {
  if (x > -700.0) return exp(x);
  return ldexp(exp(x + 200.0 * CC_LN2), -200);
}

// dpois_raw(k, lambda) with k in {alph-1, alph}: Loader form with the k-only terms taken from GammaShape
__device__ __forceinline__ double dpois_shape(double k, double stir_k, double rs2pi_k, double lambda) {
  if (lambda == 0) return (k == 0) ? 1.0 : 0.0;
  if (!isfinite(lambda)) return 0.0;
  if (k <= lambda * DBL_MIN) return exp(-lambda);
  if (lambda < k * DBL_MIN) return exp(-lambda + k * log(lambda) - lgamma(k + 1));
  (void)rs2pi_k;
  return exp_gradual(-stir_k - bd0(k, lambda)) / sqrt(CC_2PI * k);  // R_D_fexp(2 pi k, .): same rounding points as R
}

// dpois_wrap(alph, x) = dpois(alph-1, x)
__device__ __forceinline__ double dpois_wrap(const GammaShape& g, double x) {
  if (!isfinite(x)) return 0.0;
  if (g.alph > 1) return dpois_shape(g.am1, g.stir_am1, g.rs2pi_am1, x);
  const double M_cutoff = CC_LN2 * DBL_MAX_EXP / DBL_EPSILON;
  if (x > fabs(g.am1) * M_cutoff) return exp(-x - lgamma(g.alph));
  return dpois_shape(g.alph, g.stir_a, g.rs2pi_a, x) * (g.alph / x);
}

__device__ __forceinline__ double log1pmx(double x) {  // log(1+x) - x, accurate for small x
  if (x > 1 || x < -0.79149064) return log1p(x) - x;
  double r = x / (2 + x), y = r * r;
  if (fabs(x) < 1e-2) return r * ((((2.0 / 9 * y + 2.0 / 7) * y + 2.0 / 5) * y + 2.0 / 3) * y - x);
  // S(y) = sum_k y^k/(2k+3); |y| <= 1/9 on this range: 40 terms reach 1e-38
  double s = 0.0, yk = 1.0;
#pragma unroll 1
  for (int k = 0; k < 60; k++) {
    double t = yk / (double)(2 * k + 3);
    s += t;
    if (t < 1e-18 * s) break;
    yk *= y;
  }
  return r * (2 * y * s - x);
}

// log of the standard normal lower tail at x < -sqrt(32): the tail branch of R's pnorm_both() (Cody 1969, Algorithm 715
// ANORM: rational approximation in 1/x^2, argument split at 1/16 so that exp(-x^2/2) loses no bits) on the log scale
__device__ __forceinline__ double pnorm_log_lower_tail(double x) {
  const double p0 = 0.21589853405795699, p1 = 0.1274011611602473639, p2 = 0.022235277870649807, p3 = 0.001421619193227893466,
               p4 = 2.9112874951168792e-5, p5 = 0.02307344176494017303;
  const double q0 = 1.28426009614491121, q1 = 0.468238212480865118, q2 = 0.0659881378689285515, q3 = 0.00378239633202758244,
               q4 = 7.29751555083966205e-5;
  const double y = fabs(x);
  double xsq = 1.0 / (x * x);
  double xnum = p5 * xsq, xden = xsq;
  xnum = (xnum + p0) * xsq; xden = (xden + q0) * xsq;
  xnum = (xnum + p1) * xsq; xden = (xden + q1) * xsq;
  xnum = (xnum + p2) * xsq; xden = (xden + q2) * xsq;
  xnum = (xnum + p3) * xsq; xden = (xden + q3) * xsq;
  double temp = xsq * (xnum + p4) / (xden + q4);
  temp = (CC_1_SQRT_2PI - temp) / y;
  xsq = trunc(x * 16) / 16;
  const double del = (x - xsq) * (x + xsq);
  return (-xsq * xsq * 0.5) + (-del * 0.5) + log(temp);
}

// upper tail of Poisson(lambda) at x (= lower tail of gamma(alph = x+1) at lambda): R's ppois_asymp(x, lambda, FALSE, log_p).
// The log form is what pgamma_raw falls back to when the plain value comes out below DBL_MIN / DBL_EPSILON (far lower tail of
// very peaked delay distributions); it is valid for s2pt < -sqrt(32), which such a small value implies.
__device__ __noinline__ double ppois_asymp_upper(double x, double lambda, bool log_p) {
  const double coefs_a[8] = {-1e99, 2 / 3., -4 / 135., 8 / 2835., 16 / 8505., -8992 / 12629925.,
                             -334144 / 492567075., 698752 / 1477701225.};
  const double coefs_b[8] = {-1e99, 1 / 12., 1 / 288., -139 / 51840., -571 / 2488320.,
                             163879 / 209018880., 5246819 / 75246796800., -534703531 / 902961561600.};
  double dfm = lambda - x;
  double pt_ = -log1pmx(dfm / x);
  double s2pt = sqrt(2 * x * pt_);
  if (dfm < 0) s2pt = -s2pt;
  double res12 = 0, res1_term, res1_ig, res2_term, res2_ig;
  res1_ig = res1_term = sqrt(x);
  res2_ig = res2_term = s2pt;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    res12 += res1_ig * coefs_a[i];
    res12 += res2_ig * coefs_b[i];
    res1_term *= pt_ / i;
    res2_term *= 2 * pt_ / (2 * i + 1);
    res1_ig = res1_ig / x + res1_term;
    res2_ig = res2_ig / x + res2_term;
  }
  double elfb = x, elfb_term = 1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    elfb += elfb_term * coefs_b[i];
    elfb_term /= x;
  }
  elfb = -elfb;  // upper tail
  double f = res12 / elfb;
  if (log_p && s2pt < -5.656854249492380195206754896838) {
    // np = pnorm(s2pt, lower, log); dpnorm(s2pt, lower, np) = phi(s2pt) / Phi(s2pt) by the Mills-ratio series (|s2pt| > 10)
    const double np = pnorm_log_lower_tail(s2pt);
    const double ax = -s2pt;
    double n_d_over_p;
    if (ax > 10) {
      double term = 1 / ax, sum = term, x2 = ax * ax, i = 1;
#pragma unroll 1
      do { term *= -i / x2; sum += term; i += 2; } while (fabs(term) > DBL_EPSILON * sum);
      n_d_over_p = 1 / sum;
    } else {
      n_d_over_p = CC_1_SQRT_2PI * exp(-0.5 * ax * ax) / exp(np);
    }
    return np + log1p(f * n_d_over_p);
  }
  double np = 0.5 * erfc(-s2pt * 0.70710678118654752440);  // pnorm(s2pt, lower = TRUE)
  double nd = CC_1_SQRT_2PI * exp(-0.5 * s2pt * s2pt);
  return log_p ? log(np + f * nd) : np + f * nd;
}

// continued fraction of R's pd_lower_cf (only reached for alph < 1, i.e. sod > 1)
__device__ __noinline__ double pd_lower_cf(double y, double d) {
  const double scalefactor = 1.157920892373162e+77;
  double f = 0.0, of, f0, i, c2, c3, c4, a1, b1, a2, b2;
  if (y == 0) return 0;
  f0 = y / d;
  if (fabs(y - 1) < fabs(d) * DBL_EPSILON) return f0;
  if (f0 > 1.) f0 = 1.;
  c2 = y; c4 = d;
  a1 = 0; b1 = 1; a2 = y; b2 = d;
  while (b2 > scalefactor) { a1 /= scalefactor; b1 /= scalefactor; a2 /= scalefactor; b2 /= scalefactor; }
  i = 0; of = -1.;
#pragma unroll 1
  while (i < 200000) {
    i++; c2--; c3 = i * c2; c4 += 2;
    a1 = c4 * a2 + c3 * a1;
    b1 = c4 * b2 + c3 * b1;
    i++; c2--; c3 = i * c2; c4 += 2;
    a2 = c4 * a1 + c3 * a2;
    b2 = c4 * b1 + c3 * b2;
    if (b2 > scalefactor) { a1 /= scalefactor; b1 /= scalefactor; a2 /= scalefactor; b2 /= scalefactor; }
    if (b2 != 0) {
      f = a2 / b2;
      if (fabs(f - of) <= DBL_EPSILON * fmax(f0, fabs(f))) return f;
      of = f;
    }
  }
  return f;
}

__device__ __forceinline__ double pgamma_lower(const GammaShape& g, double x) {
  const double alph = g.alph;
  if (isnan(x) || isnan(alph)) return x + alph;
  if (alph < 0) return d_nan();
  if (x <= 0) return 0.0;
  if (!isfinite(x)) return 1.0;
  if (alph == 0) return 1.0;
  double res;
  if (x < 1) {
    double sum = 0, c = alph, n = 0, term;
#pragma unroll 1
    do {
      n += 1;
      c *= -x / n;
      term = c / (alph + n);
      sum += term;
    } while (fabs(term) > DBL_EPSILON * fabs(sum));
    double f1 = 1 + sum, f2;
    if (alph > 1) f2 = dpois_shape(alph, g.stir_a, g.rs2pi_a, x) * exp(x);
    else f2 = pow(x, alph) * exp(-g.lgam_a1);
    res = f1 * f2;
  } else if (x <= alph - 1 && x < 0.8 * (alph + 50)) {
    double y = alph, term = x / y, sum = term;
#pragma unroll 1
    do {
      y += 1;
      term *= x / y;
      sum += term;
    } while (term > sum * DBL_EPSILON);
    if (alph > 1 && isfinite(x)) {
      // dpois_wrap(alph, x) = dpois_raw(alph - 1, x) in its main (saddle-point) form: for x >= 1 and alph > 1 none of its
      // special cases applies.  The exponent is kept: R's pgamma_raw redoes results below DBL_MIN / DBL_EPSILON (1.002e-292)
      // in log space - exp(log(sum) + dpois_wrap(.., log = TRUE)) - because the product loses its bits to gradual underflow.
      // Only this branch can get that small for a lower tail with shape > 1.
      const double arg = -g.stir_am1 - bd0(g.am1, x);
      res = sum * (exp_gradual(arg) / sqrt(CC_2PI * g.am1));
      if (res < DBL_MIN / DBL_EPSILON) res = exp_gradual(log(sum) + (g.mhl2pi_am1 + arg));
    } else {
      res = sum * dpois_wrap(g, x);
    }
  } else if (alph - 1 < x && alph < 0.8 * (x + 50)) {
    double d = dpois_wrap(g, x), sum;
    if (alph < 1) {
      if (x * DBL_EPSILON > 1 - alph) sum = 1.0;
      else sum = pd_lower_cf(alph, x - (alph - 1)) * x / alph;
    } else {
      double y = alph - 1, term = 1, s = 0;
      const double rx = 1.0 / x;
#pragma unroll 1
      while (y >= 1 && term > s * DBL_EPSILON) {
        term *= y * rx;
        s += term;
        y -= 1;
      }
      if (y != floor(y)) s += term * pd_lower_cf(y, x + 1 - y);
      sum = 1 + s;
    }
    res = 1 - d * sum;
  } else {
    res = ppois_asymp_upper(alph - 1, x, false);
    // R's pgamma_raw: a result below DBL_MIN / DBL_EPSILON is recomputed on the log scale and exponentiated once
    if (res < DBL_MIN / DBL_EPSILON) res = exp_gradual(ppois_asymp_upper(alph - 1, x, true));
  }
  return res;
}

}  // namespace ccdev
