/*
 * oracle/nmath_port.cpp - TEST INFRASTRUCTURE ONLY (see nmath_port.h).
 *
 * Restatement of the R 4.0.x nmath routines used on the COVIDCurve hot path.  Call sites in the
 * reference: src/natcubicspline_loglike_binomial.cpp:79 (pgamma), :216 (dpois), :248,:337 (dbinom),
 * :269 (pweibull); src/natcubicspline_loglike_logit.cpp:337 (dnorm); R/Agecpp2strings.R:73-148
 * (dunif, dnorm, dbeta in the generated prior).
 *
 * pnorm: R's pnorm_both() (Cody 1969, "Rational Chebyshev approximation for the error function"; the ANORM routine of
 * Algorithm 715, whose published coefficient tables are restated below) including its log-scale tail expansion.  pgamma_raw
 * reaches it through ppois_asymp(); the log tail matters where a gamma-cdf value below 1.002e-292 coming out of that branch is
 * redone in log space (gamma shapes above ~33 500, i.e. sod < 0.0055).  The tables are pinned numerically by
 * tests/test_oracle_nmath.py: against erfc over the central range and against mpmath (50 digits) in the tails.
 */
#include "nmath_port.h"

#include <cfloat>
#include <cmath>
#include <limits>

#include "nmath_tables.h"

namespace {

const double kLnSqrt2Pi = 0.918938533204672741780329736406;  /* log(sqrt(2*pi)) */
const double kLn2Pi = 1.837877066409345483560659472811;      /* log(2*pi) */
const double k2Pi = 6.283185307179586476925286766559;
const double kLn2 = 0.693147180559945309417232121458;
const double k1Sqrt2Pi = 0.398942280401432677939946059934;   /* 1/sqrt(2*pi) */
const double kNaN = std::numeric_limits<double>::quiet_NaN();
const double kInf = std::numeric_limits<double>::infinity();

inline double d0(int lg) { return lg ? -kInf : 0.0; }
inline double d1(int lg) { return lg ? 0.0 : 1.0; }
inline double dexp_(double x, int lg) { return lg ? x : std::exp(x); }
inline double forceint(double x) { return std::nearbyint(x); }
inline bool nonint(double x) { return std::fabs(x - forceint(x)) > 1e-7 * std::fmax(1.0, std::fabs(x)); }
inline double log1_exp(double x) { return x > -kLn2 ? std::log(-std::expm1(x)) : std::log1p(-std::exp(x)); }

/* 2^256 : rescaling constant of the continued fractions */
const double kScale = 1.157920892373162e+77;

/* continued fraction for sum_{k>=0} x^k/(i+k*d), used by log1pmx / lgamma1p */
double logcf(double x, double i, double d, double eps) {
  double c1 = 2 * d, c2 = i + d, c4 = c2 + d;
  double a1 = c2, b1 = i * (c2 - i * x), b2 = d * d * x;
  double a2 = c4 * c2 - b2;
  b2 = c4 * b1 - i * b2;
  while (std::fabs(a2 * b1 - a1 * b2) > std::fabs(eps * b1 * b2)) {
    double c3 = c2 * c2 * x;
    c2 += d; c4 += d;
    a1 = c4 * a2 - c3 * a1;
    b1 = c4 * b2 - c3 * b1;
    c3 = c1 * c1 * x;
    c1 += d; c4 += d;
    a2 = c4 * a1 - c3 * a2;
    b2 = c4 * b1 - c3 * b2;
    if (std::fabs(b2) > kScale) {
      a1 /= kScale; b1 /= kScale; a2 /= kScale; b2 /= kScale;
    } else if (std::fabs(b2) < 1 / kScale) {
      a1 *= kScale; b1 *= kScale; a2 *= kScale; b2 *= kScale;
    }
  }
  return a2 / b2;
}

/* R's pnorm_both(): cum = P[X <= x], ccum = P[X > x] for X ~ N(0,1); i_tail 0: lower only, 1: upper only, 2: both */
void pnorm_both(double x, double* cum, double* ccum, int i_tail, int log_p) {
  static const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582, 18154.981253343561249,
                              0.065682337918207449113};
  static const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205, 45507.789335026729956};
  static const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979, 597.27027639480026226,
                              2494.5375852903726711, 6848.1904505362823326, 11602.651437647350124, 9842.7148383839780218,
                              1.0765576773720192317e-8};
  static const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805, 6485.558298266760755,
                              18615.571640885098091, 34900.952721145977266, 38912.003286093271411, 19685.429676859990727};
  static const double p[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807, 0.001421619193227893466,
                              2.9112874951168792e-5, 0.02307344176494017303};
  static const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515, 0.00378239633202758244,
                              7.29751555083966205e-5};
  const double kSqrt32 = 5.656854249492380195206754896838;
  double xden, xnum, temp, del, xsq;
  const double eps = DBL_EPSILON * 0.5;
  const int lower = i_tail != 1, upper = i_tail != 0;
  const double y = std::fabs(x);
  *cum = *ccum = 0.0;
  auto do_del = [&](double X) {
    xsq = std::trunc(X * 16) / 16;
    del = (X - xsq) * (X + xsq);
    if (log_p) {
      *cum = (-xsq * xsq * 0.5) + (-del * 0.5) + std::log(temp);
      if ((lower && x > 0.) || (upper && x <= 0.)) *ccum = std::log1p(-std::exp(-xsq * xsq * 0.5) * std::exp(-del * 0.5) * temp);
    } else {
      *cum = std::exp(-xsq * xsq * 0.5) * std::exp(-del * 0.5) * temp;
      *ccum = 1.0 - *cum;
    }
  };
  auto swap_tail = [&]() {
    if (x > 0.) {
      temp = *cum;
      if (lower) *cum = *ccum;
      *ccum = temp;
    }
  };
  if (y <= 0.67448975) {
    if (y > eps) {
      xsq = x * x;
      xnum = a[4] * xsq;
      xden = xsq;
      for (int i = 0; i < 3; ++i) {
        xnum = (xnum + a[i]) * xsq;
        xden = (xden + b[i]) * xsq;
      }
    } else {
      xnum = xden = 0.0;
    }
    temp = x * (xnum + a[3]) / (xden + b[3]);
    if (lower) *cum = 0.5 + temp;
    if (upper) *ccum = 0.5 - temp;
    if (log_p) {
      if (lower) *cum = std::log(*cum);
      if (upper) *ccum = std::log(*ccum);
    }
  } else if (y <= kSqrt32) {
    xnum = c[8] * y;
    xden = y;
    for (int i = 0; i < 7; ++i) {
      xnum = (xnum + c[i]) * y;
      xden = (xden + d[i]) * y;
    }
    temp = (xnum + c[7]) / (xden + d[7]);
    do_del(y);
    swap_tail();
  } else if ((log_p && y < 1e170) || (lower && -37.5193 < x && x < 8.2924) || (upper && -8.2924 < x && x < 37.5193)) {
    xsq = 1.0 / (x * x);
    xnum = p[5] * xsq;
    xden = xsq;
    for (int i = 0; i < 4; ++i) {
      xnum = (xnum + p[i]) * xsq;
      xden = (xden + q[i]) * xsq;
    }
    temp = xsq * (xnum + p[4]) / (xden + q[4]);
    temp = (k1Sqrt2Pi - temp) / y;
    do_del(x);
    swap_tail();
  } else {
    if (x > 0) { *cum = d1(log_p); *ccum = d0(log_p); }
    else { *cum = d0(log_p); *ccum = d1(log_p); }
  }
}

double pnorm_std(double x, int lower_tail, int log_p) {
  if (std::isnan(x)) return x;
  if (!std::isfinite(x)) return (x < 0) ? (lower_tail ? d0(log_p) : d1(log_p)) : (lower_tail ? d1(log_p) : d0(log_p));
  double p, cp;
  pnorm_both(x, &p, &cp, lower_tail ? 0 : 1, log_p);
  return lower_tail ? p : cp;
}

double dpnorm(double x, int lower_tail, double lp) {
  if (x < 0) { x = -x; lower_tail = !lower_tail; }
  if (x > 10 && !lower_tail) {
    double term = 1 / x, sum = term, x2 = x * x, i = 1;
    do { term *= -i / x2; sum += term; i += 2; } while (std::fabs(term) > DBL_EPSILON * sum);
    return 1 / sum;
  }
  double d = k1Sqrt2Pi * std::exp(-0.5 * x * x);
  return d / std::exp(lp);
}

/* dpois(x_plus_1 - 1, lambda) with care for x_plus_1 <= 1 */
double dpois_wrap(double x_plus_1, double lambda, int give_log) {
  const double M_cutoff = kLn2 * DBL_MAX_EXP / DBL_EPSILON;
  if (!std::isfinite(lambda)) return d0(give_log);
  if (x_plus_1 > 1) return cco_dpois_raw(x_plus_1 - 1, lambda, give_log);
  if (lambda > std::fabs(x_plus_1 - 1) * M_cutoff) return dexp_(-lambda - std::lgamma(x_plus_1), give_log);
  double d = cco_dpois_raw(x_plus_1, lambda, give_log);
  return give_log ? d + std::log(x_plus_1 / lambda) : d * (x_plus_1 / lambda);
}

/* Abramowitz & Stegun 6.5.29, x < 1 */
double pgamma_smallx(double x, double alph, int lower_tail, int log_p) {
  double sum = 0, c = alph, n = 0, term;
  do {
    n++;
    c *= -x / n;
    term = c / (alph + n);
    sum += term;
  } while (std::fabs(term) > DBL_EPSILON * std::fabs(sum));
  if (lower_tail) {
    double f1 = log_p ? std::log1p(sum) : 1 + sum, f2;
    if (alph > 1) {
      f2 = cco_dpois_raw(alph, x, log_p);
      f2 = log_p ? f2 + x : f2 * std::exp(x);
    } else if (log_p) {
      f2 = alph * std::log(x) - cco_lgamma1p(alph);
    } else {
      f2 = std::pow(x, alph) / std::exp(cco_lgamma1p(alph));
    }
    return log_p ? f1 + f2 : f1 * f2;
  }
  double lf2 = alph * std::log(x) - cco_lgamma1p(alph);
  if (log_p) return log1_exp(std::log1p(sum) + lf2);
  double f1m1 = sum, f2m1 = std::expm1(lf2);
  return -(f1m1 + f2m1 + f1m1 * f2m1);
}

/* sum_{n>=1} lambda^n / ((x+1)(x+2)...(x+n)) */
double pd_upper(double x, double lambda, int log_p) {
  x++;
  double term = lambda / x, sum = term;
  do {
    x++;
    term *= lambda / x;
    sum += term;
  } while (term > sum * DBL_EPSILON);
  return log_p ? std::log(sum) : sum;
}

/* continued fraction ~ (y/d)(1 + (1-y)/d + ...) */
double pd_lower_cf(double y, double d) {
  double f = 0.0, of, f0, i, c2, c3, c4, a1, b1, a2, b2;
  if (y == 0) return 0;
  f0 = y / d;
  if (std::fabs(y - 1) < std::fabs(d) * DBL_EPSILON) return f0;
  if (f0 > 1.) f0 = 1.;
  c2 = y; c4 = d;
  a1 = 0; b1 = 1; a2 = y; b2 = d;
  while (b2 > kScale) { a1 /= kScale; b1 /= kScale; a2 /= kScale; b2 /= kScale; }
  i = 0; of = -1.;
  while (i < 200000) {
    i++; c2--; c3 = i * c2; c4 += 2;
    a1 = c4 * a2 + c3 * a1;
    b1 = c4 * b2 + c3 * b1;
    i++; c2--; c3 = i * c2; c4 += 2;
    a2 = c4 * a1 + c3 * a2;
    b2 = c4 * b1 + c3 * b2;
    if (b2 > kScale) { a1 /= kScale; b1 /= kScale; a2 /= kScale; b2 /= kScale; }
    if (b2 != 0) {
      f = a2 / b2;
      if (std::fabs(f - of) <= DBL_EPSILON * std::fmax(f0, std::fabs(f))) return f;
      of = f;
    }
  }
  return f;
}

/* sum_{n>=0} y(y-1)...(y-n)/lambda^(n+1) (+ continued-fraction tail for non-integer y) */
double pd_lower_series(double lambda, double y) {
  double term = 1, sum = 0;
  while (y >= 1 && term > sum * DBL_EPSILON) {
    term *= y / lambda;
    sum += term;
    y--;
  }
  if (y != std::floor(y)) {
    double f = pd_lower_cf(y, lambda + 1 - y);
    sum += term * f;
  }
  return sum;
}

/* asymptotic expansion of the Poisson cdf around the mean (Temme), x near lambda and both large */
double ppois_asymp(double x, double lambda, int lower_tail, int log_p) {
  static const double coefs_a[8] = {-1e99, 2 / 3., -4 / 135., 8 / 2835., 16 / 8505., -8992 / 12629925.,
                                    -334144 / 492567075., 698752 / 1477701225.};
  static const double coefs_b[8] = {-1e99, 1 / 12., 1 / 288., -139 / 51840., -571 / 2488320.,
                                    163879 / 209018880., 5246819 / 75246796800., -534703531 / 902961561600.};
  double dfm = lambda - x;
  double pt_ = -cco_log1pmx(dfm / x);
  double s2pt = std::sqrt(2 * x * pt_);
  if (dfm < 0) s2pt = -s2pt;
  double res12 = 0, res1_term, res1_ig, res2_term, res2_ig;
  res1_ig = res1_term = std::sqrt(x);
  res2_ig = res2_term = s2pt;
  for (int i = 1; i < 8; i++) {
    res12 += res1_ig * coefs_a[i];
    res12 += res2_ig * coefs_b[i];
    res1_term *= pt_ / i;
    res2_term *= 2 * pt_ / (2 * i + 1);
    res1_ig = res1_ig / x + res1_term;
    res2_ig = res2_ig / x + res2_term;
  }
  double elfb = x, elfb_term = 1;
  for (int i = 1; i < 8; i++) {
    elfb += elfb_term * coefs_b[i];
    elfb_term /= x;
  }
  if (!lower_tail) elfb = -elfb;
  double f = res12 / elfb;
  double np = pnorm_std(s2pt, !lower_tail, log_p);
  if (log_p) {
    double n_d_over_p = dpnorm(s2pt, !lower_tail, np);
    return np + std::log1p(f * n_d_over_p);
  }
  double nd = k1Sqrt2Pi * std::exp(-0.5 * s2pt * s2pt);
  return np + f * nd;
}

double pgamma_raw(double x, double alph, int lower_tail, int log_p) {
  double res;
  if (x <= 0) return lower_tail ? d0(log_p) : d1(log_p);
  if (x >= kInf) return lower_tail ? d1(log_p) : d0(log_p);
  if (x < 1) {
    res = pgamma_smallx(x, alph, lower_tail, log_p);
  } else if (x <= alph - 1 && x < 0.8 * (alph + 50)) {
    double sum = pd_upper(alph - 1, x, log_p);
    double d = dpois_wrap(alph, x, log_p);
    if (!lower_tail) res = log_p ? log1_exp(d + sum) : 1 - d * sum;
    else res = log_p ? sum + d : sum * d;
  } else if (alph - 1 < x && alph < 0.8 * (x + 50)) {
    double sum;
    double d = dpois_wrap(alph, x, log_p);
    if (alph < 1) {
      if (x * DBL_EPSILON > 1 - alph) sum = d1(log_p);
      else {
        double f = pd_lower_cf(alph, x - (alph - 1)) * x / alph;
        sum = log_p ? std::log(f) : f;
      }
    } else {
      sum = pd_lower_series(x, alph - 1);
      sum = log_p ? std::log1p(sum) : 1 + sum;
    }
    if (!lower_tail) res = log_p ? sum + d : sum * d;
    else res = log_p ? log1_exp(d + sum) : 1 - d * sum;
  } else {
    res = ppois_asymp(alph - 1, x, !lower_tail, log_p);
  }
  if (!log_p && res < DBL_MIN / DBL_EPSILON) return std::exp(pgamma_raw(x, alph, lower_tail, 1));
  return res;
}

}  // namespace

extern "C" {

double cco_log1pmx(double x) {
  static const double minLog1Value = -0.79149064;
  if (x > 1 || x < minLog1Value) return std::log1p(x) - x;
  double r = x / (2 + x), y = r * r;
  if (std::fabs(x) < 1e-2) {
    static const double two = 2;
    return r * ((((two / 9 * y + two / 7) * y + two / 5) * y + two / 3) * y - x);
  }
  return r * (2 * y * logcf(y, 3, 2, 1e-14) - x);
}

double cco_lgamma1p(double a) {
  if (std::fabs(a) >= 0.5) return std::lgamma(a + 1);
  const double eulers_const = 0.5772156649015328606065120900824024;
  const int N = 40;
  double lgam = CCO_LGAMMA1P_C * logcf(-a / 2, N + 2, 1, 1e-14);
  for (int i = N - 1; i >= 0; i--) lgam = CCO_LGAMMA1P_COEFFS[i] - a * lgam;
  return (a * lgam - eulers_const) * a - cco_log1pmx(a);
}

double cco_stirlerr(double n) {
  const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778, S2 = 0.00079365079365079365079365,
               S3 = 0.000595238095238095238095238, S4 = 0.0008417508417508417508417508;
  double nn;
  if (n <= 15.0) {
    nn = n + n;
    if (nn == (int)nn) return CCO_SFERR_HALVES[(int)nn];
    return std::lgamma(n + 1.) - (n + 0.5) * std::log(n) + n - kLnSqrt2Pi;
  }
  nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

double cco_bd0(double x, double np) {
  if (!std::isfinite(x) || !std::isfinite(np) || np == 0.0) return kNaN;
  if (std::fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (std::fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / ((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * std::log(x / np) + np - x;
}

double cco_dpois_raw(double x, double lambda, int give_log) {
  if (lambda == 0) return (x == 0) ? d1(give_log) : d0(give_log);
  if (!std::isfinite(lambda)) return d0(give_log);
  if (x < 0) return d0(give_log);
  if (x <= lambda * DBL_MIN) return dexp_(-lambda, give_log);
  if (lambda < x * DBL_MIN) {
    if (!std::isfinite(x)) return d0(give_log);
    return dexp_(-lambda + x * std::log(lambda) - std::lgamma(x + 1), give_log);
  }
  double f = k2Pi * x, e = -cco_stirlerr(x) - cco_bd0(x, lambda);
  return give_log ? -0.5 * std::log(f) + e : std::exp(e) / std::sqrt(f);
}

double cco_dpois(double x, double lambda, int give_log) {
  if (std::isnan(x) || std::isnan(lambda)) return x + lambda;
  if (lambda < 0) return kNaN;
  if (nonint(x)) return d0(give_log); /* R also raises a warning */
  if (x < 0 || !std::isfinite(x)) return d0(give_log);
  x = forceint(x);
  return cco_dpois_raw(x, lambda, give_log);
}

double cco_dbinom_raw(double x, double n, double p, double q, int give_log) {
  double lf, lc;
  if (p == 0) return (x == 0) ? d1(give_log) : d0(give_log);
  if (q == 0) return (x == n) ? d1(give_log) : d0(give_log);
  if (x == 0) {
    if (n == 0) return d1(give_log);
    lc = (p < 0.1) ? -cco_bd0(n, n * q) - n * p : n * std::log(q);
    return dexp_(lc, give_log);
  }
  if (x == n) {
    lc = (q < 0.1) ? -cco_bd0(n, n * p) - n * q : n * std::log(p);
    return dexp_(lc, give_log);
  }
  if (x < 0 || x > n) return d0(give_log);
  lc = cco_stirlerr(n) - cco_stirlerr(x) - cco_stirlerr(n - x) - cco_bd0(x, n * p) - cco_bd0(n - x, n * q);
  lf = kLn2Pi + std::log(x) + std::log1p(-x / n);
  return dexp_(lc - 0.5 * lf, give_log);
}

double cco_dbinom(double x, double n, double p, int give_log) {
  if (std::isnan(x) || std::isnan(n) || std::isnan(p)) return x + n + p;
  if (p < 0 || p > 1 || n < 0 || nonint(n)) return kNaN;
  if (nonint(x)) return d0(give_log); /* R also raises a warning */
  if (x < 0 || !std::isfinite(x)) return d0(give_log);
  n = forceint(n);
  x = forceint(x);
  return cco_dbinom_raw(x, n, p, 1 - p, give_log);
}

double cco_pgamma(double x, double alph, double scale, int lower_tail, int log_p) {
  if (std::isnan(x) || std::isnan(alph) || std::isnan(scale)) return x + alph + scale;
  if (alph < 0. || scale <= 0.) return kNaN;
  x /= scale;
  if (std::isnan(x)) return x;
  if (alph == 0.) return (x <= 0) ? (lower_tail ? d0(log_p) : d1(log_p)) : (lower_tail ? d1(log_p) : d0(log_p));
  return pgamma_raw(x, alph, lower_tail, log_p);
}

double cco_pnorm(double x, int lower_tail, int log_p) { return pnorm_std(x, lower_tail, log_p); }

double cco_pweibull(double x, double shape, double scale, int lower_tail, int log_p) {
  if (std::isnan(x) || std::isnan(shape) || std::isnan(scale)) return x + shape + scale;
  if (shape <= 0 || scale <= 0) return kNaN;
  if (x <= 0) return lower_tail ? d0(log_p) : d1(log_p);
  x = -std::pow(x / scale, shape);
  return lower_tail ? (log_p ? log1_exp(x) : -std::expm1(x)) : dexp_(x, log_p);
}

double cco_dnorm(double x, double mu, double sigma, int give_log) {
  if (std::isnan(x) || std::isnan(mu) || std::isnan(sigma)) return x + mu + sigma;
  if (!std::isfinite(sigma)) return d0(give_log);
  if (!std::isfinite(x) && mu == x) return kNaN;
  if (sigma <= 0) {
    if (sigma < 0) return kNaN;
    return (x == mu) ? kInf : d0(give_log);
  }
  x = (x - mu) / sigma;
  if (!std::isfinite(x)) return d0(give_log);
  x = std::fabs(x);
  if (x >= 2 * std::sqrt(DBL_MAX)) return d0(give_log);
  if (give_log) return -(kLnSqrt2Pi + 0.5 * x * x + std::log(sigma));
  return k1Sqrt2Pi * std::exp(-0.5 * x * x) / sigma;
}

double cco_dunif(double x, double a, double b, int give_log) {
  if (std::isnan(x) || std::isnan(a) || std::isnan(b)) return x + a + b;
  if (b <= a) return kNaN;
  if (a <= x && x <= b) return give_log ? -std::log(b - a) : 1. / (b - a);
  return d0(give_log);
}

static double lbeta_(double a, double b) { return std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b); }

double cco_dbeta(double x, double a, double b, int give_log) {
  if (std::isnan(x) || std::isnan(a) || std::isnan(b)) return x + a + b;
  if (a < 0 || b < 0) return kNaN;
  if (x < 0 || x > 1) return d0(give_log);
  if (a == 0 || b == 0 || !std::isfinite(a) || !std::isfinite(b)) {
    if (a == 0 && b == 0) return (x == 0 || x == 1) ? kInf : d0(give_log);
    if (a == 0 || a / b == 0) return (x == 0) ? kInf : d0(give_log);
    if (b == 0 || b / a == 0) return (x == 1) ? kInf : d0(give_log);
    return (x == 0.5) ? kInf : d0(give_log);
  }
  if (x == 0) {
    if (a > 1) return d0(give_log);
    if (a < 1) return kInf;
    return give_log ? std::log(b) : b;
  }
  if (x == 1) {
    if (b > 1) return d0(give_log);
    if (b < 1) return kInf;
    return give_log ? std::log(a) : a;
  }
  double lval;
  if (a <= 2 || b <= 2) lval = (a - 1) * std::log(x) + (b - 1) * std::log1p(-x) - lbeta_(a, b);
  else lval = std::log(a + b - 1) + cco_dbinom_raw(a - 1, a + b - 2, x, 1 - x, 1);
  return dexp_(lval, give_log);
}

double cco_dgamma(double x, double shape, double scale, int give_log) {
  if (std::isnan(x) || std::isnan(shape) || std::isnan(scale)) return x + shape + scale;
  if (shape < 0 || scale <= 0) return kNaN;
  if (x < 0) return d0(give_log);
  if (shape == 0) return (x == 0) ? kInf : d0(give_log);
  if (x == 0) {
    if (shape < 1) return kInf;
    if (shape > 1) return d0(give_log);
    return give_log ? -std::log(scale) : 1 / scale;
  }
  double pr;
  if (shape < 1) {
    pr = cco_dpois_raw(shape, x / scale, give_log);
    return give_log ? pr + (std::isfinite(shape / x) ? std::log(shape / x) : std::log(shape) - std::log(x)) : pr * shape / x;
  }
  pr = cco_dpois_raw(shape - 1, x / scale, give_log);
  return give_log ? pr - std::log(scale) : pr / scale;
}

}  // extern "C"
