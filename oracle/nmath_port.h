/*
 * oracle/nmath_port.h - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
 *
 * CPU restatement of the subset of R's nmath library that the reference likelihood and the
 * generated prior call (R::pgamma, R::dpois, R::dbinom, R::pweibull, R::dnorm, R::dunif, R::dbeta):
 *   /root/reference/src/natcubicspline_loglike_binomial.cpp:79,216,248,269,337
 *   /root/reference/src/natcubicspline_loglike_logit.cpp:337
 *   /root/reference/R/Agecpp2strings.R:73-148
 * R's nmath is a third-party dependency that is NOT vendored under /root/reference (DESCRIPTION
 * only says "R >= 2.10"; the golden fit was written by R 4.0.3).  The algorithms below restate the
 * published R 4.0.x nmath algorithms (Loader's saddle-point dbinom/dpois with stirlerr/bd0, Welinder's
 * pgamma_raw); they are pinned against the 60 000 stored reference values (tests/golden) and against
 * independent mpmath evaluations (tests/test_oracle_nmath.py).
 */
#ifndef CCO_NMATH_PORT_H
#define CCO_NMATH_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

double cco_stirlerr(double n);
double cco_bd0(double x, double np);
double cco_dpois_raw(double x, double lambda, int give_log);
double cco_dpois(double x, double lambda, int give_log);
double cco_dbinom_raw(double x, double n, double p, double q, int give_log);
double cco_dbinom(double x, double n, double p, int give_log);
double cco_pgamma(double x, double alph, double scale, int lower_tail, int log_p);
double cco_pweibull(double x, double shape, double scale, int lower_tail, int log_p);
double cco_pnorm(double x, int lower_tail, int log_p); /* standard normal, R's pnorm_both (Cody 1969) */
double cco_dnorm(double x, double mu, double sigma, int give_log);
double cco_dunif(double x, double a, double b, int give_log);
double cco_dbeta(double x, double a, double b, int give_log);
double cco_dgamma(double x, double shape, double scale, int give_log);
double cco_lgamma1p(double a);
double cco_log1pmx(double x);

#ifdef __cplusplus
}
#endif
#endif
