/*
 * oracle/cc_oracle.cpp - TEST INFRASTRUCTURE ONLY (see cc_oracle.h).
 *
 * Loop-for-loop CPU restatement of the reference likelihood and of the generated prior.  It keeps the
 * reference's cost structure on purpose (per-call std::vector allocations, by-name parameter lookup,
 * the A-fold redundant serology convolution) because the same code is timed as the "reference CPU"
 * baseline; it is generalised only in A (strata) and K (knots), exactly like the
 * code that R/Agecpp2strings.R generates.
 */
#include "cc_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "nmath_port.h"

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

namespace {

/* R evaluates the likelihood with IEEE gradual underflow.  A host process that has loaded a library built with
   -ffast-math (several Python extension modules are) runs with MXCSR flush-to-zero / denormals-are-zero set, and
   threads inherit it; tiny gamma-cdf values would then flush to 0 and turn finite log-likelihoods into the in-band
   failure value.  Every oracle entry point therefore clears FTZ/DAZ for its own duration. */
struct IeeeGuard {
#if defined(__x86_64__) || defined(__i386__)
  unsigned saved;
  IeeeGuard() : saved(_mm_getcsr()) { _mm_setcsr(saved & ~0x8040u); }
  ~IeeeGuard() { _mm_setcsr(saved); }
#endif
};

/* Rcpp's params["name"]: linear scan over the names attribute (R/Agecpp2strings.R:261 etc.) */
double by_name(const cco_desc* m, const double* params, const char* name) {
  for (int i = 0; i < m->d; i++)
    if (std::strcmp(m->names[i], name) == 0) return params[i];
  throw std::out_of_range(std::string("parameter not found: ") + name);
}

bool in_list(const char* name, const char* const* list, int n) {
  for (int i = 0; i < n; i++)
    if (std::strcmp(name, list[i]) == 0) return true;
  return false;
}

/* names of `list` members in df_params row order (paramdf[paramdf$name %in% list, ]) */
std::vector<const char*> in_paramdf_order(const cco_desc* m, const char* const* list, int n) {
  std::vector<const char*> out;
  for (int i = 0; i < m->d; i++)
    if (in_list(m->names[i], list, n)) out.push_back(m->names[i]);
  return out;
}

/* R/Agecpp2strings.R:277-355: fill a role vector, optionally as scalar * reference parameter */
void fill_role(const cco_desc* m, const double* params, const char* const* role, int n, int reparam, const char* rel,
               double* dst) {
  if (!reparam) {
    for (int i = 0; i < n; i++) dst[i] = by_name(m, params, role[i]);
    return;
  }
  std::vector<const char*> ordered = in_paramdf_order(m, role, n);
  std::vector<const char*> scalars;
  for (const char* s : ordered)
    if (std::strcmp(s, rel) != 0) scalars.push_back(s);
  size_t counter = 0;
  for (int i = 0; i < n; i++) {
    if (std::strcmp(role[i], rel) == 0) dst[i] = by_name(m, params, rel);
    else dst[i] = by_name(m, params, scalars[counter++]) * by_name(m, params, rel);
  }
}

}  // namespace

extern "C" double cco_loglike(const cco_desc* m, const double* params, int /*param_i*/, cco_trace* tr) {
  IeeeGuard ieee_guard;
  /* ---- generated prologue, R/Agecpp2strings.R:255-368 ---- */
  std::vector<int> demog(m->A);
  for (int a = 0; a < m->A; a++) demog[a] = (int)m->demog[a];
  int stratlen = (int)demog.size();
  int rcensor_day = m->rcensor_day;
  int days_obsd = m->days_obsd;
  int n_knots = m->n_knots;
  int n_sero_obs = m->n_sero_obs;
  std::vector<int> sero_survey_start(m->sero_survey_start, m->sero_survey_start + n_sero_obs);
  std::vector<int> sero_survey_end(m->sero_survey_end, m->sero_survey_end + n_sero_obs);
  int max_seroday_obsd = m->max_seroday_obsd;
  bool account_serorev = m->account_serorev != 0;
  std::vector<double> node_x(n_knots), node_y(n_knots), ma(stratlen), ne(stratlen);

  double sens = by_name(m, params, "sens");
  double spec = by_name(m, params, "spec");
  double sero_con_rate = by_name(m, params, "sero_con_rate");
  double mod = by_name(m, params, m->mod_name);
  double sod = by_name(m, params, m->sod_name);
  double sero_rev_shape = -1.0, sero_rev_scale = -1.0;
  if (account_serorev) {
    sero_rev_shape = by_name(m, params, "sero_rev_shape");
    sero_rev_scale = by_name(m, params, "sero_rev_scale");
  }
  node_x[0] = 1.0;
  fill_role(m, params, m->knot_names, n_knots - 1, m->reparam_knots, m->rel_knot, node_x.data() + 1);
  fill_role(m, params, m->infxn_names, n_knots, m->reparam_infxn, m->rel_infxn, node_y.data());
  fill_role(m, params, m->ifr_names, stratlen, m->reparam_ifr, m->max_ma, ma.data());
  if (stratlen > 1) {
    /* noisevec: Noiseparams in paramdf order -> ne[0..] (R/Agecpp2strings.R:361-365) */
    std::vector<const char*> nz = in_paramdf_order(m, m->noise_names, stratlen);
    for (size_t y = 0; y < nz.size(); y++) ne[y] = by_name(m, params, nz[y]);
  }

  if (tr) { tr->L1 = tr->L2 = tr->L3 = 0.0; tr->status = 0; }

  /* ---- body, src/natcubicspline_loglike_binomial.cpp:58-347 ---- */
  /* attack-rate weights (:60-73; logit.cpp:71 fills 0 instead of 1 when stratlen == 1) */
  if (stratlen > 1) {
    double nedom = 0.0;
    for (int i = 0; i < stratlen; i++) {
      ne[i] = ne[i] * demog[i];
      nedom += ne[i];
    }
    for (int i = 0; i < stratlen; i++) ne[i] = ne[i] / nedom;
  } else {
    for (int i = 0; i < stratlen; i++) ne[i] = m->binomial_likelihood ? 1 : 0;
  }
  if (tr && tr->ne) for (int a = 0; a < stratlen; a++) tr->ne[a] = ne[a];

  /* onset-to-death gamma cdf table (:77-80) */
  std::vector<double> pgmms(days_obsd + 1);
  for (int i = 0; i < days_obsd + 1; i++)
    pgmms[i] = cco_pgamma(i, 1 / std::pow(sod, 2), mod * std::pow(sod, 2), 1, 0);

  const double OVERFLO_DOUBLE = DBL_MAX / 100.0;
  double loglik = -OVERFLO_DOUBLE;

  /* knot order (:88-96) */
  bool nodex_pass = true;
  for (size_t i = 1; i < node_x.size(); i++)
    if (node_x[i] <= node_x[i - 1]) nodex_pass = false;
  if (!nodex_pass) {
    if (tr) tr->status = 1;
    return loglik;
  }

  /* natural cubic spline coefficients (:102-139) */
  std::vector<double> denom(n_knots - 1);
  for (size_t i = 0; i < denom.size(); i++) denom[i] = node_x[i + 1] - node_x[i];
  std::vector<double> z(n_knots - 1);
  z[0] = 0;
  std::vector<double> mm(n_knots - 2);
  for (size_t i = 1; i <= mm.size(); i++)
    mm[i - 1] = (3 / denom[i]) * (node_y[i + 1] - node_y[i]) - (3 / denom[i - 1]) * (node_y[i] - node_y[i - 1]);
  std::vector<double> g(n_knots - 1), k(n_knots - 1);
  g[0] = 1;
  k[0] = 0;
  for (int i = 1; i < n_knots - 1; i++) {
    g[i] = 2 * (node_x[i + 1] - node_x[i - 1]) - denom[i - 1] * k[i - 1];
    k[i] = denom[i] / g[i];
    z[i] = (mm[i - 1] - denom[i - 1] * z[i - 1]) / g[i];
  }
  std::vector<double> sp1(n_knots - 1), sp3(n_knots - 1), sp2(n_knots);
  sp2[n_knots - 1] = 0;
  for (int i = n_knots - 2; i >= 0; i--) {
    sp2[i] = z[i] - k[i] * sp2[i + 1];
    sp1[i] = (node_y[i + 1] - node_y[i]) / denom[i] - (denom[i] * (sp2[i + 1] + 2 * sp2[i])) / 3;
    sp3[i] = (sp2[i + 1] - sp2[i]) / (3 * denom[i]);
  }

  /* spline evaluation with the interval-lag rule, then exp (:142-164) */
  std::vector<double> infxn_spline(days_obsd);
  size_t node_j = 0;
  infxn_spline[0] = node_y[0];
  for (int i = 1; i < days_obsd; i++) {
    infxn_spline[i] = node_y[node_j] + sp1[node_j] * ((i + 1) - node_x[node_j]) +
                      sp2[node_j] * std::pow((i + 1) - node_x[node_j], 2) +
                      sp3[node_j] * std::pow((i + 1) - node_x[node_j], 3);
    if (node_j < node_x.size() - 2) {
      if ((node_x[0] + i) >= node_x[node_j + 1]) node_j++;
    }
  }
  for (int i = 0; i < days_obsd; i++) infxn_spline[i] = std::exp(infxn_spline[i]);
  if (tr && tr->infxn) for (int i = 0; i < days_obsd; i++) tr->infxn[i] = infxn_spline[i];

  /* population check (:170-184) */
  bool popN_pass = true;
  std::vector<double> cum_infxn_check(stratlen);
  for (int i = 0; i < days_obsd; i++)
    for (int a = 0; a < stratlen; a++) cum_infxn_check[a] += ne[a] * infxn_spline[i];
  for (int a = 0; a < stratlen; a++)
    if (cum_infxn_check[a] > demog[a]) popN_pass = false;
  if (!popN_pass) {
    if (tr) tr->status = 2;
    return loglik;
  }

  /* death-delay convolution (:187-193) */
  std::vector<double> auc(days_obsd);
  for (int i = 0; i < days_obsd; i++) {
    for (int j = i + 1; j < days_obsd + 1; j++) {
      int delta = j - i - 1;
      auc[j - 1] += infxn_spline[i] * (pgmms[delta + 1] - pgmms[delta]);
    }
  }
  if (tr && tr->auc) for (int i = 0; i < days_obsd; i++) tr->auc[i] = auc[i];

  /* L1: Poisson on daily deaths (:199-221) */
  double L1 = 0.0, texpd = 0.0;
  std::vector<std::vector<double>> strata_expdailyd(days_obsd, std::vector<double>(stratlen));
  std::vector<double> expd(days_obsd);
  std::vector<int> obsd(days_obsd);
  for (int i = 0; i < days_obsd; i++) obsd[i] = (int)m->obs_deaths[i];
  for (int i = 0; i < days_obsd; i++) {
    for (int a = 0; a < stratlen; a++) {
      strata_expdailyd[i][a] = auc[i] * ne[a] * ma[a];
      expd[i] += strata_expdailyd[i][a];
    }
    if ((i + 1) < rcensor_day) {
      if (obsd[i] != -1) L1 += cco_dpois(obsd[i], expd[i], 1);
    }
    texpd += expd[i];
  }

  /* L2: chained binomial on the age split (:227-251) */
  double L2 = 0.0;
  std::vector<double> strata_expd(stratlen);
  for (int a = 0; a < stratlen; a++)
    for (int i = 0; i < days_obsd; i++)
      if ((i + 1) < rcensor_day) strata_expd[a] += strata_expdailyd[i][a];
  std::vector<double> paobsd(m->prop_strata_obs_deaths, m->prop_strata_obs_deaths + stratlen);
  double paobsd_sum = 0.0;
  for (int a = 0; a < stratlen; a++) paobsd_sum += paobsd[a];
  for (int a = 0; a < stratlen - 1; a++) {
    L2 += cco_dbinom(std::round(strata_expd[a]), std::round(texpd), paobsd[a] / paobsd_sum, 1);
    texpd -= strata_expd[a];
    paobsd_sum -= paobsd[a];
  }

  /* cumulative seroconversion hazard (:257-284) */
  std::vector<double> cum_serocon_hazard(max_seroday_obsd);
  if (account_serorev) {
    std::vector<double> serocon_lookup(max_seroday_obsd);
    for (int d = 0; d < max_seroday_obsd; d++) serocon_lookup[d] = (1 / sero_con_rate) * std::exp(-(d + 1) / sero_con_rate);
    std::vector<double> serorev_lookup(max_seroday_obsd);
    for (int d = 0; d < max_seroday_obsd; d++) serorev_lookup[d] = cco_pweibull(d, sero_rev_shape, sero_rev_scale, 1, 0);
    for (int d = 0; d < max_seroday_obsd; d++) {
      for (int j = 0; j < d + 1; j++) {
        int delta = d - j;
        cum_serocon_hazard[d] += serocon_lookup[j] * (1 - serorev_lookup[delta]);
      }
    }
  } else {
    for (int d = 0; d < max_seroday_obsd; d++) cum_serocon_hazard[d] = 1 - std::exp(-(d + 1) / sero_con_rate);
  }
  if (tr && tr->hazard) for (int d = 0; d < max_seroday_obsd; d++) tr->hazard[d] = cum_serocon_hazard[d];

  /* seroconverted by strata and day (:288-298) - A-fold redundant, as written */
  std::vector<std::vector<double>> sero_con_num_full(max_seroday_obsd, std::vector<double>(stratlen));
  for (int a = 0; a < stratlen; a++) {
    for (int i = 0; i < max_seroday_obsd; i++) {
      for (int j = i + 1; j < max_seroday_obsd + 1; j++) {
        int time_elapsed = j - i - 1;
        sero_con_num_full[j - 1][a] += infxn_spline[i] * ne[a] * cum_serocon_hazard[time_elapsed];
      }
    }
  }

  /* survey-window mean (:301-311) */
  std::vector<std::vector<double>> sero_con_num(n_sero_obs, std::vector<double>(stratlen));
  for (int i = 0; i < n_sero_obs; i++) {
    for (int j = 0; j < stratlen; j++) {
      for (int kk = sero_survey_start[i]; kk <= sero_survey_end[i]; kk++) sero_con_num[i][j] += sero_con_num_full[kk - 1][j];
      sero_con_num[i][j] = sero_con_num[i][j] / (sero_survey_end[i] - sero_survey_start[i] + 1);
    }
  }
  if (tr && tr->sero_num)
    for (int i = 0; i < n_sero_obs; i++)
      for (int j = 0; j < stratlen; j++) tr->sero_num[i * stratlen + j] = sero_con_num[i][j];

  /* L3 (:316-340 binomial; logit.cpp:314-339 logit-normal) */
  double L3 = 0.0;
  const double UNDERFLO_DOUBLE = DBL_MIN / 100.0;
  std::vector<double> a_raw(m->sero_a, m->sero_a + (size_t)n_sero_obs * stratlen);
  std::vector<double> b_raw(m->sero_b, m->sero_b + (size_t)n_sero_obs * stratlen);
  std::vector<std::vector<double>> da(n_sero_obs, std::vector<double>(stratlen)), db(n_sero_obs, std::vector<double>(stratlen));
  int seroiter = 0;
  for (int i = 0; i < n_sero_obs; i++) {
    for (int j = 0; j < stratlen; j++) {
      da[i][j] = a_raw[seroiter];
      db[i][j] = b_raw[seroiter];
      seroiter++;
    }
  }
  for (int i = 0; i < n_sero_obs; i++) {
    for (int j = 0; j < stratlen; j++) {
      if (m->binomial_likelihood) {
        if (da[i][j] != -1 || db[i][j] != -1) {
          double obs_prev = sens * (sero_con_num[i][j] / demog[j]) + (1 - spec) * (1 - (sero_con_num[i][j] / demog[j]));
          L3 += cco_dbinom(da[i][j], db[i][j], obs_prev, 1);
        }
      } else {
        double v = sens * (sero_con_num[i][j] / demog[j]) + (1 - spec) * (1 - (sero_con_num[i][j] / demog[j]));
        v = std::log((v + UNDERFLO_DOUBLE) / (1 - (v + UNDERFLO_DOUBLE)));
        L3 += cco_dnorm(v, da[i][j], db[i][j], 1);
      }
    }
  }

  loglik = L1 + L2 + L3;
  if (tr) { tr->L1 = L1; tr->L2 = L2; tr->L3 = L3; }
  if (!std::isfinite(loglik)) {
    loglik = -OVERFLO_DOUBLE;
    if (tr) tr->status = 3;
  }
  return loglik;
}

/* R/Agecpp2strings.R:66-219.  Sum order follows the generated expression (left to right):
   IFR dunif, knot dunif, infxn dunif, sero (con_rate dnorm, sens dbeta, spec dbeta[, rev shape, rev scale dnorm]),
   noise dnorm, mod dnorm, sod dbeta, then the Jacobian terms. */
extern "C" double cco_logprior(const cco_desc* m, const double* params, int /*param_i*/) {
  IeeeGuard ieee_guard;
  const int A = m->A, K = m->n_knots;
  auto idx = [&](const char* nm) {
    for (int i = 0; i < m->d; i++)
      if (std::strcmp(m->names[i], nm) == 0) return i;
    throw std::out_of_range(std::string("parameter not found: ") + nm);
  };
  double ret = 0.0;
  bool first = true;
  auto add = [&](double v) {
    if (first) { ret = v; first = false; } else ret = ret + v;
  };
  for (const char* nm : in_paramdf_order(m, m->ifr_names, A)) { int i = idx(nm); add(cco_dunif(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  for (const char* nm : in_paramdf_order(m, m->knot_names, K - 1)) { int i = idx(nm); add(cco_dunif(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  for (const char* nm : in_paramdf_order(m, m->infxn_names, K)) { int i = idx(nm); add(cco_dunif(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  { int i = idx("sero_con_rate"); add(cco_dnorm(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  { int i = idx("sens"); add(cco_dbeta(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  { int i = idx("spec"); add(cco_dbeta(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  if (m->account_serorev) {
    { int i = idx("sero_rev_shape"); add(cco_dnorm(params[i], m->dsc1[i], m->dsc2[i], 1)); }
    { int i = idx("sero_rev_scale"); add(cco_dnorm(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  }
  if (A > 1)
    for (const char* nm : in_paramdf_order(m, m->noise_names, A)) { int i = idx(nm); add(cco_dnorm(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  { int i = idx(m->mod_name); add(cco_dnorm(params[i], m->dsc1[i], m->dsc2[i], 1)); }
  const bool fff = !m->reparam_ifr && !m->reparam_knots && !m->reparam_infxn;
  if (!fff) {
    /* with all three flags FALSE the generated statement ends after the mod prior and the sod Beta prior is a
       dead expression statement (R/Agecpp2strings.R:202; confirmed by the stored fit's logprior string) */
    int i = idx(m->sod_name);
    add(cco_dbeta(params[i], m->dsc1[i], m->dsc2[i], 1));
  }
  if (m->reparam_ifr) add((A - 1) * std::log(params[idx(m->max_ma)]));
  if (m->reparam_knots) add((K - 2) * std::log(params[idx(m->rel_knot)]));
  if (m->reparam_infxn) add((K - 1) * std::log(params[idx(m->rel_infxn)]));
  if (!std::isfinite(ret)) ret = -(DBL_MAX / 100.0);
  return ret;
}

template <typename F>
static void run_many(long n, int nthreads, F f) {
  if (nthreads <= 1) { f(0, n); return; }
  std::vector<std::thread> th;
  long per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    long lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back(f, lo, hi);
  }
  for (auto& t : th) t.join();
}

extern "C" void cco_loglike_many(const cco_desc* m, const double* params, long n, double* out, int nthreads) {
  run_many(n, nthreads, [=](long lo, long hi) {
    for (long r = lo; r < hi; r++) out[r] = cco_loglike(m, params + r * m->d, 1, nullptr);
  });
}

extern "C" void cco_logprior_many(const cco_desc* m, const double* params, long n, double* out, int nthreads) {
  run_many(n, nthreads, [=](long lo, long hi) {
    for (long r = lo; r < hi; r++) out[r] = cco_logprior(m, params + r * m->d, 1);
  });
}

/* get_post_deaths of posterior_check_infxns_to_death (R/summarize_plot.R:605-641), loop for loop: `infxns` is the
   unlisted data frame of the per-stratum infection curves (column after column: stratum-major), the delay kernel is the
   gamma DENSITY at integer lags, out[day][stratum] += infxns[i][a] * ifr[a] * gmmlkup[delta]. */
extern "C" void cco_post_deaths(const double* infxns, const double* ifr, int daylen, int stratlen, double mod, double sod, double* out) {
  std::vector<double> gmmlkup(daylen + 1);
  for (int i = 0; i < (daylen + 1); i++) gmmlkup[i] = cco_dgamma(i, 1 / std::pow(sod, 2), mod * std::pow(sod, 2), 0);
  std::vector<std::vector<double>> infxns_strata(daylen, std::vector<double>(stratlen));
  int iter = 0;
  for (int j = 0; j < stratlen; j++)
    for (int i = 0; i < daylen; i++) {
      infxns_strata[i][j] = infxns[iter];
      iter++;
    }
  std::vector<std::vector<double>> exp_death_day_strata(daylen, std::vector<double>(stratlen, 0.0));
  for (int i = 0; i < daylen; i++)
    for (int j = i + 1; j < (daylen + 1); j++) {
      int delta = j - i - 1;
      for (int a = 0; a < stratlen; a++) exp_death_day_strata[j - 1][a] += infxns_strata[i][a] * ifr[a] * gmmlkup[delta];
    }
  for (int i = 0; i < daylen; i++)
    for (int a = 0; a < stratlen; a++) out[(long)i * stratlen + a] = exp_death_day_strata[i][a];
}
