/*
 * ccb200_shim.c - the .Call layer an R installation of COVIDCurve adds to reach libcovidcurve_b200.so.
 *
 * R and its headers are absent from this repository's build image, so the file is not linked here; it is
 * type-checked against minimal declarations of the R API (tests/r_stubs/, tests/test_abi.py::test_r_shim_compiles).  It is the
 * thin binding a maintainer drops into COVIDCurve/src/ next to RcppExports.cpp.  Every entry point only unpacks R vectors
 * into the plain-pointer C ABI of include/covidcurve_b200.h; no logic lives here.
 *
 * Replaces, on the R side:
 *   .Call("_COVIDCurve_natcubspline_loglike_binomial" | "..._logit", params, param_i, data, misc)   src/RcppExports.cpp:10-35
 *   drjacoby::run_mcmc(data, df_params, misc, loglike, logprior, ...)                               R/main.R:179-193
 *
 * Build inside the R package:  PKG_CPPFLAGS = -I<repo>/include   PKG_LIBS = -L<repo>/covidcurve_b200 -lcovidcurve_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "covidcurve_b200.h"

static void ccb_fail(const char* where) { Rf_error("%s: %s", where, cc_last_error()); }

static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  Rf_error("missing element '%s'", name);
  return R_NilValue;
}
static SEXP list_get_opt(SEXP lst, const char* name) { /* R_NilValue when absent */
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  if (Rf_isNull(names)) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(lst); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}
static int as_int(SEXP x) { return Rf_asInteger(x); }
static int opt_int(SEXP lst, const char* name, int dflt) {
  SEXP v = list_get_opt(lst, name);
  return Rf_isNull(v) ? dflt : Rf_asInteger(v);
}

static void model_finalizer(SEXP ptr) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(ptr);
  if (m) cc_model_destroy(m);
  R_ClearExternalPtr(ptr);
}

/*
 * CCB_model_create(desc): `desc` is a named list built by R/ccb200.R from exactly the objects run_IFRmodel_age already
 * builds - data_list (R/main.R:136-152), misc_list (:158-166), IFRmodel$paramdf (:172) - plus the role map and prior
 * table that R/Agecpp2strings.R otherwise bakes into C++ source strings.  Integer vectors are 0-based row indices.
 */
SEXP CCB_model_create(SEXP desc, SEXP device) {
  cc_model_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = CC_ABI_VERSION;
  d.n_params = as_int(list_get(desc, "n_params"));
  d.n_strata = as_int(list_get(desc, "n_strata"));
  d.n_knots = as_int(list_get(desc, "n_knots"));
  d.days_obsd = as_int(list_get(desc, "days_obsd"));
  d.n_sero_obs = as_int(list_get(desc, "n_sero_obs"));
  d.max_seroday_obsd = as_int(list_get(desc, "max_seroday_obsd"));
  d.rcensor_day = as_int(list_get(desc, "rcensor_day"));
  d.account_serorev = as_int(list_get(desc, "account_serorev"));
  d.binomial_likelihood = as_int(list_get(desc, "binomial_likelihood"));
  d.reparam_ifr = as_int(list_get(desc, "reparam_ifr"));
  d.reparam_knots = as_int(list_get(desc, "reparam_knots"));
  d.reparam_infxn = as_int(list_get(desc, "reparam_infxn"));
  d.demog = REAL(list_get(desc, "demog"));
  d.obs_deaths = REAL(list_get(desc, "obs_deaths"));
  d.prop_strata_obs_deaths = REAL(list_get(desc, "prop_strata_obs_deaths"));
  d.sero_a = REAL(list_get(desc, "sero_a"));
  d.sero_b = REAL(list_get(desc, "sero_b"));
  d.sero_survey_start = INTEGER(list_get(desc, "sero_survey_start"));
  d.sero_survey_end = INTEGER(list_get(desc, "sero_survey_end"));
  d.idx_sens = as_int(list_get(desc, "idx_sens"));
  d.idx_spec = as_int(list_get(desc, "idx_spec"));
  d.idx_sero_con_rate = as_int(list_get(desc, "idx_sero_con_rate"));
  d.idx_sero_rev_shape = as_int(list_get(desc, "idx_sero_rev_shape"));
  d.idx_sero_rev_scale = as_int(list_get(desc, "idx_sero_rev_scale"));
  d.idx_mod = as_int(list_get(desc, "idx_mod"));
  d.idx_sod = as_int(list_get(desc, "idx_sod"));
  d.idx_knots = INTEGER(list_get(desc, "idx_knots"));
  d.idx_infxn = INTEGER(list_get(desc, "idx_infxn"));
  d.idx_ifr = INTEGER(list_get(desc, "idx_ifr"));
  d.idx_noise = INTEGER(list_get(desc, "idx_noise"));
  d.idx_rel_knot = as_int(list_get(desc, "idx_rel_knot"));
  d.idx_rel_infxn = as_int(list_get(desc, "idx_rel_infxn"));
  d.idx_max_ma = as_int(list_get(desc, "idx_max_ma"));
  d.pmin = REAL(list_get(desc, "pmin"));
  d.pmax = REAL(list_get(desc, "pmax"));
  d.prior_family = INTEGER(list_get(desc, "prior_family"));
  d.prior_p1 = REAL(list_get(desc, "prior_p1"));
  d.prior_p2 = REAL(list_get(desc, "prior_p2"));
  SEXP jr = list_get(desc, "jac_rows"), jc = list_get(desc, "jac_counts");
  for (int j = 0; j < 3; j++) {
    d.jac_rows[j] = INTEGER(jr)[j];
    d.jac_counts[j] = REAL(jc)[j];
  }
  cc_model* m = NULL;
  if (cc_model_create(&d, as_int(device), &m) != CC_OK) ccb_fail("cc_model_create");
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, model_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* scalar evaluation on an existing model handle: list(LogLik = ) like R/RcppExports.R:4-10 */
SEXP CCB_loglike(SEXP model, SEXP params, SEXP param_i) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(model);
  double out = NA_REAL;
  if (cc_loglike(m, REAL(params), as_int(param_i), &out) != CC_OK) ccb_fail("cc_loglike");
  SEXP ret = PROTECT(Rf_allocVector(VECSXP, 1)), nm = PROTECT(Rf_allocVector(STRSXP, 1));
  SET_VECTOR_ELT(ret, 0, Rf_ScalarReal(out));
  SET_STRING_ELT(nm, 0, Rf_mkChar("LogLik"));
  Rf_setAttrib(ret, R_NamesSymbol, nm);
  UNPROTECT(2);
  return ret;
}

/*
 * The reference's own 4-argument call shape, (params, param_i, data, misc) - src/RcppExports.cpp:10-35, R/RcppExports.R:4-10,
 * and the signature of the drjacoby loglike slot (R/Agecpp2strings.R:389): a NAMED numeric vector plus the data and misc
 * lists of R/main.R:136-166.  tests/testthat/test-check_natcub_cpp.R:114-150 binds to these unchanged.  `binomial` < 0:
 * decide from the data list (obs_serologypos present -> binomial, else logit-normal).
 */
static double call_shape_loglike(SEXP params, SEXP param_i, SEXP data, SEXP misc, int binomial) {
  SEXP names = Rf_getAttrib(params, R_NamesSymbol);
  if (Rf_isNull(names)) Rf_error("params must be a named numeric vector");
  const int d = Rf_length(params);
  const char** nm = (const char**)R_alloc((size_t)d, sizeof(const char*));
  for (int i = 0; i < d; i++) nm[i] = CHAR(STRING_ELT(names, i));
  if (binomial < 0) binomial = Rf_isNull(list_get_opt(data, "obs_serologypos")) ? 0 : 1;
  int np = 0;
  SEXP p_par = PROTECT(Rf_coerceVector(params, REALSXP)); np++;
  SEXP obs = PROTECT(Rf_coerceVector(list_get(data, "obs_deaths"), REALSXP)); np++;
  SEXP prop = PROTECT(Rf_coerceVector(list_get(data, "prop_strata_obs_deaths"), REALSXP)); np++;
  SEXP sa = PROTECT(Rf_coerceVector(list_get(data, binomial ? "obs_serologypos" : "obs_serologymu"), REALSXP)); np++;
  SEXP sb = PROTECT(Rf_coerceVector(list_get(data, binomial ? "obs_serologyn" : "obs_serologyse"), REALSXP)); np++;
  SEXP demog = PROTECT(Rf_coerceVector(list_get(misc, "demog"), REALSXP)); np++;
  SEXP st = PROTECT(Rf_coerceVector(list_get(misc, "sero_survey_start"), INTSXP)); np++;
  SEXP en = PROTECT(Rf_coerceVector(list_get(misc, "sero_survey_end"), INTSXP)); np++;
  cc_call_args a;
  memset(&a, 0, sizeof a);
  a.n_params = d;
  a.names = nm;
  a.obs_deaths = REAL(obs);
  a.prop_strata_obs_deaths = REAL(prop);
  a.sero_a = REAL(sa);
  a.sero_b = REAL(sb);
  a.demog = REAL(demog);
  a.n_strata = Rf_length(demog);
  a.n_knots = as_int(list_get(misc, "n_knots"));
  a.rcensor_day = as_int(list_get(misc, "rcensor_day"));
  a.days_obsd = as_int(list_get(misc, "days_obsd"));
  a.n_sero_obs = as_int(list_get(misc, "n_sero_obs"));
  a.sero_survey_start = INTEGER(st);
  a.sero_survey_end = INTEGER(en);
  a.max_seroday_obsd = as_int(list_get(misc, "max_seroday_obsd"));
  a.account_serorev = Rf_asLogical(list_get(misc, "account_serorev")) == TRUE;
  a.binomial_likelihood = binomial;
  if (Rf_length(obs) != a.days_obsd || Rf_length(prop) != a.n_strata || Rf_length(sa) != a.n_sero_obs * a.n_strata ||
      Rf_length(sb) != a.n_sero_obs * a.n_strata || Rf_length(st) != a.n_sero_obs || Rf_length(en) != a.n_sero_obs)
    Rf_error("data / misc vectors do not have the lengths misc announces");
  double out = NA_REAL;
  const int rc = cc_loglike_call(&a, REAL(p_par), as_int(param_i), opt_int(misc, "ccb200_device", 0), &out);
  UNPROTECT(np);
  if (rc != CC_OK) ccb_fail("cc_loglike_call");
  return out;
}
static SEXP loglik_list(double v) {
  SEXP ret = PROTECT(Rf_allocVector(VECSXP, 1)), nm = PROTECT(Rf_allocVector(STRSXP, 1));
  SET_VECTOR_ELT(ret, 0, Rf_ScalarReal(v));
  SET_STRING_ELT(nm, 0, Rf_mkChar("LogLik"));
  Rf_setAttrib(ret, R_NamesSymbol, nm);
  UNPROTECT(2);
  return ret;
}
/* drop-in replacements of the two registered twins: same symbol names, same arity, same return value list(LogLik = ) */
SEXP _COVIDCurve_natcubspline_loglike_binomial(SEXP params, SEXP param_i, SEXP data, SEXP misc) {
  return loglik_list(call_shape_loglike(params, param_i, data, misc, 1));
}
SEXP _COVIDCurve_natcubspline_loglike_logit(SEXP params, SEXP param_i, SEXP data, SEXP misc) {
  return loglik_list(call_shape_loglike(params, param_i, data, misc, 0));
}
/* the drjacoby user-function shape: SEXP loglike(params, param_i, data, misc) returning a double (Agecpp2strings.R:389-401) */
SEXP CC_loglike(SEXP params, SEXP param_i, SEXP data, SEXP misc) {
  return Rf_ScalarReal(call_shape_loglike(params, param_i, data, misc, -1));
}

/* batched evaluation: theta is an n x d numeric matrix (one parameter vector per ROW, columns in df_params order); R's
   column-major storage is exactly the structure-of-arrays [d][n] layout of cc_loglike_batch, so no copy is made */
SEXP CCB_loglike_batch(SEXP model, SEXP theta) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(model);
  const int64_t n = Rf_nrows(theta);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  if (cc_loglike_batch(m, REAL(theta), n, n, REAL(out), CC_MEM_HOST, NULL) != CC_OK) ccb_fail("cc_loglike_batch");
  UNPROTECT(1);
  return out;
}
SEXP CCB_logprior_batch(SEXP model, SEXP theta) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(model);
  const int64_t n = Rf_nrows(theta);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  if (cc_logprior_batch(m, REAL(theta), n, n, REAL(out), CC_MEM_HOST, NULL) != CC_OK) ccb_fail("cc_logprior_batch");
  UNPROTECT(1);
  return out;
}

/*
 * CCB_mcmc_run(model, opts): the replacement of drjacoby::run_mcmc at R/main.R:179.  Returns a list
 *   theta [rows x d x rungs x chains] (R array), loglikelihood / logprior [rows x rungs x chains], iteration [rows],
 *   beta [rungs], mc_accept [(rungs-1) x chains]
 * which R/ccb200.R reshapes into the `drjacoby_output` data frame (chain, rung, iteration, stage, logprior, loglikelihood, ...).
 */
SEXP CCB_mcmc_run(SEXP model, SEXP opts) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(model);
  cc_mcmc_opts o;
  memset(&o, 0, sizeof o);
  o.burnin = as_int(list_get(opts, "burnin"));
  o.samples = as_int(list_get(opts, "samples"));
  o.rungs = as_int(list_get(opts, "rungs"));
  o.n_chains = as_int(list_get(opts, "chains"));
  o.chain_offset = 0;
  o.coupling_on = as_int(list_get(opts, "coupling_on"));
  o.GTI_pow = Rf_asReal(list_get(opts, "GTI_pow"));
  SEXP bm = list_get(opts, "beta_manual");
  o.beta_manual = Rf_isNull(bm) ? NULL : REAL(bm);
  o.seed = (uint64_t)Rf_asReal(list_get(opts, "seed"));
  /* rung orientation: every consumer in COVIDCurve reads the posterior from "rung1" (R/summarize_plot.R:68,164,276,591,687;
     vignettes/runningmodel.Rmd:238-299), so by default rung 1 is the cold chain (beta = 1) and the powers descend to 0 at
     rung `rungs`; opts$cold_rung_last = TRUE gives the ascending ladder instead.  rung_details (returned `beta`) always
     states which power each rung label carries. */
  o.cold_rung_last = opt_int(opts, "cold_rung_last", 0);
  o.record_rungs = opt_int(opts, "record_rungs", 1);
  o.thin = opt_int(opts, "thin", 1);
  o.bw_per_particle = opt_int(opts, "bw_per_particle", 0);
  o.no_adapt_in_sampling = opt_int(opts, "no_adapt_in_sampling", 0);
  o.init = NULL;
  o.init_default = REAL(list_get(opts, "init"));
  cc_mcmc* s = NULL;
  if (cc_mcmc_create(m, &o, &s) != CC_OK) ccb_fail("cc_mcmc_create");
  if (cc_mcmc_advance(s, o.burnin + o.samples - 1) != CC_OK) { cc_mcmc_destroy(s); ccb_fail("cc_mcmc_advance"); }
  const R_xlen_t rows = (R_xlen_t)cc_mcmc_rows(s), d = Rf_length(list_get(opts, "init")), R = o.record_rungs ? o.rungs : 1, C = o.n_chains;
  SEXP theta = PROTECT(Rf_allocVector(REALSXP, rows * d * R * C)), ll = PROTECT(Rf_allocVector(REALSXP, rows * R * C)),
       lp = PROTECT(Rf_allocVector(REALSXP, rows * R * C)), it = PROTECT(Rf_allocVector(INTSXP, rows)),
       beta = PROTECT(Rf_allocVector(REALSXP, o.rungs)), acc = PROTECT(Rf_allocVector(REALSXP, (o.rungs > 1 ? o.rungs - 1 : 1) * C));
  /* the C layout [chain][rung][row][d] read as column-major is a d x rows x rungs x chains array */
  if (cc_mcmc_fetch(s, REAL(theta), REAL(ll), REAL(lp), INTEGER(it)) != CC_OK ||
      cc_mcmc_diagnostics(s, REAL(beta), REAL(acc), NULL, NULL) != CC_OK) {
    cc_mcmc_destroy(s);
    ccb_fail("cc_mcmc_fetch");
  }
  cc_mcmc_destroy(s);
  const char* nms[] = {"theta", "loglikelihood", "logprior", "iteration", "beta", "mc_accept"};
  SEXP vals[] = {theta, ll, lp, it, beta, acc};
  SEXP ret = PROTECT(Rf_allocVector(VECSXP, 6)), nm = PROTECT(Rf_allocVector(STRSXP, 6));
  for (int i = 0; i < 6; i++) {
    SET_VECTOR_ELT(ret, i, vals[i]);
    SET_STRING_ELT(nm, i, Rf_mkChar(nms[i]));
  }
  Rf_setAttrib(ret, R_NamesSymbol, nm);
  UNPROTECT(8);
  return ret;
}

/*
 * Posterior summaries on the device (the table of get_cred_intervals / get_overall_IFR_cred_intervals, R/summarize_plot.R:136-152):
 * `draws` is a numeric array with dim c(n_params, n_iter, n_chains) - aperm() of the sampling rows of one rung - whose column-major
 * storage is the [chain][iteration][param] layout of cc_summarize_draws.  Returns a CC_SUMM_FIELDS x n_params x groups array
 * (groups = n_chains when by_chain, else 1): rows min, LCI, median, mean, UCI, max, ESS, GewekeZ, GewekeP, var, spec0, n.
 */
SEXP CCB_summarize_draws(SEXP draws, SEXP by_chain, SEXP device) {
  SEXP dim = Rf_getAttrib(draws, R_DimSymbol);
  if (Rf_isNull(dim) || Rf_length(dim) != 3) Rf_error("draws must be an array with dim c(n_params, n_iter, n_chains)");
  const int64_t P = INTEGER(dim)[0], n = INTEGER(dim)[1], C = INTEGER(dim)[2];
  const int bc = Rf_asLogical(by_chain) ? 1 : 0;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)CC_SUMM_FIELDS * P * (bc ? C : 1)));
  if (cc_summarize_draws(as_int(device), REAL(draws), C, n, (int32_t)P, P * n, P, bc, CC_MEM_HOST, NULL, REAL(out)) != CC_OK)
    ccb_fail("cc_summarize_draws");
  UNPROTECT(1);
  return out;
}

/* coda::gelman.diag ingredients + drjacoby's rhat from the same array: CC_GELMAN_FIELDS x n_params matrix (rows rhat, psrf, W, B,
   W_df, df_adj, R2_fixed, R2_random); autoburnin = TRUE drops the first half of every chain like coda does */
SEXP CCB_gelman_draws(SEXP draws, SEXP autoburnin, SEXP device) {
  SEXP dim = Rf_getAttrib(draws, R_DimSymbol);
  if (Rf_isNull(dim) || Rf_length(dim) != 3) Rf_error("draws must be an array with dim c(n_params, n_iter, n_chains)");
  const int64_t P = INTEGER(dim)[0], n = INTEGER(dim)[1], C = INTEGER(dim)[2];
  const int64_t first = Rf_asLogical(autoburnin) ? n / 2 : 0;
  double* mean = (double*)R_alloc((size_t)(C * P), sizeof(double));
  double* var = (double*)R_alloc((size_t)(C * P), sizeof(double));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)CC_GELMAN_FIELDS * P));
  if (cc_chain_moments(as_int(device), REAL(draws), C, n, (int32_t)P, P * n, P, first, CC_MEM_HOST, NULL, mean, var) != CC_OK)
    ccb_fail("cc_chain_moments");
  if (cc_gelman_from_moments(as_int(device), mean, var, C, n - first, (int32_t)P, CC_MEM_HOST, NULL, REAL(out)) != CC_OK)
    ccb_fail("cc_gelman_from_moments");
  UNPROTECT(1);
  return out;
}

/* expected deaths of posterior draws with the gamma density kernel (posterior_check_infxns_to_death, R/summarize_plot.R:605-641):
   theta is an n x d matrix of stored (lifted) draws; returns an n_strata x days x n array */
SEXP CCB_post_deaths(SEXP model, SEXP theta, SEXP n_strata, SEXP days) {
  cc_model* m = (cc_model*)R_ExternalPtrAddr(model);
  const int64_t n = Rf_nrows(theta);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n * as_int(n_strata) * as_int(days)));
  if (cc_post_deaths_batch(m, REAL(theta), n, n, REAL(out)) != CC_OK) ccb_fail("cc_post_deaths_batch");
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"CCB_model_create", (DL_FUNC)&CCB_model_create, 2},
    {"CCB_loglike", (DL_FUNC)&CCB_loglike, 3},
    {"CCB_loglike_batch", (DL_FUNC)&CCB_loglike_batch, 2},
    {"CCB_logprior_batch", (DL_FUNC)&CCB_logprior_batch, 2},
    {"CCB_mcmc_run", (DL_FUNC)&CCB_mcmc_run, 2},
    {"CCB_summarize_draws", (DL_FUNC)&CCB_summarize_draws, 3},
    {"CCB_gelman_draws", (DL_FUNC)&CCB_gelman_draws, 3},
    {"CCB_post_deaths", (DL_FUNC)&CCB_post_deaths, 4},
    /* the reference's own registration table (src/RcppExports.cpp:37-41): same names, same arity */
    {"_COVIDCurve_natcubspline_loglike_binomial", (DL_FUNC)&_COVIDCurve_natcubspline_loglike_binomial, 4},
    {"_COVIDCurve_natcubspline_loglike_logit", (DL_FUNC)&_COVIDCurve_natcubspline_loglike_logit, 4},
    {"CC_loglike", (DL_FUNC)&CC_loglike, 4},
    {NULL, NULL, 0}};

void R_init_ccb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
