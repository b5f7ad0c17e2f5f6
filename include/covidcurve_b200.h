/*
 * covidcurve_b200.h - C ABI of the B200-native (sm_100a, FP64 CUDA) COVIDCurve hot path.
 *
 * Drop-in boundary.  In the reference the hot path is reached through drjacoby's user-function slot:
 *
 *     SEXP loglike (Rcpp::NumericVector params, int param_i, Rcpp::List data, Rcpp::List misc)   R/Agecpp2strings.R:389
 *     SEXP logprior(Rcpp::NumericVector params, int param_i, Rcpp::List misc)                    R/Agecpp2strings.R:213
 *     drjacoby::run_mcmc(data, df_params, misc, loglike, logprior, burnin, samples, chains,
 *                        rungs, coupling_on, GTI_pow, beta_manual, ...)                          R/main.R:179-193
 *
 * and, for the static test twins, through
 *     .Call("_COVIDCurve_natcubspline_loglike_binomial" | "..._logit", params, param_i, data, misc)
 *                                                                     src/RcppExports.cpp:10-46, R/RcppExports.R:4-10
 *
 * This header is what an R .Call shim (covidcurve_b200/r/ccb200_shim.c), Python ctypes
 * (covidcurve_b200/_lib.py) or any other FFI binds instead.  Plain pointers and sizes only; every function
 * returns 0 on success or a negative CC_E_* code, with a message available from cc_last_error().
 * Numerical failure is NOT an error: exactly like the reference (binomial.cpp:88-96,170-184,345-347 and
 * Agecpp2strings.R:216) it is encoded in-band as the value -DBL_MAX/100.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device and fails with CC_E_CUDA otherwise.
 */
#ifndef COVIDCURVE_B200_H
#define COVIDCURVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_ABI_VERSION 1

/* compile-time capacity of one model (checked by cc_model_create) */
#define CC_MAX_PARAMS 128
#define CC_MAX_STRATA 32
#define CC_MAX_KNOTS 16
#define CC_MAX_SEROSURVEYS 16
#define CC_MAX_DAYS 1024

/* error codes */
#define CC_OK 0
#define CC_E_INVALID (-1) /* bad argument / descriptor */
#define CC_E_CUDA (-2)    /* CUDA runtime error or no device */
#define CC_E_NOMEM (-3)
#define CC_E_STATE (-4)   /* call out of order (e.g. mcmc step before init) */

/* memory space of caller-owned buffers */
#define CC_MEM_HOST 0
#define CC_MEM_DEVICE 1

/* prior families of the generated logprior (R/Agecpp2strings.R:73-148) */
#define CC_PRIOR_NONE 0
#define CC_PRIOR_DUNIF 1 /* R::dunif(x, p1, p2, true) */
#define CC_PRIOR_DNORM 2 /* R::dnorm(x, p1, p2, true) */
#define CC_PRIOR_DBETA 3 /* R::dbeta(x, p1, p2, true) */

/* The value the reference returns for an invalid proposal: -(DBL_MAX/100) */
#define CC_OVERFLO_LOGLIKE (-1.797693134862315708e+306)

/*
 * Model descriptor = data_list + misc_list + df_params + the name->array mapping and prior literals that
 * R/Agecpp2strings.R bakes into the generated C++ strings.  All pointers are HOST pointers, copied by
 * cc_model_create (the caller may free them afterwards).
 */
typedef struct cc_model_desc {
  int32_t abi_version; /* CC_ABI_VERSION */
  /* sizes */
  int32_t n_params;         /* d = nrow(df_params), R/main.R:172 */
  int32_t n_strata;         /* A = length(misc$demog) */
  int32_t n_knots;          /* K = misc$n_knots (incl. the fixed knot at day 1) */
  int32_t days_obsd;        /* T = misc$days_obsd */
  int32_t n_sero_obs;       /* S = misc$n_sero_obs */
  int32_t max_seroday_obsd; /* M = misc$max_seroday_obsd (<= T) */
  int32_t rcensor_day;      /* misc$rcensor_day (INT_MAX = no censoring) */
  int32_t account_serorev;  /* misc$account_serorev */
  int32_t binomial_likelihood; /* 1: natcubicspline_loglike_binomial.cpp, 0: ..._logit.cpp */
  int32_t reparam_ifr, reparam_knots, reparam_infxn; /* run_IFRmodel_age(reparamIFR=, reparamKnots=, reparamInfxn=) */
  /* misc / data */
  const double* demog;                  /* [A] misc$demog (popN; truncated to int like the reference, binomial.cpp:10) */
  const double* obs_deaths;             /* [T] data$obs_deaths, -1 = missing (truncated to int, binomial.cpp:204) */
  const double* prop_strata_obs_deaths; /* [A] data$prop_strata_obs_deaths */
  const double* sero_a;                 /* [S*A] data$obs_serologypos | data$obs_serologymu, survey-major */
  const double* sero_b;                 /* [S*A] data$obs_serologyn   | data$obs_serologyse */
  const int32_t* sero_survey_start;     /* [S] 1-based inclusive */
  const int32_t* sero_survey_end;       /* [S] 1-based inclusive */
  /* role map: row (0-based, df_params order) that feeds each model quantity (Agecpp2strings.R:261-368) */
  int32_t idx_sens, idx_spec, idx_sero_con_rate;
  int32_t idx_sero_rev_shape, idx_sero_rev_scale; /* -1 when !account_serorev */
  int32_t idx_mod, idx_sod;
  const int32_t* idx_knots; /* [K-1] rows of node_x[1..K-1]; scaled by params[idx_rel_knot] when reparam_knots and row != idx_rel_knot */
  const int32_t* idx_infxn; /* [K]   rows of node_y[0..K-1]; same rule with idx_rel_infxn */
  const int32_t* idx_ifr;   /* [A]   rows of ma[0..A-1];     same rule with idx_max_ma */
  const int32_t* idx_noise; /* [A]   rows of ne[0..A-1] (ignored when A == 1) */
  int32_t idx_rel_knot, idx_rel_infxn, idx_max_ma; /* -1 when the matching reparam flag is 0 */
  /* df_params bounds (drjacoby transform: all four finite/infinite combinations are supported) */
  const double* pmin; /* [d] may be -INFINITY */
  const double* pmax; /* [d] may be +INFINITY */
  /* prior table (literals of the generated logprior) */
  const int32_t* prior_family; /* [d] CC_PRIOR_* */
  const double* prior_p1;      /* [d] paramdf$dsc1 */
  const double* prior_p2;      /* [d] paramdf$dsc2 */
  int32_t jac_rows[3];         /* rows of maxMa, relKnot, relInfxn or -1 */
  double jac_counts[3];        /* number of scalars multiplied by each (A-1, K-2, K-1) */
} cc_model_desc;

typedef struct cc_model cc_model; /* opaque: owns the device copy of the descriptor and all scratch */

const char* cc_last_error(void);
int cc_abi_version(void);
int cc_device_count(void);

/* Create / destroy a model bound to CUDA device `device`. */
int cc_model_create(const cc_model_desc* desc, int device, cc_model** out);
void cc_model_destroy(cc_model* m);

/*
 * Batched drop-in for loglike(params, param_i, data, misc) (R/Agecpp2strings.R:389; param_i is ignored by the
 * reference, binomial.cpp:7, so it is not part of the batch call).
 *   theta : structure-of-arrays [d][ld], element (p, i) at theta[p*ld + i], i < n  (n contiguous -> coalesced)
 *   out   : [n] log-likelihoods
 *   mem   : CC_MEM_HOST or CC_MEM_DEVICE (pointers on `device`).  Host buffers are processed in column chunks on three
 *           streams; page-locked caller memory (cudaHostAlloc / cudaHostRegister) is read and written in place by the copy
 *           engine, pageable memory (malloc, an R vector) is staged through the library's own pinned ring.
 *   stream: cudaStream_t as void* (NULL = the library's own stream); the call is synchronous for CC_MEM_HOST and
 *           stream-ordered (asynchronous) for CC_MEM_DEVICE.  One model handle evaluates one device-pointer batch at a
 *           time: a call on another stream than the previous one is ordered behind it (they share scratch buffers); use
 *           one handle per stream for concurrent batches.
 */
int cc_loglike_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* out, int mem, void* stream);
/* Batched drop-in for logprior(params, param_i, misc) (R/Agecpp2strings.R:213). Same conventions. */
int cc_logprior_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* out, int mem, void* stream);

/* Scalar drop-ins (batch of one, host pointers) for the .Call twins _COVIDCurve_natcubspline_loglike_* . */
int cc_loglike(cc_model* m, const double* params, int param_i, double* out);
int cc_logprior(cc_model* m, const double* params, int param_i, double* out);

/*
 * The reference's own call shape, for the static twins
 *     .Call("_COVIDCurve_natcubspline_loglike_binomial" | "..._logit", params, param_i, data, misc)   src/RcppExports.cpp:10-35
 * which COVIDCurve's unit test drives directly (tests/testthat/test-check_natcub_cpp.R:114-150): a NAMED parameter vector plus
 * the data and misc lists, no model handle.  `cc_call_args` is those two lists and names(params) as plain fields; the role map
 * is the twins' fixed naming (binomial.cpp:21-55, generalised in A and K like the generator, Agecpp2strings.R:255-368):
 * sens, spec, sero_con_rate, sero_rev_shape, sero_rev_scale (read only when account_serorev), mod, sod, x1..x{K-1}, y1..yK,
 * ma1..maA, ne1..neA.  The library keeps the model handles of the most recent distinct (names, data, misc) triples, so a
 * repeated call costs one batch-of-one evaluation.  A name that is missing is an error (Rcpp's params["name"] throws).
 */
typedef struct cc_call_args {
  int32_t n_params;
  const char* const* names;             /* [n_params] names(params) */
  /* data (R/main.R:136-152) */
  const double* obs_deaths;             /* [days_obsd] */
  const double* prop_strata_obs_deaths; /* [n_strata] */
  const double* sero_a;                 /* [n_sero_obs * n_strata] obs_serologypos | obs_serologymu */
  const double* sero_b;                 /* [n_sero_obs * n_strata] obs_serologyn   | obs_serologyse */
  /* misc (R/main.R:158-166) */
  const double* demog;                  /* [n_strata] */
  int32_t n_strata, n_knots, rcensor_day, days_obsd, n_sero_obs;
  const int32_t* sero_survey_start;     /* [n_sero_obs] */
  const int32_t* sero_survey_end;       /* [n_sero_obs] */
  int32_t max_seroday_obsd, account_serorev;
  int32_t binomial_likelihood;          /* 1: natcubspline_loglike_binomial, 0: natcubspline_loglike_logit */
} cc_call_args;
int cc_loglike_call(const cc_call_args* args, const double* params, int param_i, int device, double* out);

/*
 * Intermediates of one evaluation, for the post-processing re-entry points (R/summarize_plot.R:276-884):
 * any output pointer may be NULL.  Host pointers.  Batched over n parameter vectors (SoA like above).
 *   infxn    [n][T]   exp(spline) incidence                 (draw_posterior_infxn_cubic_splines, per stratum = ne[a]*infxn)
 *   ne       [n][A]   normalised attack-rate weights
 *   auc      [n][T]   death-delay convolution (expected deaths per day before IFR weighting)
 *   sero_full[n][T]   sum_i infxn[i]*hazard[j-i] over ALL days_obsd (draw_posterior_sero_curves, x ne[a])
 */
int cc_curves_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* infxn, double* ne, double* auc,
                    double* sero_full);

/*
 * Posterior check of expected deaths (posterior_check_infxns_to_death, R/summarize_plot.R:591-680; the compiled loop at
 * :605-641): per parameter vector (draw) and stratum
 *     deaths[j][a] = sum_{i <= j} (ne[a] * infxn[i]) * ifr[a] * dgamma(j - i; shape = 1/sod^2, scale = mod * sod^2)
 * with the gamma DENSITY (R::dgamma) as the delay kernel, terms added in the reference's order (i = 0, 1, ...), each product
 * rounded.  theta: host, SoA like cc_loglike_batch, the stored (already lifted) posterior values; deaths: host [n][T][A].
 * A draw whose knots are not increasing has no curve: its block is NaN.
 */
int cc_post_deaths_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* deaths);

/* ---------------------------------------------------------------------------------------------------------
 * Posterior summaries of recorded draws ON THE DEVICE (the consumers of the fit object: R/summarize_plot.R:46-58
 * get_gelman_rubin_diagnostic -> coda::gelman.diag, :68-153 get_cred_intervals, :162-264 get_overall_IFR_cred_intervals).
 * `draws` holds element (chain c, iteration i, parameter p) at draws[c * stride_chain + i * stride_iter + p]  (doubles);
 * with mem = CC_MEM_DEVICE it is a device pointer (cc_mcmc_device_draws, or an NCCL all-gathered tensor) and `stream`
 * orders the work after whatever produced it; with CC_MEM_HOST the draws are uploaded first.  Outputs marked HOST are
 * always host arrays (they are small); every call returns after its results are complete.
 * --------------------------------------------------------------------------------------------------------- */
#define CC_SUMM_FIELDS 12
/* fields: 0 min, 1 LCI (2.5 %, type-7 quantile), 2 median, 3 mean, 4 UCI (97.5 %), 5 max, 6 ESS (coda::effectiveSize),
   7 GewekeZ (coda::geweke.diag, frac1 = 0.1, frac2 = 0.5), 8 GewekeP = dnorm(GewekeZ), 9 variance (n - 1), 10 spectral density at
   zero (coda::spectrum0.ar), 11 series length.
   by_chain != 0: one row per (chain, parameter), out HOST [n_chains][n_params][CC_SUMM_FIELDS];
   by_chain == 0: the chains' series concatenated in chain order (dplyr::group_by(param)), out HOST [1][n_params][CC_SUMM_FIELDS]. */
int cc_summarize_draws(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params,
                       int64_t stride_chain, int64_t stride_iter, int32_t by_chain, int mem, void* stream, double* out);
/* per-chain mean and unbiased variance of iterations [first_iter, n_iter): mean, var [n_chains][n_params], in the memory
   space `mem` names (device arrays can be all-gathered over the ranks before cc_gelman_from_moments) */
int cc_chain_moments(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params, int64_t stride_chain,
                     int64_t stride_iter, int64_t first_iter, int mem, void* stream, double* mean, double* var);
#define CC_GELMAN_FIELDS 8
/* from per-chain moments of n_iter_used draws each: out HOST [n_params][CC_GELMAN_FIELDS] =
   0 drjacoby's rhat (sqrt(((1 - 1/n) W + B/n) / W)), 1 coda::gelman.diag point estimate (transform = FALSE, with the
   degrees-of-freedom correction), 2 W, 3 B, 4 df of W, 5 df adjustment, 6 fixed and 7 random part of R^2: coda's upper limit is
   sqrt(f5 * (f6 + qf((1 + confidence) / 2, n_chains - 1, f4) * f7)) - one F quantile per parameter, left to the host */
int cc_gelman_from_moments(int device, const double* mean, const double* var, int64_t n_chains, int64_t n_iter_used,
                           int32_t n_params, int mem, void* stream, double* out);
/* overall IFR per draw (R/summarize_plot.R:162-264): est[c][i] = sum_a draws[.., idx_ifr[a]] * w_a with w_a = popN_a / sum popN
   (idx_noise == NULL: whichstandard = "pop") or noise_a popN_a / sum_b noise_b popN_b ("arpop"); est [n_chains][n_iter] in the
   memory space `mem` names - summarise it with cc_summarize_draws(n_params = 1) */
int cc_overall_ifr_draws(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params, int64_t stride_chain,
                         int64_t stride_iter, const int32_t* idx_ifr, const int32_t* idx_noise, const double* popN, int32_t n_strata,
                         int mem, void* stream, double* est);

/* ---------------------------------------------------------------------------------------------------------
 * GPU-resident Metropolis-coupled MCMC replacing drjacoby::run_mcmc (call site R/main.R:179-193).
 * One particle per (chain, rung); single-site Metropolis-Hastings sweep over the d parameters with
 * Robbins-Monro proposal adaptation, then one pass of adjacent-rung swaps; device RNG (Philox4x32-10 keyed
 * by seed, global chain id, rung, iteration, parameter).  Chains are independent: shard them over GPUs by
 * giving each process a [chain_offset, chain_offset + n_chains) slice - results do not depend on the split.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct cc_mcmc_opts {
  int32_t burnin;        /* drjacoby burnin  (iterations 1..burnin; iteration 1 is the initial state) */
  int32_t samples;       /* drjacoby samples */
  int32_t rungs;
  int32_t n_chains;      /* chains simulated by THIS model handle */
  int32_t chain_offset;  /* global id of the first local chain (RNG keying) */
  int32_t coupling_on;
  double GTI_pow;
  const double* beta_manual; /* NULL or [rungs] */
  uint64_t seed;
  int32_t cold_rung_last; /* 1 (default, drjacoby 1.2.0): beta ascending, rung index R-1 is the cold chain */
  int32_t record_rungs;   /* 0: record only the cold rung, 1: record every rung (drjacoby records all) */
  int32_t thin;           /* record every thin-th iteration (0/1 = all) */
  const double* init;     /* NULL (use `init_default` of every chain) or [n_chains][d] */
  const double* init_default; /* [d] df_params$init */
  int32_t bw_per_particle;      /* 0 (default): proposal bandwidths belong to the rung (temperature); 1: they travel with
                                   the particle through swaps (drjacoby 1.2.0 swaps beta between particles) */
  int32_t no_adapt_in_sampling; /* 0 (default): Robbins-Monro keeps adapting during sampling (drjacoby 1.2.0); 1: burn-in only */
} cc_mcmc_opts;

typedef struct cc_mcmc cc_mcmc; /* opaque run state, device resident */

int cc_mcmc_create(cc_model* m, const cc_mcmc_opts* opts, cc_mcmc** out);
/* advance `n_iter` iterations (sweep + swap each); records draws on the device */
int cc_mcmc_advance(cc_mcmc* s, int32_t n_iter);
/* device time (ms, CUDA events on the library's stream) of the launches of the most recent cc_mcmc_advance */
int cc_mcmc_last_advance_ms(cc_mcmc* s, double* ms);
/* iterations done so far (1 after create: the initial state is iteration 1) */
int64_t cc_mcmc_iterations(const cc_mcmc* s);
/* number of recorded rows per (chain, recorded rung) so far */
int64_t cc_mcmc_rows(const cc_mcmc* s);
/*
 * Copy recorded draws to host buffers: theta [n_chains][n_rec_rungs][rows][d], loglike/logprior
 * [n_chains][n_rec_rungs][rows], iteration [rows] (1-based).  Any pointer may be NULL.
 */
int cc_mcmc_fetch(cc_mcmc* s, double* theta, double* loglike, double* logprior, int32_t* iteration);
/* device pointers of the recorded draws (for NCCL allgather / device-side summaries):
   theta [n_chains][n_rec_rungs][rows_cap][d], loglike/logprior [n_chains][n_rec_rungs][rows_cap]; only the first
   cc_mcmc_rows() rows of each group are valid; *rows_cap = (burnin + samples) / thin */
int cc_mcmc_device_draws(cc_mcmc* s, double** theta, double** loglike, double** logprior, int64_t* rows_cap);
/* diagnostics: beta [rungs]; mc_accept [n_chains][rungs-1] swap acceptance rate over the sampling phase;
   move_accept [n_chains][rungs][d] single-site acceptance counts; bw [n_chains][rungs][d] proposal bandwidths */
int cc_mcmc_diagnostics(cc_mcmc* s, double* beta, double* mc_accept, double* move_accept, double* bw);
/* number of likelihood evaluations performed so far (full-evaluation equivalents = proposals) */
int64_t cc_mcmc_loglike_evals(const cc_mcmc* s);
void cc_mcmc_destroy(cc_mcmc* s);

/* ---------------------------------------------------------------------------------------------------------
 * Device versions of the R nmath routines the path calls (binomial.cpp:79,216,248,269,337; logit.cpp:337;
 * Agecpp2strings.R:73-148), exposed element-wise so that they can be checked against R's values in isolation.
 * Host pointers, n elements each; out[i] = fn(a[i], b[i], c[i]) with log = TRUE for the densities:
 *   CC_FN_PGAMMA   pgamma(x=a, shape=b, scale=c, lower, non-log)     CC_FN_DPOIS  dpois(x=a, lambda=b)
 *   CC_FN_DBINOM   dbinom(x=a, n=b, p=c)   CC_FN_DNORM dnorm(a, mu=b, sd=c)   CC_FN_DBETA dbeta(a, b, c)
 *   CC_FN_DUNIF    dunif(a, b, c)          CC_FN_PWEIBULL pweibull(a, shape=b, scale=c)
 *   CC_FN_STIRLERR stirlerr(a)             CC_FN_BD0   bd0(a, b)
 * --------------------------------------------------------------------------------------------------------- */
#define CC_FN_PGAMMA 1
#define CC_FN_DPOIS 2
#define CC_FN_DBINOM 3
#define CC_FN_DNORM 4
#define CC_FN_DBETA 5
#define CC_FN_DUNIF 6
#define CC_FN_PWEIBULL 7
#define CC_FN_STIRLERR 8
#define CC_FN_BD0 9
#define CC_FN_FASTLOG 10 /* the table-driven log of the Poisson hot loop (csrc/cc_fastmath.cuh), a = argument */
int cc_nmath_batch(int device, int fn, const double* a, const double* b, const double* c, int64_t n, double* out);

/* ---------------------------------------------------------------------------------------------------------
 * Forward simulator: the aggregate form of Agesim_infxn_2_death / sim_seroprev (R/simulators.R:114-263, :1-92).
 * The reference materialises one row per infection; by Poisson thinning every (day, stratum) cell of its aggregated
 * outputs is an independent Poisson count, drawn here directly on the device:
 *   deaths[r][j][a]  ~ Poisson(P_a ifr_a    sum_{t<=j} infections[t] (G(j-t+1) - G(j-t))),  G = gamma(1/s_od^2, m_od s_od^2) cdf
 *   serocon[r][j][a] ~ Poisson(P_a smplfrac sum_{t<=j} infections[t] (H(j-t+1) - H(j-t))),  H = exponential cdf, mean sero_delay_rate
 * with P_a = popN_a rho_a / sum(popN rho).  `deaths`, `serocon`: host arrays [n_rep][T][A] (daily counts; the caller
 * accumulates).  `mean_deaths`, `mean_serocon`: optional host arrays [T][A] of the cell means.
 * Seroreversion (sim_seroprev, R/simulators.R:26-33): with sero_rev_shape > 0 and sero_rev_scale > 0 the time from
 * seroconversion to reversion is Weibull; conversions and reversions are then drawn jointly as Poisson cells over the pairs
 * (conversion day, reversion day) and `serorev` [n_rep][T][A] receives the daily reversions among the sampled people
 * (pass shape = scale = 0 and serorev = NULL for no reversion).
 * Random numbers: Philox keyed by (seed, replicate, cell); parity with R's stream is distributional only.
 * --------------------------------------------------------------------------------------------------------- */
int cc_simulate_infxn_2_death(int device, const double* infections, int32_t T, const double* ifr, const double* rho,
                              const double* popN, int32_t A, double m_od, double s_od, double sero_delay_rate, double smplfrac,
                              double sero_rev_shape, double sero_rev_scale, uint64_t seed, int64_t n_rep, double* deaths,
                              double* serocon, double* serorev, double* mean_deaths, double* mean_serocon);
#define CC_FN_RPOIS 11 /* test hook: one Poisson draw, a = mean, b = replicate id, c = seed */

/* ---------------------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): live FP64 FMA peak of the device, and kernel time of the last batch call.
 * --------------------------------------------------------------------------------------------------------- */
/* Register-resident DFMA chain micro-benchmark; returns achieved TFLOP/s (FMA = 2 flops) in *tflops. */
int cc_fp64_peak_probe(int device, int iters, double* tflops, double* ms);
/* FP64 tensor-core (mma.sync m8n8k4 f64) micro-benchmark with `fma_per_mma` (0, 1, 2, 4, 8) independent DFMAs issued per
   warp-level MMA: tells whether DMMA and DFMA share a pipe on this device (death-delay convolution). */
int cc_dmma_probe(int device, int iters, int fma_per_mma, double* mma_tflops, double* fma_tflops, double* ms);
/* CUDA-event duration (ms) of the most recent cc_loglike_batch kernel on this model, and #kernels it launched. */
int cc_last_kernel_ms(cc_model* m, double* ms, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* COVIDCURVE_B200_H */
