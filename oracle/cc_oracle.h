/*
 * oracle/cc_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of COVIDCurve's hot path, used as the checker in tests/,
 * __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm.  Nothing in the
 * product path (covidcurve_b200/, include/) may include, link or call this.
 *
 * What it restates (all under /root/reference):
 *   src/natcubicspline_loglike_binomial.cpp:58-347   the likelihood body (binomial serology)
 *   src/natcubicspline_loglike_logit.cpp:314-339     the only differing block (logit-normal serology)
 *   R/Agecpp2strings.R:255-368                       parameter-name -> array mapping of the generated loglike
 *   R/Agecpp2strings.R:66-219                        the generated logprior (incl. the FFF "sod prior dropped" quirk)
 *   R/main.R:132-172                                 data_list / misc_list construction (caller side, in Python)
 *   drjacoby::run_mcmc (external, mrc-ide/drjacoby@e83884e; call site R/main.R:179-193):
 *       single-site Metropolis-Hastings + Robbins-Monro + Metropolis coupling, restated from its
 *       published algorithm - PARITY UNPINNED beyond the distributional facts of the stored fit.
 *
 * Parity pin: the likelihood + prior restatement reproduces the (loglikelihood, logprior) columns of the
 * stored reference fit data/covidcurve_modfit.rda (tests/golden/modfit_golden.npz, tests/test_oracle_golden.py).
 */
#ifndef CC_ORACLE_H
#define CC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cco_desc {
  /* df_params (row order = params vector order), R/main.R:172 */
  int d;
  const char* const* names; /* [d] */
  const double* pmin;       /* [d] */
  const double* pmax;       /* [d] */
  const double* dsc1;       /* [d] prior hyper-parameter 1 (paramdf$dsc1) */
  const double* dsc2;       /* [d] prior hyper-parameter 2 (paramdf$dsc2) */
  /* misc_list, R/main.R:158-166 */
  int A;                    /* length(demog) = stratlen */
  const double* demog;      /* [A] popN, converted to int inside like the reference */
  int n_knots;
  int rcensor_day;
  int days_obsd;
  int n_sero_obs;
  const int* sero_survey_start; /* [S] */
  const int* sero_survey_end;   /* [S] */
  int max_seroday_obsd;
  int account_serorev;
  /* data_list, R/main.R:136-152 */
  const double* obs_deaths;             /* [T], -1 = missing */
  const double* prop_strata_obs_deaths; /* [A] */
  const double* sero_a;                 /* [S*A] obs_serologypos | obs_serologymu */
  const double* sero_b;                 /* [S*A] obs_serologyn   | obs_serologyse */
  int binomial_likelihood;
  /* role lists of the R6 model, R/model.R:17-23 (model-list order) */
  const char* const* ifr_names;   /* [A]   IFRparams   */
  const char* const* knot_names;  /* [K-1] Knotparams  */
  const char* const* infxn_names; /* [K]   Infxnparams */
  const char* const* noise_names; /* [A]   Noiseparams */
  const char* mod_name;
  const char* sod_name;
  int reparam_ifr, reparam_knots, reparam_infxn;
  const char* max_ma;    /* may be NULL when !reparam_ifr   */
  const char* rel_knot;  /* may be NULL when !reparam_knots */
  const char* rel_infxn; /* may be NULL when !reparam_infxn */
} cco_desc;

/* optional intermediates (any pointer may be NULL) */
typedef struct cco_trace {
  double L1, L2, L3;
  double* infxn;    /* [T]   exp(spline)                       */
  double* auc;      /* [T]   death-delay convolution            */
  double* ne;       /* [A]   normalised attack-rate weights     */
  double* hazard;   /* [M]   cumulative seroconversion hazard   */
  double* sero_num; /* [S*A] survey-window mean seroconverted   */
  int status;       /* 0 ok, 1 knot order failed, 2 popN check failed, 3 non-finite total */
} cco_trace;

/* the generated loglike(params, param_i, data, misc); param_i is ignored like in the reference */
double cco_loglike(const cco_desc* m, const double* params, int param_i, cco_trace* trace);
/* the generated logprior(params, param_i, misc) */
double cco_logprior(const cco_desc* m, const double* params, int param_i);

/* batch helpers: params row-major [n][d]; nthreads independent std::threads (one block of rows each) */
void cco_loglike_many(const cco_desc* m, const double* params, long n, double* out, int nthreads);
void cco_logprior_many(const cco_desc* m, const double* params, long n, double* out, int nthreads);

/* get_post_deaths (R/summarize_plot.R:605-641): infxns [stratlen][daylen] (stratum-major, as R unlists the data frame), out [daylen][stratlen] */
void cco_post_deaths(const double* infxns, const double* ifr, int daylen, int stratlen, double mod, double sod, double* out);

/* drjacoby-equivalent MCMC (see cc_oracle_mcmc.cpp) */
typedef struct cco_mcmc_opts {
  int burnin, samples, rungs, chains;
  int coupling_on;
  double GTI_pow;
  const double* beta_manual; /* NULL or [rungs] */
  unsigned long long seed;
  int nthreads;              /* chains are farmed over threads like drjacoby's cluster= */
  int cold_rung_last;        /* 1: beta ascending, rung{R} is cold (drjacoby 1.2.0); 0: rung1 cold */
  int init_per_chain;        /* 0: init is [d] shared by all chains; 1: init is [chains][d] */
  int chain_offset;          /* global id of chain 0 (RNG keying) */
} cco_mcmc_opts;

/* out_theta: [chains][rungs][burnin+samples][d]; out_ll/out_lp: [chains][rungs][burnin+samples];
   out_accept: [chains][rungs-1] coupling acceptance counts over sampling (may be NULL); beta_out [rungs] */
int cco_mcmc(const cco_desc* m, const double* init, const cco_mcmc_opts* o, double* out_theta, double* out_ll,
             double* out_lp, double* out_accept, double* beta_out);
/* known-answer access to the restated Philox4x32-10 and to the (normal, uniform) draw of one proposal */
int cco_philox_kat(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);
void cco_draw(unsigned long long seed, unsigned stream, unsigned chain, unsigned id, unsigned iteration, unsigned param,
              double* z, double* u);

#ifdef __cplusplus
}
#endif
#endif
