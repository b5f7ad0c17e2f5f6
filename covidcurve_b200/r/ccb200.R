# ccb200.R - R side of the drop-in (not executed in this repository: R is absent from the build image; INTEGRATION.md).
#
# run_IFRmodel_age() keeps its signature and its returned object.  Only the block at R/main.R:179-193
#     mcmcout <- drjacoby::run_mcmc(data = data_list, df_params = df_params, misc = misc_list,
#                                   loglike = loglikfunc, logprior = logpriorfunc, ...)
# becomes
#     mcmcout <- ccb200_run_mcmc(IFRmodel, data_list, misc_list, df_params, binomial_likelihood, account_serorev,
#                                reparamIFR, reparamInfxn, reparamKnots, burnin, samples, chains, rungs,
#                                coupling_on, GTI_pow, beta_manual)
# and the two calls to make_user_Age_loglike()/make_user_Age_logprior() (R/main.R:119-125) are no longer needed: what they
# bake into C++ source strings (parameter-name -> array mapping, prior families and hyper-parameters) is passed as data.

ccb200_descriptor <- function(IFRmodel, data_list, misc_list, binomial_likelihood, account_serorev,
                              reparamIFR, reparamInfxn, reparamKnots) {
  pdf <- IFRmodel$paramdf
  row <- function(nm) match(nm, pdf$name) - 1L                      # 0-based rows of df_params
  in_paramdf_order <- function(role) pdf$name[pdf$name %in% role]
  role_rows <- function(role, reparam, rel) {                       # R/Agecpp2strings.R:277-355
    if (!reparam) return(row(role))
    scalars <- setdiff(in_paramdf_order(role), rel); k <- 0L
    vapply(role, function(nm) if (nm == rel) row(rel) else { k <<- k + 1L; row(scalars[k]) }, integer(1))
  }
  fam <- integer(nrow(pdf))                                          # 0 none, 1 dunif, 2 dnorm, 3 dbeta (Agecpp2strings.R:73-148)
  fam[row(c(IFRmodel$IFRparams, IFRmodel$Knotparams, IFRmodel$Infxnparams)) + 1L] <- 1L
  fam[row(c("sero_con_rate", IFRmodel$Noiseparams, IFRmodel$modparam)) + 1L] <- 2L
  fam[row(c("sens", "spec")) + 1L] <- 3L
  if (account_serorev) fam[row(c("sero_rev_shape", "sero_rev_scale")) + 1L] <- 2L
  # with all three flags FALSE the generated prior drops the sod Beta term (Agecpp2strings.R:202; SURVEY.md 8a item 4)
  fam[row(IFRmodel$sodparam) + 1L] <- if (reparamIFR || reparamKnots || reparamInfxn) 3L else 0L
  sero_a <- if (binomial_likelihood) data_list$obs_serologypos else data_list$obs_serologymu
  sero_b <- if (binomial_likelihood) data_list$obs_serologyn else data_list$obs_serologyse
  A <- length(misc_list$demog); K <- misc_list$n_knots
  list(n_params = nrow(pdf), n_strata = A, n_knots = K, days_obsd = misc_list$days_obsd, n_sero_obs = misc_list$n_sero_obs,
       max_seroday_obsd = misc_list$max_seroday_obsd, rcensor_day = as.integer(misc_list$rcensor_day),
       account_serorev = as.integer(account_serorev), binomial_likelihood = as.integer(binomial_likelihood),
       reparam_ifr = as.integer(reparamIFR), reparam_knots = as.integer(reparamKnots), reparam_infxn = as.integer(reparamInfxn),
       demog = as.numeric(misc_list$demog), obs_deaths = as.numeric(data_list$obs_deaths),
       prop_strata_obs_deaths = as.numeric(data_list$prop_strata_obs_deaths), sero_a = as.numeric(sero_a), sero_b = as.numeric(sero_b),
       sero_survey_start = as.integer(misc_list$sero_survey_start), sero_survey_end = as.integer(misc_list$sero_survey_end),
       idx_sens = row("sens"), idx_spec = row("spec"), idx_sero_con_rate = row("sero_con_rate"),
       idx_sero_rev_shape = if (account_serorev) row("sero_rev_shape") else -1L,
       idx_sero_rev_scale = if (account_serorev) row("sero_rev_scale") else -1L,
       idx_mod = row(IFRmodel$modparam), idx_sod = row(IFRmodel$sodparam),
       idx_knots = role_rows(IFRmodel$Knotparams, reparamKnots, IFRmodel$relKnot),
       idx_infxn = role_rows(IFRmodel$Infxnparams, reparamInfxn, IFRmodel$relInfxn),
       idx_ifr = role_rows(IFRmodel$IFRparams, reparamIFR, IFRmodel$maxMa),
       idx_noise = row(in_paramdf_order(IFRmodel$Noiseparams)),
       idx_rel_knot = if (reparamKnots) row(IFRmodel$relKnot) else -1L,
       idx_rel_infxn = if (reparamInfxn) row(IFRmodel$relInfxn) else -1L,
       idx_max_ma = if (reparamIFR) row(IFRmodel$maxMa) else -1L,
       pmin = as.numeric(pdf$min), pmax = as.numeric(pdf$max), prior_family = fam,
       prior_p1 = as.numeric(pdf$dsc1), prior_p2 = as.numeric(pdf$dsc2),
       jac_rows = c(if (reparamIFR) row(IFRmodel$maxMa) else -1L, if (reparamKnots) row(IFRmodel$relKnot) else -1L,
                    if (reparamInfxn) row(IFRmodel$relInfxn) else -1L),
       jac_counts = c(A - 1, K - 2, K - 1))
}

ccb200_run_mcmc <- function(IFRmodel, data_list, misc_list, df_params, binomial_likelihood, account_serorev,
                            reparamIFR, reparamInfxn, reparamKnots, burnin, samples, chains, rungs,
                            coupling_on, GTI_pow, beta_manual, seed = sample.int(.Machine$integer.max, 1), device = 0L,
                            cold_rung_last = FALSE, thin = 1L) {
  # Rung labels: every consumer in COVIDCurve reads the posterior from "rung1" (R/summarize_plot.R:68,164,276,591,687), so by
  # default rung1 is the cold chain (thermodynamic power 1) and the powers descend to 0 at the last rung; rung_details states
  # the power of every label either way.
  desc <- ccb200_descriptor(IFRmodel, data_list, misc_list, binomial_likelihood, account_serorev, reparamIFR, reparamInfxn, reparamKnots)
  model <- .Call("CCB_model_create", desc, as.integer(device), PACKAGE = "ccb200")
  res <- .Call("CCB_mcmc_run", model,
               list(burnin = as.integer(burnin), samples = as.integer(samples), rungs = as.integer(rungs), chains = as.integer(chains),
                    coupling_on = as.integer(coupling_on), GTI_pow = as.numeric(GTI_pow), beta_manual = beta_manual,
                    seed = as.numeric(seed), init = as.numeric(df_params$init), cold_rung_last = as.integer(cold_rung_last),
                    thin = as.integer(thin)), PACKAGE = "ccb200")
  d <- nrow(df_params); rows <- length(res$iteration)
  theta <- array(res$theta, dim = c(d, rows, rungs, chains))
  grid <- expand.grid(iteration = res$iteration, rung = seq_len(rungs), chain = seq_len(chains))
  output <- data.frame(chain = paste0("chain", grid$chain), rung = paste0("rung", grid$rung), iteration = grid$iteration,
                       stage = ifelse(grid$iteration <= burnin, "burnin", "sampling"),
                       logprior = as.numeric(res$logprior), loglikelihood = as.numeric(res$loglikelihood), stringsAsFactors = FALSE)
  pars <- matrix(aperm(theta, c(2, 3, 4, 1)), ncol = d, dimnames = list(NULL, df_params$name))
  output <- cbind(output, as.data.frame(pars))
  cold <- output[output$rung == paste0("rung", which.max(res$beta)) & output$stage == "sampling", ]
  rhat <- if (chains > 1) vapply(df_params$name, function(p) {
    x <- matrix(cold[[p]], ncol = chains); n <- nrow(x)
    W <- mean(apply(x, 2, stats::var)); B <- n * stats::var(colMeans(x)); sqrt(((1 - 1 / n) * W + B / n) / W) }, numeric(1)) else NULL
  ret <- list(output = output,
              diagnostics = list(rhat = rhat, rung_details = data.frame(rung = seq_len(rungs), thermodynamic_power = res$beta),
                                 mc_accept = if (rungs > 1) res$mc_accept else NA),
              parameters = list(data = data_list, df_params = cbind(df_params, trans_type = 3L), loglike = "covidcurve_b200",
                                logprior = "covidcurve_b200", burnin = burnin, samples = samples, rungs = rungs, chains = chains,
                                coupling_on = coupling_on, GTI_pow = GTI_pow, beta_manual = beta_manual))
  class(ret) <- "drjacoby_output"
  ret
}

# A fit written on a GPU box without R (covidcurve_b200/rdata.py: write_fit_rds) -> the `IFRmodel_inf` object every
# function of R/summarize_plot.R takes.  `mcmcout` is already a drjacoby_output; `inputs$IFRmodel` arrives as a named
# list of fields (class "IFRmodel_fields": an R6 object is an environment of closures and cannot be written without R)
# and is rebuilt here through the public setters of R/model.R:192-355, in the order the setters' own checks require.
ccb200_read_fit <- function(path) {
  fit <- readRDS(path)
  f <- fit$inputs$IFRmodel
  if (!inherits(f, "IFRmodel_fields")) return(fit)
  m <- COVIDCurve::make_IFRmodel_age$new()
  m$set_MeanTODparam(f$modparam); m$set_CoefVarOnsetTODparam(f$sodparam)
  m$set_IFRparams(f$IFRparams); m$set_maxMa(f$maxMa)
  m$set_Knotparams(f$Knotparams); m$set_relKnot(f$relKnot)
  m$set_Infxnparams(f$Infxnparams); m$set_relInfxn(f$relInfxn)
  if (!is.null(f$Noiseparams)) m$set_Noiseparams(f$Noiseparams)
  m$set_Serotestparams(f$Serotestparams)
  m$set_data(f$data); m$set_demog(f$demog); m$set_paramdf(f$paramdf)
  if (!is.null(f$rcensor_day)) m$set_rcensor_day(f$rcensor_day)
  if (!is.null(f$IFRdictkey)) m$set_IFRdictkey(f$IFRdictkey)
  fit$inputs$IFRmodel <- m
  fit
}

# Device-side replacements of the summary functions (R/summarize_plot.R:46-58, 68-153): same arguments, same columns.
# The sampling rows of one rung are reshaped to the c(n_params, n_iter, n_chains) array the library reads in place.
ccb200_draws_array <- function(IFRmodel_inf, params, whichrung = "rung1", sampling_only = TRUE) {
  out <- IFRmodel_inf$mcmcout$output
  out <- out[out$rung == whichrung & (!sampling_only | out$stage == "sampling"), ]
  chains <- unique(out$chain)
  n <- nrow(out) / length(chains)
  arr <- vapply(chains, function(ch) t(as.matrix(out[out$chain == ch, params, drop = FALSE])), matrix(0, length(params), n))
  list(draws = array(arr, dim = c(length(params), n, length(chains))), chains = chains)
}

ccb200_get_cred_intervals <- function(IFRmodel_inf, what, whichrung = "rung1", by_chain = TRUE, device = 0L) {
  m <- IFRmodel_inf$inputs$IFRmodel
  params <- switch(what, IFRparams = m$IFRparams, Knotparams = m$Knotparams, Infxnparams = m$Infxnparams,
                   Serotestparams = m$Serotestparams, Noiseparams = m$Noiseparams, DeathDelayparams = c("mod", "sod"),
                   stop("what must be one of IFRparams, Knotparams, Infxnparams, Serotestparams, Noiseparams, DeathDelayparams"))
  d <- ccb200_draws_array(IFRmodel_inf, params, whichrung)
  tab <- .Call("CCB_summarize_draws", d$draws, as.logical(by_chain), as.integer(device), PACKAGE = "ccb200")
  groups <- if (by_chain) length(d$chains) else 1L
  tab <- array(tab, dim = c(12L, length(params), groups))
  ret <- data.frame(param = rep(params, groups), min = as.numeric(tab[1, , ]), LCI = as.numeric(tab[2, , ]), median = as.numeric(tab[3, , ]),
                    mean = as.numeric(tab[4, , ]), UCI = as.numeric(tab[5, , ]), max = as.numeric(tab[6, , ]), ESS = as.numeric(tab[7, , ]),
                    GewekeZ = as.numeric(tab[8, , ]), GewekeP = as.numeric(tab[9, , ]), stringsAsFactors = FALSE)
  if (by_chain) ret <- cbind(chain = rep(d$chains, each = length(params)), ret)
  ret
}

ccb200_get_gelman_rubin_diagnostic <- function(IFRmodel_inf, confidence = 0.95, autoburnin = TRUE, device = 0L) {
  out <- IFRmodel_inf$mcmcout$output
  params <- setdiff(names(out), c("chain", "rung", "iteration", "stage", "logprior", "loglikelihood"))
  chains <- unique(out$chain)                      # like the reference: every row of a chain, no stage or rung filter (:48-51)
  n <- nrow(out) / length(chains)
  arr <- vapply(chains, function(ch) t(as.matrix(out[out$chain == ch, params, drop = FALSE])), matrix(0, length(params), n))
  g <- matrix(.Call("CCB_gelman_draws", array(arr, dim = c(length(params), n, length(chains))), as.logical(autoburnin),
                    as.integer(device), PACKAGE = "ccb200"), nrow = 8L)
  upper <- sqrt(g[6, ] * (g[7, ] + stats::qf((1 + confidence) / 2, length(chains) - 1, g[5, ]) * g[8, ]))
  list(psrf = cbind(`Point est.` = g[2, ], `Upper C.I.` = upper), rhat = stats::setNames(g[1, ], params))
}
