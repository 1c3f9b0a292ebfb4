# The two exported samplers as thin .Call drivers.  This file takes the place of the reference's R/dream_parallel.R and
# R/dream.R (delete those): same function names, same formals in the same order with the same defaults -- the API contract
# of R/dream_parallel.R:155-163 and R/dream.R:126-133 -- with the optional arguments `seed`, `device` and `series_fx` appended.  The
# bodies are gone: everything the reference does per generation runs behind .hb_drive() (hb_drive.R).

# Thin .Call driver replacing the body of HydroBayes::dream_parallel (R/dream_parallel.R:155-455).
# Formals and defaults are the reference's; new optional arguments (seed, device) come last.
dream_parallel <- function(fun, ..., lik = NULL,
                  par.info = list(initial = NULL, min = NULL, max = NULL, mu = NULL, cov = NULL, val_ini = NULL,
                                  bound = NULL, names = NULL, prior = "flat"),
                  nc, t, d,
                  burnin = 0, adapt = 0.1, updateInterval = 10, delta = 3, c_val = 0.1, c_star = 1e-12, nCR = 3,
                  p_g = 0.2, beta0 = 1, thin = 1, outlier_check = TRUE, obs = NULL, abc_rho = NULL, abc_e = NULL,
                  glue_shape = NULL, lik_fun = NULL,
                  past_sample = FALSE, m0 = NULL, archive_update = NULL, psnooker=0, mt = 1,
                  ncores = 1, checkConvergence = FALSE, verbose = TRUE,
                  seed = NULL, device = 0L, series_fx = FALSE) {
  # ncores is accepted and ignored: the population is evaluated by the device plugin, not by forked workers (:176-177)
  .hb_drive(0L, fun, list(...), lik, par.info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g,
            beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun, past_sample, m0, archive_update,
            psnooker, mt, checkConvergence, verbose, seed, device, series_fx)
}

# Thin .Call driver replacing the body of HydroBayes::dream (R/dream.R:126-316): the sequential sweep (chain j+1 sees
# chain j's update, :213-252).  Default prior "uniform" (:127); no archive, snooker or multi-try.  AR and CR come back
# as [out_t, nCR] (:161).
dream <- function(fun, ..., lik = NULL,
                  par.info = list(initial = NULL, min = NULL, max = NULL, mu = NULL, cov = NULL, val_ini = NULL,
                                  bound = NULL, names = NULL, prior = "uniform"),
                  nc, t, d,
                  burnin = 0, adapt = 0.1, updateInterval = 10, delta = 3, c_val = 0.1, c_star = 1e-12, nCR = 3,
                  p_g = 0.2, beta0 = 1, thin = 1, outlier_check = TRUE, obs = NULL,
                  abc_rho = NULL, abc_e = NULL, glue_shape = NULL, lik_fun = NULL,
                  checkConvergence = FALSE, verbose = TRUE,
                  seed = NULL, device = 0L, series_fx = FALSE) {
  if (nc <= delta * 2) stop("Argument 'nc' must be > 'delta' * 2!")      # R/dream.R:136-137
  .hb_drive(1L, fun, list(...), lik, par.info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g,
            beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun, FALSE, NULL, NULL, 0, 1,
            checkConvergence, verbose, seed, device, series_fx)
}
