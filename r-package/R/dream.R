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
                  seed = NULL, device = 0L) {
  if (nc <= delta * 2) stop("Argument 'nc' must be > 'delta' * 2!")      # R/dream.R:136-137
  .hb_drive(1L, fun, list(...), lik, par.info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g,
            beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun, FALSE, NULL, NULL, 0, 1,
            checkConvergence, verbose, seed, device)
}
