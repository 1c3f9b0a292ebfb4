# Thin .Call driver replacing the body of HydroBayes::dream_parallel (R/dream_parallel.R:155-455).
# Formals and defaults are the reference's; new optional arguments come last.  NOT RUN in the build image (no R).
dream_parallel <- function(fun, ..., lik = NULL,
                  par.info = list(initial = NULL, min = NULL, max = NULL, mu = NULL, cov = NULL, val_ini = NULL,
                                  bound = NULL, names = NULL, prior = "flat"),
                  nc, t, d,
                  burnin = 0, adapt = 0.1, updateInterval = 10, delta = 3, c_val = 0.1, c_star = 1e-12, nCR = 3,
                  p_g = 0.2, beta0 = 1, thin = 1, outlier_check = TRUE, obs = NULL, abc_rho = NULL, abc_e = NULL,
                  glue_shape = NULL, lik_fun = NULL,
                  past_sample = FALSE, m0 = NULL, archive_update = NULL, psnooker=0, mt = 1,
                  ncores = 1, checkConvergence = FALSE, verbose = TRUE,
                  seed = 1312, device = 0L) {
  timea <- Sys.time()
  prior_code <- match(par.info$prior, c("flat", "uniform", "normal")) - 1L
  if (is.na(prior_code)) stop("user prior functions stay off the device path")
  bound_code <- if (is.null(par.info$bound)) 0L else match(par.info$bound, c("bound", "reflect", "fold"))
  if (is.na(bound_code)) stop("Value 'bound' of argument list 'par.info' must be one of {'bound', 'reflect', 'fold'}!")
  cfg <- list(sweep = 0L, nc = nc, t = t, d = d, lik = lik, burnin = burnin, adapt = adapt, updateInterval = updateInterval,
              delta = delta, c_val = c_val, c_star = c_star, nCR = nCR, p_g = p_g, beta0 = beta0, thin = thin,
              outlier_check = outlier_check, prior = prior_code, bound = bound_code, past_sample = past_sample,
              m0 = m0, archive_update = archive_update, psnooker = psnooker, mt = mt, glue_shape = glue_shape,
              seed = seed, checkConvergence = checkConvergence)
  h <- .Call("hbR_create", cfg, as.integer(device))           # argument checks of :166-175 happen inside, same wording
  if (!is.null(par.info$min) && !is.null(par.info$max))
    .Call("hbR_set_bounds", h, as.numeric(par.info$min), as.numeric(par.info$max))
  if (par.info$prior == "normal") .Call("hbR_set_prior_normal", h, as.numeric(par.info$mu), as.matrix(par.info$cov))
  if (!is.null(obs)) .Call("hbR_set_obs", h, as.numeric(obs))
  if (!is.null(abc_rho) || !is.null(lik_fun))                   # names, as in the reference: get(abc_rho), get(lik_fun)
    .Call("hbR_set_lik_args", h, abc_rho, if (is.null(abc_e)) NULL else as.numeric(abc_e), lik_fun)
  dots <- list(...)
  .Call("hbR_set_target", h, fun, if (length(dots)) as.numeric(dots[[1]]) else NULL)   # get(fun)(x, ...) :219
  n0 <- if (past_sample) (if (is.null(m0)) 10 * d else m0) else nc
  z <- prior_sample(par.info, d, n0)                            # unchanged R/internals.R:26-47
  .Call("hbR_set_initial", h, z)
  if (verbose) pb <- txtProgressBar(min = 2, max = t, style = 3)
  done <- 1
  while (done < t) {                                            # the hot loop :257-435 is now hb_run slices
    done <- .Call("hbR_run", h, max(1, t %/% 50))
    if (verbose) setTxtProgressBar(pb, done)
  }
  if (verbose) close(pb)
  out_t <- if (burnin > 0 || thin > 1) (1 - burnin) * t / thin + 1 else t
  chain <- array(NA_real_, dim = c(out_t, d + 3, nc))
  parnames <- if (is.null(par.info$names)) paste0("par", 1:d) else par.info$names
  dimnames(chain) <- list(NULL, c(parnames, "lp", "ll", "lpost"), NULL)
  .Call("hbR_fetch_chain", h, chain)
  res <- .Call("hbR_fetch", h)
  out_AR <- res[[2]]; colnames(out_AR) <- c("n_eval", "accepted")
  fx <- if (fun %in% c("mixture", "t_distribution")) .Call("hbR_fetch_fx", h, array(NA_real_, dim = c(out_t, nc, 1))) else NULL
  op <- .Call("hbR_fetch_outliers", h)                         # outl[[i]] <- check_out$outliers  (:418)
  n_chk <- floor(adapt * t / updateInterval)
  outl <- list(NULL)
  if (outlier_check && !past_sample && n_chk > 0)
    for (i in updateInterval * seq_len(n_chk)) outl[[i]] <- op[[2]][op[[1]] == i]
  output <- list(chain = chain, fx = fx, runtime = Sys.time() - timea, outlier = outl, AR = out_AR, CR = res[[3]])
  if (checkConvergence && out_t > 50)
    output[["R_stat"]] <- .Call("hbR_fetch_rstat", h, matrix(NA_real_, nrow = out_t - 50, ncol = d))
  output
}

# dream(): the same driver with the sequential sweep (R/dream.R:126-316); default prior "uniform" (:127), no archive / MT
dream <- function(fun, ..., lik = NULL,
                  par.info = list(initial = NULL, min = NULL, max = NULL, mu = NULL, cov = NULL, val_ini = NULL,
                                  bound = NULL, names = NULL, prior = "uniform"),
                  nc, t, d, burnin = 0, adapt = 0.1, updateInterval = 10, delta = 3, c_val = 0.1, c_star = 1e-12,
                  nCR = 3, p_g = 0.2, beta0 = 1, thin = 1, outlier_check = TRUE, obs = NULL, abc_rho = NULL, abc_e = NULL,
                  glue_shape = NULL, lik_fun = NULL, checkConvergence = FALSE, verbose = TRUE, seed = 1312, device = 0L) {
  if (nc <= 2 * delta) stop("Argument 'nc' must be > 'delta' * 2!")      # R/dream.R:136-137
  .hb_drive(sweep = 1L, fun = fun, ..., lik = lik, par.info = par.info, nc = nc, t = t, d = d, burnin = burnin,
            adapt = adapt, updateInterval = updateInterval, delta = delta, c_val = c_val, c_star = c_star, nCR = nCR,
            p_g = p_g, beta0 = beta0, thin = thin, outlier_check = outlier_check, obs = obs, abc_rho = abc_rho,
            abc_e = abc_e, glue_shape = glue_shape, lik_fun = lik_fun, checkConvergence = checkConvergence,
            verbose = verbose, seed = seed, device = device)
  # .hb_drive is the body of dream_parallel() above with cfg$sweep taken from its first argument; AR / CR come back
  # as [out_t, nCR] (R/dream.R:161-162) -- hb_out_dims reports the shapes for either sweep
}

R_stat <- function(x) invisible(.Call("hbR_rstat_array", x))   # R/gelman_rubin.R:21
