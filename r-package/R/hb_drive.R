# Shared body of the thin .Call drivers dream() (R/dream.R:126-316) and dream_parallel() (R/dream_parallel.R:155-455).
# Everything the reference does inside `for(i in 2:t)` now happens on the device behind hbR_run; what stays in R is what
# the reference does once per call: argument fix-ups, prior_sample() (unchanged, R/internals.R:26-47), allocation of the
# result arrays (R owns them) and the assembly of the returned list.  NOT RUN in the build image (no R there); the same
# call sequence is exercised through tests/mock_r (tests/test_r_glue.py).
.hb_drive <- function(sweep, fun, dots, lik, par.info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR,
                      p_g, beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun, past_sample, m0,
                      archive_update, psnooker, mt, checkConvergence, verbose, seed, device, series_fx = FALSE) {
  timea <- Sys.time()
  if (is.null(lik)) stop("Argument 'lik' has to be one of {1,2,21,22}!")
  if (!is.character(fun) || !.Call(hbR_target_registered, fun))     # get(fun) of an arbitrary closure stays on the R path
    stop("object '", fun, "' of mode 'function' was not found in the device registry (use the reference sampler for R closures)")
  prior_code <- match(par.info$prior, c("flat", "uniform", "normal")) - 1L
  if (is.na(prior_code)) stop("user prior functions stay off the device path")
  bound_code <- if (is.null(par.info$bound)) 0L else match(par.info$bound, c("bound", "reflect", "fold"))
  if (is.na(bound_code)) stop("Value 'bound' of argument list 'par.info' must be one of {'bound', 'reflect', 'fold'}!")
  # the reference draws from R's RNG stream; the device draws from Philox keyed by one integer taken from that stream,
  # so set.seed() before the call still makes the run reproducible
  if (is.null(seed)) seed <- sample.int(.Machine$integer.max, 1)
  store_fx <- fun %in% c("mixture", "t_distribution")               # targets with scalar output (m == 1)
  cfg <- list(sweep = sweep, nc = nc, t = t, d = d, lik = lik, burnin = burnin, adapt = adapt, updateInterval = updateInterval,
              delta = delta, c_val = c_val, c_star = c_star, nCR = nCR, p_g = p_g, beta0 = beta0, thin = thin,
              outlier_check = outlier_check, prior = prior_code, bound = bound_code, past_sample = past_sample,
              m0 = m0, archive_update = archive_update, psnooker = psnooker, mt = mt, glue_shape = glue_shape,
              seed = seed, checkConvergence = checkConvergence, store_fx = store_fx)
  h <- .Call(hbR_create, cfg, as.integer(device))                   # argument checks of :166-175 happen inside, same wording
  on.exit(.Call(hbR_destroy, h), add = TRUE)
  if (!is.null(par.info$min) && !is.null(par.info$max))
    .Call(hbR_set_bounds, h, as.numeric(par.info$min), as.numeric(par.info$max))
  if (par.info$prior == "normal") .Call(hbR_set_prior_normal, h, as.numeric(par.info$mu), as.matrix(par.info$cov))
  if (!is.null(obs)) .Call(hbR_set_obs, h, as.numeric(obs))
  if (!is.null(abc_rho) || !is.null(lik_fun))                        # names, as in the reference: get(abc_rho), get(lik_fun)
    .Call(hbR_set_lik_args, h, abc_rho, if (is.null(abc_e)) NULL else as.numeric(abc_e), lik_fun)
  .Call(hbR_set_target, h, fun, if (length(dots)) as.numeric(dots[[1]]) else NULL)   # get(fun)(x, ...) :219
  n0 <- if (past_sample) (if (is.null(m0)) 10 * d else m0) else nc   # :168-171
  z <- prior_sample(par.info, d, n0)                                 # unchanged R/internals.R:26-47
  z <- as.matrix(z); storage.mode(z) <- "double"                     # the glue reads n0 * d doubles: check before handing over
  if (nrow(z) != n0 || ncol(z) != d)
    stop("initial sample has dimension ", nrow(z), " x ", ncol(z), ", expected ", n0, " x ", d, " (par.info$val_ini)")
  .Call(hbR_set_initial, h, z)

  dims <- .Call(hbR_out_dims, h)                                     # out_t, AR rows/cols, CR rows/cols, fx_m, R_stat rows
  out_t <- dims[1]
  chain <- array(NA_real_, dim = c(out_t, d + 3, nc))                # :201-203, allocated before the loop as in the reference
  parnames <- if (is.null(par.info$names)) paste0("par", 1:d) else par.info$names
  dimnames(chain) <- list(NULL, c(parnames, "lp", "ll", "lpost"), NULL)

  if (verbose) pb <- txtProgressBar(min = 2, max = t, style = 3)
  done <- 1
  while (done < t) {                                                 # the hot loop :257-435 is now hb_run slices
    done <- .Call(hbR_run, h, max(1, t %/% 50))                      # R_CheckUserInterrupt() between slices
    if (verbose) setTxtProgressBar(pb, done)
  }
  if (verbose) close(pb)

  chain <- .Call(hbR_fetch_chain, h, chain)                          # the glue fills and returns the array (no reliance on in-place mutation)
  out_AR <- .Call(hbR_fetch_ar, h, matrix(NA_real_, dims[2], dims[3]))
  out_CR <- .Call(hbR_fetch_cr, h, matrix(NA_real_, dims[4], dims[5]))
  if (sweep == 0L) colnames(out_AR) <- c("n_eval", "accepted")       # R/dream_parallel.R:205
  # fx: the reference keeps the raw target output of every stored state, fx[out_t, nc, m] (R/dream_parallel.R:230, 394).
  # Targets with scalar output (m == 1) return it as in the reference.  For series targets (HydroModel: m == T values per
  # chain and stored generation) the device keeps no series; fx is NULL unless series_fx = TRUE, which rebuilds
  # fx[it, c, ] = HydroModel(forcing, chain[it, 1:d, c]) on the device from the stored chain -- the reference's values,
  # except for the rows of a chain between an outlier reset and its next accepted proposal, where the reference's fx is
  # the stale series of the replaced state (SURVEY F9).
  fx <- if (store_fx) .Call(hbR_fetch_fx, h, array(NA_real_, dim = c(out_t, nc, 1))) else NULL
  if (is.null(fx) && series_fx && fun %in% c("HydroModel", "fun_wrap_glue") && d == 1 && length(dots)) {
    ser <- .Call(hbR_hydromodel_series, as.numeric(dots[[1]]), as.numeric(chain[, 1, ]))   # [T, out_t * nc]
    fx <- aperm(array(ser, dim = c(nrow(ser), out_t, nc)), c(2, 3, 1))
  }
  op <- .Call(hbR_fetch_outliers, h)                                 # outl[[i]] <- check_out$outliers  (:418)
  outl <- list(NULL)
  n_chk <- floor(adapt * t / updateInterval)
  if (outlier_check && !past_sample && n_chk > 0)
    for (i in updateInterval * seq_len(n_chk)) outl[[i]] <- op[[2]][op[[1]] == i]
  output <- list(chain = chain, fx = fx, runtime = Sys.time() - timea, outlier = outl, AR = out_AR, CR = out_CR)
  if (checkConvergence && dims[7] > 0)                               # SURVEY F7: R_stat(chain[1:ind_out, 1:d, ]) as dream_test does
    output[["R_stat"]] <- .Call(hbR_fetch_rstat, h, matrix(NA_real_, nrow = dims[7], ncol = d))
  output
}
