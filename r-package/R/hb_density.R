# Device log-density for the reference's OTHER consumers of a target function: the single-population samplers
# rwm() / am() / am_sd() / de_mc() (R/rwm.R:42, R/am.R:51, R/am_sd.R:59, R/de_mc.R:34,56: `pdf(x)`) and game()
# (R/game.R:163-178: get(fun) + calc_ll + prior_pdf on the m0 importance samples).  Those functions stay as they are in the
# reference; what changes is where the target is evaluated.  NOT RUN in the build image (no R there); the .Call sequence
# is exercised through tests/mock_r (tests/test_r_glue.py).

# open a log-density session: get(fun) with its `...` (forcing), lik / obs / glue_shape of calc_ll (R/internals.R:3-23),
# for points of dimension d, at most max_n points per evaluation.  Arbitrary R closures stay on the R path.
.hb_density_open <- function(fun, dots, lik, obs, glue_shape, d, max_n) {
  if (!is.character(fun) || !.Call(hbR_target_registered, fun))
    stop("object '", fun, "' of mode 'function' was not found in the device registry (use the reference's R path for closures)")
  .Call(hbR_density_open, fun, as.integer(lik), if (length(dots)) as.numeric(dots[[1]]) else NULL,
        if (is.null(obs)) NULL else as.numeric(obs), glue_shape, as.integer(d), as.numeric(max_n))
}

# points as a double matrix [n, d]: a vector is ONE point (pdf(xp) of the samplers), a matrix has one point per row
.hb_points <- function(x, d) {
  if (!is.matrix(x)) x <- matrix(x, ncol = d)
  storage.mode(x) <- "double"
  if (ncol(x) != d) stop("points have ", ncol(x), " columns, expected ", d)
  x
}

# hb_logdensity(fun, x, ..., lik): calc_ll(get(fun)(x[i, ], ...), lik, obs, NULL, NULL, glue_shape, NULL) for every row of x,
# floored at log(.Machine$double.xmin) as calc_ll does (R/internals.R:19).  One-shot: opens and closes its own session.
hb_logdensity <- function(fun, x, ..., lik, obs = NULL, glue_shape = NULL) {
  if (!is.matrix(x)) x <- matrix(x, nrow = 1)
  s <- .hb_density_open(fun, list(...), lik, obs, glue_shape, ncol(x), nrow(x))
  on.exit(.Call(hbR_density_close, s), add = TRUE)
  .Call(hbR_density_eval, s, .hb_points(x, ncol(x)))
}

# hb_pdf(fun, ..., lik, par.info, d, max_n): a closure to pass as `pdf` to the reference's unchanged rwm() / am() / am_sd() /
# de_mc() / dream_test().  pdf(x) = exp(log-likelihood + log-prior) of the point x (vector of length d) or of every row of
# the matrix x (de_mc's first call, R/de_mc.R:34: max_n >= nc); log = TRUE returns the logarithm (dream_test's pdf,
# R/dream_test.R: `p_prop - p_last`).  The session lives as long as the closure (finalizer on the external pointer).
hb_pdf <- function(fun, ..., lik, par.info = list(prior = "flat"), d, max_n = 1, obs = NULL, glue_shape = NULL, log = FALSE) {
  s <- .hb_density_open(fun, list(...), lik, obs, glue_shape, d, max_n)
  function(x) {
    x <- .hb_points(x, d)
    lpost <- .Call(hbR_density_eval, s, x) + apply(x, 1, prior_pdf, par.info, lik)   # prior_pdf: unchanged R/internals.R:49-69
    if (log) lpost else exp(lpost)
  }
}

# game(): the block R/game.R:163-178 (prior_pdf, get(fun) over the m0 samples, calc_ll, sum) as one device evaluation.
# In the reference's game() replace those lines by
#   lpost_xp <- .hb_game_lpost(theta_m0, par.info, lik, fun, list(...), obs, glue_shape)
# (abc_rho / lik_fun closures keep the R path).  Everything else of game() -- em_pmix, dmixmvn, the estimators -- is host
# work on m0 + m1 numbers and stays.
.hb_game_lpost <- function(theta_m0, par.info, lik, fun, dots, obs, glue_shape) {
  theta_m0 <- .hb_points(theta_m0, ncol(theta_m0))
  s <- .hb_density_open(fun, dots, lik, obs, glue_shape, ncol(theta_m0), nrow(theta_m0))
  on.exit(.Call(hbR_density_close, s), add = TRUE)
  .Call(hbR_density_eval, s, theta_m0) + apply(theta_m0, 1, prior_pdf, par.info, lik)
}
