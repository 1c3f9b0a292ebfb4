# The reference's single-population samplers over the device path, for REGISTERED targets (SURVEY 8f rank 4).  The
# reference's own rwm() / am() / am_sd() / de_mc() (closure arguments `prior`, `pdf`) stay in the package unchanged and can be
# given a device-backed `pdf` through hb_pdf() (hb_density.R).  The functions here are the batched forms a device pays for:
#   metropolis_replicas(): n_chains independent copies of rwm() / am() / am_sd() advancing in lockstep, ONE device
#                          evaluation per iteration for all replicas (one chain is strictly sequential: R/rwm.R:38-53);
#   de_mc_device():        de_mc() as a configuration of the dream() device path.
# NOT RUN in the build image (no R there).  The Python mirror of the same logic (hydrobayes_b200/single_chain.py, game.py)
# is what the tests run.

# kind = "rwm":   proposal N(x, (2.38^2 / d) I), fixed                                             (R/rwm.R:25-28, 40)
# kind = "am":    every 10th iteration C <- s_d * (cov(x[1:(i-1), ]) + 1e-4 I)                      (R/am.R:41-46)
# kind = "am_sd": every 25th iteration of the first half s_d <- 0.8 / 1.2 * s_d when the acceptance rate is below 20 % /
#                 above 30 %; C <- (s_d + 1e-4) I                                                   (R/am_sd.R:44-53)
# Acceptance: min(1, p_xp / p_x) > runif(1) (R/rwm.R:45-46), evaluated in logs.  The initial states and the prior come
# from par.info as for dream() (prior_sample / prior_pdf, unchanged R/internals.R:26-69).
# Value: list(chain[t, d, n_chains], density[t, n_chains]); for n_chains = 1 chain[t, d], density[t] as the reference.
metropolis_replicas <- function(kind = c("rwm", "am", "am_sd"), fun, ..., lik, par.info, t, d, n_chains = 1,
                                obs = NULL, glue_shape = NULL) {
  kind <- match.arg(kind)
  if (t < 2) stop("Argument 't' must be at least 2!")
  n <- n_chains
  s <- .hb_density_open(fun, list(...), lik, obs, glue_shape, d, n)
  on.exit(.Call(hbR_density_close, s), add = TRUE)
  lpost <- function(x) .Call(hbR_density_eval, s, x) + apply(x, 1, prior_pdf, par.info, lik)

  s_d <- rep(2.38^2 / d, n)                                          # scaling factor of the covariance matrix
  x <- array(NaN, dim = c(t, d, n))
  lp_x <- matrix(NaN, nrow = t, ncol = n)
  xi <- .hb_points(prior_sample(par.info, d, n), d)                  # x[1, 1:d] <- prior(1, d), one row per replica
  if (nrow(xi) != n) stop("initial sample has ", nrow(xi), " rows, expected ", n, " (par.info$val_ini)")
  x[1, , ] <- t(xi)
  lp_x[1, ] <- lpost(xi)
  Lc <- lapply(1:n, function(k) sqrt(s_d[k]) * diag(d))              # lower Cholesky factor of C = s_d * diag(d)
  accept <- rep(1, n)                                                # am_sd: "first sample accepted" (R/am_sd.R:38)

  for (i in 2:t) {
    if (kind == "am" && i %% 10 == 0)
      Lc <- lapply(1:n, function(k) t(chol(s_d[k] * (cov(matrix(x[1:(i - 1), , k], ncol = d)) + 1e-4 * diag(d)))))
    if (kind == "am_sd") {
      if (i %% 25 == 0 && i < t / 2) {
        AR <- 100 * (accept / (i - 1))
        s_d <- ifelse(AR < 20, 0.8 * s_d, s_d)
        s_d <- ifelse(AR > 30, 1.2 * s_d, s_d)
      }
      Lc <- lapply(1:n, function(k) sqrt(s_d[k] + 1e-4) * diag(d))
    }
    z <- matrix(rnorm(n * d), nrow = n, ncol = d)
    jump <- vapply(1:n, function(k) as.numeric(Lc[[k]] %*% z[k, ]), numeric(d))   # [d, n] (a vector when d == 1)
    xp <- xi + matrix(jump, nrow = n, ncol = d, byrow = TRUE)        # x[i-1, ] + mvrnorm(mu = 0, Sigma = C)
    lp_xp <- lpost(xp)
    acc <- pmin(0, lp_xp - lp_x[i - 1, ]) > log(runif(n))
    xi[acc, ] <- xp[acc, ]
    x[i, , ] <- t(xi)
    lp_x[i, ] <- ifelse(acc, lp_xp, lp_x[i - 1, ])
    accept <- accept + acc
  }
  if (n == 1) return(list(chain = matrix(x[, , 1], nrow = t, ncol = d), density = exp(lp_x[, 1])))
  list(chain = x, density = exp(lp_x))
}

# de_mc() (R/de_mc.R:23-70) is the sequential DREAM sweep with one pair (delta = 1), no crossover (nCR = 1: every
# coordinate moves), gamma = 2.38 / sqrt(2 d) or 1 with probability 0.1 (:26, :50), lambda = 0, zeta ~ N(0, 1e-6^2) (:53), no
# adaptation and no outlier handling: dream() with that configuration.  Value as the reference: chain[t, d, nc],
# density[t, nc] (densities, not logs).
de_mc_device <- function(fun, ..., lik, par.info, nc, t, d, obs = NULL, glue_shape = NULL, seed = NULL, device = 0L) {
  res <- dream(fun, ..., lik = lik, par.info = par.info, nc = nc, t = t, d = d, delta = 1, nCR = 1, c_val = 0, c_star = 1e-6,
               p_g = 0.1, adapt = 0, outlier_check = FALSE, obs = obs, glue_shape = glue_shape, verbose = FALSE, seed = seed,
               device = device)
  list(chain = res$chain[, 1:d, , drop = FALSE], density = exp(res$chain[, d + 3, ]))
}
