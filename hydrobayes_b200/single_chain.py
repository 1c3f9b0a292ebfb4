"""rwm() / am() / am_sd() -- the reference's single-chain Metropolis samplers (R/rwm.R:22-56, R/am.R:22-65, R/am_sd.R:22-73)
for REGISTERED device targets.

Each of them is one strictly sequential chain: x_i depends on x_{i-1}, so there is nothing to parallelise inside a chain.
What a GPU can add is replicas: ``n_chains`` independent chains advance in lockstep, the proposals of an iteration are drawn
on the host (``MASS::mvrnorm`` in the reference, numpy here: different generator, same law) and their target values come from
ONE device evaluation per iteration through a log-density session (``hb_density_open`` / ``hb_density_eval``: ``get(fun)`` + ``calc_ll``, floored at
``log(.Machine$double.xmin)`` like every likelihood in the package, R/internals.R:19).  The reference takes R closures
``prior`` / ``pdf``; those stay off the device path (like every R closure): ``fun`` names a device target and ``par_info``
gives the initial sample and the prior, as for ``dream()``.

The acceptance rule is the reference's density ratio ``min(1, p_xp / p_x) > runif(1)`` (R/rwm.R:45-46), evaluated in logs.
Returns ``dict(chain[t, d], density[t])`` for one chain (the reference's structure) and ``chain[t, d, n_chains]``,
``density[t, n_chains]`` for replicas; ``density`` holds densities, not logs, as the reference does.

Not pinned to the reference (the draws differ and there is no R here); the tests check the targets' moments.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi as A
from .game import _prior_pdf


def _open(fun, lik, d, n, target_kw):
    from .sampler import DensitySession
    return DensitySession(fun, int(lik), d, n, **target_kw)      # plugin context and device buffers live for the whole run


def _post(sess, x, lik, par_info):
    ll, _ = sess(x)
    return ll + _prior_pdf(x, par_info, lik)                     # log posterior density up to a constant (pdf() of the reference)


def _initial(par_info, d, n, rng):
    from .sampler import prior_sample
    return np.asarray(prior_sample(par_info, d, n, rng=rng), float).reshape(n, d)   # x[1, 1:d] <- prior(1, d)


def _run(kind, fun, lik, par_info, t, d, n_chains, seed, target_kw):
    if t < 2:
        raise ValueError("t must be at least 2")
    rng = np.random.default_rng(seed)
    n = int(n_chains)
    with _open(fun, lik, d, n, target_kw) as sess:
        return _chains(kind, sess, lik, par_info, t, d, n, rng)


def _chains(kind, sess, lik, par_info, t, d, n, rng):
    s_d = np.full(n, 2.38 ** 2 / d)                              # scaling factor of the covariance matrix
    x = np.full((t, d, n), np.nan); lpx = np.full((t, n), np.nan)
    x[0] = _initial(par_info, d, n, rng).T
    lpx[0] = _post(sess, x[0].T, lik, par_info)
    sig, L = np.sqrt(s_d), None                                  # chol(C) of C = s_d * diag(d) is sig * I: kept as a scale
    accept = np.ones(n)                                          # am_sd: "first sample accepted" (R/am_sd.R:38)
    S1 = x[0].T.copy(); S2 = np.einsum("ni,nj->nij", x[0].T, x[0].T)   # running sums for cov(x[1:(i-1), ])
    for i in range(2, t + 1):                                    # R's i = 2..t
        if kind == "am" and i % 10 == 0:                         # R/am.R:41-46: C <- s_d * (cov(x[1:(i-1), ]) + 1e-4 I)
            m = i - 1
            mean = S1 / m
            cov = (S2 - m * np.einsum("ni,nj->nij", mean, mean)) / (m - 1)
            L = np.linalg.cholesky(s_d[:, None, None] * (cov + 1e-4 * np.eye(d)[None]))
        if kind == "am_sd":
            if i % 25 == 0 and i < t / 2:                        # R/am_sd.R:44-51
                AR = 100.0 * accept / (i - 1)
                s_d = np.where(AR < 20, 0.8 * s_d, s_d)
                s_d = np.where(AR > 30, 1.2 * s_d, s_d)
            sig = np.sqrt(s_d + 1e-4)                            # :53  C <- (s_d + 1e-4) * diag(d)
        z = rng.standard_normal((n, d))
        jump = sig[:, None] * z if L is None else np.einsum("nij,nj->ni", L, z)
        xp = x[i - 2].T + jump                                   # x[i-1, ] + mvrnorm(mu = 0, Sigma = C)
        lpp = _post(sess, xp, lik, par_info)
        acc = np.minimum(0.0, lpp - lpx[i - 2]) > np.log(rng.uniform(size=n))   # min(1, p_xp / p_x) > runif(1)
        x[i - 1] = np.where(acc[None, :], xp.T, x[i - 2])
        lpx[i - 1] = np.where(acc, lpp, lpx[i - 2])
        accept += acc
        if kind == "am":
            xi = x[i - 1].T
            S1 += xi; S2 += np.einsum("ni,nj->nij", xi, xi)
    dens = np.exp(lpx)
    if n == 1:
        return dict(chain=x[:, :, 0], density=dens[:, 0])
    return dict(chain=x, density=dens)


def _kw(forcing, obs, glue_shape):
    return dict(forcing=forcing, obs=obs, glue_shape=float(glue_shape or 0.0))


def rwm(fun: str, *, lik: int, par_info: dict, t: int, d: int, n_chains: int = 1, seed: Optional[int] = None, forcing=None, obs=None,
        glue_shape=None):
    """Random Walk Metropolis (R/rwm.R:22-56): proposal N(x, (2.38^2 / d) I), fixed."""
    return _run("rwm", fun, lik, par_info, t, d, n_chains, seed, _kw(forcing, obs, glue_shape))


def am(fun: str, *, lik: int, par_info: dict, t: int, d: int, n_chains: int = 1, seed: Optional[int] = None, forcing=None, obs=None,
       glue_shape=None):
    """Adaptive Metropolis (R/am.R:22-65): every 10th iteration C <- (2.38^2 / d) (cov(history) + 1e-4 I)."""
    return _run("am", fun, lik, par_info, t, d, n_chains, seed, _kw(forcing, obs, glue_shape))


def am_sd(fun: str, *, lik: int, par_info: dict, t: int, d: int, n_chains: int = 1, seed: Optional[int] = None, forcing=None, obs=None,
          glue_shape=None):
    """Metropolis with an adapted scaling factor (R/am_sd.R:22-73): every 25th iteration of the first half of the run s_d is
    multiplied by 0.8 / 1.2 when the acceptance rate is below 20 % / above 30 %; C = (s_d + 1e-4) I."""
    return _run("am_sd", fun, lik, par_info, t, d, n_chains, seed, _kw(forcing, obs, glue_shape))
