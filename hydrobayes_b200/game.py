"""GAME sampling of the marginal likelihood -- host mirror of HydroBayes::game (R/game.R:79-209, Volpi et al. 2017).

The reference fits a Gaussian mixture to the posterior sample (``em_pmix``, R/internal_game.R:2-28, through the third-party
EMCluster package), draws importance samples from it and evaluates the target density on them with ``get(fun)`` +
``calc_ll`` + ``prior_pdf``.  Here the only heavy part -- the ``m0`` target evaluations of the importance / bridge
estimators (R/game.R:163-178) -- runs on the device through the log-density plugin (``hb_logdensity``); the mixture fit and
the estimators are a few lines of host code, as in R.

Not pinned to the reference: EMCluster's ``init.EM`` starts from random initialisations and is absent here, so the mixture
is fitted with scikit-learn's EM (same model: full-covariance Gaussian mixture, same selection rule: smallest BIC or
smallest variance of the density ratio over J = 1..Jmax).  What is checked (tests/test_game.py) is the estimator itself:
for a normalised target the evidence is 1.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi as A

_XMIN, _XMAX = np.finfo(float).tiny, np.finfo(float).max
_FLOOR = np.log(_XMIN)                                          # log(.Machine$double.xmin), R/internals.R:19,65


def _dmixmvn(x, gm):
    """density of the fitted mixture at the rows of x (EMCluster::dmixmvn, R/game.R:138,161)"""
    return np.exp(gm.score_samples(np.atleast_2d(x)))


def _em_pmix(theta, theta_post, J, ic, rng):
    """R/internal_game.R:2-28 with scikit-learn's EM in place of EMCluster::init.EM"""
    from sklearn.mixture import GaussianMixture
    try:
        gm = GaussianMixture(n_components=J, covariance_type="full", random_state=int(rng.integers(1 << 31)), n_init=2).fit(theta)
    except Exception:
        return dict(pmix=None, ic_val=np.nan, J=J)
    if not gm.converged_:                                        # em_res$flag != 0 -> NaN, :7-10
        return dict(pmix=gm, ic_val=np.nan, J=J)
    if "bic" in ic.lower():
        ic_val = gm.bic(theta)                                   # em.bic, :14-15
    elif "var" in ic.lower():
        ic_val = np.var(theta_post / _dmixmvn(theta, gm), ddof=1)   # :17-21
    else:
        raise ValueError("Wrong choice of Information Criterium to evaluate the EM calibration of the Gaussian mixture "
                         "distribution. Must be one oc {'BIC' or 'var'}.")
    return dict(pmix=gm, ic_val=ic_val, J=J)


def _prior_pdf(x, par_info, lik):
    """prior_pdf for the rows of x (R/internals.R:49-69): flat / uniform / normal; lik == 22 -> 0"""
    n, d = x.shape
    prior = par_info.get("prior", "flat")
    if lik == 22 or prior == "flat":
        lp = np.zeros(n)
    elif prior == "uniform":
        mn = np.broadcast_to(np.asarray(par_info["min"], float), (d,)); mx = np.broadcast_to(np.asarray(par_info["max"], float), (d,))
        inside = (x >= mn) & (x <= mx)
        with np.errstate(divide="ignore"):
            lp = np.where(inside, -np.log(mx - mn), -np.inf).sum(axis=1)
    elif prior == "normal":
        from scipy.stats import multivariate_normal
        lp = multivariate_normal(np.asarray(par_info["mu"], float), np.asarray(par_info["cov"], float)).logpdf(x)
    else:
        raise A.HbError(A.HB_ERR_UNSUPPORTED, f"prior = {prior!r}: user prior functions are R closures and stay off the device path")
    return np.maximum(lp, _FLOOR)                                # :65


def _bound(x, par_info):
    """bound_par on the importance samples (R/game.R:147-156, R/internals.R:89-113)"""
    if par_info.get("min") is None or par_info.get("max") is None or par_info.get("bound") is None:
        return x
    d = x.shape[1]
    mn = np.broadcast_to(np.asarray(par_info["min"], float), (d,)); mx = np.broadcast_to(np.asarray(par_info["max"], float), (d,))
    h = par_info["bound"]
    if h == "bound":
        return np.clip(x, mn, mx)
    out = x.copy()
    for _ in range(1024):
        lo = out < mn; hi = out > mx
        if not (lo.any() or hi.any()):
            break
        if h == "reflect":
            out = np.where(lo, mn + np.abs(mn - out), out); hi = out > mx
            out = np.where(hi, mx - np.abs(mx - out), out)
        elif h == "fold":
            out = np.where(lo, mx - np.abs(mn - out), out); hi = out > mx
            out = np.where(hi, mn + np.abs(mx - out), out)
        else:
            raise ValueError("Value 'bound' of argument list 'par.info' must be one of {'bound', 'reflect', 'fold'}!")
    return out


def game(theta, theta_post, m_sub: Optional[int] = None, m0: Optional[int] = None, *, m1: int, Jmax: int = 5, ic: str, omega: float,
         ob: bool = False, R: Optional[int] = None, par_info: Optional[dict] = None, lik: Optional[int] = None, fun: Optional[str] = None,
         forcing=None, obs=None, abc_rho=None, abc_e=None, glue_shape=None, lik_fun=None, ncores: int = 1, seed: Optional[int] = None):
    """Gaussian Mixture Importance sampling of the model evidence -- drop-in for HydroBayes::game (R/game.R:79-209).

    ``theta`` [m, d] posterior samples and ``theta_post`` [m] their posterior densities (NOT logs, as in the reference), e.g.
    ``chain[burn:, :d, :]`` and ``exp(chain[burn:, d + 2, :])`` of ``dream_parallel``.  ``fun`` is the NAME of a registered
    device target; it is evaluated on the ``m0`` importance samples on the device (``omega > 0``).  Returns
    ``dict(Z, J, ic_val)``."""
    theta = np.atleast_2d(np.asarray(theta, float)); theta_post = np.asarray(theta_post, float).ravel()
    if theta.shape[0] == 1 and theta_post.size > 1:
        theta = theta.T
    m, d = theta.shape
    # argument checks, R/game.R:83-99
    if m != theta_post.size:
        raise ValueError("Number of rows of 'theta' and length of 'theta_post' must be the same!")
    if omega < 0 or omega > 1:
        raise ValueError("Argument 'omega' must in [0,1]!")
    if "bic" not in ic.lower() and "var" not in ic.lower():
        raise ValueError("Argument 'ic' must be one of {'bic', 'var'}!")
    if m1 > m:
        raise ValueError("Argument 'm1' must not be larger than the number of rows of 'theta'!")
    if m_sub is not None and m_sub > m - m1:
        raise ValueError("Argument 'm_sub' must be <= nrow(theta)-m1!")
    if omega > 0:
        if m0 is None:
            raise ValueError("Argument 'omega' is > 0 and thus 'm0' needs to be given!")
        if par_info is None:
            raise ValueError("Argument 'omega' is > 0 and thus 'par.info' needs to be given!")
        if lik is None:
            raise ValueError("Argument 'omega' is > 0 and thus 'lik' needs to be given!")
        if fun is None:
            raise ValueError("Argument 'omega' is > 0 and thus 'fun' needs to be given!")
        if ob and R is None:
            R = 10
    rng = np.random.default_rng(seed)

    # calibrate the mixture model for different J, select the best (:104-123)
    sub = np.arange(m) if m_sub is None else rng.choice(m, m_sub, replace=False)
    fits = [_em_pmix(theta[sub], theta_post[sub], j, ic, rng) for j in range(1, Jmax + 1)]
    fits = [f for f in fits if np.isfinite(f["ic_val"])]
    if not fits:
        raise ValueError("Could not identify a suitable Gaussian mixture distribution. Change parameters 'Jmax' and/or 'ic' and try "
                         "again. However, the distribution given within 'theta' might be not describable by a Gaussian mixture "
                         "distribution!")
    best = min(fits, key=lambda f: f["ic_val"])
    gm = best["pmix"]

    # sub-sample of the posterior realisations (:126-140)
    pool = np.arange(m) if m_sub is None else np.setdiff1d(np.arange(m), sub)
    r_samp = rng.choice(pool, m1, replace=False)
    th1 = theta[r_samp]; post1 = theta_post[r_samp]
    pmix1 = np.clip(_dmixmvn(th1, gm), _XMIN, _XMAX)

    if omega == 0:
        Z = 1.0 / np.mean(pmix1 / post1)                         # reciprocal importance sampling, eq. 7 (:143-145)
    else:
        th0, _ = gm.sample(int(m0))                              # samples from q0 = the mixture (:150-153)
        th0 = _bound(np.atleast_2d(th0), par_info)               # :155-163
        pmix0 = np.clip(_dmixmvn(th0, gm), _XMIN, _XMAX)
        lp = _prior_pdf(th0, par_info, lik)                      # :170
        # get(fun) + calc_ll on the m0 samples: the device plugin (floored at log(xmin) by the library, R/internals.R:19)
        from .sampler import log_density
        if not isinstance(fun, str):
            raise A.HbError(A.HB_ERR_UNKNOWN_TARGET, "fun must be the NAME of a registered device target")
        ll, _ = log_density(fun, th0, int(lik), forcing=forcing, obs=obs, glue_shape=float(glue_shape or 0.0))
        post0 = np.exp(ll + lp)                                  # :181-183
        if omega == 1:
            Z = np.mean(post0 / pmix0)                           # importance sampling, eq. 10
        else:                                                    # bridge sampling with geometric bridge, eq. 6
            q0 = pmix0 ** (1 - omega) * post0 ** omega
            q1 = pmix1 ** (1 - omega) * post1 ** omega
            Z = np.mean(q0 / pmix0) / np.mean(q1 / post1)
        if ob:                                                   # optimal bridge, eq. 11 (:195-203)
            s0 = m0 / (m0 + m1); s1 = m1 / (m0 + m1)
            for _ in range(int(R)):
                Z = np.mean((post0 / pmix0) / (s0 * Z + s1 * post0 / pmix0)) / np.mean(1.0 / (s0 * Z + s1 * post1 / pmix1))
    return dict(Z=float(Z), J=int(best["J"]), ic_val=float(best["ic_val"]))


def de_mc(fun: str, *, lik: int, par_info: dict, nc: int, t: int, d: int, seed: Optional[int] = None, device: int = 0, **kw):
    """Differential Evolution Markov Chain (ter Braak 2006) -- device path of HydroBayes::de_mc (R/de_mc.R:23-70) for REGISTERED
    targets.  The reference's loop is the sequential DREAM sweep with one pair (delta = 1), no crossover (nCR = 1: every
    coordinate moves), gamma = 2.38/sqrt(2 d) or 1 with probability 0.1 (:26, :50), lambda = 0, zeta ~ N(0, 1e-6^2) (:53), no
    adaptation and no outlier handling -- i.e. ``dream()`` with that configuration.  The reference takes R closures ``prior`` /
    ``pdf``; those stay off the device path: here ``fun`` names a device target and ``par_info`` gives the initial sample and
    the prior, as for ``dream()``.  Returns ``dict(chain[t, d, nc], density[t, nc])`` (densities, not logs, as the reference)."""
    from .sampler import dream
    res = dream(fun, lik=lik, par_info=par_info, nc=nc, t=t, d=d, delta=1, nCR=1, c_val=0.0, c_star=1e-6, p_g=0.1, adapt=0.0,
                outlier_check=False, verbose=False, seed=seed, device=device, **kw)
    return dict(chain=np.ascontiguousarray(res["chain"][:, :d, :]), density=np.exp(res["chain"][:, d + 2, :]))
