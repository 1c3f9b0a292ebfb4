"""Shared builders for the parity tests: synthetic inputs of SURVEY 8(d) and oracle/device runners."""
import numpy as np

from hydrobayes_b200 import _abi as A

PHI = 0.6180339887498949


def frac(x):
    return np.modf(x)[0]


def x0_mixture(nc):
    j = np.arange(1, nc + 1)
    return (-20 + 40 * frac(j * PHI))[:, None]


def x0_tdist(nc, d):
    j = np.arange(1, nc + 1)[:, None]
    k = np.arange(1, d + 1)[None, :]
    return -5 + 20 * frac((j * d + k) * PHI)


def x0_hydro(nc):
    j = np.arange(1, nc + 1)
    return (0.01 + 1.99 * frac(j * PHI))[:, None]


def hydro_data(T=3653, catchment=0):
    """forcing P_t = (u_t < 0.4) * Gamma(0.7, scale 8); obs = HydroModel(P, K_c) + N(0, 0.5^2)  (SURVEY 8d)."""
    rng = np.random.default_rng(20240101 + catchment)
    u = rng.uniform(size=T)
    P = (u < 0.4) * rng.gamma(0.7, 8.0, size=T)
    K = 0.1 + 0.8 * frac(catchment * PHI)
    S = 1.0
    Y = np.empty(T)
    for i in range(T):
        S = (P[i] + S) / (1 + K)
        Y[i] = K * S
    obs = Y + rng.normal(0, 0.5, size=T)
    return P, obs


def hydro_data_batch(T, n_catchments, first=0):
    """hydro_data() for catchments first..first+n-1, simulated for all catchments at once: P [n, T], obs [n, T]."""
    P = np.empty((n_catchments, T)); noise = np.empty((n_catchments, T))
    for i in range(n_catchments):
        rng = np.random.default_rng(20240101 + first + i)
        u = rng.uniform(size=T)
        P[i] = (u < 0.4) * rng.gamma(0.7, 8.0, size=T)
        noise[i] = rng.normal(0, 0.5, size=T)
    K = 0.1 + 0.8 * frac((first + np.arange(n_catchments)) * PHI)
    S = np.ones(n_catchments)
    Y = np.empty((n_catchments, T))
    for t in range(T):
        S = (P[:, t] + S) / (1 + K)
        Y[:, t] = K * S
    return P, Y + noise


def run_device(cfg, target, x0, par_min=None, par_max=None, tape=None, forcing=None, obs=None, device=0,
               slice_generations=None, prior_mu=None, prior_cov=None, abc_e=None, abc_rho="abc_distfun", lik_fun="fun_lik"):
    """Drives the C ABI through hydrobayes_b200.Sampler and returns the same dict layout as oracle.run()."""
    from hydrobayes_b200 import Sampler
    with Sampler(cfg, device) as s:
        if par_min is not None:
            s.set_bounds(par_min, par_max)
        if prior_mu is not None:
            s.set_prior_normal(prior_mu, prior_cov)
        if obs is not None:
            s.set_obs(obs)
        if cfg.lik in (21, 22):
            s.set_lik_args(abc_rho=abc_rho, abc_e=abc_e)
        elif cfg.lik == 99:
            s.set_lik_args(lik_fun=lik_fun)
        s.set_target(target, forcing)
        s.set_initial(x0)
        if tape is not None:
            s.set_tape(tape.struct(), keepalive=tape)
        done = 1
        step = slice_generations or cfg.t
        while done < cfg.t:
            done = s.run(step)
        res = {"chain": s.fetch_chain()}
        res["AR"], res["CR"] = s.fetch_ar_cr()
        if cfg.record_accept:
            res["accept"] = s.fetch_accept()
        if cfg.store_fx:
            res["fx"] = s.fetch_fx()[:, :, 0]
        res["outlier"] = s.fetch_outliers()
        xt, lp, ll, lpost = s.fetch_state()
        res.update(xt=xt, lp=lp, ll=ll, lpost=lpost)
        if cfg.check_convergence:
            res["R_stat"] = s.fetch_rstat()
        res["timing"] = s.timing()
    return res


def assert_run_parity(dev, ora, rtol=1e-12, atol=0.0, check_fx=True):
    """accept/reject sequences identical; chain, AR, CR, final state within rtol; outlier lists identical."""
    assert ora["rc"] == 0, ora["err"]
    if "accept" in dev:
        assert np.array_equal(dev["accept"], ora["accept"]), "accept/reject sequences differ"
    assert dev["outlier"] == ora["outlier"], "outlier sets differ"
    np.testing.assert_allclose(dev["chain"], ora["chain"], rtol=rtol, atol=atol, equal_nan=True)
    np.testing.assert_allclose(dev["AR"], ora["AR"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(dev["CR"], ora["CR"], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(dev["xt"], ora["xt"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(dev["lpost"], ora["lpost"], rtol=max(rtol, 1e-12))
    if check_fx and "fx" in dev:
        np.testing.assert_allclose(dev["fx"], ora["fx"], rtol=max(rtol, 1e-12), equal_nan=True)


def copy_cfg(cfg, **kw):
    c = A.HbConfig()
    import ctypes
    ctypes.memmove(ctypes.byref(c), ctypes.byref(cfg), ctypes.sizeof(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def random_case(case):
    """Seeded sweep over the option space: both sweeps, every bound / prior / likelihood combination of the device path.
    Returns (cfg, target, x0, extra keyword arguments of oracle.run / run_device)."""
    rng = np.random.default_rng(1000 + case)
    sweep = int(rng.integers(0, 2))
    target = ["mixture", "t_distribution", "HydroModel"][int(rng.integers(0, 3))]
    delta = int(rng.integers(1, 5)); nCR = int(rng.integers(1, 7))
    nc = 2 * delta + 1 + int(rng.integers(0, 6))
    t = int(rng.integers(20, 70))
    kw = dict(nc=nc, t=t, seed=int(rng.integers(1, 10_000)), sweep=sweep, delta=delta, nCR=nCR,
              adapt=float(rng.choice([0.0, 0.3, 0.6, 1.0])), update_interval=int(rng.choice([1, 3, 5, 10])),
              p_g=float(rng.choice([0.0, 0.2, 0.7])), beta0=float(rng.choice([1.0, 0.5])), c_val=float(rng.choice([0.1, 0.4])),
              c_star=float(rng.choice([1e-12, 1e-3])), outlier_check=int(rng.integers(0, 2)))
    thin = int(rng.choice([1, 1, 2, 5]))
    if t % thin == 0:
        kw["thin"] = thin
    extra = {}
    if target == "mixture":
        kw.update(d=1, lik=1, store_fx=1)
        x0 = x0_mixture(nc)
        if rng.integers(0, 2):
            kw.update(prior=A.HB_PRIOR_UNIFORM, bound=int(rng.integers(0, 4)))
            extra.update(par_min=[-25.0], par_max=[22.0])
    elif target == "t_distribution":
        d = int(rng.integers(1, 6))
        kw.update(d=d, lik=2, store_fx=1)
        x0 = x0_tdist(nc, d)
        if rng.integers(0, 2):
            kw.update(prior=int(rng.choice([A.HB_PRIOR_FLAT, A.HB_PRIOR_UNIFORM])), bound=int(rng.integers(0, 4)))
            extra.update(par_min=list(-6.0 - np.arange(d)), par_max=list(16.0 + np.arange(d)))
    else:
        T = int(rng.integers(5, 40))
        P, obs = hydro_data(T=T, catchment=case)
        lik = int(rng.choice([31, 21, 22, 99]))
        kw.update(d=1, lik=lik, glue_shape=float(rng.choice([5.0, 50.0])), prior=A.HB_PRIOR_UNIFORM, bound=int(rng.integers(1, 4)))
        x0 = x0_hydro(nc)
        extra.update(par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
        if lik in (21, 22):
            extra["abc_e"] = float(rng.uniform(0.5, 3.0)) if rng.integers(0, 2) else rng.uniform(0.5, 3.0, size=T)
    cfg = A.default_config(**kw)
    return cfg, target, x0, extra
