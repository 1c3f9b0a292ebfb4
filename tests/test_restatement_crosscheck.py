"""Double-entry check of the CPU oracle (CPU only, no GPU): the C oracle (oracle/hb_oracle.c) and an independent
line-by-line Python transliteration of the same R code (tests/r_restatement.py) replay ONE decision tape and must agree:
identical accept/reject sequences and outlier sets, chains within 1e-10 relative.

This does not pin the oracle to a real R run (nothing can, here: no R, no reference tests -- DESIGN.md section 1); it
removes "a slip in one restatement" as a failure mode, because the two were written separately from the R sources
(R/internals.R, R/dream_parallel.R:155-455, R/dream.R:126-316, R/gelman_rubin.R:21-56).
"""
import numpy as np
import pytest

import helpers as H
import r_restatement as R
from hydrobayes_b200 import _abi as A

BOUND = {A.HB_BOUND_NONE: None, A.HB_BOUND_BOUND: "bound", A.HB_BOUND_REFLECT: "reflect", A.HB_BOUND_FOLD: "fold"}
PRIOR = {A.HB_PRIOR_FLAT: "flat", A.HB_PRIOR_UNIFORM: "uniform", A.HB_PRIOR_NORMAL: "normal"}


@pytest.fixture(scope="module")
def oracle():
    from oracle import hb_oracle_py
    hb_oracle_py.lib()
    return hb_oracle_py


def _both(oracle, cfg, target, x0, par_min=None, par_max=None, forcing=None, obs=None, abc_e=None, prior_mu=None,
          prior_cov=None):
    kw = {k: v for k, v in dict(par_min=par_min, par_max=par_max, forcing=forcing, obs=obs, abc_e=abc_e, prior_mu=prior_mu,
                                prior_cov=prior_cov).items() if v is not None}
    rec = oracle.Tape.for_config(cfg)
    r0 = oracle.run(cfg, target, x0, record=rec, **kw)                 # Philox draws, recorded as decisions
    assert r0["rc"] == 0, r0["err"]
    cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_TAPE, record_accept=1)
    ora = oracle.run(cfg_t, target, x0, tape=rec, **kw)
    assert ora["rc"] == 0, ora["err"]

    par_info = dict(prior=PRIOR[cfg.prior], bound=BOUND[cfg.bound])
    if par_min is not None:
        par_info["min"] = np.asarray(par_min, float) if np.size(par_min) > 1 else float(np.ravel(par_min)[0])
        par_info["max"] = np.asarray(par_max, float) if np.size(par_max) > 1 else float(np.ravel(par_max)[0])
    if prior_mu is not None:
        par_info["mu"], par_info["cov"] = prior_mu, prior_cov
    if target == "mixture":
        fun, dots = R.mixture, ()
    elif target == "t_distribution":
        fun, dots = R.t_distribution, ()
    else:                                                             # fun_wrap_glue(x, forc), examples/script_examples.R:495-501
        fun, dots = (lambda x, forc: R.HydroModel(forc, float(x[0]))), (np.asarray(forcing, float),)
    common = dict(burnin=cfg.burnin, adapt=cfg.adapt, updateInterval=cfg.update_interval, delta=cfg.delta, c_val=cfg.c_val,
                  c_star=cfg.c_star, nCR=cfg.nCR, p_g=cfg.p_g, beta0=cfg.beta0, thin=cfg.thin,
                  outlier_check=bool(cfg.outlier_check), obs=None if obs is None else np.asarray(obs, float),
                  abc_rho=R.abc_distfun, abc_e=abc_e, glue_shape=cfg.glue_shape, lik_fun=R.fun_lik,
                  checkConvergence=bool(cfg.check_convergence))
    if cfg.sweep == A.HB_SWEEP_SEQUENTIAL:
        py = R.dream(fun, dots, cfg.lik, par_info, cfg.nc, cfg.t, cfg.d, rec, x0, **common)
    else:
        py = R.dream_parallel(fun, dots, cfg.lik, par_info, cfg.nc, cfg.t, cfg.d, rec, x0, past_sample=bool(cfg.past_sample),
                              m0=cfg.m0 or None, archive_update=cfg.archive_update or None, psnooker=cfg.psnooker,
                              mt=cfg.mt, **common)
    return py, ora


def _agree(py, ora, rtol=1e-10, fx=True):
    assert np.array_equal(py["accept"], ora["accept"]), "accept/reject sequences differ between the two restatements"
    assert py["outlier"] == ora["outlier"], (py["outlier"], ora["outlier"])
    np.testing.assert_allclose(ora["chain"], py["chain"], rtol=rtol, equal_nan=True)
    np.testing.assert_allclose(ora["AR"], py["AR"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(ora["CR"], py["CR"], rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(ora["xt"], py["xt"], rtol=rtol)
    np.testing.assert_allclose(ora["lpost"], py["lpost"], rtol=rtol)
    np.testing.assert_allclose(ora["lp"], py["lp"], rtol=rtol)
    np.testing.assert_allclose(ora["ll"], py["ll"], rtol=rtol)
    np.testing.assert_allclose(ora["pCR"], py["pCR"], rtol=1e-9)
    np.testing.assert_allclose(ora["n_id"], py["n_id"])
    # the oracle stops accumulating J when the adaptation period ends (pCR is frozen from then on, R/dream_parallel.R:403)
    np.testing.assert_allclose(ora["J"], py["J_adapt"], rtol=1e-8)
    if fx and py["fx"].shape[2] == 1:
        np.testing.assert_allclose(ora["fx"], py["fx"][:, :, 0], rtol=rtol, equal_nan=True)


@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
def test_mixture_c1_shape(oracle, sweep):
    # BASELINE C1 (mixture, nc = 10, d = 1, lik = 1), shortened; adaptation + outlier checks inside the run
    cfg = A.default_config(nc=10, t=300, d=1, lik=1, seed=1312, sweep=sweep, adapt=0.4, store_fx=1)
    py, ora = _both(oracle, cfg, "mixture", H.x0_mixture(10))
    _agree(py, ora)
    assert 0.05 < ora["accept"].mean() < 0.95


@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
@pytest.mark.parametrize("kw", [dict(), dict(thin=5), dict(burnin=0.2, thin=4), dict(adapt=0.5, update_interval=7),
                                dict(delta=1, nCR=1), dict(delta=4, nCR=6, p_g=0.5, beta0=0.7, c_val=0.3, c_star=1e-3)])
def test_tdist_options(oracle, sweep, kw):
    nc, d = 12, 6
    cfg = A.default_config(nc=nc, t=120, d=d, lik=2, seed=11, sweep=sweep, store_fx=1, **kw)
    py, ora = _both(oracle, cfg, "t_distribution", H.x0_tdist(nc, d))
    _agree(py, ora)


@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
def test_forced_outliers(oracle, sweep):
    # F9: the reset replaces x and the last lpost row only (lp, ll, fx go stale), and the generation's output row is
    # written BEFORE the reset
    x0 = H.x0_tdist(16, 3)
    x0[3] = 400.0; x0[11] = -350.0
    cfg = A.default_config(nc=16, t=120, d=3, lik=2, seed=21, adapt=0.5, sweep=sweep, store_fx=1)
    py, ora = _both(oracle, cfg, "t_distribution", x0)
    assert len(ora["outlier"]) > 0
    _agree(py, ora)


@pytest.mark.parametrize("bound", [A.HB_BOUND_BOUND, A.HB_BOUND_REFLECT, A.HB_BOUND_FOLD])
@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
def test_hydromodel_glue_bounds_uniform_prior(oracle, bound, sweep):
    P, obs = H.hydro_data(T=60)
    cfg = A.default_config(nc=10, t=100, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=bound, seed=3,
                           adapt=0.4, sweep=sweep)
    py, ora = _both(oracle, cfg, "HydroModel", H.x0_hydro(10), par_min=[0.01], par_max=[0.6], forcing=P, obs=obs)
    _agree(py, ora, fx=False)


def test_uniform_prior_without_bound_handling(oracle):
    # prior = "uniform" with min/max but bound = NULL: proposals leave the box, dunif gives -Inf -> floored (F5)
    nc, d = 12, 3
    cfg = A.default_config(nc=nc, t=80, d=d, lik=2, seed=5, prior=A.HB_PRIOR_UNIFORM)
    py, ora = _both(oracle, cfg, "t_distribution", H.x0_tdist(nc, d) * 0.1, par_min=[-1.0, -2.0, -0.5], par_max=[1.5, 2.0, 0.5])
    _agree(py, ora)
    assert py["floored"] > 0                    # some proposals did leave the box and were floored at log(xmin)


def test_normal_prior(oracle):
    nc, d = 12, 3
    cov = np.array([[2.0, 0.3, 0.1], [0.3, 1.0, -0.2], [0.1, -0.2, 0.5]])
    cfg = A.default_config(nc=nc, t=80, d=d, lik=2, seed=9, prior=A.HB_PRIOR_NORMAL)
    py, ora = _both(oracle, cfg, "t_distribution", H.x0_tdist(nc, d) * 0.2, prior_mu=[0.5, -0.5, 0.0], prior_cov=cov)
    _agree(py, ora)


@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
@pytest.mark.parametrize("lik,abc", [(21, 0.8), (21, "vector"), (22, 2.5), (22, "vector"), (99, None)])
def test_lik_flags(oracle, lik, abc, sweep):
    T = 40
    P, obs = H.hydro_data(T=T)
    e = None if abc is None else (np.linspace(0.6, 3.0, T) if abc == "vector" else abc)
    cfg = A.default_config(nc=10, t=90, d=1, lik=lik, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_REFLECT, seed=29,
                           sweep=sweep, adapt=0.4)
    py, ora = _both(oracle, cfg, "HydroModel", H.x0_hydro(10), par_min=[0.01], par_max=[2.0], forcing=P, obs=obs, abc_e=e)
    _agree(py, ora, fx=False)


@pytest.mark.parametrize("psnooker", [0.0, 0.3])
def test_zs_archive_and_snooker(oracle, psnooker):
    nc, d, t = 3, 4, 120
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=17, past_sample=1, psnooker=psnooker)
    x0 = np.random.default_rng(2).uniform(-5, 15, size=(10 * d, d))
    py, ora = _both(oracle, cfg, "t_distribution", x0)
    _agree(py, ora)
    assert ora["outlier"] == []


@pytest.mark.parametrize("mt,d,nc", [(3, 3, 10), (2, 1, 8)])
def test_multi_try(oracle, mt, d, nc):
    if d == 1:
        cfg = A.default_config(nc=nc, t=100, d=1, lik=1, seed=37, mt=mt, adapt=0.3)
        py, ora = _both(oracle, cfg, "mixture", H.x0_mixture(nc))
    else:
        cfg = A.default_config(nc=nc, t=80, d=d, lik=2, seed=37, mt=mt, adapt=0.4)
        py, ora = _both(oracle, cfg, "t_distribution", H.x0_tdist(nc, d))
    _agree(py, ora)


def test_multi_try_zs_snooker(oracle):
    nc, d = 4, 3
    cfg = A.default_config(nc=nc, t=80, d=d, lik=2, seed=43, mt=2, past_sample=1, psnooker=0.2)
    x0 = np.random.default_rng(3).uniform(-5, 15, size=(10 * d, d))
    py, ora = _both(oracle, cfg, "t_distribution", x0)
    _agree(py, ora)


def test_check_convergence_rows_and_rstat(oracle):
    cfg = A.default_config(nc=8, t=70, d=2, lik=2, seed=23, check_convergence=1)
    py, ora = _both(oracle, cfg, "t_distribution", H.x0_tdist(8, 2))
    _agree(py, ora)
    np.testing.assert_allclose(ora["R_stat"], py["R_stat"], rtol=1e-10)
    i = np.arange(1, 101)[:, None, None]; k = np.arange(1, 3)[None, :, None]; c = np.arange(1, 5)[None, None, :]
    x = np.sin(0.37 * i * k + 1.3 * c) + 0.1 * c * k
    np.testing.assert_allclose(R.R_stat(x), [1.0113185866169283, 1.0735661244327646], rtol=1e-12)
    np.testing.assert_allclose(oracle.rstat(x), R.R_stat(x), rtol=1e-12)


@pytest.mark.parametrize("case", range(24))
def test_random_configurations(oracle, case):
    # seeded sweep over the option space (tests/helpers.py: random_case); the same cases run on the device in
    # tests/test_gpu_edge_cases.py::test_random_configurations_device
    cfg, target, x0, extra = H.random_case(case)
    py, ora = _both(oracle, cfg, target, x0, **extra)
    _agree(py, ora, fx=(target != "HydroModel"))
