"""rwm() / am() / am_sd() (R/rwm.R, R/am.R, R/am_sd.R) over the device log-density plugin.  Not pinned to the reference (no R,
different generator): the host logic is checked on CPU against a stand-in target, the device path against the targets' moments."""
import numpy as np
import pytest

import hydrobayes_b200.single_chain as SC
from hydrobayes_b200 import am, am_sd, rwm


def _std_normal_logpdf(fun, x, lik, forcing=None, obs=None, glue_shape=0.0):
    x = np.atleast_2d(x)
    return -0.5 * (x ** 2).sum(axis=1) - 0.5 * x.shape[1] * np.log(2 * np.pi), None


@pytest.mark.parametrize("sampler", [rwm, am, am_sd])
def test_single_chain_samplers_host_logic(monkeypatch, sampler):
    # the loop of the reference with the device evaluation replaced by a standard normal: shapes of the return list, the
    # first row is the initial sample, rejected iterations repeat the previous row, moments of the stationary law
    import hydrobayes_b200.sampler as S
    monkeypatch.setattr(S, "log_density", _std_normal_logpdf)
    d, t, n = 3, 3000, 16
    x0 = np.linspace(-1, 1, n * d).reshape(n, d)
    res = sampler("stand_in", lik=2, par_info=dict(initial="user", val_ini=x0, prior="flat"), t=t, d=d, n_chains=n, seed=5)
    assert res["chain"].shape == (t, d, n) and res["density"].shape == (t, n)
    assert np.array_equal(res["chain"][0], x0.T)
    moved = np.any(res["chain"][1:] != res["chain"][:-1], axis=1)
    assert 0.1 < moved.mean() < 0.7
    same = ~moved
    assert np.array_equal(res["density"][1:][same], res["density"][:-1][same])
    x = res["chain"][t // 3:]
    assert abs(x.mean()) < 0.1 and abs(x.var() - 1.0) < 0.15
    one = sampler("stand_in", lik=2, par_info=dict(initial="user", val_ini=x0[:1], prior="flat"), t=50, d=d, seed=1)
    assert one["chain"].shape == (50, d) and one["density"].shape == (50,)           # the reference's structure for one chain


def test_am_sd_scales_towards_the_target_acceptance_rate(monkeypatch):
    import hydrobayes_b200.sampler as S
    monkeypatch.setattr(S, "log_density", lambda fun, x, lik, **kw: (-0.5 * (np.atleast_2d(x) ** 2).sum(axis=1) / 1e-4, None))
    res = am_sd("narrow", lik=2, par_info=dict(initial="user", val_ini=np.zeros((8, 2)), prior="flat"), t=4000, d=2, n_chains=8, seed=2)
    early = np.any(res["chain"][1:200] != res["chain"][:199], axis=1).mean()
    late = np.any(res["chain"][3000:] != res["chain"][2999:-1], axis=1).mean()
    # sd 0.01 target: the initial scale 2.38^2 / 2 accepts almost nothing; the cumulative acceptance rate stays below 20 %, so
    # s_d shrinks every 25th iteration until C = (s_d + 1e-4) I bottoms out at sd 0.01 (R/am_sd.R:44-53)
    assert early < 0.02 and 0.3 < late < 0.7


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", [rwm, am, am_sd])
def test_single_chain_samplers_on_the_device_target(sampler):
    """d = 2 t-distribution (mean 0, Var(x_i) = 60/58 i) through the device plugin, 64 replicas"""
    import helpers as H
    d, t, n = 2, 1500, 64
    res = sampler("t_distribution", lik=2, par_info=dict(initial="user", val_ini=0.2 * H.x0_tdist(n, d), prior="flat"), t=t, d=d,
                  n_chains=n, seed=3)
    x = res["chain"][t // 2:]
    m = x.mean(axis=0)                                            # [d, n]
    assert np.all(np.abs(m.mean(axis=1)) < 4 * m.std(axis=1, ddof=1) / np.sqrt(n) + 0.05)
    v = x.var(axis=0)
    for k in range(d):
        assert abs(v[k].mean() / (60 / 58 * (k + 1)) - 1) < 0.25
    assert np.all(np.isfinite(res["density"])) and np.all(res["density"] > 0)
