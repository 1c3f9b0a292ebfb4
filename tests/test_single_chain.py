"""rwm() / am() / am_sd() (R/rwm.R, R/am.R, R/am_sd.R) over the device log-density plugin.  Not pinned to the reference (no R,
different generator): the host logic is checked on CPU against a stand-in target, the device path against the targets' moments."""
import numpy as np
import pytest

import hydrobayes_b200.single_chain as SC
from hydrobayes_b200 import am, am_sd, rwm


def _stand_in_session(logpdf):
    """a DensitySession that evaluates `logpdf` on the host: same constructor, call and context-manager protocol"""
    class Session:
        def __init__(self, fun, lik, d, max_n, forcing=None, obs=None, glue_shape=0.0):
            self.d, self.max_n = d, max_n

        def __call__(self, x):
            x = np.atleast_2d(x)
            assert x.shape[1] == self.d and x.shape[0] <= self.max_n
            return logpdf(x), None

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass
    return Session


def _std_normal_logpdf(x):
    return -0.5 * (x ** 2).sum(axis=1) - 0.5 * x.shape[1] * np.log(2 * np.pi)


@pytest.mark.parametrize("sampler", [rwm, am, am_sd])
def test_single_chain_samplers_host_logic(monkeypatch, sampler):
    # the loop of the reference with the device evaluation replaced by a standard normal: shapes of the return list, the
    # first row is the initial sample, rejected iterations repeat the previous row, moments of the stationary law
    import hydrobayes_b200.sampler as S
    monkeypatch.setattr(S, "DensitySession", _stand_in_session(_std_normal_logpdf))
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
    monkeypatch.setattr(S, "DensitySession", _stand_in_session(lambda x: -0.5 * (x ** 2).sum(axis=1) / 1e-4))
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


@pytest.mark.gpu
def test_density_session_equals_the_one_shot_evaluation():
    """hb_density_open / _eval / _close against hb_logdensity: same plugin, same kernels -> array_equal, call after call"""
    import helpers as H
    from hydrobayes_b200 import HbError, log_density
    from hydrobayes_b200.sampler import DensitySession
    rng = np.random.default_rng(11)
    for d in (1, 2, 100):
        x = rng.normal(0, 3, (257, d))
        want, want_fx = log_density("t_distribution", x, 2)
        with DensitySession("t_distribution", 2, d, 257) as s:
            for n in (257, 1, 33, 257):                              # buffers are reused: smaller batches after a full one
                ll, fx = s(x[:n])
                if n == 257:
                    assert np.array_equal(ll, want) and np.array_equal(fx, want_fx)
                else:
                    np.testing.assert_allclose(ll, want[:n], rtol=1e-12)
            with pytest.raises(HbError):
                s(np.zeros((258, d)))                                # more points than the session was opened for
    P, obs = H.hydro_data(T=200)
    k = np.linspace(0.05, 1.5, 64)[:, None]
    want, _ = log_density("HydroModel", k, 31, forcing=P, obs=obs, glue_shape=50.0)
    with DensitySession("HydroModel", 31, 1, 64, forcing=P, obs=obs, glue_shape=50.0) as s:
        assert np.array_equal(s(k)[0], want)
        np.testing.assert_allclose(s(k[:5])[0], want[:5], rtol=1e-12)
        assert np.array_equal(s(k)[0], want)
    xm = np.linspace(-12, 14, 50)[:, None]
    with DensitySession("mixture", 1, 1, 50) as s:
        assert np.array_equal(s(xm)[0], log_density("mixture", xm, 1)[0])
    with pytest.raises(HbError):
        DensitySession("my_r_closure", 2, 3, 8)
