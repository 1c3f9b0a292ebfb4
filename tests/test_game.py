"""game() -- GAME sampling of the model evidence (R/game.R:79-209).  The mixture fit is not pinned to the reference (EMCluster
is third-party and randomly initialised); the estimators are: for a NORMALISED posterior density the evidence is 1."""
import numpy as np
import pytest

import helpers as H
from hydrobayes_b200 import _abi as A
from hydrobayes_b200.game import game


def _gauss_sample(n=6000, seed=0):
    rng = np.random.default_rng(seed)
    mu = np.array([1.0, -2.0]); L = np.array([[1.0, 0.0], [0.6, 0.8]])
    x = mu + rng.standard_normal((n, 2)) @ L.T
    C = L @ L.T
    q = np.einsum("ij,jk,ik->i", x - mu, np.linalg.inv(C), x - mu)
    return x, np.exp(-0.5 * q) / (2 * np.pi * np.sqrt(np.linalg.det(C)))


def test_game_reciprocal_importance_sampling_on_a_normalised_density():
    theta, post = _gauss_sample()
    res = game(theta, post, m1=3000, Jmax=3, ic="bic", omega=0, seed=1)           # no target evaluation: host only
    assert abs(res["Z"] - 1.0) < 0.05 and 1 <= res["J"] <= 3
    res = game(theta, 3.5 * post, m_sub=2000, m1=3000, Jmax=2, ic="var", omega=0, seed=2)   # unnormalised by 3.5 -> Z = 3.5
    assert abs(res["Z"] / 3.5 - 1.0) < 0.05


def test_game_argument_checks_follow_the_reference():
    theta, post = _gauss_sample(200)
    with pytest.raises(ValueError, match="must in \\[0,1\\]"):
        game(theta, post, m1=50, ic="bic", omega=1.5)
    with pytest.raises(ValueError, match="'m1' must not be larger"):
        game(theta, post, m1=500, ic="bic", omega=0)
    with pytest.raises(ValueError, match="'m0' needs to be given"):
        game(theta, post, m1=50, ic="bic", omega=0.5)
    with pytest.raises(ValueError, match="one of \\{'bic', 'var'\\}"):
        game(theta, post, m1=50, ic="aic", omega=0)


@pytest.mark.gpu
def test_game_importance_and_bridge_sampling_with_the_device_target():
    """posterior sample of the d = 2 t-distribution from dream_parallel (flat prior: the posterior density is the normalised
    target density, so Z = 1); the m0 target evaluations of IS / bridge sampling run through the device plugin"""
    d, nc, t = 2, 64, 1500
    from hydrobayes_b200 import dream_parallel
    res = dream_parallel("t_distribution", lik=2, par_info=dict(initial="user", val_ini=H.x0_tdist(nc, d) * 0.2, prior="flat"),
                            nc=nc, t=t, d=d, seed=11, verbose=False)
    ch = res["chain"][t // 2:]
    theta = ch[:, :d, :].transpose(0, 2, 1).reshape(-1, d)
    post = np.exp(ch[:, d + 2, :].reshape(-1))
    par = dict(prior="flat")
    for omega, ob in ((1.0, False), (0.5, False), (0.5, True), (0.0, False)):
        g = game(theta, post, m_sub=8000, m0=20000, m1=8000, Jmax=4, ic="bic", omega=omega, ob=ob, par_info=par, lik=2,
                 fun="t_distribution", seed=5)
        assert abs(g["Z"] - 1.0) < 0.08, (omega, ob, g)


@pytest.mark.gpu
def test_de_mc_is_the_sequential_sweep_with_one_pair_and_no_crossover(oracle):
    """R/de_mc.R:23-70 as a configuration of the dream() device path: replay parity of that configuration with the oracle, and
    the mixture target's moments (mean 7, P(x < 1) = 1/6)"""
    from hydrobayes_b200.game import de_mc
    nc, t = 8, 300
    cfg = A.default_config(sweep=A.HB_SWEEP_SEQUENTIAL, nc=nc, t=t, d=1, lik=1, delta=1, nCR=1, c_val=0.0, c_star=1e-6, p_g=0.1,
                           adapt=0.0, outlier_check=0, seed=3, record_accept=1)
    dev = H.run_device(cfg, "mixture", H.x0_mixture(nc))
    ora = oracle.run(cfg, "mixture", H.x0_mixture(nc))
    H.assert_run_parity(dev, ora, rtol=1e-9, atol=1e-12, check_fx=False)
    nc, t = 64, 4000
    res = de_mc("mixture", lik=1, par_info=dict(initial="user", val_ini=H.x0_mixture(nc), prior="flat"), nc=nc, t=t, d=1, seed=4)
    assert res["chain"].shape == (t, 1, nc) and res["density"].shape == (t, nc)
    x = res["chain"][t // 2:, 0, :]
    m = x.mean(axis=0); p = (x < 1).mean(axis=0)
    assert abs(m.mean() - 7) < 4 * m.std(ddof=1) / np.sqrt(nc) + 0.1
    assert abs(p.mean() - 1 / 6) < 4 * p.std(ddof=1) / np.sqrt(nc) + 0.01
