"""Pins the CPU oracle against the known answers of SURVEY Appendix B (derived from the reference formulas,
the reference ships no tests) and against scipy for the third-party mvtnorm::dmvt arithmetic."""
import math

import numpy as np
import pytest

from hydrobayes_b200 import _abi as A


def test_floor_constant(oracle):
    # R/internals.R:19,65 : log(.Machine$double.xmin)
    assert oracle.log_xmin() == pytest.approx(-708.3964185322641, rel=0, abs=1e-12)


def test_philox_known_answers(oracle):
    # Random123 known-answer vectors for philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mixture(oracle):
    # R/test_mixture.R:16
    exp = {-8: 0.06649038006690544, 0: 8.420452447111373e-16, 10: 0.33245190033452726, 1: 1.0279773571668917e-18}
    for x, v in exp.items():
        assert oracle.mixture(x) == pytest.approx(v, rel=1e-13)
    logs = {-8: -2.7106980024327276, 0: -34.71069792628283, 10: -1.1012600899986273, 1: -41.418938533204674}
    for x, v in logs.items():
        ll, fx, rc = oracle.logdensity("mixture", 1, [[x]])
        assert rc == 0 and ll[0] == pytest.approx(v, rel=1e-13)


def test_mixture_floor(oracle):
    ll, fx, rc = oracle.logdensity("mixture", 1, [[60.0], [-70.0]])
    assert rc == 0
    assert np.all(ll == oracle.log_xmin())  # density underflows to 0 -> log = -Inf -> floored


def test_t_distribution_known(oracle):
    # R/test_t_distribution.R:19-46 ; SURVEY Appendix B
    assert oracle.tdist_logpdf(np.zeros((1, 1)))[0] == pytest.approx(-0.9231050070343514, rel=1e-13)
    assert oracle.tdist_logpdf(np.zeros((1, 2)))[0] == pytest.approx(-2.040609620463428, rel=1e-13)
    assert oracle.tdist_logpdf(np.zeros((1, 100)))[0] == pytest.approx(-213.43955272792834, rel=1e-13)
    assert oracle.tdist_logpdf(np.ones((1, 1)))[0] == pytest.approx(-1.4272487165462735, rel=1e-13)
    assert oracle.tdist_logpdf(np.ones((1, 2)))[0] == pytest.approx(-2.582068622322943, rel=1e-13)
    assert oracle.tdist_logpdf(np.ones((1, 100)))[0] == pytest.approx(-218.01512992854524, rel=1e-13)
    assert oracle.tdist_logpdf((np.arange(1, 3) / 2)[None])[0] == pytest.approx(-2.3125212751191637, rel=1e-13)
    assert oracle.tdist_logpdf((np.arange(1, 101) / 100)[None])[0] == pytest.approx(-213.5956709533909, rel=1e-13)


def test_t_distribution_vs_scipy(oracle):
    from scipy.stats import multivariate_t
    rng = np.random.default_rng(7)
    for d in (1, 3, 10, 100):
        i = np.arange(1, d + 1)
        Cm = (0.5 * np.eye(d) + 0.5) * np.sqrt(np.outer(i, i))
        x = rng.uniform(-5, 15, size=(16, d))
        ref = multivariate_t(loc=np.zeros(d), shape=Cm, df=60).logpdf(x)
        got = oracle.tdist_logpdf(x)
        np.testing.assert_allclose(got, np.atleast_1d(ref), rtol=1e-12)


@pytest.mark.parametrize("d", [1, 2, 3, 17, 64, 100, 130])
def test_t_distribution_structured_identity(oracle, d):
    """The device evaluates the reference's t_distribution in structured form (hydrobayes_b200/csrc/hb_targets.cu,
    hb_common.cuh: hb_tdist_acc / hb_warp_sum2 / hb_tdist_logdens): C = S (I/2 + 11'/2) S, S = diag(sqrt(i))
    (R/test_t_distribution.R:32-40), so by Sherman-Morrison x' C^-1 x = 2 (sum u^2 - (sum u)^2 / (d + 1)), u_i = x_i / sqrt(i),
    and log|C| = sum(log i) - d log 2 + log(d + 1).  Restated here in numpy with the device's order of operations (lane L of a
    warp takes the coordinate pairs 2 (32 m + L); the two halves of the warp are added first, then strides 8, 4, 2, 1) and
    checked against the oracle's dmvt (Cholesky + triangular solve) and against scipy's multivariate_t: the identity is
    pinned on the CPU, the GPU tests pin the kernels to it."""
    from scipy.stats import multivariate_t
    rng = np.random.default_rng(100 + d)
    x = np.concatenate([rng.uniform(-5, 15, size=(64, d)), rng.normal(size=(32, d)) * np.sqrt(np.arange(1, d + 1)),
                        np.zeros((1, d)), np.outer([1.0, -3.7, 30.0], np.sqrt(np.arange(1, d + 1)))])   # last rows: worst case for cancellation
    rs = 1.0 / np.sqrt(np.arange(1, d + 1, dtype=np.float64))
    df = 60.0
    ll = np.empty(x.shape[0])
    logdet = np.sum(np.log(np.arange(1, d + 1))) - d * math.log(2.0) + math.log(d + 1.0)
    konst = math.lgamma((d + df) / 2) - (math.lgamma(df / 2) + 0.5 * logdet + d / 2 * math.log(math.pi * df))
    for r, row in enumerate(x):
        s1 = np.zeros(32); s2 = np.zeros(32)
        for lane in range(32):
            for k0 in range(2 * lane, d, 64):
                u0 = row[k0] * rs[k0]; u1 = row[k0 + 1] * rs[k0 + 1] if k0 + 1 < d else 0.0
                s1[lane] = s1[lane] + (u0 + u1)
                s2[lane] = s2[lane] + u0 * u0                      # (the device uses fused multiply-adds here: one rounding less)
                s2[lane] = s2[lane] + u1 * u1
        def tree(v):
            v = v[:16] + v[16:]
            for o in (8, 4, 2, 1):
                v = v[:o] + v[o:2 * o]
            return v[0]
        a, b = tree(s1), tree(s2)
        rss = 2.0 * (b - a / (d + 1) * a)
        ll[r] = konst - 0.5 * (df + d) * math.log1p(rss / df)
    np.testing.assert_allclose(ll, oracle.tdist_logpdf(x), rtol=2e-13)
    i = np.arange(1, d + 1)
    Cm = (0.5 * np.eye(d) + 0.5) * np.sqrt(np.outer(i, i))
    np.testing.assert_allclose(ll, np.atleast_1d(multivariate_t(loc=np.zeros(d), shape=Cm, df=60).logpdf(x)), rtol=1e-12)
    assert abs(np.linalg.slogdet(Cm)[1] - logdet) < 1e-9 * max(1.0, abs(logdet))


def test_hydromodel(oracle):
    # R/test_HydroModel.R:17-36
    y = oracle.hydromodel([0, 1, 2, 0, 5], 0.5)
    np.testing.assert_allclose(y, [1 / 3, 5 / 9, 1.037037037037037, 0.691358024691358, 2.1275720164609053], rtol=1e-15)
    with pytest.raises(ValueError):
        oracle.hydromodel([0, 1], -0.1)
    with pytest.raises(ValueError):
        oracle.hydromodel([0, -1], 0.1)


def test_rstat_known(oracle):
    # R/gelman_rubin.R:21-56 ; SURVEY Appendix B deterministic array
    i = np.arange(1, 101)[:, None, None]
    k = np.arange(1, 3)[None, :, None]
    c = np.arange(1, 5)[None, None, :]
    x = np.sin(0.37 * i * k + 1.3 * c) + 0.1 * c * k
    r = oracle.rstat(x)
    np.testing.assert_allclose(r, [1.0113185866169283, 1.0735661244327646], rtol=1e-12)


def _rstat_numpy(x):
    t, d, n = x.shape
    y = x[math.ceil(t / 2) - 1:, :, :]
    xm = 2 / (t - 2) * y.sum(axis=0)              # [d, n]
    W = 2 / (n * (t - 2)) * ((y - xm[None]) ** 2).sum(axis=(0, 2))
    B = 1 / (2 * (n - 1)) * ((xm - xm.mean(axis=1, keepdims=True)) ** 2).sum(axis=1)
    sigma = (t - 2) / t * W + 2 * B
    return np.sqrt((n + 1) / n * sigma / W - (t - 2) / (n * t))


@pytest.mark.parametrize("t,d,n", [(51, 1, 3), (100, 2, 4), (201, 5, 16), (64, 100, 8)])
def test_rstat_vs_numpy(oracle, t, d, n):
    x = np.random.default_rng(t).normal(size=(t, d, n)).cumsum(axis=0)
    np.testing.assert_allclose(oracle.rstat(x), _rstat_numpy(x), rtol=1e-11)


def test_bound_par(oracle):
    # R/internals.R:89-113
    B, R, F = A.HB_BOUND_BOUND, A.HB_BOUND_REFLECT, A.HB_BOUND_FOLD
    assert oracle.bound_par(0.5, 0, 1, B) == 0.5
    assert oracle.bound_par(-0.2, 0, 1, B) == 0.0
    assert oracle.bound_par(1.7, 0, 1, B) == 1.0
    assert oracle.bound_par(-0.25, 0, 1, R) == 0.25
    assert oracle.bound_par(1.25, 0, 1, R) == 0.75
    assert oracle.bound_par(2.25, 0, 1, R) == pytest.approx(0.25)   # reflects twice
    assert oracle.bound_par(-0.25, 0, 1, F) == 0.75
    assert oracle.bound_par(1.25, 0, 1, F) == 0.25
    assert oracle.bound_par(3.25, 0, 1, F) == pytest.approx(0.25)
    assert oracle.bound_par(5.0, 0, 1, A.HB_BOUND_NONE) == 5.0


def test_quantile_type7(oracle):
    x = np.array([3.0, 1.0, 4.0, 1.5, 9.0, 2.6, 5.0])
    for p in (0.25, 0.75, 0.5, 0.0, 1.0):
        assert oracle.quantile7(x, p) == pytest.approx(np.quantile(x, p, method="linear"), rel=1e-15)
    y = np.random.default_rng(3).normal(size=1024)
    for p in (0.25, 0.75):
        assert oracle.quantile7(y, p) == pytest.approx(np.quantile(y, p), rel=1e-14)


def test_glue_likelihood(oracle):
    # R/internals.R:13-14 : glue_shape*log(max(1 - sum((res-obs)^2)/sum((obs-mean(obs))^2), 0))
    rng = np.random.default_rng(5)
    P = rng.gamma(0.7, 8, size=200) * (rng.uniform(size=200) < 0.4)
    obs = oracle.hydromodel(P, 0.4) + rng.normal(0, 0.05, size=200)
    for K in (0.3, 0.4, 1.5):
        sim = oracle.hydromodel(P, K)
        nse = 1 - ((sim - obs) ** 2).sum() / ((obs - obs.mean()) ** 2).sum()
        exp = max(50 * math.log(max(nse, 0)) if nse > 0 else -math.inf, oracle.log_xmin())
        ll, fx, rc = oracle.logdensity("HydroModel", 31, [[K]], forcing=P, obs=obs, glue_shape=50.0)
        assert rc == 0 and ll[0] == pytest.approx(exp, rel=1e-12)


def test_native_rng_transforms(oracle):
    """slot map v2 (include/hydrobayes_b200_rng.h): 64 random bits per dimension.  Independent numpy restatement of the
    three transforms + the properties the sampler relies on (open interval, symmetry, exact integer crossover test)."""
    L = oracle.lib()
    hs = [0, 1, 21845, 21846, 32767, 32768, 43690, 43691, 65535]
    for h in hs:
        assert L.hbo_rng_u16(h) == (h + 0.5) / 65536.0
        assert L.hbo_rng_lambda16(h, 0.1) == -0.1 + (2.0 * 0.1) * ((h + 0.5) / 65536.0)
        assert L.hbo_rng_lambda16(h, 0.1) == pytest.approx(-L.hbo_rng_lambda16(65535 - h, 0.1), abs=1e-16)   # symmetric about 0
    assert 0.0 < L.hbo_rng_u16(0) and L.hbo_rng_u16(65535) < 1.0
    # zt < CR  <=>  h < threshold, for every h and every CR of nCR = 1..6
    h = np.arange(65536)
    zt = (h + 0.5) / 65536.0
    for nCR in range(1, 7):
        for q in range(nCR):
            cr = (q + 1) / nCR
            thr = L.hbo_rng_zt_threshold16(cr)
            assert np.array_equal(zt < cr, h < thr), (nCR, q)
    # the integer-CLT deviate: restated bit by bit, then its first four moments over a fixed word set
    rs = np.random.RandomState(7)
    w = rs.randint(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)
    z = np.array([L.hbo_rng_normal32(int(x)) for x in w[:5000]])
    b = np.array([bin(int(x) & 0xffffff).count("1") for x in w[:5000]])
    t = ((w[:5000] >> 24) & 0xf).astype(int) + (w[:5000] >> 28).astype(int) - 15
    np.testing.assert_array_equal(z, ((b - 12) * 16 + t) * float.fromhex("0x1.9c614ae8b41e8p-6"))
    assert L.hbo_rng_normal32(0x00000000) == (-12 * 16 - 15) * float.fromhex("0x1.9c614ae8b41e8p-6")
    assert L.hbo_rng_normal32(0xffffffff) == (12 * 16 + 15) * float.fromhex("0x1.9c614ae8b41e8p-6")
    assert abs(z.mean()) < 0.05 and abs(z.var() - 1.0) < 0.06


def _hydro_np(P, K):
    S = 1.0; Y = np.empty(len(P))
    for i, p in enumerate(P):
        S = (p + S) / (1 + K); Y[i] = K * S
    return Y


def test_calc_ll_abc_and_sls_flags(oracle):
    # calc_ll, R/internals.R:8-16, restated independently in numpy: lik 21 (sum(log(abc_e)) over abc_e AS GIVEN: a scalar
    # abc_e is counted once, R recycles only inside the squared term), lik 22 = min(abc_e - rho), lik 99 = fun_lik (SLS)
    import helpers as H
    T = 40
    P, obs = H.hydro_data(T=T)
    Ks = np.array([[0.05], [0.3], [1.2]])
    ev = np.linspace(0.5, 2.0, T)
    fl = oracle.log_xmin()
    for K in Ks[:, 0]:
        Y = _hydro_np(P, K); rho = np.abs(Y - obs)
        exp21s = -T / 2 * np.log(2 * np.pi) - np.log(0.8) - 0.5 * np.sum((rho / 0.8) ** 2)
        exp21v = -T / 2 * np.log(2 * np.pi) - np.sum(np.log(ev)) - 0.5 * np.sum((rho / ev) ** 2)
        exp22s = np.min(2.5 - rho); exp22v = np.min(ev - rho)
        exp99 = -T / 2 * np.log(np.sum((Y - obs) ** 2))
        x = np.array([[K]])
        for lik, e, exp in ((21, 0.8, exp21s), (21, ev, exp21v), (22, 2.5, exp22s), (22, ev, exp22v), (99, None, exp99)):
            ll, _, rc = oracle.logdensity("HydroModel", lik, x, forcing=P, obs=obs, abc_e=e)
            assert rc == 0
            assert ll[0] == pytest.approx(max(exp, fl), rel=1e-12)


def test_multi_try_oracle_moments(oracle):
    # MT-DREAM (R/dream_parallel.R:295-341) must leave the posterior unchanged and raise the acceptance rate
    import helpers as H
    from hydrobayes_b200 import _abi as A
    acc = {}
    for mt in (1, 4):
        cfg = A.default_config(nc=32, t=1500, d=3, lik=2, seed=8, mt=mt, record_accept=1)
        r = oracle.run(cfg, "t_distribution", H.x0_tdist(32, 3))
        assert r["rc"] == 0, r["err"]
        acc[mt] = r["accept"][500:].mean()
        v = r["chain"][500:, :3, :].var(axis=(0, 2))
        np.testing.assert_allclose(v, 60.0 / 58.0 * np.arange(1, 4), rtol=0.15)
    assert acc[4] > 1.3 * acc[1]


def test_third_party_arithmetic_against_multiprecision(oracle):
    """dnorm (R nmath, split-exponent branch for |z| >= 5) and mvtnorm::dmvt are restated from recollection of their published
    sources; this pins the restatement to the EXACT values of the formulas, evaluated with 50 digits."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    xs = np.concatenate([np.linspace(-30, 32, 311), np.random.default_rng(0).uniform(-30, 32, 200)])
    for x in xs:
        X = mp.mpf(float(x))
        ref = mp.mpf(1) / 6 * mp.npdf(X, -8, 1) + mp.mpf(5) / 6 * mp.npdf(X, 10, 1)          # R/test_mixture.R:16
        if ref > mp.mpf(2.3e-308):
            assert abs((mp.mpf(oracle.mixture(float(x))) - ref) / ref) < 4e-15, x

    def dmvt_log(x):                                                                          # R/test_t_distribution.R:19-46
        d = len(x); df = mp.mpf(60)
        C = mp.matrix(d, d)
        for a in range(d):
            for b in range(d):
                C[a, b] = (mp.mpf(1) if a == b else mp.mpf("0.5")) * mp.sqrt(mp.mpf(a + 1) * mp.mpf(b + 1))
        xv = mp.matrix([mp.mpf(float(v)) for v in x])
        q = (xv.T * mp.lu_solve(C, xv))[0]
        return (mp.loggamma((df + d) / 2) - mp.loggamma(df / 2) - mp.mpf(d) / 2 * mp.log(df * mp.pi) - mp.log(mp.det(C)) / 2
                - (df + d) / 2 * mp.log(1 + q / df))
    rng = np.random.default_rng(1)
    for d in (1, 2, 5, 17, 40):
        x = rng.uniform(-5, 15, (2, d))
        got = oracle.tdist_logpdf(x)
        for r in range(2):
            ref = dmvt_log(x[r])
            assert abs((mp.mpf(float(got[r])) - ref) / ref) < 2e-14, (d, r)
