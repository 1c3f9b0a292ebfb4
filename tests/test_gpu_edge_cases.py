"""Edge cases of the generation loop on the device against the oracle (-m gpu): smallest legal sizes, degenerate
option values, the boundaries of the kernel-selection rules.  The reference has no tests; these follow what
its code does at the edges of its own argument checks (R/dream.R:136-137, R/dream_parallel.R:166-175, 195-198, 263, 403)."""
import numpy as np
import pytest

import helpers as H
from hydrobayes_b200 import _abi as A

pytestmark = pytest.mark.gpu

SYN, SEQ = A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL


@pytest.fixture(scope="module")
def hb():
    import hydrobayes_b200
    assert hydrobayes_b200.lib().hb_device_count() >= 1
    return hydrobayes_b200


def _native(oracle, cfg, target, x0, rtol=1e-9, **kw):
    cfg = H.copy_cfg(cfg, record_accept=1)
    ora = oracle.run(cfg, target, x0, **kw)
    assert ora["rc"] == 0, ora["err"]
    dev = H.run_device(cfg, target, x0, **kw)
    H.assert_run_parity(dev, ora, rtol=rtol, atol=1e-12, check_fx=False)
    return dev, ora


@pytest.mark.parametrize("sweep", [SYN, SEQ])
@pytest.mark.parametrize("kw", [
    dict(t=2),                                        # one generation: `for(i in 2:t)` runs once
    dict(t=3, thin=3),                                # out_t = 2: the initial row and generation 3
    dict(t=40, adapt=0.0),                            # no adaptation period at all: pCR stays 1/nCR, no outlier check
    dict(t=40, adapt=1.0),                            # adaptation (and outlier checks) up to the last generation
    dict(t=30, adapt=1.0, update_interval=1),         # pCR update + outlier check every generation (window ceiling(i/2):i)
    dict(t=25, update_interval=40),                   # updateInterval > t: AR/CR keep their first row only
    dict(t=70, thin=7),                               # out_t = 11
    dict(t=100, burnin=0.5, thin=5),                  # out_t = 11, first stored generation is 55
    dict(t=30, burnin=0.5),                           # out_t = 16, first stored generation is 16
    dict(t=30, outlier_check=0, adapt=0.5),
    dict(t=30, nCR=16, delta=4, adapt=0.5),           # the device limits (HB_MAX_NCR, HB_MAX_DELTA); nc must be > 8
    dict(t=30, p_g=0.0), dict(t=30, p_g=1.0),         # gamma never / always 1
    dict(t=30, c_val=0.0, c_star=0.0),                # no lambda / zeta noise
])
def test_option_edges(hb, oracle, sweep, kw):
    nc, d = 9, 3
    cfg = A.default_config(nc=nc, d=d, lik=2, seed=101, sweep=sweep, **kw)
    dev, ora = _native(oracle, cfg, "t_distribution", H.x0_tdist(nc, d))
    out_t = dev["chain"].shape[0]
    assert not np.isnan(dev["chain"][out_t - 1]).any()          # the last output row was reached


@pytest.mark.parametrize("sweep", [SYN, SEQ])
def test_smallest_population(hb, oracle, sweep):
    # nc = 2 * delta + 1 is the smallest population the argument check admits: every chain's partner pool (nc - 1 = 2 delta)
    # is exhausted when D = delta
    for delta in (1, 3):
        nc = 2 * delta + 1
        cfg = A.default_config(nc=nc, t=60, d=2, lik=2, seed=7, delta=delta, sweep=sweep, adapt=0.5)
        _native(oracle, cfg, "t_distribution", H.x0_tdist(nc, 2))
    from hydrobayes_b200 import Sampler
    with pytest.raises(A.HbError, match="'nc' must be > 'delta' \\* 2"):
        Sampler(A.default_config(nc=6, t=10, d=2, lik=2, sweep=sweep), 0)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129])
def test_dimension_boundaries_of_the_kernel_variants(hb, oracle, d):
    # lanes-per-chain variants switch at d = 1, 2, 4, 8, 16, 32, 64, 128 (hb_launch_k1); the row pitch is d rounded up to 4
    nc = 10
    cfg = A.default_config(nc=nc, t=25, d=d, lik=2, seed=d, adapt=0.4, update_interval=5)
    _native(oracle, cfg, "t_distribution", H.x0_tdist(nc, d))


@pytest.mark.parametrize("nc", [37888 - 32, 37888, 37888 + 17])
def test_population_size_where_the_wide_kernels_take_over(hb, oracle, nc):
    # from sm_count * 256 chains (37 888 on a B200) the warp-per-32-chains proposal / accept kernels are selected without
    # the test hook; ragged last batch (nc not a multiple of 32); outlier check at this size (device sort + select)
    d = 20
    cfg = A.default_config(nc=nc, t=12, d=d, lik=2, seed=5, adapt=0.5, update_interval=5, thin=4)
    x0 = H.x0_tdist(nc, d)
    x0[123] = 250.0
    dev, ora = _native(oracle, cfg, "t_distribution", x0, rtol=1e-9)
    assert (5, 123) in ora["outlier"]


def test_sequential_sweep_largest_run(hb, oracle):
    # the sequential sweep admits nc <= 2048 chains per run (one K1/K2/K3 round per chain)
    nc, d = 2048, 2
    cfg = A.default_config(nc=nc, t=3, d=d, lik=2, seed=3, sweep=SEQ)
    _native(oracle, cfg, "t_distribution", H.x0_tdist(nc, d))
    from hydrobayes_b200 import Sampler
    with pytest.raises(A.HbError, match="nc <= 2048"):
        Sampler(A.default_config(nc=2049, t=3, d=d, lik=2, sweep=SEQ), 0)


@pytest.mark.parametrize("kw", [dict(t=30, thin=4, burnin=0.1),      # 27/4 + 1 = 7.75 -> 7 rows, 8 stores wanted
                                dict(t=30, burnin=0.9)])              # (1 - 0.9) * 30 + 1 = 3.9999999999999996 -> 3 rows, 4 stores
def test_non_integer_out_t_is_the_callers_problem(hb, oracle, kw):
    # (1 - burnin) * t / thin not an integer: R's array() truncates the row count and a late store fails with "subscript
    # out of bounds" (R/dream_parallel.R:195-203, :388); oracle and device report the same condition (no write past the end)
    from hydrobayes_b200 import Sampler
    cfg = A.default_config(nc=8, d=2, lik=2, **kw)
    x0 = H.x0_tdist(8, 2)
    ora = oracle.run(cfg, "t_distribution", x0)
    assert ora["rc"] != 0 and "subscript out of bounds" in ora["err"]
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        with pytest.raises(A.HbError, match="subscript out of bounds"):
            s.run(cfg.t)


@pytest.mark.parametrize("sweep", [SYN, SEQ])
def test_identical_chains_give_zero_spread(hb, oracle, sweep):
    # all chains start in the same point: every jump differential is 0 and std_x = 0, so R's J becomes NaN and
    # `if(any(J > 0))` stops with "missing value where TRUE/FALSE needed"  -- during adaptation only
    nc, d = 8, 2
    x0 = np.tile([[1.0, -2.0]], (nc, 1))
    cfg = A.default_config(nc=nc, t=30, d=d, lik=2, seed=1, sweep=sweep, adapt=0.5, c_star=0.0)
    ora = oracle.run(cfg, "t_distribution", x0)
    from hydrobayes_b200 import Sampler
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        if ora["rc"] != 0:
            with pytest.raises(A.HbError, match="missing value where TRUE/FALSE needed"):
                s.run(30)
        else:
            s.run(30)
            np.testing.assert_allclose(s.fetch_chain(), ora["chain"], rtol=1e-12, equal_nan=True)
    cfg0 = A.default_config(nc=nc, t=30, d=d, lik=2, seed=1, sweep=sweep, adapt=0.0, c_star=0.0)   # outside adaptation J is never read
    _native(oracle, cfg0, "t_distribution", x0)


@pytest.mark.parametrize("case", range(24))
def test_random_configurations_device(hb, oracle, case):
    # the seeded option-space sweep of tests/test_restatement_crosscheck.py, replayed on the device: the oracle records
    # the decision tape of its Philox run, oracle and device replay it (accept/reject identical, chain within 1e-10)
    cfg, target, x0, extra = H.random_case(case)
    rec = oracle.Tape.for_config(cfg)
    r0 = oracle.run(cfg, target, x0, record=rec, **extra)
    assert r0["rc"] == 0, r0["err"]
    cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_TAPE, record_accept=1)
    ora = oracle.run(cfg_t, target, x0, tape=rec, **extra)
    dev = H.run_device(cfg_t, target, x0, tape=rec, **extra)
    H.assert_run_parity(dev, ora, rtol=1e-10, check_fx=(target != "HydroModel"))
    # and with native Philox draws
    cfg_n = H.copy_cfg(cfg, record_accept=1)
    H.assert_run_parity(H.run_device(cfg_n, target, x0, **extra), oracle.run(cfg_n, target, x0, **extra), rtol=1e-9, atol=1e-12,
                        check_fx=(target != "HydroModel"))


@pytest.mark.parametrize("bound", [A.HB_BOUND_NONE, A.HB_BOUND_BOUND, A.HB_BOUND_REFLECT, A.HB_BOUND_FOLD])
@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("d", [20, 40, 100])
@pytest.mark.parametrize("persist", ["1", "0"])
def test_fused_small_population_generation(hb, oracle, monkeypatch, bound, native, d, persist):
    # Small populations outside the adaptation period (the default path for them): proposal + log-density + accept fused
    # with a double-buffered population -- as ONE cooperative launch per slice with a grid barrier per generation
    # (hb_persist_small_kernel, HB_PERSIST unset / 1) or one launch per generation (hb_fused_small_kernel, HB_PERSIST=0):
    # parity with the oracle, and identical decisions to the three-kernel path (HB_FUSED=0)
    monkeypatch.setenv("HB_PERSIST", persist)
    nc, t = 37, 90
    kw = {}
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=9 + d, adapt=0.2, thin=3, store_fx=1, record_accept=1,
                           prior=A.HB_PRIOR_UNIFORM if bound else A.HB_PRIOR_FLAT, bound=bound)
    if bound:
        kw.update(par_min=list(-3.0 - 0.01 * np.arange(d)), par_max=list(4.0 + 0.02 * np.arange(d)))    # the box bites
    x0 = 0.3 * H.x0_tdist(nc, d)
    monkeypatch.setenv("HB_FUSED", "1")
    if native:
        ora = oracle.run(cfg, "t_distribution", x0, **kw)
        dev = H.run_device(cfg, "t_distribution", x0, **kw)
        H.assert_run_parity(dev, ora, rtol=1e-9, atol=1e-12)
        cfg_t, tape_kw = cfg, {}
    else:
        rec = oracle.Tape.for_config(cfg)
        assert oracle.run(cfg, "t_distribution", x0, record=rec, **kw)["rc"] == 0
        cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_TAPE)
        ora = oracle.run(cfg_t, "t_distribution", x0, tape=rec, **kw)
        dev = H.run_device(cfg_t, "t_distribution", x0, tape=rec, slice_generations=7, **kw)
        H.assert_run_parity(dev, ora, rtol=1e-11)
        tape_kw = dict(tape=rec)
    assert 0.03 < ora["accept"].mean() < 0.9
    monkeypatch.setenv("HB_FUSED", "0")                            # the same run through K1 / K2 / K3
    ref = H.run_device(cfg_t, "t_distribution", x0, **tape_kw, **kw)
    assert np.array_equal(ref["accept"], dev["accept"])
    np.testing.assert_allclose(dev["chain"], ref["chain"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(dev["AR"], ref["AR"])
    assert dev["timing"][1] < ref["timing"][1] - (t - 1 - int(0.2 * t)) , "the fused path was not taken"


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("d,nc,slice_g", [(100, 64, None), (40, 37, 7), (20, 200, 13)])
def test_persistent_kernel_in_adaptation_period(hb, oracle, monkeypatch, native, d, nc, slice_g):
    # The adaptation period of a small population inside the persistent kernel (hb_persist_small_kernel with adapt = 1):
    # per-CTA std_x / jump statistics, CTA 0 folds them and updates J / pCR / the AR-CR rows inside the grid barrier, one
    # cooperative launch per updateInterval with the outlier check (R/dream_parallel.R:414-419) between the launches.
    # Parity with the oracle (forced outlier resets included) and with the per-generation kernels (HB_PERSIST_ADAPT=0).
    t = 120
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=70 + d, adapt=0.6, thin=2, record_accept=1, update_interval=10)
    x0 = 0.3 * H.x0_tdist(nc, d); x0[1] = 250.0; x0[nc - 2] = -180.0          # two chains far out: outlier resets
    kw = dict(slice_generations=slice_g) if slice_g else {}
    monkeypatch.setenv("HB_PERSIST_ADAPT", "1")
    if native:
        cfg_t, tape_kw = cfg, {}
        ora = oracle.run(cfg, "t_distribution", x0)
        dev = H.run_device(cfg, "t_distribution", x0, **kw)
        H.assert_run_parity(dev, ora, rtol=1e-9, atol=1e-12)
    else:
        rec = oracle.Tape.for_config(cfg)
        assert oracle.run(cfg, "t_distribution", x0, record=rec)["rc"] == 0
        cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_TAPE)
        ora = oracle.run(cfg_t, "t_distribution", x0, tape=rec)
        tape_kw = dict(tape=rec)
        dev = H.run_device(cfg_t, "t_distribution", x0, **tape_kw, **kw)
        H.assert_run_parity(dev, ora, rtol=1e-10)
    assert len(ora["outlier"]) > 0
    monkeypatch.setenv("HB_PERSIST_ADAPT", "0")
    ref = H.run_device(cfg_t, "t_distribution", x0, **tape_kw, **kw)
    assert np.array_equal(ref["accept"], dev["accept"]) and ref["outlier"] == dev["outlier"]
    np.testing.assert_allclose(dev["chain"], ref["chain"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(dev["CR"], ref["CR"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(dev["AR"], ref["AR"])
    assert dev["timing"][1] < ref["timing"][1] - 3 * int(0.6 * t) // 2, "the adaptation period did not run in the persistent kernel"


def test_series_fx_opt_in(hb, oracle):
    """fx of a series target (HydroModel: m = T values per stored state, R/dream_parallel.R:230, 394) is not kept on the
    device; series_fx=True rebuilds it from the stored chain, and it equals the reference's fx wherever that is not stale
    (no outlier reset in this run, so everywhere)."""
    P, obs = H.hydro_data(T=60)
    nc, t = 12, 40
    par = dict(initial="user", val_ini=H.x0_hydro(nc), prior="uniform", min=[0.01], max=[2.0], bound="bound")
    kw = dict(lik=31, par_info=par, nc=nc, t=t, d=1, glue_shape=50.0, obs=obs, forcing=P, seed=3, verbose=False, outlier_check=False)
    res = hb.dream_parallel("HydroModel", **kw)
    assert res["fx"] is None
    res = hb.dream_parallel("HydroModel", series_fx=True, **kw)
    assert res["fx"].shape == (t, nc, 60)
    np.testing.assert_array_equal(res["fx"][7, 3], hb.HydroModel(P, res["chain"][7, 0, 3]))
    cfg = A.default_config(nc=nc, t=t, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_BOUND, seed=3, outlier_check=0)
    ora = oracle.run(cfg, "HydroModel", H.x0_hydro(nc), par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
    np.testing.assert_allclose(res["chain"], ora["chain"], rtol=1e-9, equal_nan=True)
