"""Sharded population on ONE GPU (-m gpu): the ranks of a multi-GPU run emulated as handles on one device
(hb_group_attach / hb_group_run, SURVEY section 4-iv "one kernel over all ranks' data").  The kernels, the delta log, the
send / apply kernels and the arena collectives are those of N processes on N GPUs; only the wiring differs (plain device
pointers instead of CUDA IPC mappings, one stream).  Every shard must equal the same chains of the single-handle run bit
for bit (native Philox draws are keyed by the global chain id), the population-wide statistics (AR, CR, R_stat) to
rounding, with outlier resets crossing shard borders, the adaptation statistics carried by the per-interval all-reduce,
the DREAM-ZS archive and snooker updates, and 1..3 pieces per generation."""
import os

import numpy as np
import pytest

import helpers as H
from hydrobayes_b200 import _abi as A

pytestmark = pytest.mark.gpu


def _single(cfg, target, x0, kw):
    return H.run_device(cfg, target, x0, **kw)


def _group(cfg, target, x0, kw, world, slice_gens=37):
    from hydrobayes_b200.sampler import GroupSampler

    def configure(s):
        if "par_min" in kw:
            s.set_bounds(kw["par_min"], kw["par_max"])
        if "obs" in kw:
            s.set_obs(kw["obs"])
        s.set_target(target, kw.get("forcing"))

    nl = cfg.nc // world
    with GroupSampler(cfg, world, 0, configure) as g:
        g.set_initial(x0)
        done = 1
        while done < cfg.t:
            done = g.run(slice_gens)
        out = []
        for r, s in enumerate(g.ranks):
            res = {"chain": s.fetch_chain(nl)}
            res["AR"], res["CR"] = s.fetch_ar_cr()
            res["accept"] = s.fetch_accept(nl)
            xt, lp, ll, lpost = s.fetch_state(nl)
            res.update(xt=xt, lpost=lpost, outlier=s.fetch_outliers())
            if cfg.check_convergence:
                res["R_stat"] = s.fetch_rstat()
            out.append(res)
    return out


def _compare(full, shards, cfg, world):
    nl = cfg.nc // world
    for r, sh in enumerate(shards):
        lo = r * nl
        assert np.array_equal(sh["accept"], full["accept"][:, lo:lo + nl]), f"rank {r}: accept/reject sequences differ"
        assert np.array_equal(sh["chain"], full["chain"][:, :, lo:lo + nl], equal_nan=True), f"rank {r}: chain differs"
        assert np.array_equal(sh["xt"], full["xt"]), f"rank {r}: the replica differs from the single-GPU population"
        assert np.array_equal(sh["lpost"], full["lpost"][lo:lo + nl])
        assert sh["outlier"] == full["outlier"]
        np.testing.assert_allclose(sh["AR"], full["AR"], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(sh["CR"], full["CR"], rtol=1e-9, equal_nan=True)
        if "R_stat" in full:
            np.testing.assert_allclose(sh["R_stat"], full["R_stat"], rtol=1e-9, equal_nan=True)


@pytest.fixture
def pieces_env():
    # bit-for-bit comparisons need the same kernel variant on both sides: the single-handle reference run is forced onto the
    # three-kernel path the sharded run uses (a small single population would otherwise take the fused generation kernel,
    # whose log-density sums the quadratic form in FMA order -- same decisions, last-bit differences in ll)
    old = {k: os.environ.get(k) for k in ("HB_PIECES", "HB_FUSED")}
    os.environ["HB_FUSED"] = "0"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("world,pieces", [(2, 1), (2, 3), (4, 2), (8, 1)])
def test_group_tdist_d100_outliers_adaptation_rstat(world, pieces, pieces_env):
    nc, d = 64 * world, 100
    x0 = H.x0_tdist(nc, d); x0[3] = 300.0; x0[nc - 2] = -250.0     # forces outlier resets across shards
    cfg = A.default_config(nc=nc, t=150, d=d, lik=2, seed=5, adapt=0.4, record_accept=1, check_convergence=1)
    full = _single(cfg, "t_distribution", x0, {})
    assert len(full["outlier"]) > 0 and np.isfinite(full["R_stat"]).any()
    os.environ["HB_PIECES"] = str(pieces)
    shards = _group(cfg, "t_distribution", x0, {}, world)
    _compare(full, shards, cfg, world)


@pytest.mark.parametrize("world,pieces", [(2, 2), (4, 1)])
def test_group_mixture_d1(world, pieces, pieces_env):
    nc = 512 * world
    cfg = A.default_config(nc=nc, t=200, d=1, lik=1, seed=9, adapt=0.3, record_accept=1)
    x0 = H.x0_mixture(nc)
    full = _single(cfg, "mixture", x0, {})
    os.environ["HB_PIECES"] = str(pieces)
    shards = _group(cfg, "mixture", x0, {}, world)
    _compare(full, shards, cfg, world)


@pytest.mark.parametrize("world,pieces", [(2, 2), (4, 1)])
def test_group_zs_archive_snooker(world, pieces, pieces_env):
    ncz, dz, m0 = 8 * world, 12, 200
    cfg = A.default_config(nc=ncz, t=160, d=dz, lik=2, seed=11, past_sample=1, psnooker=0.2, m0=m0, archive_update=5, record_accept=1)
    xz = np.random.default_rng(7).uniform(-5, 15, size=(m0, dz))
    full = _single(cfg, "t_distribution", xz, {})
    os.environ["HB_PIECES"] = str(pieces)
    shards = _group(cfg, "t_distribution", xz, {}, world)
    _compare(full, shards, cfg, world)


@pytest.mark.parametrize("world,pieces", [(2, 2), (4, 3)])
def test_group_wide_kernels(world, pieces, pieces_env):
    """the large-population kernels (warp per 32 chains, TMA rings, log slots allocated per batch) on a sharded population"""
    L = A.lib()
    nc, d = 96 * world, 40
    x0 = H.x0_tdist(nc, d)
    cfg = A.default_config(nc=nc, t=90, d=d, lik=2, seed=21, adapt=0.5, record_accept=1, thin=3)
    prev = L.hb_test_force_wide(1)
    try:
        full = _single(cfg, "t_distribution", x0, {})
        os.environ["HB_PIECES"] = str(pieces)
        shards = _group(cfg, "t_distribution", x0, {}, world)
    finally:
        L.hb_test_force_wide(prev)
    _compare(full, shards, cfg, world)


def test_group_hydromodel_bounds(pieces_env):
    world = 2
    nc = 64 * world
    P, obs = H.hydro_data(T=200)
    cfg = A.default_config(nc=nc, t=120, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_REFLECT,
                           seed=3, adapt=0.5, record_accept=1)
    kw = dict(par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
    full = _single(cfg, "HydroModel", H.x0_hydro(nc), kw)
    os.environ["HB_PIECES"] = "2"
    shards = _group(cfg, "HydroModel", H.x0_hydro(nc), kw, world)
    _compare(full, shards, cfg, world)


def test_group_error_is_reported_by_every_rank(pieces_env):
    """a shard-local failure (HydroModel's `param < 0` stop() for a chain of rank 1 only) must come back from EVERY rank's
    slice: the error masks are OR-ed over the ranks before hb_run returns (otherwise the other ranks would enter
    the next slice and wait for a peer that has already returned)"""
    from hydrobayes_b200.sampler import GroupSampler
    world = 2
    nc = 16 * world
    P, obs = H.hydro_data(T=50)
    x0 = H.x0_hydro(nc)
    x0[nc - 1] = -0.5                                   # a chain of rank 1: K < 0  (R/test_HydroModel.R:18)
    cfg = A.default_config(nc=nc, t=30, d=1, lik=31, glue_shape=50.0, seed=2)

    def configure(s):
        s.set_obs(obs); s.set_target("HydroModel", P)

    with GroupSampler(cfg, world, 0, configure) as g:
        g.set_initial(x0)
        with pytest.raises(A.HbError):
            g.run(10)
        msgs = [g.L.hb_last_error(s.h).decode() for s in g.ranks]
    # rank 0's chains are all valid: whatever it reports reached it through the OR-reduce of the error masks
    assert msgs[0] == msgs[1] and ("has to be >= zero" in msgs[0] or "not finite" in msgs[0]), msgs
