"""Host-side mirror of HydroBayes' R API on top of the C ABI (include/hydrobayes_b200.h).

``dream_parallel`` / ``dream`` / ``R_stat`` keep the formals, defaults, error behaviour and return structures of
R/dream_parallel.R:155-163, R/dream.R:126-133 and R/gelman_rubin.R:21; new optional arguments come last and default
to the reference behaviour.  The thin R drivers that a HydroBayes maintainer would ship are in r-package/ (R is not
installed in the build image, so this Python mirror is what the parity tests drive).

Everything numerical happens in libhydrobayes_b200.so on the GPU; nothing here computes a proposal, a density or an
acceptance on the CPU, and there is no fallback when the library or a device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional

import numpy as np

from . import _abi as A

_dp = C.POINTER(C.c_double)

_PRIOR = {"flat": A.HB_PRIOR_FLAT, "uniform": A.HB_PRIOR_UNIFORM, "normal": A.HB_PRIOR_NORMAL}
_BOUND = {None: A.HB_BOUND_NONE, "bound": A.HB_BOUND_BOUND, "reflect": A.HB_BOUND_REFLECT, "fold": A.HB_BOUND_FOLD}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


_PINNED = {}     # page-locked result arrays (address -> array) waiting for their background cudaHostUnregister


def _unpin_later(arr):
    """The result array was page-locked for the fetch; unlocking 10 GB takes a few 100 ms, so it happens on a helper thread
    after the call has returned (the array stays valid and usable throughout; the thread holds a reference to it)."""
    import threading
    if _PINNED.pop(arr.ctypes.data, None) is None:
        return
    L = A.lib()

    def work(a=arr):
        L.hb_host_unregister(a.ctypes.data)
    threading.Thread(target=work, daemon=False).start()


class Sampler:
    """Owns an ``hb_handle``.  Thin: every method is one C-ABI call."""

    def __init__(self, cfg: A.HbConfig, device: int = 0):
        self.L = A.lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.hb_create(C.byref(cfg), device, C.byref(self.h))
        if rc != A.HB_OK:
            raise A.HbError(rc, self.L.hb_last_error(None).decode())
        self._keep = []

    def _check(self, rc):
        if rc != A.HB_OK:
            raise A.HbError(rc, self.L.hb_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.hb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- configuration --------------------------------------------------------------------------------
    def set_bounds(self, par_min, par_max):
        mn = np.ascontiguousarray(np.atleast_1d(np.asarray(par_min, np.float64)))
        mx = np.ascontiguousarray(np.atleast_1d(np.asarray(par_max, np.float64)))
        if mn.size != mx.size:
            raise ValueError("min and max must have the same length")
        self._check(self.L.hb_set_bounds(self.h, _ptr(mn), _ptr(mx), int(mn.size)))

    def set_prior_normal(self, mu, cov):
        """par.info$mu / $cov of prior = "normal" (dmvnorm, R/internals.R:58-59)."""
        d = self.cfg.d
        m = np.ascontiguousarray(np.asarray(mu, np.float64).reshape(d))
        c = np.asfortranarray(np.asarray(cov, np.float64).reshape(d, d))
        self._check(self.L.hb_set_prior_normal(self.h, _ptr(m), _ptr(c)))

    def set_lik_args(self, abc_rho=None, abc_e=None, lik_fun=None):
        """abc_rho= / abc_e= (lik 21, 22) and lik_fun= (lik 99): function NAMES resolved in the device registry
        ("abc_distfun", "fun_lik": the closures of examples/script_examples.R:494, 503-507).  Before set_target."""
        e = None if abc_e is None else np.ascontiguousarray(np.atleast_1d(np.asarray(abc_e, np.float64)))
        self._check(self.L.hb_set_lik_args(self.h, abc_rho.encode() if abc_rho else None, _ptr(e), 0 if e is None else e.size,
                                           lik_fun.encode() if lik_fun else None))

    def set_obs(self, obs):
        o = np.ascontiguousarray(np.asarray(obs, np.float64))
        self._check(self.L.hb_set_obs(self.h, _ptr(o), o.size))

    def set_target_batched(self, name: str, data, obs):
        """Per-run target data for batched independent runs: data [n_runs, len], obs [n_runs, n_obs]."""
        a = np.ascontiguousarray(np.asarray(data, np.float64)); o = np.ascontiguousarray(np.asarray(obs, np.float64))
        if a.ndim != 2 or o.ndim != 2 or a.shape[0] != self.cfg.n_runs or o.shape[0] != self.cfg.n_runs:
            raise ValueError("data and obs must be [n_runs, length] arrays")
        self._check(self.L.hb_set_target_batched(self.h, name.encode(), _ptr(a), a.shape[1], _ptr(o), o.shape[1]))

    def set_target(self, name: str, data=None):
        if data is None:
            self._check(self.L.hb_set_target(self.h, name.encode(), None, None, 0))
        else:
            a = np.ascontiguousarray(np.asarray(data, np.float64))
            dims = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
            self._check(self.L.hb_set_target(self.h, name.encode(), _ptr(a), dims, max(a.ndim, 1)))

    def set_initial(self, x0):
        d = self.cfg.d
        x = np.asarray(x0, np.float64).reshape(-1, d)
        if x.flags["F_CONTIGUOUS"] and not x.flags["C_CONTIGUOUS"]:     # R's layout: pass through
            self._check(self.L.hb_set_initial(self.h, x.ctypes.data_as(_dp)))
        else:                                                          # numpy's layout: no host transposition
            x = np.ascontiguousarray(x)
            self._check(self.L.hb_set_initial_rows(self.h, x.ctypes.data_as(_dp)))

    def set_tape(self, tape_struct: A.HbTape, keepalive=None):
        self._keep.append(keepalive)
        self._check(self.L.hb_set_tape(self.h, C.byref(tape_struct)))

    def comm_init(self, rank: int, world: int, unique_id: Optional[bytes] = None):
        """unique_id: the 128-byte NCCL id of hb_comm_unique_id (not needed for HB_EXCHANGE_DELTA, which runs its
        barriers and small collectives over peer-mapped memory)."""
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._check(self.L.hb_comm_init(self.h, rank, world, buf))

    def peer_setup(self, world: int, group=None):
        """HB_EXCHANGE_PEER: exchange the CUDA IPC handles of the double-buffered shards between the ranks."""
        import torch.distributed as dist
        mine = C.create_string_buffer(A.HB_PEER_EXPORT_BYTES)
        self._check(self.L.hb_peer_export(self.h, mine))
        parts = [None] * world
        dist.all_gather_object(parts, bytes(mine.raw), group=group)
        allh = C.create_string_buffer(b"".join(parts), A.HB_PEER_EXPORT_BYTES * world)
        self._check(self.L.hb_peer_import(self.h, allh))
        dist.barrier(group=group)

    # -- run ------------------------------------------------------------------------------------------
    def run(self, n_generations: int) -> int:
        done = C.c_int64()
        self._check(self.L.hb_run(self.h, int(n_generations), C.byref(done)))
        return done.value

    def profile(self, enable: bool):
        self._check(self.L.hb_profile(self.h, int(enable)))

    def kernel_ms(self, name: str):
        ms = C.c_double(); n = C.c_int64()
        self._check(self.L.hb_get_kernel_ms(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timing(self):
        ms = C.c_double(); n = C.c_int64()
        self._check(self.L.hb_get_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- results --------------------------------------------------------------------------------------
    def out_dims(self):
        v = [C.c_int64() for _ in range(7)]
        self._check(self.L.hb_out_dims(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("out_t", "ar_rows", "ar_cols", "cr_rows", "cr_cols", "fx_m", "rstat_rows"), (x.value for x in v)))

    def ind_out(self) -> int:
        return self.L.hb_ind_out(self.h)

    def n_local(self) -> int:
        return self.cfg.nc * max(self.cfg.n_runs, 1)   # sharded handles: pass n_chains = nc / world to the fetchers

    def alloc_chain(self, n_chains: Optional[int] = None, prefault: bool = True):
        """R allocates chain <- array(NA, c(out_t, d+3, nc)) BEFORE the generation loop (R/dream_parallel.R:201-203), so
        its pages are resident when the results are written.  Same here: the (uninitialised) result array is allocated up
        front and its pages are touched by a helper thread while the device runs; returns (array, thread-or-None)."""
        import threading
        dm = self.out_dims()
        n = n_chains or self.n_local()
        out = np.empty((dm["out_t"], self.cfg.d + 3, n), order="F")
        th = None
        if prefault and out.nbytes >= (64 << 20) and os.environ.get("HB_NO_PREFAULT") != "1":
            L = self.L
            # page-locking the result for a direct-DMA fetch is OFF by default: measured on the C4 result (9.5 GB),
            # cudaHostRegister pins at ~5 GB/s and serialises with the kernel launches of the running loop (the enqueue of
            # 999 generations took 1.3 s instead of 0.06 s), which costs more than the 35 ms the direct fetch saves
            pin = os.environ.get("HB_PIN") == "1"

            def touch_and_pin():                          # ctypes releases the GIL for the duration of the calls
                L.hb_host_prefault(out.ctypes.data, out.nbytes, 0)
                if pin and L.hb_host_register(out.ctypes.data, out.nbytes) == A.HB_OK:
                    _PINNED[out.ctypes.data] = out        # hb_fetch_chain then writes it by direct DMA; released by _unpin_later
            th = threading.Thread(target=touch_and_pin, daemon=True)
            th.start()
        return out, th

    def fetch_chain(self, n_chains: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        dm = self.out_dims()
        n = n_chains or self.n_local()
        if out is None:
            out = np.empty((dm["out_t"], self.cfg.d + 3, n), order="F")
        elif out.shape != (dm["out_t"], self.cfg.d + 3, n) or not out.flags["F_CONTIGUOUS"] or out.dtype != np.float64:
            raise ValueError("out must be a Fortran-ordered float64 array [out_t, d+3, n_chains]")
        self._check(self.L.hb_fetch_chain(self.h, _ptr(out)))
        return out

    def fetch_chain_rows(self, row0: int, nrows: int, n_chains: Optional[int] = None) -> np.ndarray:
        n = n_chains or self.n_local()
        out = np.empty((nrows, self.cfg.d + 3, n), order="F")
        self._check(self.L.hb_fetch_chain_rows(self.h, row0, nrows, _ptr(out)))
        return out

    def fetch_fx(self, n_chains: Optional[int] = None) -> np.ndarray:
        dm = self.out_dims()
        n = n_chains or self.n_local()
        out = np.empty((dm["out_t"], n, 1), order="F")
        self._check(self.L.hb_fetch_fx(self.h, _ptr(out)))
        return out

    def fetch_ar_cr(self):
        dm = self.out_dims()
        nr = max(self.cfg.n_runs, 1)
        ar = np.empty((nr, dm["ar_cols"], dm["ar_rows"]))      # run-major blocks of column-major [rows, cols]
        cr = np.empty((nr, dm["cr_cols"], dm["cr_rows"]))
        self._check(self.L.hb_fetch_ar(self.h, _ptr(ar)))
        self._check(self.L.hb_fetch_cr(self.h, _ptr(cr)))
        ar = np.transpose(ar, (0, 2, 1)); cr = np.transpose(cr, (0, 2, 1))
        if nr == 1:
            return np.asfortranarray(ar[0]), np.asfortranarray(cr[0])
        return ar, cr

    def fetch_rstat(self) -> np.ndarray:
        """out_rstat[out_t - 50, d] (R/dream.R:163, R/dream_parallel.R:207); batched runs: [n_runs, out_t - 50, d]."""
        dm = self.out_dims()
        nr = max(self.cfg.n_runs, 1)
        out = np.empty((nr, self.cfg.d, dm["rstat_rows"]))          # run-major blocks of column-major [rows, d]
        self._check(self.L.hb_fetch_rstat(self.h, _ptr(out)))
        out = np.transpose(out, (0, 2, 1))
        return np.asfortranarray(out[0]) if nr == 1 else out

    def fetch_state(self, n_chains: Optional[int] = None):
        d = self.cfg.d
        n = n_chains or self.n_local()
        peer_shard = n_chains is not None and self.cfg.exchange == A.HB_EXCHANGE_PEER   # peer mode: the shard only
        nc = n if peer_shard else self.n_local()
        xt = np.empty((nc, d), order="F")
        lp = np.empty(n); ll = np.empty(n); lpost = np.empty(n)
        self._check(self.L.hb_fetch_state(self.h, _ptr(xt), _ptr(lp), _ptr(ll), _ptr(lpost)))
        return xt, lp, ll, lpost

    def fetch_accept(self, n_chains: Optional[int] = None) -> np.ndarray:
        n = n_chains or self.n_local()
        out = np.zeros((self.cfg.t - 1, n), np.uint8)
        self._check(self.L.hb_fetch_accept(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def fetch_outliers(self):
        n = C.c_int64()
        self._check(self.L.hb_fetch_outliers(self.h, C.byref(n), None, None, 0))
        g = np.zeros(max(n.value, 1), np.int64); c = np.zeros(max(n.value, 1), np.int64)
        i64 = C.POINTER(C.c_int64)
        self._check(self.L.hb_fetch_outliers(self.h, C.byref(n), g.ctypes.data_as(i64), c.ctypes.data_as(i64), n.value))
        return list(zip(g[:n.value].tolist(), c[:n.value].tolist()))

    def rstat(self) -> np.ndarray:
        out = np.empty(self.cfg.d)
        self._check(self.L.hb_rstat(self.h, _ptr(out)))
        return out

    def rstat_runs(self) -> np.ndarray:
        """R_stat of every independent run of a batched handle: [n_runs, d] (one device launch for all runs)."""
        out = np.empty((max(self.cfg.n_runs, 1), self.cfg.d))
        self._check(self.L.hb_rstat_runs(self.h, _ptr(out)))
        return out

    def run_to_convergence(self, threshold: float = 1.2, check_every: int = 10, hold: int = 3, first_check: int = 50):
        """Runs until every independent run has max_k R_stat_k < threshold at `hold` consecutive checks (a plain first
        crossing gives false early convergence when all chains are equally over-dispersed, SURVEY 8d metric 2), checked
        every `check_every` stored generations from ind_out > first_check (R/dream.R:270).  Returns a dict with the
        wall seconds, the generation every run converged at (0 = not converged within t) and the last R_stat.
        A single population (n_runs == 1, either sweep) is monitored through ``rstat()`` -- on a sharded population that is
        a collective: every rank calls this method and takes the same decisions."""
        t0 = time.perf_counter()
        nr = max(self.cfg.n_runs, 1)
        population = nr == 1 and self.cfg.sweep == A.HB_SWEEP_SYNCHRONOUS
        streak = np.zeros(nr, np.int64); conv_gen = np.zeros(nr, np.int64)
        done = self.run(max(first_check * self.cfg.thin, 1))
        r = None; checks = 0; check_seconds = 0.0
        while True:
            if self.ind_out() > first_check:
                tc = time.perf_counter()
                r = self.rstat()[None, :].max(axis=1) if population else self.rstat_runs().max(axis=1)
                check_seconds += time.perf_counter() - tc; checks += 1
                ok = r < threshold                                   # NaN (degenerate run) never counts as converged
                streak = np.where(ok, streak + 1, 0)
                newly = (streak >= hold) & (conv_gen == 0)
                conv_gen[newly] = done
                if (conv_gen > 0).all():
                    break
            if done >= self.cfg.t:
                break
            done = self.run(check_every * self.cfg.thin)
        return {"seconds": time.perf_counter() - t0, "generations": int(done), "converged_at": conv_gen,
                "all_converged": bool((conv_gen > 0).all()), "R_stat_max": r, "checks": checks,
                "check_seconds": check_seconds}


class GroupSampler:
    """Rank emulation on ONE device (testing hook: ``hb_group_attach`` / ``hb_group_run``): ``world`` handles, rank r owning
    chains [r*nc/world, (r+1)*nc/world), HB_EXCHANGE_DELTA, driven in lockstep on one stream.  The kernels and the arena
    protocol are those of ``world`` processes on ``world`` GPUs; a one-GPU box uses it to check that the shards of a
    sharded run equal the single-GPU run bit for bit."""

    def __init__(self, cfg: A.HbConfig, world: int, device: int = 0, configure=None):
        self.world = world
        self.ranks = []
        for r in range(world):
            c = A.HbConfig.from_buffer_copy(bytes(cfg))
            c.exchange = A.HB_EXCHANGE_DELTA
            s = Sampler(c, device)
            if configure is not None:
                configure(s)
            s.comm_init(r, world, None)
            self.ranks.append(s)
        self.L = A.lib()
        self._arr = (C.c_void_p * world)(*[s.h for s in self.ranks])

    def set_initial(self, x0):
        for s in self.ranks:
            s.set_initial(x0)
        rc = self.L.hb_group_attach(self._arr, self.world)
        if rc != A.HB_OK:
            raise A.HbError(rc, "; ".join(self.L.hb_last_error(s.h).decode() for s in self.ranks))

    def run(self, n_generations: int) -> int:
        done = C.c_int64()
        rc = self.L.hb_group_run(self._arr, self.world, int(n_generations), C.byref(done))
        if rc != A.HB_OK:
            raise A.HbError(rc, "; ".join(self.L.hb_last_error(s.h).decode() for s in self.ranks))
        return done.value

    def close(self):
        for s in self.ranks:
            s.close()
        self.ranks = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---------------------------------------------------------------------------------------------------------
# prior_sample (R/internals.R:26-47) stays on the host, exactly as it would in the R driver
# ---------------------------------------------------------------------------------------------------------
def prior_sample(par_info: dict, d: int, nc: int, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    rng = rng or np.random.default_rng()
    initial = par_info.get("initial")
    if initial == "uniform":
        mn = np.broadcast_to(np.asarray(par_info["min"], float), (d,))
        mx = np.broadcast_to(np.asarray(par_info["max"], float), (d,))
        xt = rng.uniform(mn, mx, size=(nc, d))                         # :29
    elif initial == "normal":
        xt = rng.multivariate_normal(np.asarray(par_info["mu"], float), np.asarray(par_info["cov"], float), size=nc)  # :32
    elif initial == "latin":
        mn = np.broadcast_to(np.asarray(par_info["min"], float), (d,))
        mx = np.broadcast_to(np.asarray(par_info["max"], float), (d,))
        # lhs::randomLHS: one stratum per point and dimension, uniformly inside the stratum (:34-35)
        u = (np.stack([rng.permutation(nc) for _ in range(d)], axis=1) + rng.uniform(size=(nc, d))) / nc
        xt = mn + (mx - mn) * u
    elif initial == "user":
        xt = np.asarray(par_info["val_ini"], float)                    # :38
    else:
        raise ValueError("Value 'initial' of argument list 'par.info' must be one of "
                         "{'uniform', 'normal', 'latin', 'user'}!")     # :40
    return np.asarray(xt, float).reshape(-1, d)                        # :42-43


# targets whose raw output is a whole series (m = T): fx[out_t, nc, m] is not kept on the device for them
_SERIES_TARGETS = ("HydroModel", "fun_wrap_glue")


def _make_config(sweep, lik, par_info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g, beta0,
                 thin, outlier_check, glue_shape, past_sample, m0, archive_update, psnooker, mt, checkConvergence,
                 seed, tape, record_accept, store_fx, default_prior) -> A.HbConfig:
    if lik is None:
        raise ValueError("Argument 'lik' has to be one of {1,2,21,22}!")  # calc_ll with lik = NULL errors in R
    prior = par_info.get("prior", default_prior)
    if prior not in _PRIOR:
        raise A.HbError(A.HB_ERR_UNSUPPORTED, f"prior = {prior!r}: user prior functions are R closures and stay off "
                                              "the device path")
    bound = par_info.get("bound")
    if bound not in _BOUND:
        raise ValueError("Value 'bound' of argument list 'par.info' must be one of {'bound', 'reflect', 'fold'}!")
    return A.default_config(
        sweep=sweep, nc=int(nc), t=int(t), d=int(d), lik=int(lik), burnin=float(burnin), adapt=float(adapt),
        update_interval=int(updateInterval), delta=int(delta), c_val=float(c_val), c_star=float(c_star), nCR=int(nCR),
        thin=int(thin), p_g=float(p_g), beta0=float(beta0), outlier_check=int(bool(outlier_check)),
        prior=_PRIOR[prior], bound=_BOUND[bound], past_sample=int(bool(past_sample)),
        m0=int(m0) if m0 is not None else 0, archive_update=int(archive_update) if archive_update is not None else 0,
        mt=int(mt), psnooker=float(psnooker), glue_shape=float(glue_shape) if glue_shape is not None else 0.0,
        rng_mode=A.HB_RNG_TAPE if tape is not None else A.HB_RNG_PHILOX,
        seed=int(seed) if seed is not None else 1312, record_accept=int(bool(record_accept)),
        check_convergence=int(bool(checkConvergence)), store_fx=int(bool(store_fx)))


def _drive(sweep, fun, extra, lik, par_info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g,
           beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun, past_sample, m0, archive_update,
           psnooker, mt, checkConvergence, verbose, seed, tape, device, record_accept, store_fx, slice_generations,
           default_prior, forcing, distributed=False, exchange=None, series_fx=False):
    if not isinstance(fun, str):
        raise A.HbError(A.HB_ERR_UNKNOWN_TARGET, "fun must be the NAME of a registered device target "
                                                 "(arbitrary closures stay off the GPU path)")
    par_info = dict(par_info or {})
    timea = time.time()
    cfg = _make_config(sweep, lik, par_info, nc, t, d, burnin, adapt, updateInterval, delta, c_val, c_star, nCR, p_g,
                       beta0, thin, outlier_check, glue_shape, past_sample, m0, archive_update, psnooker, mt,
                       checkConvergence, seed, tape, record_accept, store_fx and fun not in _SERIES_TARGETS, default_prior)
    rng = np.random.default_rng(seed)
    n0 = (m0 if m0 is not None else 10 * d) if past_sample else nc
    rank, world = 0, 1
    if distributed:   # one process per GPU: this rank owns chains [rank*nc/world, (rank+1)*nc/world)
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if world > 1:
            if sweep != A.HB_SWEEP_SYNCHRONOUS:
                raise A.HbError(A.HB_ERR_UNSUPPORTED, "only dream_parallel (synchronous sweep) shards across GPUs")
            # delta exchange also with the DREAM-ZS archive (replicated per rank); MT-DREAM all-gathers the state and the
            # candidate population zt (R/dream_parallel.R:316 draws the reference proposals' partners from all of zt)
            cfg.exchange = (A.HB_EXCHANGE_ALLGATHER if mt > 1 else A.HB_EXCHANGE_DELTA) if exchange is None else exchange
    n_shard = shard_range(nc, rank, world)[1] if world > 1 else None
    _t_in = time.perf_counter()
    with Sampler(cfg, device) as s:
        _timing_note("py: hb_create", _t_in)
        if par_info.get("min") is not None and par_info.get("max") is not None:
            s.set_bounds(par_info["min"], par_info["max"])
        if cfg.prior == A.HB_PRIOR_NORMAL:
            if par_info.get("mu") is None or par_info.get("cov") is None:
                raise ValueError("prior = 'normal' needs par.info$mu and par.info$cov")
            s.set_prior_normal(par_info["mu"], par_info["cov"])
        if obs is not None:
            s.set_obs(obs)
        if abc_rho is not None or abc_e is not None or lik_fun is not None:
            for nm, f in (("abc_rho", abc_rho), ("lik_fun", lik_fun)):
                if f is not None and not isinstance(f, str):
                    raise A.HbError(A.HB_ERR_UNKNOWN_TARGET, f"{nm} must be the NAME of a registered device function "
                                                             "(arbitrary closures stay off the GPU path)")
            s.set_lik_args(abc_rho, abc_e, lik_fun)
        data = forcing if forcing is not None else (extra[0] if extra else None)
        s.set_target(fun, data)
        if world > 1:
            uid = None if cfg.exchange == A.HB_EXCHANGE_DELTA else broadcast_unique_id(lambda: _unique_id(s.L), rank)
            s.comm_init(rank, world, uid)
        _timing_note("py: ... target / communicator set up", _t_in)
        x0 = prior_sample(par_info, d, n0, rng)   # same seed on every rank: every rank holds the whole initial population
        if x0.shape[0] != n0:
            raise ValueError(f"initial sample has {x0.shape[0]} rows, expected {n0}")
        s.set_initial(x0)
        if world > 1 and cfg.exchange in (A.HB_EXCHANGE_PEER, A.HB_EXCHANGE_DELTA):
            s.peer_setup(world)
        if tape is not None:
            s.set_tape(tape.struct(), keepalive=tape)
        step = int(slice_generations) if slice_generations else (max(1, t // 50) if verbose else t)
        chain_buf, toucher = s.alloc_chain(n_shard)                 # chain <- array(NA, ...) before the loop, :201-203
        _timing_note("py: ... initial population uploaded, result allocated", _t_in)
        done = 1
        _tn = _timing_note
        _t0 = time.perf_counter()
        while done < t:
            done = s.run(step)          # the R driver ticks txtProgressBar / checks interrupts between slices
        _tn("py: generation loop", _t0); _t0 = time.perf_counter()
        if toucher is not None:
            toucher.join()
        _tn("py: wait for the pre-fault thread", _t0); _t0 = time.perf_counter()
        chain = s.fetch_chain(n_shard, out=chain_buf)   # sharded: this rank's chains, chain[out_t, d+3, nc/world]
        _tn("py: fetch_chain", _t0)
        _unpin_later(chain_buf)
        ar, cr = s.fetch_ar_cr()
        out = {"chain": chain, "rank": rank, "world": world}
        names = par_info.get("names") or [f"par{k + 1}" for k in range(d)]
        out["chain_colnames"] = list(names) + ["lp", "ll", "lpost"]
        # fx[out_t, nc, m] (R/dream_parallel.R:230, 394): scalar targets as in the reference; series targets (HydroModel,
        # m == T) return None unless series_fx=True, which rebuilds fx[it, c, :] = HydroModel(forcing, chain[it, 0, c]) on the
        # device from the stored chain (the reference's values except for the stale rows of SURVEY F9, see r-package/R/hb_drive.R)
        out["fx"] = s.fetch_fx(n_shard) if cfg.store_fx else None
        if out["fx"] is None and series_fx and fun in _SERIES_TARGETS and d == 1 and data is not None:
            ser = HydroModel(data, np.ascontiguousarray(chain[:, 0, :]).ravel())          # [out_t * n, T]
            out["fx"] = ser.reshape(chain.shape[0], chain.shape[2], -1)
        pairs = s.fetch_outliers()
        out["outlier"] = _outlier_list(pairs, cfg, sweep)
        out["AR"] = ar
        out["CR"] = cr
        if sweep == A.HB_SWEEP_SYNCHRONOUS:
            out["AR_colnames"] = ["n_eval", "accepted"]
        if checkConvergence:
            out["R_stat"] = s.fetch_rstat()
        if record_accept:
            out["accept"] = s.fetch_accept(n_shard)
        loop_ms, launches = s.timing()
        out["device_loop_seconds"] = loop_ms * 1e-3
        out["gpu_launches"] = launches
        _timing_note("py: ... results fetched", _t_in)
    _timing_note("py: ... handle closed (total)", _t_in)
    out["runtime"] = time.time() - timea
    return out


def _timing_note(what, t0):
    """HB_TIMING=1: host-side phases of the drivers on stderr (next to the library's own `[hb timing]` lines)"""
    if os.environ.get("HB_TIMING") == "1":
        import sys
        print(f"[hb timing] {what:38s} {(time.perf_counter() - t0) * 1e3:8.1f} ms", file=sys.stderr, flush=True)


def _outlier_list(pairs, cfg, sweep):
    """R: outl[[i]] <- check_out$outliers (integer(0) when none); entries of other generations stay NULL."""
    last = 0
    i = cfg.update_interval
    checks = []
    while i <= cfg.adapt * cfg.t and i <= cfg.t:
        checks.append(i)
        i += cfg.update_interval
    if not cfg.outlier_check or cfg.past_sample or not checks:
        return [None]
    last = checks[-1]
    lst = [None] * last
    for g in checks:
        lst[g - 1] = np.array([c for (gg, c) in pairs if gg == g], dtype=np.int64)
    return lst


def dream_parallel(fun, *extra, lik=None, par_info=None, nc, t, d, burnin=0, adapt=0.1, updateInterval=10, delta=3,
                   c_val=0.1, c_star=1e-12, nCR=3, p_g=0.2, beta0=1, thin=1, outlier_check=True, obs=None, abc_rho=None,
                   abc_e=None, glue_shape=None, lik_fun=None, past_sample=False, m0=None, archive_update=None,
                   psnooker=0, mt=1, ncores=1, checkConvergence=False, verbose=True,
                   seed=None, tape=None, device=0, record_accept=False, store_fx=True, slice_generations=None,
                   forcing=None, distributed=False, exchange=None, series_fx=False):
    """DREAM, synchronous population sweep -- drop-in for HydroBayes::dream_parallel (R/dream_parallel.R:155-455).

    ``fun`` is the name of a registered device target ("mixture", "t_distribution", "HydroModel", or "fun_wrap_glue", the
    HydroModel wrapper of the reference's example script).  ``ncores`` is
    accepted and ignored.  Chain indices in ``outlier`` are 0-based.

    ``distributed=True`` (one process per GPU, ``torch.distributed`` initialised): the population is block-partitioned
    across the ranks, every rank returns the ``chain`` / ``fx`` of its own chains (``gather_chain`` assembles the full
    array); AR, CR, outlier and R_stat are identical on every rank.
    """
    return _drive(A.HB_SWEEP_SYNCHRONOUS, fun, extra, lik, par_info, nc, t, d, burnin, adapt, updateInterval, delta,
                  c_val, c_star, nCR, p_g, beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun,
                  past_sample, m0, archive_update, psnooker, mt, checkConvergence, verbose, seed, tape, device,
                  record_accept, store_fx, slice_generations, "flat", forcing, distributed, exchange, series_fx)


def dream(fun, *extra, lik=None, par_info=None, nc, t, d, burnin=0, adapt=0.1, updateInterval=10, delta=3, c_val=0.1,
          c_star=1e-12, nCR=3, p_g=0.2, beta0=1, thin=1, outlier_check=True, obs=None, abc_rho=None, abc_e=None,
          glue_shape=None, lik_fun=None, checkConvergence=False, verbose=True,
          seed=None, tape=None, device=0, record_accept=False, store_fx=True, slice_generations=None, forcing=None,
          series_fx=False):
    """DREAM, sequential chain sweep -- drop-in for HydroBayes::dream (R/dream.R:126-316)."""
    return _drive(A.HB_SWEEP_SEQUENTIAL, fun, extra, lik, par_info, nc, t, d, burnin, adapt, updateInterval, delta,
                  c_val, c_star, nCR, p_g, beta0, thin, outlier_check, obs, abc_rho, abc_e, glue_shape, lik_fun,
                  False, None, None, 0, 1, checkConvergence, verbose, seed, tape, device, record_accept, store_fx,
                  slice_generations, "uniform", forcing, series_fx=series_fx)


def R_stat(x) -> np.ndarray:
    """Gelman-Rubin diagnostic of x[samples, parameters, chains] -- drop-in for R_stat (R/gelman_rubin.R:21)."""
    x = np.asfortranarray(np.asarray(x, np.float64))
    if x.ndim != 3:
        raise ValueError("x must be a 3-d array [samples, parameters, chains]")
    t, d, n = x.shape
    out = np.empty(d)
    err = C.create_string_buffer(256)
    rc = A.lib().hb_rstat_array(x.ctypes.data_as(_dp), t, d, n, _ptr(out), err, 256)
    if rc != A.HB_OK:
        raise A.HbError(rc, err.value.decode())
    return out


def log_density(fun: str, x, lik: int, forcing=None, obs=None, glue_shape: float = 0.0):
    """calc_ll(get(fun)(x_i, ...), lik, obs, glue_shape) for the rows of x on the device (R/internals.R:3-23)."""
    x = np.atleast_2d(np.asarray(x, np.float64))
    n, d = x.shape
    xf = np.asfortranarray(x)
    ll = np.empty(n); fx = np.empty(n)
    err = C.create_string_buffer(256)
    data = None if forcing is None else np.ascontiguousarray(np.asarray(forcing, np.float64))
    o = None if obs is None else np.ascontiguousarray(np.asarray(obs, np.float64))
    dims = (C.c_int64 * 1)(data.size if data is not None else 0)
    rc = A.lib().hb_logdensity(fun.encode(), int(lik), _ptr(data), dims, 1 if data is not None else 0, _ptr(o),
                               o.size if o is not None else 0, float(glue_shape), xf.ctypes.data_as(_dp), n, d,
                               _ptr(ll), _ptr(fx), err, 256)
    if rc != A.HB_OK:
        raise A.HbError(rc, err.value.decode())
    return ll, fx


def mixture(x) -> np.ndarray:
    """mixture(x) = 1/6 dnorm(x, -8, 1) + 5/6 dnorm(x, 10, 1), the exported test target of R/test_mixture.R:16, evaluated on the
    device for every element of x (a DENSITY, as in the reference; `lik = 1` takes its log inside the samplers)."""
    a = np.asarray(x, np.float64)
    _, fx = log_density("mixture", a.reshape(-1, 1), 1)
    return fx.reshape(a.shape) if a.ndim else float(fx[0])


def t_distribution(x):
    """t_distribution(x), the exported d-variate Student target of R/test_t_distribution.R:19-46 (60 degrees of freedom,
    C[i, j] = 0.5 sqrt(i j), C[i, i] = i): the LOG-density of one point x[d] or of every row of x[n, d], on the device."""
    a = np.asarray(x, np.float64)
    _, fx = log_density("t_distribution", np.atleast_2d(a), 2)
    return fx if a.ndim > 1 else float(fx[0])


class DensitySession:
    """log_density() with the plugin context and the device buffers kept between calls (``hb_density_open`` / ``_eval`` /
    ``_close``): for callers that evaluate the target once per iteration -- ``pdf(xp)`` of rwm() / am() / am_sd()
    (R/rwm.R:42) -- where the per-call setup of log_density() (covariance factor, forcing upload, allocations) would dominate."""

    def __init__(self, fun: str, lik: int, d: int, max_n: int, forcing=None, obs=None, glue_shape: float = 0.0):
        if not isinstance(fun, str):
            raise A.HbError(A.HB_ERR_UNKNOWN_TARGET, "fun must be the NAME of a registered device target")
        data = None if forcing is None else np.ascontiguousarray(np.asarray(forcing, np.float64))
        o = None if obs is None else np.ascontiguousarray(np.asarray(obs, np.float64))
        dims = (C.c_int64 * 1)(data.size if data is not None else 0)
        self._s = C.c_void_p()
        self._L = A.lib()
        self.d, self.max_n = int(d), int(max_n)
        err = C.create_string_buffer(256)
        rc = self._L.hb_density_open(fun.encode(), int(lik), _ptr(data), dims, 1 if data is not None else 0, _ptr(o),
                                     o.size if o is not None else 0, float(glue_shape), self.d, self.max_n, C.byref(self._s),
                                     err, 256)
        if rc != A.HB_OK:
            self._s = C.c_void_p()
            raise A.HbError(rc, err.value.decode())

    def __call__(self, x):
        """(ll, fx) for the rows of x[n, d], n <= max_n"""
        x = np.atleast_2d(np.asarray(x, np.float64))
        n, d = x.shape
        if d != self.d:
            raise ValueError(f"x has {d} columns, the session was opened for d = {self.d}")
        xf = np.asfortranarray(x)
        ll = np.empty(n); fx = np.empty(n)
        err = C.create_string_buffer(256)
        rc = self._L.hb_density_eval(self._s, xf.ctypes.data_as(_dp), n, _ptr(ll), _ptr(fx), err, 256)
        if rc != A.HB_OK:
            raise A.HbError(rc, err.value.decode())
        return ll, fx

    def close(self):
        if getattr(self, "_s", None):
            self._L.hb_density_close(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def HydroModel(forcing, param) -> np.ndarray:
    """HydroModel(forcing, param) (R/test_HydroModel.R:17-36) evaluated on the device; param may be a vector."""
    f = np.ascontiguousarray(np.asarray(forcing, np.float64))
    p = np.ascontiguousarray(np.atleast_1d(np.asarray(param, np.float64)))
    out = np.empty((p.size, f.size))
    err = C.create_string_buffer(256)
    rc = A.lib().hb_hydromodel_series(_ptr(f), f.size, _ptr(p), p.size, _ptr(out), err, 256)
    if rc != A.HB_OK:
        raise A.HbError(rc, err.value.decode())
    return out[0] if np.ndim(param) == 0 else out


# ---------------------------------------------------------------------------------------------------------
# multi-GPU host plumbing (one process per GPU; torch.distributed is plumbing only)
# ---------------------------------------------------------------------------------------------------------
def _unique_id(L) -> bytes:
    buf = C.create_string_buffer(128)
    rc = L.hb_comm_unique_id(buf)
    if rc != A.HB_OK:
        raise A.HbError(rc, L.hb_last_error(None).decode())
    return bytes(buf.raw)


def shard_range(nc: int, rank: int, world: int):
    """Block partition of the population: rank r owns chains [r*nc/world, (r+1)*nc/world) (hb_comm_init)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if nc % world:
        raise ValueError("nc must be divisible by the number of GPUs")
    nl = nc // world
    return rank * nl, nl


def broadcast_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the 128-byte communicator id (hb_comm_unique_id) and every rank receives it."""
    import torch.distributed as dist
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    if not isinstance(obj[0], (bytes, bytearray)) or len(obj[0]) != 128:
        raise ValueError("communicator id must be 128 bytes")
    return bytes(obj[0])


def gather_chain(shard: np.ndarray, world: int, group=None) -> np.ndarray:
    """All ranks contribute chain[out_t, d+3, nc/world]; returns chain[out_t, d+3, nc] (chains are the slowest
    dimension of R's layout, so rank shards concatenate along the last axis)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(shard, 2, 0)))          # [nl, out_t, d+3]
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    full = torch.cat(parts, dim=0).cpu().numpy()                                    # [nc, out_t, d+3]
    return np.asfortranarray(np.moveaxis(full, 0, 2))
