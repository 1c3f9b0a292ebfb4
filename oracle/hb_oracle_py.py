"""ctypes binding of the CPU oracle (oracle/hb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs,
never by the hydrobayes_b200 package.  PARITY UNPINNED (see oracle/hb_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hydrobayes_b200._abi import HbConfig, HbTape

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TARGET_KIND = {"mixture": 0, "t_distribution": 1, "HydroModel": 2}

_dp = C.POINTER(C.c_double)


class HboTarget(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("forcing", _dp), ("T", C.c_int64),
                ("obs", _dp), ("n_obs", C.c_int64), ("abc_e", _dp), ("n_abc_e", C.c_int64)]


class HboIo(C.Structure):
    _fields_ = [("chain", _dp), ("fx", _dp), ("AR", _dp), ("CR", _dp), ("accept", C.POINTER(C.c_uint8)),
                ("outlier_gen", C.POINTER(C.c_int64)), ("outlier_chain", C.POINTER(C.c_int64)),
                ("outlier_cap", C.c_int64), ("n_outlier_pairs", C.c_int64), ("rstat", _dp),
                ("xt", _dp), ("lp", _dp), ("ll", _dp), ("lpost", _dp), ("pCR", _dp), ("J", _dp), ("n_id", _dp),
                ("record", C.POINTER(HbTape)), ("run", C.c_int64), ("err", C.c_char * 256)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libhb_oracle.so")
    src = os.path.join(_HERE, "hb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CC": "gcc"})
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.hbo_log_xmin.restype = C.c_double
        L.hbo_dnorm.restype = C.c_double
        L.hbo_dnorm.argtypes = [C.c_double] * 3
        L.hbo_mixture.restype = C.c_double
        L.hbo_mixture.argtypes = [C.c_double]
        L.hbo_tdist_logpdf.argtypes = [_dp, C.c_int64, C.c_int32, _dp]
        L.hbo_hydromodel.argtypes = [_dp, C.c_int64, C.c_double, _dp]
        L.hbo_calc_ll.restype = C.c_double
        L.hbo_calc_ll.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_double, C.POINTER(C.c_int)]
        L.hbo_bound_par.restype = C.c_double
        L.hbo_bound_par.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int32]
        L.hbo_quantile7.restype = C.c_double
        L.hbo_quantile7.argtypes = [_dp, C.c_int64, C.c_double]
        L.hbo_rstat.argtypes = [_dp, C.c_int64, C.c_int32, C.c_int64, _dp]
        L.hbo_logdensity.argtypes = [C.POINTER(HboTarget), C.c_int32, C.c_double, _dp, C.c_int64, C.c_int32, _dp, _dp]
        L.hbo_out_dims.argtypes = [C.POINTER(HbConfig)] + [C.POINTER(C.c_int64)] * 5
        L.hbo_run.argtypes = [C.POINTER(HbConfig), C.POINTER(HboTarget), _dp, _dp, C.c_int, _dp, _dp, _dp,
                              C.POINTER(HbTape), C.POINTER(HboIo)]
        L.hbo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.hbo_set_threads.argtypes = [C.c_int]
        L.hbo_rng_u16.restype = C.c_double; L.hbo_rng_u16.argtypes = [C.c_uint32]
        L.hbo_rng_lambda16.restype = C.c_double; L.hbo_rng_lambda16.argtypes = [C.c_uint32, C.c_double]
        L.hbo_rng_normal32.restype = C.c_double; L.hbo_rng_normal32.argtypes = [C.c_uint32]
        L.hbo_rng_zt_threshold16.restype = C.c_uint32; L.hbo_rng_zt_threshold16.argtypes = [C.c_double]
        L.hbo_last_loop_seconds.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def log_xmin() -> float:
    return lib().hbo_log_xmin()


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().hbo_philox4x32_10(c, k, o)
    return list(o)


def mixture(x: float) -> float:
    return lib().hbo_mixture(float(x))


def tdist_logpdf(x) -> np.ndarray:
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, d = x.shape
    xf = np.asfortranarray(x)
    out = np.empty(n)
    rc = lib().hbo_tdist_logpdf(xf.ctypes.data_as(_dp), n, d, _p(out))
    assert rc == 0
    return out


def hydromodel(forcing, param: float) -> np.ndarray:
    f = _f64(forcing)
    out = np.empty(f.size)
    rc = lib().hbo_hydromodel(_p(f), f.size, float(param), _p(out))
    if rc:
        raise ValueError("Argument 'param' has to be >= zero! / forcing must be >= zero")
    return out


def bound_par(par, mn, mx, handle) -> float:
    return lib().hbo_bound_par(float(par), float(mn), float(mx), int(handle))


def quantile7(x, p) -> float:
    x = _f64(x)
    return lib().hbo_quantile7(_p(x), x.size, float(p))


def rstat(x) -> np.ndarray:
    """x: numpy array [t, d, n] (any memory order)."""
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    t, d, n = x.shape
    out = np.empty(d)
    lib().hbo_rstat(x.ctypes.data_as(_dp), t, d, n, _p(out))
    return out


def make_target(name, forcing=None, obs=None, abc_e=None):
    t = HboTarget()
    t.kind = TARGET_KIND[name]
    keep = []
    if forcing is not None:
        f = _f64(forcing); keep.append(f)
        t.forcing = _p(f); t.T = f.size
    if obs is not None:
        o = _f64(obs); keep.append(o)
        t.obs = _p(o); t.n_obs = o.size
    if abc_e is not None:
        e = _f64(np.atleast_1d(abc_e)); keep.append(e)
        t.abc_e = _p(e); t.n_abc_e = e.size
    return t, keep


def logdensity(name, lik, x, forcing=None, obs=None, glue_shape=0.0, abc_e=None):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, d = x.shape
    xf = np.asfortranarray(x)
    tgt, keep = make_target(name, forcing, obs, abc_e)
    ll = np.empty(n); fx = np.empty(n)
    rc = lib().hbo_logdensity(C.byref(tgt), lik, glue_shape, xf.ctypes.data_as(_dp), n, d, _p(ll), _p(fx))
    return ll, fx, rc


class Tape:
    """Owns the numpy arrays behind an hb_tape (layout of SURVEY Appendix D)."""

    def __init__(self, n_gen, nct, d, delta, n_events, snooker=False, mt=1):
        self.n_gen, self.nct, self.d, self.delta, self.n_events = n_gen, nct, d, delta, n_events
        self.n_sets = 2 * mt - 1 if mt > 1 else 1      # MT-DREAM: proposal sets per generation (hb_tape.n_sets)
        rows = n_gen * self.n_sets
        self.u_snooker = np.ones((rows, nct)) if snooker else None
        self.D = np.zeros((rows, nct), np.uint8)
        self.partners = np.full((rows, nct, 2 * delta), -1, np.int32)
        self.id = np.zeros((rows, nct), np.uint8)
        self.zt = np.zeros((rows, nct, d))
        self.lambda_ = np.zeros((rows, nct, d))
        self.g_is_one = np.zeros((rows, nct), np.uint8)
        self.zeta = np.zeros((rows, nct, d))
        self.g_snooker = np.full((rows, nct), 1.7) if snooker else None
        self.u_accept = np.zeros((n_gen, nct))
        self.outlier_new = np.full((max(n_events, 1), nct), -1, np.int32)
        self.mt_pick = np.zeros((n_gen, nct), np.uint8) if mt > 1 else None

    def struct(self) -> HbTape:
        t = HbTape()
        t.n_gen, t.nct, t.d, t.delta, t.n_events = self.n_gen, self.nct, self.d, self.delta, self.n_events
        u8 = C.POINTER(C.c_uint8); i32 = C.POINTER(C.c_int32)
        t.u_snooker = _p(self.u_snooker)
        t.D = self.D.ctypes.data_as(u8)
        t.partners = self.partners.ctypes.data_as(i32)
        t.id = self.id.ctypes.data_as(u8)
        t.zt = _p(self.zt)
        t.lambda_ = _p(self.lambda_)
        t.g_is_one = self.g_is_one.ctypes.data_as(u8)
        t.zeta = _p(self.zeta)
        t.g_snooker = _p(self.g_snooker)
        t.u_accept = _p(self.u_accept)
        t.outlier_new = self.outlier_new.ctypes.data_as(i32)
        t.n_sets = self.n_sets
        t.mt_pick = self.mt_pick.ctypes.data_as(u8) if self.mt_pick is not None else None
        return t

    @classmethod
    def for_config(cls, cfg: HbConfig):
        nct = max(cfg.n_runs, 1) * cfg.nc
        n_events = cfg.t // cfg.update_interval
        return cls(cfg.t - 1, nct, cfg.d, cfg.delta, n_events, snooker=cfg.psnooker > 0, mt=max(cfg.mt, 1))


def out_dims(cfg: HbConfig):
    v = [C.c_int64() for _ in range(5)]
    lib().hbo_out_dims(C.byref(cfg), *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def run(cfg: HbConfig, target: str, x0, par_min=None, par_max=None, prior_mu=None, prior_cov=None, tape: Tape = None,
        record: Tape = None, forcing=None, obs=None, run: int = 0, threads: int = 1, outputs: bool = True, abc_e=None):
    """Run the restated generation loop.  x0: [n0, d] array.  Returns a dict shaped like the reference's list."""
    L = lib()
    d, nc, t = cfg.d, cfg.nc, cfg.t
    out_t, ar_rows, ar_cols, cr_rows, cr_cols = out_dims(cfg)
    x0f = np.asfortranarray(np.asarray(x0, dtype=np.float64).reshape(-1, d))
    tgt, keep = make_target(target, forcing, obs, abc_e)
    io = HboIo()
    res = {}
    if outputs:
        res["chain"] = np.empty((out_t, d + 3, nc), order="F")
        res["fx"] = np.empty((out_t, nc), order="F")
        res["AR"] = np.empty((ar_rows, ar_cols), order="F")
        res["CR"] = np.empty((cr_rows, cr_cols), order="F")
        res["accept"] = np.zeros((t - 1, nc), np.uint8)
        cap = max(1, (t // cfg.update_interval) * nc)
        og = np.zeros(cap, np.int64); oc = np.zeros(cap, np.int64)
        res["xt"] = np.empty((nc, d), order="F")
        for k in ("lp", "ll", "lpost"):
            res[k] = np.empty(nc)
        for k in ("pCR", "J", "n_id"):
            res[k] = np.empty(cfg.nCR)
        io.chain = _p(res["chain"]); io.fx = _p(res["fx"]); io.AR = _p(res["AR"]); io.CR = _p(res["CR"])
        io.accept = res["accept"].ctypes.data_as(C.POINTER(C.c_uint8))
        io.outlier_gen = og.ctypes.data_as(C.POINTER(C.c_int64))
        io.outlier_chain = oc.ctypes.data_as(C.POINTER(C.c_int64))
        io.outlier_cap = cap
        io.xt = _p(res["xt"]); io.lp = _p(res["lp"]); io.ll = _p(res["ll"]); io.lpost = _p(res["lpost"])
        io.pCR = _p(res["pCR"]); io.J = _p(res["J"]); io.n_id = _p(res["n_id"])
        if cfg.check_convergence and out_t > 50:
            res["R_stat"] = np.empty((out_t - 50, d), order="F")
            io.rstat = _p(res["R_stat"])
    io.run = run
    rec_struct = None
    if record is not None:
        rec_struct = record.struct()
        io.record = C.pointer(rec_struct)
    tape_struct = tape.struct() if tape is not None else None
    pmin = _f64(par_min); pmax = _f64(par_max)
    nb = 0 if pmin is None else int(np.size(pmin))
    mu = _f64(prior_mu); cov = None if prior_cov is None else np.asfortranarray(np.asarray(prior_cov, np.float64))
    L.hbo_set_threads(int(threads))
    rc = L.hbo_run(C.byref(cfg), C.byref(tgt), _p(pmin), _p(pmax), nb, _p(mu), _p(cov), _p(x0f),
                   C.byref(tape_struct) if tape_struct is not None else None, C.byref(io))
    L.hbo_set_threads(1)
    res["rc"] = rc
    res["err"] = io.err.decode()
    res["loop_seconds"] = L.hbo_last_loop_seconds()
    if outputs:
        n = io.n_outlier_pairs
        res["outlier"] = list(zip(og[:n].tolist(), oc[:n].tolist()))
    return res
