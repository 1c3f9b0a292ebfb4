"""ctypes mirror of include/hydrobayes_b200.h and the loader of libhydrobayes_b200.so.

There is no CPU fallback: if the shared library (built by ``__graft_entry__.build()`` /
``hydrobayes_b200/csrc/Makefile``) is missing, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HB_ABI_VERSION = 1

HB_OK = 0
HB_ERR_INVALID = 1
HB_ERR_NO_DEVICE = 2
HB_ERR_CUDA = 3
HB_ERR_NONFINITE = 4
HB_ERR_UNKNOWN_TARGET = 5
HB_ERR_STATE = 6
HB_ERR_NCCL = 7
HB_ERR_UNSUPPORTED = 8

HB_SWEEP_SYNCHRONOUS = 0
HB_SWEEP_SEQUENTIAL = 1
HB_PRIOR_FLAT, HB_PRIOR_UNIFORM, HB_PRIOR_NORMAL = 0, 1, 2
HB_BOUND_NONE, HB_BOUND_BOUND, HB_BOUND_REFLECT, HB_BOUND_FOLD = 0, 1, 2, 3
HB_RNG_TAPE, HB_RNG_PHILOX = 0, 1
HB_EXCHANGE_ALLGATHER, HB_EXCHANGE_PEER, HB_EXCHANGE_DELTA = 0, 1, 2
HB_PEER_EXPORT_BYTES = 640   # 10 CUDA IPC handles of 64 bytes (include/hydrobayes_b200.h)

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class HbConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("sweep", C.c_int32),
        ("nc", C.c_int64), ("t", C.c_int64),
        ("d", C.c_int32), ("lik", C.c_int32),
        ("n_runs", C.c_int64),
        ("burnin", C.c_double), ("adapt", C.c_double),
        ("update_interval", C.c_int32), ("delta", C.c_int32),
        ("c_val", C.c_double), ("c_star", C.c_double),
        ("nCR", C.c_int32), ("thin", C.c_int32),
        ("p_g", C.c_double), ("beta0", C.c_double),
        ("outlier_check", C.c_int32), ("prior", C.c_int32),
        ("bound", C.c_int32), ("past_sample", C.c_int32),
        ("m0", C.c_int64),
        ("archive_update", C.c_int32), ("mt", C.c_int32),
        ("psnooker", C.c_double), ("glue_shape", C.c_double),
        ("rng_mode", C.c_int32), ("record_accept", C.c_int32),
        ("seed", C.c_uint64),
        ("check_convergence", C.c_int32), ("store_fx", C.c_int32),
        ("exchange", C.c_int32), ("run_offset", C.c_int32),
    ]


class HbTape(C.Structure):
    _fields_ = [
        ("n_gen", C.c_int64), ("nct", C.c_int64),
        ("d", C.c_int32), ("delta", C.c_int32),
        ("u_snooker", _dp), ("D", _u8p), ("partners", _i32p), ("id", _u8p),
        ("zt", _dp), ("lambda_", _dp), ("g_is_one", _u8p), ("zeta", _dp),
        ("g_snooker", _dp), ("u_accept", _dp), ("outlier_new", _i32p),
        ("n_events", C.c_int64),
        ("n_sets", C.c_int32), ("reserved", C.c_int32), ("mt_pick", _u8p),
    ]


def default_config(**kw) -> HbConfig:
    """Reference defaults of dream_parallel() (R/dream_parallel.R:155-163)."""
    c = HbConfig()
    c.abi_version = HB_ABI_VERSION
    c.sweep = HB_SWEEP_SYNCHRONOUS
    c.n_runs = 1
    c.burnin = 0.0
    c.adapt = 0.1
    c.update_interval = 10
    c.delta = 3
    c.c_val = 0.1
    c.c_star = 1e-12
    c.nCR = 3
    c.thin = 1
    c.p_g = 0.2
    c.beta0 = 1.0
    c.outlier_check = 1
    c.prior = HB_PRIOR_FLAT
    c.bound = HB_BOUND_NONE
    c.mt = 1
    c.rng_mode = HB_RNG_PHILOX
    c.seed = 1312
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(f"hb_config has no field {k!r}")
        setattr(c, k, v)
    return c


class HbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


_LIB = None
LIB_NAME = "libhydrobayes_b200.so"


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def lib() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise HbError(HB_ERR_NO_DEVICE,
                      f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    hp = C.c_void_p
    cp = C.c_char_p
    sig = {
        "hb_register_target": ([C.c_void_p], C.c_int),
        "hb_target_registered": ([cp], C.c_int),
        "hb_config_default": ([C.POINTER(HbConfig)], None),
        "hb_device_count": ([], C.c_int),
        "hb_create": ([C.POINTER(HbConfig), C.c_int, C.POINTER(hp)], C.c_int),
        "hb_destroy": ([hp], None),
        "hb_last_error": ([hp], cp),
        "hb_set_bounds": ([hp, _dp, _dp, C.c_int], C.c_int),
        "hb_set_prior_normal": ([hp, _dp, _dp], C.c_int),
        "hb_set_lik_args": ([hp, C.c_char_p, _dp, C.c_int64, C.c_char_p], C.c_int),
        "hb_set_obs": ([hp, _dp, C.c_int64], C.c_int),
        "hb_set_target": ([hp, cp, _dp, _i64p, C.c_int], C.c_int),
        "hb_set_target_batched": ([hp, cp, _dp, C.c_int64, _dp, C.c_int64], C.c_int),
        "hb_set_initial": ([hp, _dp], C.c_int),
        "hb_set_initial_rows": ([hp, _dp], C.c_int),
        "hb_set_tape": ([hp, C.POINTER(HbTape)], C.c_int),
        "hb_comm_unique_id": ([C.c_void_p], C.c_int),
        "hb_comm_init": ([hp, C.c_int, C.c_int, C.c_void_p], C.c_int),
        "hb_peer_export": ([hp, C.c_void_p], C.c_int),
        "hb_peer_import": ([hp, C.c_void_p], C.c_int),
        "hb_run": ([hp, C.c_int64, _i64p], C.c_int),
        "hb_group_attach": ([C.POINTER(hp), C.c_int], C.c_int),
        "hb_group_run": ([C.POINTER(hp), C.c_int, C.c_int64, _i64p], C.c_int),
        "hb_sync": ([hp], C.c_int),
        "hb_out_dims": ([hp] + [_i64p] * 7, C.c_int),
        "hb_fetch_chain": ([hp, _dp], C.c_int),
        "hb_fetch_chain_rows": ([hp, C.c_int64, C.c_int64, _dp], C.c_int),
        "hb_fetch_fx": ([hp, _dp], C.c_int),
        "hb_fetch_ar": ([hp, _dp], C.c_int),
        "hb_fetch_cr": ([hp, _dp], C.c_int),
        "hb_fetch_rstat": ([hp, _dp], C.c_int),
        "hb_fetch_state": ([hp, _dp, _dp, _dp, _dp], C.c_int),
        "hb_fetch_accept": ([hp, _u8p], C.c_int),
        "hb_fetch_outliers": ([hp, _i64p, _i64p, _i64p, C.c_int64], C.c_int),
        "hb_rstat": ([hp, _dp], C.c_int),
        "hb_rstat_runs": ([hp, _dp], C.c_int),
        "hb_wait_idle": ([], None),
        "hb_host_prefault": ([C.c_void_p, C.c_size_t, C.c_int], C.c_int),
        "hb_host_register": ([C.c_void_p, C.c_size_t], C.c_int),
        "hb_host_unregister": ([C.c_void_p], C.c_int),
        "hb_test_force_wide": ([C.c_int], C.c_int),
        "hb_ind_out": ([hp], C.c_int64),
        "hb_get_timing": ([hp, _dp, _i64p], C.c_int),
        "hb_profile": ([hp, C.c_int], C.c_int),
        "hb_get_kernel_ms": ([hp, cp, _dp, _i64p], C.c_int),
        "hb_rstat_array": ([_dp, C.c_int64, C.c_int32, C.c_int64, _dp, C.c_char_p, C.c_int], C.c_int),
        "hb_logdensity": ([cp, C.c_int32, _dp, _i64p, C.c_int, _dp, C.c_int64, C.c_double, _dp, C.c_int64,
                           C.c_int32, _dp, _dp, C.c_char_p, C.c_int], C.c_int),
        "hb_density_open": ([cp, C.c_int32, _dp, _i64p, C.c_int, _dp, C.c_int64, C.c_double, C.c_int32, C.c_int64,
                             C.POINTER(C.c_void_p), C.c_char_p, C.c_int], C.c_int),
        "hb_density_eval": ([C.c_void_p, _dp, C.c_int64, _dp, _dp, C.c_char_p, C.c_int], C.c_int),
        "hb_density_dim": ([C.c_void_p], C.c_int32),
        "hb_density_close": ([C.c_void_p], None),
        "hb_hydromodel_series": ([_dp, C.c_int64, _dp, C.c_int64, _dp, C.c_char_p, C.c_int], C.c_int),
        "hb_fp64_peaks": ([_dp, _dp], C.c_int),
        "hb_divcheck": ([_dp, _dp, C.c_int64, _i64p], C.c_int),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)  # AttributeError here == the library does not export a declared symbol
        f.argtypes = args
        f.restype = res
    _LIB = L
    return L


EXPORTED_SYMBOLS = [
    "hb_register_target", "hb_target_registered", "hb_config_default", "hb_device_count", "hb_create", "hb_destroy",
    "hb_last_error", "hb_set_bounds", "hb_set_prior_normal", "hb_set_lik_args", "hb_set_obs", "hb_set_target", "hb_set_target_batched",
    "hb_set_initial", "hb_set_initial_rows", "hb_set_tape", "hb_comm_unique_id", "hb_comm_init", "hb_peer_export", "hb_peer_import", "hb_run", "hb_group_attach", "hb_group_run", "hb_sync", "hb_out_dims",
    "hb_fetch_chain", "hb_fetch_chain_rows", "hb_fetch_fx", "hb_fetch_ar", "hb_fetch_cr", "hb_fetch_rstat",
    "hb_fetch_state", "hb_fetch_accept", "hb_fetch_outliers", "hb_rstat", "hb_rstat_runs", "hb_wait_idle", "hb_host_prefault", "hb_host_register", "hb_host_unregister", "hb_test_force_wide", "hb_ind_out", "hb_get_timing", "hb_profile",
    "hb_get_kernel_ms", "hb_rstat_array", "hb_logdensity", "hb_density_open", "hb_density_eval", "hb_density_dim", "hb_density_close", "hb_hydromodel_series", "hb_fp64_peaks", "hb_divcheck",
]
