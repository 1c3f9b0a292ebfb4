"""The plugin boundary: `get(fun)(x, ...)` by name (R/dream_parallel.R:279-283) is a name-keyed registry of device targets
(`hb_target_vtable`, `hb_register_target`, include/hydrobayes_b200.h).  CPU: the registry's own rules, and that a third-party
target builds against the public header alone.  GPU: that target (tests/plugin/gauss_plugin.cu, compiled outside the library)
is registered at run time and driven through the same entry points as the built-in targets -- one-shot log-density,
log-density session, dream_parallel() and dream()."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hydrobayes_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN_DIR = os.path.join(ROOT, "tests", "plugin")
PLUGIN_SO = os.path.join(PLUGIN_DIR, "libhb_gauss_plugin.so")
FLOOR = np.log(np.finfo(float).tiny)                            # log(.Machine$double.xmin), R/internals.R:19

_INIT = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int,
                    C.POINTER(C.c_double), C.c_int64, C.c_double, C.c_char_p, C.c_int)
_LAUNCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_void_p)
_DESTROY = C.CFUNCTYPE(None, C.c_void_p)


class _VTable(C.Structure):                                     # hb_target_vtable, include/hydrobayes_b200.h
    _fields_ = [("name", C.c_char_p), ("init", _INIT), ("launch", _LAUNCH), ("destroy", _DESTROY), ("work", C.c_void_p),
                ("launch_items", C.c_void_p)]


_KEEP = []                                                      # the registry keeps the function pointers: keep the callbacks alive


def _stub_vtable(name, with_launch=True):
    init = _INIT(lambda *a: A.HB_ERR_INVALID)
    launch = _LAUNCH(lambda *a: A.HB_ERR_INVALID) if with_launch else _LAUNCH()
    destroy = _DESTROY(lambda ctx: None)
    vt = _VTable(name, init, launch, destroy, None, None)
    _KEEP.extend([init, launch, destroy, vt])
    return vt


def test_registry_rules():
    L = A.lib()
    reg = L.hb_register_target
    reg.argtypes = [C.c_void_p]; reg.restype = C.c_int
    assert L.hb_target_registered(b"cpu_stub_target") == 0
    assert reg(C.addressof(_stub_vtable(b"cpu_stub_target"))) == A.HB_OK
    assert L.hb_target_registered(b"cpu_stub_target") == 1
    assert reg(C.addressof(_stub_vtable(b"cpu_stub_target"))) == A.HB_ERR_INVALID       # names are unique
    assert reg(C.addressof(_stub_vtable(b"t_distribution"))) == A.HB_ERR_INVALID        # built-ins cannot be shadowed
    assert reg(C.addressof(_stub_vtable(b"no_launch", with_launch=False))) == A.HB_ERR_INVALID
    assert reg(None) == A.HB_ERR_INVALID
    assert L.hb_target_registered(b"no_launch") == 0


def _build_plugin():
    src = os.path.join(PLUGIN_DIR, "gauss_plugin.cu")
    deps = [src, os.path.join(ROOT, "include", "hydrobayes_b200.h"), A.lib_path()]
    if not os.path.exists(PLUGIN_SO) or any(os.path.getmtime(PLUGIN_SO) < os.path.getmtime(p) for p in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                        "-shared", "-I" + os.path.join(ROOT, "include"), src, "-o", PLUGIN_SO,
                        "-L" + os.path.dirname(A.lib_path()), "-lhydrobayes_b200",
                        "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../hydrobayes_b200"], check=True)
    return PLUGIN_SO


def test_plugin_builds_against_the_public_header_only():
    # nvcc cross-compiles sm_100a without a GPU: the third-party target needs nothing but include/hydrobayes_b200.h and
    # takes exactly one symbol from the library
    so = _build_plugin()
    out = subprocess.run(["nm", "-D", "--defined-only", so], check=True, capture_output=True, text=True).stdout
    assert "hb_gauss_plugin_register" in out
    und = subprocess.run(["nm", "-D", "--undefined-only", so], check=True, capture_output=True, text=True).stdout
    assert [ln.split()[-1] for ln in und.splitlines() if " hb_" in ln] == ["hb_register_target"]


@pytest.fixture(scope="module")
def gauss_plugin():
    A.lib()                                                     # the library first: the plugin resolves hb_register_target in it
    try:
        so = _build_plugin()
    except (OSError, subprocess.CalledProcessError) as e:       # no nvcc on this machine: use the library build()'s copy
        if not os.path.exists(PLUGIN_SO):
            pytest.skip(f"third-party target not built and cannot be built here: {e}")
        so = PLUGIN_SO
    P = C.CDLL(so)
    P.hb_gauss_plugin_register.restype = C.c_int
    rc = P.hb_gauss_plugin_register()
    assert rc == A.HB_OK or A.lib().hb_target_registered(b"gauss_shifted") == 1
    return P


def test_registered_plugin_is_visible_by_name(gauss_plugin):
    assert A.lib().hb_target_registered(b"gauss_shifted") == 1
    assert gauss_plugin.hb_gauss_plugin_register() == A.HB_ERR_INVALID               # a second registration of the name is refused


@pytest.mark.gpu
def test_third_party_target_through_the_log_density_entry_points(gauss_plugin):
    from hydrobayes_b200 import HbError, log_density
    from hydrobayes_b200.sampler import DensitySession
    rng = np.random.default_rng(8)
    for d in (1, 3, 100):
        mu = np.arange(d, dtype=float) - 1.5
        x = rng.normal(0, 2, (70, d)) + mu
        want = -0.5 * ((x - mu) ** 2).sum(axis=1) - 0.5 * d * np.log(2 * np.pi)
        ll, fx = log_density("gauss_shifted", x, 2, forcing=mu)
        np.testing.assert_allclose(ll, np.maximum(want, FLOOR), rtol=1e-12, atol=1e-12)  # calc_ll's floor is the library's
        np.testing.assert_allclose(fx, want, rtol=1e-12, atol=1e-12)                     # the raw output, not floored
        with DensitySession("gauss_shifted", 2, d, 70, forcing=mu) as s:
            assert np.array_equal(s(x)[0], ll)
        ll0, _ = log_density("gauss_shifted", x, 2)                                       # no `...`: mu = 0
        np.testing.assert_allclose(ll0, np.maximum(-0.5 * (x ** 2).sum(axis=1) - 0.5 * d * np.log(2 * np.pi), FLOOR),
                                   rtol=1e-12, atol=1e-12)
    x1 = rng.normal(0, 1, (20, 2))
    ll1, fx1 = log_density("gauss_shifted", x1, 1)                                        # lik = 1: the function returns a likelihood
    np.testing.assert_allclose(fx1, np.exp(ll1), rtol=1e-12)
    with pytest.raises(HbError, match="lik has to be 1 or 2"):                             # the plugin's own init error comes through
        log_density("gauss_shifted", x1, 31)
    with pytest.raises(HbError, match="length d"):
        log_density("gauss_shifted", x1, 2, forcing=np.zeros(5))


@pytest.mark.gpu
@pytest.mark.parametrize("sweep", ["dream_parallel", "dream"])
def test_third_party_target_through_the_samplers(gauss_plugin, sweep):
    import hydrobayes_b200 as hb
    d, nc, t = 3, 48, 2000
    mu = np.array([-4.0, 0.5, 7.0])
    x0 = np.random.default_rng(1).uniform(-10, 10, (nc, d))
    res = getattr(hb, sweep)("gauss_shifted", mu, lik=2, par_info=dict(initial="user", val_ini=x0, prior="flat"), nc=nc, t=t, d=d,
                             verbose=False, seed=12)
    chain = res["chain"]
    assert chain.shape == (t, d + 3, nc)
    x = chain[t // 2:, :d, :]
    m = x.mean(axis=(0, 2)); v = x.var(axis=(0, 2))
    assert np.all(np.abs(m - mu) < 0.1), m
    assert np.all(np.abs(v - 1.0) < 0.15), v
    # ll column = the plugin's value at the stored state, lp = 0 (flat prior), lpost = lp + ll  (R/dream_parallel.R:386-394)
    xs = chain[-1, :d, :].T
    want = -0.5 * ((xs - mu) ** 2).sum(axis=1) - 0.5 * d * np.log(2 * np.pi)
    np.testing.assert_allclose(chain[-1, d + 1, :], want, rtol=1e-12, atol=1e-12)
    assert np.all(chain[-1, d, :] == 0.0) and np.array_equal(chain[-1, d + 2, :], chain[-1, d + 1, :])
