"""CPU-side checks: the C-ABI library loads here (no GPU), exports every symbol include/hydrobayes_b200.h declares,
fails loudly without a device, and the host-side logic of the R-API mirror."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hydrobayes_b200 as hb
from hydrobayes_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(A.lib_path()):
        import __graft_entry__ as g
        g.build()
    return A.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "hydrobayes_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr)) - {"hb_target_vtable"})
    declared = [s for s in declared if not s.endswith("_t")]
    assert len(declared) >= 30
    raw = C.CDLL(A.lib_path())
    for s in declared:
        assert hasattr(raw, s), f"{s} is declared in the header but not exported"
    assert sorted(A.EXPORTED_SYMBOLS) == declared


def test_struct_layout_matches_header(L):
    c = A.HbConfig()
    L.hb_config_default(C.byref(c))
    ref = A.default_config()
    for f, _ in A.HbConfig._fields_:
        assert getattr(c, f) == getattr(ref, f), f
    assert c.adapt == 0.1 and c.update_interval == 10 and c.delta == 3 and c.nCR == 3 and c.p_g == 0.2
    assert c.c_star == 1e-12 and c.seed == 1312 and c.abi_version == A.HB_ABI_VERSION


def test_builtin_targets_registered(L):
    for n in ("mixture", "t_distribution", "HydroModel"):
        assert L.hb_target_registered(n.encode()) == 1
    assert L.hb_target_registered(b"some_r_closure") == 0


@pytest.mark.skipif(A.lib().hb_device_count() > 0 if os.path.exists(A.lib_path()) else False, reason="GPU present")
def test_no_cpu_fallback(L):
    with pytest.raises(hb.HbError) as e:
        hb.dream_parallel("mixture", lik=1, par_info=dict(initial="user", val_ini=np.zeros((10, 1)), prior="flat"),
                          nc=10, t=10, d=1, verbose=False)
    assert e.value.code == A.HB_ERR_NO_DEVICE
    with pytest.raises(hb.HbError) as e:
        hb.R_stat(np.zeros((10, 1, 3)))
    assert e.value.code == A.HB_ERR_NO_DEVICE
    with pytest.raises(hb.HbError):
        hb.log_density("mixture", [[0.0]], lik=1)


def test_argument_checks_before_device(L):
    # R/dream.R:136-137 : nc must be > 2*delta -- checked before any device work
    cfg = A.default_config(nc=6, t=10, d=1, lik=1)
    h = C.c_void_p()
    assert L.hb_create(C.byref(cfg), 0, C.byref(h)) == A.HB_ERR_INVALID
    assert b"'nc' must be > 'delta' * 2" in L.hb_last_error(None)
    cfg = A.default_config(nc=10, t=10, d=1, lik=7)
    assert L.hb_create(C.byref(cfg), 0, C.byref(h)) == A.HB_ERR_INVALID
    assert b"'lik' has to be one of" in L.hb_last_error(None)
    cfg = A.default_config(nc=10, t=10, d=1, lik=1, mt=3, sweep=A.HB_SWEEP_SEQUENTIAL)   # MT-DREAM: dream_parallel only
    assert L.hb_create(C.byref(cfg), 0, C.byref(h)) == A.HB_ERR_UNSUPPORTED


def test_prior_sample_host_logic():
    rng = np.random.default_rng(0)
    x = hb.prior_sample(dict(initial="uniform", min=[-1, 0], max=[1, 10]), 2, 500, rng)
    assert x.shape == (500, 2) and x[:, 0].min() >= -1 and x[:, 1].max() <= 10
    x = hb.prior_sample(dict(initial="latin", min=-20, max=20), 1, 40, rng)
    assert np.array_equal(np.sort(np.floor((x[:, 0] + 20) / 1.0)), np.arange(40))   # one point per stratum
    x = hb.prior_sample(dict(initial="normal", mu=[0, 5], cov=[[1, 0], [0, 4]]), 2, 4000, rng)
    assert abs(x[:, 1].mean() - 5) < 0.2 and abs(x[:, 1].std() - 2) < 0.15
    v = np.arange(6.0).reshape(3, 2)
    assert np.array_equal(hb.prior_sample(dict(initial="user", val_ini=v), 2, 3), v)
    with pytest.raises(ValueError):
        hb.prior_sample(dict(initial="bogus"), 1, 3)


def test_outlier_list_structure():
    from hydrobayes_b200.sampler import _outlier_list
    cfg = A.default_config(nc=10, t=100, d=1, lik=1)
    lst = _outlier_list([(10, 3), (10, 7)], cfg, A.HB_SWEEP_SYNCHRONOUS)
    assert len(lst) == 10 and lst[0] is None and lst[9].tolist() == [3, 7]
    cfg = A.default_config(nc=10, t=100, d=1, lik=1, outlier_check=0)
    assert _outlier_list([], cfg, A.HB_SWEEP_SYNCHRONOUS) == [None]


def test_header_is_plain_c_and_a_c_consumer_links(tmp_path):
    # the boundary is a C ABI: the public headers must compile as C99 and as C++ on their own, and a plain C program (what
    # the R glue, cgo or JNI would be) links against the shared library with nothing but -lhydrobayes_b200
    import subprocess
    inc = os.path.join(ROOT, "include")
    for hdr, cxx in (("hydrobayes_b200.h", "c++11"), ("hydrobayes_b200_rng.h", "c++17")):   # the draw spec uses hex float literals
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, hdr)], check=True)
        subprocess.run(["g++", "-std=" + cxx, "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", os.path.join(inc, hdr)], check=True)
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "hydrobayes_b200.h"
int main(void) {
  hb_config c;
  hb_config_default(&c);
  if (c.abi_version != HB_ABI_VERSION || c.delta != 3 || c.nCR != 3 || c.update_interval != 10 || c.thin != 1) return 10;
  c.nc = 8; c.t = 10; c.d = 2; c.lik = 2;
  hb_handle* h = NULL;
  int rc = hb_create(&c, 0, &h);
  if (hb_device_count() == 0) {                 /* no CUDA device: must fail loudly, there is no CPU path */
    if (rc != HB_ERR_NO_DEVICE || h != NULL) return 11;
    if (!strstr(hb_last_error(NULL), "no CPU fallback")) return 12;
  } else {
    if (rc != HB_OK || h == NULL) return 13;
    hb_destroy(h);
    hb_wait_idle();
  }
  c.nc = 6;                                     /* nc <= 2 * delta: the reference's stop() wording (R/dream.R:136-137) */
  rc = hb_create(&c, 0, &h);
  if (rc != HB_ERR_INVALID || !strstr(hb_last_error(NULL), "'nc' must be > 'delta' * 2")) return 14;
  printf("targets: %d %d %d %d\n", hb_target_registered("mixture"), hb_target_registered("t_distribution"),
         hb_target_registered("HydroModel"), hb_target_registered("no_such_closure"));
  return 0;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.join(ROOT, "hydrobayes_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + inc, str(src), "-o", str(exe), "-L" + libdir, "-lhydrobayes_b200",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.strip() == "targets: 1 1 1 0"


def test_exported_target_callables_shapes(monkeypatch):
    # mixture(x) / t_distribution(x) of the reference's NAMESPACE (R/test_mixture.R:16, R/test_t_distribution.R:19-46): the
    # argument handling of the Python callables, with the device evaluation replaced by numpy
    import hydrobayes_b200 as hb
    import hydrobayes_b200.sampler as S
    calls = []

    def fake(fun, x, lik, **kw):
        x = np.atleast_2d(np.asarray(x, float)); calls.append((fun, x.shape, lik))
        v = x.sum(axis=1)
        return v - 1.0, v
    monkeypatch.setattr(S, "log_density", fake)
    assert hb.mixture(2.0) == 2.0 and calls[-1] == ("mixture", (1, 1), 1)                     # a scalar: one point
    m = hb.mixture(np.arange(6.0).reshape(2, 3))                                              # elementwise, shape kept
    assert m.shape == (2, 3) and np.array_equal(m, np.arange(6.0).reshape(2, 3)) and calls[-1] == ("mixture", (6, 1), 1)
    assert hb.t_distribution([1.0, 2.0, 3.0]) == 6.0 and calls[-1] == ("t_distribution", (1, 3), 2)   # a vector: ONE point of dimension d
    t = hb.t_distribution(np.ones((4, 5)))                                                    # a matrix: one point per row
    assert t.shape == (4,) and np.all(t == 5.0) and calls[-1] == ("t_distribution", (4, 5), 2)
