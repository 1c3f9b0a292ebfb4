"""The reference-side binding (r-package/): `.Call` glue compiled against a stand-in for R's C API (tests/mock_r) --
R itself is absent from the build image.  CPU tests: the glue compiles, registers every symbol the R drivers `.Call`
with the arity they pass, and turns a C-ABI failure into an R error (stop) with the library's message.  GPU test: the
entry points, driven in the order r-package/R/hb_drive.R calls them with R's column-major arrays, give exactly what the
Python host mirror gives for the same seed (both sit on the same C ABI) and agree with the oracle."""
import os
import re
import sys

import numpy as np
import pytest

import helpers as H
from hydrobayes_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "mock_r"))
import rmock  # noqa: E402


@pytest.fixture(scope="module")
def R():
    r = rmock.MockR()
    yield r
    r.free_all()


def _r_sources():
    src = {}
    for f in sorted(os.listdir(os.path.join(ROOT, "r-package", "R"))):
        src[f] = open(os.path.join(ROOT, "r-package", "R", f)).read()
    return src


def test_every_dotcall_in_the_r_drivers_is_registered_with_matching_arity(R):
    entries = R.entries()
    assert R.L.mockR_dynamic_symbols() == 0                       # R_useDynamicSymbols(dll, FALSE)
    used = {}
    for f, text in _r_sources().items():
        for m in re.finditer(r"\.Call\((hbR_\w+)((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)", text):
            name, rest = m.group(1), m.group(2)
            depth = 0; nargs = 0
            for ch in rest:                                       # top-level commas = number of arguments after the symbol
                depth += ch in "([" ; depth -= ch in ")]"
                nargs += (ch == "," and depth == 0)
            used.setdefault(name, set()).add(nargs)
    assert used, "no .Call sites found"
    for name, arities in used.items():
        assert name in entries, f"{name} is .Call'ed by the R drivers but not registered in hb_rglue.c"
        assert arities == {entries[name]}, (name, arities, entries[name])
    assert set(entries) == set(used), "registered but never called from R: " + ", ".join(sorted(set(entries) - set(used)))


def test_r_formals_match_the_reference_signature():
    # R/dream_parallel.R:155-163 and R/dream.R:126-133: same formals, same order, same defaults; additions come last
    src = _r_sources()
    want_par = ["fun", "...", "lik", "par.info", "nc", "t", "d", "burnin", "adapt", "updateInterval", "delta", "c_val", "c_star",
                "nCR", "p_g", "beta0", "thin", "outlier_check", "obs", "abc_rho", "abc_e", "glue_shape", "lik_fun"]
    tails = {"dream_parallel": ["past_sample", "m0", "archive_update", "psnooker", "mt", "ncores", "checkConvergence", "verbose",
                                "seed", "device", "series_fx"],
             "dream": ["checkConvergence", "verbose", "seed", "device", "series_fx"]}
    text = src["hb_entry_points.R"]
    heads = {}
    for fn, tail in tails.items():
        start = text.index(fn + " <- function(") + len(fn + " <- function(")
        head = text[start: text.index(") {", start)]
        heads[fn] = head
        head = re.sub(r"list\([^)]*\)", "LIST", head)
        names = [a.split("=")[0].strip() for a in head.split(",")]
        assert names == want_par + tail, (fn, names)
    assert 'prior = "flat"' in heads["dream_parallel"] and 'prior = "uniform"' in heads["dream"]
    for frag in ["burnin = 0", "adapt = 0.1", "updateInterval = 10", "delta = 3", "c_val = 0.1", "c_star = 1e-12", "nCR = 3",
                 "p_g = 0.2", "beta0 = 1", "thin = 1", "outlier_check = TRUE"]:
        assert frag in heads["dream_parallel"] and frag in heads["dream"], frag


def test_dotcall_checks_symbol_and_arity_like_r(R):
    with pytest.raises(rmock.RError, match="not in DLL"):
        R.call("hbR_no_such_entry", R.nil)
    with pytest.raises(rmock.RError, match="Incorrect number of arguments"):
        R.call("hbR_run", R.nil)


@pytest.mark.skipif(A.lib().hb_device_count() > 0, reason="needs a machine WITHOUT a CUDA device")
def test_no_device_becomes_an_r_error(R):
    cfg = R.named_list(sweep=0, nc=8, t=10, d=2, lik=2)
    with pytest.raises(rmock.RError, match="(?i)cuda|device"):     # there is no CPU fallback behind the glue either
        R.call("hbR_create", cfg, R.integer(0))
    x = R.real(np.zeros((60, 2, 3)))
    with pytest.raises(rmock.RError, match="(?i)cuda|device"):
        R.call("hbR_rstat_array", x)
    with pytest.raises(rmock.RError, match="3-d array"):
        R.call("hbR_rstat_array", R.real(np.zeros(10)))
    with pytest.raises(rmock.RError, match="(?i)cuda|device"):
        R.call("hbR_density_open", R.string("t_distribution"), R.integer(2), R.nil, R.nil, R.nil, R.integer(3), R.real([8.0]))
    assert R.to_numpy(R.call("hbR_target_registered", R.string("t_distribution")))[0] == 1
    assert R.to_numpy(R.call("hbR_target_registered", R.string("my_r_closure")))[0] == 0
    # the wrapper name of the reference's example script resolves like the package's own targets (script_examples.R:495-501)
    assert R.to_numpy(R.call("hbR_target_registered", R.string("fun_wrap_glue")))[0] == 1


def _hb_drive(R, sweep, fun, dots, lik, par_info, nc, t, d, z, seed, adapt=0.1, updateInterval=10, obs=None, glue_shape=None,
              checkConvergence=False, past_sample=False, prior="flat", bound=None, **cfg_kw):
    """r-package/R/hb_drive.R, statement by statement, with the mock's .Call."""
    prior_code = ["flat", "uniform", "normal"].index(prior)
    bound_code = 0 if bound is None else ["bound", "reflect", "fold"].index(bound) + 1
    store_fx = fun in ("mixture", "t_distribution")
    cfg = R.named_list(sweep=sweep, nc=nc, t=t, d=d, lik=lik, adapt=adapt, updateInterval=updateInterval, prior=prior_code,
                       bound=bound_code, past_sample=past_sample, m0=None, archive_update=None, glue_shape=glue_shape,
                       seed=seed, checkConvergence=checkConvergence, store_fx=store_fx, **cfg_kw)
    h = R.call("hbR_create", cfg, R.integer(0))
    try:
        if par_info.get("min") is not None:
            R.call("hbR_set_bounds", h, R.real(np.atleast_1d(par_info["min"])), R.real(np.atleast_1d(par_info["max"])))
        if obs is not None:
            R.call("hbR_set_obs", h, R.real(obs))
        R.call("hbR_set_target", h, R.string(fun), R.real(dots[0]) if dots else R.nil)
        R.call("hbR_set_initial", h, R.real(z))                     # an R matrix: column-major
        dims = R.to_numpy(R.call("hbR_out_dims", h)).astype(int)
        out_t = dims[0]
        chain = R.real_na(out_t, d + 3, nc)                          # array(NA_real_, dim = c(out_t, d + 3, nc))
        done, slices = 1, 0
        while done < t:
            done = int(R.to_numpy(R.call("hbR_run", h, R.real([max(1, t // 50)])))[0]); slices += 1
        R.call("hbR_fetch_chain", h, chain)
        out = dict(chain=R.to_numpy(chain, (out_t, d + 3, nc)), slices=slices)
        out["AR"] = R.to_numpy(R.call("hbR_fetch_ar", h, R.real_na(dims[1], dims[2])), (dims[1], dims[2]))
        out["CR"] = R.to_numpy(R.call("hbR_fetch_cr", h, R.real_na(dims[3], dims[4])), (dims[3], dims[4]))
        if store_fx:
            out["fx"] = R.to_numpy(R.call("hbR_fetch_fx", h, R.real_na(out_t, nc, 1)), (out_t, nc, 1))
        op = R.call("hbR_fetch_outliers", h)
        out["outlier"] = list(zip(R.to_numpy(R.list_elt(op, 0)).tolist(), R.to_numpy(R.list_elt(op, 1)).tolist()))
        if checkConvergence and dims[6] > 0:
            out["R_stat"] = R.to_numpy(R.call("hbR_fetch_rstat", h, R.real_na(dims[6], d)), (dims[6], d))
    finally:
        R.call("hbR_destroy", h)                                     # on.exit(.Call(hbR_destroy, h))
    with pytest.raises(rmock.RError, match="already destroyed"):
        R.call("hbR_run", h, R.real([1.0]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("sweep", [A.HB_SWEEP_SYNCHRONOUS, A.HB_SWEEP_SEQUENTIAL])
def test_dotcall_sequence_matches_host_mirror_and_oracle(R, sweep):
    from oracle import hb_oracle_py as oracle
    nc, t, d, seed = 24, 130, 6, 77
    x0 = H.x0_tdist(nc, d)
    x0[5] = 300.0                                                    # one chain far out: an outlier reset must happen
    got = _hb_drive(R, sweep, "t_distribution", (), 2, {}, nc, t, d, x0, seed, adapt=0.5, checkConvergence=True)
    assert got["slices"] == -(-(t - 1) // max(1, t // 50))
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=seed, sweep=sweep, adapt=0.5, check_convergence=1, store_fx=1)
    dev = H.run_device(cfg, "t_distribution", x0)
    assert np.array_equal(got["chain"], dev["chain"], equal_nan=True)         # same C ABI underneath: bit-identical
    assert np.array_equal(got["AR"], dev["AR"], equal_nan=True) and np.array_equal(got["CR"], dev["CR"], equal_nan=True)
    assert np.array_equal(got["fx"][:, :, 0], dev["fx"], equal_nan=True)
    assert np.array_equal(got["R_stat"], dev["R_stat"], equal_nan=True)
    assert len(got["outlier"]) > 0
    assert got["outlier"] == [(g, c + 1) for g, c in dev["outlier"]]             # R side: 1-based chain numbers
    ora = oracle.run(cfg, "t_distribution", x0)
    np.testing.assert_allclose(got["chain"], ora["chain"], rtol=1e-9, atol=1e-12, equal_nan=True)
    r = R.to_numpy(R.call("hbR_rstat_array", R.real(got["chain"][:, :d, :])))  # R_stat(chain[, 1:d, ])
    np.testing.assert_allclose(r, oracle.rstat(got["chain"][:, :d, :]), rtol=1e-10)


@pytest.mark.gpu
def test_dotcall_hydromodel_with_bounds_and_library_errors(R):
    from oracle import hb_oracle_py as oracle
    P, obs = H.hydro_data(T=120)
    nc, t = 16, 90
    x0 = H.x0_hydro(nc)
    got = _hb_drive(R, 0, "HydroModel", (P,), 31, dict(min=0.01, max=2.0), nc, t, 1, x0, 5, adapt=0.4, obs=obs, glue_shape=50.0,
                    prior="uniform", bound="reflect")
    cfg = A.default_config(nc=nc, t=t, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_REFLECT, seed=5,
                           adapt=0.4)
    ora = oracle.run(cfg, "HydroModel", x0, par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
    np.testing.assert_allclose(got["chain"], ora["chain"], rtol=1e-9, atol=1e-12, equal_nan=True)
    # stop() sites of the reference, worded by the library: nc <= 2 * delta (R/dream_parallel.R:166-167), unknown target
    with pytest.raises(rmock.RError, match="'nc' must be > 'delta' \\* 2"):
        R.call("hbR_create", R.named_list(sweep=0, nc=6, t=10, d=2, lik=2), R.integer(0))
    h = R.call("hbR_create", R.named_list(sweep=0, nc=8, t=10, d=2, lik=2), R.integer(0))
    with pytest.raises(rmock.RError, match="my_r_closure"):
        R.call("hbR_set_target", h, R.string("my_r_closure"), R.nil)
    R.call("hbR_destroy", h)


@pytest.mark.gpu
def test_dotcall_density_session(R):
    """hbR_density_open / _eval / _close as r-package/R/hb_density.R calls them: R matrices in, the library's log-density out"""
    from hydrobayes_b200 import log_density
    rng = np.random.default_rng(4)
    x = rng.normal(0, 2, (40, 5))
    s = R.call("hbR_density_open", R.string("t_distribution"), R.integer(2), R.nil, R.nil, R.nil, R.integer(5), R.real([40.0]))
    got = R.to_numpy(R.call("hbR_density_eval", s, R.real(x)))
    assert np.array_equal(got, log_density("t_distribution", x, 2)[0])
    one = R.to_numpy(R.call("hbR_density_eval", s, R.real(x[:1])))           # pdf(xp) of rwm(): one point, a 1 x d matrix
    np.testing.assert_allclose(one, got[:1], rtol=1e-12)
    with pytest.raises(rmock.RError, match="matrix"):
        R.call("hbR_density_eval", s, R.real(x[0]))
    with pytest.raises(rmock.RError, match="max_n"):
        R.call("hbR_density_eval", s, R.real(np.zeros((41, 5))))
    with pytest.raises(rmock.RError, match="opened for d = 5"):
        R.call("hbR_density_eval", s, R.real(np.zeros((4, 3))))              # fewer columns than d: refused, not read past the end
    R.call("hbR_density_close", s)
    with pytest.raises(rmock.RError, match="already closed"):
        R.call("hbR_density_eval", s, R.real(x))
    P, obs = H.hydro_data(T=150)
    k = np.linspace(0.05, 1.5, 24)[:, None]
    s = R.call("hbR_density_open", R.string("HydroModel"), R.integer(31), R.real(P), R.real(obs), R.real([50.0]), R.integer(1),
               R.real([24.0]))
    got = R.to_numpy(R.call("hbR_density_eval", s, R.real(k)))
    assert np.array_equal(got, log_density("HydroModel", k, 31, forcing=P, obs=obs, glue_shape=50.0)[0])
    R.call("hbR_density_close", s)
    with pytest.raises(rmock.RError, match="unknown target"):
        R.call("hbR_density_open", R.string("my_r_closure"), R.integer(2), R.nil, R.nil, R.nil, R.integer(3), R.real([8.0]))


def test_r_sources_are_lexically_balanced():
    # no R parser here: at least every bracket of the R drivers closes (strings and comments skipped), the files end at
    # top level, and every function the drivers call is either base R, the reference's own (kept) internals, or defined here
    for f, text in _r_sources().items():
        stack, i, n = [], 0, len(text)
        pairs = {")": "(", "]": "[", "}": "{"}
        while i < n:
            ch = text[i]
            if ch == "#":
                while i < n and text[i] != "\n":
                    i += 1
                continue
            if ch in "\"'":
                q = ch; i += 1
                while i < n and text[i] != q:
                    i += 2 if text[i] == "\\" else 1
            elif ch in "([{":
                stack.append((ch, text.count("\n", 0, i) + 1))
            elif ch in ")]}":
                assert stack and stack[-1][0] == pairs[ch], f"{f}: unbalanced {ch!r} at line {text.count(chr(10), 0, i) + 1}"
                stack.pop()
            i += 1
        assert not stack, f"{f}: unclosed {stack[-1][0]!r} opened at line {stack[-1][1]}"
    text = "\n".join(_r_sources().values())
    defined = set(re.findall(r"^\s*([.\w]+) <- function\(", text, flags=re.M))
    assert {"dream", "dream_parallel", "R_stat", ".hb_drive", "hb_logdensity", "hb_pdf", ".hb_game_lpost", "metropolis_replicas",
            "de_mc_device"} <= defined
    called = set(re.findall(r"(?<![\w.$])([.A-Za-z][\w.]*)\(", re.sub(r"#.*", "", text)))
    known = defined | {"function", "list", "c", "if", "for", "while", "stop", "match", "is.na", "is.null", "is.character", "as.integer",
                       "as.numeric", "as.matrix", "aperm", "nrow", "ncol", "storage.mode", "array", "matrix", "paste0", "dimnames", "colnames", "length", "floor", "max",
                       "seq_len", "invisible", "on.exit", "Sys.time", "sample.int", "txtProgressBar", "setTxtProgressBar", "close",
                       "rep", "lapply", "vapply", "apply", "sqrt", "diag", "t", "chol", "cov", "ifelse", "rnorm", "runif", "pmin",
                       "log", "exp", "numeric", "is.matrix", "match.arg", "return",
                       "prior_sample", "prior_pdf"}   # the reference's own R/internals.R:26-69, kept
    unknown = {c for c in called if c not in known and not c.startswith(".Call")}
    assert not unknown, f"R drivers call functions that are neither base R nor defined in the package: {sorted(unknown)}"
