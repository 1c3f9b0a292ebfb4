"""Converts the draw log written by tools/record_reference.R (a real R run of the reference with runif / rnorm / sample
masked by recording wrappers) into the decision-level tape of SURVEY Appendix D (oracle.Tape / hb_tape), and back.

Draw order of dream_parallel (mt = 1, no archive), R/dream_parallel.R:269 + R/internals.R:115-219, :343 + :241, :417 + :81:
  per generation i = 2..t:
    for chain j = 1..nc (calc_prop):  runif(1) [snooker test, always drawn :119] ; sample(1:delta, 1) -> D [:123] ;
        sample((1:nc)[-j], 2D) -> partners, first D = a, next D = b [:142-144] ; sample(1:nCR, 1, prob = pCR) -> id [:173] ;
        runif(d) -> zt [:175] ; runif(d*, -c, c) -> lambda [:188] ; sample(c(gamma_d, 1), 1, prob) -> g [:192] ;
        rnorm(d*, sd = c_star) -> zeta [:194]           (d* = #{zt < CR[id]}, or 1 when that set is empty :181-184)
    for chain j = 1..nc:  runif(1) -> u_accept [:241]
    if i <= adapt*t and i %% updateInterval == 0 and outliers exist:  sample((1:N)[-outliers], n_out) [:81]
The log stores the RESULTS of the calls, so the tape does not depend on R's sample() internals or the R version.

`tape_to_drawlog` writes the same format from an oracle tape: it exists so that the parser can be tested without R.
"""
import numpy as np


def _fmt(v):
    return " ".join(repr(float(x)) for x in np.atleast_1d(v))


def tape_to_drawlog(path, tape, cfg, outlier_events=None):
    """outlier_events: {generation i: n_out} (only generations where outliers existed draw a sample())."""
    d, nc, nCR = cfg.d, cfg.nc, cfg.nCR
    CR = (np.arange(nCR) + 1) / nCR
    outlier_events = outlier_events or {}
    with open(path, "w") as f:
        for g in range(cfg.t - 1):
            i = g + 2
            for j in range(nc):
                D = int(tape.D[g, j]); idq = int(tape.id[g, j])
                f.write(f"runif,1,{_fmt(tape.u_snooker[g, j] if tape.u_snooker is not None else 0.5)}\n")
                f.write(f"sample,1,{D}\n")
                f.write(f"sample,{2 * D},{_fmt(tape.partners[g, j, :2 * D] + 1)}\n")          # R ids are 1-based
                f.write(f"sample,1,{idq + 1}\n")
                zt = tape.zt[g, j]
                f.write(f"runif,{d},{_fmt(zt)}\n")
                ds = max(int((zt < CR[idq]).sum()), 1)
                f.write(f"runif,{ds},{_fmt(tape.lambda_[g, j, :ds])}\n")
                gamma_d = 2.38 / np.sqrt(2.0 * D * ds)
                f.write(f"sample,1,{_fmt(1.0 if tape.g_is_one[g, j] else gamma_d)}\n")
                f.write(f"rnorm,{ds},{_fmt(tape.zeta[g, j, :ds])}\n")
            for j in range(nc):
                f.write(f"runif,1,{_fmt(tape.u_accept[g, j])}\n")
            if i in outlier_events:
                e = i // cfg.update_interval - 1
                n_out = outlier_events[i]
                f.write(f"sample,{n_out},{_fmt(tape.outlier_new[e, :n_out] + 1)}\n")


def drawlog_to_tape(path, cfg, Tape):
    """Tape: the oracle's Tape class.  Returns (tape, outlier_events)."""
    d, nc, nCR = cfg.d, cfg.nc, cfg.nCR
    CR = (np.arange(nCR) + 1) / nCR
    calls = []
    with open(path) as f:
        for line in f:
            name, n, vals = line.rstrip("\n").split(",", 2)
            calls.append((name, int(n), np.array([float(v) for v in vals.split()])))
    pos = 0

    def take(name, n=None):
        nonlocal pos
        if pos >= len(calls):
            raise ValueError("draw log ended early")
        nm, k, v = calls[pos]
        if nm != name or (n is not None and k != n):
            raise ValueError(f"draw log entry {pos}: expected {name}({n}), found {nm}({k})")
        pos += 1
        return v

    tape = Tape.for_config(cfg)
    if tape.u_snooker is None:
        tape.u_snooker = np.ones((tape.n_gen, tape.nct))
    outlier_events = {}
    for g in range(cfg.t - 1):
        i = g + 2
        for j in range(nc):
            tape.u_snooker[g, j] = take("runif", 1)[0]
            D = int(take("sample", 1)[0])
            tape.D[g, j] = D
            tape.partners[g, j, :2 * D] = take("sample", 2 * D).astype(np.int32) - 1
            idq = int(take("sample", 1)[0]) - 1
            tape.id[g, j] = idq
            zt = take("runif", d)
            tape.zt[g, j] = zt
            ds = max(int((zt < CR[idq]).sum()), 1)
            tape.lambda_[g, j, :ds] = take("runif", ds)
            tape.g_is_one[g, j] = 1 if take("sample", 1)[0] == 1.0 else 0
            tape.zeta[g, j, :ds] = take("rnorm", ds)
        for j in range(nc):
            tape.u_accept[g, j] = take("runif", 1)[0]
        if pos < len(calls) and calls[pos][0] == "sample" and i <= cfg.adapt * cfg.t and i % cfg.update_interval == 0:
            # an outlier event: the next generation would start with runif(1) (snooker test), never with sample()
            v = take("sample")
            e = i // cfg.update_interval - 1
            tape.outlier_new[e, :v.size] = v.astype(np.int32) - 1
            outlier_events[i] = int(v.size)
    if pos != len(calls):
        raise ValueError(f"{len(calls) - pos} unread entries at the end of the draw log")
    return tape, outlier_events
