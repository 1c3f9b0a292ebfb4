"""Generates the committed fixtures under tests/golden/ (run from the repo root: python tools/make_golden.py).

Provenance -- read before trusting them: the reference (tpilz/HydroBayes) is interpreted R with no test suite, no
fixtures and no golden numbers, and R is not installed in the build image (F3).  These vectors are therefore
outputs of the CPU ORACLE (oracle/hb_oracle.c, a restatement of the reference by code reading), frozen so that
  * the oracle cannot drift silently (tests/test_golden.py, CPU), and
  * the CUDA path is compared with numbers that do not depend on the oracle build of the GPU box (tests, -m gpu).
They do NOT pin the oracle to the real R package: parity stays "unpinned" until someone runs tools/record_reference.R.
known_answers.json holds SURVEY Appendix B (values derived from the reference's formulas with numpy/scipy).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from hydrobayes_b200 import _abi as A
from oracle import hb_oracle_py as oracle

OUT = os.path.join(ROOT, "tests", "golden")
TAPE_FIELDS = ("D", "partners", "id", "zt", "lambda_", "g_is_one", "zeta", "u_accept", "outlier_new")
CFG_FIELDS = ("nc", "t", "d", "lik", "seed", "adapt", "thin", "burnin", "update_interval", "delta", "nCR", "c_val", "c_star",
              "p_g", "beta0", "prior", "bound", "glue_shape", "store_fx")


def cfg_dict(cfg):
    return {k: getattr(cfg, k) for k in CFG_FIELDS}


def save_case(name, cfg, target, x0, replay, **kw):
    if replay:   # the oracle draws with Philox and records the decision tape; the fixture is the replay of that tape
        rec = oracle.Tape.for_config(cfg)
        r0 = oracle.run(cfg, target, x0, record=rec, **kw)
        assert r0["rc"] == 0, r0["err"]
        cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_TAPE, record_accept=1)
        res = oracle.run(cfg_t, target, x0, tape=rec, **kw)
        tape = {f"tape_{f}": getattr(rec, f) for f in TAPE_FIELDS}
        tape["tape_n_events"] = np.array(rec.n_events)
    else:
        cfg_t = H.copy_cfg(cfg, rng_mode=A.HB_RNG_PHILOX, record_accept=1)
        res = oracle.run(cfg_t, target, x0, **kw)
        tape = {}
    assert res["rc"] == 0, res["err"]
    og = np.array(res["outlier"], dtype=np.int64).reshape(-1, 2)
    extra = {f"kw_{k}": np.asarray(v, dtype=np.float64) for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg_dict(cfg)), target=target, x0=x0,
                        replay=np.array(int(replay)), chain=res["chain"], AR=res["AR"], CR=res["CR"], accept=res["accept"],
                        outlier=og, xt=res["xt"], lpost=res["lpost"], **tape, **extra)
    print(f"{name}: chain {res['chain'].shape}, accept rate {res['accept'].mean():.3f}, outliers {len(og)}")


def main():
    os.makedirs(OUT, exist_ok=True)
    # BASELINE C1 shape (mixture, nc = 10, d = 1), shortened to t = 400
    save_case("c1_mixture_replay", A.default_config(nc=10, t=400, d=1, lik=1, seed=1312, store_fx=1), "mixture",
              H.x0_mixture(10), True)
    # d = 100 t-distribution (C2/C4 shape) with few chains: exercises the wide kernels and the DMMA plugin
    save_case("tdist_d100_replay", A.default_config(nc=24, t=60, d=100, lik=2, seed=7, adapt=0.5), "t_distribution",
              H.x0_tdist(24, 100), True)
    # native Philox draws (slot map v2): no tape, the draws themselves are part of the fixture
    save_case("tdist_d8_native", A.default_config(nc=16, t=120, d=8, lik=2, seed=2024, adapt=0.5), "t_distribution",
              H.x0_tdist(16, 8), False)
    # HydroModel with the GLUE likelihood, uniform prior and reflecting bounds (C3 shape, short series)
    P, obs = H.hydro_data(T=200)
    save_case("hydro_glue_replay", A.default_config(nc=12, t=150, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM,
                                                    bound=A.HB_BOUND_REFLECT, seed=3, adapt=0.4), "HydroModel", H.x0_hydro(12),
              True, par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
    known = {
        "_provenance": "SURVEY.md Appendix B: derived from the reference's formulas with numpy/scipy, not from reference tests",
        "log_xmin": -708.3964185322641,
        "mixture": {"-8": 0.06649038006690544, "0": 8.420452447111373e-16, "10": 0.33245190033452726, "1": 1.0279773571668917e-18},
        "t_distribution_zeros": {"1": -0.9231050070343514, "2": -2.040609620463428, "100": -213.43955272792834},
        "t_distribution_ones": {"1": -1.4272487165462735, "2": -2.582068622322943, "100": -218.01512992854524},
        "hydromodel_forcing_0_1_2_0_5_K_0.5": [1 / 3, 5 / 9, 1.037037037037037, 0.691358024691358, 2.1275720164609053],
        "rstat_sin_case": [1.0113185866169283, 1.0735661244327646],
        "philox4x32_10": [{"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
                          {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
                          {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
                           "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]}],
    }
    json.dump(known, open(os.path.join(OUT, "known_answers.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
