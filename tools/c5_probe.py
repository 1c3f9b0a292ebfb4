"""python tools/c5_probe.py : a short slice of BASELINE C5 (4 096 batched HydroModel runs, sequential sweep) for ncu captures
of the batched target kernel (hb_hydro_runs_tma_kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

n_runs, nc, T, t = 4096, 10, 3653, 400
P, O = H.hydro_data_batch(T, n_runs, first=0)
cfg = A.default_config(nc=nc, t=t, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_BOUND,
                       seed=1312, sweep=A.HB_SWEEP_SEQUENTIAL, n_runs=n_runs)
x0 = np.concatenate([H.x0_hydro(nc) for _ in range(n_runs)])
with Sampler(cfg, 0) as s:
    s.set_bounds([0.01], [2.0]); s.set_target_batched("HydroModel", P, O); s.set_initial(x0)
    s.run(45)
    t0 = time.perf_counter(); s.run(20); dt = time.perf_counter() - t0
    print(f"C5 slice: {dt / 20 * 1e3:.3f} ms per generation ({n_runs} runs x {nc} chains), launches {s.timing()[1]}", flush=True)
