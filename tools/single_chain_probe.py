"""Per-iteration cost of the single-chain samplers' target evaluation: log-density session (hb_density_eval) against the
one-shot hb_logdensity (context + buffers set up per call).  python tools/single_chain_probe.py  (needs a GPU)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hydrobayes_b200 as hb  # noqa: E402
from hydrobayes_b200.sampler import DensitySession  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as H  # noqa: E402

for fun, d, kw, lik in (("t_distribution", 100, {}, 2), ("HydroModel", 1, None, 31)):
    if kw is None:
        P, obs = H.hydro_data(T=3653)
        kw = dict(forcing=P, obs=obs, glue_shape=50.0)
    for n in (1, 1024):
        rng = np.random.default_rng(0)
        x = np.abs(rng.normal(0.5, 0.1, (n, d))) if fun == "HydroModel" else rng.normal(0, 1, (n, d))
        hb.log_density(fun, x, lik, **kw)                                  # warm-up (context, module load)
        t0 = time.perf_counter()
        for _ in range(50):
            hb.log_density(fun, x, lik, **kw)
        one_shot = (time.perf_counter() - t0) / 50
        with DensitySession(fun, lik, d, n, **kw) as s:
            s(x)
            t0 = time.perf_counter()
            for _ in range(500):
                s(x)
            sess = (time.perf_counter() - t0) / 500
        print(f"{fun:15s} d={d:3d} points per call {n:5d}: one-shot {one_shot * 1e6:8.1f} us, session {sess * 1e6:8.1f} us per evaluation")
t0 = time.perf_counter()
res = hb.rwm("t_distribution", lik=2, par_info=dict(initial="user", val_ini=0.2 * H.x0_tdist(64, 100), prior="flat"), t=2000, d=100,
             n_chains=64, seed=1)
dt = time.perf_counter() - t0
print(f"rwm t_distribution d=100, 64 replicas, t=2000: {dt:.3f} s = {dt / 1999 * 1e6:.1f} us per iteration")
