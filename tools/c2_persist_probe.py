"""python tools/c2_persist_probe.py : BASELINE C2 (t_distribution d = 100, nc = 1024) through the default small-population path
(persistent generation loop) -- a short run for ncu captures of hb_persist_small_kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc, d, t = 1024, 100, 1000
cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1, thin=10)
with Sampler(cfg, 0) as s:
    s.set_target("t_distribution"); s.set_initial(H.x0_tdist(nc, d))
    s.run(120)                                     # adaptation period (100 generations) + warm-up
    for _ in range(3):
        t0 = time.perf_counter(); s.run(200); dt = time.perf_counter() - t0
        print(f"persistent slice of 200 generations: {dt / 200 * 1e6:.2f} us/generation", flush=True)
