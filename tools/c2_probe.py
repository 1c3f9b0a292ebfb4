"""python tools/c2_probe.py : microseconds per generation of BASELINE C2 (t_distribution d = 100, nc = 1024) after the
adaptation period, experimental fused small-population kernel (HB_FUSED=1) against the default three-kernel path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc, d, t = 1024, 100, 2400
x0 = H.x0_tdist(nc, d)
for fused in ("0", "1", "0", "1"):
    os.environ["HB_FUSED"] = fused
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1, thin=10)
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        s.run(400)                                   # adaptation period (240 generations) + warm-up
        l0 = s.timing()
        t0 = time.perf_counter(); s.run(1500); dt = time.perf_counter() - t0
        l1 = s.timing()
    print(f"HB_FUSED={fused}: {dt / 1500 * 1e6:.2f} us/generation wall, {(l1[0] - l0[0]) / 1500 * 1e3:.2f} us device, "
          f"{(l1[1] - l0[1]) / 1500:.2f} launches/generation, {nc * 1500 / dt / 1e6:.1f} M chain-gen/s", flush=True)
