"""python tools/c2_probe.py : microseconds per generation of BASELINE C2 (t_distribution d = 100, nc = 1024) after the
adaptation period: three-kernel path (HB_FUSED=0), one fused launch per generation (HB_PERSIST=0) and the default persistent
generation loop (one cooperative launch per slice, grid barrier per generation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc, d, t = 1024, 100, 2400
x0 = H.x0_tdist(nc, d)
for fused, persist in (("0", "0"), ("1", "0"), ("1", "1"), ("0", "0"), ("1", "0"), ("1", "1")):
    os.environ["HB_FUSED"] = fused; os.environ["HB_PERSIST"] = persist
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1, thin=10)
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        ta = time.perf_counter(); s.run(239); ta = (time.perf_counter() - ta) / 238   # generations 2..240: the adaptation period
        print(f"HB_FUSED={fused} HB_PERSIST={persist}: adaptation period {ta * 1e6:.2f} us/generation wall", flush=True)
        s.run(161)                                   # warm-up behind the adaptation period
        l0 = s.timing()
        t0 = time.perf_counter(); s.run(1500); dt = time.perf_counter() - t0
        l1 = s.timing()
    print(f"HB_FUSED={fused} HB_PERSIST={persist}: {dt / 1500 * 1e6:.2f} us/generation wall, {(l1[0] - l0[0]) / 1500 * 1e3:.2f} us device, "
          f"{(l1[1] - l0[1]) / 1500:.2f} launches/generation, {nc * 1500 / dt / 1e6:.1f} M chain-gen/s", flush=True)
