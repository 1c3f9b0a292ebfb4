"""python tools/c2_adapt_probe.py : BASELINE C2 inside the adaptation period -- microseconds per generation with the persistent
kernel (default) and with the per-generation kernels (HB_PERSIST_ADAPT=0), outlier check on and off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc, d, t = 1024, 100, 20000
x0 = H.x0_tdist(nc, d)
for pa, oc in (("1", 1), ("0", 1), ("1", 0), ("0", 0), ("1", 1)):
    os.environ["HB_PERSIST_ADAPT"] = pa
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1, thin=10, outlier_check=oc)
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        s.run(100)
        l0 = s.timing()
        t0 = time.perf_counter(); s.run(1500); dt = time.perf_counter() - t0
        l1 = s.timing()
    print(f"HB_PERSIST_ADAPT={pa} outlier_check={oc}: {dt / 1500 * 1e6:.2f} us/generation wall, {(l1[0] - l0[0]) / 1500 * 1e3:.2f} us device, "
          f"{(l1[1] - l0[1]) / 1500:.2f} launches/generation", flush=True)
