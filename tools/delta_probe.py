"""python tools/delta_probe.py [nc] : one GPU, a shard-sized population (default 131072 chains = one rank of C4 on 8 GPUs):
plain three-kernel path against the stage-and-apply pipeline forced on one GPU (HB_FORCE_DELTA=1, no peers) with 1 / 2 / 4
pieces -- device time per steady-state generation, host enqueue cost, launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
d, t = 100, 700
x0 = H.x0_tdist(nc, d)
modes = [("plain", None, None)] + [(f"delta P={p}", "1", str(p)) for p in (1, 2, 4)]
if len(sys.argv) > 2:
    modes = [m for m in modes if m[0].startswith(sys.argv[2])]
for name, force, pieces in modes:
    for k, v in (("HB_FORCE_DELTA", force), ("HB_PIECES", pieces)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1, thin=100)
    with Sampler(cfg, 0) as s:
        s.set_target("t_distribution"); s.set_initial(x0)
        s.run(19)
        ta = time.perf_counter(); s.run(40); ta = (time.perf_counter() - ta) / 40      # generations 21..60: adaptation period (i <= 70)
        s.profile(True); s.run(10); s.profile(False)
        ka = {k: round(s.kernel_ms(k)[0] / max(s.kernel_ms(k)[1], 1) * 1e3, 1) for k in ("K1", "K2", "K3", "FIN", "OUTLIER", "XCHG", "APPLY") if s.kernel_ms(k)[1]}
        print(f"{name:12s}: adaptation period {ta * 1e6:8.1f} us/gen wall; us/launch {ka}", flush=True)
        s.run(31)                                    # leave the adaptation period (70 generations) + warm-up
        l0 = s.timing()
        t0 = time.perf_counter(); s.run(500); dt = time.perf_counter() - t0
        l1 = s.timing()
        s.profile(True); s.run(20); s.profile(False)
        ks = {k: round(s.kernel_ms(k)[0] / max(s.kernel_ms(k)[1], 1) * 1e3, 1) for k in ("K1", "K2", "K3", "XCHG", "APPLY") if s.kernel_ms(k)[1]}
    print(f"{name:12s}: {dt / 500 * 1e6:8.1f} us/gen wall, {(l1[0] - l0[0]) / 500 * 1e3:8.1f} us/gen device, {(l1[1] - l0[1]) / 500:5.1f} launches/gen, "
          f"{nc * 500 / dt / 1e6:7.1f} M chain-gen/s; us/launch {ks}", flush=True)
