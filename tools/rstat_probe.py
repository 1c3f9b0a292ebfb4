"""python tools/rstat_probe.py : K4 (R_stat, R/gelman_rubin.R:21-56) on a C4-sized stored history (nc = 2^20, d = 100,
12 stored generations = 10 GB) -- CUDA-event time per call and a target for ncu captures of hb_rstat_*."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A

nc, d, t = 1 << 20, 100, 12
cfg = A.default_config(nc=nc, t=t, d=d, lik=2, seed=1312, thin=1)
with Sampler(cfg, 0) as s:
    s.set_target("t_distribution"); s.set_initial(H.x0_tdist(nc, d))
    s.run(t)
    s.rstat()
    s.profile(True)
    t0 = time.perf_counter()
    for _ in range(3):
        r = s.rstat()
    dt = (time.perf_counter() - t0) / 3
    ms, n = s.kernel_ms("K4")
    rows = s.ind_out()
    algo = 8.0 * (rows // 2 + 1) * d * nc
    print(f"R_stat over {rows} stored rows x {nc} chains x d={d}: {dt * 1e3:.3f} ms wall per call, K4 device {ms / max(n, 1) * 2:.3f} ms per call; "
          f"algorithmic bytes {algo / 1e9:.2f} GB -> {algo / (ms / max(n, 1) * 2 * 1e-3) / 1e9:.0f} GB/s; max R_stat {r.max():.3f}", flush=True)
