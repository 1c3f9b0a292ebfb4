"""HB_TIMING=1 python tools/e2e_api.py : phase times of the public dream_parallel() call on C4 (one GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, bench, hydrobayes_b200 as hb
w = bench.workload("C4"); c = w["cfg"]; x0 = w["x0"]()
for rep in range(4):
    t0 = time.perf_counter()
    res = hb.dream_parallel(w["target"], lik=c.lik, par_info=dict(initial="user", val_ini=x0, prior="flat"), nc=c.nc, t=c.t, d=c.d,
                            thin=c.thin, verbose=False, seed=1312, store_fx=False)
    print(f"rep {rep}: e2e {time.perf_counter() - t0:.3f} s, device loop {res['device_loop_seconds']:.3f} s", flush=True)
    del res
