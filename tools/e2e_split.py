import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, bench
from hydrobayes_b200 import Sampler
w = bench.workload("C4"); cfg = w["cfg"]
x0 = w["x0"]()
T = {}
t0 = time.perf_counter(); s = Sampler(cfg, 0); s.set_target(w["target"]); T["create"] = time.perf_counter() - t0
t0 = time.perf_counter(); s.set_initial(x0); T["set_initial"] = time.perf_counter() - t0
t0 = time.perf_counter(); s.run(cfg.t); T["run"] = time.perf_counter() - t0
t0 = time.perf_counter(); ch = s.fetch_chain(); T["fetch_chain"] = time.perf_counter() - t0
t0 = time.perf_counter(); s.close(); T["close"] = time.perf_counter() - t0
print({k: round(v, 3) for k, v in T.items()}, "chain GB", ch.nbytes / 1e9)
