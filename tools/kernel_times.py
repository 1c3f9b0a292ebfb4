"""Per-kernel device time (CUDA events on the library stream) for a workload, inside and after the adaptation period."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import bench
w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
from hydrobayes_b200 import Sampler
cfg = w["cfg"]
s = Sampler(cfg, 0)
kw = w["kw"]
if "par_min" in kw: s.set_bounds(kw["par_min"], kw["par_max"])
if "obs" in kw: s.set_obs(kw["obs"])
s.set_target(w["target"], kw.get("forcing"))
s.set_initial(w["x0"]())
s.run(10)
def prof(n, label):
    before = {k: s.kernel_ms(k) for k in ("K1","K2","K3","FIN","OUTLIER")}
    s.profile(True); s.run(n); s.profile(False)
    out = {}
    for k,(m0,n0) in before.items():
        m1,n1 = s.kernel_ms(k)
        if n1>n0: out[k] = round((m1-m0)/(n1-n0)*1e3,1)
    print(label, "us/launch", out, flush=True)
prof(20, "adaptation gens")
adapt_end = int(cfg.adapt*cfg.t)
done = s.run(max(0, adapt_end + 5 - 31))
prof(20, "post-adaptation gens")
import time
t0=time.perf_counter(); s.run(100); dt=time.perf_counter()-t0
print("plain 100 gens: %.1f us/gen" % (dt/100*1e6))
