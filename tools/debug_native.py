import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np
import helpers as H
from hydrobayes_b200 import _abi as A
from oracle import hb_oracle_py as oracle
P, obs = H.hydro_data(T=200)
cfg = A.default_config(nc=20, t=80, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_REFLECT, seed=8, record_accept=1)
tgt = "HydroModel"; x0 = H.x0_hydro(20); kw = dict(par_min=[0.01], par_max=[2.0], forcing=P, obs=obs)
rec = oracle.Tape.for_config(cfg)
ora = oracle.run(cfg, tgt, x0, record=rec, **kw)
dev = H.run_device(cfg, tgt, x0, **kw)
a, b = dev["accept"], ora["accept"]
bad = np.argwhere(a != b)
print("n mismatched accepts", len(bad), "first", bad[:5])
ch_d, ch_o = dev["chain"], ora["chain"]
diff = np.abs(ch_d - ch_o) > 1e-9 * np.abs(ch_o) + 1e-12
rows = np.argwhere(diff.any(axis=(1,)))
print("first differing (row, chain):", rows[:5])
if len(rows):
    r, c = rows[0]
    print("row", r, "chain", c, "dev", ch_d[r, :, c], "ora", ch_o[r, :, c])
    print("prev row dev", ch_d[r-1, :, c], "ora", ch_o[r-1, :, c])
    g = r - 1
    print("tape: D", rec.D[g, c], "partners", rec.partners[g, c], "id", rec.id[g, c], "zt", rec.zt[g, c], "lam", rec.lambda_[g, c], "g1", rec.g_is_one[g, c], "zeta", rec.zeta[g, c], "u", rec.u_accept[g, c], "log u", np.log(rec.u_accept[g, c]))
    print("partner x (ora prev row):", ch_o[r-1, 0, rec.partners[g, c][:2*rec.D[g,c]]])
    print("accept dev", a[g, c], "ora", b[g, c])
    # evaluate the proposal density on both
    from hydrobayes_b200 import log_density
print("outliers dev", dev["outlier"][:10], "ora", ora["outlier"][:10])
