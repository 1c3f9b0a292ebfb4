"""Multi-GPU identity check (run under torchrun on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py
Every rank runs its shard of one population; rank r also runs the whole population alone on its own GPU and compares
its shard of the chain / state bit for bit (native Philox draws are keyed by the global chain id, so the result must not
depend on the number of GPUs).  Covers adaptation (statistics all-reduce), outlier resets and the all-gather exchange."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import helpers as H
from hydrobayes_b200 import Sampler, _abi as A, lib


def run(cfg, target, x0, local, rank=0, world=1, uid=None, kw={}, peer=False, shard_state=False):
    with Sampler(cfg, local) as s:
        if "par_min" in kw:
            s.set_bounds(kw["par_min"], kw["par_max"])
        if "obs" in kw:
            s.set_obs(kw["obs"])
        s.set_target(target, kw.get("forcing"))
        if world > 1:
            s.comm_init(rank, world, uid)
        s.set_initial(x0)
        if peer:
            s.peer_setup(world)
        done = 1
        while done < cfg.t:
            done = s.run(37)
        nl = cfg.nc // world
        chain = s.fetch_chain(nl)
        ar, cr = s.fetch_ar_cr()
        acc = s.fetch_accept(nl)
        xt, lp, ll, lpost = s.fetch_state(nl)
        if not shard_state and world > 1:
            xt = xt[rank * nl:(rank + 1) * nl]
        out = dict(chain=chain, AR=ar, CR=cr, accept=acc, xt=xt, lpost=lpost, outlier=s.fetch_outliers())
        if cfg.check_convergence:   # R_stat of the WHOLE population: collective on a sharded handle
            out["R_stat"] = s.fetch_rstat(); out["rstat_now"] = s.rstat()
        return out


def main():
    # same kernel variant on both sides of the bit-for-bit comparison (the single-GPU reference run of these small
    # populations would otherwise take the fused generation kernel: same decisions, last-bit differences in ll)
    os.environ["HB_FUSED"] = "0"
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = []
    nc, d = 64 * world, 100
    x0 = H.x0_tdist(nc, d); x0[3] = 300.0; x0[nc - 2] = -250.0     # forces outlier resets across shards
    cases.append(("tdist d=100", A.default_config(nc=nc, t=150, d=d, lik=2, seed=5, adapt=0.4, record_accept=1, check_convergence=1), "t_distribution", x0, {}))
    nc1 = 512 * world
    cases.append(("mixture d=1", A.default_config(nc=nc1, t=200, d=1, lik=1, seed=9, adapt=0.3, record_accept=1), "mixture", H.x0_mixture(nc1), {}))
    cases = [c + (0,) for c in cases] + [c + (1,) for c in cases if c[1].d > 16] + [c + (2,) for c in cases]
    # DREAM-ZS: archive sampling + snooker on a sharded population (all-gather and delta exchange; the peer gather does not
    # cover the archive).  x0 is the archive (m0 rows), the chains start at its last nc rows (R/dream_parallel.R:213).
    ncz, dz, m0 = 8 * world, 12, 200
    zs = A.default_config(nc=ncz, t=160, d=dz, lik=2, seed=11, past_sample=1, psnooker=0.2, m0=m0, archive_update=5, record_accept=1)
    xz = np.random.default_rng(7).uniform(-5, 15, size=(m0, dz))
    cases += [("tdist d=12 ZS+snooker", zs, "t_distribution", xz, {}, 0), ("tdist d=12 ZS+snooker", zs, "t_distribution", xz, {}, 2)]
    # MT-DREAM (mt > 1) on a sharded population: all-gather exchange of the state and of the candidate population zt
    ncm, dm = 16 * world, 20
    mtc = A.default_config(nc=ncm, t=120, d=dm, lik=2, seed=13, mt=3, adapt=0.3, record_accept=1)
    cases += [("tdist d=20 MT-DREAM mt=3", mtc, "t_distribution", H.x0_tdist(ncm, dm), {}, 0)]
    for name, cfg, tgt, x0, kw, xmode in cases:
        peer = xmode != 0
        if xmode:
            cfg = H.copy_cfg(cfg, exchange=xmode); name += [" ", " [peer exchange]", " [delta exchange]"][xmode]
        buf = C.create_string_buffer(128)
        if rank == 0:
            assert lib().hb_comm_unique_id(buf) == 0
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        sh = run(cfg, tgt, x0, local, rank, world, obj[0], kw, peer=peer, shard_state=(xmode == 1))
        full = run(cfg, tgt, x0, local, kw=kw)
        nl = cfg.nc // world; lo = rank * nl
        same = (np.array_equal(sh["chain"], full["chain"][:, :, lo:lo + nl], equal_nan=True)
                and np.array_equal(sh["accept"], full["accept"][:, lo:lo + nl])
                and np.array_equal(sh["xt"], full["xt"][lo:lo + nl]) and np.array_equal(sh["lpost"], full["lpost"][lo:lo + nl])
                and sh["outlier"] == full["outlier"])
        close = np.allclose(sh["AR"], full["AR"], rtol=1e-12, equal_nan=True) and np.allclose(sh["CR"], full["CR"], rtol=1e-9, equal_nan=True)
        if "R_stat" in full:
            rs_ok = (np.allclose(sh["R_stat"], full["R_stat"], rtol=1e-9, equal_nan=True) and np.isfinite(full["R_stat"]).any()
                     and np.allclose(sh["rstat_now"], full["rstat_now"], rtol=1e-9))
            name += f" [sharded R_stat close: {rs_ok}]"
            close = close and rs_ok
        print(f"[rank {rank}] {name}: shard == single-GPU run: {same}; AR/CR close: {close}; outliers {len(full['outlier'])}; "
              f"accept rate {full['accept'].mean():.3f}", flush=True)
        ok = ok and same and close
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if t.item() == 1.0 else "FAIL", f"world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
