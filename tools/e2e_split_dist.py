"""Host-side phase times of the distributed dream_parallel() call (run under torchrun with HB_TIMING=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, torch, torch.distributed as dist
import bench, hydrobayes_b200 as hb
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = bench.workload("C4"); c = w["cfg"]; x0 = w["x0"]()
par = dict(initial="user", val_ini=x0, prior="flat")
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = hb.dream_parallel(w["target"], lik=c.lik, par_info=par, nc=c.nc, t=c.t, d=c.d, thin=c.thin, verbose=False, seed=1312,
                            store_fx=False, device=local, distributed=True)
    dist.barrier()
    if rank == 0:
        print(f"rep {rep}: e2e {time.perf_counter() - t0:.3f} s, device loop {res['device_loop_seconds']:.3f} s", flush=True)
    del res
dist.destroy_process_group()
