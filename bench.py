#!/usr/bin/env python
"""bench.py -- chain-generations/s of the DREAM generation step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|C3] [--impl reference]

One "step" is one slice of GENS_PER_STEP = 10 synchronous DREAM generations (one updateInterval) of the whole
population through the C ABI (hb_run): K1 proposal -> K2 log-density plugin -> K3 accept/update [-> finalize].
Default workload (BASELINE C4, the configuration the north-star metric is quoted on, it fits one GPU): d = 100
t_distribution, nc = 2^20 chains, t = 1000, thin = 100; at N > 1 the population is sharded across ranks (strong
scaling) with a per-generation NCCL all-gather of the chain state.  Inputs (839 MB of chain state) are larger than
the 126 MB L2, so no explicit L2 flush is needed between steps.

`value`  : nc * generations / seconds, state resident in HBM, timed between two barrier+synchronize pairs, max over ranks.
`e2e`    : the same metric through the public API call dream_parallel() with HOST buffers: x0 upload, the full run,
           and the device->host fetch of chain / AR / CR inside the timed region.
`--impl reference`: the CPU oracle (restated reference, NOT R: R is not installed) on all host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENS_PER_STEP = 10
METRIC = "chain-generations/sec"
UNIT = "chain-generations/s"


def workload(name):
    import helpers as H
    from hydrobayes_b200 import _abi as A
    if name == "C4":
        nc, d, t = 1 << 20, 100, 1000
        cfg = A.default_config(nc=nc, t=t, d=d, lik=2, thin=100, seed=1312)
        return dict(cfg=cfg, target="t_distribution", x0=lambda: H.x0_tdist(nc, d), kw={},
                    desc="C4: dream_parallel t_distribution d=100 nc=2^20 t=1000 thin=100")
    if name == "C2":
        nc, d, t = 1024, 100, 20000
        cfg = A.default_config(nc=nc, t=t, d=d, lik=2, thin=1, seed=1312)
        return dict(cfg=cfg, target="t_distribution", x0=lambda: H.x0_tdist(nc, d), kw={},
                    desc="C2: t_distribution d=100 nc=1024 t=20000 (synchronous sweep)")
    if name == "C3":
        nc, t = 65536, 500
        P, obs = H.hydro_data(T=3653)
        cfg = A.default_config(nc=nc, t=t, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_BOUND,
                               thin=10, seed=1312)
        return dict(cfg=cfg, target="HydroModel", x0=lambda: H.x0_hydro(nc),
                    kw=dict(par_min=[0.01], par_max=[2.0], forcing=P, obs=obs),
                    desc="C3: HydroModel (d=1, T=3653 daily forcing, GLUE lik=31) nc=65536")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") to stdout while the first
        # communicator is created, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return rank, world, local, torch


def barrier_sync(torch, world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(torch, world, v):
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_sampler(w, local, rank, world, torch):
    from hydrobayes_b200 import Sampler, lib
    cfg = w["cfg"]
    s = Sampler(cfg, local)
    kw = w["kw"]
    if "par_min" in kw:
        s.set_bounds(kw["par_min"], kw["par_max"])
    if "obs" in kw:
        s.set_obs(kw["obs"])
    s.set_target(w["target"], kw.get("forcing"))
    if world > 1 and cfg.exchange == 2:      # HB_EXCHANGE_DELTA: no NCCL communicator inside the library
        s.comm_init(rank, world, None)
    elif world > 1:
        import ctypes as C
        import torch.distributed as dist
        buf = C.create_string_buffer(128)
        if rank == 0:
            rc = lib().hb_comm_unique_id(buf)
            assert rc == 0, lib().hb_last_error(None)
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        s.comm_init(rank, world, obj[0])
    s.set_initial(w["x0"]())
    if world > 1 and cfg.exchange in (1, 2):
        s.peer_setup(world)
    return s


def sharded_check(local, rank, world, torch, nc=1 << 16, d=100, gens=30):
    """N > 1, before the timed region: `gens` generations of a 2^16-chain population (adaptation statistics, outlier check
    and the chain-state exchange all active) run SHARDED over the ranks and, on every rank's own GPU, alone; the shard must
    equal the same chains of the single-GPU run bit for bit (chain, accept/reject record, final state), in every
    exchange mode.  A mismatch fails the whole bench run."""
    import helpers as H
    import torch.distributed as dist
    from hydrobayes_b200 import Sampler, _abi as A, lib
    x0 = H.x0_tdist(nc, d); x0[5] = 300.0; x0[nc - 3] = -250.0        # outlier resets across shard borders
    base = A.default_config(nc=nc, t=gens + 1, d=d, lik=2, seed=77, adapt=0.7, thin=5, record_accept=1)
    nl = nc // world

    def run_one(cfg, shard):
        with Sampler(cfg, local) as s:
            s.set_target("t_distribution")
            if shard:
                uid = None
                if cfg.exchange != A.HB_EXCHANGE_DELTA:
                    import ctypes as C
                    buf = C.create_string_buffer(128)
                    if rank == 0:
                        assert lib().hb_comm_unique_id(buf) == 0
                    obj = [bytes(buf.raw)]
                    dist.broadcast_object_list(obj, src=0)
                    uid = obj[0]
                s.comm_init(rank, world, uid)
            s.set_initial(x0)
            if shard and cfg.exchange in (A.HB_EXCHANGE_PEER, A.HB_EXCHANGE_DELTA):
                s.peer_setup(world)
            s.run(gens)
            n = nl if shard else None
            chain = s.fetch_chain(n); acc = s.fetch_accept(n)
            xt, _, _, lpost = s.fetch_state(n)
            return chain, acc, xt, lpost, s.fetch_outliers()

    full = run_one(H.copy_cfg(base), False)
    lo = rank * nl
    modes = {"delta": A.HB_EXCHANGE_DELTA, "allgather": A.HB_EXCHANGE_ALLGATHER}
    res = {}
    for name, mode in modes.items():
        sh = run_one(H.copy_cfg(base, exchange=mode), True)
        same = (np.array_equal(sh[0], full[0][:, :, lo:lo + nl], equal_nan=True) and np.array_equal(sh[1], full[1][:, lo:lo + nl])
                and np.array_equal(sh[2], full[2]) and np.array_equal(sh[3], full[3][lo:lo + nl]) and sh[4] == full[4])
        t = torch.tensor([1.0 if same else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res[name] = bool(t.item() == 1.0)
    return {"bit_identical": all(res.values()), "modes": sorted(res), "per_mode": res, "chains": nc, "d": d, "generations": gens,
            "outlier_resets": len(full[4]), "acceptance_rate": float(full[1].mean()),
            "what": "every rank compares its shard of the sharded run with the same chains of a single-GPU run on its own GPU"}


def roofline_for(kname, per_launch_ms, n_local, vt_work, peaks, dmma_peak, acc_rate=0.16, density_in_k1=False):
    """algorithmic bytes/flops per chain-generation x chains per launch / launch duration"""
    d = 100
    if kname == "K1":
        # gather of 1 + 2 D-bar rows + the proposal row written: 48 d + 16 bytes; with the log-density folded in (built-in
        # t_distribution, structured evaluation) 8 more for ll -- the 808 bytes per chain the plugin would re-read are not moved
        bytes_ = (48 * d + 16 + (8 if density_in_k1 else 0)) * n_local
        a = bytes_ / (per_launch_ms * 1e-3) / 1e9
        return {"kernel": "hb_k1_pair_kernel<2, true> (proposal + folded t-distribution log-density)" if density_in_k1 else "hb_k1_kernel (proposal)",
                "bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": a / peaks["hbm_gbs"], "peak_source": peaks["source"], "algorithmic_bytes_per_chain": bytes_ / n_local}
    if kname == "K3":
        # steady state (thin >> 1, outside the adaptation period): only accepted rows move -- read xp, write x -- plus
        # the per-chain scalars lp_xp, ll_raw, lpost, id and the accepted lp/ll/lpost stores
        bytes_ = (16 * d * acc_rate + 28 + 24 * acc_rate) * n_local
        a = bytes_ / (per_launch_ms * 1e-3) / 1e9
        return {"kernel": "hb_k3_kernel (accept)", "bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": a / peaks["hbm_gbs"], "peak_source": peaks["source"]}
    if os.environ.get("HB_TDIST_DENSE") != "1" and vt_work[0] == d * d + 3.0 * d:
        # the t_distribution plugin in its structured evaluation (HB_K1_DENSITY=0 runs it as its own launch): a stream over
        # the proposals, 8 d + 8 bytes per chain
        bytes_ = (8 * d + 8) * n_local
        a = bytes_ / (per_launch_ms * 1e-3) / 1e9
        return {"kernel": "hb_tdist_struct_kernel (log-density, structured evaluation)", "bound": "hbm", "achieved": a,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"], "peak_source": peaks["source"]}
    flops = vt_work[0] * n_local
    a = flops / (per_launch_ms * 1e-3) / 1e12
    return {"kernel": "K2 log-density plugin", "bound": "tensor", "achieved": a, "peak": dmma_peak, "unit": "TFLOP/s",
            "frac": a / dmma_peak if dmma_peak else None, "peak_source": "fp64 DMMA micro-benchmark hb_fp64_peaks() on this box"}


NVLINK_PEER_GBS = 770.0   # measured peer-copy bandwidth per direction on this pool (B200_PROFILING.md); nominal 900


def exchange_roofline(xchg_ms_per_launch, launches_per_generation, n_local, world, d, acc_rate):
    """HB_EXCHANGE_DELTA: every accepted row (ldx doubles + a 4-byte chain index) is stored once into each of the G - 1
    peers by the send kernels; `achieved` is that egress over the send kernels' own device time (they run on a side
    stream concurrently with the next piece's proposal / density / accept kernels)."""
    ldx = d if d == 1 else (d + 3) // 4 * 4
    bytes_gen = acc_rate * n_local * (8 * ldx + 4) * (world - 1)
    ms_gen = xchg_ms_per_launch * launches_per_generation
    a = bytes_gen / (ms_gen * 1e-3) / 1e9
    return {"kernel": "hb_delta_send_kernel (accepted rows -> every peer's arena, posted NVLink stores)", "bound": "nvlink",
            "achieved": a, "peak": NVLINK_PEER_GBS, "unit": "GB/s", "frac": a / NVLINK_PEER_GBS,
            "peak_source": "measured peer copy per direction, B200_PROFILING.md (nominal 900)",
            "bytes_pushed_per_generation": bytes_gen, "bytes_pushed_per_launch": bytes_gen / launches_per_generation,
            "send_launches_per_generation": launches_per_generation, "ms_per_generation_in_send_kernels": ms_gen,
            "ingress_bytes_per_generation": bytes_gen, "acceptance_rate": acc_rate}


def rscript_probe():
    """is the reference's own runtime on this box?"""
    import shutil
    p = shutil.which("Rscript")
    if not p:
        return "absent on GPU box (no Rscript on PATH): the reference cannot be executed, CPU baseline is the restated oracle"
    try:
        v = subprocess.run([p, "--version"], capture_output=True, text=True, timeout=20)
        return f"present: {p} ({(v.stdout + v.stderr).strip().splitlines()[0]})"
    except Exception as e:
        return f"present: {p} (version query failed: {e!r})"


def dgemm_cublas_tflops(torch, n=8192, reps=3):
    """cuBLAS DGEMM n^3 through torch.matmul (fp64): an independent anchor for the fp64 tensor-pipe denominator next to the
    library's own DMMA micro-benchmark.  Measurement only -- nothing on the product path calls cuBLAS."""
    try:
        a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


def cpu_baseline(w, seconds_budget=20.0, threads=None):
    """The oracle (restated reference, kind 'port') on the host cores, bounded sample of the same workload."""
    from oracle import hb_oracle_py as oracle
    import helpers as H
    from hydrobayes_b200 import _abi as A
    cfg = w["cfg"]
    threads = threads or os.cpu_count() or 1
    nc_s = min(cfg.nc, 8192 if cfg.d > 1 else 4096)
    x0 = w["x0"]()[:nc_s]
    gens = GENS_PER_STEP
    c = H.copy_cfg(cfg, nc=nc_s, t=gens + 1, thin=1, rng_mode=A.HB_RNG_PHILOX, n_runs=1)
    kw = dict(w["kw"])
    t0 = time.time(); done = 0; loop_s = 0.0; reps = 0
    while time.time() - t0 < seconds_budget and reps < 50:
        r = oracle.run(c, w["target"], x0, threads=threads, outputs=False, **kw)
        assert r["rc"] == 0, r["err"]
        loop_s += r["loop_seconds"]; done += nc_s * gens; reps += 1
    out = {"value": done / loop_s, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{reps} x ({nc_s} chains x {gens} generations) of the same workload, restated CPU oracle (C, OpenMP over "
                     f"chains), not R (R is not installed)"}
    if threads > 1:   # the reference's generation loop is single-threaded -- the same sample on ONE core beside it
        t0 = time.time(); done1 = 0; loop1 = 0.0; reps1 = 0
        while time.time() - t0 < seconds_budget / 4 and reps1 < 10:
            r = oracle.run(c, w["target"], x0, threads=1, outputs=False, **kw)
            assert r["rc"] == 0, r["err"]
            loop1 += r["loop_seconds"]; done1 += nc_s * gens; reps1 += 1
        out["one_core"] = {"value": done1 / loop1, "unit": UNIT, "cores": 1, "sample": f"{reps1} x the same sample"}
    return out, nc_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.workload)
    from oracle import hb_oracle_py as oracle
    import helpers as H
    from hydrobayes_b200 import _abi as A
    cfg = w["cfg"]
    threads = os.cpu_count() or 1
    nc_s = min(cfg.nc, 8192 if cfg.d > 1 else 4096)
    x0 = w["x0"]()[:nc_s]
    c = H.copy_cfg(cfg, nc=nc_s, t=GENS_PER_STEP + 1, thin=1, rng_mode=A.HB_RNG_PHILOX, n_runs=1)
    for _ in range(args.warmup):
        oracle.run(c, w["target"], x0, threads=threads, outputs=False, **w["kw"])
    loop = 0.0
    for _ in range(args.steps):
        r = oracle.run(c, w["target"], x0, threads=threads, outputs=False, **w["kw"])
        assert r["rc"] == 0, r["err"]
        loop += r["loop_seconds"]
    value = nc_s * GENS_PER_STEP * args.steps / loop
    sample = (f"each step = {nc_s} chains x {GENS_PER_STEP} generations of the workload (bounded sample of nc={cfg.nc}); restated "
              f"CPU oracle (C + OpenMP over chains), not R: R is not installed and the reference has no native code")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": loop / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "generations_per_step": GENS_PER_STEP, "chains": nc_s, "d": cfg.d,
                   "chains_of_the_workload": cfg.nc, "same_config": nc_s == cfg.nc,
                   "note": "bounded sample: the CPU state fits its caches, which favours the CPU arm"},
        "rscript": rscript_probe(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def quick_rate(name, local, torch, gens=200, warm=30):
    """short side measurement of another BASELINE config on one GPU (reported under `other_workloads`): the rate inside
    the adaptation period (statistics + finalize every generation, outlier check every updateInterval) and the
    steady-state rate after it (90 % of a run with the reference's default adapt = 0.1)"""
    w = workload(name)
    cfg = w["cfg"]
    s = make_sampler(w, local, 0, 1, torch)

    def timed(n):
        torch.cuda.synchronize()
        ms0, l0 = s.timing()
        t0 = time.perf_counter()
        s.run(n)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms1, l1 = s.timing()
        return wall, ms1 - ms0, l1 - l0

    s.run(warm)
    adapt_end = int(cfg.adapt * cfg.t)
    n_adapt = max(0, min(gens, adapt_end - warm - 1))
    out = {"workload": w["desc"], "unit": UNIT}
    if n_adapt >= 20:
        wall, dev, _ = timed(n_adapt)
        out["adaptation_period"] = {"value": cfg.nc * n_adapt / wall, "generations": n_adapt, "us_per_generation": wall / n_adapt * 1e6}
    done = warm + n_adapt + 1
    if done <= adapt_end + 2:
        s.run(adapt_end + 2 - done)          # leave the adaptation period
    wall, dev, launches = timed(gens)
    out.update({"value": cfg.nc * gens / wall, "generations": gens, "us_per_generation": wall / gens * 1e6,
                "device_us_per_generation": dev / gens * 1e3, "gpu_launches": launches,
                "note": "steady state (after the adaptation period)"})
    s.close()
    return out


def convergence_c5(local, torch, n_runs=4096, nc=10, T=3653, t=3000, rank=0, world=1):
    """BASELINE config 5 on one GPU: n_runs independent HydroModel DREAM runs (distinct synthetic catchments, nc = 10
    chains each, dream() = sequential sweep), every run driven until R_stat < 1.2 held at 3 consecutive checks
    (metric 2 of SURVEY 8d: seconds to Gelman-Rubin R-hat < 1.2)."""
    import helpers as H
    from hydrobayes_b200 import Sampler, _abi as A
    # independent runs shard with no communication: rank r owns catchments [r*n/world, (r+1)*n/world); run_offset keys
    # the Philox streams by the global run index, so the result does not depend on the number of GPUs
    n_loc = n_runs // world
    first = rank * n_loc
    P, O = H.hydro_data_batch(T, n_loc, first=first)
    cfg = A.default_config(nc=nc, t=t, d=1, lik=31, glue_shape=50.0, prior=A.HB_PRIOR_UNIFORM, bound=A.HB_BOUND_BOUND,
                           seed=1312, sweep=A.HB_SWEEP_SEQUENTIAL, n_runs=n_loc, run_offset=first)
    x0 = np.concatenate([H.x0_hydro(nc) for _ in range(n_loc)])
    with Sampler(cfg, local) as s:
        s.set_bounds([0.01], [2.0])
        s.set_target_batched("HydroModel", P, O)
        s.set_initial(x0)
        barrier_sync(torch, world)
        t0 = time.perf_counter()
        res = s.run_to_convergence(threshold=1.2, check_every=10, hold=3)
        barrier_sync(torch, world)
        seconds = max_over_ranks(torch, world, time.perf_counter() - t0)
        ms, launches = s.timing()
    conv = res["converged_at"]
    n_conv = int((conv > 0).sum()); gens = res["generations"]
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([n_conv, gens], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        n_conv = int(tt[0].item()); gens_sum = tt[1].item()
    else:
        gens_sum = gens
    return {"workload": f"C5: {n_runs} independent HydroModel DREAM runs (dream(), nc={nc}, T={T}, lik=31), each to R-hat<1.2"
                        + (f", {n_loc} runs per GPU, no communication" if world > 1 else ""),
            "seconds_to_rhat_below_1.2": seconds, "all_converged": n_conv == n_loc * world,
            "runs_converged": n_conv, "generations_run": gens,
            "median_generations_to_converge": float(np.median(conv[conv > 0])) if (conv > 0).any() else None,
            "value": n_loc * nc * (gens_sum - world) / seconds, "unit": UNIT,
            "device_loop_seconds": ms * 1e-3, "gpu_launches": launches,
            "rule": "max_k R_stat_k < 1.2 at 3 consecutive checks, every 10 stored generations from ind_out > 50"}


def convergence_population(name, local, rank, world, torch, t_cap, thin, check_every):
    """BASELINE metric 2 for one population (C2 on one GPU, C4 on N GPUs): wall seconds of dream_parallel's generation loop
    until max_k R_stat_k < 1.2 held at 3 consecutive checks of the stored (thinned) chain (R/gelman_rubin.R:21-56 on
    chain[1:ind_out, 1:d, ], as R/dream_test.R:344), checked every `check_every` stored generations from ind_out > 50
    (R/dream.R:270).  On a sharded population R_stat is the collective hb_rstat (per-chain statistics of the shards,
    two small all-reduces): every rank takes the same decisions."""
    import helpers as H
    w = workload(name)
    cfg = H.copy_cfg(w["cfg"], t=t_cap, thin=thin)
    if world > 1:
        cfg.exchange = 2
    w = dict(w, cfg=cfg)
    s = make_sampler(w, local, rank, world, torch)
    try:
        barrier_sync(torch, world)
        t0 = time.perf_counter()
        res = s.run_to_convergence(threshold=1.2, check_every=check_every, hold=3)
        barrier_sync(torch, world)
        seconds = max_over_ranks(torch, world, time.perf_counter() - t0)
        ms, launches = s.timing()
    finally:
        s.close()
    conv = int(res["converged_at"][0])
    return {"workload": w["desc"] + f" (t <= {t_cap}, thin = {thin}" + (f", sharded over {world} GPUs, delta exchange)" if world > 1 else ")"),
            "seconds_to_rhat_below_1.2": seconds if conv > 0 else None, "converged": conv > 0, "converged_at_generation": conv or None,
            "generations_run": res["generations"], "seconds_run": seconds, "R_stat_max_last": None if res["R_stat_max"] is None else float(res["R_stat_max"][0]),
            "rstat_checks": res["checks"], "seconds_in_rstat_checks": res["check_seconds"], "device_loop_seconds": ms * 1e-3,
            "value": cfg.nc * (res["generations"] - 1) / seconds, "unit": UNIT, "gpu_launches": launches,
            "rule": f"max_k R_stat_k < 1.2 at 3 consecutive checks, every {check_every} stored generations (thin {thin}) from ind_out > 50"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=97)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / other workloads (profiling runs)")
    ap.add_argument("--exchange", default="delta", choices=["delta", "peer", "allgather"], help="multi-GPU chain-state exchange")
    ap.add_argument("--quick", action="store_true", help="keep the per-kernel roofline, skip e2e / cpu_baseline / others")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3 if not args.no_extras else args.warmup

    rank, world, local, torch = dist_setup(args.gpus)
    import hydrobayes_b200 as hb
    from hydrobayes_b200 import lib
    assert lib().hb_device_count() >= 1, "no CUDA device: there is no CPU fallback"
    w = workload(args.workload)
    cfg = w["cfg"]
    if world > 1 and args.exchange == "peer" and cfg.d > 16:
        cfg.exchange = 1   # HB_EXCHANGE_PEER
    if world > 1 and args.exchange == "delta":
        cfg.exchange = 2   # HB_EXCHANGE_DELTA
    steady_steps = 0 if args.no_extras else 20             # a second, steady-state window after the profiled generations
    need = (args.steps + args.warmup + (0 if args.no_extras else 2) + steady_steps) * GENS_PER_STEP + 1   # + the profiled generations
    if need > cfg.t:
        cfg.t = need
    sharded = None
    if world > 1 and not args.no_extras:
        sharded = sharded_check(local, rank, world, torch)
        if not sharded["bit_identical"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "sharded_check": sharded,
                                  "error": "a shard of the sharded run differs from the single-GPU run: no number is reported"}))
            sys.exit(1)
    s = make_sampler(w, local, rank, world, torch)
    n_local = cfg.nc // world

    for _ in range(args.warmup):
        s.run(GENS_PER_STEP)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier_sync(torch, world)
    ms0, l0 = s.timing()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.run(GENS_PER_STEP)
    barrier_sync(torch, world)
    wall = time.perf_counter() - t0
    ms1, l1 = s.timing()
    clk = clocks.stop() if rank == 0 else None
    wall = max_over_ranks(torch, world, wall)
    dev_ms = max_over_ranks(torch, world, ms1 - ms0)
    gens = args.steps * GENS_PER_STEP
    value = cfg.nc * gens / wall

    # per-kernel durations (CUDA events on the library's stream, after the timed region)
    roof = None; kernels = {}; acc_rate = 0.16
    if not args.no_extras:
        s.profile(True)
        s.run(2 * GENS_PER_STEP)
        s.profile(False)
        for k in ("K1", "K2", "K3", "FIN", "XCHG", "APPLY", "OUTLIER"):
            ms, n = s.kernel_ms(k)
            if n:
                kernels[k] = {"ms_per_launch": ms / n, "launches": n}
        # steady state: `value` above is timed right after the warm-up, i.e. partly inside the adaptation period (statistics
        # pass over every row, outlier checks, and twice the acceptance rate = twice the rows to exchange); this second
        # window lies behind it (and behind the profiled generations)
        steady = None
        adapt_end = int(cfg.adapt * cfg.t)
        done_gens = 1 + (args.warmup + args.steps + 2) * GENS_PER_STEP
        if done_gens > adapt_end + GENS_PER_STEP:
            barrier_sync(torch, world)
            t0s = time.perf_counter()
            for _ in range(steady_steps):
                s.run(GENS_PER_STEP)
            barrier_sync(torch, world)
            ws = max_over_ranks(torch, world, time.perf_counter() - t0s)
            steady = {"value": cfg.nc * steady_steps * GENS_PER_STEP / ws, "unit": UNIT, "ms_per_generation": ws / (steady_steps * GENS_PER_STEP) * 1e3,
                      "generations": f"{done_gens + 1}..{done_gens + steady_steps * GENS_PER_STEP} (the adaptation period ends at {adapt_end})"}
        try:   # acceptance rate of the last complete updateInterval (AR column 2, R/dream_parallel.R:431)
            ar, _ = s.fetch_ar_cr()
            col = ar[:, 1][np.isfinite(ar[:, 1])]
            acc_rate = float(col[-1]) if col.size else 0.16
        except Exception:
            acc_rate = 0.16
    s.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall / args.steps * 1e3, "device_ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "generations_per_step": GENS_PER_STEP, "chains": cfg.nc, "d": cfg.d,
                       "chains_per_gpu": n_local, "rng": "philox4x32-10", "l2": "inputs larger than L2 (no flush)"
                       if cfg.nc * cfg.d * 8 > 126e6 else "state is L2-resident by construction (latency-bound config)",
                       "exchange": ("none" if world == 1 else "delta (stage-and-apply): full replica per GPU; accepted rows go to a delta log, a send kernel on a side stream stores each piece's log into every peer's arena over NVLink while the next piece computes, an apply kernel folds all logs into the replica; adaptation sums all-reduced once per updateInterval" if cfg.exchange == 2 else "peer gather over NVLink (bulk TMA reads of partner rows) + 17-word all-reduce per generation" if cfg.exchange == 1 else "ncclAllGather per generation")},
            "gpu_launches": l1 - l0, "rscript": rscript_probe()}
    if sharded is not None:
        line["sharded_check"] = sharded
    if not args.no_extras and steady is not None:
        line["steady_state"] = steady
        line["timed_window"] = (f"generations {2 + args.warmup * GENS_PER_STEP}..{1 + (args.warmup + args.steps) * GENS_PER_STEP}; the adaptation "
                                f"period (statistics pass, outlier checks, higher acceptance rate) ends at generation {int(cfg.adapt * cfg.t)}")
    if rank == 0 and not args.no_extras:
        peaks = load_peaks()
        import ctypes as C
        dm, df = C.c_double(), C.c_double()
        lib().hb_fp64_peaks(C.byref(dm), C.byref(df))
        flops = {"t_distribution": (cfg.d * cfg.d + 3.0 * cfg.d,), "HydroModel": (5.0 * 3653,), "mixture": (40.0,)}[w["target"]]
        if kernels:
            dom = max(("K1", "K2", "K3"), key=lambda k: kernels.get(k, {"ms_per_launch": 0})["ms_per_launch"])
            pieces = max(1, round(kernels["K1"]["launches"] / (2 * GENS_PER_STEP))) if "K1" in kernels else 1
            chains_per_launch = n_local / pieces
            # no K2 launches on a t_distribution workload: the proposal kernel wrote the log-density itself
            folded = w["target"] == "t_distribution" and "K2" not in kernels
            roof = roofline_for(dom, kernels[dom]["ms_per_launch"], chains_per_launch, flops, peaks, dm.value, acc_rate, folded)
            # DRAM traffic of the dominant kernel is NOT measured inside this run (that needs ncu): at N = 1 the figure of
            # the committed ncu capture of the same kernel on the same workload is quoted and labelled; never at N > 1,
            # where a launch covers a different number of chains
            roof["traffic"] = None
            prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if world == 1 and args.workload == "C4" and os.path.exists(prof):
                pj = json.load(open(prof))
                roof["traffic"] = pj.get(dom)
                roof["traffic_source"] = f"not measured in this run: {pj.get('_source', 'profiles/ncu_traffic.json')} (N = 1, {cfg.nc} chains per launch)"
            roof["all_kernels"] = {k: roofline_for(k, kernels[k]["ms_per_launch"], chains_per_launch, flops, peaks, dm.value, acc_rate, folded)
                                   for k in ("K1", "K2", "K3") if k in kernels}
            if folded:
                roof["log_density"] = ("t_distribution in its structured evaluation (x' C^-1 x = 2 (sum u^2 - (sum u)^2 / (d + 1)), u_i = x_i / sqrt(i): "
                                       "4 d flops per chain), accumulated by the proposal kernel on the proposal it holds in registers; no K2 launch. "
                                       "HB_K1_DENSITY=0 runs it as a separate kernel, HB_TDIST_DENSE=1 restores dmvt's triangular product on the fp64 tensor cores")
            roof["kernel_ms_per_launch"] = {k: v["ms_per_launch"] for k, v in kernels.items()}
            roof["launches_per_generation"] = {k: v["launches"] / (2 * GENS_PER_STEP) for k, v in kernels.items()}
            roof["chains_per_launch"] = chains_per_launch
            dg = dgemm_cublas_tflops(torch)
            roof["fp64_peaks_measured_tflops"] = {"dmma": dm.value, "dfma": df.value, "dgemm_cublas_tflops": dg,
                                                  "note": "dmma / dfma: the library's micro-benchmarks; dgemm_cublas: torch.matmul fp64 8192^3 (measurement only)"}
            if dg and "K2" in roof["all_kernels"]:
                roof["all_kernels"]["K2"]["frac_of_cublas_dgemm"] = roof["all_kernels"]["K2"]["achieved"] / dg
            roof["acceptance_rate"] = acc_rate
            if w["target"] == "t_distribution" and world == 1:
                # the least a generation has to move (gather + accepted rows + stored rows), against the wall time of a generation
                single_pass = (8 * cfg.d * 5 + 8 * cfg.d * acc_rate + 8 * (cfg.d + 3) / max(1, cfg.thin)) * cfg.nc
                gen_s = (steady["ms_per_generation"] if steady else wall / gens * 1e3) * 1e-3
                roof["generation_vs_single_pass_bound"] = {"bytes": single_pass, "generation_ms": gen_s * 1e3,
                                                           "frac": single_pass / gen_s / 1e9 / peaks["hbm_gbs"]}
            if world > 1 and cfg.exchange == 2 and "XCHG" in kernels:
                roof["exchange"] = exchange_roofline(kernels["XCHG"]["ms_per_launch"], kernels["XCHG"]["launches"] / (2 * GENS_PER_STEP),
                                                     n_local, world, cfg.d, acc_rate)
                gen_ms = wall / gens * 1e3
                roof["exchange"]["frac_of_generation_time_if_not_overlapped"] = roof["exchange"]["ms_per_generation_in_send_kernels"] / gen_ms
                roof["generation_ms"] = gen_ms
        line["roofline"] = roof
        line["clocks"] = clk
    if not args.no_extras and not args.quick:
        # end to end through the public API with host buffers (full C4 run: x0 upload, t-1 generations, chain fetch);
        # at N > 1 every rank calls dream_parallel(distributed=True) and fetches the chain of its own shard
        w2 = workload(args.workload)
        c2 = w2["cfg"]
        x0 = w2["x0"]()
        par = dict(initial="user", val_ini=x0, prior={0: "flat", 1: "uniform"}[c2.prior])
        if "par_min" in w2["kw"]:
            par.update(min=w2["kw"]["par_min"], max=w2["kw"]["par_max"], bound="bound")
        # The call is host-bound outside its device loop (9.5 GB of result pages to fault and fill, x0 through a pinned ring)
        # and varies from call to call on a shared box (1.6 .. 3 s measured for the same code): three complete calls, every
        # one reported, the MEDIAN is the value.
        reps = []
        d2h = 0
        for _ in range(3):
            lib().hb_wait_idle()   # the teardown of the previous sampler (deferred frees) is not part of the call timed below
            barrier_sync(torch, world)
            t0 = time.perf_counter()
            res = hb.dream_parallel(w2["target"], lik=c2.lik, par_info=par, nc=c2.nc, t=c2.t, d=c2.d, thin=c2.thin,
                                    glue_shape=c2.glue_shape, obs=w2["kw"].get("obs"), forcing=w2["kw"].get("forcing"),
                                    verbose=False, seed=1312, store_fx=False, device=local, distributed=world > 1,
                                    exchange=cfg.exchange if world > 1 else None)
            barrier_sync(torch, world)
            reps.append(max_over_ranks(torch, world, time.perf_counter() - t0))
            d2h = res["chain"].nbytes + res["AR"].nbytes + res["CR"].nbytes
            del res
        e2e_s = sorted(reps)[1]
        n_steps = (c2.t - 1) / GENS_PER_STEP
        if rank == 0:
            line["e2e"] = {"value": c2.nc * (c2.t - 1) / e2e_s, "unit": UNIT,
                           "h2d_bytes_per_step": world * x0.nbytes / n_steps,
                           "d2h_bytes_per_step": world * d2h / n_steps,
                           "seconds": e2e_s, "seconds_of_every_call": reps, "statistic": "median of 3 complete calls",
                           "call": "hydrobayes_b200.dream_parallel(...) full run incl. x0 upload and chain fetch"
                                   + (" (distributed=True: every rank uploads x0 and fetches the chain of its shard)" if world > 1 else "")}
    if world > 1 and not args.no_extras and not args.quick:
        try:
            c5 = convergence_c5(local, torch, rank=rank, world=world)
        except Exception as e:  # report, never hide
            c5 = {"error": repr(e)}
        try:   # BASELINE metric 2 on the sharded C4 population (collective R_stat)
            c4m2 = convergence_population("C4", local, rank, world, torch, t_cap=8001, thin=100, check_every=5)
        except Exception as e:
            c4m2 = {"error": repr(e)}
        if rank == 0:
            line["other_workloads"] = {"C5": c5, "C4_seconds_to_rhat": c4m2}
    if rank == 0 and world == 1 and not args.no_extras and not args.quick:
        cb, _ = cpu_baseline(w)
        line["cpu_baseline"] = cb
        others = {}
        for name in ("C2", "C3"):
            if name != args.workload:
                try:
                    others[name] = quick_rate(name, local, torch)
                except Exception as e:  # report, never hide
                    others[name] = {"error": repr(e)}
        if "C2" in others and "error" not in others["C2"]:
            # the default C2 path is the persistent generation loop (one cooperative launch per slice); next to it the
            # one-launch-per-generation fused kernel and the three-kernel path; a failure is reported, never hidden
            for label, env in (("fused_one_launch_per_generation", {"HB_PERSIST": "0"}), ("three_kernel_path", {"HB_FUSED": "0"})):
                os.environ.update(env)
                try:
                    f = quick_rate("C2", local, torch)
                    others["C2"][label] = {k: f[k] for k in ("value", "us_per_generation", "device_us_per_generation", "gpu_launches", "generations")}
                    others["C2"][label]["enable"] = " ".join(f"{k}={v}" for k, v in env.items())
                except Exception as e:
                    others["C2"][label] = {"error": repr(e)}
                finally:
                    for k in env:
                        os.environ.pop(k, None)
        if args.workload == "C4":
            # The same workload with the target evaluated by dmvt's own route (chol(C)^-1, triangular product on the fp64 tensor
            # cores, its own launch) and with the structured evaluation as a separate launch: what the default path -- the
            # structured form folded into the proposal kernel -- is to be read against.
            for label, env, what in (("C4_dense_log_density", {"HB_TDIST_DENSE": "1"}, "K1 -> hb_tdist_dmma_kernel -> K3"),
                                     ("C4_structured_log_density_own_launch", {"HB_K1_DENSITY": "0"}, "K1 -> hb_tdist_struct_kernel -> K3")):
                os.environ.update(env)
                try:
                    f = quick_rate("C4", local, torch, gens=60, warm=110)
                    others[label] = {"path": what, "enable": " ".join(f"{k}={v}" for k, v in env.items()),
                                     **{k: f[k] for k in ("value", "us_per_generation", "device_us_per_generation", "gpu_launches", "generations", "note")}}
                except Exception as e:
                    others[label] = {"error": repr(e)}
                finally:
                    for k in env:
                        os.environ.pop(k, None)
        try:
            others["C5"] = convergence_c5(local, torch)
        except Exception as e:
            others["C5"] = {"error": repr(e)}
        for key, args_m2 in (("C2_seconds_to_rhat", dict(name="C2", t_cap=20000, thin=10, check_every=10)),
                             ("C4_seconds_to_rhat", dict(name="C4", t_cap=8001, thin=100, check_every=5))):
            try:   # BASELINE metric 2 (seconds to R-hat < 1.2) for the single-population configs
                lib().hb_wait_idle()
                others[key] = convergence_population(local=local, rank=0, world=1, torch=torch, **args_m2)
            except Exception as e:
                others[key] = {"error": repr(e)}
        line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
