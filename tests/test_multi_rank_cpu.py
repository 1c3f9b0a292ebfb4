"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: block partition, communicator-id
broadcast, shard gather into R's chain[out_t, d+3, nc] layout.  The device side of the N > 1 path is checked on the GPU
box by tools/multi_gpu_check.py (sharded run == single-GPU run, bit for bit)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hydrobayes_b200.sampler import broadcast_unique_id, gather_chain, shard_range
    nc, out_t, d = 12, 5, 3
    c0, nl = shard_range(nc, rank, world)
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
    full = np.arange(out_t * (d + 3) * nc, dtype=np.float64).reshape(out_t, d + 3, nc)
    shard = np.asfortranarray(full[:, :, c0:c0 + nl])
    got = gather_chain(shard, world)
    q.put((rank, c0, nl, uid == bytes(range(128)), np.array_equal(got, full), got.flags["F_CONTIGUOUS"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plumbing_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[0], r[1], r[2]) for r in res] == [(0, 0, 6), (1, 6, 6)]
    assert all(r[3] and r[4] and r[5] for r in res)


def test_shard_range_errors():
    from hydrobayes_b200.sampler import shard_range
    assert shard_range(1 << 20, 7, 8) == (7 * (1 << 17), 1 << 17)
    with pytest.raises(ValueError):
        shard_range(10, 0, 3)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
