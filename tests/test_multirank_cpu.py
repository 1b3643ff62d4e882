"""world_size-2 gloo test of the N>1 host logic of bench.py: independent chains per rank (no data-path
collective), barrier + max-over-ranks timing, whole-job value = all ranks' steps / max time."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    steps = 3
    ms_local = torch.tensor([100.0 * (rank + 1)], dtype=torch.float64)  # rank 1 is the slow one
    ms = bench.reduce_max_ms(ms_local, world)
    val = bench.whole_job_value(world, steps, ms)
    seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(seeds, torch.tensor([bench.chain_seed(rank)], dtype=torch.int64))
    if rank == 0:
        out.put((ms, val, [int(s.item()) for s in seeds]))
    dist.destroy_process_group()


def test_world_size_2_timing_and_seeds():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, val, seeds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 200.0                       # max over ranks
    assert abs(val - 2 * 3 / 0.2) < 1e-9     # all ranks' steps / max time
    assert len(set(seeds)) == 2              # different chains per rank
