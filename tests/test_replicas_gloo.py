"""N>1 host logic on CPU: world_size-2 gloo run of the replica plumbing bench.py uses."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xivo_b200 import replicas


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = replicas.max_over_ranks(10.0 + 5.0 * rank)
    streams = replicas.assign_streams(rank, world, 4, 8)
    ids = replicas.global_sequence_ids(rank, world, 4)
    dist.barrier()
    q.put((rank, ms, streams, ids))
    dist.destroy_process_group()


def test_two_rank_gloo_replicas():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(abs(r[1] - 15.0) < 1e-12 for r in res)  # max over ranks, identical on every rank
    assert res[0][2] == [0, 1, 2, 3] and res[1][2] == [4, 5, 6, 7]  # disjoint streams
    assert sorted(res[0][3] + res[1][3]) == list(range(8))  # every sequence owned exactly once
    assert replicas.aggregate_frames_per_second(4 * 10, 2, 15.0) == 2 * 40 / 0.015
