"""N>1 path on CPU: world_size-2 gloo run of the frame-index scatter + sequence merge (rendezvous on 127.0.0.1)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ultragrid_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, frames, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    for step in range(steps):
        out = torch.zeros(frames, dtype=torch.int32)
        sharding.scatter_assignment(step, frames, out)
        got.append(out.tolist())
    gathered = [None] * world
    dist.all_gather_object(gathered, [(i, f"frame{i}@rank{rank}") for s in got for i in s])
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_frame_assignment_partitions_the_queue():
    a = sharding.frame_assignment(3, 4, 5)
    assert a.shape == (4, 5) and a.flatten().tolist() == list(range(60, 80))


def test_scatter_and_merge_world_size_2():
    world, frames, steps = 2, 4, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per_rank = [sorted(g) for g in gathered]
    assert [i for i, _ in per_rank[0]][:4] == [0, 1, 2, 3] and [i for i, _ in per_rank[1]][:4] == [4, 5, 6, 7]
    merged = sharding.merge_in_sequence(per_rank)
    assert len(merged) == world * frames * steps
    assert merged[5] == "frame5@rank1" and merged[8] == "frame8@rank0"


def test_single_process_fallback():
    out = torch.zeros(6, dtype=torch.int32)
    sharding.scatter_assignment(2, 6, out)
    assert out.tolist() == list(range(12, 18))
