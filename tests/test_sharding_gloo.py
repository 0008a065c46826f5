"""CPU: the N>1 path's host logic with the gloo backend, world_size 2 (no GPU needed): block partitioning, shard-
independent input streams, and the layout of the per-cycle joint all-gather."""
import os
import socket

import numpy as np
import pytest

from syropod_highlevel_controller_b200.parallel import all_shards, shard_robots


def test_block_partition_covers_batch_exactly():
    for n, w in ((1048576, 8), (1000, 3), (7, 8), (4096, 1)):
        shards = all_shards(n, w)
        assert shards[0].offset == 0 and sum(s.count for s in shards) == n
        for a, b in zip(shards, shards[1:]):
            assert a.offset + a.count == b.offset
    s = shard_robots(1048576, 5, 8)
    assert (s.offset, s.count) == (5 * 131072, 131072)  # BASELINE configs[4]: 131072 robots per GPU
    with pytest.raises(ValueError):
        shard_robots(10, 3, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, cycles, q):
    import torch
    import torch.distributed as dist
    from syropod_highlevel_controller_b200.parallel import JointGather, shard_robots
    from syropod_highlevel_controller_b200.streams import CommandStream

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, D = 6, 3
        sh = shard_robots(n_global, rank, world)
        cs = CommandStream(sh.count, robot_offset=sh.offset)
        gather = JointGather(sh, L, D, "cpu")
        ok = True
        for c in range(cycles):
            cmd = torch.from_numpy(cs.next())
            out = gather.next_local_buffer()
            # stand-in for the kernel: a deterministic function of (global robot id, command, cycle)
            ids = torch.arange(sh.offset, sh.offset + sh.count, dtype=torch.float32)
            out.copy_((ids[:, None, None] * 1e-3 + cmd[:, :1, None] + c) * torch.ones(1, L, D))
            full = gather.gather()
            # the reference stream for ALL robots, generated without sharding
            if c == 0:
                ref_cs = CommandStream(n_global)
            ref_cmd = torch.from_numpy(ref_cs.next())
            ref_ids = torch.arange(n_global, dtype=torch.float32)
            ref = (ref_ids[:, None, None] * 1e-3 + ref_cmd[:, :1, None] + c) * torch.ones(1, L, D)
            ok = ok and torch.equal(full, ref)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_all_gather_layout_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 64, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
