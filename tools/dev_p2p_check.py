"""Multi-GPU check (run under torchrun with >= 2 GPUs): the fused peer-memory all-gather delivers, on every rank, exactly
the joint angles every rank computed (compared with a plain torch.distributed all_gather of the local results), for every
buffer-mapping path the box offers.  Usage: dev_p2p_check.py [robots per rank] [paths, comma separated: auto,symm-unicast,ipc]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.parallel import shard_robots
from syropod_highlevel_controller_b200.streams import CommandStream

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 + 17
paths = (sys.argv[2] if len(sys.argv) > 2 else "auto,symm-unicast,ipc").split(",")
cfg = hexapod_config("tripod_gait")
shard = shard_robots(n * world, rank, world)
all_ok = True
for path in paths:
    cs = CommandStream(n, robot_offset=shard.offset, min_len=10, max_len=30)
    K = 48  # three times around the 16-buffer ring
    cmds = torch.from_numpy(np.stack([cs.next() for _ in range(K)])).to(dev)
    ref = Engine(cfg, n, device=local)   # same shard, plain steps
    eng = Engine(cfg, n, device=local)
    try:
        g = eng.init_gather_fused(rank, world, mode=path)
    except RuntimeError as ex:
        if rank == 0:
            print(f"path {path}: unavailable ({ex})", flush=True)
        eng.close(); ref.close()
        continue
    worst = 0.0
    for k0 in range(0, K, 8):
        for k in range(k0, k0 + 8):
            eng.gather_step(cmds[k])
            j = ref.step(cmds[k])
        b = eng.gather_sync()
        torch.cuda.synchronize()
        eng.gather_status()
        parts = [torch.empty_like(j) for _ in range(world)]
        dist.all_gather(parts, j.contiguous())
        want = torch.stack(parts)  # [world, n, L, D]
        got = g[b]
        worst = max(worst, float((got - want).abs().max()))
        dist.barrier()
    ok = worst == 0.0
    all_ok = all_ok and ok
    print(f"rank {rank}: path {path} [{eng.gather_mode}]: fused gather vs all_gather of per-rank results: max |diff| = {worst} -> "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    eng.close(); ref.close()
    dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)
