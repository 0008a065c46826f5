"""Multi-GPU check (run under torchrun with >= 2 GPUs): the fused peer-memory all-gather delivers, on every rank, exactly
the joint angles every rank computed (compared with a plain torch.distributed all_gather of the local results)."""
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
cfg = hexapod_config("tripod_gait")
shard = shard_robots(n * world, rank, world)
cs = CommandStream(n, robot_offset=shard.offset, min_len=10, max_len=30)
K = 40
cmds = torch.from_numpy(np.stack([cs.next() for _ in range(K)])).to(dev)
# reference engine: same shard, plain steps
ref = Engine(cfg, n, device=local)
eng = Engine(cfg, n, device=local)
eng.init_nccl(rank, world)
g = eng.init_gather_fused(rank, world)
worst = 0.0
for k0 in range(0, K, 8):
    b = eng.rollout_gather_fused(cmds[k0:k0 + 8])
    for k in range(k0, k0 + 8):
        j = ref.step(cmds[k])
    torch.cuda.synchronize()
    parts = [torch.empty_like(j) for _ in range(world)]
    dist.all_gather(parts, j.contiguous())
    want = torch.stack(parts)  # [world, n, L, D]
    got = g[b]
    worst = max(worst, float((got - want).abs().max()))
    dist.barrier()
ok = worst == 0.0
print(f"rank {rank}: fused gather vs all_gather of per-rank results: max |diff| = {worst} -> {'OK' if ok else 'MISMATCH'}", flush=True)
eng.close(); ref.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
