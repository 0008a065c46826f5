"""Data-parallel plumbing over the robot batch: robots are independent, so the batch is cut into one
contiguous block per rank with no data-path collective; the only exchange is one all-gather of the joint angles per
control cycle (BASELINE.json configs[4]).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    n_global: int
    offset: int  # first global robot id of this rank
    count: int   # robots on this rank


def shard_robots(n_global: int, rank: int, world_size: int) -> Shard:
    """Contiguous block partition; the first n_global % world_size ranks hold one extra robot."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_global, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return Shard(rank, world_size, n_global, offset, count)


def all_shards(n_global: int, world_size: int) -> List[Shard]:
    return [shard_robots(n_global, r, world_size) for r in range(world_size)]


class JointGather:
    """Per-cycle all-gather of the joint-angle slices, double buffered on a side stream so that the gather of cycle t
    overlaps the kernel of cycle t+1 (nothing in cycle t+1 depends on the gathered result)."""

    def __init__(self, shard: Shard, legs: int, dof: int, device, group=None, buffers: int = 2):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.shard, self.group = shard, group
        if shard.n_global % shard.world_size:
            raise ValueError("all_gather_into_tensor needs equal shards; pad the batch to a multiple of the world size")
        self.local = [torch.empty((shard.count, legs, dof), dtype=torch.float32, device=device) for _ in range(buffers)]
        self.full = [torch.empty((shard.n_global, legs, dof), dtype=torch.float32, device=device) for _ in range(buffers)]
        self.cuda = torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [None] * buffers
        self.cycle = 0

    def next_local_buffer(self):
        """Output buffer for this cycle's kernel; waits (on the current stream) until its previous gather has drained."""
        b = self.cycle % len(self.local)
        if self.cuda and self.done[b] is not None:
            self.torch.cuda.current_stream().wait_event(self.done[b])
        return self.local[b]

    def gather(self):
        """Enqueue the all-gather of the buffer just written; returns the full [n_global, L, D] tensor it lands in."""
        torch, dist = self.torch, self.dist
        b = self.cycle % len(self.local)
        self.cycle += 1
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                dist.all_gather_into_tensor(self.full[b], self.local[b], group=self.group)
                self.done[b] = torch.cuda.Event()
                self.done[b].record()
        else:
            parts = [torch.empty_like(self.local[b]) for _ in range(self.shard.world_size)]
            dist.all_gather(parts, self.local[b], group=self.group)
            self.full[b].copy_(torch.cat(parts, dim=0))
        return self.full[b]

    def wait(self):
        if self.cuda:
            self.torch.cuda.current_stream().wait_stream(self.side)
