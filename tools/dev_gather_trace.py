"""Tuning (torchrun, >= 2 GPUs, SHC_GATHER_TRACE=1): timeline of the fused gather protocol on rank 0 — per cycle the end of
the control-cycle kernel, start and end of its landed-signal kernel; per device-side wait its start and end."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SHC_GATHER_TRACE"] = "1"
import torch
import torch.distributed as dist
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine, lib
from syropod_highlevel_controller_b200.streams import CommandStream

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 131072
path = sys.argv[1] if len(sys.argv) > 1 else "auto"
calls = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "40,20,20").split(",")]
eng = Engine(hexapod_config("tripod_gait"), n, device=local, precision="f64")
eng.init_gather_fused(rank, world, mode=path)
cs = CommandStream(n, robot_offset=rank * n)
cmds = torch.from_numpy(np.stack([cs.next() for _ in range(sum(calls))])).to(dev)
k0 = 0
for k in calls:
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.rollout_gather_fused(cmds[k0:k0 + k])
    e1.record()
    torch.cuda.synchronize()
    print(f"rank {rank}: call of {k} cycles: {e0.elapsed_time(e1) * 1e3:.1f} us = {e0.elapsed_time(e1) * 1e3 / k:.1f} us/cycle", flush=True)
    k0 += k
torch.cuda.synchronize()
L = lib()
L.shc_gather_trace.restype = C.POINTER(C.c_uint64)
L.shc_gather_trace.argtypes = [C.c_void_p]
tr = np.ctypeslib.as_array(L.shc_gather_trace(eng._h), shape=(4096, 4)).copy()
if rank == 0:
    t0 = tr[0, 0]
    print(f"path {path} [{eng.gather_mode}] calls {calls}; microseconds since the end of cycle 0's kernel")
    print("cycle  kernel_end  signal_start  signal_end  (signal lag behind its kernel)")
    for c in range(sum(calls)):
        ke, ss, se = (tr[c, :3].astype(np.int64) - int(t0)) / 1e3
        print(f"{c:4d} {ke:10.1f} {ss:12.1f} {se:11.1f}   {se - ke:8.1f}")
    print("wait  start  end  duration  need")
    for w in range(64):
        s, e_, need = tr[2048 + w, :3]
        if s == 0:
            break
        print(f"{w:3d} {(int(s) - int(t0)) / 1e3:10.1f} {(int(e_) - int(t0)) / 1e3:10.1f} {(int(e_) - int(s)) / 1e3:8.1f} {int(need)}")
eng.close()
dist.destroy_process_group()
