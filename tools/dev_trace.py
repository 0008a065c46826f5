"""Kernel-tuning tool (run under gpurun with a -DSHC_TRACE build: tools/dev_build_variants.sh "trace:-DSHC_TRACE", then
SHC_B200_LIB=$PWD/gpurun_in_lib_trace.so python tools/dev_trace.py): per-tile timeline of one control-cycle launch in
the bench workload's steady state, from globaltimer stamps written by lane 0 of every warp.  Prints per-stage
durations (ns; median / p10 / p90 over tiles), split into the first wave of warps and the later ones."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200 import engine as E
from syropod_highlevel_controller_b200.streams import CommandStream

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
cfg = hexapod_config("tripod_gait")
L = cfg.leg_count
eng = E.Engine(cfg, n, precision=prec)
cs = CommandStream(n)
pre = 300
cmds = torch.from_numpy(np.stack([cs.next() for _ in range(pre + 4)])).cuda()
for i in range(pre):
    eng.step(cmds[i])
torch.cuda.synchronize()
tiles = (n + 31) // 32
buf = torch.zeros((tiles, 32), dtype=torch.int64, device="cuda")
lib = E.lib()
have_trace = hasattr(lib, "shc_debug_trace")
if have_trace:
    lib.shc_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.shc_debug_trace(eng._h, ctypes.c_void_p(buf.data_ptr()))
eng.step(cmds[pre]); eng.step(cmds[pre + 1])
buf.zero_()
eng.step(cmds[pre + 2])
torch.cuda.synchronize()
snap = buf.cpu().numpy().astype(np.int64)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(20):
    eng.step(cmds[pre + (i & 3)])
ev1.record()
torch.cuda.synchronize()
print("event-timed us/step over 20 steps:", ev0.elapsed_time(ev1) / 20 * 1e3)
if not have_trace:
    sys.exit("not a -DSHC_TRACE build")
t = snap
t0 = t[:, 0].min()
t = np.where(t > 0, t - t0, -1)
END = 31
def stat(name, a):
    a = a.astype(float)
    print(f"{name:34s} med {np.median(a):9.0f}  p10 {np.percentile(a,10):9.0f}  p90 {np.percentile(a,90):9.0f}  max {a.max():9.0f}")
print("kernel span (ns):", t[:, END].max())
first = t[:, 0] < np.median(t[:, END] - t[:, 0]) * 0.5
print("tiles in the first wave:", int(first.sum()), "of", tiles)
for name, sel in (("first wave", first), ("later", ~first)):
    if sel.sum() == 0:
        continue
    stat(f"[{name}] robot-level loads (0->2)", t[sel, 2] - t[sel, 0])
    stat(f"[{name}] robot-level stage (2->3)", t[sel, 3] - t[sel, 2])
    for l in range(L):
        stat(f"[{name}] leg {l} slot wait", t[sel, 5 + 3 * l] - t[sel, 4 + 3 * l])
    stat(f"[{name}] leg stepper (wait->release)", np.concatenate([t[sel, 6 + 3 * l] - t[sel, 5 + 3 * l] for l in range(L)]))
    stat(f"[{name}] leg ik (release->next)", np.concatenate([t[sel, 4 + 3 * (l + 1)] - t[sel, 6 + 3 * l] for l in range(L - 1)]))
    stat(f"[{name}] legs total (3->30)", t[sel, 30] - t[sel, 3])
    stat(f"[{name}] output (30->31)", t[sel, END] - t[sel, 30])
    stat(f"[{name}] tile total", t[sel, END] - t[sel, 0])
starts, ends = np.sort(t[:, 0]), np.sort(t[:, END])
for q in (0.1, 0.25, 0.5, 0.75, 0.9, 1.0):
    tt = t[:, END].max() * q
    print(f"t={tt:8.0f} ns: started {np.searchsorted(starts, tt, 'right'):5d} finished {np.searchsorted(ends, tt, 'right'):5d}")
