"""Kernel-tuning tool: end-to-end step time through shc_step_host with page-locked buffers (zero-copy output vs the
tile-range / copy-engine path is chosen by SHC_HOST_ZEROCOPY in the environment)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.streams import CommandStream

n = 131072
cfg = hexapod_config("tripod_gait")
eng = Engine(cfg, n)
ref = Engine(cfg, n, startup=eng.startup())
cs = CommandStream(n)
cmd = eng.pinned_host(n, 3); out = eng.pinned_host(n, 6, 3)
for i in range(40):
    c = cs.next()
    cmd[:] = c
    j = eng.step_host(cmd, out=out)
    jr = ref.step(torch.from_numpy(c).cuda()).cpu().numpy()
    assert np.array_equal(j, jr), i
t0 = time.perf_counter()
for i in range(30):
    eng.step_host(cmd, out=out)
dt = (time.perf_counter() - t0) / 30
print(f"SHC_HOST_ZEROCOPY={os.environ.get('SHC_HOST_ZEROCOPY', '(default)')}: bit-exact over 40 steps; {dt*1e6:.1f} us/step, {n/dt:.4g} steps/s end to end")
