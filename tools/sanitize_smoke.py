"""Small end-to-end pass over every kernel of the library, meant to run under `compute-sanitizer --tool memcheck` (and
`--tool racecheck`) on a GPU box: control cycle in all three kernel modes and both precisions, host-buffer step, status
flags, sequences, start-up, workspace sweep, message packing, stand-alone IK.  Prints `sanitize smoke ok`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config  # noqa: E402
from syropod_highlevel_controller_b200.engine import Engine  # noqa: E402
from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream  # noqa: E402

n, cycles = 70, 12  # not a multiple of 32: the tail tile is exercised
dev = torch.device("cuda:0")
configs = [(hexapod_config("tripod_gait"), False), (hexapod_config("wave_gait", auto_posing=1), False),
           (octopod_config("tripod_gait"), True), (octopod_config("ripple_gait", use_joint_effort=1), True),
           (octopod_config("tripod_gait", gravity_aligned_tips=1), True), (hexapod_config("tripod_gait", gravity_aligned_tips=1), False),
           (hexapod_config("amble_gait", rough_terrain_mode=1, step_depth=0.01), False)]
for cfg, sensors in configs:
    for precision in ("f64", "mixed"):
        L, D = cfg.leg_count, cfg.joint_count
        eng = Engine(cfg, n, precision=precision)
        eng.set_options(1)  # status flags
        cs, ims, fs = CommandStream(n), ImuStream(n), ForceStream(n, L)
        if cfg.use_joint_effort:
            eng.set_joint_efforts(torch.randn(n, L, D, device=dev))
        for c in range(cycles):
            cmd = torch.from_numpy(cs.next()).to(dev)
            imu = torch.from_numpy(ims.next(cfg.time_delta)).to(dev) if sensors else None
            force = torch.from_numpy(fs.next()).to(dev) if (sensors or cfg.rough_terrain_mode) else None
            eng.step(cmd, imu, force)
        eng.step_host(cs.next(), ims.next(cfg.time_delta) if sensors else None, fs.next() if (sensors or cfg.rough_terrain_mode) else None)
        eng.status_flags()
        st = eng.get_state()
        eng.set_state(st)
        eng.pack_messages(0, 5)
        if not (cfg.gravity_aligned_tips and D > 3):
            for _ in range(3):
                eng.step_to_new_stance()
                eng.execute_sequence(False)
            eng.pack_legs(0.1)
            eng.unpack_legs(0.1)
        eng.startup_begin(None)
        eng.startup_step()
        eng.generate_workspaces(False, 4)
        torch.cuda.synchronize()
        assert torch.isfinite(eng.joints).all()
        eng.close()
print("sanitize smoke ok")
