"""Generates the golden fixtures under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and cannot be built in this container (parity is unpinned), so
these fixtures pin the ORACLE's output at the time of writing: they guard the oracle against regressions and give the
GPU box (where /root/reference does not exist) fixed vectors to compare the CUDA path with.

    python tests/golden/make_golden.py

Fixtures (BASELINE.json configs[0]): single hexapod, tripod gait, default.yaml parameters, 1000 cycles:
  config1_100hz_straight.npz   time_delta 0.01, command (1,0,0) for cycles 0-599 then (0,0,0)
  config1_100hz_cruise.npz     time_delta 0.01, command (1,0,0.5) (default.yaml:93-94) for 0-599 then zero
  config1_50hz_straight.npz    time_delta 0.02 as shipped
plus one 600-cycle rollout per gait at 50 Hz (wave / amble / ripple) and an octopod rollout with IMU + tip forces.
Each file: joints [cycles, L, D] (every cycle), tips [cycles, L, 3] (stepper tip positions), walk_state [cycles],
cmd [cycles, 3] and, for the octopod, imu [cycles, 10] and force [cycles, L, 3].
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config  # noqa: E402
from syropod_highlevel_controller_b200.streams import ForceStream, ImuStream  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rollout(cfg, cmds, imu=None, force=None):
    L, D = cfg.leg_count, cfg.joint_count
    ob = O.OracleBatch(cfg, 1)
    n = len(cmds)
    joints = np.zeros((n, L, D))
    tips = np.zeros((n, L, 3))
    ws = np.zeros(n, dtype=np.int32)
    for c in range(n):
        ob.step(cmds[c][None].astype(np.float64), None if imu is None else imu[c][None].astype(np.float64),
                None if force is None else force[c][None].astype(np.float64))
        joints[c] = ob.joints()[0]
        st = ob.get_state()[0]
        ws[c] = st.walk_state
        for l in range(L):
            tips[c, l] = list(st.legs[l].tip_position)
    ob.close()
    return joints, tips, ws


def commands(n, first, switch=600):
    cmd = np.zeros((n, 3), dtype=np.float32)
    cmd[:switch] = np.asarray(first, dtype=np.float32)
    return cmd


def main():
    cases = {
        "config1_100hz_straight": (hexapod_config("tripod_gait", 0.01), commands(1000, (1.0, 0.0, 0.0))),
        "config1_100hz_cruise": (hexapod_config("tripod_gait", 0.01), commands(1000, (1.0, 0.0, 0.5))),
        "config1_50hz_straight": (hexapod_config("tripod_gait", 0.02), commands(1000, (1.0, 0.0, 0.0))),
        "wave_50hz": (hexapod_config("wave_gait", 0.02), commands(600, (0.6, 0.3, -0.2), 420)),
        "amble_50hz": (hexapod_config("amble_gait", 0.02), commands(600, (-0.5, 0.4, 0.3), 420)),
        "ripple_50hz": (hexapod_config("ripple_gait", 0.02), commands(600, (0.2, -0.7, 0.1), 420)),
    }
    for name, (cfg, cmd) in cases.items():
        joints, tips, ws = rollout(cfg, cmd)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), joints=joints, tips=tips, walk_state=ws, cmd=cmd)
        print(name, joints.shape, "walk states seen", sorted(set(ws.tolist())))
    cfg = octopod_config("tripod_gait", 0.02)
    n = 500
    cmd = commands(n, (0.7, 0.2, 0.2), 350)
    ims, fs = ImuStream(1), ForceStream(1, 8)
    imu = np.stack([ims.next(cfg.time_delta)[0] for _ in range(n)])
    force = np.stack([fs.next()[0] for _ in range(n)])
    joints, tips, ws = rollout(cfg, cmd, imu, force)
    np.savez_compressed(os.path.join(HERE, "octopod_50hz.npz"), joints=joints, tips=tips, walk_state=ws, cmd=cmd, imu=imu,
                        force=force)
    print("octopod_50hz", joints.shape, sorted(set(ws.tolist())))


if __name__ == "__main__":
    main()
