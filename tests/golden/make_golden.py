"""Generates the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

The reference ships no golden vectors.  These are outputs of the reference's own control code —
oracle/_ref/libshc_ref.so: its unmodified state_controller.cpp, model.cpp, walk_controller.cpp, pose_controller.cpp,
admittance_controller.cpp compiled from /root/reference against stand-in ROS / Eigen / Boost headers (oracle/Makefile.ref,
oracle/ref_py.py) — run in this container.  /root/reference does not exist on the GPU box, so the vectors are committed:
they are what the oracle (tests/test_oracle_*.py, CPU) and the CUDA path (`-m gpu`) are compared with there.  The
script also runs the restated oracle on the same inputs and reports the difference (zero on the hexapod, ~1e-15 rad on
the octopod).

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
from oracle import ref_py as R  # noqa: E402
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config  # noqa: E402
from syropod_highlevel_controller_b200.streams import ForceStream, ImuStream  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rollout(cfg, cmds, imu=None, force=None):
    """The reference's rollout; the restated oracle runs beside it and the largest joint difference is printed."""
    L, D = cfg.leg_count, cfg.joint_count
    ref = R.RefRobot(cfg)
    ob = O.OracleBatch(cfg, 1)
    n = len(cmds)
    joints = np.zeros((n, L, D))
    tips = np.zeros((n, L, 3))
    ws = np.zeros(n, dtype=np.int32)
    worst = 0.0
    for c in range(n):
        i = None if imu is None else imu[c].astype(np.float64)
        f = None if force is None else force[c].astype(np.float64)
        ref.step(cmds[c].astype(np.float64), i, f)
        ob.step(cmds[c][None].astype(np.float64), None if i is None else i[None], None if f is None else f[None])
        joints[c] = ref.joints()
        worst = max(worst, float(np.abs(joints[c] - ob.joints()[0]).max()))
        st = ref.get_state()
        ws[c] = st.walk_state
        for l in range(L):
            tips[c, l] = list(st.legs[l].tip_position)
    n_assert, first = ref.assert_failures()
    assert n_assert == 0, first
    ref.close(); ob.close()
    print(f"  restated oracle vs reference: worst joint difference {worst:.2e} rad")
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
        "autopose_tripod_50hz": (hexapod_config("tripod_gait", 0.02, auto_posing=1), commands(600, (0.8, -0.2, 0.3), 420)),
    }
    for name, (cfg, cmd) in cases.items():
        joints, tips, ws = rollout(cfg, cmd)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), joints=joints, tips=tips, walk_state=ws, cmd=cmd, source="reference")
        print(name, joints.shape, "walk states seen", sorted(set(ws.tolist())))
    cfg = octopod_config("tripod_gait", 0.02)
    n = 500
    cmd = commands(n, (0.7, 0.2, 0.2), 350)
    ims, fs = ImuStream(1), ForceStream(1, 8)
    imu = np.stack([ims.next(cfg.time_delta)[0] for _ in range(n)])
    force = np.stack([fs.next()[0] for _ in range(n)])
    joints, tips, ws = rollout(cfg, cmd, imu, force)
    np.savez_compressed(os.path.join(HERE, "octopod_50hz.npz"), joints=joints, tips=tips, walk_state=ws, cmd=cmd, imu=imu,
                        force=force, source="reference")
    print("octopod_50hz", joints.shape, sorted(set(ws.tolist())))


if __name__ == "__main__":
    main()
