#!/usr/bin/env python
"""Runs a configuration through the reference's own code (oracle/_ref), the restated oracle and — when a CUDA device is
present — the B200 engine, on the same synthetic command streams, and prints the largest difference per state field.

    python tools/check_against_reference.py [--default default.yaml --gait-file gait.yaml --auto-pose-file auto_pose.yaml]
                                            [--gait tripod_gait] [--robots 4] [--cycles 600]

Without file arguments the reference's shipped config/ directory is used when /root/reference is here, else the built-in
hexapod.  Test / integration tooling: it imports oracle/, so it is not part of the product."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from gpu_common import state_diff
    from oracle import oracle_py as O
    from oracle import ref_py as R
    from syropod_highlevel_controller_b200.config import ShcRobotState, hexapod_config, load_reference_yaml
    from syropod_highlevel_controller_b200.streams import CommandStream

    ap = argparse.ArgumentParser()
    ap.add_argument("--default")
    ap.add_argument("--gait-file")
    ap.add_argument("--auto-pose-file")
    ap.add_argument("--gait", default="tripod_gait")
    ap.add_argument("--robots", type=int, default=4)
    ap.add_argument("--cycles", type=int, default=600)
    a = ap.parse_args()
    d = os.path.join(R.REFERENCE_ROOT, "config")
    if a.default or os.path.isdir(d):
        cfg = load_reference_yaml(a.default or os.path.join(d, "default.yaml"), a.gait_file or os.path.join(d, "gait.yaml"),
                                  a.auto_pose_file or os.path.join(d, "auto_pose.yaml"), gait=a.gait)
        src = a.default or os.path.join(d, "default.yaml")
    else:
        cfg, src = hexapod_config(a.gait), "built-in hexapod"
    n, L, D = a.robots, cfg.leg_count, cfg.joint_count
    refs = [R.RefRobot(cfg) for _ in range(n)]
    ob = O.OracleBatch(cfg, n)
    eng = None
    try:
        import torch

        if torch.cuda.is_available():
            from syropod_highlevel_controller_b200.engine import Engine

            eng = Engine(cfg, n, precision="f64", startup=refs[0].startup())
    except Exception as ex:  # no CUDA device / extension: the CPU comparison still runs
        print(f"(engine not compared: {ex})")
    cs = CommandStream(n, min_len=60, max_len=200)
    worst_o, worst_e, wj = {}, {}, 0.0

    def ref_states():
        arr = (ShcRobotState * n)()
        for i, r in enumerate(refs):
            arr[i] = r.get_state()
        return arr

    for c in range(a.cycles):
        cmd = cs.next()
        for i, r in enumerate(refs):
            r.step(cmd[i].astype(np.float64))
        ob.step(cmd.astype(np.float64))
        sr = ref_states()
        for k, v in state_diff(ob.get_state(), sr, L, D).items():
            worst_o[k] = max(worst_o.get(k, 0.0), v)
        if eng is not None:
            j = eng.step(torch.from_numpy(cmd).cuda()).cpu().numpy().astype(np.float64)
            wj = max(wj, float(np.abs(j - np.stack([r.joints() for r in refs])).max()))
            if c % 10 == 9:
                for k, v in state_diff(eng.get_state(), sr, L, D).items():
                    worst_e[k] = max(worst_e.get(k, 0.0), v)
    print(f"{src}, {a.gait}: {n} robots x {a.cycles} cycles, {L} legs x {D} joints")
    print("restated oracle vs the reference's own code, largest difference per field (0 = bit-identical):")
    for k, v in sorted(worst_o.items()):
        if v != 0.0:
            print(f"  {k:32s} {v:.3e}")
    print(f"  fields that differ: {sum(1 for v in worst_o.values() if v != 0.0)} of {len(worst_o)}")
    if eng is not None:
        print(f"B200 engine vs the reference's own code (free-running; float32 joint output): joints {wj:.3e} rad; state fields:")
        for k, v in sorted(worst_e.items()):
            if v > 1e-9:
                print(f"  {k:32s} {v:.3e}")
    n_assert, first = refs[0].assert_failures()
    print(f"ROS_ASSERT violations inside the reference: {n_assert} {first}")


if __name__ == "__main__":
    main()
