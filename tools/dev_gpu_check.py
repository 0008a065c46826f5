"""Development check run under gpurun: parity of the CUDA path vs the oracle + crude timing.  Not part of the product."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.streams import CommandStream, ImuStream, ForceStream
from oracle import oracle_py as O


def state_diff(se, so, L, D):
    out = {}
    def upd(k, a, b):
        d = float(np.max(np.abs(np.array(a, dtype=float) - np.array(b, dtype=float))))
        out[k] = max(out.get(k, 0.0), d)
    for r in range(len(se)):
        a, b = se[r], so[r]
        for f in ("desired_linear_velocity", "walk_plane", "walk_plane_normal", "odometry_ideal", "walk_plane_pose",
                  "origin_walk_plane_pose", "manual_pose", "imu_pose", "inclination_pose", "auto_pose", "rotation_absement_error",
                  "rotation_velocity_error", "current_pose"):
            upd(f, list(getattr(a, f)), list(getattr(b, f)))
        upd("desired_angular_velocity", a.desired_angular_velocity, b.desired_angular_velocity)
        for f in ("walk_state", "legs_at_correct_phase", "legs_completed_first_step", "return_to_default_attempted", "pose_state"):
            upd(f, getattr(a, f), getattr(b, f))
        for l in range(L):
            la, lb = a.legs[l], b.legs[l]
            upd("joint_position", list(la.joint_position)[:D], list(lb.joint_position)[:D])
            upd("joint_velocity", list(la.joint_velocity)[:D], list(lb.joint_velocity)[:D])
            for f in ("tip_position", "tip_velocity", "swing_origin_position", "stance_origin_position",
                      "default_tip_position", "target_tip_position", "stride_vector", "walk_plane", "walk_plane_normal",
                      "admittance_state", "admittance_delta", "model_tip_position"):
                upd(f, list(getattr(la, f)), list(getattr(lb, f)))
            for f in ("phase", "step_state", "at_correct_phase", "completed_first_step", "swing_progress", "stance_progress"):
                upd(f, getattr(la, f), getattr(lb, f))
    return out


def rollout(cfg, n, cycles, precision, use_imu=False, use_force=False, check_every=50, tag=""):
    ob = O.OracleBatch(cfg, n)
    eng = Engine(cfg, n, precision=precision, startup=ob.startup())
    L, D = cfg.leg_count, cfg.joint_count
    cs = CommandStream(n, min_len=40, max_len=160)
    ims = ImuStream(n) if use_imu else None
    fs = ForceStream(n, L) if use_force else None
    maxq = 0.0
    worst = {}
    for c in range(cycles):
        cmd = cs.next()
        imu = ims.next(cfg.time_delta) if ims else None
        force = fs.next() if fs else None
        j = eng.step(torch.from_numpy(cmd).cuda(), None if imu is None else torch.from_numpy(imu).cuda(),
                     None if force is None else torch.from_numpy(force).cuda())
        ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                None if force is None else force.astype(np.float64), threads=8)
        dq = np.abs(j.cpu().numpy().astype(np.float64) - ob.joints())
        maxq = max(maxq, dq.max())
        if (c + 1) % check_every == 0 or c == cycles - 1:
            sd = state_diff(eng.get_state(), ob.get_state(), L, D)
            for k, v in sd.items():
                worst[k] = max(worst.get(k, 0.0), v)
    print(f"[{tag}] {precision} n={n} cycles={cycles}: max |dq| (f32 out) = {maxq:.3e}")
    print("    state diffs:", {k: float(f"{v:.2e}") for k, v in sorted(worst.items()) if v > 0})
    eng.close(); ob.close()
    return maxq


def timing(cfg, n, precision, steps=20, tag=""):
    eng = Engine(cfg, n, precision=precision)
    cmd = torch.from_numpy(CommandStream(n).next()).cuda()
    for _ in range(5):
        eng.step(cmd)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        eng.step(cmd)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    b = eng.bytes_per_step_algorithmic
    print(f"[{tag}] {precision} n={n}: {ms*1e3:.1f} us/step, {n/ms*1e3:.3e} steps/s, alg {b} B/step -> {n*b/ms/1e6:.1f} GB/s "
          f"({n*b/ms/1e6/6553.6*100:.1f}% of 6553.6), device bytes {eng.bytes_per_step_device}")
    eng.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    t0 = time.time()
    for prec in ("f64", "mixed"):
        rollout(hexapod_config("tripod_gait"), 64, 600, prec, tag="hex tripod")
    for gait in ("wave_gait", "ripple_gait", "amble_gait"):
        rollout(hexapod_config(gait), 32, 500, "f64", tag="hex " + gait)
        rollout(hexapod_config(gait), 32, 500, "mixed", tag="hex " + gait)
    for prec in ("f64", "mixed"):
        rollout(octopod_config("tripod_gait"), 32, 400, prec, use_imu=True, use_force=True, tag="octo")
    print("parity time", time.time() - t0)
    for prec in ("f64", "mixed"):
        for n in (4096, 65536, 131072, 262144):
            timing(hexapod_config(), n, prec, tag="hex")
        timing(octopod_config(), 262144, prec, tag="octo")
