"""GPU: the CUDA control-cycle path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): |joint angle difference| <= 1e-6 rad per joint.  The parity mode is
SHC_PRECISION_F64.  Three layers:
  * one cycle from identical state, sampled all along the rollouts: EVERY state field and joint to 1e-11 (the engine
    computes the same function as the reference);
  * free-running rollouts: every open-loop state field (tip trajectories, body velocity, poses, phases, walk state,
    admittance / IMU states) to 1e-8 at all times;
  * free-running joint angles: <= 1e-6 rad except inside the reference's own numerically unstable chatter windows
    (gpu_common.JointErrors), where any two double-precision implementations differ — bounded and counted.
The reference cannot be built on the GPU box (nor in the build container: ROS/Eigen/Boost absent), so the oracle —
pinned by tests/test_oracle_*.py and the committed fixtures under tests/golden/ — is the checker."""
import glob
import os

import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

from gpu_common import JOINT_FIELDS, JointErrors, assert_state_close, run_both

pytestmark = pytest.mark.gpu
TOL = 1e-6          # rad, north_star
STATE_TOL = 1e-8    # double state fields in f64 mode
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(cfg, n, oracle_batch, precision="f64"):
    from syropod_highlevel_controller_b200.engine import Engine

    return Engine(cfg, n, precision=precision, startup=oracle_batch.startup())


def _cfg_for(name):
    if name.startswith("config1_100hz"):
        return hexapod_config("tripod_gait", 0.01)
    if name.startswith("config1_50hz"):
        return hexapod_config("tripod_gait", 0.02)
    if name.startswith("octopod"):
        return octopod_config("tripod_gait", 0.02)
    return hexapod_config(name.split("_")[0] + "_gait", 0.02)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=os.path.basename)
def test_golden_fixture_rollouts(shc_lib, oracle, path):
    """BASELINE configs[0]: single robot, 1000 cycles (and one rollout per gait + the octopod), against the committed
    golden vectors, every cycle."""
    import torch

    g = np.load(path)
    cfg = _cfg_for(os.path.basename(path)[:-4])
    ob = oracle.OracleBatch(cfg, 1)
    eng = _engine(cfg, 1, ob)
    errs = JointErrors()
    for c in range(len(g["cmd"])):
        j = eng.step(torch.from_numpy(g["cmd"][c][None]).cuda(),
                     torch.from_numpy(g["imu"][c][None]).cuda() if "imu" in g else None,
                     torch.from_numpy(g["force"][c][None]).cuda() if "force" in g else None)
        errs.add(np.abs(j.cpu().numpy()[0].astype(np.float64) - g["joints"][c]))
        if c % 50 == 49:
            st = eng.get_state()[0]
            tips = np.array([list(st.legs[l].tip_position) for l in range(cfg.leg_count)])
            assert np.abs(tips - g["tips"][c]).max() < STATE_TOL
            assert st.walk_state == g["walk_state"][c]
    errs.check(max_fraction=0.01)
    eng.close(); ob.close()


def test_config2_batch4096_tripod(shc_lib, oracle):
    """BASELINE configs[1]: 4096 hexapods, tripod gait, per-robot random command streams (splitmix64), 300 cycles."""
    cfg = hexapod_config("tripod_gait")
    n = 4096
    ob = oracle.OracleBatch(cfg, n)
    eng = _engine(cfg, n, ob)
    errs = run_both(eng, ob, 300, CommandStream(n, min_len=50, max_len=200))
    errs.check(max_fraction=1e-3)
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    eng.close(); ob.close()


@pytest.mark.parametrize("gait", ["wave_gait", "amble_gait", "ripple_gait", "tripod_gait"])
def test_config3_gait_sweep(shc_lib, oracle, gait):
    """BASELINE configs[2] at oracle-sized batch: every gait, long enough for STARTING -> MOVING -> STOPPING -> STOPPED."""
    cfg = hexapod_config(gait)
    n = 512
    ob = oracle.OracleBatch(cfg, n)
    eng = _engine(cfg, n, ob)
    seen = set()

    def watch(c, jg, o):
        if c % 25 == 0:
            seen.update(int(s.walk_state) for s in o.get_state())

        if c % 100 == 99:
            assert_state_close(eng.get_state(), o.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)

    errs = run_both(eng, ob, 900, CommandStream(n, min_len=60, max_len=360), per_cycle=watch)
    errs.check(max_fraction=1e-3)
    assert seen == {0, 1, 2, 3}
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    eng.close(); ob.close()


def test_config4_octopod_admittance_imu_inclination(shc_lib, oracle):
    """BASELINE configs[3] at oracle-sized batch: 8 legs x 5 DOF, admittance + IMU PID + inclination posing."""
    cfg = octopod_config("tripod_gait")
    n = 512
    ob = oracle.OracleBatch(cfg, n)
    eng = _engine(cfg, n, ob)
    errs = run_both(eng, ob, 400, CommandStream(n, min_len=50, max_len=200), ImuStream(n), ForceStream(n, 8), dt=cfg.time_delta)
    errs.check(max_fraction=1e-3)
    d = assert_state_close(eng.get_state(), ob.get_state(), 8, 5, STATE_TOL, skip=JOINT_FIELDS)
    assert d["admittance_state"] < 1e-12 and d["imu_pose"] < 1e-12
    eng.close(); ob.close()


def test_auto_posing_and_100hz(shc_lib, oracle):
    """Auto posing (pose_controller.cpp:1134-1187, 1338-1439, 1716-1778) for every gait, at time_delta 0.01."""
    for gait in ("tripod_gait", "wave_gait", "ripple_gait", "amble_gait"):
        cfg = hexapod_config(gait, 0.01, auto_posing=1)
        n = 64
        ob = oracle.OracleBatch(cfg, n)
        eng = _engine(cfg, n, ob)
        errs = run_both(eng, ob, 1200, CommandStream(n, min_len=150, max_len=500), dt=0.01)
        errs.check(max_fraction=5e-3)
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
        eng.close(); ob.close()


def test_other_parameter_variants(shc_lib, oracle):
    """real-velocity input mode, force_normal_touchdown, swing width / stance span, no manual posing, unclamped joints."""
    variants = [dict(velocity_input_mode=1), dict(force_normal_touchdown=1), dict(swing_width=0.01, stance_span_modifier=0.3),
                dict(manual_posing=0), dict(clamp_joint_positions=0, clamp_joint_velocities=0), dict(body_velocity_scaler=0.7)]
    for kw in variants:
        cfg = hexapod_config("ripple_gait", **kw)
        n = 96
        ob = oracle.OracleBatch(cfg, n)
        eng = _engine(cfg, n, ob)
        cs = CommandStream(n, min_len=60, max_len=240)
        if kw.get("velocity_input_mode") == 1:  # "real" mode takes m/s and rad/s
            base = cs.next
            cs.next = lambda: base() * np.array([0.08, 0.08, 0.5], dtype=np.float32)
        errs = run_both(eng, ob, 500, cs)
        errs.check(max_fraction=2e-3)
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
        eng.close(); ob.close()


def test_manual_pose_inputs_and_reset_modes(shc_lib, oracle):
    """PoseController::updateManualPose (pose_controller.cpp:863-1003) driven by joystick-style inputs."""
    import torch

    cfg = hexapod_config("tripod_gait")
    n = 32
    ob = oracle.OracleBatch(cfg, n)
    eng = _engine(cfg, n, ob)
    rng = np.random.default_rng(11)
    cs = CommandStream(n, min_len=50, max_len=150)
    man = np.zeros((n, 6), dtype=np.float32)
    errs = JointErrors()
    for c in range(400):
        if c % 40 == 0:
            man = rng.choice([-1.0, 0.0, 0.5, 1.0], size=(n, 6)).astype(np.float32)
        cmd = cs.next()
        j = eng.step(torch.from_numpy(cmd).cuda(), manual=torch.from_numpy(man).cuda())
        ob.step(cmd.astype(np.float64), manual=man.astype(np.float64))
        errs.add(np.abs(j.cpu().numpy().astype(np.float64) - ob.joints()))
    errs.check(max_fraction=2e-3)
    d = assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    assert max(abs(v) for s in ob.get_state() for v in list(s.manual_pose)[:3]) > 0.01  # the pose really moved
    eng.close(); ob.close()


def test_single_step_parity_all_modes(shc_lib, oracle):
    """One cycle from identical state: every state field, both precisions.  Mixed precision is held to 1e-6 here;
    over long rollouts it cannot follow the reference's stand-still limit cycle (DESIGN.md "Precision")."""
    import torch
    from syropod_highlevel_controller_b200.engine import Engine

    def f64(a):
        return None if a is None else a.astype(np.float64)

    def dev(a):
        return None if a is None else torch.from_numpy(a).cuda()

    for cfg, L, D, sensors in ((hexapod_config("tripod_gait"), 6, 3, False), (hexapod_config("wave_gait"), 6, 3, False),
                               (octopod_config("tripod_gait"), 8, 5, True)):
        n = 128
        ob = oracle.OracleBatch(cfg, n)  # the trajectory generator: supplies realistic states
        cs = CommandStream(n, min_len=30, max_len=120)
        ims = ImuStream(n) if sensors else None
        fs = ForceStream(n, L) if sensors else None
        e64 = Engine(cfg, n, precision="f64", startup=ob.startup())
        emx = Engine(cfg, n, precision="mixed", startup=ob.startup())
        ref = oracle.OracleBatch(cfg, n)  # re-seated on the fp32-rounded state for the mixed comparison
        for c in range(260):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            sample = c % 4 == 3
            if sample:
                snap = ob.get_state()
                e64.set_state(snap)
                emx.set_state(snap)
                ref.set_state(emx.get_state())
                j64 = e64.step(dev(cmd), dev(imu), dev(force)).cpu().numpy().astype(np.float64)
                jmx = emx.step(dev(cmd), dev(imu), dev(force)).cpu().numpy().astype(np.float64)
                ref.step(f64(cmd), f64(imu), f64(force), threads=8)
            ob.step(f64(cmd), f64(imu), f64(force), threads=8)
            if sample:
                assert np.abs(j64 - ob.joints()).max() <= 1e-7, c  # float32 output rounding only
                assert_state_close(e64.get_state(), ob.get_state(), L, D, 1e-11, vel_tol=1e-9)
                assert np.abs(jmx - ref.joints()).max() <= TOL, c
                assert_state_close(emx.get_state(), ref.get_state(), L, D, 2e-6, vel_tol=2e-4, skip=("odometry_ideal",))
        for x in (e64, emx, ob, ref):
            x.close()


def test_mixed_precision_rollout_statistics(shc_lib, oracle):
    """Mixed precision over a long rollout: the typical joint error stays far below 1e-6 rad; excursions are bounded by
    the amplitude of the reference's own period-2 joint chatter (~2.3e-3 rad peak to peak), which fp32 state cannot
    phase-track (DESIGN.md "Precision").  This documents the throughput mode; the parity mode is f64 (tests above)."""
    cfg = hexapod_config("tripod_gait")
    n = 256
    ob = oracle.OracleBatch(cfg, n)
    eng = _engine(cfg, n, ob, "mixed")
    errs = []
    run_both(eng, ob, 600, CommandStream(n), per_cycle=lambda c, jg, o: errs.append(np.abs(jg - o.joints()).reshape(n, -1).max(axis=1)))
    errs = np.array(errs)
    assert np.median(errs) < 5e-7
    assert np.quantile(errs, 0.9) < 2e-6
    assert errs.max() < JointErrors.CHATTER_BOUND
    # stepper tips are open-loop and accumulate in double: they stay close regardless
    d = assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1.0, skip=())
    assert d["tip_position"] < 1e-6 and d["int:phase"] == 0 and d["int:walk_state"] == 0
    eng.close(); ob.close()
