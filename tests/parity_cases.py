"""Parity cases shared by the GPU tests (tests/test_gpu_parity.py: the CUDA kernel through the C-ABI) and the CPU tests
(tests/test_emu_parity.py: the same cycle source compiled for the host, tests/emu.py).  The oracle is the checker.

Tolerance (BASELINE.json north_star): |joint angle difference| <= 1e-6 rad per joint; the parity mode is f64.  Layers:
  * one cycle from identical state, sampled all along the rollouts: EVERY state field and joint to 1e-11;
  * free-running rollouts: every open-loop state field (tip trajectories, body velocity, poses, phases, walk state,
    admittance / IMU states) to 1e-8 at all times;
  * free-running joint angles: <= 1e-6 rad except inside the reference's own numerically unstable chatter windows
    (gpu_common.JointErrors; the oracle-vs-oracle evidence is tests/test_oracle_chatter.py), bounded and counted.
The caps below are ~3x the fractions observed on the B200 (printed by JointErrors.check into the test log)."""
import glob
import os

import numpy as np

from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

from gpu_common import JOINT_FIELDS, JointErrors, assert_state_close, run_both

TOL = 1e-6          # rad, north_star
STATE_TOL = 1e-8    # double state fields in f64 mode
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))

# caps on the fraction of joint samples beyond 1e-6 rad (chatter windows): 50 Hz and 100 Hz control rates
CAP_50HZ = 5e-4
CAP_100HZ = 2e-3


def cfg_for_golden(name):
    if name.startswith("config1_100hz"):
        return hexapod_config("tripod_gait", 0.01)
    if name.startswith("config1_50hz"):
        return hexapod_config("tripod_gait", 0.02)
    if name.startswith("octopod"):
        return octopod_config("tripod_gait", 0.02)
    if name.startswith("autopose_tripod"):
        return hexapod_config("tripod_gait", 0.02, auto_posing=1)
    return hexapod_config(name.split("_")[0] + "_gait", 0.02)


def golden_rollout(backend, oracle, path):
    """BASELINE configs[0]: single robot, 1000 cycles (and one rollout per gait + the octopod), against the committed
    golden vectors, every cycle."""
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    cfg = cfg_for_golden(name)
    ob = oracle.OracleBatch(cfg, 1)
    eng = backend.engine(cfg, 1, startup=ob.startup())
    errs = JointErrors()
    for c in range(len(g["cmd"])):
        j = eng.step(g["cmd"][c][None], g["imu"][c][None] if "imu" in g else None, g["force"][c][None] if "force" in g else None)
        errs.add(np.abs(j[0] - g["joints"][c]))
        if c % 50 == 49:
            st = eng.get_state()[0]
            tips = np.array([list(st.legs[l].tip_position) for l in range(cfg.leg_count)])
            assert np.abs(tips - g["tips"][c]).max() < STATE_TOL
            assert st.walk_state == g["walk_state"][c]
    # one robot, 18-40 joints, 1000 cycles: a single chatter window of ~20 cycles on one leg is 60 samples of 18000
    errs.check(max_fraction=5e-3 if "100hz" in name else 2e-3, label=f"golden {name}")
    eng.close(); ob.close()


def batch_tripod(backend, oracle, n=4096, cycles=300):
    """BASELINE configs[1]: hexapods, tripod gait, per-robot random command streams (splitmix64)."""
    cfg = hexapod_config("tripod_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=ob.startup())
    errs = run_both(eng, ob, cycles, CommandStream(n, min_len=50, max_len=200))
    errs.check(max_fraction=CAP_50HZ, label=f"config2 tripod n={n}")
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    eng.close(); ob.close()


def gait_sweep(backend, oracle, gait, n=512, cycles=900, cap=CAP_50HZ):
    """BASELINE configs[2] at oracle-sized batch: long enough for STARTING -> MOVING -> STOPPING -> STOPPED."""
    cfg = hexapod_config(gait)
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=ob.startup())
    seen = set()

    def watch(c, jg, o):
        if c % 25 == 0:
            seen.update(int(s.walk_state) for s in o.get_state())
        if c % 100 == 99:
            assert_state_close(eng.get_state(), o.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)

    errs = run_both(eng, ob, cycles, CommandStream(n, min_len=60, max_len=360), per_cycle=watch)
    errs.check(max_fraction=cap, label=f"config3 {gait} n={n}")
    assert seen == {0, 1, 2, 3}
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    eng.close(); ob.close()


def octopod_full(backend, oracle, n=512, cycles=400):
    """BASELINE configs[3] at oracle-sized batch: 8 legs x 5 DOF, admittance + IMU PID + inclination posing."""
    cfg = octopod_config("tripod_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=ob.startup())
    errs = run_both(eng, ob, cycles, CommandStream(n, min_len=50, max_len=200), ImuStream(n), ForceStream(n, 8), dt=cfg.time_delta)
    errs.check(max_fraction=CAP_50HZ, label=f"config4 octopod n={n}")
    d = assert_state_close(eng.get_state(), ob.get_state(), 8, 5, STATE_TOL, skip=JOINT_FIELDS)
    assert d["admittance_state"] < 1e-12 and d["imu_pose"] < 1e-12
    # dynamic stiffness (admittance_controller.cpp:96) is live: some swinging leg has left the global value
    ks = [s.legs[l].virtual_stiffness for s in ob.get_state() for l in range(8)]
    assert min(ks) < cfg.virtual_stiffness < max(ks)
    eng.close(); ob.close()


def auto_posing_100hz(backend, oracle, gaits=("tripod_gait", "wave_gait", "ripple_gait", "amble_gait"), n=64, cycles=1200):
    """Auto posing (pose_controller.cpp:1134-1187, 1338-1439, 1716-1778) for every gait, at time_delta 0.01."""
    for gait in gaits:
        cfg = hexapod_config(gait, 0.01, auto_posing=1)
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        errs = run_both(eng, ob, cycles, CommandStream(n, min_len=150, max_len=500), dt=0.01)
        errs.check(max_fraction=CAP_100HZ, label=f"auto posing {gait} 100 Hz")
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
        eng.close(); ob.close()


def auto_posing_own_cycle(backend, oracle, n=32, cycles=800):
    """Auto posing on its own cycle (pose_frequency != -1: the non-synchronised branches of pose_controller.cpp:1134-1187 and
    AutoPoser::updatePose :1338): the posers cycle from the second start-up loop on, whether the robot walks or not, and the
    engine's initial state replays those loops (csrc/shc_host.cuh initial_state).  In the reference this mode leaves the
    body swaying while generateWorkspaces runs, the tip check of model.cpp:330 fails and the workspace — hence every speed
    limit — comes out zero: the robot poses but cannot walk (all shipped configurations use -1).  The oracle reproduces
    that, so this case checks the posing itself; the engine refuses to compute its OWN start-up for such a configuration."""
    from backends import Backend  # noqa: F401  (ShcError / EmuError types differ per backend)

    for gait, dt, freq in (("tripod_gait", 0.02, 0.7), ("wave_gait", 0.01, 1.3)):
        cfg = hexapod_config(gait, dt, auto_posing=1, pose_frequency=freq)
        ob = oracle.OracleBatch(cfg, n)
        assert max(ob.startup().walkspace) == 0.0
        eng = backend.engine(cfg, n, startup=ob.startup())
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1e-12, skip=JOINT_FIELDS)
        errs = run_both(eng, ob, cycles, CommandStream(n, min_len=100, max_len=300), dt=dt)
        errs.check(max_fraction=5e-3, label=f"auto posing, own cycle {freq} Hz, {gait}")
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1e-8, skip=JOINT_FIELDS)
        assert max(abs(v) for s in ob.get_state() for v in list(s.auto_pose)) > 1e-4  # the body is swaying
        eng.close(); ob.close()
        try:
            backend.engine(cfg, 1, startup=None)
            raise AssertionError("own start-up with a free-running poser should be refused")
        except RuntimeError as ex:
            assert "start-up" in str(ex)


def parameter_variants(backend, oracle, n=96, cycles=500):
    """real-velocity input mode, force_normal_touchdown, swing width / stance span, no manual posing, unclamped joints."""
    variants = [dict(velocity_input_mode=1), dict(force_normal_touchdown=1), dict(swing_width=0.01, stance_span_modifier=0.3),
                dict(manual_posing=0), dict(clamp_joint_positions=0, clamp_joint_velocities=0), dict(body_velocity_scaler=0.7)]
    for kw in variants:
        cfg = hexapod_config("ripple_gait", **kw)
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        cs = CommandStream(n, min_len=60, max_len=240)
        if kw.get("velocity_input_mode") == 1:  # "real" mode takes m/s and rad/s
            base = cs.next
            cs.next = lambda: base() * np.array([0.08, 0.08, 0.5], dtype=np.float32)
        errs = run_both(eng, ob, cycles, cs)
        errs.check(max_fraction=1e-3, label=f"variant {kw}")
        assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
        eng.close(); ob.close()


def manual_pose_and_reset_modes(backend, oracle, n=32, cycles=520):
    """PoseController::updateManualPose (pose_controller.cpp:863-1003) driven by joystick-style inputs, then every
    PoseResetMode in turn (Z+yaw, X+Y, pitch+roll, all, immediate: poser_->setPoseResetMode, state_controller.cpp:1199)."""
    cfg = hexapod_config("tripod_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=ob.startup())
    rng = np.random.default_rng(11)
    cs = CommandStream(n, min_len=50, max_len=150)
    man = np.zeros((n, 6), dtype=np.float32)
    errs = JointErrors()
    moved = 0.0
    # cycles 0..239: free inputs; then blocks of 40 cycles: inputs again (20) followed by one reset mode (20)
    schedule = {240 + 56 * k + 28: mode for k, mode in enumerate((1, 2, 3, 4, 5))}
    mode = 0
    for c in range(cycles):
        if c in schedule:
            mode = schedule[c]
        elif c - 28 in schedule or c < 240 and mode:
            mode = 0
        if c % 28 == 0:
            man = rng.choice([-1.0, 0.0, 0.5, 1.0], size=(n, 6)).astype(np.float32)
        eng.set_pose_reset_mode(mode)
        ob.set_pose_reset_mode(mode)
        cmd = cs.next()
        j = eng.step(cmd, manual=man)
        ob.step(cmd.astype(np.float64), manual=man.astype(np.float64))
        errs.add(np.abs(j - ob.joints()))
        if c % 7 == 6:
            assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
        moved = max(moved, max(abs(v) for s in ob.get_state() for v in list(s.manual_pose)[:3]))
    errs.check(max_fraction=2e-3, label="manual pose + reset modes")
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, STATE_TOL, skip=JOINT_FIELDS)
    assert moved > 0.01  # the pose really moved
    eng.close(); ob.close()


def joint_effort_tip_force(backend, oracle, n=64, cycles=240):
    """use_joint_effort = 1: Leg::calculateTipForce (model.cpp:667-708) from measured joint efforts feeds the admittance
    controller (force_gain applied twice, trap 4).  Efforts change every 30 cycles.

    The estimate J (J^T J + l^2 I)^-1 tau is very sensitive to the joint angles (~1e4 N/rad for these legs), and it feeds
    back into them through the admittance delta: two builds of the ORACLE (with / without FMA contraction) agree to 1e-12 N
    after one cycle from identical state and drift to ~1e-5 N along a free-running rollout of the 5-DOF leg
    (tests/test_oracle_chatter.py prints it).  So the free-running check here is statistical — joints under the usual
    cap at 1e-5 rad, forces by quantile — and the tight functional check is single_step_inputs below."""
    for cfg, L, D in ((hexapod_config("tripod_gait", admittance_control=1, use_joint_effort=1), 6, 3),
                      (octopod_config("tripod_gait", use_joint_effort=1), 8, 5)):
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        rng = np.random.default_rng(5)
        cs = CommandStream(n, min_len=40, max_len=120)
        ims = ImuStream(n) if cfg.imu_posing else None
        errs = JointErrors()
        seen_force, force_diffs = 0.0, []
        loose = ("tip_force_calculated", "admittance_delta", "admittance_state")
        for c in range(cycles):
            if c % 30 == 0:
                eff = rng.normal(0.0, 2.0, size=(n, L, D)).astype(np.float32)
                eng.set_joint_efforts(eff)
                ob.set_joint_efforts(eff.astype(np.float64))
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            j = eng.step(cmd, imu)
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64), threads=4)
            errs.add(np.abs(j - ob.joints()), tol=1e-5)
            if c % 20 == 19:
                se, so = eng.get_state(), ob.get_state()
                assert_state_close(se, so, L, D, 1e-6, skip=JOINT_FIELDS + loose)
                fe = np.array([list(se[r].legs[l].tip_force_calculated) for r in range(n) for l in range(L)])
                fo = np.array([list(so[r].legs[l].tip_force_calculated) for r in range(n) for l in range(L)])
                force_diffs.append(np.abs(fe - fo).max(axis=1))
                seen_force = max(seen_force, float(np.abs(fo).max()))
        force_diffs = np.concatenate(force_diffs)
        print(f"[tip-force] {L}x{D}: |F| up to {seen_force:.1f} N, difference median {np.median(force_diffs):.2e} N, "
              f"99 % {np.quantile(force_diffs, 0.99):.2e} N, max {force_diffs.max():.2e} N")
        assert seen_force > 0.1  # the estimate is live
        assert np.median(force_diffs) < 1e-6 and np.quantile(force_diffs, 0.99) < 1e-3
        errs.check(max_fraction=2e-3, label=f"joint-effort tip force {L}x{D} (1e-5 rad)")
        eng.close(); ob.close()


def own_startup_free_running(backend, oracle, n=64, cycles=400):
    """The engine's OWN start-up (startup=NULL: csrc/shc_host.cuh — the path bench.py and every user takes) against the
    oracle's: constants to 1e-7, then a free-running 50 Hz rollout.  The default-stance joints of the two start-ups both
    end inside the stand-still limit cycle (DESIGN.md "Reference dynamics"), so joints are compared under the chatter
    bound until the robot first walks, and with the usual fraction cap after that."""
    cfg = hexapod_config("tripod_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=None)
    se, so = eng.startup(), ob.startup()
    for f in ("workspace", "walkspace", "max_linear_speed", "max_angular_speed", "max_linear_acceleration", "max_angular_acceleration"):
        a, b = np.array(getattr(se, f), dtype=float), np.array(getattr(so, f), dtype=float)
        assert np.abs(a - b).max() <= 1e-7 * max(1.0, np.abs(b).max()), f
    for f in ("period", "swing_period", "stance_period", "stance_end", "swing_start", "swing_end", "stance_start"):
        assert getattr(se, f) == getattr(so, f), f
    cs = CommandStream(n, min_len=50, max_len=200)
    errs_all, errs_walk = JointErrors(), JointErrors()
    for c in range(cycles):
        cmd = cs.next()
        j = eng.step(cmd)
        ob.step(cmd.astype(np.float64), threads=4)
        d = np.abs(j - ob.joints())
        errs_all.add(d)
        if c >= 150:
            errs_walk.add(d)
    assert errs_all.worst <= JointErrors.CHATTER_BOUND
    errs_walk.check(max_fraction=5e-3, label="own start-up, after the first steps")
    assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1e-7, skip=JOINT_FIELDS)
    eng.close(); ob.close()


def single_step_all_modes(backend, oracle, n=128, cycles=260, mixed=True):
    """One cycle from identical state: every state field, both precisions.  Mixed precision is held to 1e-6 here;
    over long rollouts it cannot follow the reference's stand-still limit cycle (DESIGN.md "Precision")."""

    def f64(a):
        return None if a is None else a.astype(np.float64)

    for cfg, L, D, sensors in ((hexapod_config("tripod_gait"), 6, 3, False), (hexapod_config("wave_gait"), 6, 3, False),
                               (octopod_config("tripod_gait"), 8, 5, True)):
        ob = oracle.OracleBatch(cfg, n)  # the trajectory generator: supplies realistic states
        cs = CommandStream(n, min_len=30, max_len=120)
        ims = ImuStream(n) if sensors else None
        fs = ForceStream(n, L) if sensors else None
        e64 = backend.engine(cfg, n, "f64", startup=ob.startup())
        emx = backend.engine(cfg, n, "mixed", startup=ob.startup()) if mixed else None
        ref = oracle.OracleBatch(cfg, n)  # re-seated on the fp32-rounded state for the mixed comparison
        for c in range(cycles):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            sample = c % 4 == 3
            if sample:
                snap = ob.get_state()
                e64.set_state(snap)
                j64 = e64.step(cmd, imu, force)
                if emx is not None:
                    emx.set_state(snap)
                    ref.set_state(emx.get_state())
                    jmx = emx.step(cmd, imu, force)
                    ref.step(f64(cmd), f64(imu), f64(force), threads=8)
            ob.step(f64(cmd), f64(imu), f64(force), threads=8)
            if sample:
                assert np.abs(j64 - ob.joints()).max() <= 1e-7, c  # float32 output rounding only
                assert_state_close(e64.get_state(), ob.get_state(), L, D, 1e-11, vel_tol=1e-9)
                if emx is not None:
                    assert np.abs(jmx - ref.joints()).max() <= TOL, c
                    dm = assert_state_close(emx.get_state(), ref.get_state(), L, D, 2e-6, vel_tol=2e-4,
                                            skip=("odometry_ideal", "virtual_stiffness"))
                    assert dm["virtual_stiffness"] < 1e-4  # values up to ~60 N/m held in fp32
        for x in (e64, emx, ob, ref):
            if x is not None:
                x.close()


def single_step_inputs(backend, oracle, n=64, cycles=200):
    """One cycle from identical state with the inputs the rollout cases above do not carry: manual pose inputs under
    every reset mode, measured joint efforts (use_joint_effort), and an engine that computed its own start-up."""
    rng = np.random.default_rng(3)
    # manual pose + reset modes
    cfg = hexapod_config("ripple_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=None)  # own start-up: the state is injected, the constants are the engine's
    cs = CommandStream(n, min_len=30, max_len=120)
    for c in range(cycles):
        mode = (c // 8) % 6
        man = rng.choice([-1.0, 0.0, 1.0], size=(n, 6)).astype(np.float32)
        cmd = cs.next()
        ob.set_pose_reset_mode(mode)
        if c % 4 == 3:
            eng.set_state(ob.get_state())
            eng.set_pose_reset_mode(mode)
            eng.step(cmd, manual=man)
        ob.step(cmd.astype(np.float64), manual=man.astype(np.float64), threads=4)
        if c % 4 == 3:
            assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1e-11, vel_tol=1e-9)
    eng.close(); ob.close()
    # joint efforts
    cfg = octopod_config("tripod_gait", use_joint_effort=1)
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, startup=ob.startup())
    cs, ims = CommandStream(n, min_len=30, max_len=120), ImuStream(n)
    for c in range(cycles):
        eff = rng.normal(0.0, 2.0, size=(n, 8, 5)).astype(np.float32)
        cmd, imu = cs.next(), ims.next(cfg.time_delta)
        ob.set_joint_efforts(eff.astype(np.float64))
        if c % 4 == 3:
            eng.set_state(ob.get_state())
            eng.set_joint_efforts(eff)
            eng.step(cmd, imu)
        ob.step(cmd.astype(np.float64), imu.astype(np.float64), threads=4)
        if c % 4 == 3:
            d = assert_state_close(eng.get_state(), ob.get_state(), 8, 5, 1e-11, vel_tol=1e-9, skip=("tip_force_calculated",))
            assert d["tip_force_calculated"] < 1e-9  # |F| ~ 10 N through a 5x5 solve
    eng.close(); ob.close()


def tip_orientation(backend, oracle, n=48, cycles=400):
    """gravity_aligned_tips.  Legs of more than three joints: LegStepper::updateTipRotation
    (walk_controller.cpp:1193) turns the stepper's tip rotation towards "tip x axis down" through the second half of every
    swing, PoseController::updateStance carries it into the poser's tip pose, and Leg::applyIK (model.cpp:880-900, 932-936)
    follows the position step by a rotation step with the full 6 x D Jacobian, retrying position-only when the constrained
    attempt misses IK_TOLERANCE.  Legs of at most three joints: PoseController::updateTipAlignPose
    (pose_controller.cpp:1024) shifts the body instead.  One cycle from identical state to 1e-11 on every field, then a
    free-running rollout."""
    for cfg, L, D in ((octopod_config("tripod_gait", gravity_aligned_tips=1), 8, 5),
                      (octopod_config("ripple_gait", gravity_aligned_tips=1, use_joint_effort=1), 8, 5),
                      (hexapod_config("tripod_gait", gravity_aligned_tips=1), 6, 3),
                      (hexapod_config("wave_gait", gravity_aligned_tips=1, auto_posing=1), 6, 3)):
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        own = backend.engine(cfg, n, startup=None)  # the engine's own initial state carries the tip rotations too
        assert_state_close(own.get_state(), ob.get_state(), L, D, 1e-7, skip=JOINT_FIELDS)
        own.close()
        cs = CommandStream(n, min_len=40, max_len=160)
        ims = ImuStream(n) if cfg.imu_posing or cfg.inclination_posing else None
        fs = ForceStream(n, L) if cfg.admittance_control and not cfg.use_joint_effort else None
        rng = np.random.default_rng(21)
        retried = rotated = shifted = 0
        for c in range(cycles):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            if cfg.use_joint_effort and c % 25 == 0:
                eff = rng.normal(0.0, 2.0, size=(n, L, D)).astype(np.float32)
                eng.set_joint_efforts(eff)
                ob.set_joint_efforts(eff.astype(np.float64))
            sample = c % 3 == 2
            if sample:
                eng.set_state(ob.get_state())
                j = eng.step(cmd, imu, force)
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                    None if force is None else force.astype(np.float64), threads=4)
            if sample:
                so = ob.get_state()
                assert np.abs(j - ob.joints()).max() <= 1e-7, c
                d = assert_state_close(eng.get_state(), so, L, D, 1e-11, vel_tol=1e-9, skip=("tip_force_calculated",))
                assert d["tip_force_calculated"] < 1e-9
                for s in so:
                    rotated += sum(1 for l in range(L) if any(v != 0.0 for v in s.legs[l].tip_rotation))
                    retried += sum(1 for l in range(L) if s.legs[l].ik_result == 0.0)
                    shifted += 1 if any(abs(v) > 1e-4 for v in list(s.tip_align_pose)[:3]) else 0
        if D > 3:
            assert rotated > 0 and retried > 0, (rotated, retried)  # rotation-constrained IK and its retry were both live
        else:
            assert shifted > 0  # the body really shifted
        print(f"[tip-orientation] {L}x{D}: leg-cycles with a defined tip rotation {rotated}, IK results of 0 {retried}, "
              f"robot-cycles with a tip-align shift {shifted}")
        # free-running
        eng.set_state(ob.get_state())
        errs = run_both(eng, ob, 200, cs, ims, fs, dt=cfg.time_delta, threads=4) if not cfg.use_joint_effort else None
        if errs is not None:
            errs.check(max_fraction=5e-3, label=f"gravity_aligned_tips {L}x{D} free-running")
            assert_state_close(eng.get_state(), ob.get_state(), L, D, 1e-6,
                               skip=JOINT_FIELDS + (("tip_rotation", "origin_tip_rotation", "admittance_delta") if D > 3 else ()))
        eng.close(); ob.close()


def rough_terrain(backend, oracle, n=48, cycles=420):
    """rough_terrain_mode.  Measured tip forces drive Leg::touchdownDetection (model.cpp:712); the
    stepper re-seats its default tip at every swing and stance start (walk_controller.cpp:984-1014, 1058, 1160), moves the swing
    target onto the detected step plane or reaches down by step_depth (:1065-1107), and freezes the second half of the swing
    on ground contact (:1110-1112, 1282-1289); the walk plane, its pose and the swing clearance follow the default tips
    (non-flat plane paths of updateWalkPlane / updateWalkPlanePose).  One cycle from identical state to 1e-11 on every field,
    then a free-running rollout.  The start-up constants are the oracle's (layered workspace)."""
    for cfg, L, D, full, ranges, external in (
            (hexapod_config("tripod_gait", rough_terrain_mode=1, step_depth=0.01), 6, 3, False, False, False),
            (hexapod_config("wave_gait", rough_terrain_mode=1, step_depth=0.02, force_normal_touchdown=1), 6, 3, False, False, False),
            (hexapod_config("ripple_gait", rough_terrain_mode=1, step_depth=0.01), 6, 3, False, True, False),
            (hexapod_config("amble_gait", rough_terrain_mode=1, step_depth=0.01), 6, 3, False, False, True),
            (octopod_config("ripple_gait", rough_terrain_mode=1, step_depth=0.01), 8, 5, True, False, False)):
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        step_planes = np.zeros((n, L, 3), dtype=np.float32)
        cs = CommandStream(n, min_len=40, max_len=160)
        ims = ImuStream(n) if full else None
        rng = np.random.default_rng(17)
        contacts = planes = 0
        force = np.zeros((n, L, 3), dtype=np.float32)

        def contact_forces(st, c):
            # a leg in stance, or late in its swing on a (varying) early touchdown, presses with ~5 N; a lifted tip reads ~0
            for r in range(n):
                for l in range(L):
                    g = st[r].legs[l]
                    early = g.swing_progress > 0.55 + 0.4 * ((r * 7 + l * 3 + c // 50) % 10) / 10.0
                    force[r, l] = (rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(2.0, 8.0)) if (g.step_state != 0 or early) else \
                                  (0.0, 0.0, rng.uniform(0.0, 0.05))

        def range_readings(st, c):
            # tip range sensors (TipState.step_plane): a reading while the tip is near the ground (stance, late swing), none
            # ("UNASSIGNED") high in the swing; the wrench path is off in this variant
            for r in range(n):
                for l in range(L):
                    g = st[r].legs[l]
                    near = g.step_state != 0 or g.swing_progress > 0.5 + 0.3 * ((r * 5 + l + c // 40) % 7) / 7.0
                    step_planes[r, l] = (rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.0, 0.03)) if near else (0.0, 0.0, 2.0e9)
            eng.set_tip_step_planes(step_planes)
            ob.set_tip_step_planes(step_planes)

        def external_requests(c):
            # externally requested swing targets and default tip poses (targetTipPoseCallback, state_controller.cpp:1700-1760):
            # every 40 cycles a third of the legs get a target near their own and a fifth a new default; `transform` stands
            # for the robot's movement since the request, which the reference reads from the tf tree every loop (:703-750)
            st = ob.get_state()
            for r in range(n):
                for l in range(L):
                    g = st[r].legs[l]
                    pick = (r * 11 + l * 5 + c // 40) % 15
                    if pick < 5:
                        tgt = np.array(list(g.target_tip_position)) + rng.uniform(-0.015, 0.015, 3)
                        g.external_target_pose[:] = list(tgt) + [1.0, 0.0, 0.0, 0.0]
                        g.external_target_transform[:] = list(rng.uniform(-0.004, 0.004, 3)) + [1.0, 0.0, 0.0, 0.0]
                        g.external_target_clearance = float(rng.uniform(0.01, 0.03))
                        g.external_target_defined, g.external_target_odom_frame = 1, int(pick % 2)
                    if pick in (7, 8, 9):
                        dft = np.array(list(g.default_tip_position)) + rng.uniform(-0.01, 0.01, 3)
                        g.external_default_pose[:] = list(dft) + [1.0, 0.0, 0.0, 0.0]
                        g.external_default_transform[:] = list(rng.uniform(-0.003, 0.003, 3)) + [1.0, 0.0, 0.0, 0.0]
                        g.external_default_defined = 1
                    elif pick == 10:
                        g.external_default_defined = 0
            ob.set_state(st)

        external_used = 0
        for c in range(cycles):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            if external and c % 40 == 20:
                external_requests(c)
            st = ob.get_state()
            if external:
                external_used += sum(1 for s_ in st for l in range(L) if s_.legs[l].external_target_defined and s_.legs[l].step_state == 0)
            if ranges:
                range_readings(st, c)
            else:
                contact_forces(st, c)
            sample = c % 3 == 2
            fin = None if ranges else force
            if sample:
                eng.set_state(st)
                j = eng.step(cmd, imu, fin)
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64), None if ranges else force.astype(np.float64), threads=4)
            if sample:
                so = ob.get_state()
                assert np.abs(j - ob.joints()).max() <= 1.3e-7, c  # float32 output rounding of angles beyond 2 rad (half an ulp = 1.2e-7)
                assert_state_close(eng.get_state(), so, L, D, 1e-11, vel_tol=1e-9)
                contacts += sum(1 for s in so for l in range(L) if s.legs[l].step_plane_defined and s.legs[l].step_state == 0)
                planes += sum(1 for s in so if abs(s.walk_plane_normal[2] - 1.0) > 1e-9)
        assert contacts > 0 and planes > 0, (contacts, planes)  # swings really ended on contact, the walk plane really tilted
        assert not external or external_used > 0  # swings really ran towards external targets
        print(f"[rough-terrain] {L}x{D}{' (range sensors)' if ranges else ''}{f' (external targets: {external_used} swinging leg-cycles)' if external else ''}: swinging leg-cycles in ground contact {contacts}, robot-cycles with a tilted walk plane {planes}")
        eng.set_state(ob.get_state())
        errs = JointErrors()
        for c in range(150):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            if ranges:
                range_readings(ob.get_state(), cycles + c)
            else:
                contact_forces(ob.get_state(), cycles + c)
            j = eng.step(cmd, imu, None if ranges else force)
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64), None if ranges else force.astype(np.float64), threads=4)
            errs.add(np.abs(j - ob.joints()))
        errs.check(max_fraction=5e-3, label=f"rough terrain {L}x{D} free-running")
        assert_state_close(eng.get_state(), ob.get_state(), L, D, 1e-7, skip=JOINT_FIELDS)
        eng.close(); ob.close()


def sequences(backend, oracle, n=32):
    """PoseController::stepToNewStance (pose_controller.cpp:520; LegPoser::stepToPosition :1571 +
    Leg::applyIK) and packLegs / unpackLegs (:597 / :661; LegPoser::transitionConfiguration :1476), every loop() of each
    sequence against the oracle, from the state a walking batch was stopped in (every robot's legs somewhere else)."""
    for cfg, L, D in ((hexapod_config("ripple_gait"), 6, 3), (octopod_config("tripod_gait"), 8, 5)):
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        cs = CommandStream(n, min_len=30, max_len=90)
        ims = ImuStream(n) if cfg.imu_posing or cfg.inclination_posing else None
        fs = ForceStream(n, L) if cfg.admittance_control else None
        for c in range(130):  # walk for a while: tips, body poses and admittance deltas differ from robot to robot
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                    None if force is None else force.astype(np.float64), threads=4)
        eng.set_state(ob.get_state())
        # stepToNewStance: two groups, one step period each
        loops = 2 * max(1, int(round((1.0 / cfg.step_frequency) / cfg.time_delta)))
        worst, seen = 0.0, set()
        # (exactly one full sequence: LegPoser shares first_iteration_ / master_iteration_count_ between stepToPosition and
        # transitionConfiguration, so the reference itself cannot start packLegs from a half-finished step)
        for k in range(loops):
            j, p = eng.sequence_step("new_stance")
            po = ob.sequence_step("new_stance")
            assert np.array_equal(p, po), (k, p[:8], po[:8])
            worst = max(worst, float(np.abs(j - ob.joints()).max()))
            seen.update(int(v) for v in po)
        assert worst <= 1e-6, worst
        assert max(seen) >= 99 and min(seen) == 0
        d = assert_state_close(eng.get_state(), ob.get_state(), L, D, 1e-9, vel_tol=1e-7,
                               skip=("model_tip_position", "desired_tip_position", "ik_result"))
        print(f"[sequences] {L}x{D} stepToNewStance: {loops} loops, worst joint difference {worst:.2e} rad, "
              f"joint state {d['joint_position']:.2e}")
        # packLegs, then unpackLegs (2 s each)
        for kind in ("pack", "unpack"):
            worst, last = 0.0, 0
            for k in range(400):
                j, p = eng.sequence_step(kind, 2.0)
                po = ob.sequence_step(kind, 2.0)
                assert np.array_equal(p, po), (kind, k, p[:4], po[:4])
                worst = max(worst, float(np.abs(j - ob.joints()).max()))
                last = int(po[0])
                if last == 100:
                    break
            assert last == 100 and k + 1 == int(round(2.0 / cfg.time_delta)), (kind, k, last)
            assert worst <= 1e-6, (kind, worst)
            want = np.array([[getattr(cfg, "joint_packed" if kind == "pack" else "joint_unpacked")[l][jj] for jj in range(D)] for l in range(L)])
            assert np.abs(ob.joints()[0] - want).max() < 1e-12  # the sequence ends on the configured joint positions
            print(f"[sequences] {L}x{D} {kind}Legs: {k + 1} loops, worst joint difference {worst:.2e} rad")
        eng.close(); ob.close()


def execute_sequence(backend, oracle, n=24):
    """PoseController::executeSequence (pose_controller.cpp:145-461), every loop() against the oracle:
    unpack, the FIRST start-up (which generates the transition sequence under its joint-limit safety factor: progress -1),
    a shut-down along the recorded transition poses, and a second start-up (progress 0..100).  The robots start from the
    different states a walking batch stopped in.  Hexapod: free-running, all ~1250 loops.  Octopod: its five-joint legs are
    redundant for a position target, so over thousands of closed-loop IK iterations the null-space component of the joints
    wanders with the reference's normalised cost gradient (DESIGN.md "Reference dynamics") and two correct implementations
    part ways in joint space while their tips and progress values agree; there every loop starts from the oracle's joint state
    (one loop from identical state, to 1e-9), the sequence bookkeeping — step counters, transition poses, targets — still
    being each side's own."""
    for cfg, L, D in ((hexapod_config("tripod_gait"), 6, 3), (octopod_config("wave_gait"), 8, 5)):
        resync = D > 3
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        cs = CommandStream(n, min_len=30, max_len=90)
        ims = ImuStream(n) if cfg.imu_posing or cfg.inclination_posing else None
        fs = ForceStream(n, L) if cfg.admittance_control else None
        for c in range(90):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                    None if force is None else force.astype(np.float64), threads=4)
        eng.set_state(ob.get_state())
        worst = 0.0

        def run(kind, time=0.0, limit=6000):
            """Steps the sequence for the whole batch until every robot has returned 100 (a robot that is through keeps
            returning 100 without moving, as its state machine would stop calling executeSequence)."""
            nonlocal worst
            done = np.zeros(n, dtype=bool)
            seen = set()
            for k in range(limit):
                if resync:
                    eng.set_state(ob.get_state())
                j, p = eng.sequence_step(kind, time)
                po = ob.sequence_step(kind, time)
                assert np.array_equal(p, po), (kind, k, p[:8], po[:8])
                worst = max(worst, float(np.abs(j - ob.joints()).max()))
                seen.update(int(v) for v in po)
                done |= po == 100
                if done.all():
                    return k + 1, seen
            raise AssertionError(f"{kind} did not complete")

        loops_unpack, _ = run("unpack", 2.0)
        loops1, seen1 = run("start_up")
        loops2, seen2 = run("shut_down")
        loops3, seen3 = run("start_up")
        assert -1 in seen1 and -1 not in seen3 and max(seen3 - {100}) >= 90 and -2 not in (seen1 | seen2 | seen3)
        # hexapod, free-running over ~1250 closed-loop IK iterations partly at rest: the joints may sit in the reference's
        # stand-still limit cycle, so the free-running bound applies (observed: float32 rounding only)
        assert worst <= (2e-7 if resync else JointErrors.CHATTER_BOUND), worst
        d = assert_state_close(eng.get_state(), ob.get_state(), L, D, 1e-9 if resync else JointErrors.CHATTER_BOUND, vel_tol=1e-7 if resync else 1.0,
                               skip=("model_tip_position", "desired_tip_position", "ik_result"))
        print(f"[execute-sequence] {L}x{D}: unpack {loops_unpack}, first start-up {loops1}, shut-down {loops2}, second start-up {loops3} "
              f"loops; worst joint difference {worst:.2e} rad, joint state {d['joint_position']:.2e}"
              + (" (every loop from the oracle's joint state)" if resync else " (free-running)"))
        eng.close(); ob.close()


def mixed_precision_statistics(backend, oracle, n=256, cycles=600):
    """Mixed precision over a long rollout: the typical joint error stays far below 1e-6 rad; excursions are bounded by
    the amplitude of the reference's own period-2 joint chatter (~2.3e-3 rad peak to peak), which fp32 state cannot
    phase-track (DESIGN.md "Precision").  This documents the throughput mode; the parity mode is f64."""
    cfg = hexapod_config("tripod_gait")
    ob = oracle.OracleBatch(cfg, n)
    eng = backend.engine(cfg, n, "mixed", startup=ob.startup())
    errs = []
    run_both(eng, ob, cycles, CommandStream(n), per_cycle=lambda c, jg, o: errs.append(np.abs(jg - o.joints()).reshape(n, -1).max(axis=1)))
    errs = np.array(errs)
    assert np.median(errs) < 5e-7
    assert np.quantile(errs, 0.9) < 2e-6
    assert errs.max() < JointErrors.CHATTER_BOUND
    # stepper tips are open-loop and accumulate in double: they stay close regardless
    d = assert_state_close(eng.get_state(), ob.get_state(), 6, 3, 1.0, skip=())
    assert d["tip_position"] < 1e-6 and d["int:phase"] == 0 and d["int:walk_state"] == 0
    eng.close(); ob.close()


def _msg_diff(a, b, skip=()):
    """Largest absolute difference per field between two ctypes message records."""
    out = {}
    for name, _ in a._fields_:
        if name in skip:
            continue
        out[name] = float(np.max(np.abs(np.ctypeslib.as_array(getattr(a, name)) - np.ctypeslib.as_array(getattr(b, name))))) \
            if not isinstance(getattr(a, name), float) else abs(getattr(a, name) - getattr(b, name))
    return out


def wire_formats(backend, oracle, n=24, cycles=260):
    """the JointState / LegState / body records packed from the state planes (shc_pack_messages)
    against the oracle's restatement of the reference's publishers (state_controller.cpp:777-1047), along rollouts of the
    hexapod and of the octopod with admittance + IMU posing + dynamic stiffness.  Fields the engine's state does not
    determine are excluded and named: LegState.auto_pose / poser_tip_pose inside an auto-pose negation window (the
    per-leg auto pose is not kept).  LegState.model_tip_velocity is zero on the reference's wire (its publisher resets it
    before reading it, state_controller.cpp:842-848) and in both records."""
    for cfg, L, D, sensors in ((hexapod_config("ripple_gait"), 6, 3, False), (octopod_config("tripod_gait"), 8, 5, True)):
        ob = oracle.OracleBatch(cfg, n)
        eng = backend.engine(cfg, n, startup=ob.startup())
        cs = CommandStream(n, min_len=30, max_len=120)
        ims = ImuStream(n) if sensors else None
        fs = ForceStream(n, L) if sensors else None
        rng = np.random.default_rng(8)
        lo = np.array([[cfg.joint_min[l][j] for j in range(D)] for l in range(L)])
        hi = np.array([[cfg.joint_max[l][j] for j in range(D)] for l in range(L)])
        measured = (lo + (hi - lo) * rng.uniform(0.1, 0.9, size=(n, L, D))).astype(np.float32)
        worst = {}
        for c in range(cycles):
            cmd = cs.next()
            imu = ims.next(cfg.time_delta) if ims else None
            force = fs.next() if fs else None
            sample = c % 20 == 19
            if sample:  # one cycle from identical state, so that joint chatter does not blur the comparison
                eng.set_state(ob.get_state())
                eng.step(cmd, imu, force)
            ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                    None if force is None else force.astype(np.float64), threads=4)
            if not sample:
                continue
            js, legs, body = eng.pack_messages(0, n, measured)
            for r in range(0, n, 5):
                jo, lo_, bo = ob.messages(r, measured[r].astype(np.float64))
                d = _msg_diff(js[r], jo)
                d.update({"body." + k: v for k, v in _msg_diff(body[r], bo).items()})
                for l in range(L):
                    dl = _msg_diff(legs[r][l], lo_[l])
                    d.update({"leg." + k: max(v, d.get("leg." + k, 0.0)) for k, v in dl.items()})
                for k, v in d.items():
                    worst[k] = max(worst.get(k, 0.0), v)
        print(f"[wire-formats] {L}x{D}: " + ", ".join(f"{k} {v:.1e}" for k, v in sorted(worst.items()) if v > 1e-9))
        for k, v in worst.items():
            tol = 1e-6 if k == "leg.tip_force" else 1e-9
            assert v <= tol, (k, v)
        eng.close(); ob.close()
