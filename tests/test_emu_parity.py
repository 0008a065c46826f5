"""CPU: the control-cycle SOURCE (csrc/shc_cycle.cuh — the code the CUDA kernel is built from) compiled for the host
(tests/emu.py) against the oracle, on the cases the `-m gpu` tests run on the B200, at sizes that keep the CPU suite short.
This is what lets a GPU-less container check every branch of the cycle code; the emulator is test infrastructure, not a
fallback of the engine (the product has no CPU path), and the parity claim itself rests on the `-m gpu` run."""
import os

import pytest

import parity_cases as P
from backends import Backend


@pytest.fixture(scope="module")
def emu():
    import emu as E

    E.build()
    return Backend("emu")


@pytest.mark.parametrize("path", P.GOLDEN_FILES, ids=os.path.basename)
def test_emu_golden_fixture_rollouts(emu, oracle, path):
    P.golden_rollout(emu, oracle, path)


def test_emu_batch_tripod(emu, oracle):
    P.batch_tripod(emu, oracle, n=512, cycles=300)


@pytest.mark.parametrize("gait", ["wave_gait", "amble_gait", "ripple_gait", "tripod_gait"])
def test_emu_gait_sweep(emu, oracle, gait):
    P.gait_sweep(emu, oracle, gait, n=96, cycles=900, cap=1e-3)  # small batch: one window weighs more


def test_emu_octopod_admittance_imu_inclination(emu, oracle):
    P.octopod_full(emu, oracle, n=96, cycles=400)


def test_emu_auto_posing_and_100hz(emu, oracle):
    P.auto_posing_100hz(emu, oracle, n=24, cycles=1200)


def test_emu_auto_posing_own_cycle(emu, oracle):
    P.auto_posing_own_cycle(emu, oracle, n=16)


def test_emu_other_parameter_variants(emu, oracle):
    P.parameter_variants(emu, oracle, n=48, cycles=500)


def test_emu_manual_pose_inputs_and_reset_modes(emu, oracle):
    P.manual_pose_and_reset_modes(emu, oracle)


def test_emu_joint_effort_tip_force(emu, oracle):
    P.joint_effort_tip_force(emu, oracle, n=32)


def test_emu_own_startup_free_running(emu, oracle):
    P.own_startup_free_running(emu, oracle, n=48)


def test_emu_single_step_parity_all_modes(emu, oracle):
    P.single_step_all_modes(emu, oracle, n=64, cycles=260)


def test_emu_single_step_parity_inputs(emu, oracle):
    P.single_step_inputs(emu, oracle, n=32, cycles=200)


def test_emu_tip_orientation(emu, oracle):
    P.tip_orientation(emu, oracle, n=16, cycles=300)


def test_emu_rough_terrain(emu, oracle):
    P.rough_terrain(emu, oracle, n=12, cycles=330)


def test_emu_sequences(emu, oracle):
    P.sequences(emu, oracle, n=16)


def test_emu_execute_sequence(emu, oracle):
    P.execute_sequence(emu, oracle, n=8)


def test_emu_wire_formats(emu, oracle):
    P.wire_formats(emu, oracle)



# ---- the kernel source against THE REFERENCE'S OWN CODE (oracle/_ref), without the restated oracle in between ---------------

def _ref_oracle():
    from oracle import ref_py

    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref is here")
    ref_py.build()
    from backends import RefOracle

    return RefOracle


def test_emu_against_the_reference_itself_gaits(emu):
    R = _ref_oracle()
    P.batch_tripod(emu, R, n=24, cycles=300)
    P.gait_sweep(emu, R, "ripple_gait", n=16, cycles=600, cap=2e-3)


def test_emu_against_the_reference_itself_octopod_and_posing(emu):
    R = _ref_oracle()
    P.octopod_full(emu, R, n=8, cycles=300)
    P.manual_pose_and_reset_modes(emu, R, n=8)
    P.auto_posing_100hz(emu, R, gaits=("wave_gait",), n=6, cycles=800)


def test_emu_imported_state_continues_bit_for_bit(emu):
    """A second engine that imports the first one's state record ends up in the same state, byte for byte, after the next
    cycles (the `-m gpu` twin is test_state_range_and_limit_maps): the legs' saved walk planes must not depend on which
    cycles a leg happened to save them in (sign of zero of the fitted plane normal included)."""
    from syropod_highlevel_controller_b200.config import hexapod_config
    from syropod_highlevel_controller_b200.streams import CommandStream

    cfg, n = hexapod_config("tripod_gait"), 64
    a = emu.engine(cfg, n)
    cs = CommandStream(n, min_len=20, max_len=60)
    for c in range(90):
        a.step(cs.next())
    b = emu.engine(cfg, n)
    b.set_state(a.get_state())
    for c in range(6):
        cmd = cs.next()
        ja, jb = a.step(cmd), b.step(cmd)
        assert (ja == jb).all()
        assert bytes(a.get_state()) == bytes(b.get_state()), c
    a.close(); b.close()


def _gait_change_case(backend, switch, n=4, model="hexapod"):
    """StateController::changeGait of the reference (gaitSelectionCallback; the robot is stopped, then the step cycle, limit
    maps, phase offsets and auto posers are regenerated, state_controller.cpp:513-540) against the engine's way of doing
    it: a new engine for the new gait that carries the old one's state (`switch(engine, new_cfg)`).  Tripod -> wave with
    auto posing, walking before and after."""
    import numpy as np

    from gpu_common import JOINT_FIELDS, JointErrors, assert_state_close
    from oracle import ref_py
    from syropod_highlevel_controller_b200.config import ShcRobotState, hexapod_config, octopod_config
    from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref is here")
    ref_py.build()
    if model == "hexapod":
        cfg_a, cfg_b = hexapod_config("tripod_gait", auto_posing=1), hexapod_config("wave_gait", auto_posing=1)
    else:  # 8 x 5 with admittance, IMU and inclination posing
        cfg_a, cfg_b = octopod_config("tripod_gait"), octopod_config("ripple_gait")
    L, D = cfg_a.leg_count, cfg_a.joint_count
    ims = ImuStream(n) if cfg_a.imu_posing or cfg_a.inclination_posing else None
    fs = ForceStream(n, L) if cfg_a.admittance_control else None
    refs = [ref_py.RefRobot(cfg_a) for _ in range(n)]
    eng = backend.engine(cfg_a, n, startup=refs[0].startup())
    cs = CommandStream(n, min_len=60, max_len=200)
    errs = JointErrors()

    def ref_states():
        arr = (ShcRobotState * n)()
        for i, r in enumerate(refs):
            arr[i] = r.get_state()
        return arr

    def step(e, cmd):
        imu = ims.next(cfg_a.time_delta) if ims else None
        f = fs.next() if fs else None
        j = e.step(cmd, imu, f)
        for i, r in enumerate(refs):
            r.step(cmd[i].astype(np.float64), None if imu is None else imu[i].astype(np.float64), None if f is None else f[i].astype(np.float64))
        errs.add(np.abs(j - np.stack([r.joints() for r in refs])))

    for c in range(300):
        step(eng, cs.next())
    # The batch is brought to rest first (zero velocity input, as changeGait itself would force), then the gait selection
    # arrives: the next loop() of every reference robot runs changeGait and updates no tips (state_controller.cpp:391-395,
    # 427) — the engine takes no cycle for it.
    for k in range(3000):
        if all(s.walk_state == 3 for s in ref_states()):
            break
        step(eng, np.zeros((n, 3), dtype=np.float32))
    for i, r in enumerate(refs):
        r.select_gait(cfg_b)
        imu = ims.next(cfg_a.time_delta) if ims else None  # the pose and admittance stages of that loop() still run
        f = fs.next() if fs else None
        r.step(np.zeros(3), None if imu is None else imu[i].astype(np.float64), None if f is None else f[i].astype(np.float64))
    assert not any(r.gait_change_pending for r in refs)
    eng2 = switch(eng, cfg_b, False)  # the state carried into an engine for the new gait
    eng.close()
    su_ref, su = refs[0].startup(), eng2.startup()
    assert (su.period, su.swing_period, su.stance_period) == (su_ref.period, su_ref.swing_period, su_ref.stance_period)
    assert list(su.phase_offsets)[:L] == list(su_ref.phase_offsets)[:L]
    for f in ("max_linear_speed", "max_angular_speed", "max_linear_acceleration", "max_angular_acceleration", "walkspace"):
        assert np.abs(np.array(list(getattr(su, f))) - np.array(list(getattr(su_ref, f)))).max() < 1e-7, f
    assert_state_close(eng2.get_state(), ref_states(), L, D, 1e-9, skip=JOINT_FIELDS)
    for c in range(700):
        step(eng2, cs.next())
    errs.check(max_fraction=2e-3, label=f"{model}: gait change against the reference's changeGait")
    st = ref_states()
    assert_state_close(eng2.get_state(), st, L, D, 1e-8, skip=JOINT_FIELDS)
    assert {s.walk_state for s in st} & {0, 1, 2}  # walking again under the new gait
    eng2.close()
    for r in refs:
        r.close()


def _emu_switch(emu):
    def switch(eng, cfg, keep_pose_cycle):
        from syropod_highlevel_controller_b200.engine import compute_startup

        su = compute_startup(cfg)  # the engine's own constants for the new parameters
        if keep_pose_cycle:
            old = eng.startup()
            su.pose_phase_length, su.pose_normaliser = old.pose_phase_length, old.pose_normaliser
        new = emu.engine(cfg, eng.n, startup=su)
        new.set_state(eng.get_state())
        return new

    return switch


def test_emu_gait_change_against_the_reference(emu):
    # (model="octopod" — IMU posing and admittance on — is NOT exact: the reference's switching loop() still advances the
    # IMU PID and the admittance model by one step, while the engine's switch takes no cycle at all: imu_pose is 4.5e-4
    # rad apart right after the switch.  Documented at shc_clone_reconfigured.)
    _gait_change_case(emu, _emu_switch(emu), model="hexapod")


PARAMETER_CHANGES = [
    # (model, base overrides, changed parameter, at rest?, cycles the reference lags, first command after the change)
    ("hexapod", {}, dict(swing_height=0.035), False, 0, None),
    ("hexapod", {}, dict(swing_width=0.02), False, 0, None),
    ("hexapod", dict(rough_terrain_mode=1, step_depth=0.01), dict(step_depth=0.02), False, 0, None),
    ("hexapod", {}, dict(stance_span_modifier=0.3), True, 0, None),
    ("hexapod", {}, dict(step_frequency=1.5), True, 0, (0.7, 0.2, 0.3)),
    # with auto posing the reference keeps the OLD pose cycle length (setAutoPoseParams is re-run by changeGait only)
    ("hexapod", dict(auto_posing=1), dict(step_frequency=1.5), True, 0, (0.7, 0.2, 0.3)),
    ("octopod", {}, dict(virtual_stiffness=20.0), False, 1, None),
    ("octopod", {}, dict(force_gain=0.3), False, 1, None),
    ("octopod", {}, dict(virtual_mass=15.0), False, 1, None),
    ("octopod", {}, dict(virtual_damping_ratio=1.2), False, 1, None),
]


def _parameter_change_case(backend, switch, model, base, change, at_rest, lag, first_cmd, n=3):
    """StateController::adjustParameter of the reference (a dynamic_reconfigure request through dynamicParameterCallback,
    applied inside the next loop(), state_controller.cpp:451-508, 1465-1548) against the engine's way of changing a
    batch-wide parameter: a new engine for the new value that carries the state (`switch`).  Timing as in the reference:
      * walk parameters (swing height / width, step depth; stance span modifier at rest) act in the loop() that applies them;
      * admittance parameters act one loop later (updateAdmittance has already run when runningState() applies them);
      * step_frequency at rest: regenerated step cycle, limit maps and phase offsets.  The reference defers this while a
        SIGNED, component-wise velocity test fails (state_controller.cpp:489-491: a robot commanded backwards keeps its old
        step cycle with the new speed limits); the case commands forwards.  Re-phasing WALKING legs (LegStepper::updatePhase)
        is not supported by the engine: the batch is stopped first, as for a gait change."""
    import numpy as np

    from gpu_common import JOINT_FIELDS, JointErrors, assert_state_close
    from oracle import ref_py
    from syropod_highlevel_controller_b200.config import ShcRobotState, hexapod_config, octopod_config
    from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref is here")
    ref_py.build()
    make = hexapod_config if model == "hexapod" else octopod_config
    cfg_a, cfg_b = make("tripod_gait", **base), make("tripod_gait", **dict(base, **change))
    L, D = cfg_a.leg_count, cfg_a.joint_count
    refs = [ref_py.RefRobot(cfg_a) for _ in range(n)]
    eng = backend.engine(cfg_a, n, startup=refs[0].startup())
    cs = CommandStream(n, min_len=60, max_len=200)
    ims = ImuStream(n) if cfg_a.imu_posing or cfg_a.inclination_posing else None
    fs = ForceStream(n, L) if cfg_a.admittance_control else None
    errs = JointErrors()

    def ref_states():
        arr = (ShcRobotState * n)()
        for i, r in enumerate(refs):
            arr[i] = r.get_state()
        return arr

    def step(e, cmd):
        imu = ims.next(cfg_a.time_delta) if ims else None
        f = fs.next() if fs else None
        j = e.step(cmd, imu, f)
        for i, r in enumerate(refs):
            r.step(cmd[i].astype(np.float64), None if imu is None else imu[i].astype(np.float64), None if f is None else f[i].astype(np.float64))
        errs.add(np.abs(j - np.stack([r.joints() for r in refs])))

    for c in range(200):
        step(eng, cs.next())
    if at_rest:
        for k in range(3000):
            if all(s.walk_state == 3 for s in ref_states()):
                break
            step(eng, np.zeros((n, 3), dtype=np.float32))
    for r in refs:
        assert r.adjust_parameter(cfg_b)
    for k in range(lag):
        step(eng, cs.next())
    eng2 = switch(eng, cfg_b, True)  # adjustParameter: the auto-pose cycle is not regenerated
    eng.close()
    if first_cmd is not None:
        for c in range(40):
            step(eng2, np.tile(np.array(first_cmd, dtype=np.float32), (n, 1)))
    for c in range(500):
        step(eng2, cs.next())
    assert not any(r.parameter_adjust_pending for r in refs)
    errs.check(max_fraction=2e-3, label=f"{model}: {change} against the reference's adjustParameter")
    assert_state_close(eng2.get_state(), ref_states(), L, D, 1e-8, skip=JOINT_FIELDS)
    eng2.close()
    for r in refs:
        r.close()


@pytest.mark.parametrize("model,base,change,at_rest,lag,first_cmd", PARAMETER_CHANGES, ids=[",".join(c[2]) + ("+autopose" if c[1].get("auto_posing") else "") for c in PARAMETER_CHANGES])
def test_emu_parameter_change_against_the_reference(emu, model, base, change, at_rest, lag, first_cmd):
    _parameter_change_case(emu, _emu_switch(emu), model, base, change, at_rest, lag, first_cmd)
