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


def _gait_change_case(backend, switch, n=4):
    """StateController::changeGait of the reference (gaitSelectionCallback; the robot is stopped, then the step cycle, limit
    maps, phase offsets and auto posers are regenerated, state_controller.cpp:513-540) against the engine's way of doing
    it: a new engine for the new gait that carries the old one's state (`switch(engine, new_cfg)`).  Tripod -> wave with
    auto posing, walking before and after."""
    import numpy as np

    from gpu_common import JOINT_FIELDS, JointErrors, assert_state_close
    from oracle import ref_py
    from syropod_highlevel_controller_b200.config import ShcRobotState, hexapod_config
    from syropod_highlevel_controller_b200.streams import CommandStream

    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref is here")
    ref_py.build()
    cfg_a, cfg_b = hexapod_config("tripod_gait", auto_posing=1), hexapod_config("wave_gait", auto_posing=1)
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
        j = e.step(cmd)
        for i, r in enumerate(refs):
            r.step(cmd[i].astype(np.float64))
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
    for r in refs:
        r.select_gait(cfg_b)
        r.step(np.zeros(3))
    assert not any(r.gait_change_pending for r in refs)
    eng2 = switch(eng, cfg_b)  # the state carried into an engine for the new gait
    eng.close()
    su_ref, su = refs[0].startup(), eng2.startup()
    assert (su.period, su.swing_period, su.stance_period) == (su_ref.period, su_ref.swing_period, su_ref.stance_period)
    assert list(su.phase_offsets)[:6] == list(su_ref.phase_offsets)[:6]
    for f in ("max_linear_speed", "max_angular_speed", "max_linear_acceleration", "max_angular_acceleration", "walkspace"):
        assert np.abs(np.array(list(getattr(su, f))) - np.array(list(getattr(su_ref, f)))).max() < 1e-7, f
    assert_state_close(eng2.get_state(), ref_states(), 6, 3, 1e-9, skip=JOINT_FIELDS)
    for c in range(700):
        step(eng2, cs.next())
    errs.check(max_fraction=2e-3, label="gait change tripod -> wave against the reference's changeGait")
    st = ref_states()
    assert_state_close(eng2.get_state(), st, 6, 3, 1e-8, skip=JOINT_FIELDS)
    assert {s.walk_state for s in st} & {0, 1, 2}  # walking again under the new gait
    eng2.close()
    for r in refs:
        r.close()


def test_emu_gait_change_against_the_reference(emu):
    def switch(eng, cfg):
        new = emu.engine(cfg, eng.n, startup=None)  # the engine's own constants for the new gait
        new.set_state(eng.get_state())
        return new

    _gait_change_case(emu, switch)
