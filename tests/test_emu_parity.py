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
