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

