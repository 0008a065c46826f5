"""GPU: the CUDA control-cycle path (through the C-ABI) against the CPU oracle on identical inputs — the parity tests
proper.  The cases live in tests/parity_cases.py (shared with the host-emulator CPU tests); the reference cannot be
built on the GPU box (nor in the build container: ROS/Eigen/Boost absent), so the oracle — pinned by
tests/test_oracle_*.py and the committed fixtures under tests/golden/ — is the checker."""
import os

import pytest

import parity_cases as P
from backends import Backend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(shc_lib):
    return Backend("gpu")


@pytest.mark.parametrize("path", P.GOLDEN_FILES, ids=os.path.basename)
def test_golden_fixture_rollouts(gpu, oracle, path):
    P.golden_rollout(gpu, oracle, path)


def test_config2_batch4096_tripod(gpu, oracle):
    P.batch_tripod(gpu, oracle, n=4096, cycles=300)


@pytest.mark.parametrize("gait", ["wave_gait", "amble_gait", "ripple_gait", "tripod_gait"])
def test_config3_gait_sweep(gpu, oracle, gait):
    P.gait_sweep(gpu, oracle, gait, n=512, cycles=900)


def test_config4_octopod_admittance_imu_inclination(gpu, oracle):
    P.octopod_full(gpu, oracle, n=512, cycles=400)


def test_auto_posing_and_100hz(gpu, oracle):
    P.auto_posing_100hz(gpu, oracle)


def test_auto_posing_own_cycle(gpu, oracle):
    P.auto_posing_own_cycle(gpu, oracle)


def test_other_parameter_variants(gpu, oracle):
    P.parameter_variants(gpu, oracle)


def test_manual_pose_inputs_and_reset_modes(gpu, oracle):
    P.manual_pose_and_reset_modes(gpu, oracle)


def test_joint_effort_tip_force(gpu, oracle):
    P.joint_effort_tip_force(gpu, oracle)


def test_own_startup_free_running(gpu, oracle):
    P.own_startup_free_running(gpu, oracle)


def test_single_step_parity_all_modes(gpu, oracle):
    P.single_step_all_modes(gpu, oracle)


def test_single_step_parity_inputs(gpu, oracle):
    P.single_step_inputs(gpu, oracle)


def test_mixed_precision_rollout_statistics(gpu, oracle):
    P.mixed_precision_statistics(gpu, oracle)


def test_tip_orientation(gpu, oracle):
    P.tip_orientation(gpu, oracle, n=48, cycles=400)


def test_rough_terrain(gpu, oracle):
    P.rough_terrain(gpu, oracle, n=48, cycles=420)


def test_sequences(gpu, oracle):
    P.sequences(gpu, oracle, n=96)


def test_execute_sequence(gpu, oracle):
    P.execute_sequence(gpu, oracle, n=48)


def test_wire_formats(gpu, oracle):
    P.wire_formats(gpu, oracle)

