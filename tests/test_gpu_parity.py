"""GPU: the CUDA control-cycle path (through the C-ABI) against the CPU oracle on identical inputs — the parity tests
proper.  The cases live in tests/parity_cases.py (shared with the host-emulator CPU tests).  The oracle is the checker: it
is pinned to the reference's own code (tests/test_reference_pin.py), the fixtures under tests/golden/ are the reference's
outputs, and the last two tests compare the CUDA path with the reference's own code directly (oracle/_ref is compiled
from /root/reference in the build container; the prebuilt library travels to the GPU box)."""
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



# ---- the CUDA path against THE REFERENCE'S OWN CODE (the prebuilt oracle/_ref travels to the GPU box) ----------------------

def _ref_oracle():
    from oracle import ref_py

    if not ref_py.available():
        pytest.skip("oracle/_ref (the reference's own sources, compiled where /root/reference exists) is not here")
    ref_py.build()
    from backends import RefOracle

    return RefOracle


def test_gpu_against_the_reference_itself_gaits(gpu):
    R = _ref_oracle()
    P.batch_tripod(gpu, R, n=48, cycles=300)
    for gait in ("ripple_gait", "wave_gait", "amble_gait"):
        P.gait_sweep(gpu, R, gait, n=24, cycles=600, cap=2e-3)


def test_gpu_against_the_reference_itself_octopod_and_posing(gpu):
    R = _ref_oracle()
    P.octopod_full(gpu, R, n=12, cycles=300)
    P.manual_pose_and_reset_modes(gpu, R, n=12)
    P.auto_posing_100hz(gpu, R, gaits=("wave_gait", "tripod_gait"), n=8, cycles=800)
    P.joint_effort_tip_force(gpu, R, n=12)
