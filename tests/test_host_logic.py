"""CPU: the product's host side — the C-ABI library loads and exports every symbol include/shc_b200.h declares, the
ctypes mirrors match the C structs, the engine's own start-up restatement agrees with the oracle, the shared
kinematics code agrees with the oracle, and there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from syropod_highlevel_controller_b200 import config as cfgmod
from syropod_highlevel_controller_b200.config import (ShcConfig, ShcRobotState, ShcStartup, hexapod_config,
                                                      octopod_config)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(shc_lib):
    header = open(os.path.join(ROOT, "include", "shc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(shc_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(shc_lib, n), f"libshc_b200.so does not export {n}"


def test_struct_sizes_match(shc_lib, oracle):
    assert shc_lib.shc_sizeof_config() == C.sizeof(ShcConfig) == oracle.lib().shc_oracle_config_size()
    assert shc_lib.shc_sizeof_startup() == C.sizeof(ShcStartup) == oracle.lib().shc_oracle_startup_size()
    assert shc_lib.shc_sizeof_robot_state() == C.sizeof(ShcRobotState) == oracle.lib().shc_oracle_state_record_size()


def test_no_cpu_fallback(shc_lib):
    """Without a CUDA device the engine refuses to exist (SHC_E_CUDA); with one this test is skipped."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from syropod_highlevel_controller_b200.engine import Engine, ShcError

    with pytest.raises(ShcError):
        Engine(hexapod_config(), 4)
    h = C.c_void_p()
    cfg = hexapod_config()
    rc = shc_lib.shc_create(C.byref(cfg), None, 4, 0, 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in shc_lib.shc_last_error()


def test_unsupported_configs_are_rejected(shc_lib):
    from syropod_highlevel_controller_b200.engine import ShcError, compute_startup

    with pytest.raises(ShcError):  # stance span change over the layered workspace is not built
        compute_startup(hexapod_config(rough_terrain_mode=1, stance_span_modifier=0.3))
    bad = hexapod_config()
    bad.leg_count = 9
    with pytest.raises(ShcError):
        compute_startup(bad)


@pytest.mark.parametrize("make,gait,dt", [(hexapod_config, "tripod_gait", 0.02), (hexapod_config, "wave_gait", 0.01),
                                          (hexapod_config, "ripple_gait", 0.02), (hexapod_config, "amble_gait", 0.02),
                                          (octopod_config, "tripod_gait", 0.02)])
def test_engine_startup_matches_oracle(shc_lib, oracle, make, gait, dt):
    """generateStepCycle / directStartup / generateWorkspaces / generateWalkspace / generateLimits, engine vs oracle.

    Limit tables, walkspace and workspace agree to rounding; see the note on the default-stance joints below."""
    from syropod_highlevel_controller_b200.engine import compute_startup

    cfg = make(gait, dt)
    su = compute_startup(cfg)
    so = oracle.OracleBatch(cfg, 1).startup()
    for f in ("period", "swing_period", "stance_period", "stance_end", "swing_start", "swing_end", "stance_start",
              "pose_phase_length", "pose_normaliser", "auto_pose_reference_leg"):
        assert getattr(su, f) == getattr(so, f), f
    assert list(su.phase_offsets) == list(so.phase_offsets)
    assert su.step_frequency == so.step_frequency
    L, D = cfg.leg_count, cfg.joint_count
    a = np.array([list(r)[:D] for r in su.default_joint][:L]); b = np.array([list(r)[:D] for r in so.default_joint][:L])
    # Default-stance joints: the simulated start-up ends inside the reference's stand-still limit cycle (joints
    # alternate by ~1e-3 rad every cycle, DESIGN.md "Reference dynamics"), whose phase is decided by rounding-level
    # differences — so two correct implementations agree only to that amplitude, and their tips to ~0.1 mm.
    assert np.abs(a - b).max() < 5e-3
    for l in range(L):
        assert np.abs(oracle.fk(cfg, l, a[l]) - oracle.fk(cfg, l, b[l])).max() < 5e-4
    if make is hexapod_config and dt == 0.02:
        assert np.abs(a - b).max() < 1e-7  # the shipped configuration happens to land on the same phase
    assert np.abs(np.array([list(r) for r in su.workspace]) - np.array([list(r) for r in so.workspace])).max() < 1e-7
    for f in ("walkspace", "max_linear_speed", "max_angular_speed", "max_linear_acceleration", "max_angular_acceleration"):
        assert np.abs(np.array(list(getattr(su, f))) - np.array(list(getattr(so, f)))).max() < 1e-7, f


def test_engine_startup_rough_terrain_matches_oracle(shc_lib, oracle):
    """rough_terrain_mode: the start-up reads the layered workspace (Leg::generateWorkspace in full, model.cpp:309-510)
    through Leg::getWorkplane at the height of the default tips (model.cpp:514); walkspace and limit tables vs the oracle.
    (The oracle's start-up record has no simple workspace in this mode: shc_startup.workspace is the engine's interpolated
    workplane, checked against the layered workspace of the oracle directly.)"""
    from syropod_highlevel_controller_b200.engine import compute_startup

    for cfg in (hexapod_config("tripod_gait", rough_terrain_mode=1), octopod_config("tripod_gait", rough_terrain_mode=1)):
        su = compute_startup(cfg)
        so = oracle.OracleBatch(cfg, 1).startup()
        for f in ("walkspace", "max_linear_speed", "max_angular_speed", "max_linear_acceleration", "max_angular_acceleration"):
            assert np.abs(np.array(list(getattr(su, f))) - np.array(list(getattr(so, f)))).max() < 1e-6, f
        assert max(su.walkspace) > 0.01
        for leg in (0, cfg.leg_count - 1):
            h, r = oracle.workspace(cfg, leg, True, 24)
            k = int(np.argmin(np.abs(h)))  # the plane the search put at height ~0
            assert abs(h[k]) < 1e-9
            assert np.abs(np.array(list(su.workspace[leg])) - r[k]).max() < 1e-6


def test_shared_kinematics_match_oracle(shc_lib, oracle):
    """The Leg::applyIK routine shared by the kernels and the start-up code, evaluated on the host, vs the oracle."""
    from syropod_highlevel_controller_b200.engine import host_apply_ik

    L_ = oracle.lib()
    dp = C.POINTER(C.c_double)
    L_.shc_oracle_apply_ik.restype = C.c_double
    L_.shc_oracle_apply_ik.argtypes = [C.POINTER(ShcConfig), C.c_int, dp, dp, dp, C.c_int, dp]
    rng = np.random.default_rng(5)
    for cfg in (hexapod_config(), octopod_config()):
        D = cfg.joint_count
        for leg in range(cfg.leg_count):
            lo = np.array([cfg.joint_min[leg][j] for j in range(D)]); hi = np.array([cfg.joint_max[leg][j] for j in range(D)])
            for trial in range(8):
                q = lo + (hi - lo) * rng.uniform(0.05, 0.95, D)
                qd = rng.normal(0, 1.0, D)
                des = oracle.fk(cfg, leg, q) + rng.normal(0, 0.004, 3)
                for sim in (0, 1):
                    qo, qdo, tipo = q.copy(), qd.copy(), np.empty(3)
                    ro = L_.shc_oracle_apply_ik(C.byref(cfg), leg, qo.ctypes.data_as(dp), qdo.ctypes.data_as(dp),
                                                np.ascontiguousarray(des).ctypes.data_as(dp), sim, tipo.ctypes.data_as(dp))
                    qe, qde, tipe, re_ = host_apply_ik(cfg, leg, q, qd, des, bool(sim))
                    assert np.abs(qe - qo).max() < 1e-12 and np.abs(qde - qdo).max() < 1e-10
                    assert np.abs(tipe - tipo).max() < 1e-13
                    assert re_ == pytest.approx(ro, abs=1e-12)


def test_reference_yaml_loader_matches_builtin():
    """load_reference_yaml on the reference's own config files reproduces hexapod_config() (skipped on the GPU box,
    where /root/reference does not exist)."""
    base = "/root/reference/config"
    if not os.path.isdir(base):
        pytest.skip("reference checkout not present")
    for gait in ("tripod_gait", "wave_gait", "ripple_gait", "amble_gait"):
        a = cfgmod.load_reference_yaml(f"{base}/default.yaml", f"{base}/gait.yaml", f"{base}/auto_pose.yaml", gait=gait)
        b = hexapod_config(gait)
        assert cfgmod.config_to_dict(a) == cfgmod.config_to_dict(b), gait


def test_command_stream_is_deterministic_and_shard_independent():
    from syropod_highlevel_controller_b200.streams import CommandStream

    full = CommandStream(64)
    lo, hi = CommandStream(32, robot_offset=0), CommandStream(32, robot_offset=32)
    zeros = 0
    for c in range(500):
        f = full.next()
        assert np.array_equal(f[:32], lo.next()) and np.array_equal(f[32:], hi.next())
        assert np.all(np.hypot(f[:, 0], f[:, 1]) <= 1.0 + 1e-6) and np.all(np.abs(f[:, 2]) <= 1.0)
        zeros += np.sum(np.all(f == 0, axis=1))
    assert 0.1 < zeros / (500 * 64) < 0.3  # about 20 % of the segments are all-zero


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm) runs without a GPU and prints the contract's JSON line."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2
    from oracle import ref_py

    # the reference's own code where oracle/_ref exists (with the restated oracle's figure beside it), else the port
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_py.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    if ref_py.available():
        assert line["cpu_baseline"]["port_value"] > 0 and line["config"]["same_config"] is False
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
