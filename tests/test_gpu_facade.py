"""GPU: the C++ façade (include/shc_facade.hpp) driven with the reference's loop()/runningState() call sequence by a
compiled harness (tests/cpp/facade_harness.cpp) reproduces the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.streams import CommandStream

from gpu_common import JointErrors

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_harness_matches_oracle(shc_lib, oracle, tmp_path):
    from syropod_highlevel_controller_b200.build import LIB_PATH, PKG_DIR

    exe = str(tmp_path / "facade_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_harness.cpp"), "-o", exe,
                           "-L", PKG_DIR, "-l:libshc_b200.so", "-Wl,-rpath," + PKG_DIR])
    cfg = hexapod_config("tripod_gait")
    n, cycles = 48, 300
    ob = oracle.OracleBatch(cfg, n)
    su = ob.startup()
    (tmp_path / "cfg.bin").write_bytes(bytes(cfg))
    (tmp_path / "su.bin").write_bytes(bytes(su))
    cs = CommandStream(n, min_len=40, max_len=120)
    cmds = np.stack([cs.next() for _ in range(cycles)]).astype(np.float32)
    cmds.tofile(tmp_path / "cmd.bin")
    out = subprocess.run([exe, str(tmp_path / "cfg.bin"), str(tmp_path / "su.bin"), str(n), str(cycles),
                          str(tmp_path / "cmd.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"{cycles} engine steps" in out.stdout  # one fused launch per control cycle for the whole batch
    rows = np.fromfile(tmp_path / "out.bin", dtype=np.float64).reshape(-1, n * 18 + n)
    seq_rows, rows = rows[cycles:], rows[:cycles]
    errs = JointErrors()
    for c in range(cycles):
        ob.step(cmds[c].astype(np.float64), threads=8)
        errs.add(np.abs(rows[c, : n * 18].reshape(n, 6, 3) - ob.joints()))
        ws = np.array([s.walk_state for s in ob.get_state()])
        assert np.array_equal(rows[c, n * 18:].astype(int), ws)
    errs.check(max_fraction=1e-3)
    # the sequence tail: stepToNewStance (two groups x one step period), then packLegs(2 s), one device step per loop
    num = int(round((1.0 / cfg.step_frequency) / cfg.time_delta))
    pack = int(round(2.0 / cfg.time_delta))
    assert len(seq_rows) == 2 * num + pack
    assert f"{2 * num + pack} sequence loops in {2 * num + pack} device steps" in out.stdout
    errs = JointErrors()  # the sequences start from the free-running joint state: same chatter allowance as above
    for k, row in enumerate(seq_rows):
        po = ob.sequence_step("new_stance") if k < 2 * num else ob.sequence_step("pack", 2.0)
        assert np.array_equal(row[n * 18:].astype(int), po), k
        errs.add(np.abs(row[: n * 18].reshape(n, 6, 3) - ob.joints()))
    errs.check(max_fraction=5e-3, label="facade sequences")
