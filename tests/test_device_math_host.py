"""CPU: the device-only math routines (bounded-argument sincos_, refined-seed rsqrt_ of csrc/shc_math.cuh) replayed on the
host with exact fma() and compared with long double: the accuracy DESIGN.md quotes (<= 1.6 ulp / < 1 ulp).  The kernels
themselves are checked against the oracle on the GPU (tests/test_gpu_parity.py); this guards the constants."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sincos_and_rsqrt_sequences_are_accurate(tmp_path):
    exe = str(tmp_path / "device_math_check")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-o", exe, os.path.join(ROOT, "tests", "cpp", "device_math_check.c"), "-lm"])
    es, ec, er = map(float, subprocess.check_output([exe], text=True).split())
    assert es <= 1.7 and ec <= 1.7, (es, ec)
    assert er < 1.0, er


def test_constants_match_the_device_source():
    """Every floating-point literal of the host replay appears verbatim in shc_math.cuh."""
    dev = open(os.path.join(ROOT, "syropod_highlevel_controller_b200", "csrc", "shc_math.cuh")).read()
    chk = open(os.path.join(ROOT, "tests", "cpp", "device_math_check.c")).read()
    body = chk[chk.index("static void sincos_dev"):chk.index("static double ulp_err")]
    lits = set(re.findall(r"-?\d+\.\d+(?:e[+-]?\d+)?", body))
    assert len(lits) >= 15
    for lit in lits:
        assert lit.lstrip("-") in dev, lit
