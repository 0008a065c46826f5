"""Kernel tuning (run under gpurun): for every gpurun_in_lib_<name>.so built by tools/dev_build_variants.sh, a parity smoke
check against the oracle and one bench line per precision in DEV_PRECISIONS (default f64)."""
import glob
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
precisions = os.environ.get("DEV_PRECISIONS", "f64").split(",")
for lib in sorted(glob.glob(os.path.join(root, "gpurun_in_lib_*.so"))):
    env = dict(os.environ, SHC_B200_LIB=lib)
    sm = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], env=env, capture_output=True, text=True, cwd=root)
    ok = "smoke ok" in sm.stdout
    print(os.path.basename(lib), "smoke", "ok" if ok else "FAILED " + (sm.stdout + sm.stderr)[-300:], flush=True)
    for prec in precisions:
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "30", "--no-cpu-baseline", "--precision", prec] + sys.argv[1:],
                             env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(os.path.basename(lib), prec, "value %.4g  %.1f us/step  kernel %.1f us  frac %.3f" %
                  (d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"]), flush=True)
        except Exception:
            print(os.path.basename(lib), prec, "FAILED", out.stderr[-300:], flush=True)
