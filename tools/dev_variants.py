import sys, os, subprocess, glob, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(root, "gpurun_in_lib_*.so"))):
    env = dict(os.environ, SHC_B200_LIB=lib)
    for prec in ("f64", "mixed"):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "30", "--no-cpu-baseline", "--precision", prec],
                             env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(os.path.basename(lib), prec, "value %.4g  %.1f us/step  frac %.3f" % (d["value"], d["ms_per_step"] * 1e3, d["roofline"]["frac"]))
        except Exception as ex:
            print(os.path.basename(lib), prec, "FAILED", out.stderr[-300:])
