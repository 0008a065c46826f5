import sys, os, subprocess, glob
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(root, "gpurun_in_lib_*.so"))):
    env = dict(os.environ, SHC_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", """
import sys; sys.path.insert(0, %r)
import torch
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from tools.dev_gpu_check import timing
for prec in ('f64','mixed'):
    timing(hexapod_config(), 131072, prec, tag='hex')
""" % root], env=env, capture_output=True, text=True)
    print(os.path.basename(lib)); print(out.stdout.strip()); print(out.stderr.strip()[-300:])
