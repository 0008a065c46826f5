import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from tools.dev_gpu_check import timing
print(torch.cuda.get_device_name(0))
for prec in ("f64", "mixed"):
    for n in (4096, 65536, 131072, 262144, 1048576):
        timing(hexapod_config(), n, prec, tag="hex")
    timing(octopod_config(), 262144, prec, tag="octo")
