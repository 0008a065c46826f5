import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine
from oracle import oracle_py as O
from tools.dev_gpu_check import state_diff
cfg = hexapod_config()
for n in (1, 2, 64):
    ob = O.OracleBatch(cfg, n); eng = Engine(cfg, n, precision="f64", startup=ob.startup())
    sd = state_diff(eng.get_state(), ob.get_state(), 6, 3)
    print("n", n, "initial state diffs", {k: v for k, v in sd.items() if v > 0})
    for c in range(4):
        cmd = np.tile(np.array([[0.5, 0, 0]], dtype=np.float32), (n, 1))
        j = eng.step(torch.from_numpy(cmd).cuda()).cpu().numpy().astype(np.float64)
        ob.step(cmd.astype(np.float64))
        sd = state_diff(eng.get_state(), ob.get_state(), 6, 3)
        print("  c", c, "dq", np.abs(j - ob.joints()).max(), {k: float(f"{v:.2e}") for k, v in sd.items() if v > 1e-9})
    eng.close(); ob.close()
