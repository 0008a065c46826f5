import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from syropod_highlevel_controller_b200.engine import Engine
from oracle import oracle_py as O

def run(cfg, cmdfun, cycles, precision, tag):
    n = 1
    ob = O.OracleBatch(cfg, n); eng = Engine(cfg, n, precision=precision, startup=ob.startup())
    hist = []
    for c in range(cycles):
        cmd = np.array([cmdfun(c)], dtype=np.float32)
        j = eng.step(torch.from_numpy(cmd).cuda()).cpu().numpy().astype(np.float64)
        ob.step(cmd.astype(np.float64))
        jo = ob.joints()
        so = ob.get_state()[0]
        hist.append((c, so.walk_state, np.abs(j - jo).max(), max(abs(v) for l in range(cfg.leg_count) for v in list(so.legs[l].joint_velocity)[:cfg.joint_count]),
                     min(so.legs[l].ik_result for l in range(cfg.leg_count))))
    big = [h for h in hist if h[2] > 1e-6]
    print(f"[{tag}] {precision}: max dq {max(h[2] for h in hist):.3e}; first >1e-6 at", big[0] if big else None)
    for c in range(0, cycles, max(1, cycles // 25)):
        h = hist[c]
        print(f"   c={h[0]:4d} ws={h[1]} dq={h[2]:.2e} max|qd|_oracle={h[3]:.3e} min_ik={h[4]:.3f}")
    eng.close(); ob.close()

cfg = hexapod_config()
run(cfg, lambda c: (0, 0, 0), 300, "mixed", "standstill")
run(cfg, lambda c: (0.5, 0, 0), 500, "mixed", "walk 0.5")
run(cfg, lambda c: (1.0, 0, 0.5) if c < 400 else (0, 0, 0), 700, "mixed", "walk cruise then stop")
run(cfg, lambda c: (0.5, 0, 0) if c < 300 else (0, 0, 0), 800, "f64", "walk then stop f64")
