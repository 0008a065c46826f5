import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.streams import CommandStream, ImuStream, ForceStream
from oracle import oracle_py as O

def run(cfg, n, cycles, precision, tag, imu=False, force=False):
    L, D = cfg.leg_count, cfg.joint_count
    ob = O.OracleBatch(cfg, n); eng = Engine(cfg, n, precision=precision, startup=ob.startup())
    cs = CommandStream(n); ims = ImuStream(n) if imu else None; fs = ForceStream(n, L) if force else None
    err = np.zeros((cycles, n)); dev = np.zeros((cycles, n)); lim = np.zeros((cycles, n), bool)
    for c in range(cycles):
        cmd = cs.next(); im = ims.next(cfg.time_delta) if ims else None; fo = fs.next() if fs else None
        j = eng.step(torch.from_numpy(cmd).cuda(), None if im is None else torch.from_numpy(im).cuda(), None if fo is None else torch.from_numpy(fo).cuda()).cpu().numpy().astype(np.float64)
        ob.step(cmd.astype(np.float64), None if im is None else im.astype(np.float64), None if fo is None else fo.astype(np.float64), threads=8)
        err[c] = np.abs(j - ob.joints()).reshape(n, -1).max(axis=1)
        st = ob.get_state()
        for r in range(n):
            d = 0.0; z = False
            for l in range(L):
                lg = st[r].legs[l]
                d = max(d, max(abs(lg.model_tip_position[k] - lg.desired_tip_position[k]) for k in range(3)))
                z = z or lg.ik_result == 0.0
            dev[c, r] = d; lim[c, r] = z
    print(f"[{tag}] {precision} n={n} cycles={cycles}: overall max err {err.max():.2e}; frac robot-cycles with dev>5mm {np.mean(dev>0.005):.4f}; ik_result==0 {np.mean(lim):.4f}")
    for thr in (0.005, 0.003, 0.002):
        for W in (1, 20, 60, 150):
            bad = dev > thr
            # healthy[c] = no bad in the last W cycles (inclusive)
            cs_ = np.cumsum(np.vstack([np.zeros((1, n)), bad]), axis=0)
            recent = np.zeros_like(bad)
            for c in range(cycles):
                lo = max(0, c - W + 1)
                recent[c] = (cs_[c + 1] - cs_[lo]) > 0
            healthy = ~recent
            print(f"    thr={thr} W={W}: healthy frac {healthy.mean():.3f}, max err healthy {err[healthy].max():.2e}, p99.9 {np.quantile(err[healthy], 0.999):.2e}; max err unhealthy {err[~healthy].max() if (~healthy).any() else 0:.2e}")
    eng.close(); ob.close()

run(hexapod_config(), 192, 1200, "mixed", "hex tripod")
run(hexapod_config(), 192, 1200, "f64", "hex tripod")
run(octopod_config(), 96, 800, "mixed", "octo", imu=True, force=True)
run(octopod_config(), 96, 800, "f64", "octo", imu=True, force=True)
