import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.streams import CommandStream
from oracle import oracle_py as O

def run(cfg, n, cycles, precision, cmds=None, tag=""):
    L, D = 6, 3
    ob = O.OracleBatch(cfg, n); eng = Engine(cfg, n, precision=precision, startup=ob.startup())
    cs = CommandStream(n)
    err = np.zeros((cycles, n, L)); qdn = np.zeros((cycles, n, L))
    for c in range(cycles):
        cmd = cs.next() if cmds is None else cmds[c][None].repeat(n, 0)
        j = eng.step(torch.from_numpy(cmd).cuda()).cpu().numpy().astype(np.float64)
        ob.step(cmd.astype(np.float64), threads=8)
        err[c] = np.abs(j - ob.joints()).max(axis=2)
        st = ob.get_state()
        for r in range(n):
            for l in range(L):
                qdn[c, r, l] = np.linalg.norm(list(st[r].legs[l].joint_velocity)[:D])
    print(f"[{tag}] {precision} dt={cfg.time_delta} n={n}: max err {err.max():.2e}; frac leg-cycles >1e-6: {np.mean(err>1e-6):.5f}, >1e-7: {np.mean(err>1e-7):.5f}")
    # for excursions: min qd norm in the preceding W cycles
    for W in (10, 30, 60):
        mins = []
        for c, r, l in zip(*np.nonzero(err > 1e-6)):
            mins.append(qdn[max(0, c - W):c + 1, r, l].min())
        if mins:
            mins = np.array(mins)
            print(f"    W={W}: excursions {len(mins)}; min-qd-in-window quantiles: 50% {np.quantile(mins,0.5):.3f} 90% {np.quantile(mins,0.9):.3f} 99% {np.quantile(mins,0.99):.3f} max {mins.max():.3f}")
    # how common are low-velocity windows overall
    for thr in (0.1, 0.2, 0.3, 0.5):
        print(f"    frac leg-cycles with qd norm < {thr}: {np.mean(qdn < thr):.3f}")
    eng.close(); ob.close()

g = np.load("tests/golden/config1_100hz_straight.npz")
run(hexapod_config("tripod_gait", 0.01), 1, 1000, "f64", cmds=g["cmd"], tag="golden 100hz straight")
run(hexapod_config("tripod_gait", 0.01), 256, 1000, "f64", tag="random 100hz")
run(hexapod_config("tripod_gait", 0.02), 256, 1000, "f64", tag="random 50hz")
run(hexapod_config("tripod_gait", 0.02), 256, 1000, "mixed", tag="random 50hz")
