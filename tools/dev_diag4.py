import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syropod_highlevel_controller_b200.config import hexapod_config
from syropod_highlevel_controller_b200.engine import Engine
from syropod_highlevel_controller_b200.streams import CommandStream
from oracle import oracle_py as O
cfg = hexapod_config(); n = 96; cycles = 900; L, D = 6, 3
ob = O.OracleBatch(cfg, n); eng = Engine(cfg, n, precision="mixed", startup=ob.startup())
cs = CommandStream(n)
rec = []
for c in range(cycles):
    cmd = cs.next()
    j = eng.step(torch.from_numpy(cmd).cuda()).cpu().numpy().astype(np.float64)
    ob.step(cmd.astype(np.float64), threads=8)
    jo = ob.joints()
    st = ob.get_state(); se = eng.get_state()
    err = np.abs(j - jo).reshape(n, -1).max(axis=1)
    row = []
    for r in range(n):
        qdmax = max(abs(v) for l in range(L) for v in list(st[r].legs[l].joint_velocity)[:D])
        ikmin = min(st[r].legs[l].ik_result for l in range(L))
        tiperr = max(abs(se[r].legs[l].tip_position[k] - st[r].legs[l].tip_position[k]) for l in range(L) for k in range(3))
        # which leg/joint has the max error
        e = np.abs(j[r] - jo[r]); li, ji = np.unravel_index(np.argmax(e), e.shape)
        row.append((err[r], st[r].walk_state, tuple(np.round(cmd[r], 3)), qdmax, ikmin, tiperr, li, ji,
                    st[r].legs[li].step_state, st[r].legs[li].phase, jo[r, li, ji]))
    rec.append(row)
first = {}
for c in range(cycles):
    for r in range(n):
        if rec[c][r][0] > 2e-6 and r not in first:
            first[r] = c
print("robots exceeding 2e-6:", len(first), "of", n)
for r, c0 in list(first.items())[:6]:
    print(f"robot {r} first exceed at cycle {c0}")
    for c in range(max(0, c0 - 6), min(cycles, c0 + 10)):
        e, ws, cmd, qd, ik, te, li, ji, ss, ph, q = rec[c][r]
        print(f"   c={c} err={e:.2e} ws={ws} cmd={cmd} |qd|max={qd:.3f} ikmin={ik:.3f} tiperr={te:.1e} leg={li} joint={ji} step_state={ss} phase={ph} q={q:.4f}")
