"""CPU: behavioural invariants of the oracle's rollouts and the committed golden fixtures — outputs of
the reference's own code (oracle/_ref, tests/golden/make_golden.py), which the oracle must reproduce to 1e-12 rad."""
import glob
import os

import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg_for(name):
    import parity_cases

    return parity_cases.cfg_for_golden(name)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=os.path.basename)
def test_oracle_reproduces_golden(oracle, path):
    """The oracle reproduces the vectors the REFERENCE produced (tests/golden/make_golden.py runs oracle/_ref)."""
    assert str(np.load(path)["source"]) == "reference"
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    cfg = _cfg_for(name)
    ob = oracle.OracleBatch(cfg, 1)
    for c in range(len(g["cmd"])):
        ob.step(g["cmd"][c][None].astype(np.float64), g["imu"][c][None].astype(np.float64) if "imu" in g else None,
                g["force"][c][None].astype(np.float64) if "force" in g else None)
        assert np.abs(ob.joints()[0] - g["joints"][c]).max() < 1e-12, (name, c)
        assert ob.get_state()[0].walk_state == g["walk_state"][c]


def test_startup_reaches_default_stance(oracle):
    cfg = hexapod_config()
    ob = oracle.OracleBatch(cfg, 1)
    assert ob.startup_loops == 301  # 1 (UNKNOWN->PACKED) + time_to_start / time_delta
    st = ob.get_state()[0]
    for l in range(6):
        tip = np.array(st.legs[l].model_tip_position)
        want = np.array([cfg.stance_x[l], cfg.stance_y[l], -cfg.body_clearance])
        assert np.abs(tip - want).max() < 1e-4  # default-stance tips end at z = -clearance in the body frame (§3.2)
        assert st.legs[l].step_state == 1 and st.legs[l].phase == 0
    su = ob.startup()
    # tripod @ 0.02: max_stance_extension = 26 -> time_to_max_stride = (26+52+52)*0.02 = 2.6 s
    for b in range(9):
        assert su.max_linear_acceleration[b] == pytest.approx(su.max_linear_speed[b] / 2.6, rel=1e-12)
    assert su.walkspace[0] == su.walkspace[8] == su.walkspace[4]  # symmetric, and 360 == 0


def test_walk_state_machine_and_ik_closure(oracle):
    cfg = hexapod_config()
    ob = oracle.OracleBatch(cfg, 1)
    states, max_err = [], 0.0
    for c in range(900):
        cmd = np.array([[0.5, 0.0, 0.0]]) if c < 500 else np.zeros((1, 3))
        ob.step(cmd)
        st = ob.get_state()[0]
        states.append(st.walk_state)
        for l in range(6):
            e = np.abs(np.array(st.legs[l].model_tip_position) - np.array(st.legs[l].desired_tip_position)).max()
            max_err = max(max_err, e)
    changes = [s for i, s in enumerate(states) if i == 0 or s != states[i - 1]]
    assert changes == [0, 1, 2, 3]  # STARTING -> MOVING -> STOPPING -> STOPPED
    assert states[0] == 0 and states[-1] == 3
    assert max_err < 0.005  # inside the workspace IK closes to IK_TOLERANCE (model.cpp:916-929)


def test_stance_velocity_is_constant_and_tripod_groups_alternate(oracle):
    cfg = hexapod_config()
    ob = oracle.OracleBatch(cfg, 1)
    su = ob.startup()
    hist = []
    for c in range(700):
        ob.step(np.array([[0.5, 0.0, 0.0]]))
        hist.append(ob.get_state()[0])
    # steady state: stance tip velocity == -stride / stance_time (walk_controller.cpp:1295-1310)
    stance_time = su.stance_period * cfg.time_delta
    checked = 0
    for st in hist[400:]:
        for l in range(6):
            leg = st.legs[l]
            if leg.step_state == 1 and 0.1 < leg.stance_progress < 0.9:
                v = np.array(leg.tip_velocity); s = np.array(leg.stride_vector)
                assert np.allclose(v, -s / stance_time, atol=1e-12)
                checked += 1
    assert checked > 500
    # tripod: legs {0,2,4} and {1,3,5} are half a period apart
    st = hist[-1]
    ph = [st.legs[l].phase for l in range(6)]
    assert ph[0] == ph[2] == ph[4] and ph[1] == ph[3] == ph[5]
    assert (ph[1] - ph[0]) % su.period == su.period // 2


def test_swing_is_continuous_at_both_joins(oracle):
    """C0/C1 at the stance->swing and swing->stance joins (walk_controller.cpp:1249-1260, 1270-1280)."""
    cfg = hexapod_config()
    ob = oracle.OracleBatch(cfg, 1)
    prev = None
    worst_jump = 0.0
    for c in range(600):
        ob.step(np.array([[0.4, 0.1, 0.0]]))
        st = ob.get_state()[0]
        if prev is not None and c > 300:
            for l in range(6):
                dv = np.array(st.legs[l].tip_velocity) - np.array(prev.legs[l].tip_velocity)
                worst_jump = max(worst_jump, np.abs(dv).max())
        prev = st
    # tip speeds are ~0.05-0.3 m/s; velocity changes between consecutive 20 ms cycles stay small everywhere
    assert worst_jump < 0.05


def test_oracle_state_round_trip(oracle):
    cfg = octopod_config()
    a = oracle.OracleBatch(cfg, 1)
    b = oracle.OracleBatch(cfg, 1)
    rng = np.random.default_rng(0)
    for c in range(150):
        a.step(np.array([[0.6, -0.2, 0.3]]))
    b.set_state(a.get_state())
    for c in range(100):
        cmd = np.array([[0.6, -0.2, 0.3]]) if c < 50 else np.zeros((1, 3))
        a.step(cmd); b.step(cmd)
        assert np.abs(a.joints() - b.joints()).max() < 1e-12
