"""CPU: the oracle against every known answer derivable by hand from the reference code
and against independent numpy restatements.  The reference ships no tests or golden vectors; the pin to the reference's own code is tests/test_reference_pin.py."""
import ctypes as C
import math

import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# dt, gait, period, swing_start, swing_end, stance_start, stance_period, swing_period, offsets (AR,BR,CR,CL,BL,AL)
STEP_CYCLES = [
    (0.02, "wave_gait", 312, 130, 182, 182, 260, 52, [104, 156, 208, 52, 0, 260]),
    (0.02, "tripod_gait", 104, 26, 78, 78, 52, 52, [0, 52, 0, 52, 0, 52]),
    (0.02, "ripple_gait", 156, 52, 104, 104, 104, 52, [52, 0, 104, 26, 78, 130]),
    (0.02, "amble_gait", 150, 50, 100, 100, 100, 50, [50, 100, 0, 50, 100, 0]),
    (0.01, "wave_gait", 600, 250, 350, 350, 500, 100, [200, 300, 400, 100, 0, 500]),
    (0.01, "tripod_gait", 200, 50, 150, 150, 100, 100, [0, 100, 0, 100, 0, 100]),
    (0.01, "ripple_gait", 300, 100, 200, 200, 200, 100, [100, 0, 200, 50, 150, 250]),
    (0.01, "amble_gait", 300, 100, 200, 200, 200, 100, [100, 200, 0, 100, 200, 0]),
]


@pytest.mark.parametrize("dt,gait,period,ss,se,sts,stp,swp,offs", STEP_CYCLES)
def test_step_cycle_table(oracle, dt, gait, period, ss, se, sts, stp, swp, offs):
    """WalkController::generateStepCycle (walk_controller.cpp:365-389) + phase offsets (:237-278)."""
    s = oracle.step_cycle(hexapod_config(gait, dt))
    assert (s.period, s.swing_start, s.swing_end, s.stance_start, s.stance_period, s.swing_period) == \
        (period, ss, se, sts, stp, swp)
    assert list(s.phase_offsets)[:6] == offs
    assert s.stance_end == s.swing_start
    assert s.step_frequency == pytest.approx(1.0 / (period * dt), rel=1e-15)


def test_scalar_helpers(oracle):
    L = oracle.lib()
    # mod is Euclidean (standard_includes.h:76)
    assert L.shc_oracle_mod(-1, 360) == 359 and L.shc_oracle_mod(725, 360) == 5 and L.shc_oracle_mod(-45, 360) == 315
    # roundToInt rounds half away from zero (:93)
    assert [L.shc_oracle_round_to_int(x) for x in (0.5, 1.49, -0.5, -1.5, 2.5)] == [1, 1, -1, -2, 3]
    # roundToEvenInt is int(x) when even, else int(x)+1 — NOT nearest even (:98)
    assert [L.shc_oracle_round_to_even_int(x) for x in (2.9, 3.1, 4.0, 5.999, 26.0)] == [2, 4, 4, 6, 26]
    for c, v in ((0.0, 0.0), (0.5, 0.5), (1.0, 1.0), (0.25, 0.103515625)):
        assert L.shc_oracle_smooth_step(c) == pytest.approx(v, abs=1e-15)


def test_dh_matrix(oracle):
    """createDHMatrix(0, theta, r, 0) is a planar rotation with translation (r cos, r sin, 0)."""
    out = np.empty(16)
    oracle.lib().shc_oracle_dh(0.0, 0.3, 0.05, 0.0, _dp(out))
    m = out.reshape(4, 4)
    ref = np.array([[math.cos(0.3), -math.sin(0.3), 0, 0.05 * math.cos(0.3)],
                    [math.sin(0.3), math.cos(0.3), 0, 0.05 * math.sin(0.3)], [0, 0, 1, 0], [0, 0, 0, 1]])
    assert np.allclose(m, ref, atol=1e-16)
    # general case against Rz(theta) Tz(d) Tx(r) Rx(alpha)
    d, th, r, al = 0.02, -0.7, 0.11, 1.571

    def rz(a): return np.array([[math.cos(a), -math.sin(a), 0, 0], [math.sin(a), math.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    def rx(a): return np.array([[1, 0, 0, 0], [0, math.cos(a), -math.sin(a), 0], [0, math.sin(a), math.cos(a), 0], [0, 0, 0, 1.0]])
    def tr(x, y, z): m = np.eye(4); m[:3, 3] = (x, y, z); return m
    oracle.lib().shc_oracle_dh(d, th, r, al, _dp(out))
    assert np.allclose(out.reshape(4, 4), rz(th) @ tr(0, 0, d) @ tr(r, 0, 0) @ rx(al), atol=1e-15)


FK_KNOWN = [  # leg AR of default.yaml
    ((0.0, 0.0, -0.1), (0.214856298691, -0.123871196145, -0.019866932667)),
    ((0.0, 0.785, -1.138), (0.195169268563, -0.112523286003, -0.008425232172)),
    ((0.2, -0.3, -1.0), (0.152152381157, -0.061379602448, -0.113320980981)),
]


@pytest.mark.parametrize("q,tip", FK_KNOWN)
def test_forward_kinematics_known(oracle, q, tip):
    assert np.allclose(oracle.fk(hexapod_config(), 0, q), tip, atol=2e-12)


def _numpy_chain(cfg, leg, q):
    """Independent restatement of the DH chain (not from the reference): list of 4x4 transforms T1, T2.., Ttip."""
    def dh(d, th, r, al):
        return np.array([[math.cos(th), -math.sin(th) * math.cos(al), math.sin(th) * math.sin(al), r * math.cos(th)],
                         [math.sin(th), math.cos(th) * math.cos(al), -math.cos(th) * math.sin(al), r * math.sin(th)],
                         [0, math.sin(al), math.cos(al), d], [0, 0, 0, 1.0]])
    D = cfg.joint_count
    ts = [dh(cfg.link_d[leg][0], cfg.link_theta[leg][0], cfg.link_r[leg][0], cfg.link_alpha[leg][0])]
    for j in range(1, D + 1):
        ts.append(dh(cfg.link_d[leg][j], cfg.link_theta[leg][j] + q[j - 1], cfg.link_r[leg][j], cfg.link_alpha[leg][j]))
    return ts


def _numpy_dls(cfg, leg, q, qd, delta):
    """Independent numpy DLS step: J^T (J J^T + l^2 I)^-1 delta + (I - J^+ J) g  (SURVEY.md §3.3)."""
    D = cfg.joint_count
    ts = _numpy_chain(cfg, leg, q)
    frames = [np.eye(4)]
    for t in ts[1:]:
        frames.append(frames[-1] @ t)
    pe = frames[-1][:3, 3]
    J = np.zeros((6, D))
    for i in range(D):
        z, p = frames[i][:3, 2], frames[i][:3, 3]
        J[:3, i] = np.cross(z, pe - p)
    Jinv = J.T @ np.linalg.inv(J @ J.T + 0.02 ** 2 * np.eye(6))
    w = 0.1
    lo = np.array([cfg.joint_min[leg][j] for j in range(D)]); hi = np.array([cfg.joint_max[leg][j] for j in range(D)])
    vm = np.array([cfg.joint_max_vel[leg][j] for j in range(D)])
    rng, ctr = hi - lo, lo + (hi - lo) / 2
    pc = np.sum((w * (q - ctr) / rng) ** 2); vc = np.sum((w * qd / (2 * vm)) ** 2)
    gp = -w * w * (q - ctr) / rng ** 2 * (0 if pc == 0 else 1 / math.sqrt(pc))
    gv = -w * w * qd / (2 * vm) ** 2 * (0 if vc == 0 else 1 / math.sqrt(vc))
    g = 0.25 * gp + 0.75 * gv
    d6 = np.concatenate([delta, np.zeros(3)])
    return Jinv @ d6 + (np.eye(D) - Jinv @ J) @ g


def test_solve_ik_known_and_numpy(oracle):
    cfg = hexapod_config()
    dq = oracle.solve_ik(cfg, 0, [0, 0.785, -1.138], [0, 0, 0], [0.001, 0.0005, -0.002])
    assert np.allclose(dq, (0.002813657956, -0.029328982122, 0.019660649581), atol=2e-12)
    rng = np.random.default_rng(1)
    for cfg in (hexapod_config(), octopod_config()):
        D = cfg.joint_count
        for leg in range(cfg.leg_count):
            lo = np.array([cfg.joint_min[leg][j] for j in range(D)]); hi = np.array([cfg.joint_max[leg][j] for j in range(D)])
            q = lo + (hi - lo) * rng.uniform(0.1, 0.9, D)
            qd = rng.normal(0, 0.5, D)
            delta = rng.normal(0, 0.002, 3)
            assert np.allclose(oracle.solve_ik(cfg, leg, q, qd, delta), _numpy_dls(cfg, leg, q, qd, delta), atol=1e-12)


def _closed_form_ik_3dof(cfg, leg, tip_robot):
    """Closed-form IK of the shipped 3-DOF leg (yaw - twist alpha - pitch - pitch; SURVEY.md §4 "Cross-check"), derived
    independently of the DLS iteration.  In the leg frame (T1 removed) the tip is
        p = Rz(q1) [ (r1 + u, v cos(alpha), v sin(alpha)) ],  (u, v) = the planar 2R point of femur and tibia,
    so v = p_z / sin(alpha), (r1 + u)^2 = p_x^2 + p_y^2 - (v cos(alpha))^2, and q2, q3 follow from the planar 2R problem
    (knee branch: tibia angle negative, as its joint limits demand)."""
    th0, r0 = cfg.link_theta[leg][0], cfg.link_r[leg][0]
    x, y, z = tip_robot[0] - r0 * math.cos(th0), tip_robot[1] - r0 * math.sin(th0), tip_robot[2] - cfg.link_d[leg][0]
    px, py, pz = math.cos(th0) * x + math.sin(th0) * y, -math.sin(th0) * x + math.cos(th0) * y, z
    r1, al = cfg.link_r[leg][1], cfg.link_alpha[leg][1]
    r2, r3, off3 = cfg.link_r[leg][2], cfg.link_r[leg][3], cfg.link_theta[leg][3]
    assert cfg.link_theta[leg][1] == 0 and cfg.link_theta[leg][2] == 0 and cfg.link_alpha[leg][2] == 0 and cfg.link_alpha[leg][3] == 0
    v = pz / math.sin(al)
    w = v * math.cos(al)
    u = math.sqrt(px * px + py * py - w * w) - r1
    q1 = math.atan2(py, px) - math.atan2(w, r1 + u)
    c3 = (u * u + v * v - r2 * r2 - r3 * r3) / (2 * r2 * r3)
    a3 = -math.acos(max(-1.0, min(1.0, c3)))
    q2 = math.atan2(v, u) - math.atan2(r3 * math.sin(a3), r2 + r3 * math.cos(a3))
    return np.array([q1, q2, a3 - off3])


def test_dls_fixed_point_is_the_closed_form_ik(oracle):
    """the 3-DOF leg has a closed-form IK; the fixed point of the reference's DLS iteration (Leg::applyIK
    repeated on one target) must be that solution — up to the amplitude of its own stand-still limit cycle (DESIGN.md
    "Reference dynamics": ~1e-3 rad, i.e. ~0.2 mm at the tip)."""
    cfg = hexapod_config()
    L_ = oracle.lib()
    dp = C.POINTER(C.c_double)
    L_.shc_oracle_apply_ik.restype = C.c_double
    L_.shc_oracle_apply_ik.argtypes = [C.POINTER(type(cfg)), C.c_int, dp, dp, dp, C.c_int, dp]
    rng = np.random.default_rng(9)
    worst_q = worst_tip = 0.0
    for leg in range(cfg.leg_count):
        for _ in range(6):
            q_true = np.array([rng.uniform(-0.45, 0.45), rng.uniform(-0.2, 1.1), rng.uniform(-2.0, -0.5)])
            target = oracle.fk(cfg, leg, q_true)
            # the closed form inverts the forward kinematics exactly ...
            assert np.abs(_closed_form_ik_3dof(cfg, leg, target) - q_true).max() < 1e-10
            # ... and the DLS iteration converges onto it from a perturbed start
            q = q_true + rng.uniform(-0.08, 0.08, 3)
            q[2] = min(q[2], -0.15)
            qd, tip = np.zeros(3), np.zeros(3)
            for _ in range(400):
                L_.shc_oracle_apply_ik(C.byref(cfg), leg, q.ctypes.data_as(dp), qd.ctypes.data_as(dp),
                                       np.ascontiguousarray(target).ctypes.data_as(dp), 0, tip.ctypes.data_as(dp))
            worst_q, worst_tip = max(worst_q, float(np.abs(q - q_true).max())), max(worst_tip, float(np.abs(tip - target).max()))
    # The iteration does not stop exactly on the solution: with damping, (I - J^+ J) is not zero for a square J, so the
    # normalised cost gradient (a push of fixed size, model.cpp:788-790) keeps a small standing offset and the stand-still
    # limit cycle on top of it.  Both stay far inside IK_TOLERANCE (5 mm).
    print(f"[closed-form IK] worst joint deviation {worst_q:.2e} rad, worst tip deviation {worst_tip:.2e} m")
    assert worst_q < 1.5e-2 and worst_tip < 5e-4


def test_fk_matches_numpy_chain(oracle):
    rng = np.random.default_rng(2)
    for cfg in (hexapod_config(), octopod_config()):
        D = cfg.joint_count
        for leg in range(cfg.leg_count):
            q = rng.uniform(-1, 1, D)
            t = np.eye(4)
            for m in _numpy_chain(cfg, leg, q):
                t = t @ m
            assert np.allclose(oracle.fk(cfg, leg, q), t[:3, 3], atol=1e-14)


def test_admittance_affine_map(oracle):
    """30 RK4 steps of the virtual spring-damper = one affine map; numbers of SURVEY.md A.6 (default.yaml:126-130)."""
    cfg = hexapod_config()
    x_in = np.array([0.01, -0.02]); out = np.empty(2)
    # force only on the last axis so the 2-state is integrated twice with F=0 and once with F=3 (trap 2)
    P = np.array([[0.8883377, 0.31682983], [-0.38019579, 0.3330262]]); q = np.array([-0.00930519, -0.03168298])
    f = np.array([3.0, 0.0, 0.0])
    oracle.lib().shc_oracle_admittance(C.byref(cfg), _dp(x_in), _dp(f), _dp(out))
    x = P @ x_in + q * 3.0
    x = P @ x
    x = P @ x
    assert np.allclose(out, x, atol=2e-8)
    # brute-force RK4 check of one pass quoted in the survey
    assert np.allclose(P @ x_in + q * 3.0, (-0.02536879, -0.10551143), atol=1e-7)


def test_euler_quaternion_round_trip(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for intrinsic in (0, 1):
        for _ in range(200):
            e = rng.uniform(-1.2, 1.2, 3)
            q = np.empty(4); e2 = np.empty(3)
            L.shc_oracle_euler_to_quat(_dp(e), intrinsic, _dp(q))
            assert abs(np.linalg.norm(q) - 1) < 1e-15
            L.shc_oracle_quat_to_euler(_dp(q), intrinsic, _dp(e2))
            assert np.allclose(e, e2, atol=1e-13), (intrinsic, e, e2)
    # identity and the zero ("undefined") quaternion both read as zero angles (toRotationMatrix of 0 is the identity)
    for qv in ((1.0, 0, 0, 0), (0.0, 0, 0, 0)):
        q = np.array(qv); e2 = np.empty(3)
        L.shc_oracle_quat_to_euler(_dp(q), 0, _dp(e2))
        assert np.all(e2 == 0)


def test_euler_extrinsic_matches_rotation_composition(oracle):
    """eulerAnglesToQuaternion(extrinsic) = Rz(yaw) Ry(pitch) Rx(roll) (standard_includes.h:237)."""
    from scipy.spatial.transform import Rotation as R
    L = oracle.lib()
    e = np.array([0.3, -0.2, 0.7]); q = np.empty(4)
    L.shc_oracle_euler_to_quat(_dp(e), 0, _dp(q))
    ref = (R.from_euler("z", 0.7) * R.from_euler("y", -0.2) * R.from_euler("x", 0.3)).as_quat()  # x y z w
    assert np.allclose(q, (ref[3], ref[0], ref[1], ref[2]), atol=1e-15)
    L.shc_oracle_euler_to_quat(_dp(e), 1, _dp(q))
    ref = (R.from_euler("x", 0.3) * R.from_euler("y", -0.2) * R.from_euler("z", 0.7)).as_quat()
    assert np.allclose(q, (ref[3], ref[0], ref[1], ref[2]), atol=1e-15)


def test_pose_algebra(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(4)
    for _ in range(50):
        def rp():
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            return np.concatenate([rng.normal(size=3), q])
        a, b = rp(), rp()
        add, rem, inv = np.empty(7), np.empty(7), np.empty(7)
        L.shc_oracle_pose_ops(_dp(a), _dp(b), _dp(add), _dp(rem), _dp(inv))
        # ~(~p) == p
        inv2 = np.empty(7); junk = np.empty(7)
        L.shc_oracle_pose_ops(_dp(inv), _dp(b), _dp(junk), _dp(junk.copy()), _dp(inv2))
        assert np.allclose(inv2, a, atol=1e-14)
        # addPose then removePose of the same pose restores the rotation; positions follow pose.h:167-184 literally
        back = np.empty(7)
        L.shc_oracle_pose_ops(_dp(add), _dp(b), _dp(junk), _dp(back), _dp(junk.copy()))
        assert np.allclose(back[3:], a[3:], atol=1e-14)


def test_quartic_bezier(oracle):
    L = oracle.lib()
    nodes = np.arange(15, dtype=float).reshape(5, 3) ** 1.5
    p, d = np.empty(3), np.empty(3)
    for t in (0.0, 0.3, 1.0):
        L.shc_oracle_quartic_bezier(_dp(nodes.ravel().copy()), t, _dp(p), _dp(d))
        s = 1 - t
        w = [s ** 4, 4 * t * s ** 3, 6 * t * t * s * s, 4 * t ** 3 * s, t ** 4]
        assert np.allclose(p, sum(wi * n for wi, n in zip(w, nodes)), atol=1e-12)
        h = 1e-6
        if 0 < t < 1:
            def B(tt):
                ss = 1 - tt
                ww = [ss ** 4, 4 * tt * ss ** 3, 6 * tt * tt * ss * ss, 4 * tt ** 3 * ss, tt ** 4]
                return sum(wi * n for wi, n in zip(ww, nodes))
            assert np.allclose(d, (B(t + h) - B(t - h)) / (2 * h), rtol=1e-6)
    # evenly spaced nodes -> constant derivative 4 * separation (stance curve, walk_controller.cpp:1295)
    sep = np.array([0.01, -0.02, 0.0]); o = np.array([0.1, 0.2, 0.0])
    nodes = np.stack([o + k * sep for k in range(5)])
    L.shc_oracle_quartic_bezier(_dp(nodes.ravel().copy()), 0.37, _dp(p), _dp(d))
    assert np.allclose(d, 4 * sep, atol=1e-15)


def test_from_two_vectors_and_slerp(oracle):
    L = oracle.lib()
    from scipy.spatial.transform import Rotation as R
    a, b = np.array([0, 0, 1.0]), np.array([0.1, -0.2, 0.97]); q = np.empty(4)
    L.shc_oracle_from_two_vectors(_dp(a), _dp(b), _dp(q))
    rot = R.from_quat([q[1], q[2], q[3], q[0]])
    assert np.allclose(rot.apply(a), b / np.linalg.norm(b), atol=1e-15)
    qa = np.array([1.0, 0, 0, 0]); out = np.empty(4)
    L.shc_oracle_slerp(_dp(qa), 0.25, _dp(q), _dp(out))
    half = R.from_rotvec(rot.as_rotvec() * 0.25).as_quat()
    assert np.allclose(out, (half[3], half[0], half[1], half[2]), atol=1e-14)
