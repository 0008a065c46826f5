"""ctypes binding of the CPU parity oracle (oracle/libshc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.  The restatement is PINNED to the reference's own
code: oracle/ref_py.py runs the reference's unmodified sources (compiled against stand-in ROS / Eigen / Boost headers) and
tests/test_reference_pin.py requires the two to agree on every state field of every cycle; tests/golden/*.npz are the
reference's outputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from syropod_highlevel_controller_b200.config import (ShcBodyMsg, ShcConfig, ShcJointStateMsg, ShcLegStateMsg, ShcRobotState,
                                                      ShcStartup)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libshc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (a few seconds).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def variant_lib(name: str, cxxflags: str):
    """A second build of the same oracle sources with other compiler flags (oracle/_build/libshc_oracle_<name>.so), e.g.
    -ffp-contract=fast: used by tests/test_oracle_chatter.py to show how far two correct double-precision builds of
    the reference arithmetic drift apart inside the stand-still limit cycle."""
    out = os.path.join(_HERE, "_build", f"libshc_oracle_{name}.so")
    srcs = [os.path.join(_HERE, f) for f in ("shc_oracle_model.cpp", "shc_oracle_walk.cpp", "shc_oracle_pose.cpp", "shc_oracle_capi.cpp")]
    deps = srcs + [os.path.join(_HERE, f) for f in ("oracle_math.hpp", "shc_oracle.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++"] + cxxflags.split() + ["-std=c++17", "-fPIC", "-pthread", "-shared", "-o", out] + srcs)
    return _bind(C.CDLL(out))


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _bind(C.CDLL(_LIB_PATH))
    return _lib


def _bind(L):
    if True:
        dp = C.POINTER(C.c_double)
        L.shc_oracle_batch_create.restype = C.c_void_p
        L.shc_oracle_batch_create.argtypes = [C.POINTER(ShcConfig), C.c_int]
        L.shc_oracle_batch_destroy.argtypes = [C.c_void_p]
        L.shc_oracle_batch_size.argtypes = [C.c_void_p]
        L.shc_oracle_startup_loops.argtypes = [C.c_void_p]
        L.shc_oracle_get_startup.argtypes = [C.c_void_p, C.POINTER(ShcStartup)]
        L.shc_oracle_batch_step.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int]
        L.shc_oracle_batch_run.restype = C.c_double
        L.shc_oracle_batch_run.argtypes = [C.c_void_p, dp, C.c_int, C.c_int]
        L.shc_oracle_batch_run_seq.restype = C.c_double
        L.shc_oracle_batch_run_seq.argtypes = [C.c_void_p, dp, C.c_int, C.c_int]
        L.shc_oracle_batch_get_joints.argtypes = [C.c_void_p, dp]
        L.shc_oracle_batch_sequence_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int)]
        L.shc_oracle_batch_get_state.argtypes = [C.c_void_p, C.POINTER(ShcRobotState)]
        L.shc_oracle_batch_set_state.argtypes = [C.c_void_p, C.POINTER(ShcRobotState)]
        L.shc_oracle_batch_set_pose_reset_mode.argtypes = [C.c_void_p, C.c_int]
        L.shc_oracle_batch_set_joint_efforts.argtypes = [C.c_void_p, dp]
        L.shc_oracle_batch_set_step_planes.argtypes = [C.c_void_p, dp]
        L.shc_oracle_smooth_step.restype = C.c_double
        L.shc_oracle_smooth_step.argtypes = [C.c_double]
        L.shc_oracle_round_to_int.argtypes = [C.c_double]
        L.shc_oracle_round_to_even_int.argtypes = [C.c_double]
        L.shc_oracle_quat_to_euler.argtypes = [dp, C.c_int, dp]
        L.shc_oracle_euler_to_quat.argtypes = [dp, C.c_int, dp]
        L.shc_oracle_dh.argtypes = [C.c_double] * 4 + [dp]
        L.shc_oracle_quartic_bezier.argtypes = [dp, C.c_double, dp, dp]
        L.shc_oracle_from_two_vectors.argtypes = [dp, dp, dp]
        L.shc_oracle_slerp.argtypes = [dp, C.c_double, dp, dp]
        L.shc_oracle_pose_ops.argtypes = [dp, dp, dp, dp, dp]
        L.shc_oracle_fk.argtypes = [C.POINTER(ShcConfig), C.c_int, dp, dp]
        L.shc_oracle_solve_ik.argtypes = [C.POINTER(ShcConfig), C.c_int, dp, dp, dp, dp]
        L.shc_oracle_step_cycle.argtypes = [C.POINTER(ShcConfig), C.POINTER(ShcStartup)]
        L.shc_oracle_admittance.argtypes = [C.POINTER(ShcConfig), dp, dp, dp]
        L.shc_oracle_batch_get_messages.argtypes = [C.c_void_p, C.c_int, dp, C.POINTER(ShcJointStateMsg), C.POINTER(ShcLegStateMsg),
                                                    C.POINTER(ShcBodyMsg)]
        L.shc_oracle_workspace.argtypes = [C.POINTER(ShcConfig), C.c_int, C.c_int, C.c_int, dp, dp]
        L.shc_oracle_startup_trajectory.argtypes = [C.POINTER(ShcConfig), dp, C.c_int, dp]
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class OracleBatch:
    """N independent copies of the reference controller, started up once and cloned."""

    def __init__(self, cfg: ShcConfig, n_robots: int = 1, library=None):
        self.cfg = cfg
        self.n = n_robots
        self.L, self.D = cfg.leg_count, cfg.joint_count
        self._lib = library if library is not None else lib()
        self._h = self._lib.shc_oracle_batch_create(C.byref(cfg), n_robots)

    def close(self):
        if self._h:
            self._lib.shc_oracle_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def startup_loops(self) -> int:
        return self._lib.shc_oracle_startup_loops(self._h)

    def startup(self) -> ShcStartup:
        s = ShcStartup()
        self._lib.shc_oracle_get_startup(self._h, C.byref(s))
        return s

    def step(self, cmd, imu=None, tip_force=None, manual=None, threads: int = 1):
        cmd, imu, tip_force, manual = _arr(cmd), _arr(imu), _arr(tip_force), _arr(manual)
        assert cmd.shape == (self.n, 3)
        self._lib.shc_oracle_batch_step(self._h, _dp(cmd), _dp(imu), _dp(tip_force), _dp(manual), threads)

    def run(self, cmd, cycles: int, threads: int = 1) -> float:
        cmd = _arr(cmd)
        return self._lib.shc_oracle_batch_run(self._h, _dp(cmd), cycles, threads)

    def run_seq(self, cmd_seq, threads: int = 1) -> float:
        """All cycles of cmd_seq [cycles, n, 3] inside the C library (threads spawned once); returns wall seconds."""
        cmd_seq = _arr(cmd_seq)
        assert cmd_seq.ndim == 3 and cmd_seq.shape[1:] == (self.n, 3)
        return self._lib.shc_oracle_batch_run_seq(self._h, _dp(cmd_seq), int(cmd_seq.shape[0]), threads)

    def set_pose_reset_mode(self, mode: int):
        """poser_->setPoseResetMode (state_controller.cpp:1199): 0 none, 1 Z+yaw, 2 X+Y, 3 pitch+roll, 4 all, 5 immediate."""
        self._lib.shc_oracle_batch_set_pose_reset_mode(self._h, int(mode))

    def set_joint_efforts(self, efforts):
        """Measured joint efforts [n, L, D] (jointStatesCallback, state_controller.cpp:1565) for calculateTipForce."""
        e = _arr(efforts)
        assert e is None or e.shape == (self.n, self.L, self.D)
        self._lib.shc_oracle_batch_set_joint_efforts(self._h, _dp(e))

    def sequence_step(self, kind: str, time: float = 0.0) -> np.ndarray:
        """One loop() of PoseController::stepToNewStance ("new_stance", pose_controller.cpp:520), packLegs(time) ("pack",
        :597) or unpackLegs(time) ("unpack", :661) for every robot; returns each robot's progress value."""
        out = np.zeros(self.n, dtype=np.int32)
        self._lib.shc_oracle_batch_sequence_step(self._h, {"new_stance": 0, "pack": 1, "unpack": 2, "start_up": 3, "shut_down": 4}[kind], float(time),
                                                 out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def set_tip_step_planes(self, step_planes):
        """Tip range-sensor readings [n, L, 3] (TipState.step_plane; z >= 1e9 stands for the reference's UNASSIGNED_VALUE), or None."""
        if step_planes is None:
            self._lib.shc_oracle_batch_set_step_planes(self._h, None)
            return
        sp = np.array(step_planes, dtype=np.float64)
        assert sp.shape == (self.n, self.L, 3)
        sp[..., 2] = np.where(sp[..., 2] >= 1e9, float(2 ** 31 - 1), sp[..., 2])
        self._lib.shc_oracle_batch_set_step_planes(self._h, _dp(np.ascontiguousarray(sp)))

    def joints(self) -> np.ndarray:
        out = np.empty((self.n, self.L, self.D), dtype=np.float64)
        self._lib.shc_oracle_batch_get_joints(self._h, _dp(out))
        return out

    def get_state(self):
        arr = (ShcRobotState * self.n)()
        self._lib.shc_oracle_batch_get_state(self._h, arr)
        return arr

    def messages(self, index: int, measured=None):
        """The reference's publishers (state_controller.cpp:777-1047) for one robot: (JointState, LegState x L, body)."""
        js, legs, body = ShcJointStateMsg(), (ShcLegStateMsg * self.L)(), ShcBodyMsg()
        m = _arr(measured)
        self._lib.shc_oracle_batch_get_messages(self._h, index, _dp(m), C.byref(js), legs, C.byref(body))
        return js, legs, body

    def set_state(self, arr):
        self._lib.shc_oracle_batch_set_state(self._h, arr)


def fk(cfg, leg, q):
    q = _arr(q)
    out = np.empty(3)
    lib().shc_oracle_fk(C.byref(cfg), leg, _dp(q), _dp(out))
    return out


def solve_ik(cfg, leg, q, qd, delta):
    q, qd, delta = _arr(q), _arr(qd), _arr(delta)
    out = np.empty(len(q))
    lib().shc_oracle_solve_ik(C.byref(cfg), leg, _dp(q), _dp(qd), _dp(delta), _dp(out))
    return out


def step_cycle(cfg) -> ShcStartup:
    s = ShcStartup()
    lib().shc_oracle_step_cycle(C.byref(cfg), C.byref(s))
    return s


def workspace(cfg, leg: int, full: bool, max_planes: int = 16):
    """Leg::generateWorkspace (model.cpp:309-510) of one leg after the direct start-up: (heights [P], radii [P, 9])."""
    h = np.zeros(max_planes)
    r = np.zeros((max_planes, 9))
    n = lib().shc_oracle_workspace(C.byref(cfg), leg, int(full), max_planes, _dp(h), _dp(r))
    assert n <= max_planes
    return h[:n], r[:n]


def startup_trajectory(cfg, q_init=None, max_loops: int = 2000):
    """Joint commands of every loop() of the direct start-up from joint angles q_init [L, D] (None: defaults): [loops, L, D]."""
    L, D = cfg.leg_count, cfg.joint_count
    out = np.zeros((max_loops, L, D))
    q = None if q_init is None else np.ascontiguousarray(q_init, dtype=np.float64)
    n = lib().shc_oracle_startup_trajectory(C.byref(cfg), _dp(q), max_loops, _dp(out))
    return out[:n]
