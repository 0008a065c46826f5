"""ctypes binding of oracle/_ref/libshc_ref.so: the REFERENCE'S OWN control code (its unmodified state_controller.cpp,
model.cpp, walk_controller.cpp, pose_controller.cpp, admittance_controller.cpp compiled from /root/reference against the
stand-in ROS / Eigen / Boost headers of oracle/shim/, see oracle/Makefile.ref and oracle/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY: used by tests/ and tests/golden/make_ref_golden.py to pin the restated oracle.  The library can
only be BUILT where /root/reference exists (this container); the built .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from syropod_highlevel_controller_b200.config import (ShcBodyMsg, ShcConfig, ShcJointStateMsg, ShcLegStateMsg, ShcRobotState,
                                                      ShcStartup)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libshc_ref.so")
REFERENCE_ROOT = os.environ.get("SHC_REFERENCE_ROOT", "/root/reference")
_lib = None


def can_build() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def available() -> bool:
    return os.path.exists(_LIB_PATH) or can_build()


def build(force: bool = False) -> str:
    """Compile the reference's own sources (a few seconds).  Needs /root/reference; otherwise the prebuilt .so is used."""
    if can_build():
        subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.ref", "-s", "-j8", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    elif not os.path.exists(_LIB_PATH):
        raise RuntimeError("oracle/_ref/libshc_ref.so is absent and the reference sources are not here to build it")
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.shc_ref_create.restype = C.c_void_p
        L.shc_ref_create.argtypes = [C.POINTER(ShcConfig)]
        L.shc_ref_destroy.argtypes = [C.c_void_p]
        L.shc_ref_startup_loops.argtypes = [C.c_void_p]
        L.shc_ref_assert_failures.restype = C.c_long
        L.shc_ref_assert_failures.argtypes = [C.c_char_p, C.c_int]
        L.shc_ref_set_pose_reset_mode.argtypes = [C.c_void_p, C.c_int]
        L.shc_ref_set_joint_efforts.argtypes = [C.c_void_p, dp]
        L.shc_ref_step.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.shc_ref_sequence_step.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.shc_ref_set_joint_state.argtypes = [C.c_void_p, dp, dp]
        L.shc_ref_get_messages.argtypes = [C.c_void_p, dp, C.POINTER(ShcJointStateMsg), C.POINTER(ShcLegStateMsg), C.POINTER(ShcBodyMsg)]
        ip = C.POINTER(C.c_int)
        L.shc_ref_request_tip_targets.argtypes = [C.c_void_p, ip, dp, dp, dp, ip, ip, dp, dp]
        L.shc_ref_batch_run_seq.restype = C.c_double
        L.shc_ref_batch_run_seq.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp, C.c_int, C.c_int]
        L.shc_ref_startup_trajectory.argtypes = [C.POINTER(ShcConfig), dp, C.c_int, dp]
        L.shc_ref_select_gait.argtypes = [C.c_void_p, C.POINTER(ShcConfig)]
        L.shc_ref_gait_change_pending.argtypes = [C.c_void_p]
        L.shc_ref_adjust_parameter.argtypes = [C.c_void_p, C.POINTER(ShcConfig)]
        L.shc_ref_parameter_adjust_pending.argtypes = [C.c_void_p]
        L.shc_ref_get_joints.argtypes = [C.c_void_p, dp]
        L.shc_ref_get_state.argtypes = [C.c_void_p, C.POINTER(ShcRobotState)]
        L.shc_ref_get_startup.argtypes = [C.c_void_p, C.POINTER(ShcStartup)]
        L.shc_ref_workspace.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class RefRobot:
    """One instance of the reference's StateController, taken through its direct start-up and put in RUNNING state."""

    def __init__(self, cfg: ShcConfig, transition_through_loop: bool = False):
        """transition_through_loop: READY -> RUNNING by robotStateCallback + loop() as the node does it (that loop() already
        runs one control cycle with a zero command); default: RUNNING set directly, cycle 0 is the caller's."""
        self.cfg = cfg
        self.L, self.D = cfg.leg_count, cfg.joint_count
        self._lib = lib()
        self._lib.shc_ref_transition_through_loop(int(transition_through_loop))
        self._h = self._lib.shc_ref_create(C.byref(cfg))
        self._lib.shc_ref_transition_through_loop(0)

    def close(self):
        if self._h:
            self._lib.shc_ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def startup_loops(self) -> int:
        return self._lib.shc_ref_startup_loops(self._h)

    def startup(self) -> ShcStartup:
        s = ShcStartup()
        self._lib.shc_ref_get_startup(self._h, C.byref(s))
        return s

    def step(self, cmd, imu=None, tip_force=None, manual=None, step_plane=None):
        cmd, imu, tip_force, manual, step_plane = _arr(cmd), _arr(imu), _arr(tip_force), _arr(manual), _arr(step_plane)
        assert cmd.shape == (3,)
        self._lib.shc_ref_step(self._h, _dp(cmd), _dp(imu), _dp(tip_force), _dp(manual), _dp(step_plane))

    def set_pose_reset_mode(self, mode: int):
        self._lib.shc_ref_set_pose_reset_mode(self._h, int(mode))

    def set_joint_efforts(self, efforts):
        e = _arr(efforts)
        assert e.shape == (self.L, self.D)
        self._lib.shc_ref_set_joint_efforts(self._h, _dp(e))

    def sequence_step(self, kind: str, time: float = 0.0) -> int:
        """One loop() of stepToNewStance / packLegs / unpackLegs / executeSequence on the reference's PoseController."""
        return self._lib.shc_ref_sequence_step(self._h, {"new_stance": 0, "pack": 1, "unpack": 2, "start_up": 3, "shut_down": 4}[kind], float(time))

    def set_joint_state(self, position, velocity):
        """Desired joint positions / velocities [L, D] overwritten, forward kinematics re-run (see ref_harness.cpp)."""
        p, v = _arr(position), _arr(velocity)
        assert p.shape == v.shape == (self.L, self.D)
        self._lib.shc_ref_set_joint_state(self._h, _dp(p), _dp(v))

    def messages(self, measured=None):
        """What the reference's own publishers put on the wire for this robot: (JointState, LegState x L, body)."""
        js, legs, body = ShcJointStateMsg(), (ShcLegStateMsg * self.L)(), ShcBodyMsg()
        m = _arr(measured)
        self._lib.shc_ref_get_messages(self._h, _dp(m), C.byref(js), legs, C.byref(body))
        return js, legs, body

    def request_tip_targets(self, target_defined, target_pose, target_transform, clearance, odom_frame, default_defined,
                            default_pose, default_transform):
        """One TargetTipPose message (targetTipPoseCallback) plus the tf answers for its requests; arrays per leg."""
        ints = [np.ascontiguousarray(a, dtype=np.int32) for a in (target_defined, odom_frame, default_defined)]
        dbl = [_arr(a) for a in (target_pose, target_transform, clearance, default_pose, default_transform)]
        ipt = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._lib.shc_ref_request_tip_targets(self._h, ipt(ints[0]), _dp(dbl[0]), _dp(dbl[1]), _dp(dbl[2]), ipt(ints[1]), ipt(ints[2]),
                                              _dp(dbl[3]), _dp(dbl[4]))

    def select_gait(self, new_cfg: ShcConfig):
        """gaitSelectionCallback: the following loop() calls stop the robot, then StateController::changeGait switches."""
        self.cfg = new_cfg
        return self._lib.shc_ref_select_gait(self._h, C.byref(new_cfg))

    @property
    def gait_change_pending(self) -> bool:
        return bool(self._lib.shc_ref_gait_change_pending(self._h))

    def adjust_parameter(self, new_cfg: ShcConfig) -> bool:
        """One dynamic_reconfigure request with new_cfg's adjustable parameters (the first that differs is taken);
        StateController::adjustParameter applies it inside the following loop()."""
        self.cfg = new_cfg
        return bool(self._lib.shc_ref_adjust_parameter(self._h, C.byref(new_cfg)))

    @property
    def parameter_adjust_pending(self) -> bool:
        return bool(self._lib.shc_ref_parameter_adjust_pending(self._h))

    def joints(self) -> np.ndarray:
        out = np.empty((self.L, self.D), dtype=np.float64)
        self._lib.shc_ref_get_joints(self._h, _dp(out))
        return out

    def get_state(self) -> ShcRobotState:
        s = ShcRobotState()
        self._lib.shc_ref_get_state(self._h, C.byref(s))
        return s

    def workspace(self, leg: int, max_planes: int = 16):
        h = np.zeros(max_planes)
        r = np.zeros((max_planes, 9))
        n = self._lib.shc_ref_workspace(self._h, leg, max_planes, _dp(h), _dp(r))
        assert n <= max_planes
        return h[:n], r[:n]

    def assert_failures(self):
        buf = C.create_string_buffer(512)
        n = self._lib.shc_ref_assert_failures(buf, 512)
        return int(n), buf.value.decode()


def batch_run_seq(robots, cmd_seq, threads: int = 1) -> float:
    """All cycles of cmd_seq [cycles, n, 3] for the given RefRobots inside the C library (threads spawned once, robots
    partitioned over them); returns wall seconds.  The CPU baseline of bench.py: the reference's own loop()."""
    cmd_seq = _arr(cmd_seq)
    n = len(robots)
    assert cmd_seq.ndim == 3 and cmd_seq.shape[1:] == (n, 3)
    hs = (C.c_void_p * n)(*[r._h for r in robots])
    return lib().shc_ref_batch_run_seq(hs, n, _dp(cmd_seq), int(cmd_seq.shape[0]), int(threads))


def startup_trajectory(cfg, q_init=None, max_loops: int = 2000):
    """Joint commands of every loop() of the reference's direct start-up from joint angles q_init [L, D] (None: defaults)."""
    L, D = cfg.leg_count, cfg.joint_count
    out = np.zeros((max_loops, L, D))
    q = None if q_init is None else np.ascontiguousarray(q_init, dtype=np.float64)
    n = lib().shc_ref_startup_trajectory(C.byref(cfg), _dp(q), max_loops, _dp(out))
    return out[:n]
