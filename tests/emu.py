"""Host emulator of the control-cycle code (TEST INFRASTRUCTURE ONLY).

tests/cpp/shc_emu.cpp compiles the very header the CUDA kernel is built from (csrc/shc_cycle.cuh) for the host with plain
g++ and runs the lanes of a tile one after the other.  The CPU test-suite uses it to check the cycle source against the
oracle in this GPU-less container; `-m gpu` tests check the real kernel on the B200 through the C-ABI.  Nothing in the
product package imports this module, and the product has no CPU path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from syropod_highlevel_controller_b200.config import (ShcBodyMsg, ShcConfig, ShcJointStateMsg, ShcLegStateMsg, ShcRobotState,
                                                      ShcStartup)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpp", "shc_emu.cpp")
_LIB = os.path.join(_HERE, "cpp", "_build", "libshc_emu.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "syropod_highlevel_controller_b200", "csrc")
_INC = os.path.join(os.path.dirname(_HERE), "include")
_lib = None
PRECISION = {"f64": 0, "mixed": 1}


def build(force: bool = False) -> str:
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_INC, f) for f in os.listdir(_INC)]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps)
    if stale:
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", _LIB, _SRC])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, fp = C.c_void_p, C.POINTER(C.c_float)
        L.shc_emu_last_error.restype = C.c_char_p
        L.shc_emu_create.argtypes = [C.POINTER(ShcConfig), C.POINTER(ShcStartup), C.c_int, C.c_int, C.POINTER(vp)]
        L.shc_emu_destroy.argtypes = [vp]
        L.shc_emu_destroy.restype = None
        L.shc_emu_get_startup.argtypes = [vp, C.POINTER(ShcStartup)]
        L.shc_emu_set_options.argtypes = [vp, C.c_int]
        L.shc_emu_set_pose_reset_mode.argtypes = [vp, C.c_int]
        L.shc_emu_set_joint_efforts.argtypes = [vp, fp]
        L.shc_emu_set_tip_step_planes.argtypes = [vp, fp]
        L.shc_emu_get_status_flags.argtypes = [vp, C.POINTER(C.c_int)]
        L.shc_emu_get_state.argtypes = [vp, C.POINTER(ShcRobotState), C.c_size_t]
        L.shc_emu_set_state.argtypes = [vp, C.POINTER(ShcRobotState), C.c_size_t]
        L.shc_emu_step.argtypes = [vp, fp, fp, fp, fp, fp]
        L.shc_emu_sequence_reset.argtypes = [vp]
        L.shc_emu_sequence_step.argtypes = [vp, C.c_int, C.c_double, fp, C.POINTER(C.c_int)]
        L.shc_emu_pack_messages.argtypes = [vp, C.c_int, fp, C.POINTER(ShcJointStateMsg), C.POINTER(ShcLegStateMsg), C.POINTER(ShcBodyMsg)]
        _lib = L
    return _lib


class EmuError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise EmuError(f"shc_emu error {rc}: {lib().shc_emu_last_error().decode()}")


def _f32(a, shape):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.shape == shape, (a.shape, shape)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class EmuEngine:
    """Same surface as syropod_highlevel_controller_b200.engine.Engine, numpy in / numpy out."""

    def __init__(self, cfg, n_robots, precision="f64", startup=None):
        self.cfg, self.n, self.L, self.D = cfg, int(n_robots), cfg.leg_count, cfg.joint_count
        self._h = C.c_void_p()
        _check(lib().shc_emu_create(C.byref(cfg), C.byref(startup) if startup is not None else None, self.n,
                                    PRECISION[precision], C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().shc_emu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def startup(self):
        s = ShcStartup()
        _check(lib().shc_emu_get_startup(self._h, C.byref(s)))
        return s

    def set_options(self, o):
        _check(lib().shc_emu_set_options(self._h, o))

    def set_pose_reset_mode(self, m):
        _check(lib().shc_emu_set_pose_reset_mode(self._h, m))

    def set_joint_efforts(self, eff):
        a, p = _f32(eff, (self.n, self.L, self.D))
        _check(lib().shc_emu_set_joint_efforts(self._h, p))

    def set_tip_step_planes(self, sp):
        a, p = _f32(sp, (self.n, self.L, 3))
        _check(lib().shc_emu_set_tip_step_planes(self._h, p))

    def status_flags(self):
        out = np.empty(self.n, dtype=np.int32)
        _check(lib().shc_emu_get_status_flags(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def get_state(self):
        arr = (ShcRobotState * self.n)()
        _check(lib().shc_emu_get_state(self._h, arr, self.n))
        return arr

    def set_state(self, arr):
        _check(lib().shc_emu_set_state(self._h, arr, self.n))

    def pack_messages(self, first=0, count=None, measured_joint_positions=None):
        count = self.n - first if count is None else count
        m, pm = _f32(measured_joint_positions, (self.n, self.L, self.D))
        js, legs, body = (ShcJointStateMsg * count)(), ((ShcLegStateMsg * self.L) * count)(), (ShcBodyMsg * count)()
        for k in range(count):
            _check(lib().shc_emu_pack_messages(self._h, first + k, pm, C.byref(js[k]), legs[k], C.byref(body[k])))
        return js, legs, body

    def sequence_reset(self):
        _check(lib().shc_emu_sequence_reset(self._h))

    def sequence_step(self, kind, time=0.0):
        """One loop() of stepToNewStance ("new_stance") / packLegs ("pack") / unpackLegs ("unpack"): (joints [n, L, D], progress [n])."""
        out = np.zeros((self.n, self.L, self.D), dtype=np.float32)
        prog = np.zeros(self.n, dtype=np.int32)
        _check(lib().shc_emu_sequence_step(self._h, {"new_stance": 0, "pack": 1, "unpack": 2, "start_up": 3, "shut_down": 4}[kind], float(time),
                                           out.ctypes.data_as(C.POINTER(C.c_float)), prog.ctypes.data_as(C.POINTER(C.c_int))))
        return out.astype(np.float64), prog

    def step(self, cmd, imu=None, tip_force=None, manual=None):
        cmd, pc = _f32(cmd, (self.n, 3))
        imu, pi_ = _f32(imu, (self.n, 10))
        tip_force, pf = _f32(tip_force, (self.n, self.L, 3))
        manual, pm = _f32(manual, (self.n, 6))
        out = np.empty((self.n, self.L, self.D), dtype=np.float32)
        _check(lib().shc_emu_step(self._h, pc, pi_, pf, pm, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out
