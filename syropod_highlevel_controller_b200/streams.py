"""Synthetic input streams of SURVEY.md §8(d): per-robot body-velocity commands, IMU and tip-force signals.

Commands: piecewise-constant segments of U{100..400} cycles, (vx, vy) uniform in the unit disc, wz in U[-1, 1], 20 % of
the segments all-zero; one splitmix64 generator per robot seeded with 0x5EED0000 + robot_id, so a robot's stream does
not depend on the batch size or on how the batch is sharded across GPUs.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0000
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(state: np.ndarray):
    """Vectorised splitmix64: advances `state` in place and returns the next 64-bit outputs."""
    with np.errstate(over="ignore"):
        state += np.uint64(0x9E3779B97F4A7C15)
        z = state.copy()
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _u01(state: np.ndarray) -> np.ndarray:
    return (_splitmix64(state) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class CommandStream:
    """next() -> float32 [n, 3] body velocity commands for robots robot_offset .. robot_offset + n - 1."""

    def __init__(self, n: int, robot_offset: int = 0, seed_base: int = SEED_BASE, zero_fraction: float = 0.2,
                 min_len: int = 100, max_len: int = 400):
        self.n = n
        self.state = (np.arange(n, dtype=np.uint64) + np.uint64(robot_offset) + np.uint64(seed_base)).astype(np.uint64)
        self.remaining = np.zeros(n, dtype=np.int64)
        self.cmd = np.zeros((n, 3), dtype=np.float32)
        self.zero_fraction, self.min_len, self.max_len = zero_fraction, min_len, max_len

    def next(self) -> np.ndarray:
        new = self.remaining <= 0
        if new.any():
            idx = np.nonzero(new)[0]
            st = self.state[idx]
            length = self.min_len + (_u01(st) * (self.max_len - self.min_len + 1)).astype(np.int64)
            zero = _u01(st) < self.zero_fraction
            rad = np.sqrt(_u01(st))
            ang = 2.0 * np.pi * _u01(st)
            wz = 2.0 * _u01(st) - 1.0
            c = np.stack([rad * np.cos(ang), rad * np.sin(ang), wz], axis=1)
            c[zero] = 0.0
            self.state[idx] = st
            self.cmd[idx] = c.astype(np.float32)
            self.remaining[idx] = np.minimum(length, self.max_len)
        self.remaining -= 1
        return self.cmd.copy()


class ImuStream:
    """roll/pitch = 0.1 sin(2 pi t / 500 + phi_r) rad, gyro = analytic derivative.

    next(dt) -> float32 [n, 10]: quaternion (w,x,y,z) of R = Rz(0) Ry(pitch) Rx(roll), gyro (3), accel (3)."""

    def __init__(self, n: int, robot_offset: int = 0, seed_base: int = SEED_BASE ^ 0x1111):
        st = (np.arange(n, dtype=np.uint64) + np.uint64(robot_offset) + np.uint64(seed_base)).astype(np.uint64)
        self.phi_r = 2.0 * np.pi * _u01(st)
        self.phi_p = 2.0 * np.pi * _u01(st)
        self.t = 0

    def next(self, dt: float) -> np.ndarray:
        w = 2.0 * np.pi / 500.0
        roll = 0.1 * np.sin(w * self.t + self.phi_r)
        pitch = 0.1 * np.sin(w * self.t + self.phi_p)
        droll = 0.1 * w * np.cos(w * self.t + self.phi_r) / dt
        dpitch = 0.1 * w * np.cos(w * self.t + self.phi_p) / dt
        self.t += 1
        cr, sr, cp, sp = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2)
        q = np.stack([cp * cr, cp * sr, sp * cr, -sp * sr], axis=1)  # q = qy(pitch) * qx(roll)
        out = np.zeros((len(roll), 10), dtype=np.float32)
        out[:, 0:4] = q
        out[:, 4] = droll
        out[:, 5] = dpitch
        out[:, 9] = 9.81
        return out


class ForceStream:
    """Tip force z ~ U[0, 10] N per leg (masked by `stance_mask` [n, L] when given); x, y = 0."""

    def __init__(self, n: int, legs: int, robot_offset: int = 0, seed_base: int = SEED_BASE ^ 0x2222):
        self.n, self.L = n, legs
        self.state = (np.arange(n, dtype=np.uint64) + np.uint64(robot_offset) + np.uint64(seed_base)).astype(np.uint64)

    def next(self, stance_mask=None) -> np.ndarray:
        out = np.zeros((self.n, self.L, 3), dtype=np.float32)
        for l in range(self.L):
            out[:, l, 2] = (10.0 * _u01(self.state)).astype(np.float32)
        if stance_mask is not None:
            out[:, :, 2] *= stance_mask.astype(np.float32)
        return out
