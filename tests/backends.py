"""The things the parity cases (tests/parity_cases.py) can be run against, behind one numpy-in / numpy-out surface:

  * GpuBackend — the product: the CUDA control-cycle kernel through the C-ABI (`-m gpu` tests, needs a B200);
  * EmuBackend — tests/emu.py: the same cycle SOURCE compiled for the host (CPU tests in this GPU-less container);
  * RefBackend — oracle/_ref: the REFERENCE'S OWN control code (its unmodified sources compiled against stand-in ROS /
    Eigen / Boost headers, oracle/ref_py.py), one StateController per robot.  Running the cases with this backend checks
    the restated oracle against the reference itself (tests/test_reference_pin.py): it is what pins the oracle.

The oracle is the checker in the first two cases; neither backend touches it."""
import numpy as np


class _GpuStepper:
    def __init__(self, cfg, n, precision, startup):
        import torch
        from syropod_highlevel_controller_b200.engine import Engine

        self.torch = torch
        self.eng = Engine(cfg, n, precision=precision, startup=startup)
        self.n = n

    def _dev(self, a):
        return None if a is None else self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()

    def step(self, cmd, imu=None, tip_force=None, manual=None):
        j = self.eng.step(self._dev(cmd), self._dev(imu), self._dev(tip_force), self._dev(manual))
        return j.cpu().numpy().astype(np.float64)

    def set_joint_efforts(self, eff):
        self.eng.set_joint_efforts(None if eff is None else self._dev(eff))

    def set_tip_step_planes(self, sp):
        self.eng.set_tip_step_planes(None if sp is None else self._dev(sp))

    def pack_messages(self, first=0, count=None, measured_joint_positions=None):
        return self.eng.pack_messages(first, count, None if measured_joint_positions is None else self._dev(measured_joint_positions))

    def __getattr__(self, name):  # get_state, set_state, set_options, set_pose_reset_mode, status_flags, startup, close
        return getattr(self.eng, name)


class _EmuStepper:
    def __init__(self, cfg, n, precision, startup):
        from emu import EmuEngine

        self.eng = EmuEngine(cfg, n, precision=precision, startup=startup)
        self.n = n

    def step(self, cmd, imu=None, tip_force=None, manual=None):
        return self.eng.step(cmd, imu, tip_force, manual).astype(np.float64)

    def __getattr__(self, name):
        return getattr(self.eng, name)


class _RefStepper:
    """n independent instances of the reference's StateController (each through its own direct start-up)."""

    def __init__(self, cfg, n, precision, startup):
        from oracle import ref_py

        assert precision == "f64"
        self.cfg, self.n = cfg, n
        self.L, self.D = cfg.leg_count, cfg.joint_count
        self.robots = [ref_py.RefRobot(cfg) for _ in range(n)]
        self._planes = None

    @staticmethod
    def _row(a, i):
        return None if a is None else np.asarray(a[i], dtype=np.float64)

    def step(self, cmd, imu=None, tip_force=None, manual=None):
        out = np.empty((self.n, self.L, self.D))
        for i, r in enumerate(self.robots):
            r.step(self._row(cmd, i), self._row(imu, i), self._row(tip_force, i), self._row(manual, i), self._row(self._planes, i))
            out[i] = r.joints()
        return out

    def get_state(self):
        from syropod_highlevel_controller_b200.config import ShcRobotState

        arr = (ShcRobotState * self.n)()
        for i, r in enumerate(self.robots):
            arr[i] = r.get_state()
        return arr

    def startup(self):
        return self.robots[0].startup()

    def set_pose_reset_mode(self, mode):
        for r in self.robots:
            r.set_pose_reset_mode(mode)

    def set_joint_efforts(self, eff):
        for i, r in enumerate(self.robots):
            r.set_joint_efforts(np.asarray(eff[i], dtype=np.float64))

    def set_tip_step_planes(self, sp):
        if sp is None:
            self._planes = None
            return
        sp = np.array(sp, dtype=np.float64)
        sp[..., 2] = np.where(sp[..., 2] >= 1e9, float(2 ** 31 - 1), sp[..., 2])  # the reference's UNASSIGNED_VALUE
        self._planes = sp

    def set_options(self, *a, **k):
        pass

    def set_joint_state_from(self, states):
        """Joint positions / velocities of shc_robot_state records (the oracle's) into the reference's joints."""
        for i, r in enumerate(self.robots):
            pos = np.array([list(states[i].legs[l].joint_position)[:self.D] for l in range(self.L)])
            vel = np.array([list(states[i].legs[l].joint_velocity)[:self.D] for l in range(self.L)])
            r.set_joint_state(pos, vel)

    def sequence_step(self, kind, time=0.0):
        done = getattr(self, "_seq_done", None)
        if done is None:
            done = self._seq_done = [None] * self.n
        prog = np.zeros(self.n, dtype=np.int32)
        out = np.empty((self.n, self.L, self.D))
        for i, r in enumerate(self.robots):
            if kind in ("start_up", "shut_down") and done[i] == kind:
                prog[i] = 100  # through this sequence: StateController no longer calls executeSequence (state_controller.cpp:314-350)
            else:
                prog[i] = r.sequence_step(kind, time)
                if kind in ("start_up", "shut_down"):
                    done[i] = kind if prog[i] == 100 else None
            out[i] = r.joints()
        return out, prog

    def assert_failures(self):
        return self.robots[0].assert_failures()

    def close(self):
        for r in self.robots:
            r.close()
        self.robots = []


class Backend:
    def __init__(self, kind):
        assert kind in ("gpu", "emu", "ref")
        self.kind = kind

    def engine(self, cfg, n, precision="f64", startup=None):
        return {"gpu": _GpuStepper, "emu": _EmuStepper, "ref": _RefStepper}[self.kind](cfg, n, precision, startup)


class RefOracle:
    """The reference's own code behind the surface the parity cases expect of the ORACLE module (`OracleBatch`): lets a
    case compare the engine under test with the reference directly instead of through the restated oracle.  Cases that
    need the oracle's state import (set_state) cannot use it."""

    class OracleBatch:
        def __init__(self, cfg, n_robots=1):
            self._s = _RefStepper(cfg, n_robots, "f64", None)
            self.cfg, self.n = cfg, n_robots
            self.L, self.D = cfg.leg_count, cfg.joint_count
            self._j = np.stack([r.joints() for r in self._s.robots])

        def step(self, cmd, imu=None, tip_force=None, manual=None, threads=1):
            self._j = self._s.step(cmd, imu, tip_force, manual)

        def joints(self):
            return self._j

        def get_state(self):
            return self._s.get_state()

        def startup(self):
            return self._s.startup()

        @property
        def startup_loops(self):
            return self._s.robots[0].startup_loops

        def set_pose_reset_mode(self, mode):
            self._s.set_pose_reset_mode(mode)

        def set_joint_efforts(self, eff):
            self._s.set_joint_efforts(eff)

        def set_tip_step_planes(self, sp):
            self._s.set_tip_step_planes(sp)

        def close(self):
            self._s.close()
