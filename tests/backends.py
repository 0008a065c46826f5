"""The two things the parity cases (tests/parity_cases.py) can be run against, behind one numpy-in / numpy-out surface:

  * GpuBackend — the product: the CUDA control-cycle kernel through the C-ABI (`-m gpu` tests, needs a B200);
  * EmuBackend — tests/emu.py: the same cycle SOURCE compiled for the host (CPU tests in this GPU-less container).

The oracle is the checker in both cases; neither backend touches it."""
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


class Backend:
    def __init__(self, kind):
        assert kind in ("gpu", "emu")
        self.kind = kind

    def engine(self, cfg, n, precision="f64", startup=None):
        return (_GpuStepper if self.kind == "gpu" else _EmuStepper)(cfg, n, precision, startup)
