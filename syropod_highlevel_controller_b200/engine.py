"""ctypes binding of libshc_b200.so — the Python face of the C-ABI in include/shc_b200.h.

PyTorch provides device buffers, streams and (in bench.py / multi-GPU runs) torch.distributed; every computation
happens in the hand-written CUDA kernels behind the C-ABI.  There is no CPU or eager fallback: if the shared library
is missing or no CUDA device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .config import ShcBodyMsg, ShcConfig, ShcJointStateMsg, ShcLegStateMsg, ShcRobotState, ShcStartup

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHC_B200_LIB", os.path.join(_PKG, "libshc_b200.so"))  # override: kernel-tuning builds only
_lib = None

PRECISION = {"f64": 0, "mixed": 1}
OPT_STATUS_FLAGS = 1
FLAG_IK_DEVIATION, FLAG_POSITION_CLAMP, FLAG_VELOCITY_CLAMP, FLAG_IMU_UNSTABLE = 1, 2, 4, 8


class ShcError(RuntimeError):
    pass


def lib():
    """Load libshc_b200.so (built in-tree by `python -m syropod_highlevel_controller_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ShcError(f"{LIB_PATH} is missing: build the CUDA extension first "
                           "(python -m syropod_highlevel_controller_b200.build); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        fp, vp = C.POINTER(C.c_float), C.c_void_p
        L.shc_create.argtypes = [C.POINTER(ShcConfig), C.POINTER(ShcStartup), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.shc_destroy.argtypes = [vp]
        L.shc_destroy.restype = None
        L.shc_last_error.restype = C.c_char_p
        L.shc_compute_startup.argtypes = [C.POINTER(ShcConfig), C.POINTER(ShcStartup)]
        dp = C.POINTER(C.c_double)
        L.shc_host_apply_ik.argtypes = [C.POINTER(ShcConfig), C.c_int, dp, dp, dp, C.c_int, dp, dp]
        L.shc_get_startup.argtypes = [vp, C.POINTER(ShcStartup)]
        L.shc_pack_messages.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp, vp, vp, vp]
        L.shc_startup_begin.argtypes = [vp, vp]
        L.shc_startup_step.argtypes = [vp, vp, vp]
        L.shc_direct_startup.argtypes = [vp, vp, vp, vp]
        L.shc_set_tip_step_planes.argtypes = [vp, vp]
        L.shc_set_state_range.argtypes = [vp, C.c_size_t, C.c_size_t, C.POINTER(ShcRobotState)]
        L.shc_clone_reconfigured.argtypes = [vp, C.POINTER(ShcConfig), C.POINTER(ShcStartup), C.c_int, C.POINTER(vp)]
        L.shc_step_to_new_stance.argtypes = [vp, vp, vp, vp]
        L.shc_sequence_reset.argtypes = [vp]
        L.shc_execute_sequence.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_int), vp]
        L.shc_transition_begin.argtypes = [vp, C.POINTER(C.c_double), C.c_double]
        L.shc_transition_step.argtypes = [vp, vp, vp]
        L.shc_pack_legs.argtypes = [vp, C.c_double, vp, vp]
        L.shc_unpack_legs.argtypes = [vp, C.c_double, vp, vp]
        L.shc_generate_workspaces.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int)]
        L.shc_host_generate_workspaces.argtypes = [C.POINTER(ShcConfig), C.POINTER(ShcStartup), C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int)]
        L.shc_n_robots.argtypes = [vp]
        L.shc_options.argtypes = [vp]
        L.shc_set_options.argtypes = [vp, C.c_int]
        L.shc_set_pose_reset_mode.argtypes = [vp, C.c_int]
        L.shc_get_state.argtypes = [vp, C.POINTER(ShcRobotState), C.c_size_t]
        L.shc_set_state.argtypes = [vp, C.POINTER(ShcRobotState), C.c_size_t]
        L.shc_get_state_range.argtypes = [vp, C.c_size_t, C.c_size_t, C.POINTER(ShcRobotState)]
        L.shc_set_limit_maps.argtypes = [vp, dp, dp, dp, dp]
        L.shc_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.shc_step_host.argtypes = [vp, fp, fp, fp, fp, fp]
        L.shc_rollout.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.shc_set_joint_efforts.argtypes = [vp, vp]
        L.shc_nccl_unique_id.argtypes = [vp]
        L.shc_nccl_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.shc_allgather_joints.argtypes = [vp, vp, vp, vp]
        L.shc_rollout_allgather.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.shc_gather_bytes.argtypes = [vp, C.c_int]
        L.shc_gather_bytes.restype = C.c_size_t
        L.shc_gather_attach.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), vp]
        L.shc_gather_step.argtypes = [vp, vp, vp, vp, vp, vp]
        L.shc_gather_sync.argtypes = [vp, C.POINTER(C.c_int), vp]
        L.shc_gather_status.argtypes = [vp]
        L.shc_gather_alloc.argtypes = [vp, vp, C.POINTER(vp)]
        L.shc_gather_open_peer.argtypes = [vp, C.c_int, vp]
        L.shc_rollout_gather_fused.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int), vp]
        L.shc_stream.argtypes = [vp]
        L.shc_stream.restype = vp
        L.shc_synchronize.argtypes = [vp]
        L.shc_status_flags_device.argtypes = [vp]
        L.shc_status_flags_device.restype = vp
        L.shc_get_status_flags.argtypes = [vp, C.POINTER(C.c_int)]
        L.shc_apply_ik.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp]
        for name in ("shc_sizeof_config", "shc_sizeof_startup", "shc_sizeof_robot_state"):
            getattr(L, name).restype = C.c_size_t
        for name in ("shc_bytes_per_step_device", "shc_bytes_per_step_algorithmic"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [vp]
        if L.shc_sizeof_config() != C.sizeof(ShcConfig) or L.shc_sizeof_startup() != C.sizeof(ShcStartup) or \
                L.shc_sizeof_robot_state() != C.sizeof(ShcRobotState):
            raise ShcError("ctypes struct layout does not match include/shc_config.h / shc_state.h")
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise ShcError(f"shc error {rc}: {lib().shc_last_error().decode()}")


def compute_startup(cfg: ShcConfig) -> ShcStartup:
    """Start-up constants (host arithmetic only, no GPU needed)."""
    s = ShcStartup()
    _check(lib().shc_compute_startup(C.byref(cfg), C.byref(s)))
    return s


def host_generate_workspaces(cfg: ShcConfig, full: bool = False, max_planes: int = 16, startup: Optional[ShcStartup] = None):
    """Leg::generateWorkspace (model.cpp:309-510) for every leg on the host with the routine the device sweep runs.
    Returns (heights [L, P], radii [L, P, 9], n_planes [L]); planes in generation order."""
    dp = C.POINTER(C.c_double)
    L = cfg.leg_count
    h, r, n = np.zeros((L, max_planes)), np.zeros((L, max_planes, 9)), np.zeros(L, dtype=np.int32)
    _check(lib().shc_host_generate_workspaces(C.byref(cfg), C.byref(startup) if startup is not None else None, int(full), max_planes,
                                              h.ctypes.data_as(dp), r.ctypes.data_as(dp), n.ctypes.data_as(C.POINTER(C.c_int))))
    return h, r, n


def host_apply_ik(cfg: ShcConfig, leg: int, q, qd, desired, simulation: bool = True):
    """One Leg::applyIK on the host with the kernels' own kinematics code (double).  Returns (q, qd, tip, result)."""
    dp = C.POINTER(C.c_double)
    q = np.array(q, dtype=np.float64)
    qd = np.array(qd, dtype=np.float64)
    des = np.ascontiguousarray(desired, dtype=np.float64)
    tip = np.empty(3)
    res = C.c_double()
    _check(lib().shc_host_apply_ik(C.byref(cfg), leg, q.ctypes.data_as(dp), qd.ctypes.data_as(dp), des.ctypes.data_as(dp),
                                   int(simulation), tip.ctypes.data_as(dp), C.byref(res)))
    return q, qd, tip, res.value


def _stream_handle(torch, device, stream):
    """cudaStream_t for the C-ABI: torch's current stream unless given.  torch's default stream is the legacy default
    stream, whose handle is 0 — spelled cudaStreamLegacy (0x1) here because NULL means "the engine's own stream"."""
    st = torch.cuda.current_stream(device).cuda_stream if stream is None else stream
    return C.c_void_p(st if st else 1)


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """N robots of one morphology resident on one B200.  One `step()` = one control cycle for every robot."""

    def __init__(self, cfg: ShcConfig, n_robots: int, device: int = 0, precision: str = "mixed",
                 startup: Optional[ShcStartup] = None):
        import torch

        if not torch.cuda.is_available():
            raise ShcError("no CUDA device: the SHC engine has no CPU fallback")
        self.torch = torch
        self.cfg = cfg
        self.n = int(n_robots)
        self.L, self.D = cfg.leg_count, cfg.joint_count
        self.device = torch.device("cuda", device)
        self.precision = precision
        self._h = C.c_void_p()
        _check(lib().shc_create(C.byref(cfg), C.byref(startup) if startup is not None else None, self.n, device,
                                PRECISION[precision], C.byref(self._h)))
        self.joints = torch.empty((self.n, self.L, self.D), dtype=torch.float32, device=self.device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().shc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reconfigured(self, cfg: ShcConfig, startup: Optional[ShcStartup] = None, keep_pose_cycle: bool = False) -> "Engine":
        """A NEW engine for `cfg` (same model, batch size, device, precision) that carries this engine's state: the batch-
        wide gait switch of StateController::changeGait (state_controller.cpp:513-540) or a change of a constants-only
        parameter.  Call it in place of the cycle in which the reference switches (that loop() updates no tips), with every
        robot STOPPED for a gait change, and close() this engine afterwards.  keep_pose_cycle: adjustParameter semantics (the
        auto-pose cycle length is not regenerated; the reference does that in changeGait only).  shc_clone_reconfigured in
        include/shc_b200.h."""
        new = Engine.__new__(Engine)
        new.torch, new.cfg, new.n = self.torch, cfg, self.n
        new.L, new.D, new.device, new.precision = self.L, self.D, self.device, self.precision
        new._h = C.c_void_p()
        _check(lib().shc_clone_reconfigured(self._h, C.byref(cfg), C.byref(startup) if startup is not None else None,
                                            1 if keep_pose_cycle else 0, C.byref(new._h)))
        new.joints = self.torch.empty((self.n, self.L, self.D), dtype=self.torch.float32, device=self.device)
        return new

    # ---- introspection -------------------------------------------------------------------------------------------
    def startup(self) -> ShcStartup:
        s = ShcStartup()
        _check(lib().shc_get_startup(self._h, C.byref(s)))
        return s

    @property
    def bytes_per_step_device(self) -> int:
        return lib().shc_bytes_per_step_device(self._h)

    @property
    def bytes_per_step_algorithmic(self) -> int:
        return lib().shc_bytes_per_step_algorithmic(self._h)

    def set_options(self, options: int):
        _check(lib().shc_set_options(self._h, options))

    def set_pose_reset_mode(self, mode: int):
        _check(lib().shc_set_pose_reset_mode(self._h, mode))

    def status_flags(self) -> np.ndarray:
        """int32 [N] status words of the last cycle (needs OPT_STATUS_FLAGS)."""
        out = np.empty(self.n, dtype=np.int32)
        _check(lib().shc_get_status_flags(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    # ---- output wire formats ----------------------------------------------------------------------------------------
    def pack_messages(self, first: int = 0, count: Optional[int] = None, measured_joint_positions=None):
        """JointState / LegState / body records (include/shc_msgs.h) of the robots [first, first + count), packed by one
        kernel straight into page-locked host memory.  Returns (joint_state [count], leg_state [count][L], body [count])
        as ctypes arrays viewing pinned memory owned by the engine object."""
        torch = self.torch
        count = self.n - first if count is None else count
        sizes = (C.sizeof(ShcJointStateMsg) * count, C.sizeof(ShcLegStateMsg) * count * self.L, C.sizeof(ShcBodyMsg) * count)
        bufs = [torch.empty(sz, dtype=torch.uint8, pin_memory=True) for sz in sizes]
        m = self._f32(measured_joint_positions, (self.n, self.L, self.D))
        _check(lib().shc_pack_messages(self._h, first, count, _ptr(m), bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(),
                                       _stream_handle(torch, self.device, None)))
        torch.cuda.synchronize(self.device)
        self._msg_keep = bufs
        js = (ShcJointStateMsg * count).from_address(bufs[0].data_ptr())
        legs = ((ShcLegStateMsg * self.L) * count).from_address(bufs[1].data_ptr())
        body = (ShcBodyMsg * count).from_address(bufs[2].data_ptr())
        return js, legs, body

    # ---- start-up on the device -------------------------------------------------------------------------------------
    def generate_workspaces(self, full: bool = False, max_planes: int = 16):
        """Leg::generateWorkspace for every leg on the device (one block per leg, eight lanes = eight bearings).
        Returns (heights [L, P], radii [L, P, 9], n_planes [L])."""
        dp = C.POINTER(C.c_double)
        h, r, n = np.zeros((self.L, max_planes)), np.zeros((self.L, max_planes, 9)), np.zeros(self.L, dtype=np.int32)
        _check(lib().shc_generate_workspaces(self._h, int(full), max_planes, h.ctypes.data_as(dp), r.ctypes.data_as(dp),
                                             n.ctypes.data_as(C.POINTER(C.c_int))))
        return h, r, n

    def startup_begin(self, joint_positions=None):
        """PoseController::directStartup begins: joint_positions [N, L, D] float64 device tensor (measured joint states) or None."""
        if joint_positions is not None:
            t = self.torch
            assert joint_positions.is_cuda and joint_positions.dtype == t.float64 and joint_positions.is_contiguous()
            assert tuple(joint_positions.shape) == (self.n, self.L, self.D)
        _check(lib().shc_startup_begin(self._h, _ptr(joint_positions)))

    def startup_step(self, out=None, stream=None) -> int:
        """One loop() of the start-up for the batch; joint commands into `out` (default: self.joints).  Returns the progress
        (100 = complete)."""
        out = self._out(out)
        rc = lib().shc_startup_step(self._h, _ptr(out), _stream_handle(self.torch, self.device, stream))
        if rc < 0:
            _check(rc)
        return rc

    def direct_startup(self, joint_positions=None, out=None, stream=None):
        out = self._out(out)
        _check(lib().shc_direct_startup(self._h, _ptr(joint_positions), _ptr(out), _stream_handle(self.torch, self.device, stream)))
        return out

    # ---- stepping / joint-space sequences --------------------------------------------------------
    def step_to_new_stance(self, out=None, progress_out=None, stream=None) -> int:
        """One loop() of PoseController::stepToNewStance (pose_controller.cpp:520) for the batch; joint commands into `out`
        (default: self.joints), per-robot progress into `progress_out` (int32 device tensor [N]) when given.  Returns the
        smallest progress over the batch."""
        out = self._out(out)
        rc = lib().shc_step_to_new_stance(self._h, _ptr(out), _ptr(progress_out), _stream_handle(self.torch, self.device, stream))
        if rc < 0:
            _check(rc)
        return rc

    def execute_sequence(self, shut_down: bool = False, out=None, progress_out=None, stream=None) -> int:
        """One loop() of PoseController::executeSequence (pose_controller.cpp:145) for the batch: the start-up (or shut-down)
        sequence.  Returns the smallest per-robot value: -1 while a first start-up generates its sequence, 0..100, -2 failed."""
        out = self._out(out)
        mn = C.c_int(0)
        _check(lib().shc_execute_sequence(self._h, int(bool(shut_down)), _ptr(out), _ptr(progress_out), C.byref(mn),
                                          _stream_handle(self.torch, self.device, stream)))
        return int(mn.value)

    def sequence_reset(self):
        _check(lib().shc_sequence_reset(self._h))

    def transition_begin(self, desired_configuration, transition_time: float):
        """PoseController::transitionConfiguration (pose_controller.cpp:703) begins: desired joint positions [L, D]."""
        d = np.ascontiguousarray(desired_configuration, dtype=np.float64)
        assert d.shape == (self.L, self.D)
        _check(lib().shc_transition_begin(self._h, d.ctypes.data_as(C.POINTER(C.c_double)), float(transition_time)))

    def transition_step(self, out=None, stream=None) -> int:
        out = self._out(out)
        rc = lib().shc_transition_step(self._h, _ptr(out), _stream_handle(self.torch, self.device, stream))
        if rc < 0:
            _check(rc)
        return rc

    def pack_legs(self, time_to_pack: float, out=None, stream=None) -> int:
        """One loop() of PoseController::packLegs (pose_controller.cpp:597); returns the progress (100 = complete)."""
        out = self._out(out)
        rc = lib().shc_pack_legs(self._h, float(time_to_pack), _ptr(out), _stream_handle(self.torch, self.device, stream))
        if rc < 0:
            _check(rc)
        return rc

    def unpack_legs(self, time_to_unpack: float, out=None, stream=None) -> int:
        """One loop() of PoseController::unpackLegs (pose_controller.cpp:661); returns the progress (100 = complete)."""
        out = self._out(out)
        rc = lib().shc_unpack_legs(self._h, float(time_to_unpack), _ptr(out), _stream_handle(self.torch, self.device, stream))
        if rc < 0:
            _check(rc)
        return rc

    def sequence_step(self, kind: str, time: float = 0.0):
        """Test-facing form shared with the host emulator: one loop() of "new_stance" / "pack" / "unpack";
        returns (joints float64 numpy [N, L, D], per-robot progress int32 numpy [N])."""
        t = self.torch
        if kind == "new_stance":
            prog = t.zeros(self.n, dtype=t.int32, device=self.device)
            self.step_to_new_stance(progress_out=prog)
        elif kind in ("start_up", "shut_down"):
            prog = t.zeros(self.n, dtype=t.int32, device=self.device)
            self.execute_sequence(kind == "shut_down", progress_out=prog)
        else:
            p = self.pack_legs(time) if kind == "pack" else self.unpack_legs(time)
            prog = t.full((self.n,), p, dtype=t.int32, device=self.device)
        t.cuda.synchronize(self.device)
        return self.joints.cpu().numpy().astype(np.float64), prog.cpu().numpy()

    # ---- state ------------------------------------------------------------------------------------------------------
    def get_state(self):
        arr = (ShcRobotState * self.n)()
        _check(lib().shc_get_state(self._h, arr, self.n))
        return arr

    def set_state(self, arr):
        _check(lib().shc_set_state(self._h, arr, self.n))

    def set_state_range(self, first: int, records):
        """Replaces the records of the robots [first, first + len(records))."""
        _check(lib().shc_set_state_range(self._h, first, len(records), records))

    def get_state_range(self, first: int, count: int = 1):
        """Records of the robots [first, first + count) only (three small copies, whatever the batch size)."""
        arr = (ShcRobotState * count)()
        _check(lib().shc_get_state_range(self._h, first, count, arr))
        return arr

    def set_limit_maps(self, max_linear_speed=None, max_angular_speed=None, max_linear_acceleration=None,
                       max_angular_acceleration=None):
        """WalkController::set*LimitMap (walk_controller.h:126-141): 9 values per table (bearings 0..360 step 45)."""
        dp = C.POINTER(C.c_double)
        arrs = [None if m is None else np.ascontiguousarray(m, dtype=np.float64) for m in
                (max_linear_speed, max_angular_speed, max_linear_acceleration, max_angular_acceleration)]
        assert all(a is None or a.shape == (9,) for a in arrs)
        _check(lib().shc_set_limit_maps(self._h, *[None if a is None else a.ctypes.data_as(dp) for a in arrs]))

    # ---- stepping --------------------------------------------------------------------------------------------------
    def _f32(self, t, shape):
        if t is None:
            return None
        torch = self.torch
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        assert tuple(t.shape) == shape, (tuple(t.shape), shape)
        return t

    def _out(self, out):
        """The [N, L, D] float32 device tensor a kernel writes the joint angles to (the kernel trusts this pointer)."""
        if out is None:
            return self.joints
        torch = self.torch
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == self.device and out.dtype == torch.float32
                and out.is_contiguous() and tuple(out.shape) == (self.n, self.L, self.D)):
            raise ShcError(f"out must be a contiguous float32 tensor of shape {(self.n, self.L, self.D)} on {self.device}")
        return out

    def _seq(self, t, per_cycle_shape, k):
        """A [k, ...] float32 per-cycle input sequence on this engine's device (or None)."""
        if t is None:
            return None
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device and t.dtype == torch.float32
                and t.is_contiguous() and tuple(t.shape) == (k,) + tuple(per_cycle_shape)):
            raise ShcError(f"sequence must be a contiguous float32 tensor of shape {(k,) + tuple(per_cycle_shape)} on {self.device}")
        return t

    def step(self, cmd, imu=None, tip_force=None, manual=None, out=None, stream=None):
        """One control cycle with device-resident inputs (torch CUDA tensors).  Asynchronous on `stream` (default:
        torch's current stream).  Returns the [N, L, D] float32 joint-angle tensor."""
        torch = self.torch
        cmd = self._f32(cmd, (self.n, 3))
        imu = self._f32(imu, (self.n, 10))
        tip_force = self._f32(tip_force, (self.n, self.L, 3))
        manual = self._f32(manual, (self.n, 6))
        out = self._out(out)
        _check(lib().shc_step(self._h, _ptr(cmd), _ptr(imu), _ptr(tip_force), _ptr(manual), _ptr(out),
                              _stream_handle(torch, self.device, stream)))
        self._keep = (cmd, imu, tip_force, manual)  # keep inputs alive until the launch has consumed them
        return out

    def pinned_host(self, *shape) -> np.ndarray:
        """A page-locked float32 host array (numpy view of a pinned torch tensor): shc_step_host moves such buffers by
        DMA directly instead of staging them."""
        t = self.torch.empty(shape, dtype=self.torch.float32, pin_memory=True)
        a = t.numpy()
        self._pinned = getattr(self, "_pinned", []) + [t]  # keep the owner alive
        return a

    def step_host(self, cmd, imu=None, tip_force=None, manual=None, out=None) -> np.ndarray:
        """One control cycle with HOST numpy buffers through the C-ABI (H2D, tile-range kernels overlapped with the D2H of
        the joint angles, sync).  Page-locked arrays (pinned_host) are transferred in place; others go through staging."""
        fp = C.POINTER(C.c_float)

        def h(a, shape):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.shape == shape
            return a, a.ctypes.data_as(fp)

        cmd, pc = h(cmd, (self.n, 3))
        imu, pi_ = h(imu, (self.n, 10))
        tip_force, pf = h(tip_force, (self.n, self.L, 3))
        manual, pm = h(manual, (self.n, 6))
        if out is None:
            out = np.empty((self.n, self.L, self.D), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (self.n, self.L, self.D) and out.flags["C_CONTIGUOUS"]
        _check(lib().shc_step_host(self._h, pc, pi_, pf, pm, out.ctypes.data_as(fp)))
        return out

    def rollout(self, cmd_seq, imu_seq=None, force_seq=None, out=None, stream=None):
        """k cycles with per-cycle device-resident inputs cmd_seq [k, N, 3]; one CUDA graph launch."""
        torch = self.torch
        k = int(cmd_seq.shape[0])
        cmd_seq = self._seq(cmd_seq, (self.n, 3), k)
        imu_seq = self._seq(imu_seq, (self.n, 10), k)
        force_seq = self._seq(force_seq, (self.n, self.L, 3), k)
        out = self._out(out)
        _check(lib().shc_rollout(self._h, k, _ptr(cmd_seq), _ptr(imu_seq), _ptr(force_seq), _ptr(out),
                                 _stream_handle(torch, self.device, stream)))
        self._keep = (cmd_seq, imu_seq, force_seq)
        return out

    # ---- multi-GPU ---------------------------------------------------------------------------------------------------
    def init_nccl(self, rank: int, world_size: int):
        """Joins this engine to an NCCL communicator over the ranks of the default torch.distributed group (the 128-byte
        unique id travels through torch.distributed; the collective itself is issued by the library)."""
        import torch.distributed as dist

        torch = self.torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _check(lib().shc_nccl_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(self.device)
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _check(lib().shc_nccl_init(self._h, C.c_char_p(raw), rank, world_size))
        self.world_size = world_size
        self._nccl_ready = True

    def rollout_allgather(self, cmd_seq, local2, full2, stream=None):
        """k cycles, each followed by the all-gather of its joint angles (overlapped, double buffered)."""
        torch = self.torch
        k = int(cmd_seq.shape[0])
        cmd_seq = self._seq(cmd_seq, (self.n, 3), k)
        per = self.n * self.L * self.D
        for t, cnt in ((local2, 2 * per), (full2, 2 * per * self.world_size)):
            if not (t.is_cuda and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == cnt):
                raise ShcError("rollout_allgather: local2 / full2 must be contiguous float32 [2, n, L, D] / [2, world * n, L, D]")
        _check(lib().shc_rollout_allgather(self._h, k, _ptr(cmd_seq), _ptr(local2), _ptr(full2),
                                           _stream_handle(torch, self.device, stream)))
        self._keep = (cmd_seq, local2, full2)

    def init_gather_fused(self, rank: int, world_size: int, mode: str = "auto"):
        """Sets up the fused all-gather over peer memory.  Collective: every rank must call it.

        mode "auto": torch symmetric memory (torch.distributed._symmetric_memory) allocates one buffer per rank, maps every
        rank's buffer on every rank and — when the fabric has NVLS — provides their NVSwitch multicast mapping, which the
        kernel then stores to (one multimem.st per 16 bytes, replicated by the switch).  If symmetric memory is not
        available on any rank, falls back to "ipc": engine-owned cudaMalloc buffers exchanged as CUDA-IPC handles (after
        init_nccl), unicast TMA bulk stores.  "symm" / "symm-unicast" / "ipc" force one path.  If no path works on every rank,
        every rank raises RuntimeError together (so that callers can fall back to the NCCL gather together).
        Returns the own buffer as a torch tensor [buffers, world, n, L, D]; self.gather_mode says which path is live."""
        import torch.distributed as dist

        torch = self.torch

        def agree(ok: bool) -> bool:
            flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return int(flag.item()) == 1

        nb = int(lib().shc_gather_buffers())
        shape = (nb, world_size, self.n, self.L, self.D)
        n_data = nb * world_size * self.n * self.L * self.D
        self.world_size = world_size
        if mode in ("auto", "symm", "symm-unicast"):
            ok, err, peers, mc, buf, hdl = True, "", [], 0, None, None
            try:
                import torch.distributed._symmetric_memory as symm_mem

                nbytes = int(lib().shc_gather_bytes(self._h, world_size))
                buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=self.device)
                hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
                peers = [int(p) for p in hdl.buffer_ptrs]
                mc = int(hdl.multicast_ptr) if mode != "symm-unicast" else 0
            except Exception as ex:  # noqa: BLE001 - any failure means "not available here"
                ok, err = False, f"{type(ex).__name__}: {ex}"
            if agree(ok):
                use_mc = agree(mc != 0)  # multicast only if every rank has the mapping
                arr = (C.c_void_p * world_size)(*peers)
                rc = lib().shc_gather_attach(self._h, rank, world_size, arr, C.c_void_p(mc if use_mc else 0))
                if not agree(rc == 0):
                    raise RuntimeError("fused gather: shc_gather_attach failed on at least one rank")
                dist.barrier()  # every rank's landed counters are zeroed before anyone's first cycle
                self._gather_keep = (buf, hdl)
                self.gather = buf[:n_data].view(shape)
                self.gather_mode = "multicast (NVLS multimem.st)" if use_mc else "unicast (TMA bulk stores, symmetric memory)"
                return self.gather
            if mode != "auto":
                raise RuntimeError(f"fused gather: torch symmetric memory unavailable on at least one rank ({err})")
        if not hasattr(self, "_nccl_ready"):
            self.init_nccl(rank, world_size)
        handle = (C.c_char * 64)()
        bufp = C.c_void_p()
        rc = lib().shc_gather_alloc(self._h, handle, C.byref(bufp))
        if not agree(rc == 0):
            raise RuntimeError("fused gather unavailable on at least one rank (shc_gather_alloc)")
        handles = [None] * world_size
        dist.all_gather_object(handles, bytes(handle.raw))
        ok = True
        for p, hb in enumerate(handles):
            if p != rank and lib().shc_gather_open_peer(self._h, p, C.c_char_p(hb)) != 0:
                ok = False
        if not agree(ok):
            raise RuntimeError("fused gather unavailable on at least one rank (shc_gather_open_peer)")
        dist.barrier()

        class _Dev:  # __cuda_array_interface__ view of the library's buffer
            __cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(bufp.value), False), "version": 2}

        self.gather = torch.as_tensor(_Dev(), device=self.device)
        self.gather_mode = "unicast (TMA bulk stores, CUDA IPC)"
        return self.gather

    def gather_step(self, cmd, imu=None, tip_force=None, manual=None, stream=None):
        """One control cycle whose joint commands land in every rank's gather buffer (no wait: see gather_sync)."""
        torch = self.torch
        cmd = self._f32(cmd, (self.n, 3))
        imu = self._f32(imu, (self.n, 10))
        tip_force = self._f32(tip_force, (self.n, self.L, 3))
        manual = self._f32(manual, (self.n, 6))
        _check(lib().shc_gather_step(self._h, _ptr(cmd), _ptr(imu), _ptr(tip_force), _ptr(manual),
                                     _stream_handle(torch, self.device, stream)))
        self._keep = (cmd, imu, tip_force, manual)

    def gather_sync(self, stream=None) -> int:
        """The stream continues once every rank's shard of every cycle issued so far has landed here.  Returns the index
        of the buffer that holds the last cycle."""
        last = C.c_int(-1)
        _check(lib().shc_gather_sync(self._h, C.byref(last), _stream_handle(self.torch, self.device, stream)))
        return int(last.value)

    def gather_status(self):
        """Raises ShcError if a device-side wait for a peer gave up (a rank stopped signalling)."""
        _check(lib().shc_gather_status(self._h))

    def rollout_gather_fused(self, cmd_seq, stream=None) -> int:
        """k cycles whose joint commands land in every rank's gather buffer from inside the kernel (stores over NVLink),
        then the wait for the last cycle.  Returns the index of the buffer that holds the last cycle."""
        torch = self.torch
        k = int(cmd_seq.shape[0])
        assert cmd_seq.is_cuda and cmd_seq.dtype == torch.float32 and cmd_seq.is_contiguous()
        assert tuple(cmd_seq.shape[1:]) == (self.n, 3) and cmd_seq.device == self.device
        last = C.c_int(-1)
        _check(lib().shc_rollout_gather_fused(self._h, k, _ptr(cmd_seq), C.byref(last), _stream_handle(torch, self.device, stream)))
        self._keep = (cmd_seq,)
        return int(last.value)

    def set_joint_efforts(self, efforts):
        efforts = self._f32(efforts, (self.n, self.L, self.D))
        self._efforts = efforts
        _check(lib().shc_set_joint_efforts(self._h, _ptr(efforts)))

    def set_tip_step_planes(self, step_planes):
        """Tip range-sensor readings [N, L, 3] (x, y slopes, z range; z >= 1e9 = no reading) for rough-terrain mode, or None."""
        step_planes = self._f32(step_planes, (self.n, self.L, 3))
        self._step_planes = step_planes
        _check(lib().shc_set_tip_step_planes(self._h, _ptr(step_planes)))

    def apply_ik(self, leg_id, q, qd, desired_tip, simulation: bool = True):
        """Stand-alone batched Leg::applyIK in double.  Returns (q, qd, tip, ik_result) as torch tensors."""
        torch = self.torch
        leg_id = torch.as_tensor(leg_id, dtype=torch.int32, device=self.device).contiguous()
        n = leg_id.numel()
        q = torch.as_tensor(q, dtype=torch.float64, device=self.device).clone().contiguous()
        qd = torch.as_tensor(qd, dtype=torch.float64, device=self.device).clone().contiguous()
        des = torch.as_tensor(desired_tip, dtype=torch.float64, device=self.device).contiguous()
        tip = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        res = torch.empty((n,), dtype=torch.float64, device=self.device)
        _check(lib().shc_apply_ik(self._h, n, _ptr(leg_id), _ptr(q), _ptr(qd), _ptr(des), int(simulation), _ptr(tip),
                                  _ptr(res), _stream_handle(torch, self.device, None)))
        return q, qd, tip, res

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)
        _check(lib().shc_synchronize(self._h))
