"""Helpers shared by the `-m gpu` parity tests (the oracle is the checker, the CUDA path is what is checked)."""
import numpy as np

STATE_FIELDS_ROBOT = ("desired_linear_velocity", "walk_plane", "walk_plane_normal", "odometry_ideal", "walk_plane_pose",
                      "origin_walk_plane_pose", "manual_pose", "imu_pose", "inclination_pose", "auto_pose",
                      "rotation_absement_error", "rotation_position_error", "rotation_velocity_error", "current_pose",
                      "tip_align_pose", "origin_tip_align_pose")
INT_FIELDS_ROBOT = ("walk_state", "legs_at_correct_phase", "legs_completed_first_step", "return_to_default_attempted",
                    "pose_state", "auto_posing_state", "pose_phase")
STATE_FIELDS_LEG = ("tip_position", "tip_velocity", "swing_origin_position", "stance_origin_position",
                    "default_tip_position", "target_tip_position", "stride_vector", "walk_plane", "walk_plane_normal",
                    "admittance_state", "admittance_delta", "tip_force_calculated", "virtual_stiffness", "model_tip_position")
# LegStepper tip rotations: only state of an engine with gravity_aligned_tips on legs of more than three joints (the
# reference keeps updating origin_tip_pose_.rotation_ in every configuration, but nothing reads it elsewhere): compared
# when the record under test carries a target rotation
ROTATION_FIELDS_LEG = ("tip_rotation", "origin_tip_rotation", "target_tip_rotation")
INT_FIELDS_LEG = ("phase", "step_state", "at_correct_phase", "completed_first_step", "negate_auto_pose")
# rough-terrain state: compared when the record under test is in touchdown-detection mode
ROUGH_FIELDS_LEG = ("step_plane_position", "external_target_pose", "external_target_transform", "external_target_clearance",
                    "external_default_pose", "external_default_transform")
ROUGH_INT_FIELDS_LEG = ("step_plane_defined", "touchdown_detection", "external_target_defined", "external_target_odom_frame",
                        "external_default_defined")


def state_diff(se, so, L, D):
    """Max abs difference per field between two shc_robot_state arrays; integer fields must match exactly."""
    out = {}

    def upd(k, a, b):
        d = float(np.max(np.abs(np.asarray(a, dtype=float) - np.asarray(b, dtype=float))))
        out[k] = max(out.get(k, 0.0), d)

    for r in range(len(se)):
        a, b = se[r], so[r]
        for f in STATE_FIELDS_ROBOT:
            upd(f, list(getattr(a, f)), list(getattr(b, f)))
        upd("desired_angular_velocity", a.desired_angular_velocity, b.desired_angular_velocity)
        for f in INT_FIELDS_ROBOT:
            upd("int:" + f, getattr(a, f), getattr(b, f))
        for l in range(L):
            la, lb = a.legs[l], b.legs[l]
            upd("joint_position", list(la.joint_position)[:D], list(lb.joint_position)[:D])
            upd("joint_velocity", list(la.joint_velocity)[:D], list(lb.joint_velocity)[:D])
            with_rot = any(v != 0.0 for v in la.target_tip_rotation)
            rough = bool(la.touchdown_detection)
            for f in STATE_FIELDS_LEG + (ROTATION_FIELDS_LEG if with_rot else ()) + (ROUGH_FIELDS_LEG if rough else ()):
                va, vb = getattr(la, f), getattr(lb, f)
                upd(f, va if isinstance(va, float) else list(va), vb if isinstance(vb, float) else list(vb))
            upd("swing_progress", la.swing_progress, lb.swing_progress)
            upd("stance_progress", la.stance_progress, lb.stance_progress)
            for f in INT_FIELDS_LEG + (ROUGH_INT_FIELDS_LEG if rough else ()):
                upd("int:" + f, getattr(la, f), getattr(lb, f))
    return out


JOINT_FIELDS = ("joint_position", "joint_velocity", "model_tip_position")


def assert_state_close(se, so, L, D, tol, vel_tol=None, skip=()):
    """All fields within tol (integers exact).  Free-running rollouts pass skip=JOINT_FIELDS: the joint state is judged
    by JointErrors instead (see there); every other field is open-loop and must match to tol at all times."""
    d = state_diff(se, so, L, D)
    vel_tol = tol * 100 if vel_tol is None else vel_tol
    for k, v in d.items():
        if k in skip:
            continue
        if k.startswith("int:"):
            assert v == 0, (k, v)
        elif k in ("joint_velocity", "tip_velocity"):
            assert v <= vel_tol, (k, v)
        else:
            assert v <= tol, (k, v)
    return d


class JointErrors:
    """Joint-angle differences of a free-running rollout.

    The reference's joint update is not a contraction everywhere: the joint-limit cost gradient of Leg::solveIK is
    normalised (model.cpp:788-790), so whenever a leg's joint velocity gets small the null-space term turns into a
    fixed-magnitude push against the previous motion and the joints settle into a period-2 limit cycle of ~1e-3 rad
    (every standing robot does this).  While a leg enters that regime, differences in the last bit are multiplied by
    ~4 per cycle until they saturate at the limit-cycle amplitude, then die out again once the leg moves.  Two correct
    double-precision implementations (or the reference built with two compilers) therefore agree to ~1e-9 rad almost
    always and differ by up to the chatter amplitude in rare short windows (DESIGN.md "Reference dynamics").  Hence:
    every difference must stay below CHATTER_BOUND, and all but a small fraction below the 1e-6 rad tolerance."""

    CHATTER_BOUND = 6e-3

    def __init__(self):
        self.count = 0
        self.exceed = 0
        self.worst = 0.0

    def add(self, diff, tol=1e-6):
        self.count += diff.size
        self.exceed += int(np.sum(diff > tol))
        self.worst = max(self.worst, float(diff.max()))

    @property
    def exceed_fraction(self):
        return self.exceed / max(self.count, 1)

    def check(self, max_fraction, label=""):
        # printed so that the tail of a GPU test run records what was observed against the cap
        print(f"[joint-errors] {label}: worst {self.worst:.3e} rad, {self.exceed}/{self.count} samples beyond 1e-6 "
              f"(fraction {self.exceed_fraction:.2e}, cap {max_fraction:.1e})")
        assert self.worst <= self.CHATTER_BOUND, self.worst
        assert self.exceed_fraction <= max_fraction, (self.exceed_fraction, self.exceed, self.count, self.worst)


def run_both(eng, ob, cycles, cmd_stream, imu_stream=None, force_stream=None, dt=0.02, threads=8, per_cycle=None):
    """Steps the engine under test (a tests/backends.py stepper: numpy in, float64 joints out) and the oracle on identical
    inputs; returns JointErrors over all cycles/joints (float32 output vs oracle double).  per_cycle(c, joints, oracle) is
    called after every cycle when given."""
    errs = JointErrors()
    for c in range(cycles):
        cmd = cmd_stream.next()
        imu = imu_stream.next(dt) if imu_stream is not None else None
        force = force_stream.next() if force_stream is not None else None
        jg = eng.step(cmd, imu, force)
        ob.step(cmd.astype(np.float64), None if imu is None else imu.astype(np.float64),
                None if force is None else force.astype(np.float64), threads=threads)
        errs.add(np.abs(jg - ob.joints()))
        if per_cycle is not None:
            per_cycle(c, jg, ob)
    return errs
