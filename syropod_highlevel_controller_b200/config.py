"""Parameter blocks for the batched SHC engine.

`ShcConfig` / `ShcStartup` / `ShcRobotState` mirror the C structs of include/shc_config.h and
include/shc_state.h field by field (ctypes).  `hexapod_config()` carries the values of the reference's
config/default.yaml + config/gait.yaml + config/auto_pose.yaml (cited per block); `octopod_config()` is the
synthetic 8-leg x 5-DOF robot of BASELINE.json configs[3] (no such config ships with the reference —
SURVEY.md §8).  `load_reference_yaml()` reads the reference's own three-file YAML schema, which is how a user
of the reference brings their robot over.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

MAX_LEGS = 8
MAX_DOF = 5
MAX_AUTO_POSERS = 8
N_BEARINGS = 9

_d = C.c_double
_i = C.c_int


class ShcConfig(C.Structure):
    _fields_ = [
        ("time_delta", _d),
        ("manual_posing", _i), ("auto_posing", _i), ("rough_terrain_mode", _i),
        ("admittance_control", _i), ("inclination_posing", _i), ("imu_posing", _i),
        ("leg_count", _i), ("joint_count", _i),
        ("joint_min", (_d * MAX_DOF) * MAX_LEGS),
        ("joint_max", (_d * MAX_DOF) * MAX_LEGS),
        ("joint_max_vel", (_d * MAX_DOF) * MAX_LEGS),
        ("joint_offset", (_d * MAX_DOF) * MAX_LEGS),
        ("link_d", (_d * (MAX_DOF + 1)) * MAX_LEGS),
        ("link_theta", (_d * (MAX_DOF + 1)) * MAX_LEGS),
        ("link_r", (_d * (MAX_DOF + 1)) * MAX_LEGS),
        ("link_alpha", (_d * (MAX_DOF + 1)) * MAX_LEGS),
        ("clamp_joint_positions", _i), ("clamp_joint_velocities", _i),
        ("body_clearance", _d), ("step_frequency", _d), ("swing_height", _d), ("swing_width", _d),
        ("step_depth", _d), ("stance_span_modifier", _d),
        ("velocity_input_mode", _i),
        ("body_velocity_scaler", _d),
        ("stance_x", _d * MAX_LEGS), ("stance_y", _d * MAX_LEGS),
        ("overlapping_walkspaces", _i), ("force_normal_touchdown", _i), ("gravity_aligned_tips", _i),
        ("stance_phase", _i), ("swing_phase", _i), ("phase_offset", _i),
        ("offset_multiplier", _i * MAX_LEGS),
        ("time_to_start", _d),
        ("rotation_pid_p", _d), ("rotation_pid_i", _d), ("rotation_pid_d", _d),
        ("max_translation", _d * 3), ("max_rotation", _d * 3),
        ("max_translation_velocity", _d), ("max_rotation_velocity", _d),
        ("pose_frequency", _d),
        ("pose_phase_length", _i), ("auto_poser_count", _i),
        ("pose_phase_starts", _i * MAX_AUTO_POSERS), ("pose_phase_ends", _i * MAX_AUTO_POSERS),
        ("pose_negation_phase_starts", _i * MAX_LEGS), ("pose_negation_phase_ends", _i * MAX_LEGS),
        ("negation_transition_ratio", _d * MAX_LEGS),
        ("x_amplitudes", _d * MAX_AUTO_POSERS), ("y_amplitudes", _d * MAX_AUTO_POSERS),
        ("z_amplitudes", _d * MAX_AUTO_POSERS), ("gravity_amplitudes", _d * MAX_AUTO_POSERS),
        ("roll_amplitudes", _d * MAX_AUTO_POSERS), ("pitch_amplitudes", _d * MAX_AUTO_POSERS),
        ("yaw_amplitudes", _d * MAX_AUTO_POSERS),
        ("dynamic_stiffness", _i), ("use_joint_effort", _i),
        ("integrator_step_time", _d), ("virtual_mass", _d), ("virtual_stiffness", _d),
        ("virtual_damping_ratio", _d), ("force_gain", _d),
        ("load_stiffness_scaler", _d), ("swing_stiffness_scaler", _d),
        ("joint_packed", (_d * MAX_DOF) * MAX_LEGS), ("joint_unpacked", (_d * MAX_DOF) * MAX_LEGS),
        ("touchdown_threshold", _d), ("liftoff_threshold", _d),
    ]


class ShcStartup(C.Structure):
    _fields_ = [
        ("default_joint", (_d * MAX_DOF) * MAX_LEGS),
        ("workspace", (_d * N_BEARINGS) * MAX_LEGS),
        ("walkspace", _d * N_BEARINGS),
        ("max_linear_speed", _d * N_BEARINGS),
        ("max_angular_speed", _d * N_BEARINGS),
        ("max_linear_acceleration", _d * N_BEARINGS),
        ("max_angular_acceleration", _d * N_BEARINGS),
        ("step_frequency", _d),
        ("period", _i), ("swing_period", _i), ("stance_period", _i), ("stance_end", _i),
        ("swing_start", _i), ("swing_end", _i), ("stance_start", _i),
        ("phase_offsets", _i * MAX_LEGS),
        ("pose_phase_length", _i), ("pose_normaliser", _i), ("auto_pose_reference_leg", _i),
        ("startup_loops", _i),
    ]


class ShcLegState(C.Structure):
    _fields_ = [
        ("joint_position", _d * MAX_DOF), ("joint_velocity", _d * MAX_DOF),
        ("tip_position", _d * 3), ("tip_velocity", _d * 3),
        ("swing_origin_position", _d * 3), ("swing_origin_velocity", _d * 3),
        ("stance_origin_position", _d * 3),
        ("default_tip_position", _d * 3), ("target_tip_position", _d * 3), ("stride_vector", _d * 3),
        ("walk_plane", _d * 3), ("walk_plane_normal", _d * 3),
        ("swing_progress", _d), ("stance_progress", _d),
        ("phase", _i), ("step_state", _i), ("at_correct_phase", _i), ("completed_first_step", _i),
        ("admittance_state", _d * 2), ("admittance_delta", _d * 3), ("tip_force_calculated", _d * 3),
        ("virtual_stiffness", _d),
        ("negate_auto_pose", _i), ("pad0", _i),
        ("tip_rotation", _d * 4), ("origin_tip_rotation", _d * 4), ("target_tip_rotation", _d * 4),
        ("step_plane_position", _d * 3), ("step_plane_defined", _i), ("touchdown_detection", _i),
        ("external_target_pose", _d * 7), ("external_target_transform", _d * 7), ("external_target_clearance", _d),
        ("external_target_defined", _i), ("external_target_odom_frame", _i),
        ("external_default_pose", _d * 7), ("external_default_transform", _d * 7),
        ("external_default_defined", _i), ("pad1", _i),
        ("model_tip_position", _d * 3), ("desired_tip_position", _d * 3), ("ik_result", _d),
    ]


class ShcRobotState(C.Structure):
    _fields_ = [
        ("desired_linear_velocity", _d * 2), ("desired_angular_velocity", _d),
        ("walk_state", _i), ("legs_at_correct_phase", _i), ("legs_completed_first_step", _i),
        ("return_to_default_attempted", _i), ("pose_state", _i), ("pad0", _i),
        ("walk_plane", _d * 3), ("walk_plane_normal", _d * 3),
        ("odometry_ideal", _d * 7),
        ("walk_plane_pose", _d * 7), ("origin_walk_plane_pose", _d * 7), ("manual_pose", _d * 7),
        ("imu_pose", _d * 7), ("inclination_pose", _d * 7), ("auto_pose", _d * 7),
        ("rotation_absement_error", _d * 3), ("rotation_position_error", _d * 3),
        ("rotation_velocity_error", _d * 3),
        ("tip_align_pose", _d * 7), ("origin_tip_align_pose", _d * 7),
        ("auto_posing_state", _i), ("pose_phase", _i),
        ("auto_poser_flags", _i * MAX_AUTO_POSERS),
        ("current_pose", _d * 7),
        ("status_flags", _i), ("pad1", _i),
        ("legs", ShcLegState * MAX_LEGS),
    ]


# include/shc_msgs.h ----------------------------------------------------------------------------------------------
class ShcJointStateMsg(C.Structure):
    _fields_ = [("position", _d * (MAX_LEGS * MAX_DOF)), ("velocity", _d * (MAX_LEGS * MAX_DOF)),
                ("effort", _d * (MAX_LEGS * MAX_DOF)), ("position_command", _d * (MAX_LEGS * MAX_DOF))]


class ShcLegStateMsg(C.Structure):
    _fields_ = [("walker_tip_pose", _d * 7), ("target_tip_pose", _d * 7), ("poser_tip_pose", _d * 7),
                ("model_tip_pose", _d * 7), ("actual_tip_pose", _d * 7), ("model_tip_velocity", _d * 3),
                ("joint_positions", _d * MAX_DOF), ("joint_velocities", _d * MAX_DOF), ("joint_efforts", _d * MAX_DOF),
                ("stance_progress", _d), ("swing_progress", _d), ("time_to_swing_end", _d),
                ("pose_delta", _d * 7), ("auto_pose", _d * 7), ("tip_force", _d * 3), ("admittance_delta", _d * 3),
                ("virtual_stiffness", _d), ("joint_transform", (_d * 7) * MAX_DOF), ("tip_transform", _d * 7)]


class ShcBodyMsg(C.Structure):
    _fields_ = [("velocity", _d * 6), ("pose", _d * 6), ("rotation_pose_error", _d * 9),
                ("odom_ideal_to_base_link", _d * 7), ("base_link_to_walk_plane", _d * 7)]


# --------------------------------------------------------------------------------------------------------------
# Gait tables — values of /root/reference/config/gait.yaml:13-50 (leg order AR,BR,CR,CL,BL,AL as default.yaml:26)
# --------------------------------------------------------------------------------------------------------------
HEXAPOD_LEGS = ["AR", "BR", "CR", "CL", "BL", "AL"]
GAITS: Dict[str, dict] = {
    "wave_gait": dict(stance_phase=10, swing_phase=2, phase_offset=2,
                      offset_multiplier=dict(AR=2, BR=3, CR=4, CL=1, BL=0, AL=5)),
    "tripod_gait": dict(stance_phase=2, swing_phase=2, phase_offset=2,
                        offset_multiplier=dict(AR=0, BR=1, CR=0, CL=1, BL=0, AL=1)),
    "ripple_gait": dict(stance_phase=4, swing_phase=2, phase_offset=1,
                        offset_multiplier=dict(AR=2, BR=0, CR=4, CL=1, BL=3, AL=5)),
    "amble_gait": dict(stance_phase=2, swing_phase=1, phase_offset=1,
                       offset_multiplier=dict(AR=1, BR=2, CR=0, CL=1, BL=2, AL=0)),
}
# Auto-pose tables — values of /root/reference/config/auto_pose.yaml (one entry per gait, "<gait>_pose").
AUTO_POSES: Dict[str, dict] = {
    "wave_gait": dict(
        pose_frequency=-1.0, pose_phase_length=12,
        pose_phase_starts=[1, 3, 5, 7, 9, 11], pose_phase_ends=[3, 5, 7, 9, 11, 1],
        pose_negation_phase_starts=dict(AR=1, BR=11, CR=9, CL=3, BL=5, AL=7),
        pose_negation_phase_ends=dict(AR=3, BR=1, CR=11, CL=5, BL=7, AL=9),
        negation_transition_ratio=dict(AR=0, BR=0, CR=0, CL=0, BL=0, AL=0),
        roll_amplitudes=[-0.015, 0.015, 0.015, 0.015, -0.015, -0.015],
        pitch_amplitudes=[0.020, -0.020, 0.000, 0.020, -0.020, 0.000],
        yaw_amplitudes=[0.0] * 6, x_amplitudes=[0.0] * 6, y_amplitudes=[0.0] * 6, z_amplitudes=[0.0] * 6,
        gravity_amplitudes=[0.0] * 6),
    "tripod_gait": dict(
        pose_frequency=-1.0, pose_phase_length=4,
        pose_phase_starts=[1, 3], pose_phase_ends=[3, 1],
        pose_negation_phase_starts=dict(AR=1, BR=3, CR=1, CL=3, BL=1, AL=3),
        pose_negation_phase_ends=dict(AR=3, BR=1, CR=3, CL=1, BL=3, AL=1),
        negation_transition_ratio=dict(AR=0, BR=0, CR=0, CL=0, BL=0, AL=0),
        roll_amplitudes=[-0.015, 0.015], pitch_amplitudes=[0.0, 0.0], yaw_amplitudes=[0.0, 0.0],
        x_amplitudes=[0.0, 0.0], y_amplitudes=[0.0, 0.0], z_amplitudes=[0.020, 0.020],
        gravity_amplitudes=[0.0, 0.0]),
    "ripple_gait": dict(
        pose_frequency=-1.0, pose_phase_length=6,
        pose_phase_starts=[0, 1, 2, 3, 4, 5], pose_phase_ends=[2, 3, 4, 5, 0, 1],
        pose_negation_phase_starts=dict(AR=0, BR=2, CR=4, CL=1, BL=5, AL=3),
        pose_negation_phase_ends=dict(AR=2, BR=4, CR=0, CL=3, BL=1, AL=5),
        negation_transition_ratio=dict(AR=0, BR=0, CR=0, CL=0, BL=0, AL=0),
        roll_amplitudes=[-0.015, 0.015, -0.015, 0.015, -0.015, 0.015],
        pitch_amplitudes=[-0.020, 0.020, 0.000, -0.020, 0.020, 0.000],
        yaw_amplitudes=[0.0] * 6, x_amplitudes=[0.0] * 6, y_amplitudes=[0.0] * 6, z_amplitudes=[0.0] * 6,
        gravity_amplitudes=[0.0] * 6),
    "amble_gait": dict(
        pose_frequency=-1.0, pose_phase_length=3,
        pose_phase_starts=[0, 1, 2], pose_phase_ends=[1, 2, 0],
        pose_negation_phase_starts=dict(AR=0, BR=2, CR=1, CL=0, BL=2, AL=1),
        pose_negation_phase_ends=dict(AR=1, BR=0, CR=2, CL=1, BL=0, AL=2),
        negation_transition_ratio=dict(AR=0, BR=0, CR=0, CL=0, BL=0, AL=0),
        roll_amplitudes=[0.0] * 3, pitch_amplitudes=[0.0] * 3, yaw_amplitudes=[0.0] * 3,
        x_amplitudes=[0.0] * 3, y_amplitudes=[0.0] * 3, z_amplitudes=[0.0] * 3, gravity_amplitudes=[0.0] * 3),
}


def _set_gait(cfg: ShcConfig, legs: Sequence[str], gait: dict) -> None:
    cfg.stance_phase = int(gait["stance_phase"])
    cfg.swing_phase = int(gait["swing_phase"])
    cfg.phase_offset = int(gait["phase_offset"])
    for i, name in enumerate(legs):
        cfg.offset_multiplier[i] = int(gait["offset_multiplier"][name])


def _set_auto_pose(cfg: ShcConfig, legs: Sequence[str], ap: dict) -> None:
    cfg.pose_frequency = float(ap["pose_frequency"])
    cfg.pose_phase_length = int(ap["pose_phase_length"])
    n = len(ap["pose_phase_starts"])
    if n > MAX_AUTO_POSERS:
        raise ValueError("too many auto posers")
    cfg.auto_poser_count = n
    for k in range(n):
        cfg.pose_phase_starts[k] = int(ap["pose_phase_starts"][k])
        cfg.pose_phase_ends[k] = int(ap["pose_phase_ends"][k])
        for key in ("x", "y", "z", "gravity", "roll", "pitch", "yaw"):
            getattr(cfg, key + "_amplitudes")[k] = float(ap[key + "_amplitudes"][k])
    for i, name in enumerate(legs):
        cfg.pose_negation_phase_starts[i] = int(ap["pose_negation_phase_starts"][name])
        cfg.pose_negation_phase_ends[i] = int(ap["pose_negation_phase_ends"][name])
        cfg.negation_transition_ratio[i] = float(ap["negation_transition_ratio"][name])


def _common_defaults(cfg: ShcConfig) -> None:
    """Scalar parameters of /root/reference/config/default.yaml:9-15, 75-132."""
    cfg.time_delta = 0.02
    cfg.manual_posing, cfg.auto_posing, cfg.rough_terrain_mode = 1, 0, 0
    cfg.admittance_control, cfg.inclination_posing, cfg.imu_posing = 0, 0, 0
    cfg.clamp_joint_positions, cfg.clamp_joint_velocities = 1, 1
    cfg.body_clearance = 0.100
    cfg.step_frequency = 1.000
    cfg.swing_height = 0.020
    cfg.swing_width = 0.000
    cfg.step_depth = 0.000
    cfg.stance_span_modifier = 0.000
    cfg.velocity_input_mode = 0  # throttle
    cfg.body_velocity_scaler = 1.000
    cfg.overlapping_walkspaces, cfg.force_normal_touchdown, cfg.gravity_aligned_tips = 0, 0, 0
    cfg.time_to_start = 6.000
    cfg.rotation_pid_p = cfg.rotation_pid_i = cfg.rotation_pid_d = 0.0
    for k in range(3):
        cfg.max_translation[k] = 0.025
        cfg.max_rotation[k] = 0.250
    cfg.max_translation_velocity = 0.050
    cfg.max_rotation_velocity = 0.200
    cfg.dynamic_stiffness, cfg.use_joint_effort = 1, 0
    cfg.integrator_step_time = 0.500
    cfg.virtual_mass = 10.00
    cfg.virtual_stiffness = 12.00
    cfg.virtual_damping_ratio = 0.800
    cfg.force_gain = 0.100
    cfg.load_stiffness_scaler = 5.000
    cfg.swing_stiffness_scaler = 0.100
    cfg.touchdown_threshold, cfg.liftoff_threshold = 0.9, 0.1


def hexapod_config(gait: str = "tripod_gait", time_delta: float = 0.02, **overrides) -> ShcConfig:
    """The reference's shipped robot: 6 legs x 3 DOF (config/default.yaml:25-73, 97-102)."""
    cfg = ShcConfig()
    _common_defaults(cfg)
    cfg.time_delta = time_delta
    cfg.leg_count, cfg.joint_count = 6, 3
    base_theta = dict(AR=-0.523, BR=-1.571, CR=-2.617, AL=0.523, BL=1.571, CL=2.617)
    stance = dict(AR=(0.130, -0.075), BR=(0.000, -0.150), CR=(-0.130, -0.075),
                  CL=(-0.130, 0.075), BL=(0.000, 0.150), AL=(0.130, 0.075))
    jmin, jmax = (-0.550, -1.500, -2.355), (0.550, 1.500, -0.100)
    for i, name in enumerate(HEXAPOD_LEGS):
        for j in range(3):
            cfg.joint_min[i][j], cfg.joint_max[i][j] = jmin[j], jmax[j]
            cfg.joint_max_vel[i][j], cfg.joint_offset[i][j] = 5.000, 0.000
            cfg.joint_packed[i][j], cfg.joint_unpacked[i][j] = (-1.571, 1.900, 1.200)[j], (0.000, 0.785, -1.138)[j]
        # links: base, coxa, femur, tibia  -> (d, theta, r, alpha)
        links = [(0.0, base_theta[name], 0.050, 0.000), (0.0, 0.0, 0.050, 1.571),
                 (0.0, 0.0, 0.050, 0.000), (0.0, -0.100, 0.100, 0.000)]
        for k, (d, th, r, al) in enumerate(links):
            cfg.link_d[i][k], cfg.link_theta[i][k], cfg.link_r[i][k], cfg.link_alpha[i][k] = d, th, r, al
        cfg.stance_x[i], cfg.stance_y[i] = stance[name]
    _set_gait(cfg, HEXAPOD_LEGS, GAITS[gait])
    _set_auto_pose(cfg, HEXAPOD_LEGS, AUTO_POSES[gait])
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


OCTOPOD_LEGS = ["AR", "BR", "CR", "DR", "DL", "CL", "BL", "AL"]


def octopod_config(gait: str = "tripod_gait", time_delta: float = 0.02, **overrides) -> ShcConfig:
    """Synthetic 8-leg x 5-DOF robot for BASELINE.json configs[3] (authored here).

    Each leg is the hexapod leg with two extra joints: a second yaw joint after the coxa and a second pitch
    joint before the tip (coxa, coxa2, femur, tibia, tarsus).  Legs are spread clockwise from front-right.
    Gaits map the 8 legs onto the reference's offset patterns by alternating groups.
    """
    cfg = ShcConfig()
    _common_defaults(cfg)
    cfg.time_delta = time_delta
    cfg.leg_count, cfg.joint_count = 8, 5
    n = 8
    for i, name in enumerate(OCTOPOD_LEGS):
        right = i < 4
        k = i if right else 7 - i  # 0..3 front -> back on each side
        ang = math.radians(30.0 + 40.0 * k)  # angle from +x towards the side
        theta0 = -ang if right else ang
        #            d      theta    r      alpha
        links = [(0.0, round(theta0, 3), 0.060, 0.000),   # base -> coxa joint
                 (0.0, 0.000, 0.030, 0.000),              # coxa (yaw)
                 (0.0, 0.000, 0.040, 1.571),              # coxa2 (yaw) then twist to pitch axes
                 (0.0, 0.000, 0.060, 0.000),              # femur (pitch)
                 (0.0, 0.000, 0.060, 0.000),              # tibia (pitch)
                 (0.0, -0.100, 0.070, 0.000)]             # tarsus (pitch) -> tip
        for kk, (d, th, r, al) in enumerate(links):
            cfg.link_d[i][kk], cfg.link_theta[i][kk], cfg.link_r[i][kk], cfg.link_alpha[i][kk] = d, th, r, al
        jmin = (-0.500, -0.500, -1.500, -2.000, -1.600)
        jmax = (0.500, 0.500, 1.500, -0.100, 0.600)
        for j in range(5):
            cfg.joint_min[i][j], cfg.joint_max[i][j] = jmin[j], jmax[j]
            cfg.joint_max_vel[i][j], cfg.joint_offset[i][j] = 5.000, 0.000
            cfg.joint_packed[i][j] = (-1.400 if right else 1.400, 0.000, 1.450, -1.900, 0.500)[j]
            cfg.joint_unpacked[i][j] = (0.000, 0.000, 0.700, -1.200, -0.300)[j]
        rad = 0.200
        cfg.stance_x[i] = round(rad * math.cos(theta0), 3)
        cfg.stance_y[i] = round(rad * math.sin(theta0), 3)
    # gaits: alternate legs around the body (tripod -> "tetrapod" alternating groups)
    g = GAITS[gait]
    base_mult = [g["offset_multiplier"][nm] for nm in HEXAPOD_LEGS]
    cfg.stance_phase, cfg.swing_phase, cfg.phase_offset = g["stance_phase"], g["swing_phase"], g["phase_offset"]
    if gait == "tripod_gait":
        mult = [i % 2 for i in range(n)]
    elif gait == "wave_gait":
        mult = [2, 3, 4, 5, 0, 1, 2, 3]
    else:
        mult = [base_mult[i % 6] for i in range(n)]
    for i in range(n):
        cfg.offset_multiplier[i] = mult[i]
    ap = AUTO_POSES[gait]
    legs6 = HEXAPOD_LEGS
    ap8 = dict(ap)
    ap8["pose_negation_phase_starts"] = {nm: ap["pose_negation_phase_starts"][legs6[i % 6]] for i, nm in enumerate(OCTOPOD_LEGS)}
    ap8["pose_negation_phase_ends"] = {nm: ap["pose_negation_phase_ends"][legs6[i % 6]] for i, nm in enumerate(OCTOPOD_LEGS)}
    ap8["negation_transition_ratio"] = {nm: 0.0 for nm in OCTOPOD_LEGS}
    _set_auto_pose(cfg, OCTOPOD_LEGS, ap8)
    # BASELINE.json configs[3]: admittance + IMU + inclination posing, non-zero PID gains
    cfg.admittance_control, cfg.imu_posing, cfg.inclination_posing = 1, 1, 1
    cfg.rotation_pid_p, cfg.rotation_pid_i, cfg.rotation_pid_d = 0.20, 0.05, 0.01
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def load_reference_yaml(default_yaml: str, gait_yaml: str, auto_pose_yaml: Optional[str] = None,
                        gait: Optional[str] = None) -> ShcConfig:
    """Build an ShcConfig from the reference's own YAML schema (config/default.yaml, gait.yaml, auto_pose.yaml).

    Mirrors StateController::initParameters / initGaitParameters / initAutoPoseParameters
    (/root/reference/src/state_controller.cpp:1771-2001).  Adjustable parameters take their `default` value.
    """
    import yaml

    with open(default_yaml) as f:
        p = yaml.safe_load(f)["syropod"]["parameters"]
    with open(gait_yaml) as f:
        gaits = yaml.safe_load(f)["syropod"]["gait_parameters"]
    cfg = ShcConfig()
    legs: List[str] = list(p["leg_id"])
    joints: List[str] = list(p["joint_id"])
    links: List[str] = list(p["link_id"])
    dofs = {p["leg_DOF"][n] for n in legs}
    if len(dofs) != 1:
        raise ValueError("the batched engine needs the same DOF on every leg")
    D = dofs.pop()
    if len(legs) > MAX_LEGS or D > MAX_DOF:
        raise ValueError("robot exceeds 8 legs x 5 DOF")
    adj = lambda v: float(v["default"]) if isinstance(v, dict) else float(v)  # noqa: E731
    cfg.time_delta = float(p["time_delta"])
    for key in ("manual_posing", "auto_posing", "rough_terrain_mode", "admittance_control", "inclination_posing",
                "imu_posing", "clamp_joint_positions", "clamp_joint_velocities", "overlapping_walkspaces",
                "force_normal_touchdown", "gravity_aligned_tips", "dynamic_stiffness", "use_joint_effort"):
        setattr(cfg, key, int(bool(p[key])))
    cfg.leg_count, cfg.joint_count = len(legs), D
    for i, leg in enumerate(legs):
        for j in range(D):
            jp = p[f"{leg}_{joints[j]}_joint_parameters"]
            cfg.joint_min[i][j], cfg.joint_max[i][j] = float(jp["min"]), float(jp["max"])
            cfg.joint_max_vel[i][j], cfg.joint_offset[i][j] = float(jp["max_vel"]), float(jp["offset"])
            cfg.joint_packed[i][j], cfg.joint_unpacked[i][j] = float(jp.get("packed", 0.0)), float(jp.get("unpacked", 0.0))
        for k in range(D + 1):
            lp = p[f"{leg}_{links[k]}_link_parameters"]
            cfg.link_d[i][k], cfg.link_theta[i][k] = float(lp["d"]), float(lp["theta"])
            cfg.link_r[i][k], cfg.link_alpha[i][k] = float(lp["r"]), float(lp["alpha"])
        sp = p[f"{leg}_stance_position"]
        cfg.stance_x[i], cfg.stance_y[i] = float(sp["x"]), float(sp["y"])
    cfg.body_clearance = float(p["body_clearance"])
    cfg.touchdown_threshold, cfg.liftoff_threshold = float(p.get("touchdown_threshold", 0.9)), float(p.get("liftoff_threshold", 0.1))
    for key in ("step_frequency", "swing_height", "swing_width", "step_depth", "stance_span_modifier",
                "virtual_mass", "virtual_stiffness", "virtual_damping_ratio", "force_gain"):
        setattr(cfg, key, adj(p[key]))
    cfg.velocity_input_mode = 0 if p["velocity_input_mode"] == "throttle" else 1
    cfg.body_velocity_scaler = float(p["body_velocity_scaler"])
    cfg.time_to_start = float(p["time_to_start"])
    cfg.rotation_pid_p = float(p["rotation_pid_gains"]["p"])
    cfg.rotation_pid_i = float(p["rotation_pid_gains"]["i"])
    cfg.rotation_pid_d = float(p["rotation_pid_gains"]["d"])
    for k, ax in enumerate(("x", "y", "z")):
        cfg.max_translation[k] = float(p["max_translation"][ax])
    for k, ax in enumerate(("roll", "pitch", "yaw")):
        cfg.max_rotation[k] = float(p["max_rotation"][ax])
    cfg.max_translation_velocity = float(p["max_translation_velocity"])
    cfg.max_rotation_velocity = float(p["max_rotation_velocity"])
    cfg.integrator_step_time = float(p["integrator_step_time"])
    cfg.load_stiffness_scaler = float(p["load_stiffness_scaler"])
    cfg.swing_stiffness_scaler = float(p["swing_stiffness_scaler"])
    gait_name = gait or p["gait_type"]
    _set_gait(cfg, legs, gaits[gait_name])
    if auto_pose_yaml is not None:
        with open(auto_pose_yaml) as f:
            aps = yaml.safe_load(f)["syropod"]["auto_pose_parameters"]
        key = gait_name + "_pose" if p.get("auto_pose_type", "auto") == "auto" else p["auto_pose_type"]
        _set_auto_pose(cfg, legs, aps[key])
    return cfg


def config_to_dict(cfg) -> dict:
    """Plain nested-list/dict view of a ctypes struct (for comparisons and JSON fixtures)."""
    def conv(v):
        if isinstance(v, C.Array):
            return [conv(x) for x in v]
        if isinstance(v, C.Structure):
            return config_to_dict(v)
        return v
    return {name: conv(getattr(cfg, name)) for name, _ in cfg._fields_}
