/*
 * shc_config.h — plain-C parameter block for the batched SHC hot path.
 *
 * This is the data contract both the CUDA engine (include/shc_b200.h) and the CPU
 * parity oracle (oracle/) are configured with, so that "same inputs" is literal.
 * It carries exactly the fields of the reference's `Parameters` struct that the
 * per-control-cycle path reads (reference: include/syropod_highlevel_controller/
 * parameters_and_states.h:271-381, values in config/default.yaml, config/gait.yaml,
 * config/auto_pose.yaml).  No code lives here — only POD layout.
 */
#ifndef SHC_CONFIG_H
#define SHC_CONFIG_H

#ifdef __cplusplus
extern "C" {
#endif

#define SHC_MAX_LEGS 8        /* parameters_and_states.h:298 */
#define SHC_MAX_DOF 5         /* parameters_and_states.h:299 (joint_parameters[8][6]; README: <= 5 DOF) */
#define SHC_MAX_AUTO_POSERS 8 /* one AutoPoser per pose_phase_starts entry (pose_controller.cpp:83) */
#define SHC_N_BEARINGS 9      /* LimitMap keys 0,45,...,360 (model.h:22 BEARING_STEP) */

enum { SHC_VELOCITY_THROTTLE = 0, SHC_VELOCITY_REAL = 1 }; /* velocity_input_mode (default.yaml:90) */
/* a tip range-sensor reading at or above this value stands for the reference's UNASSIGNED_VALUE ("no reading") */
#define SHC_RANGE_UNASSIGNED 1.0e9f

typedef struct shc_config {
  /* ---- control flags (default.yaml:9-15) ---- */
  double time_delta;
  int manual_posing;
  int auto_posing;
  int rough_terrain_mode; /* layered workspace, default-tip updates, touchdown detection (tip forces / range sensors),
                           * external targets and defaults, target shifting */
  int admittance_control;
  int inclination_posing;
  int imu_posing;

  /* ---- model (default.yaml:25-73) ---- */
  int leg_count;   /* L */
  int joint_count; /* D, uniform over legs in this engine */
  double joint_min[SHC_MAX_LEGS][SHC_MAX_DOF];
  double joint_max[SHC_MAX_LEGS][SHC_MAX_DOF];
  double joint_max_vel[SHC_MAX_LEGS][SHC_MAX_DOF];
  double joint_offset[SHC_MAX_LEGS][SHC_MAX_DOF]; /* added at output only (state_controller.cpp:795) */
  /* link 0 = "base" link (constant transform), links 1..D actuated by joints 1..D (model.cpp:221-239) */
  double link_d[SHC_MAX_LEGS][SHC_MAX_DOF + 1];
  double link_theta[SHC_MAX_LEGS][SHC_MAX_DOF + 1];
  double link_r[SHC_MAX_LEGS][SHC_MAX_DOF + 1];
  double link_alpha[SHC_MAX_LEGS][SHC_MAX_DOF + 1];
  int clamp_joint_positions;
  int clamp_joint_velocities;

  /* ---- walker (default.yaml:79-108) ---- */
  double body_clearance;
  double step_frequency; /* current_value of the adjustable parameter */
  double swing_height;
  double swing_width;
  double step_depth;
  double stance_span_modifier;
  int velocity_input_mode;
  double body_velocity_scaler;
  double stance_x[SHC_MAX_LEGS];
  double stance_y[SHC_MAX_LEGS];
  int overlapping_walkspaces;
  int force_normal_touchdown;
  int gravity_aligned_tips;

  /* ---- gait (gait.yaml) ---- */
  int stance_phase;
  int swing_phase;
  int phase_offset;
  int offset_multiplier[SHC_MAX_LEGS];

  /* ---- poser (default.yaml:112-120) ---- */
  double time_to_start;
  double rotation_pid_p, rotation_pid_i, rotation_pid_d;
  double max_translation[3]; /* x y z */
  double max_rotation[3];    /* roll pitch yaw */
  double max_translation_velocity;
  double max_rotation_velocity;

  /* ---- auto pose (auto_pose.yaml) ---- */
  double pose_frequency; /* -1.0 = sync with step cycle */
  int pose_phase_length;
  int auto_poser_count;
  int pose_phase_starts[SHC_MAX_AUTO_POSERS];
  int pose_phase_ends[SHC_MAX_AUTO_POSERS];
  int pose_negation_phase_starts[SHC_MAX_LEGS];
  int pose_negation_phase_ends[SHC_MAX_LEGS];
  double negation_transition_ratio[SHC_MAX_LEGS];
  double x_amplitudes[SHC_MAX_AUTO_POSERS];
  double y_amplitudes[SHC_MAX_AUTO_POSERS];
  double z_amplitudes[SHC_MAX_AUTO_POSERS];
  double gravity_amplitudes[SHC_MAX_AUTO_POSERS];
  double roll_amplitudes[SHC_MAX_AUTO_POSERS];
  double pitch_amplitudes[SHC_MAX_AUTO_POSERS];
  double yaw_amplitudes[SHC_MAX_AUTO_POSERS];

  /* ---- admittance (default.yaml:124-132) ---- */
  int dynamic_stiffness;
  int use_joint_effort;
  double integrator_step_time;
  double virtual_mass;
  double virtual_stiffness;
  double virtual_damping_ratio;
  double force_gain;
  double load_stiffness_scaler;
  double swing_stiffness_scaler;

  /* ---- pack / unpack (default.yaml:31-48 "packed" / "unpacked": Joint::packed_positions_ with one pack step,
   *      Joint::unpacked_position_; model.cpp:1019-1036) ---- */
  double joint_packed[SHC_MAX_LEGS][SHC_MAX_DOF];
  double joint_unpacked[SHC_MAX_LEGS][SHC_MAX_DOF];

  /* ---- touchdown detection (default.yaml:107-108; Leg::touchdownDetection, model.cpp:712) ---- */
  double touchdown_threshold; /* N: |measured tip force| above it defines the step plane at the tip */
  double liftoff_threshold;   /* N: below it the step plane is forgotten */
} shc_config;

/* Constants produced by the reference's start-up path (state_controller.cpp:263-281:
 * directStartup -> updateDefaultConfiguration -> generateWorkspaces -> generateWalkspace ->
 * generateLimits) that the per-cycle path then treats as read-only. */
typedef struct shc_startup {
  double default_joint[SHC_MAX_LEGS][SHC_MAX_DOF];     /* joint angles of the default stance */
  double workspace[SHC_MAX_LEGS][SHC_N_BEARINGS];      /* simple workspace radii (model.cpp:309) */
  double walkspace[SHC_N_BEARINGS];                    /* walk_controller.cpp:57 */
  double max_linear_speed[SHC_N_BEARINGS];             /* walk_controller.cpp:231 */
  double max_angular_speed[SHC_N_BEARINGS];
  double max_linear_acceleration[SHC_N_BEARINGS];
  double max_angular_acceleration[SHC_N_BEARINGS];
  /* StepCycle (walk_controller.h:23-33) */
  double step_frequency;
  int period, swing_period, stance_period, stance_end, swing_start, swing_end, stance_start;
  int phase_offsets[SHC_MAX_LEGS];
  /* auto-pose cycle (pose_controller.cpp:44-63) */
  int pose_phase_length, pose_normaliser;
  int auto_pose_reference_leg;
  /* number of StateController::loop() calls the direct start-up takes until READY (state_controller.cpp:254-281,
   * pose_controller.cpp:463): an auto poser with its own cycle (pose_frequency != -1) keeps cycling through them */
  int startup_loops;
} shc_startup;

#ifdef __cplusplus
}
#endif
#endif /* SHC_CONFIG_H */
