/*
 * shc_state.h — host exchange format for the per-robot state that carries across control cycles.
 *
 * The engine keeps state on the device as struct-of-arrays planes (DESIGN.md "Data layout in HBM");
 * shc_get_state()/shc_set_state() (include/shc_b200.h) convert to and from this array-of-structs record,
 * in double, one record per robot.  The parity oracle exports the same record, so a test can start both
 * from one snapshot and compare every field after a step.  Poses are 7 doubles: px py pz qw qx qy qz.
 *
 * Each field names the reference member it mirrors (file:line under /root/reference).
 */
#ifndef SHC_STATE_H
#define SHC_STATE_H

#include "shc_config.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct shc_leg_state {
  /* Joint (model.h:635-636) */
  double joint_position[SHC_MAX_DOF]; /* desired_position_ */
  double joint_velocity[SHC_MAX_DOF]; /* desired_velocity_ */
  /* LegStepper (walk_controller.h:493-531) */
  double tip_position[3];           /* current_tip_pose_.position_ (walk-plane frame) */
  double tip_velocity[3];           /* current_tip_velocity_ */
  double swing_origin_position[3];  /* swing_origin_tip_position_ */
  double swing_origin_velocity[3];  /* swing_origin_tip_velocity_ */
  double stance_origin_position[3]; /* stance_origin_tip_position_ */
  double default_tip_position[3];   /* default_tip_pose_.position_ */
  double target_tip_position[3];    /* target_tip_pose_.position_ */
  double stride_vector[3];          /* stride_vector_ */
  double walk_plane[3];             /* walk_plane_ (saved copy) */
  double walk_plane_normal[3];      /* walk_plane_normal_ */
  double swing_progress;            /* swing_progress_ */
  double stance_progress;           /* stance_progress_ */
  int phase;                        /* phase_ */
  int step_state;                   /* step_state_: 0 SWING 1 STANCE 2 FORCE_STANCE 3 FORCE_STOP */
  int at_correct_phase;             /* at_correct_phase_ */
  int completed_first_step;         /* completed_first_step_ */
  /* Leg admittance / force (model.h:514-531) */
  double admittance_state[2];       /* admittance_state_ */
  double admittance_delta[3];       /* admittance_delta_ */
  double tip_force_calculated[3];   /* tip_force_calculated_ */
  double virtual_stiffness;         /* virtual_stiffness_ (model.h:516): dynamic stiffness of AdmittanceController::
                                     * updateStiffness (admittance_controller.cpp:96); only read by LegState.msg (trap 3);
                                     * never initialised by the reference before the first update - 0 here */
  /* LegPoser auto-pose negation latch (pose_controller.h:575) */
  int negate_auto_pose;
  int pad0;
  /* LegStepper tip rotations, w x y z (walk_controller.h:513-516; walk_controller.cpp:1193 updateTipRotation): only live with
   * gravity_aligned_tips on legs of more than three joints; all four components zero = UNDEFINED_ROTATION */
  double tip_rotation[4];           /* current_tip_pose_.rotation_ */
  double origin_tip_rotation[4];    /* origin_tip_pose_.rotation_ */
  double target_tip_rotation[4];    /* target_tip_pose_.rotation_ (a constant of the configuration without rough-terrain
                                     * targets: written by shc_get_state, ignored by shc_set_state) */
  /* rough-terrain mode */
  double step_plane_position[3];    /* Leg::step_plane_pose_.position_ (model.h:529): the tip pose at touchdown, base_link frame */
  int step_plane_defined;           /* step_plane_pose_ != Pose::Undefined() */
  int touchdown_detection;          /* LegStepper::touchdown_detection_ (walk_controller.h:495): tip state inputs have arrived */
  /* externally requested swing target and default tip pose (LegStepper::external_target_ / external_default_,
   * walk_controller.h:36-44, set by targetTipPoseCallback state_controller.cpp:1700-1760; `transform` = the robot's movement
   * since the request, which the reference refreshes from the tf tree every loop, :703-750): rough-terrain mode only.
   * Poses are px py pz qw qx qy qz. */
  double external_target_pose[7];
  double external_target_transform[7];
  double external_target_clearance;  /* ExternalTarget::swing_clearance_ */
  int external_target_defined;       /* cleared by the stepper at the start of the next stance */
  int external_target_odom_frame;    /* frame_id_ == "odom_ideal": the target is led by the odometry until the swing ends */
  double external_default_pose[7];
  double external_default_transform[7];
  int external_default_defined;
  int pad1;
  /* outputs of the last cycle (recomputed every cycle; not algorithmic state) */
  double model_tip_position[3];     /* Leg::current_tip_pose_.position_ after applyFK (base_link frame) */
  double desired_tip_position[3];   /* Leg::desired_tip_pose_.position_ */
  double ik_result;                 /* return value of Leg::applyIK (model.cpp:861): 0.0 = failed this cycle */
} shc_leg_state;

typedef struct shc_robot_state {
  /* WalkController (walk_controller.h:245-268) */
  double desired_linear_velocity[2];
  double desired_angular_velocity;
  int walk_state;                   /* 0 STARTING 1 MOVING 2 STOPPING 3 STOPPED */
  int legs_at_correct_phase;
  int legs_completed_first_step;
  int return_to_default_attempted;
  int pose_state;                   /* walker's copy of the auto posing state */
  int pad0;
  double walk_plane[3];
  double walk_plane_normal[3];
  double odometry_ideal[7];
  /* PoseController (pose_controller.h:278-317) */
  double walk_plane_pose[7];
  double origin_walk_plane_pose[7];
  double manual_pose[7];
  double imu_pose[7];
  double inclination_pose[7];
  double auto_pose[7];
  double rotation_absement_error[3];
  double rotation_position_error[3];
  double rotation_velocity_error[3];
  /* tip-align posing (pose_controller.cpp:1024 updateTipAlignPose): gravity_aligned_tips on legs of at most three joints */
  double tip_align_pose[7];         /* tip_align_pose_ */
  double origin_tip_align_pose[7];  /* origin_tip_align_pose_ */
  int auto_posing_state;            /* 0 POSING 1 STOP_POSING 2 POSING_COMPLETE */
  int pose_phase;
  /* AutoPoser latches (pose_controller.h:408-410): bit0 start_check, bit1 end_check.first, bit2 end_check.second,
   * bit3 allow_posing */
  int auto_poser_flags[SHC_MAX_AUTO_POSERS];
  /* output of the pose stage (recomputed every cycle) */
  double current_pose[7];           /* Model::current_pose_ (walk-plane -> base_link) */
  int status_flags;                 /* engine status word, see SHC_FLAG_* in shc_b200.h; 0 in oracle records */
  int pad1;
  shc_leg_state legs[SHC_MAX_LEGS];
} shc_robot_state;

#ifdef __cplusplus
}
#endif
#endif /* SHC_STATE_H */
