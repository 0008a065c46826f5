/* shc_msgs.h — the reference's OUTPUT WIRE FORMATS as plain C records: what
 * StateController's publishers assemble per cycle with host-side loops over legs and joints —
 *   publishDesiredJointState  state_controller.cpp:777-805   (Leg::generateDesiredJointStateMsg, model.cpp:605-617)
 *   publishLegState           state_controller.cpp:809-893   (msg/LegState.msg)
 *   publishVelocity / publishPose / publishRotationPoseError  :897-961
 *   publishFrameTransforms    state_controller.cpp:963-1047  (odom_ideal -> base_link, base_link -> walk_plane,
 *                                                             base_link -> every joint frame and tip frame)
 * — are packed by ONE kernel (shc_pack_messages, shc_b200.h) straight from the engine's state planes into these records,
 * for any range of robots, into device memory or page-locked host memory (then there is no copy at all).
 * Every pose is [x y z  qw qx qy qz] (orientation w first, as everywhere in this interface; geometry_msgs/Pose carries
 * x y z w).  Strings (frame ids, joint names) and time stamps are the caller's: they are constants per joint. */
#ifndef SHC_MSGS_H
#define SHC_MSGS_H
#include "shc_config.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sensor_msgs/JointState of publishDesiredJointState (combined interface), joints in leg-major order (leg 0 joint 1..D,
 * leg 1 ...: the order Leg::generateDesiredJointStateMsg appends them), plus the individual interface's command. */
typedef struct shc_joint_state_msg {
  double position[SHC_MAX_LEGS * SHC_MAX_DOF]; /* Joint::desired_position_ */
  double velocity[SHC_MAX_LEGS * SHC_MAX_DOF]; /* Joint::desired_velocity_ */
  double effort[SHC_MAX_LEGS * SHC_MAX_DOF];   /* Joint::desired_effort_ (the controller never commands one: 0) */
  double position_command[SHC_MAX_LEGS * SHC_MAX_DOF]; /* desired_position_ + offset_ (std_msgs/Float64 per joint, :795) */
} shc_joint_state_msg;

/* msg/LegState.msg, numeric fields in message order, then the leg's frames of publishFrameTransforms. */
typedef struct shc_leg_state_msg {
  double walker_tip_pose[7];     /* LegStepper::current_tip_pose_ (walk_plane frame; rotation undefined = 0 0 0 0) */
  double target_tip_pose[7];     /* LegStepper::target_tip_pose_ */
  double poser_tip_pose[7];      /* LegPoser::current_tip_pose_ (base_link frame): PoseController::updateStance */
  double model_tip_pose[7];      /* Leg::current_tip_pose_: forward kinematics of the desired joint positions */
  double actual_tip_pose[7];     /* Leg::applyFK(false, true): forward kinematics of the MEASURED joint positions */
  double model_tip_velocity[3];  /* Leg::current_tip_velocity_ as published: zero (state_controller.cpp:842 resets it before :846 reads it) */
  double joint_positions[SHC_MAX_DOF], joint_velocities[SHC_MAX_DOF], joint_efforts[SHC_MAX_DOF];
  double stance_progress, swing_progress;
  double time_to_swing_end;      /* :866-877 */
  double pose_delta[7];          /* WalkController::calculateOdometry(time_to_swing_end) */
  double auto_pose[7];           /* LegPoser::auto_pose_ */
  double tip_force[3];           /* tip_force_calculated_ * force_gain */
  double admittance_delta[3];
  double virtual_stiffness;
  /* publishFrameTransforms :1017-1047 */
  double joint_transform[SHC_MAX_DOF][7]; /* base_link -> joint i: Joint::getPoseRobotFrame() turned by the joint angle */
  double tip_transform[7];                /* base_link -> tip: Tip::getPoseRobotFrame() */
} shc_leg_state_msg;

typedef struct shc_body_msg {
  double velocity[6];                /* publishVelocity (geometry_msgs/Twist): linear x y 0, angular 0 0 z */
  double pose[6];                    /* publishPose: Model::current_pose_ position, then its Euler angles (roll pitch yaw) */
  double rotation_pose_error[9];     /* publishRotationPoseError: absement (3), position (3), velocity (3) */
  double odom_ideal_to_base_link[7]; /* publishFrameTransforms :965-990 */
  double base_link_to_walk_plane[7]; /* :993-1004: inverse of Model::current_pose_ */
} shc_body_msg;

#ifdef __cplusplus
}
#endif
#endif /* SHC_MSGS_H */
