// oracle/shim/syropod_highlevel_controller/msgs_generated.h — TEST INFRASTRUCTURE ONLY: what catkin would generate from the
// reference's msg/LegState.msg, msg/TipState.msg, msg/TargetTipPose.msg and config/Dynamic.cfg (field names and types
// only; generated code is not part of the reference tree).
#ifndef SHC_SHIM_SHC_MSGS_GENERATED_H
#define SHC_SHIM_SHC_MSGS_GENERATED_H
#include "msg_common.h"
namespace syropod_highlevel_controller {
struct LegState {
  std_msgs::Header header;
  std::string name;
  geometry_msgs::PoseStamped walker_tip_pose, target_tip_pose, poser_tip_pose, model_tip_pose, actual_tip_pose;
  geometry_msgs::TwistStamped model_tip_velocity;
  std::vector<double> joint_positions, joint_velocities, joint_efforts;
  double stance_progress = 0, swing_progress = 0;
  double time_to_swing_end = 0;
  geometry_msgs::Pose pose_delta;
  geometry_msgs::Pose auto_pose;
  geometry_msgs::Vector3 tip_force, admittance_delta;
  double virtual_stiffness = 0;
};
struct TargetTipPose {
  std_msgs::Header header;
  std::vector<std::string> name;
  std::vector<geometry_msgs::PoseStamped> target, stance;
  std::vector<double> swing_clearance;
};
struct TipState {
  std_msgs::Header header;
  std::vector<std::string> name;
  std::vector<geometry_msgs::Wrench> wrench;
  std::vector<geometry_msgs::Vector3> step_plane;
};
struct DynamicConfig {
  double step_frequency = 0, swing_height = 0, swing_width = 0, step_depth = 0, stance_span_modifier = 0, virtual_mass = 0,
         virtual_stiffness = 0, virtual_damping_ratio = 0, force_gain = 0;
};
}  // namespace syropod_highlevel_controller
#endif
