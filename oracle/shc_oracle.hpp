// shc_oracle.hpp — TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).  See oracle_math.hpp for the rules:
// nothing in the product path may include, link or call this.
//
// PARITY PINNED TO THE REFERENCE'S OWN BUILD.  The reference ships no tests or golden vectors and its
// catkin build needs ROS, tf2, Eigen and Boost, none of which are in the image — but its control sources compile
// unmodified against self-written stand-ins for those headers (oracle/shim/, oracle/Makefile.ref -> oracle/_ref/
// libshc_ref.so, driven by oracle/ref_harness.cpp).  tests/test_reference_pin.py runs that library and this restatement
// side by side: equal start-up constants, every state field of every cycle equal (bit for bit on 3-DOF legs, <= 1e-13
// rad on 5-DOF legs) over gaits, rates, posing modes, admittance, tip orientation, rough terrain, sequences, publishers;
// tests/golden/*.npz are outputs of the reference.  What stays restated-from-publication is third-party arithmetic
// (Eigen kernels, Boost.Odeint RK4): the stand-ins implement the same published formulas.
// This file is a line-by-line restatement, in IEEE double, of the reference arithmetic for one robot:
//   Model/Leg/Joint/Link/Tip   src/model.cpp, include/.../model.h
//   WalkController/LegStepper  src/walk_controller.cpp
//   PoseController/AutoPoser/LegPoser  src/pose_controller.cpp
//   AdmittanceController       src/admittance_controller.cpp
//   StateController::init/loop/transitionRobotState/runningState   src/state_controller.cpp:127-447
// Containers (std::map<int,shared_ptr<..>>) are replaced by fixed arrays iterated in the same (id) order.
#pragma once
#include <map>
#include <vector>

#include "../include/shc_config.h"
#include "oracle_math.hpp"

namespace shc_oracle {
using namespace om;

constexpr double IK_TOLERANCE = 0.005;           // model.h:17
constexpr double DLS_COEFFICIENT = 0.02;         // model.h:19
constexpr double JOINT_LIMIT_COST_WEIGHT = 0.1;  // model.h:20
constexpr int BEARING_STEP = 45;                 // model.h:22
constexpr double MAX_POSITION_DELTA = 0.002;     // model.h:23
constexpr double MAX_WORKSPACE_RADIUS = 1.0;     // model.h:24
constexpr int WORKSPACE_LAYERS = 10;             // model.h:25
constexpr double JOINT_TOLERANCE = 0.01;         // pose_controller.h:18
constexpr double TIP_TOLERANCE = 0.01;           // pose_controller.h:19
constexpr double STABILITY_THRESHOLD = 100;      // pose_controller.h:23
constexpr double IMU_POSING_DEADBAND = 0.0;      // pose_controller.h:25
constexpr double ADMITTANCE_DEADBAND = 0.0;      // admittance_controller.h:18
constexpr double GRAVITY_ACCELERATION = -9.81;   // standard_includes.h:59
constexpr int PROGRESS_COMPLETE = 100;           // standard_includes.h:53
constexpr double SAFETY_FACTOR = 0.15;           // pose_controller.h:20
constexpr double HORIZONTAL_TRANSITION_TIME = 1.0;  // pose_controller.h:21
constexpr double VERTICAL_TRANSITION_TIME = 3.0;    // pose_controller.h:22
constexpr int TRANSITION_STEP_THRESHOLD = 20;    // pose_controller.h:24
constexpr double HALF_BODY_DEPTH = 0.05;         // model.h:18
enum SequenceSelection { START_UP = 0, SHUT_DOWN = 1 };  // parameters_and_states.h

enum RobotState { PACKED, READY, RUNNING, ROBOT_STATE_COUNT, UNKNOWN = -1, OFF = -2 };  // parameters_and_states.h:27
enum LegState { WALKING, MANUAL, LEG_STATE_COUNT, WALKING_TO_MANUAL = -1, MANUAL_TO_WALKING = -2 };
enum WalkState { STARTING, MOVING, STOPPING, STOPPED };
enum StepState { SWING, STANCE, FORCE_STANCE, FORCE_STOP };
enum PosingState { POSING, STOP_POSING, POSING_COMPLETE };
enum PoseResetMode { NO_RESET, Z_AND_YAW_RESET, X_AND_Y_RESET, PITCH_AND_ROLL_RESET, ALL_RESET, IMMEDIATE_ALL_RESET };

typedef std::map<int, double> LimitMap;   // walk_controller.h:18 / model.h:190 Workplane
typedef std::map<double, LimitMap> Workspace;

struct StepCycle {  // walk_controller.h:23
  double frequency_ = 0;
  int period_ = 0, swing_period_ = 0, stance_period_ = 0, stance_end_ = 0, swing_start_ = 0, swing_end_ = 0,
      stance_start_ = 0;
};

struct ImuData {
  Quat orientation = UndefinedRotation();
  Vec3 linear_acceleration, angular_velocity;
};

struct Robot;
struct Leg;

// ---- Joint / Link (model.h:542-652) — flattened ---------------------------------------------------------------
struct Joint {
  double min_position_ = 0, max_position_ = 0, offset_ = 0, max_angular_speed_ = 0;
  double desired_position_ = 0, desired_velocity_ = 0, desired_effort_ = 0, prev_desired_position_ = 0;
  double current_position_ = UNASSIGNED_VALUE, current_velocity_ = 0, current_effort_ = 0;
  double default_position_ = UNASSIGNED_VALUE, default_velocity_ = 0, default_effort_ = 0;
  double packed_position_ = 0, unpacked_position_ = 0;  // packed_positions_[0] / unpacked_position_ (model.cpp:1035-1056)
  Mat4 current_transform_ = Mat4::Identity();  // transform from previous joint to this joint
};
struct Link {
  double d = 0, theta = 0, r = 0, alpha = 0;
};

// ExternalTarget (walk_controller.h:36-44).  frame_id_ == "odom_ideal" is kept as a flag; time_ only serves the tf lookup
// (state_controller.cpp:717-725), whose result the harness supplies as transform_.
struct ExternalTarget {
  Pose pose_;
  double swing_clearance_ = 0.0;
  bool odom_ideal_frame_ = false;
  Pose transform_ = Pose::Identity();
  bool defined_ = false;
};

// ---- LegStepper (walk_controller.h:286-535) -------------------------------------------------------------------
struct LegStepper {
  Robot* robot = nullptr;
  Leg* leg_ = nullptr;
  bool at_correct_phase_ = false, completed_first_step_ = false;
  bool touchdown_detection_ = false;  // walk_controller.h:495: set once tip state messages arrive (state_controller.cpp:1640)
  ExternalTarget external_target_, external_default_;  // walk_controller.h:533-534
  int phase_ = 0, phase_offset_ = 0;
  double step_progress_ = 0.0, swing_progress_ = -1.0, stance_progress_ = -1.0;
  StepState step_state_ = STANCE;
  Vec3 swing_1_nodes_[5], swing_2_nodes_[5], stance_nodes_[5];
  Vec3 walk_plane_, walk_plane_normal_, stride_vector_, swing_clearance_;
  double swing_delta_t_ = 0.0, stance_delta_t_ = 0.0;
  Pose identity_tip_pose_, default_tip_pose_, current_tip_pose_, origin_tip_pose_, target_tip_pose_;
  Vec3 current_tip_velocity_, swing_origin_tip_position_, swing_origin_tip_velocity_, stance_origin_tip_position_;

  void construct(Robot* r, Leg* leg, const Pose& identity_tip_pose);
  void iteratePhase();
  void updateStepState();
  void updateStride();
  void updateDefaultTipPosition();
  Vec3 calculateStanceSpanChange();
  void updateTipPosition();
  void updateTipRotation();
  void generatePrimarySwingControlNodes();
  void generateSecondarySwingControlNodes(bool ground_contact);
  void generateStanceControlNodes(double stride_scaler);
  void forceNormalTouchdown();
};

// ---- LegPoser (pose_controller.h:425-600) ---------------------------------------------------------------------
struct LegPoser {
  Robot* robot = nullptr;
  Leg* leg_ = nullptr;
  Pose auto_pose_ = Pose::Identity();
  bool negate_auto_pose_ = false;
  int pose_negation_phase_start_ = 0, pose_negation_phase_end_ = 0;
  double negation_transition_ratio_ = 0.0;
  bool first_iteration_ = true;
  int master_iteration_count_ = 0;
  std::vector<double> desired_configuration_, origin_configuration_;
  Pose origin_tip_pose_, current_tip_pose_ = Pose::Undefined(), target_tip_pose_ = Pose::Undefined();
  bool leg_completed_step_ = false;          // pose_controller.h:596
  std::vector<Pose> transition_poses_;       // pose_controller.h:591
  int resetStepToPosition() { first_iteration_ = true; return PROGRESS_COMPLETE; }  // pose_controller.h:535

  int transitionConfiguration(double transition_time);
  int stepToPosition(const Pose& target_tip_pose, const Pose& target_pose, double lift_height, double time_to_step,
                     bool apply_delta = true);
  void updateAutoPose(int phase);
};

// ---- Leg (model.h:196-535) ------------------------------------------------------------------------------------
struct Leg {
  Robot* robot = nullptr;
  int id_number_ = 0, joint_count_ = 0, group_ = 0;
  LegState leg_state_ = WALKING;
  Joint joints[SHC_MAX_DOF + 1];  // index 1..D (index 0 = the null joint acting as chain origin)
  Link links[SHC_MAX_DOF + 1];    // index 0..D
  Mat4 tip_transform_ = Mat4::Identity();
  LegStepper stepper;
  LegPoser poser;
  Workspace workspace_;
  Vec3 admittance_delta_;
  double virtual_mass_ = 0, virtual_stiffness_ = 0, virtual_damping_ratio_ = 0;
  double admittance_state_[2] = {0, 0};
  Pose desired_tip_pose_ = Pose::Undefined(), current_tip_pose_ = Pose::Undefined();
  Vec3 desired_tip_velocity_, current_tip_velocity_;
  Vec3 tip_force_calculated_, tip_force_measured_, tip_torque_calculated_, tip_torque_measured_;
  Pose step_plane_pose_ = Pose::Undefined();
  double last_ik_result_ = 1.0;  // return value of the most recent applyIK (harness read-out only)

  void construct(Robot* r, int id);
  // Joint::getTransformFromJoint / Tip::getTransformFromJoint (model.h:594-599, 674-679)
  Mat4 jointTransformFrom(int joint_id, int target_joint_id) const;
  Mat4 tipTransformFrom(int target_joint_id) const;
  Pose jointPoseRobotFrame(int joint_id, const Pose& p = Pose::Identity()) const;   // model.h:604
  Pose jointPoseJointFrame(int joint_id, const Pose& p = Pose::Identity()) const;   // model.h:613
  Pose tipPoseRobotFrame(const Pose& p = Pose::Identity()) const;                   // model.h:684
  void init(bool use_default_joint_positions);                                      // model.cpp:286
  Workspace generateWorkspace();                                                    // model.cpp:309
  LimitMap getWorkplane(double height);                                             // model.cpp:514
  void updateDefaultConfiguration();                                                // model.cpp:593
  void setDesiredTipPose(const Pose& tip_pose = Pose::Undefined(), bool apply_delta = true);  // model.cpp:653
  void calculateTipForce();                                                         // model.cpp:667
  void touchdownDetection();                                                        // model.cpp:712
  void solveIK(const double delta[6], bool solve_rotation, double* out);            // model.cpp:726
  double updateJointPositions(const double* delta, bool simulation);                // model.cpp:799
  double applyIK(bool simulation = false);                                          // model.cpp:861
  Pose applyFK(bool set_current = true, bool use_actual = false);                   // model.cpp:945
  void setAdmittanceDelta(const Vec3& delta) {                                      // model.h:365
    admittance_delta_ = getProjection(delta, current_tip_pose_.rotation_.transformVector(UnitX()));
  }
};

// ---- AutoPoser (pose_controller.h:333-420) --------------------------------------------------------------------
struct AutoPoser {
  Robot* robot = nullptr;
  int id_number_ = 0, start_phase_ = 0, end_phase_ = 0;
  bool start_check_ = false;
  bool end_check_first = false, end_check_second = false;
  bool allow_posing_ = false;
  double x_amplitude_ = 0, y_amplitude_ = 0, z_amplitude_ = 0, gravity_amplitude_ = 0, roll_amplitude_ = 0,
         pitch_amplitude_ = 0, yaw_amplitude_ = 0;
  Pose updatePose(int phase);
};

// ---- one robot: Model + WalkController + PoseController + AdmittanceController + StateController harness -------
struct Robot {
  shc_config params_;

  // Model (model.h:58-182)
  int leg_count_ = 0;
  double time_delta_ = 0;
  Pose current_pose_ = Pose::Identity(), default_pose_model_ = Pose::Identity();
  ImuData imu_data_;
  Leg legs[SHC_MAX_LEGS];

  // WalkController (walk_controller.h:240-273)
  WalkState walk_state_ = STOPPED;
  PosingState pose_state_ = POSING_COMPLETE;
  StepCycle step_;
  LimitMap walkspace_;
  Vec3 walk_plane_, walk_plane_normal_;
  double desired_linear_velocity_[2] = {0, 0};
  double desired_angular_velocity_ = 0;
  Pose odometry_ideal_ = Pose::Identity();
  LimitMap max_linear_speed_, max_angular_speed_, max_linear_acceleration_, max_angular_acceleration_;
  int legs_at_correct_phase_ = 0, legs_completed_first_step_ = 0;
  bool return_to_default_attempted_ = false;

  // PoseController (pose_controller.h:262-321)
  PoseResetMode pose_reset_mode_ = NO_RESET;
  Vec3 translation_velocity_input_, rotation_velocity_input_;
  Pose manual_pose_, auto_pose_, imu_pose_, inclination_pose_, default_pose_, tip_align_pose_, origin_tip_align_pose_,
      walk_plane_pose_, origin_walk_plane_pose_;
  bool executing_transition_ = false;
  int legs_completed_step_ = 0, current_group_ = 0, pack_step_ = 0;
  bool reset_transition_sequence_ = true;
  int transition_step_ = 0, transition_step_count_ = 0;  // pose_controller.h:296-304
  bool set_target_ = true, proximity_alert_ = false, horizontal_transition_complete_ = false,
       vertical_transition_complete_ = false, first_sequence_execution_ = true;
  bool sequence_failed_ = false;  // set where the reference would ROS_FATAL + shutdown (pose_controller.cpp:438-442)
  int auto_pose_reference_leg_ = 0;
  std::vector<AutoPoser> auto_posers_;
  PosingState auto_posing_state_ = POSING_COMPLETE;
  int pose_phase_ = 0;
  double pose_frequency_ = 0.0;
  int pose_phase_length_ = 0, normaliser_ = 1;
  Vec3 rotation_absement_error_, rotation_position_error_, rotation_velocity_error_;
  bool imu_unstable_ = false;  // set where the reference would ROS_FATAL + shutdown (pose_controller.cpp:1228)

  // StateController (state_controller.h:330-370)
  RobotState robot_state_ = UNKNOWN, new_robot_state_ = UNKNOWN;
  bool transition_state_flag_ = false;
  double linear_velocity_input_[2] = {0, 0};  // trap 12: uninitialised in the reference; the harness zeroes it
  double angular_velocity_input_ = 0;

  explicit Robot(const shc_config& cfg);
  // Copyable member-wise; the copy's back-pointers must then be re-seated (see cloneRobot in shc_oracle_capi.cpp).

  // Model
  void initLegs(bool use_default);
  void updateDefaultConfiguration();
  void generateWorkspaces();
  void updateModel();
  Vec3 estimateGravity();
  ImuData getImuData() const;
  void setImuData(const Quat& q, const Vec3& acc, const Vec3& gyro);

  // WalkController
  void walkerInit();
  void generateWalkspace();
  void generateLimits();
  StepCycle generateStepCycle(bool set_step_cycle = true);
  double getLimit(const double lin[2], double ang, const LimitMap& limit);
  void updateWalk(const double lin[2], double ang);
  void updateWalkPlane();
  Pose calculateOdometry(double time_period);

  // PoseController
  void poserInit();
  void setAutoPoseParams();
  void updateStance();
  int directStartup();
  int stepToNewStance();
  int executeSequence(SequenceSelection sequence);
  bool legsBearingLoad();
  int packLegs(double time_to_pack);
  int unpackLegs(double time_to_unpack);
  void updateCurrentPose(RobotState robot_state);
  void updateManualPose();
  void updateTipAlignPose();
  void updateWalkPlanePose();
  void updateAutoPose();
  void updateIMUPose();
  void updateInclinationPose();

  // AdmittanceController
  void updateAdmittance();
  void updateStiffness();

  // StateController harness
  void stateInit();
  void requestRobotState(RobotState input_state);  // robotStateCallback, state_controller.cpp:1098
  void setBodyVelocityInput(double vx, double vy, double wz);  // bodyVelocityInputCallback :1127
  void loop();
  void transitionRobotState();
  void runningState();
  // convenience: run the direct start-up until READY, then enter RUNNING (returns number of loop() calls)
  int startUp();
};

}  // namespace shc_oracle
