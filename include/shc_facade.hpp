// shc_facade.hpp — C++ façade over the C-ABI (shc_b200.h) that keeps the reference's class and method names, so the
// per-cycle call sequence of StateController::loop() / runningState() (state_controller.cpp:162-193, 379-447) reads the
// same against the batched B200 engine.  Header-only; link with libshc_b200.so.
//
// One `shc_b200::Batch` owns one engine (N robots on one GPU) and hands out, per robot index, the objects
// StateController holds: Model, WalkController, PoseController, AdmittanceController.  The per-cycle methods record
// their inputs; the last call of the cycle, Model::updateModel(), runs ONE fused control-cycle launch for the whole
// batch once every robot of the batch has reached it (immediately for N = 1), through shc_step_host.  Getters read a
// host copy of ONE robot's record, fetched on first use after a cycle (shc_get_state_range: three small copies however
// large the batch is), replacing the member reads of the reference's publishers (state_controller.cpp:777-1078).  Eigen is not required: Vector2d/Vector3d/Quaterniond/Pose below are plain structs
// with the member names the reference code uses (position_, rotation_, x(), w() ...).
//
// Reference interfaces mirrored (file:line under the reference tree):
//   Model            include/syropod_highlevel_controller/model.h:58-182
//   Leg / Joint      model.h:196-535, 573-652
//   WalkController   walk_controller.h:54-277          LegStepper  walk_controller.h:286-535
//   PoseController   pose_controller.h:36-321          AdmittanceController  admittance_controller.h
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "shc_b200.h"

namespace shc_b200 {

struct Vector2d {
  double v[2] = {0, 0};
  Vector2d() = default;
  Vector2d(double x, double y) : v{x, y} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1]); }
};
struct Vector3d {
  double v[3] = {0, 0, 0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  explicit Vector3d(const double* p) : v{p[0], p[1], p[2]} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Quaterniond {
  double w_ = 1, x_ = 0, y_ = 0, z_ = 0;
  Quaterniond() = default;
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
};
struct Pose {  // pose.h:17
  Vector3d position_;
  Quaterniond rotation_;
  Pose() = default;
  explicit Pose(const double* p7) : position_(p7), rotation_(p7[3], p7[4], p7[5], p7[6]) {}
};

// parameters_and_states.h:27-143
enum RobotState { PACKED, READY, RUNNING, ROBOT_STATE_COUNT, UNKNOWN = -1, OFF = -2 };
enum WalkState { STARTING, MOVING, STOPPING, STOPPED, WALK_STATE_COUNT };
enum StepState { SWING, STANCE, FORCE_STANCE, FORCE_STOP, STEP_STATE_COUNT };
enum PosingState { POSING, STOP_POSING, POSING_COMPLETE, POSING_STATE_COUNT };
enum SequenceSelection { START_UP, SHUT_DOWN, SEQUENCE_SELECTION_COUNT };  // parameters_and_states.h:183
enum PoseResetMode { NO_RESET, Z_AND_YAW_RESET, X_AND_Y_RESET, PITCH_AND_ROLL_RESET, ALL_RESET, IMMEDIATE_ALL_RESET };
enum LegState { WALKING, MANUAL, LEG_STATE_COUNT, WALKING_TO_MANUAL = -1, MANUAL_TO_WALKING = -2 };

struct StepCycle {  // walk_controller.h:23
  double frequency_;
  int period_, swing_period_, stance_period_, stance_end_, swing_start_, swing_end_, stance_start_;
};

struct Parameters {  // parameters_and_states.h:271 (the fields the hot path reads)
  shc_config cfg;
};

class Batch;

// ---------------------------------------------------------------------------------------------------------------------
class Joint {  // model.h:573
 public:
  std::string id_name_;
  double desired_position_ = 0, desired_velocity_ = 0, offset_ = 0;
  double min_position_ = 0, max_position_ = 0, max_angular_speed_ = 0;
};

class LegStepper {  // walk_controller.h:286 — read-only view
 public:
  LegStepper(Batch* b, int robot, int leg) : b_(b), robot_(robot), leg_(leg) {}
  inline Pose getCurrentTipPose();
  inline Pose getDefaultTipPose();
  inline Pose getTargetTipPose();
  inline Vector3d getStrideVector();
  inline Vector3d getWalkPlane();
  inline Vector3d getWalkPlaneNormal();
  inline StepState getStepState();
  inline WalkState getWalkState();
  inline int getPhase();
  inline double getSwingProgress();
  inline double getStanceProgress();
  inline bool hasCompletedFirstStep();
  inline bool isAtCorrectPhase();
  inline Pose getIdentityTipPose();  // walk_controller.h:323 (stance position of the leg, rotation undefined)
  inline Vector3d getSwingOriginTipPosition();
  inline Vector3d getSwingOriginTipVelocity();
  inline Vector3d getStanceOriginTipPosition();
  /// Control nodes of the leg's quartic Beziers (walk_controller.h:327-333), i = 0..4, rebuilt from the leg's record with
  /// the reference's formulas (generatePrimary/SecondarySwingControlNodes :1238-1291, generateStanceControlNodes :1295).
  /// The reference regenerates them every cycle the leg is on that curve and keeps the last set otherwise (RViz shows
  /// stale nodes of the other curve); here they are always those of the leg's CURRENT record.
  inline Vector3d getSwing1ControlNode(int i);
  inline Vector3d getSwing2ControlNode(int i);
  inline Vector3d getStanceControlNode(int i);

 private:
  inline void swingNodes(Vector3d n1[5], Vector3d n2[5]);
  Batch* b_;
  int robot_, leg_;
};

class Leg {  // model.h:196
 public:
  Leg(Batch* b, int robot, int leg);
  int getIDNumber() const { return leg_; }
  int getJointCount() const { return int(joints_.size()); }
  LegState getLegState() const { return WALKING; }
  std::shared_ptr<LegStepper> getLegStepper() { return stepper_; }
  inline std::shared_ptr<Joint> getJointByIDNumber(int joint_id_number);  // 1-based, as the reference
  inline Pose getCurrentTipPose();
  inline Pose getDesiredTipPose();
  inline Vector3d getAdmittanceDelta();
  inline Vector3d getTipForceCalculated();
  inline double getVirtualStiffness();  // model.h:264 (dynamic stiffness of AdmittanceController::updateStiffness)
  inline void setTipForceMeasured(const Vector3d& f);  // tipStatesCallback (state_controller.cpp:1645)

 private:
  friend class Batch;
  Batch* b_;
  int robot_, leg_;
  std::vector<std::shared_ptr<Joint>> joints_;
  std::shared_ptr<LegStepper> stepper_;
};

class Model {  // model.h:58
 public:
  Model(Batch* b, int robot);
  int getLegCount() const { return int(legs_.size()); }
  std::shared_ptr<Leg> getLegByIDNumber(int id) { return legs_.at(id); }
  inline double getTimeDelta() const;
  inline Pose getCurrentPose();
  inline void setImuData(const Quaterniond& orientation, const Vector3d& linear_acceleration,
                         const Vector3d& angular_velocity);  // model.h:146
  /// Model::updateModel (model.cpp:142): the last call of runningState(); runs the fused control cycle for the batch.
  inline void updateModel();

 private:
  friend class Batch;
  Batch* b_;
  int robot_;
  std::vector<std::shared_ptr<Leg>> legs_;
};

class WalkController {  // walk_controller.h:54
 public:
  WalkController(Batch* b, int robot) : b_(b), robot_(robot) {}
  /// WalkController::updateWalk (walk_controller.cpp:440): records the command; executed inside the fused cycle.
  inline void updateWalk(const Vector2d& linear_velocity_input, const double& angular_velocity_input);
  /// updateManual overloads (walk_controller.cpp:652, :711): no leg can be in MANUAL state in the batched engine.
  void updateManual(int, const Vector3d&, int, const Vector3d&) {}
  void updateManual(int, const Pose&, int, const Pose&) {}
  void setPoseState(const PosingState&) {}  // carried inside the engine state (walk_controller.h:120)
  inline WalkState getWalkState();
  inline Vector2d getDesiredLinearVelocity();
  inline double getDesiredAngularVelocity();
  inline Vector3d getWalkPlane();
  inline Vector3d getWalkPlaneNormal();
  inline Pose getOdometryIdeal();
  inline StepCycle getStepCycle();
  inline std::array<double, SHC_N_BEARINGS> getWalkspace();  // radii at bearings 0,45,...,360 (LimitMap)
  inline double getTimeDelta() const;
  inline double getStepClearance() const;
  inline double getBodyClearance() const;
  /// WalkController::getLimit (walk_controller.cpp:414) on the robot's current tips and limit tables.
  inline double getLimit(const Vector2d& linear_velocity_input, const double& angular_velocity_input,
                         const std::array<double, SHC_N_BEARINGS>& limit);
  inline std::array<double, SHC_N_BEARINGS> getLinearSpeedLimitMap();
  inline std::array<double, SHC_N_BEARINGS> getAngularSpeedLimitMap();
  inline std::array<double, SHC_N_BEARINGS> getLinearAccelerationLimitMap();
  inline std::array<double, SHC_N_BEARINGS> getAngularAccelerationLimitMap();
  /// walk_controller.h:126-141 — batch-wide in the engine (every robot shares the constants block)
  inline void setLinearSpeedLimitMap(const std::array<double, SHC_N_BEARINGS>& m);
  inline void setAngularSpeedLimitMap(const std::array<double, SHC_N_BEARINGS>& m);
  inline void setLinearAccelerationLimitMap(const std::array<double, SHC_N_BEARINGS>& m);
  inline void setAngularAccelerationLimitMap(const std::array<double, SHC_N_BEARINGS>& m);

 private:
  Batch* b_;
  int robot_;
};

class PoseController {  // pose_controller.h:36
 public:
  PoseController(Batch* b, int robot) : b_(b), robot_(robot) {}
  /// PoseController::updateCurrentPose (pose_controller.cpp:811): executed inside the fused cycle.
  void updateCurrentPose(const RobotState&) {}
  /// PoseController::updateStance (pose_controller.cpp:110): executed inside the fused cycle.
  void updateStance() {}
  /// Sequences (one loop() per call, pose_controller.cpp:463 / :520 / :597 / :661): they act on the whole batch — the first
  /// robot to call in a loop runs the device step for everyone, the other robots' calls of that loop read their own result.
  inline int directStartup();
  /// PoseController::executeSequence (pose_controller.cpp:145): START_UP / SHUT_DOWN; -1 while the first start-up generates
  /// its sequence, else 0..100.  A robot that is through a sequence is not stepped by the other robots' further loops.
  inline int executeSequence(const SequenceSelection& sequence);
  inline int stepToNewStance();
  inline int packLegs(const double& time_to_pack);
  inline int unpackLegs(const double& time_to_unpack);
  inline void setManualPoseInput(const Vector3d& translation, const Vector3d& rotation);  // pose_controller.h:116
  inline void setPoseResetMode(const PoseResetMode& mode);                                // pose_controller.h:105
  inline PosingState getAutoPoseState();
  inline Pose getAutoPose();
  inline Vector3d getRotationAbsementError();
  inline Vector3d getRotationVelocityError();
  inline Pose getWalkPlanePose();
  inline Pose getManualPose();
  inline Pose getImuPose();
  inline Pose getInclinationPose();

 private:
  Batch* b_;
  int robot_;
};

class AdmittanceController {  // admittance_controller.h
 public:
  AdmittanceController(Batch* b, int robot) : b_(b), robot_(robot) {}
  void updateAdmittance() {}                            // admittance_controller.cpp:22, inside the fused cycle
  void updateStiffness(std::shared_ptr<WalkController>) {}  // :96, inside the fused cycle (dynamic_stiffness); read it
                                                            // back with Leg::getVirtualStiffness

 private:
  Batch* b_;
  int robot_;
};

struct Controllers {
  std::shared_ptr<Model> model_;
  std::shared_ptr<WalkController> walker_;
  std::shared_ptr<PoseController> poser_;
  std::shared_ptr<AdmittanceController> admittance_;
};

// ---------------------------------------------------------------------------------------------------------------------
class Batch {
 public:
  Batch(const Parameters& params, int n_robots, int device = 0, int precision = SHC_PRECISION_F64,
        const shc_startup* startup = nullptr)
      : params_(params), n_(n_robots) {
    if (shc_create(&params.cfg, startup, n_robots, device, precision, &e_) != SHC_OK)
      throw std::runtime_error(std::string("shc_create: ") + shc_last_error());
    const int L = params.cfg.leg_count, D = params.cfg.joint_count;
    cmd_.assign(size_t(n_) * 3, 0.f);
    joints_.assign(size_t(n_) * L * D, 0.f);
    reached_.assign(n_, 0);
    state_.resize(n_);
    fetched_.assign(n_, -1);
    for (int r = 0; r < n_; ++r) {
      Controllers c;
      c.model_ = std::make_shared<Model>(this, r);
      c.walker_ = std::make_shared<WalkController>(this, r);
      c.poser_ = std::make_shared<PoseController>(this, r);
      c.admittance_ = std::make_shared<AdmittanceController>(this, r);
      robots_.push_back(c);
    }
  }
  ~Batch() { shc_destroy(e_); }
  Batch(const Batch&) = delete;
  Batch& operator=(const Batch&) = delete;

  /// StateController::changeGait for the whole batch (state_controller.cpp:513-540), or a change of a constants-only
  /// adjustable parameter: the engine is replaced by one for the new parameters that carries the state over
  /// (shc_clone_reconfigured).  As in the reference, switch once every robot's walker has STOPPED, in place of a cycle.
  void changeGait(const Parameters& params, const shc_startup* startup = nullptr) { reconfigure(params, startup, 0); }
  /// StateController::adjustParameter for the whole batch (state_controller.cpp:451-508): as changeGait, but the auto-pose
  /// cycle keeps its length (the reference regenerates it in changeGait only).
  void adjustParameter(const Parameters& params, const shc_startup* startup = nullptr) {
    reconfigure(params, startup, SHC_RECONF_KEEP_POSE_CYCLE);
  }

 private:
  void reconfigure(const Parameters& params, const shc_startup* startup, int flags) {
    shc_engine* e2 = nullptr;
    if (shc_clone_reconfigured(e_, &params.cfg, startup, flags, &e2) != SHC_OK)
      throw std::runtime_error(std::string("shc_clone_reconfigured: ") + shc_last_error());
    shc_destroy(e_);
    e_ = e2;
    params_ = params;
    std::fill(fetched_.begin(), fetched_.end(), -1);
  }

 public:
  Controllers& robot(int r) { return robots_.at(r); }
  int size() const { return n_; }
  shc_engine* engine() { return e_; }
  const Parameters& params() const { return params_; }
  /// Desired joint positions (+offset) of the last cycle, [N][L][D] — what publishDesiredJointState sends
  /// (state_controller.cpp:777-805).
  const std::vector<float>& desiredJointPositions() const { return joints_; }

  // ---- used by the façade classes ----
  void setCommand(int r, double vx, double vy, double wz) {
    cmd_[size_t(r) * 3 + 0] = float(vx);
    cmd_[size_t(r) * 3 + 1] = float(vy);
    cmd_[size_t(r) * 3 + 2] = float(wz);
  }
  void setImu(int r, const Quaterniond& q, const Vector3d& acc, const Vector3d& gyro) {
    if (imu_.empty()) {
      imu_.assign(size_t(n_) * 10, 0.f);
      for (int i = 0; i < n_; ++i) imu_[size_t(i) * 10] = 1.f;
    }
    float* m = &imu_[size_t(r) * 10];
    m[0] = float(q.w()); m[1] = float(q.x()); m[2] = float(q.y()); m[3] = float(q.z());
    for (int k = 0; k < 3; ++k) { m[4 + k] = float(gyro[k]); m[7 + k] = float(acc[k]); }
  }
  void setTipForce(int r, int leg, const Vector3d& f) {
    const int L = params_.cfg.leg_count;
    if (force_.empty()) force_.assign(size_t(n_) * L * 3, 0.f);
    for (int k = 0; k < 3; ++k) force_[(size_t(r) * L + leg) * 3 + k] = float(f[k]);
  }
  void setManual(int r, const Vector3d& t, const Vector3d& rot) {
    if (manual_.empty()) manual_.assign(size_t(n_) * 6, 0.f);
    for (int k = 0; k < 3; ++k) { manual_[size_t(r) * 6 + k] = float(t[k]); manual_[size_t(r) * 6 + 3 + k] = float(rot[k]); }
  }
  void setPoseResetMode(int mode) { shc_set_pose_reset_mode(e_, mode); }
  /// Called by Model::updateModel of robot r; steps the batch once every robot has arrived.
  void arrive(int r) {
    if (!reached_[r]) { reached_[r] = 1; ++arrived_; }
    if (arrived_ == n_) step();
  }
  void step() {
    if (shc_step_host(e_, cmd_.data(), imu_.empty() ? nullptr : imu_.data(), force_.empty() ? nullptr : force_.data(),
                      manual_.empty() ? nullptr : manual_.data(), joints_.data()) != SHC_OK)
      throw std::runtime_error(std::string("shc_step_host: ") + shc_last_error());
    std::fill(reached_.begin(), reached_.end(), 0);
    arrived_ = 0;
    ++cycles_;
  }
  /// Record of robot r as of the last cycle: fetched from the device the first time it is asked for after a cycle.
  const shc_robot_state& state(int r) {
    if (fetched_.at(r) != cycles_) {
      if (shc_get_state_range(e_, size_t(r), 1, &state_[r]) != SHC_OK)
        throw std::runtime_error(std::string("shc_get_state_range: ") + shc_last_error());
      fetched_[r] = cycles_;
    }
    return state_[r];
  }
  /// One loop() of a sequence as seen by robot r (see PoseController::directStartup ...).
  int sequence(int r, int kind, double time) {
    if (seq_used_.empty()) { seq_used_.assign(n_, 1); seq_progress_.assign(n_, 0); }
    if (seq_used_[r] || kind != seq_kind_) {
      if (shc_sequence_step_host(e_, kind, time, joints_.data(), seq_progress_.data()) < 0)
        throw std::runtime_error(std::string("shc_sequence_step_host: ") + shc_last_error());
      std::fill(seq_used_.begin(), seq_used_.end(), 0);
      seq_kind_ = kind;
      ++cycles_;
    }
    seq_used_[r] = 1;
    return seq_progress_[r];
  }
  void setLimitMaps(const double* ls, const double* as, const double* la, const double* aa) {
    if (shc_set_limit_maps(e_, ls, as, la, aa) != SHC_OK) throw std::runtime_error(std::string("shc_set_limit_maps: ") + shc_last_error());
    have_startup_ = false;
  }
  const shc_startup& startup() {
    if (!have_startup_) { shc_get_startup(e_, &startup_); have_startup_ = true; }
    return startup_;
  }
  long cycles() const { return cycles_; }

 private:
  Parameters params_;
  int n_;
  shc_engine* e_ = nullptr;
  std::vector<Controllers> robots_;
  std::vector<float> cmd_, imu_, force_, manual_, joints_;
  std::vector<char> reached_, seq_used_;
  std::vector<int> seq_progress_;
  int seq_kind_ = -1;
  int arrived_ = 0;
  std::vector<shc_robot_state> state_;
  std::vector<long> fetched_;  // cycle count at which state_[r] was fetched (-1: never)
  bool have_startup_ = false;
  shc_startup startup_;
  long cycles_ = 0;
};

// ---- inline definitions ---------------------------------------------------------------------------------------------
inline Leg::Leg(Batch* b, int robot, int leg) : b_(b), robot_(robot), leg_(leg) {
  const shc_config& c = b->params().cfg;
  for (int j = 0; j < c.joint_count; ++j) {
    auto joint = std::make_shared<Joint>();
    joint->id_name_ = "leg" + std::to_string(leg) + "_joint" + std::to_string(j + 1);
    joint->offset_ = c.joint_offset[leg][j];
    joint->min_position_ = c.joint_min[leg][j];
    joint->max_position_ = c.joint_max[leg][j];
    joint->max_angular_speed_ = c.joint_max_vel[leg][j];
    joints_.push_back(joint);
  }
  stepper_ = std::make_shared<LegStepper>(b, robot, leg);
}
inline std::shared_ptr<Joint> Leg::getJointByIDNumber(int id) {
  const shc_leg_state& s = b_->state(robot_).legs[leg_];
  auto j = joints_.at(id - 1);
  j->desired_position_ = s.joint_position[id - 1];
  j->desired_velocity_ = s.joint_velocity[id - 1];
  return j;
}
inline Pose Leg::getCurrentTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].model_tip_position); return p; }
inline Pose Leg::getDesiredTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].desired_tip_position); return p; }
inline Vector3d Leg::getAdmittanceDelta() { return Vector3d(b_->state(robot_).legs[leg_].admittance_delta); }
inline Vector3d Leg::getTipForceCalculated() { return Vector3d(b_->state(robot_).legs[leg_].tip_force_calculated); }
inline double Leg::getVirtualStiffness() { return b_->state(robot_).legs[leg_].virtual_stiffness; }
inline void Leg::setTipForceMeasured(const Vector3d& f) { b_->setTipForce(robot_, leg_, f); }

inline Model::Model(Batch* b, int robot) : b_(b), robot_(robot) {
  for (int l = 0; l < b->params().cfg.leg_count; ++l) legs_.push_back(std::make_shared<Leg>(b, robot, l));
}
inline double Model::getTimeDelta() const { return b_->params().cfg.time_delta; }
inline Pose Model::getCurrentPose() { return Pose(b_->state(robot_).current_pose); }
inline void Model::setImuData(const Quaterniond& q, const Vector3d& acc, const Vector3d& gyro) { b_->setImu(robot_, q, acc, gyro); }
inline void Model::updateModel() { b_->arrive(robot_); }

// (tip rotations are all zero = UNDEFINED_ROTATION unless gravity_aligned_tips is live on legs of more than three joints)
inline Pose LegStepper::getCurrentTipPose() {
  const shc_leg_state& g = b_->state(robot_).legs[leg_];
  Pose p; p.position_ = Vector3d(g.tip_position); p.rotation_ = Quaterniond(g.tip_rotation[0], g.tip_rotation[1], g.tip_rotation[2], g.tip_rotation[3]);
  return p;
}
inline Pose LegStepper::getDefaultTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].default_tip_position); p.rotation_ = Quaterniond(0, 0, 0, 0); return p; }
inline Pose LegStepper::getTargetTipPose() {
  const shc_leg_state& g = b_->state(robot_).legs[leg_];
  Pose p; p.position_ = Vector3d(g.target_tip_position);
  p.rotation_ = Quaterniond(g.target_tip_rotation[0], g.target_tip_rotation[1], g.target_tip_rotation[2], g.target_tip_rotation[3]);
  return p;
}
inline Vector3d LegStepper::getStrideVector() { return Vector3d(b_->state(robot_).legs[leg_].stride_vector); }
inline Vector3d LegStepper::getWalkPlane() { return Vector3d(b_->state(robot_).legs[leg_].walk_plane); }
inline Vector3d LegStepper::getWalkPlaneNormal() { return Vector3d(b_->state(robot_).legs[leg_].walk_plane_normal); }
inline StepState LegStepper::getStepState() { return StepState(b_->state(robot_).legs[leg_].step_state); }
inline WalkState LegStepper::getWalkState() { return WalkState(b_->state(robot_).walk_state); }
inline int LegStepper::getPhase() { return b_->state(robot_).legs[leg_].phase; }
inline double LegStepper::getSwingProgress() { return b_->state(robot_).legs[leg_].swing_progress; }
inline double LegStepper::getStanceProgress() { return b_->state(robot_).legs[leg_].stance_progress; }
inline bool LegStepper::hasCompletedFirstStep() { return b_->state(robot_).legs[leg_].completed_first_step != 0; }
inline bool LegStepper::isAtCorrectPhase() { return b_->state(robot_).legs[leg_].at_correct_phase != 0; }

inline Pose LegStepper::getIdentityTipPose() {
  Pose p;
  p.position_ = Vector3d(b_->params().cfg.stance_x[leg_], b_->params().cfg.stance_y[leg_], 0.0);
  p.rotation_ = Quaterniond(0, 0, 0, 0);
  return p;
}
inline Vector3d LegStepper::getSwingOriginTipPosition() { return Vector3d(b_->state(robot_).legs[leg_].swing_origin_position); }
inline Vector3d LegStepper::getSwingOriginTipVelocity() { return Vector3d(b_->state(robot_).legs[leg_].swing_origin_velocity); }
inline Vector3d LegStepper::getStanceOriginTipPosition() { return Vector3d(b_->state(robot_).legs[leg_].stance_origin_position); }
inline void LegStepper::swingNodes(Vector3d n1[5], Vector3d n2[5]) {
  const shc_config& c = b_->params().cfg;
  const shc_startup& su = b_->startup();
  const shc_robot_state& rs = b_->state(robot_);
  const shc_leg_state& g = rs.legs[leg_];
  const double dt = c.time_delta;
  // LegStepper::updateTipPosition (walk_controller.cpp:1035-1041): iteration counts in double, as written there
  int swing_iterations = int((double(su.swing_period) / su.period) / (su.step_frequency * dt));
  swing_iterations = (swing_iterations % 2 == 0) ? swing_iterations : swing_iterations + 1;
  const double swing_dt = 1.0 / (swing_iterations / 2.0);
  const bool standard = g.step_state == SWING || g.completed_first_step;
  int stance_period = standard ? ((su.stance_end - su.stance_start) % su.period + su.period) % su.period
                               : ((su.stance_end - su.phase_offsets[leg_]) % su.period + su.period) % su.period;
  if (stance_period == 0) stance_period = su.period;
  const int stance_iterations = int((double(stance_period) / su.period) / (su.step_frequency * dt));
  const double stance_dt = 1.0 / stance_iterations;
  auto V = [](const double* p) { return Vector3d(p); };
  auto add = [](Vector3d a, Vector3d b) { return Vector3d(a[0] + b[0], a[1] + b[1], a[2] + b[2]); };
  auto sub = [](Vector3d a, Vector3d b) { return Vector3d(a[0] - b[0], a[1] - b[1], a[2] - b[2]); };
  auto mul = [](Vector3d a, double k) { return Vector3d(a[0] * k, a[1] * k, a[2] * k); };
  Vector3d nrm = V(g.walk_plane_normal);
  const double nn = nrm.norm();
  if (nn > 0) nrm = mul(nrm, 1.0 / nn);
  const Vector3d clearance = mul(nrm, c.swing_height);  // updateStride (:944)
  const Vector3d origin = V(g.swing_origin_position), target = V(g.target_tip_position);
  Vector3d mid = mul(add(origin, target), 0.5);
  mid[2] = std::max(origin[2], target[2]);
  mid = add(mid, clearance);
  mid[1] += (c.stance_y[leg_] > 0 ? 1.0 : -1.0) * c.swing_width;
  const Vector3d sep1 = mul(V(g.swing_origin_velocity), 0.25 * (dt / swing_dt));
  n1[0] = origin;
  n1[1] = add(origin, sep1);
  n1[2] = add(origin, mul(sep1, 2.0));
  n1[3] = mul(add(mid, n1[2]), 0.5);
  n1[3][2] = mid[2];
  n1[4] = mid;
  const Vector3d final_tip_velocity = mul(V(g.stride_vector), -(stance_dt / dt));
  const Vector3d sep2 = mul(final_tip_velocity, 0.25 * (dt / swing_dt));
  n2[0] = n1[4];
  n2[1] = sub(n1[4], sub(n1[3], n1[4]));
  n2[2] = sub(target, mul(sep2, 2.0));
  n2[3] = sub(target, sep2);
  n2[4] = target;
  if (c.force_normal_touchdown) {  // forceNormalTouchdown (:1314)
    Vector3d bo = sub(target, mul(sep2, 4.0));
    bo[2] = std::max(origin[2], target[2]);
    bo = add(bo, clearance);
    n1[4] = bo;
    n2[0] = bo;
    n1[3] = sub(n2[0], mul(sub(n2[2], bo), 0.5));
    n2[1] = add(n2[0], mul(sub(n2[2], bo), 0.5));
  }
}
inline Vector3d LegStepper::getSwing1ControlNode(int i) { Vector3d a[5], b[5]; swingNodes(a, b); return a[i]; }
inline Vector3d LegStepper::getSwing2ControlNode(int i) { Vector3d a[5], b[5]; swingNodes(a, b); return b[i]; }
inline Vector3d LegStepper::getStanceControlNode(int i) {
  const shc_startup& su = b_->startup();
  const shc_leg_state& g = b_->state(robot_).legs[leg_];
  const bool standard = g.step_state == SWING || g.completed_first_step;
  const int std_period = ((su.stance_end - su.stance_start) % su.period + su.period) % su.period;
  int mod_period = ((su.stance_end - su.phase_offsets[leg_]) % su.period + su.period) % su.period;
  if (mod_period == 0) mod_period = su.period;
  const double scaler = standard ? 1.0 : double(mod_period) / std_period;
  Vector3d o(g.stance_origin_position), st(g.stride_vector);
  return Vector3d(o[0] - st[0] * scaler * 0.25 * i, o[1] - st[1] * scaler * 0.25 * i, o[2] - st[2] * scaler * 0.25 * i);
}

inline void WalkController::updateWalk(const Vector2d& lin, const double& ang) { b_->setCommand(robot_, lin[0], lin[1], ang); }
inline WalkState WalkController::getWalkState() { return WalkState(b_->state(robot_).walk_state); }
inline Vector2d WalkController::getDesiredLinearVelocity() { const auto& s = b_->state(robot_); return Vector2d(s.desired_linear_velocity[0], s.desired_linear_velocity[1]); }
inline double WalkController::getDesiredAngularVelocity() { return b_->state(robot_).desired_angular_velocity; }
inline Vector3d WalkController::getWalkPlane() { return Vector3d(b_->state(robot_).walk_plane); }
inline Vector3d WalkController::getWalkPlaneNormal() { return Vector3d(b_->state(robot_).walk_plane_normal); }
inline Pose WalkController::getOdometryIdeal() { return Pose(b_->state(robot_).odometry_ideal); }
inline StepCycle WalkController::getStepCycle() {
  const shc_startup& s = b_->startup();
  return StepCycle{s.step_frequency, s.period, s.swing_period, s.stance_period, s.stance_end, s.swing_start, s.swing_end, s.stance_start};
}
inline std::array<double, SHC_N_BEARINGS> WalkController::getWalkspace() {
  std::array<double, SHC_N_BEARINGS> w;
  for (int i = 0; i < SHC_N_BEARINGS; ++i) w[i] = b_->startup().walkspace[i];
  return w;
}
inline double WalkController::getTimeDelta() const { return b_->params().cfg.time_delta; }
inline double WalkController::getStepClearance() const { return b_->params().cfg.swing_height; }
inline double WalkController::getBodyClearance() const { return b_->params().cfg.body_clearance; }

inline double WalkController::getLimit(const Vector2d& lin, const double& ang, const std::array<double, SHC_N_BEARINGS>& limit) {
  const shc_robot_state& s = b_->state(robot_);
  double min_limit = 2147483647.0;  // UNASSIGNED_VALUE
  const double pi = 3.14159265358979323846;
  for (int l = 0; l < b_->params().cfg.leg_count; ++l) {
    const double* tip = s.legs[l].tip_position;
    const double sx = lin[0] + ang * (-tip[1]), sy = lin[1] + ang * tip[0];
    const double deg = std::atan2(sy, sx) / pi * 180.0;
    int bearing = int(deg >= 0 ? deg + 0.5 : -(0.5 - deg));          // roundToInt
    bearing = ((bearing % 360) + 360) % 360;                           // mod
    const int lower = bearing / 45 * 45, upper = lower + 45;
    // interpolation with int / int (trap 1): the weight is 0, the result the lower bucket's value
    const double v = limit[lower / 45] + (bearing - lower) / (upper - lower) * (limit[upper / 45 % SHC_N_BEARINGS] - limit[lower / 45]);
    min_limit = std::min(min_limit, v);
  }
  return min_limit;
}
inline std::array<double, SHC_N_BEARINGS> WalkController::getLinearSpeedLimitMap() { std::array<double, SHC_N_BEARINGS> m; for (int i = 0; i < SHC_N_BEARINGS; ++i) m[i] = b_->startup().max_linear_speed[i]; return m; }
inline std::array<double, SHC_N_BEARINGS> WalkController::getAngularSpeedLimitMap() { std::array<double, SHC_N_BEARINGS> m; for (int i = 0; i < SHC_N_BEARINGS; ++i) m[i] = b_->startup().max_angular_speed[i]; return m; }
inline std::array<double, SHC_N_BEARINGS> WalkController::getLinearAccelerationLimitMap() { std::array<double, SHC_N_BEARINGS> m; for (int i = 0; i < SHC_N_BEARINGS; ++i) m[i] = b_->startup().max_linear_acceleration[i]; return m; }
inline std::array<double, SHC_N_BEARINGS> WalkController::getAngularAccelerationLimitMap() { std::array<double, SHC_N_BEARINGS> m; for (int i = 0; i < SHC_N_BEARINGS; ++i) m[i] = b_->startup().max_angular_acceleration[i]; return m; }
inline void WalkController::setLinearSpeedLimitMap(const std::array<double, SHC_N_BEARINGS>& m) { b_->setLimitMaps(m.data(), nullptr, nullptr, nullptr); }
inline void WalkController::setAngularSpeedLimitMap(const std::array<double, SHC_N_BEARINGS>& m) { b_->setLimitMaps(nullptr, m.data(), nullptr, nullptr); }
inline void WalkController::setLinearAccelerationLimitMap(const std::array<double, SHC_N_BEARINGS>& m) { b_->setLimitMaps(nullptr, nullptr, m.data(), nullptr); }
inline void WalkController::setAngularAccelerationLimitMap(const std::array<double, SHC_N_BEARINGS>& m) { b_->setLimitMaps(nullptr, nullptr, nullptr, m.data()); }

inline int PoseController::directStartup() { return b_->sequence(robot_, SHC_SEQ_DIRECT_STARTUP, 0.0); }
inline int PoseController::executeSequence(const SequenceSelection& sequence) {
  return b_->sequence(robot_, sequence == START_UP ? SHC_SEQ_START_UP : SHC_SEQ_SHUT_DOWN, 0.0);
}
inline int PoseController::stepToNewStance() { return b_->sequence(robot_, SHC_SEQ_NEW_STANCE, 0.0); }
inline int PoseController::packLegs(const double& t) { return b_->sequence(robot_, SHC_SEQ_PACK, t); }
inline int PoseController::unpackLegs(const double& t) { return b_->sequence(robot_, SHC_SEQ_UNPACK, t); }
inline void PoseController::setManualPoseInput(const Vector3d& t, const Vector3d& r) { b_->setManual(robot_, t, r); }
inline void PoseController::setPoseResetMode(const PoseResetMode& mode) { b_->setPoseResetMode(int(mode)); }
inline PosingState PoseController::getAutoPoseState() { return PosingState(b_->state(robot_).auto_posing_state); }
inline Pose PoseController::getAutoPose() { return Pose(b_->state(robot_).auto_pose); }
inline Vector3d PoseController::getRotationAbsementError() { return Vector3d(b_->state(robot_).rotation_absement_error); }
inline Vector3d PoseController::getRotationVelocityError() { return Vector3d(b_->state(robot_).rotation_velocity_error); }
inline Pose PoseController::getWalkPlanePose() { return Pose(b_->state(robot_).walk_plane_pose); }
inline Pose PoseController::getManualPose() { return Pose(b_->state(robot_).manual_pose); }
inline Pose PoseController::getImuPose() { return Pose(b_->state(robot_).imu_pose); }
inline Pose PoseController::getInclinationPose() { return Pose(b_->state(robot_).inclination_pose); }

}  // namespace shc_b200
