// shc_facade.hpp — C++ façade over the C-ABI (shc_b200.h) that keeps the reference's class and method names, so the
// per-cycle call sequence of StateController::loop() / runningState() (state_controller.cpp:162-193, 379-447) reads the
// same against the batched B200 engine.  Header-only; link with libshc_b200.so.
//
// One `shc_b200::Batch` owns one engine (N robots on one GPU) and hands out, per robot index, the objects
// StateController holds: Model, WalkController, PoseController, AdmittanceController.  The per-cycle methods record
// their inputs; the last call of the cycle, Model::updateModel(), runs ONE fused control-cycle launch for the whole
// batch once every robot of the batch has reached it (immediately for N = 1), through shc_step_host.  Getters read a
// lazily refreshed host copy of the state (shc_get_state), replacing the member reads of the reference's publishers
// (state_controller.cpp:777-1078).  Eigen is not required: Vector2d/Vector3d/Quaterniond/Pose below are plain structs
// with the member names the reference code uses (position_, rotation_, x(), w() ...).
//
// Reference interfaces mirrored (file:line under the reference tree):
//   Model            include/syropod_highlevel_controller/model.h:58-182
//   Leg / Joint      model.h:196-535, 573-652
//   WalkController   walk_controller.h:54-277          LegStepper  walk_controller.h:286-535
//   PoseController   pose_controller.h:36-321          AdmittanceController  admittance_controller.h
#pragma once
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

 private:
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
  inline void setManualPoseInput(const Vector3d& translation, const Vector3d& rotation);  // pose_controller.h:116
  inline void setPoseResetMode(const PoseResetMode& mode);                                // pose_controller.h:105
  inline PosingState getAutoPoseState();
  inline Pose getAutoPose();
  inline Vector3d getRotationAbsementError();
  inline Vector3d getRotationVelocityError();

 private:
  Batch* b_;
  int robot_;
};

class AdmittanceController {  // admittance_controller.h
 public:
  AdmittanceController(Batch* b, int robot) : b_(b), robot_(robot) {}
  void updateAdmittance() {}                            // admittance_controller.cpp:22, inside the fused cycle
  void updateStiffness(std::shared_ptr<WalkController>) {}  // :96 — only feeds LegState.msg in the reference (trap 3)

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
    for (int r = 0; r < n_; ++r) {
      Controllers c;
      c.model_ = std::make_shared<Model>(this, r);
      c.walker_ = std::make_shared<WalkController>(this, r);
      c.poser_ = std::make_shared<PoseController>(this, r);
      c.admittance_ = std::make_shared<AdmittanceController>(this, r);
      robots_.push_back(c);
    }
    refresh();
  }
  ~Batch() { shc_destroy(e_); }
  Batch(const Batch&) = delete;
  Batch& operator=(const Batch&) = delete;

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
    stale_ = true;
    ++cycles_;
  }
  const shc_robot_state& state(int r) {
    if (stale_) refresh();
    return state_[r];
  }
  const shc_startup& startup() {
    if (!have_startup_) { shc_get_startup(e_, &startup_); have_startup_ = true; }
    return startup_;
  }
  long cycles() const { return cycles_; }

 private:
  void refresh() {
    if (shc_get_state(e_, state_.data(), state_.size()) != SHC_OK)
      throw std::runtime_error(std::string("shc_get_state: ") + shc_last_error());
    stale_ = false;
  }
  Parameters params_;
  int n_;
  shc_engine* e_ = nullptr;
  std::vector<Controllers> robots_;
  std::vector<float> cmd_, imu_, force_, manual_, joints_;
  std::vector<char> reached_;
  int arrived_ = 0;
  std::vector<shc_robot_state> state_;
  bool stale_ = true, have_startup_ = false;
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
inline void Leg::setTipForceMeasured(const Vector3d& f) { b_->setTipForce(robot_, leg_, f); }

inline Model::Model(Batch* b, int robot) : b_(b), robot_(robot) {
  for (int l = 0; l < b->params().cfg.leg_count; ++l) legs_.push_back(std::make_shared<Leg>(b, robot, l));
}
inline double Model::getTimeDelta() const { return b_->params().cfg.time_delta; }
inline Pose Model::getCurrentPose() { return Pose(b_->state(robot_).current_pose); }
inline void Model::setImuData(const Quaterniond& q, const Vector3d& acc, const Vector3d& gyro) { b_->setImu(robot_, q, acc, gyro); }
inline void Model::updateModel() { b_->arrive(robot_); }

inline Pose LegStepper::getCurrentTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].tip_position); p.rotation_ = Quaterniond(0, 0, 0, 0); return p; }
inline Pose LegStepper::getDefaultTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].default_tip_position); p.rotation_ = Quaterniond(0, 0, 0, 0); return p; }
inline Pose LegStepper::getTargetTipPose() { Pose p; p.position_ = Vector3d(b_->state(robot_).legs[leg_].target_tip_position); p.rotation_ = Quaterniond(0, 0, 0, 0); return p; }
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

inline void PoseController::setManualPoseInput(const Vector3d& t, const Vector3d& r) { b_->setManual(robot_, t, r); }
inline void PoseController::setPoseResetMode(const PoseResetMode& mode) { b_->setPoseResetMode(int(mode)); }
inline PosingState PoseController::getAutoPoseState() { return PosingState(b_->state(robot_).auto_posing_state); }
inline Pose PoseController::getAutoPose() { return Pose(b_->state(robot_).auto_pose); }
inline Vector3d PoseController::getRotationAbsementError() { return Vector3d(b_->state(robot_).rotation_absement_error); }
inline Vector3d PoseController::getRotationVelocityError() { return Vector3d(b_->state(robot_).rotation_velocity_error); }

}  // namespace shc_b200
