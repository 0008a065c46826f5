// shc_oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.  extern "C" surface of the parity oracle for ctypes
// (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).  PINNED to the reference's own code (oracle/_ref, tests/test_reference_pin.py; see shc_oracle.hpp).
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "../include/shc_msgs.h"
#include "../include/shc_state.h"
#include "shc_oracle.hpp"

using namespace shc_oracle;

namespace {

void fixPointers(Robot* r) {
  for (int i = 0; i < SHC_MAX_LEGS; ++i) {
    r->legs[i].robot = r;
    r->legs[i].stepper.robot = r;
    r->legs[i].stepper.leg_ = &r->legs[i];
    r->legs[i].poser.robot = r;
    r->legs[i].poser.leg_ = &r->legs[i];
  }
  for (auto& ap : r->auto_posers_) ap.robot = r;
}

Robot* cloneRobot(const Robot& src) {
  Robot* r = new Robot(src);
  fixPointers(r);
  return r;
}

void putPose(double* o, const Pose& p) {
  o[0] = p.position_.x; o[1] = p.position_.y; o[2] = p.position_.z;
  o[3] = p.rotation_.w; o[4] = p.rotation_.x; o[5] = p.rotation_.y; o[6] = p.rotation_.z;
}
Pose getPose(const double* o) { return Pose(Vec3(o[0], o[1], o[2]), Quat(o[3], o[4], o[5], o[6])); }
void putQuat(double* o, const Quat& q) { o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; }
void put3(double* o, const Vec3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
Vec3 get3(const double* o) { return Vec3(o[0], o[1], o[2]); }

void exportState(const Robot& r, shc_robot_state* s) {
  std::memset(s, 0, sizeof(*s));
  s->desired_linear_velocity[0] = r.desired_linear_velocity_[0];
  s->desired_linear_velocity[1] = r.desired_linear_velocity_[1];
  s->desired_angular_velocity = r.desired_angular_velocity_;
  s->walk_state = r.walk_state_;
  s->legs_at_correct_phase = r.legs_at_correct_phase_;
  s->legs_completed_first_step = r.legs_completed_first_step_;
  s->return_to_default_attempted = r.return_to_default_attempted_;
  s->pose_state = r.pose_state_;
  put3(s->walk_plane, r.walk_plane_);
  put3(s->walk_plane_normal, r.walk_plane_normal_);
  putPose(s->odometry_ideal, r.odometry_ideal_);
  putPose(s->walk_plane_pose, r.walk_plane_pose_);
  putPose(s->origin_walk_plane_pose, r.origin_walk_plane_pose_);
  putPose(s->manual_pose, r.manual_pose_);
  putPose(s->imu_pose, r.imu_pose_);
  putPose(s->inclination_pose, r.inclination_pose_);
  putPose(s->auto_pose, r.auto_pose_);
  put3(s->rotation_absement_error, r.rotation_absement_error_);
  put3(s->rotation_position_error, r.rotation_position_error_);
  put3(s->rotation_velocity_error, r.rotation_velocity_error_);
  putPose(s->tip_align_pose, r.tip_align_pose_);
  putPose(s->origin_tip_align_pose, r.origin_tip_align_pose_);
  s->auto_posing_state = r.auto_posing_state_;
  s->pose_phase = r.pose_phase_;
  for (size_t k = 0; k < r.auto_posers_.size() && k < SHC_MAX_AUTO_POSERS; ++k) {
    const AutoPoser& ap = r.auto_posers_[k];
    s->auto_poser_flags[k] = (ap.start_check_ ? 1 : 0) | (ap.end_check_first ? 2 : 0) | (ap.end_check_second ? 4 : 0) |
                             (ap.allow_posing_ ? 8 : 0);
  }
  putPose(s->current_pose, r.current_pose_);
  for (int i = 0; i < r.leg_count_; ++i) {
    const Leg& leg = r.legs[i];
    const LegStepper& st = leg.stepper;
    shc_leg_state& o = s->legs[i];
    for (int j = 0; j < leg.joint_count_; ++j) {
      o.joint_position[j] = leg.joints[j + 1].desired_position_;
      o.joint_velocity[j] = leg.joints[j + 1].desired_velocity_;
    }
    put3(o.tip_position, st.current_tip_pose_.position_);
    put3(o.tip_velocity, st.current_tip_velocity_);
    put3(o.swing_origin_position, st.swing_origin_tip_position_);
    put3(o.swing_origin_velocity, st.swing_origin_tip_velocity_);
    put3(o.stance_origin_position, st.stance_origin_tip_position_);
    put3(o.default_tip_position, st.default_tip_pose_.position_);
    put3(o.target_tip_position, st.target_tip_pose_.position_);
    put3(o.stride_vector, st.stride_vector_);
    put3(o.walk_plane, st.walk_plane_);
    put3(o.walk_plane_normal, st.walk_plane_normal_);
    o.swing_progress = st.swing_progress_;
    o.stance_progress = st.stance_progress_;
    o.phase = st.phase_;
    o.step_state = st.step_state_;
    o.at_correct_phase = st.at_correct_phase_;
    o.completed_first_step = st.completed_first_step_;
    o.admittance_state[0] = leg.admittance_state_[0];
    o.admittance_state[1] = leg.admittance_state_[1];
    put3(o.admittance_delta, leg.admittance_delta_);
    put3(o.tip_force_calculated, leg.tip_force_calculated_);
    o.virtual_stiffness = leg.virtual_stiffness_;
    o.negate_auto_pose = leg.poser.negate_auto_pose_;
    putQuat(o.tip_rotation, st.current_tip_pose_.rotation_);
    putQuat(o.origin_tip_rotation, st.origin_tip_pose_.rotation_);
    putQuat(o.target_tip_rotation, st.target_tip_pose_.rotation_);
    o.step_plane_defined = leg.step_plane_pose_ != Pose::Undefined();
    if (o.step_plane_defined) put3(o.step_plane_position, leg.step_plane_pose_.position_);
    o.touchdown_detection = st.touchdown_detection_;
    putPose(o.external_target_pose, st.external_target_.pose_);
    putPose(o.external_target_transform, st.external_target_.transform_);
    o.external_target_clearance = st.external_target_.swing_clearance_;
    o.external_target_defined = st.external_target_.defined_;
    o.external_target_odom_frame = st.external_target_.odom_ideal_frame_;
    putPose(o.external_default_pose, st.external_default_.pose_);
    putPose(o.external_default_transform, st.external_default_.transform_);
    o.external_default_defined = st.external_default_.defined_;
    put3(o.model_tip_position, leg.current_tip_pose_.position_);
    put3(o.desired_tip_position, leg.desired_tip_pose_.position_);
    o.ik_result = leg.last_ik_result_;
  }
}

void importState(Robot& r, const shc_robot_state* s) {
  r.desired_linear_velocity_[0] = s->desired_linear_velocity[0];
  r.desired_linear_velocity_[1] = s->desired_linear_velocity[1];
  r.desired_angular_velocity_ = s->desired_angular_velocity;
  r.walk_state_ = WalkState(s->walk_state);
  r.legs_at_correct_phase_ = s->legs_at_correct_phase;
  r.legs_completed_first_step_ = s->legs_completed_first_step;
  r.return_to_default_attempted_ = s->return_to_default_attempted != 0;
  r.pose_state_ = PosingState(s->pose_state);
  r.walk_plane_ = get3(s->walk_plane);
  r.walk_plane_normal_ = get3(s->walk_plane_normal);
  r.odometry_ideal_ = getPose(s->odometry_ideal);
  r.walk_plane_pose_ = getPose(s->walk_plane_pose);
  r.origin_walk_plane_pose_ = getPose(s->origin_walk_plane_pose);
  r.manual_pose_ = getPose(s->manual_pose);
  r.imu_pose_ = getPose(s->imu_pose);
  r.inclination_pose_ = getPose(s->inclination_pose);
  r.auto_pose_ = getPose(s->auto_pose);
  r.rotation_absement_error_ = get3(s->rotation_absement_error);
  r.rotation_position_error_ = get3(s->rotation_position_error);
  r.rotation_velocity_error_ = get3(s->rotation_velocity_error);
  r.tip_align_pose_ = getPose(s->tip_align_pose);
  r.origin_tip_align_pose_ = getPose(s->origin_tip_align_pose);
  r.auto_posing_state_ = PosingState(s->auto_posing_state);
  r.pose_phase_ = s->pose_phase;
  for (size_t k = 0; k < r.auto_posers_.size() && k < SHC_MAX_AUTO_POSERS; ++k) {
    AutoPoser& ap = r.auto_posers_[k];
    ap.start_check_ = s->auto_poser_flags[k] & 1;
    ap.end_check_first = s->auto_poser_flags[k] & 2;
    ap.end_check_second = s->auto_poser_flags[k] & 4;
    ap.allow_posing_ = s->auto_poser_flags[k] & 8;
  }
  r.current_pose_ = getPose(s->current_pose);
  for (int i = 0; i < r.leg_count_; ++i) {
    Leg& leg = r.legs[i];
    LegStepper& st = leg.stepper;
    const shc_leg_state& o = s->legs[i];
    for (int j = 0; j < leg.joint_count_; ++j) {
      leg.joints[j + 1].desired_position_ = o.joint_position[j];
      leg.joints[j + 1].desired_velocity_ = o.joint_velocity[j];
    }
    leg.applyFK();  // current_tip_pose_ is a function of the joint positions
    st.current_tip_pose_.position_ = get3(o.tip_position);
    st.current_tip_velocity_ = get3(o.tip_velocity);
    st.swing_origin_tip_position_ = get3(o.swing_origin_position);
    st.swing_origin_tip_velocity_ = get3(o.swing_origin_velocity);
    st.stance_origin_tip_position_ = get3(o.stance_origin_position);
    st.default_tip_pose_.position_ = get3(o.default_tip_position);
    st.target_tip_pose_.position_ = get3(o.target_tip_position);
    st.stride_vector_ = get3(o.stride_vector);
    st.walk_plane_ = get3(o.walk_plane);
    st.walk_plane_normal_ = get3(o.walk_plane_normal);
    st.swing_progress_ = o.swing_progress;
    st.stance_progress_ = o.stance_progress;
    st.phase_ = o.phase;
    st.step_state_ = StepState(o.step_state);
    st.at_correct_phase_ = o.at_correct_phase != 0;
    st.completed_first_step_ = o.completed_first_step != 0;
    leg.admittance_state_[0] = o.admittance_state[0];
    leg.admittance_state_[1] = o.admittance_state[1];
    leg.admittance_delta_ = get3(o.admittance_delta);
    leg.tip_force_calculated_ = get3(o.tip_force_calculated);
    leg.virtual_stiffness_ = o.virtual_stiffness;
    leg.poser.negate_auto_pose_ = o.negate_auto_pose != 0;
    st.touchdown_detection_ = o.touchdown_detection != 0;
    st.external_target_.pose_ = getPose(o.external_target_pose);
    st.external_target_.transform_ = getPose(o.external_target_transform);
    st.external_target_.swing_clearance_ = o.external_target_clearance;
    st.external_target_.defined_ = o.external_target_defined != 0;
    st.external_target_.odom_ideal_frame_ = o.external_target_odom_frame != 0;
    st.external_default_.pose_ = getPose(o.external_default_pose);
    st.external_default_.transform_ = getPose(o.external_default_transform);
    st.external_default_.defined_ = o.external_default_defined != 0;
    // (the rotation of the step plane pose is never read: walk_controller.cpp:1085-1110 uses its position and whether it is defined)
    leg.step_plane_pose_ = o.step_plane_defined ? Pose(get3(o.step_plane_position), leg.current_tip_pose_.rotation_) : Pose::Undefined();
    // target_tip_pose_.rotation_ is a constant of the configuration (no rough-terrain targets): left as constructed
    st.current_tip_pose_.rotation_ = Quat(o.tip_rotation[0], o.tip_rotation[1], o.tip_rotation[2], o.tip_rotation[3]);
    st.origin_tip_pose_.rotation_ = Quat(o.origin_tip_rotation[0], o.origin_tip_rotation[1], o.origin_tip_rotation[2], o.origin_tip_rotation[3]);
  }
}

void fillStartup(const Robot& r, shc_startup* out) {
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < r.leg_count_; ++i) {
    for (int j = 0; j < r.legs[i].joint_count_; ++j) out->default_joint[i][j] = r.legs[i].joints[j + 1].default_position_;
    if (!r.legs[i].workspace_.empty()) {
      const LimitMap& wp = r.legs[i].workspace_.begin()->second;
      for (int b = 0; b < SHC_N_BEARINGS; ++b) out->workspace[i][b] = wp.count(b * 45) ? wp.at(b * 45) : 0.0;
    }
    out->phase_offsets[i] = r.legs[i].stepper.phase_offset_;
  }
  for (int b = 0; b < SHC_N_BEARINGS; ++b) {
    int k = b * 45;
    out->walkspace[b] = r.walkspace_.count(k) ? r.walkspace_.at(k) : 0.0;
    out->max_linear_speed[b] = r.max_linear_speed_.count(k) ? r.max_linear_speed_.at(k) : 0.0;
    out->max_angular_speed[b] = r.max_angular_speed_.count(k) ? r.max_angular_speed_.at(k) : 0.0;
    out->max_linear_acceleration[b] = r.max_linear_acceleration_.count(k) ? r.max_linear_acceleration_.at(k) : 0.0;
    out->max_angular_acceleration[b] = r.max_angular_acceleration_.count(k) ? r.max_angular_acceleration_.at(k) : 0.0;
  }
  out->step_frequency = r.step_.frequency_;
  out->period = r.step_.period_;
  out->swing_period = r.step_.swing_period_;
  out->stance_period = r.step_.stance_period_;
  out->stance_end = r.step_.stance_end_;
  out->swing_start = r.step_.swing_start_;
  out->swing_end = r.step_.swing_end_;
  out->stance_start = r.step_.stance_start_;
  out->pose_phase_length = r.pose_phase_length_;
  out->pose_normaliser = r.normaliser_;
  out->auto_pose_reference_leg = r.auto_pose_reference_leg_;
}

struct Batch {
  shc_config cfg;
  std::vector<Robot*> robots;
  int startup_loops = 0;
  // harness state of shc_oracle_batch_sequence_step: 1 = the robot's start-up sequence has completed, 2 = its shut-down
  // (StateController stops calling executeSequence once it has returned 100, state_controller.cpp:314-350)
  std::vector<int> sequence_done;
  std::vector<double> step_planes;  // latched TipState.step_plane readings [n][L][3] (shc_oracle_batch_set_step_planes)
};

void stepOne(Robot& r, const double* cmd, const double* imu, const double* tip_force, const double* manual,
             const double* step_plane = nullptr) {
  // Inputs arrive through callbacks in ros::spinOnce() before the next loop() (main.cpp:130).
  if (imu) r.setImuData(Quat(imu[0], imu[1], imu[2], imu[3]), Vec3(imu[7], imu[8], imu[9]), Vec3(imu[4], imu[5], imu[6]));
  if (tip_force)  // tipStatesCallback with wrench values (state_controller.cpp:1636-1648)
    for (int l = 0; l < r.leg_count_; ++l) {
      r.legs[l].stepper.touchdown_detection_ = true;
      r.legs[l].tip_force_measured_ = Vec3(tip_force[l * 3], tip_force[l * 3 + 1], tip_force[l * 3 + 2]);
      r.legs[l].touchdownDetection();
    }
  if (step_plane)  // tipStatesCallback with step-plane values (state_controller.cpp:1650-1675)
    for (int l = 0; l < r.leg_count_; ++l) {
      Leg& leg = r.legs[l];
      leg.stepper.touchdown_detection_ = true;
      const double* sp = step_plane + 3 * l;
      if (sp[2] != UNASSIGNED_VALUE) {
        Vec3 step_plane_position(sp[2], 0.0, 0.0);
        Vec3 step_plane_normal(sp[0], sp[1], -1.0);
        Quat step_plane_orientation = fromTwoVectors(Vec3(0, 0, 1.0), -step_plane_normal);
        leg.step_plane_pose_ = leg.tipPoseRobotFrame(Pose(step_plane_position, step_plane_orientation));
      } else {
        leg.step_plane_pose_ = Pose::Undefined();
      }
    }
  if (manual) {  // poser_->setManualPoseInput (state_controller.cpp:1148)
    r.translation_velocity_input_ = Vec3(manual[0], manual[1], manual[2]);
    r.rotation_velocity_input_ = Vec3(manual[3], manual[4], manual[5]);
  }
  r.setBodyVelocityInput(cmd[0], cmd[1], cmd[2]);
  r.loop();
}

}  // namespace

extern "C" {

// One robot is taken through StateController::init + the direct start-up (PACKED -> READY) once, then cloned n times
// and put in RUNNING state; engine cycle 0 corresponds to the first full loop() in RUNNING state.
void* shc_oracle_batch_create(const shc_config* cfg, int n_robots) {
  Batch* b = new Batch();
  b->cfg = *cfg;
  Robot proto(*cfg);
  proto.stateInit();
  int loops = 0;
  while (proto.robot_state_ != READY && loops < 100000) {
    proto.requestRobotState(RUNNING);
    proto.loop();
    ++loops;
  }
  b->startup_loops = loops;
  proto.robot_state_ = RUNNING;
  proto.new_robot_state_ = RUNNING;
  proto.transition_state_flag_ = false;
  for (int i = 0; i < n_robots; ++i) b->robots.push_back(cloneRobot(proto));
  return b;
}

void shc_oracle_batch_destroy(void* h) {
  Batch* b = static_cast<Batch*>(h);
  for (Robot* r : b->robots) delete r;
  delete b;
}

int shc_oracle_batch_size(void* h) { return int(static_cast<Batch*>(h)->robots.size()); }
int shc_oracle_startup_loops(void* h) { return static_cast<Batch*>(h)->startup_loops; }

void shc_oracle_get_startup(void* h, shc_startup* out) {
  fillStartup(*static_cast<Batch*>(h)->robots[0], out);
  out->startup_loops = static_cast<Batch*>(h)->startup_loops;
}

// cmd [n][3]; imu [n][10] (quat wxyz, gyro xyz, accel xyz) or NULL; tip_force [n][L][3] or NULL; manual [n][6] or NULL.
void shc_oracle_batch_step(void* h, const double* cmd, const double* imu, const double* tip_force, const double* manual,
                           int n_threads) {
  Batch* b = static_cast<Batch*>(h);
  const int n = int(b->robots.size());
  const int L = b->cfg.leg_count;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i)
      stepOne(*b->robots[i], cmd + 3 * i, imu ? imu + 10 * i : nullptr, tip_force ? tip_force + 3 * L * i : nullptr,
              manual ? manual + 6 * i : nullptr, b->step_planes.empty() ? nullptr : b->step_planes.data() + (size_t)3 * L * i);
  };
  if (n_threads <= 1 || n < 2 * n_threads) {
    work(0, n);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(work, int((long long)n * t / n_threads), int((long long)n * (t + 1) / n_threads));
  for (auto& t : th) t.join();
}

// Runs `cycles` control cycles with a constant command per robot and returns wall seconds (CPU baseline timing).
// Robots are partitioned over n_threads std::threads; each thread runs all cycles of its own robots (robots are
// independent, so no per-cycle synchronisation is needed).
double shc_oracle_batch_run(void* h, const double* cmd, int cycles, int n_threads) {
  Batch* b = static_cast<Batch*>(h);
  const int n = int(b->robots.size());
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i)
      for (int c = 0; c < cycles; ++c) stepOne(*b->robots[i], cmd + 3 * i, nullptr, nullptr, nullptr);
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, int((long long)n * t / n_threads), int((long long)n * (t + 1) / n_threads));
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Runs `cycles` control cycles with PER-CYCLE commands cmd_seq [cycles][n][3] (pre-generated by the caller, so that the
// timed region holds nothing but the control cycles) and returns wall seconds.  Robots are partitioned over n_threads
// std::threads once; each thread runs all cycles of its own robots.
double shc_oracle_batch_run_seq(void* h, const double* cmd_seq, int cycles, int n_threads) {
  Batch* b = static_cast<Batch*>(h);
  const int n = int(b->robots.size());
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i)
      for (int c = 0; c < cycles; ++c) stepOne(*b->robots[i], cmd_seq + ((size_t)c * n + i) * 3, nullptr, nullptr, nullptr);
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, int((long long)n * t / n_threads), int((long long)n * (t + 1) / n_threads));
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// One loop() of a stepping / joint-space sequence for every robot: kind 0 = stepToNewStance
// (pose_controller.cpp:520), 1 = packLegs(time) (:597), 2 = unpackLegs(time) (:661), 3 / 4 = executeSequence(START_UP / SHUT_DOWN) (:145).  progress_out [n] gets each robot's return value.
void shc_oracle_batch_sequence_step(void* h, int kind, double time, int* progress_out) {
  Batch* b = static_cast<Batch*>(h);
  b->sequence_done.resize(b->robots.size(), 0);
  for (size_t i = 0; i < b->robots.size(); ++i) {
    Robot& r = *b->robots[i];
    int p;
    if (kind == 3 || kind == 4) {
      if (b->sequence_done[i] == kind - 2) {
        p = PROGRESS_COMPLETE;  // already through this sequence: its state machine no longer calls executeSequence
      } else {
        b->sequence_done[i] = 0;
        p = r.executeSequence(kind == 3 ? START_UP : SHUT_DOWN);
        if (p == PROGRESS_COMPLETE) b->sequence_done[i] = kind - 2;
        if (r.sequence_failed_) p = -2;
      }
    } else {
      p = kind == 0 ? r.stepToNewStance() : kind == 1 ? r.packLegs(time) : r.unpackLegs(time);
    }
    if (progress_out) progress_out[i] = p;
  }
}

void shc_oracle_batch_get_joints(void* h, double* out) {  // [n][L][D]
  Batch* b = static_cast<Batch*>(h);
  const int L = b->cfg.leg_count, D = b->cfg.joint_count;
  for (size_t i = 0; i < b->robots.size(); ++i)
    for (int l = 0; l < L; ++l)
      for (int j = 0; j < D; ++j) out[(i * L + l) * D + j] = b->robots[i]->legs[l].joints[j + 1].desired_position_;
}

void shc_oracle_batch_get_state(void* h, shc_robot_state* out) {
  Batch* b = static_cast<Batch*>(h);
  for (size_t i = 0; i < b->robots.size(); ++i) exportState(*b->robots[i], out + i);
}

void shc_oracle_batch_set_state(void* h, const shc_robot_state* in) {
  Batch* b = static_cast<Batch*>(h);
  for (size_t i = 0; i < b->robots.size(); ++i) importState(*b->robots[i], in + i);
}

// poser_->setPoseResetMode (state_controller.cpp:1199) for every robot of the batch.
void shc_oracle_batch_set_pose_reset_mode(void* h, int mode) {
  Batch* b = static_cast<Batch*>(h);
  for (Robot* r : b->robots) r->pose_reset_mode_ = PoseResetMode(mode);
}

// jointStatesCallback (state_controller.cpp:1565-1590): measured joint efforts [n][L][D] (NULL = zero), read by
// Leg::calculateTipForce (model.cpp:667) when use_joint_effort is set.
// Tip range-sensor readings [n][L][3] (x, y, z; z = UNASSIGNED_VALUE for "no reading"), latched; NULL = none.
void shc_oracle_batch_set_step_planes(void* h, const double* sp) {
  Batch* b = static_cast<Batch*>(h);
  if (!sp) { b->step_planes.clear(); return; }
  b->step_planes.assign(sp, sp + b->robots.size() * b->cfg.leg_count * 3);
}

void shc_oracle_batch_set_joint_efforts(void* h, const double* eff) {
  Batch* b = static_cast<Batch*>(h);
  const int L = b->cfg.leg_count, D = b->cfg.joint_count;
  for (size_t i = 0; i < b->robots.size(); ++i)
    for (int l = 0; l < L; ++l)
      for (int j = 0; j < D; ++j) b->robots[i]->legs[l].joints[j + 1].current_effort_ = eff ? eff[(i * L + l) * D + j] : 0.0;
}

int shc_oracle_state_record_size(void) { return int(sizeof(shc_robot_state)); }
int shc_oracle_config_size(void) { return int(sizeof(shc_config)); }
int shc_oracle_startup_size(void) { return int(sizeof(shc_startup)); }

// ---- unit-test hooks ------------------------------------------------------------------------------------------
int shc_oracle_mod(int a, int b) { return om::mod(a, b); }
int shc_oracle_round_to_int(double x) { return om::roundToInt(x); }
int shc_oracle_round_to_even_int(double x) { return om::roundToEvenInt(x); }
double shc_oracle_smooth_step(double c) { return om::smoothStep(c); }
void shc_oracle_quat_to_euler(const double q[4], int intrinsic, double out[3]) {
  Vec3 e = om::quaternionToEulerAngles(Quat(q[0], q[1], q[2], q[3]), intrinsic != 0);
  put3(out, e);
}
void shc_oracle_euler_to_quat(const double e[3], int intrinsic, double out[4]) {
  Quat q = om::eulerAnglesToQuaternion(Vec3(e[0], e[1], e[2]), intrinsic != 0);
  out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
}
void shc_oracle_dh(double d, double theta, double r, double alpha, double out[16]) {
  Mat4 m = om::createDHMatrix(d, theta, r, alpha);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = m.m[i][j];
}
void shc_oracle_quartic_bezier(const double nodes[15], double t, double out[3], double out_dot[3]) {
  Vec3 n[5];
  for (int i = 0; i < 5; ++i) n[i] = get3(nodes + 3 * i);
  put3(out, om::quarticBezier(n, t));
  put3(out_dot, om::quarticBezierDot(n, t));
}
void shc_oracle_from_two_vectors(const double a[3], const double b[3], double out[4]) {
  Quat q = om::fromTwoVectors(get3(a), get3(b));
  out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
}
void shc_oracle_slerp(const double a[4], double t, const double b[4], double out[4]) {
  Quat q = om::slerp(Quat(a[0], a[1], a[2], a[3]), t, Quat(b[0], b[1], b[2], b[3]));
  out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
}
void shc_oracle_pose_ops(const double a[7], const double b[7], double add_out[7], double remove_out[7], double inv_out[7]) {
  Pose A = getPose(a), B = getPose(b);
  putPose(add_out, A.addPose(B));
  putPose(remove_out, A.removePose(B));
  putPose(inv_out, ~A);
}
// Forward kinematics of one leg at joint angles q (model.cpp:945): tip position in the base_link frame.
void shc_oracle_fk(const shc_config* cfg, int leg, const double* q, double out_tip[3]) {
  Robot r(*cfg);
  Leg& l = r.legs[leg];
  for (int j = 0; j < l.joint_count_; ++j) l.joints[j + 1].desired_position_ = q[j];
  Pose p = l.applyFK();
  put3(out_tip, p.position_);
}
// One Leg::solveIK call (model.cpp:726) at joint state (q, qd) for a leg-frame position delta.
void shc_oracle_solve_ik(const shc_config* cfg, int leg, const double* q, const double* qd, const double delta[3],
                         double* out_dq) {
  Robot r(*cfg);
  Leg& l = r.legs[leg];
  for (int j = 0; j < l.joint_count_; ++j) {
    l.joints[j + 1].desired_position_ = q[j];
    l.joints[j + 1].desired_velocity_ = qd[j];
  }
  l.applyFK();
  double d6[6] = {delta[0], delta[1], delta[2], 0, 0, 0};
  l.solveIK(d6, false, out_dq);
}
// generateStepCycle only (walk_controller.cpp:365) + phase offsets (walk_controller.cpp:237-278).
void shc_oracle_step_cycle(const shc_config* cfg, shc_startup* out) {
  Robot r(*cfg);
  r.walkerInit();
  r.generateLimits();  // walkspace empty: only sets the phase offsets
  fillStartup(r, out);
}
// One admittance update for one leg from state x with measured force f (admittance_controller.cpp:22).
void shc_oracle_admittance(const shc_config* cfg, const double x_in[2], const double force[3], double x_out[2]) {
  Robot r(*cfg);
  r.legs[0].admittance_state_[0] = x_in[0];
  r.legs[0].admittance_state_[1] = x_in[1];
  // undo the force gain so that `force` is the value seen by the integrator
  r.params_.force_gain = 1.0;
  r.legs[0].tip_force_measured_ = get3(force);
  r.params_.use_joint_effort = 0;
  r.leg_count_ = 1;
  r.legs[0].current_tip_pose_ = Pose::Identity();
  r.updateAdmittance();
  x_out[0] = r.legs[0].admittance_state_[0];
  x_out[1] = r.legs[0].admittance_state_[1];
}

}  // extern "C"

extern "C" {
// One Leg::applyIK (model.cpp:861) from joint state (q, qd) toward `desired` (base_link frame).
double shc_oracle_apply_ik(const shc_config* cfg, int leg, double* q, double* qd, const double desired[3], int simulation,
                           double tip_out[3]) {
  Robot r(*cfg);
  Leg& l = r.legs[leg];
  for (int j = 0; j < l.joint_count_; ++j) {
    l.joints[j + 1].desired_position_ = q[j];
    l.joints[j + 1].desired_velocity_ = qd[j];
  }
  l.applyFK();
  l.setDesiredTipPose(Pose(get3(desired), UndefinedRotation()));
  double res = l.applyIK(simulation != 0);
  for (int j = 0; j < l.joint_count_; ++j) {
    q[j] = l.joints[j + 1].desired_position_;
    qd[j] = l.joints[j + 1].desired_velocity_;
  }
  put3(tip_out, l.current_tip_pose_.position_);
  return res;
}
}

extern "C" {
// Leg::generateWorkspace (model.cpp:309-510) of leg `leg` for a robot that has gone through the direct start-up: the
// simple (full = 0) or the layered rough-terrain workspace (full = 1), planes in ascending height: heights [max_planes],
// radii [max_planes][9].  Returns the number of planes.
int shc_oracle_workspace(const shc_config* cfg, int leg, int full, int max_planes, double* heights, double* radii) {
  Robot r(*cfg);
  r.stateInit();
  int loops = 0;
  while (r.robot_state_ != READY && loops < 100000) {
    r.requestRobotState(RUNNING);
    r.loop();
    ++loops;
  }
  const int saved = r.params_.rough_terrain_mode;
  r.params_.rough_terrain_mode = full ? 1 : 0;  // generateWorkspace: simple_workspace = !rough_terrain_mode (model.cpp:313)
  Leg search_leg = r.legs[leg];                 // as Model::generateWorkspaces (model.cpp:120) does
  search_leg.workspace_.clear();
  search_leg.stepper.leg_ = &search_leg;
  search_leg.poser.leg_ = &search_leg;
  search_leg.init(true);
  Workspace ws = search_leg.generateWorkspace();
  r.params_.rough_terrain_mode = saved;
  int p = 0;
  for (auto& kv : ws) {
    if (p < max_planes) {
      heights[p] = kv.first;
      for (int b = 0; b < SHC_N_BEARINGS; ++b) radii[p * SHC_N_BEARINGS + b] = kv.second.count(b * 45) ? kv.second.at(b * 45) : 0.0;
    }
    ++p;
  }
  return p;
}

// The joint commands of every loop() of the direct start-up (state_controller.cpp:254-281, pose_controller.cpp:463) of a
// robot whose joints are at q_init [L][D] when it begins (NULL: the default joint positions): out [max_loops][L][D], one row
// per loop() from the first one that moves the joints.  Returns the number of rows.
int shc_oracle_startup_trajectory(const shc_config* cfg, const double* q_init, int max_loops, double* out) {
  Robot r(*cfg);
  r.stateInit();
  const int L = r.leg_count_, D = cfg->joint_count;
  if (q_init) {  // jointStatesCallback + initModel(false) (state_controller.cpp:1565-1590, model.cpp:286)
    for (int l = 0; l < L; ++l) {
      for (int j = 0; j < D; ++j) {
        Joint& jt = r.legs[l].joints[j + 1];
        jt.current_position_ = q_init[l * D + j];
        jt.desired_position_ = jt.current_position_;
        jt.prev_desired_position_ = jt.desired_position_;
      }
      r.legs[l].applyFK();
      r.legs[l].desired_tip_pose_ = r.legs[l].current_tip_pose_;
    }
  }
  int rows = 0, loops = 0;
  while (r.robot_state_ != READY && loops < 100000) {
    r.requestRobotState(RUNNING);
    const bool moving = r.robot_state_ == PACKED;  // the first loop() only leaves UNKNOWN
    r.loop();
    ++loops;
    if (moving && rows < max_loops) {
      for (int l = 0; l < L; ++l)
        for (int j = 0; j < D; ++j) out[((size_t)rows * L + l) * D + j] = r.legs[l].joints[j + 1].desired_position_;
      ++rows;
    }
  }
  return rows;
}
}

extern "C" {
// The reference's publishers (state_controller.cpp:777-1047) for robot `index` of the batch, restated on the oracle's robot:
// publishDesiredJointState, publishLegState (msg/LegState.msg), publishVelocity / Pose / RotationPoseError and
// publishFrameTransforms, into the records of include/shc_msgs.h.  measured [L][D] = measured joint positions or NULL.
void shc_oracle_batch_get_messages(void* h, int index, const double* measured, shc_joint_state_msg* js, shc_leg_state_msg* legs,
                                   shc_body_msg* body) {
  Batch* b = static_cast<Batch*>(h);
  Robot& r = *b->robots[index];
  const int L = r.leg_count_, D = b->cfg.joint_count;
  std::memset(js, 0, sizeof(*js));
  for (int l = 0; l < L; ++l) {
    Leg& leg = r.legs[l];
    shc_leg_state_msg& m = legs[l];
    std::memset(&m, 0, sizeof(m));
    putPose(m.walker_tip_pose, leg.stepper.current_tip_pose_);
    putPose(m.target_tip_pose, leg.stepper.target_tip_pose_);
    putPose(m.poser_tip_pose, leg.poser.current_tip_pose_);
    putPose(m.model_tip_pose, leg.current_tip_pose_);
    if (measured)
      for (int j = 0; j < D; ++j) leg.joints[j + 1].current_position_ = measured[l * D + j];
    else
      for (int j = 0; j < D; ++j) leg.joints[j + 1].current_position_ = leg.joints[j + 1].desired_position_;
    putPose(m.actual_tip_pose, leg.applyFK(false, true));
    // state_controller.cpp:842 restores the transforms of the desired positions with a plain applyFK(): set_current is
    // true, so Leg::current_tip_velocity_ becomes (tip - tip) / time_delta = 0 BEFORE :846-848 read it — the reference
    // publishes a zero model_tip_velocity (found by running its own publishers, tests/test_reference_pin.py).
    leg.applyFK();
    put3(m.model_tip_velocity, leg.current_tip_velocity_);
    for (int j = 0; j < D; ++j) {
      const Joint& jt = leg.joints[j + 1];
      m.joint_positions[j] = jt.desired_position_;
      m.joint_velocities[j] = jt.desired_velocity_;
      m.joint_efforts[j] = jt.desired_effort_;
      js->position[l * D + j] = jt.desired_position_;   // Leg::generateDesiredJointStateMsg (model.cpp:605)
      js->velocity[l * D + j] = jt.desired_velocity_;
      js->effort[l * D + j] = jt.desired_effort_;
      js->position_command[l * D + j] = jt.desired_position_ + jt.offset_;  // :795
      Pose jp = leg.jointPoseRobotFrame(j + 1);
      Pose turned(jp.position_, jp.rotation_ * quatFromAngleAxis(jt.desired_position_, UnitZ()));  // :1030
      putPose(m.joint_transform[j], turned);
    }
    putPose(m.tip_transform, leg.tipPoseRobotFrame());
    m.swing_progress = leg.stepper.swing_progress_;
    m.stance_progress = leg.stepper.stance_progress_;
    const StepCycle& step = r.step_;
    double swing_time = (double(step.swing_period_) / step.period_) / step.frequency_;
    double stance_time = (double(step.stance_period_) / step.period_) / step.frequency_;
    double time_to_swing_end;
    if (leg.stepper.stance_progress_ >= 0.0) time_to_swing_end = stance_time * (1.0 - leg.stepper.stance_progress_) + swing_time;
    else time_to_swing_end = swing_time * (1.0 - leg.stepper.swing_progress_);
    m.time_to_swing_end = time_to_swing_end;
    putPose(m.pose_delta, r.calculateOdometry(time_to_swing_end));
    putPose(m.auto_pose, leg.poser.auto_pose_);
    m.tip_force[0] = leg.tip_force_calculated_[0] * r.params_.force_gain;
    m.tip_force[1] = leg.tip_force_calculated_[1] * r.params_.force_gain;
    m.tip_force[2] = leg.tip_force_calculated_[2] * r.params_.force_gain;
    put3(m.admittance_delta, leg.admittance_delta_);
    m.virtual_stiffness = leg.virtual_stiffness_;
  }
  std::memset(body, 0, sizeof(*body));
  body->velocity[0] = r.desired_linear_velocity_[0];
  body->velocity[1] = r.desired_linear_velocity_[1];
  body->velocity[5] = r.desired_angular_velocity_;
  put3(body->pose, r.current_pose_.position_);
  Vec3 e = om::quaternionToEulerAngles(r.current_pose_.rotation_, false);
  put3(body->pose + 3, e);
  put3(body->rotation_pose_error, r.rotation_absement_error_);
  put3(body->rotation_pose_error + 3, r.rotation_position_error_);
  put3(body->rotation_pose_error + 6, r.rotation_velocity_error_);
  putPose(body->odom_ideal_to_base_link, r.odometry_ideal_.addPose(r.current_pose_));
  putPose(body->base_link_to_walk_plane, ~r.current_pose_);
}
}

