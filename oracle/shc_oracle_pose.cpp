// shc_oracle_pose.cpp — TEST INFRASTRUCTURE ONLY.  Restates the per-cycle and direct-start-up parts of
// /root/reference/src/pose_controller.cpp, all of src/admittance_controller.cpp and the StateController
// init/loop/transitionRobotState/runningState call order (src/state_controller.cpp:127-447) in IEEE double.
// PINNED to the reference's own code (oracle/_ref, tests/test_reference_pin.py; see shc_oracle.hpp).
#include "shc_oracle.hpp"

namespace shc_oracle {

// ---------------------------------------------------------------------------------------------------------------
// PoseController
// ---------------------------------------------------------------------------------------------------------------
void Robot::poserInit() {  // pose_controller.cpp:30
  for (int i = 0; i < leg_count_; ++i) {
    legs[i].poser = LegPoser();
    legs[i].poser.robot = this;
    legs[i].poser.leg_ = &legs[i];
  }
  setAutoPoseParams();
  walk_plane_pose_.position_ = Vec3(0.0, 0.0, params_.body_clearance);
  origin_walk_plane_pose_ = walk_plane_pose_;
}

void Robot::setAutoPoseParams() {  // pose_controller.cpp:44
  double raw_phase_length;
  int base_phase_length;
  pose_frequency_ = params_.pose_frequency;
  if (pose_frequency_ == -1.0) {
    base_phase_length = params_.stance_phase + params_.swing_phase;
    double swing_ratio = double(params_.swing_phase) / base_phase_length;
    raw_phase_length = ((1.0 / params_.step_frequency) / params_.time_delta) / swing_ratio;
  } else {
    base_phase_length = params_.pose_phase_length;
    raw_phase_length = ((1.0 / pose_frequency_) / params_.time_delta);
  }
  pose_phase_length_ = roundToEvenInt(raw_phase_length / base_phase_length) * base_phase_length;
  normaliser_ = pose_phase_length_ / base_phase_length;

  for (int i = 0; i < leg_count_; ++i) {
    LegPoser& leg_poser = legs[i].poser;
    leg_poser.pose_negation_phase_start_ = params_.pose_negation_phase_starts[i];
    leg_poser.pose_negation_phase_end_ = params_.pose_negation_phase_ends[i];
    leg_poser.negation_transition_ratio_ = params_.negation_transition_ratio[i];
    if (params_.offset_multiplier[i] == 0) auto_pose_reference_leg_ = i;
  }
  auto_posers_.clear();
  for (int i = 0; i < params_.auto_poser_count; ++i) {
    AutoPoser ap;
    ap.robot = this;
    ap.id_number_ = i;
    ap.start_phase_ = params_.pose_phase_starts[i];
    ap.end_phase_ = params_.pose_phase_ends[i];
    ap.x_amplitude_ = params_.x_amplitudes[i];
    ap.y_amplitude_ = params_.y_amplitudes[i];
    ap.z_amplitude_ = params_.z_amplitudes[i];
    ap.gravity_amplitude_ = params_.gravity_amplitudes[i];
    ap.roll_amplitude_ = params_.roll_amplitudes[i];
    ap.pitch_amplitude_ = params_.pitch_amplitudes[i];
    ap.yaw_amplitude_ = params_.yaw_amplitudes[i];
    ap.start_check_ = false;  // resetChecks (pose_controller.h:387)
    ap.end_check_first = ap.end_check_second = false;
    auto_posers_.push_back(ap);
  }
}

void Robot::updateStance() {  // pose_controller.cpp:110
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    LegStepper& leg_stepper = leg.stepper;
    LegPoser& leg_poser = leg.poser;
    Pose current_pose = current_pose_;
    LegState leg_state = leg.leg_state_;
    if (leg_state == WALKING || leg_state == MANUAL_TO_WALKING) {
      current_pose = current_pose.removePose(auto_pose_);
      current_pose = current_pose.addPose(leg_poser.auto_pose_);
      Vec3 new_tip_position = current_pose.inverseTransformVector(leg_stepper.current_tip_pose_.position_);
      Quat new_tip_rotation = current_pose.rotation_.inverse() * leg_stepper.current_tip_pose_.rotation_;
      leg_poser.current_tip_pose_ = Pose(new_tip_position, new_tip_rotation);
    } else if (leg_state == MANUAL || leg_state == WALKING_TO_MANUAL) {
      leg_poser.current_tip_pose_ = leg_stepper.current_tip_pose_;
    }
  }
}

int Robot::directStartup() {  // pose_controller.cpp:463
  int progress = 0;
  double time_to_start = params_.time_to_start;
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    LegPoser& leg_poser = leg.poser;
    LegStepper& leg_stepper = leg.stepper;
    if (!executing_transition_) {
      Leg test_leg = leg;  // Leg copy-ctor + generate(leg) (model.cpp:192-282)
      test_leg.workspace_.clear();
      test_leg.stepper.leg_ = &test_leg;
      test_leg.poser.leg_ = &test_leg;
      test_leg.init(true);
      Pose default_tip_pose = leg_stepper.default_tip_pose_;
      while (progress != PROGRESS_COMPLETE) {
        LegPoser& test_leg_poser = test_leg.poser;
        progress = test_leg_poser.stepToPosition(default_tip_pose, current_pose_, 0.0, time_to_start);
        test_leg.setDesiredTipPose(test_leg_poser.current_tip_pose_, true);
        test_leg.applyIK(true);
      }
      std::vector<double> default_configuration(test_leg.joint_count_, UNASSIGNED_VALUE);
      for (int j = 1; j <= test_leg.joint_count_; ++j) default_configuration[j - 1] = test_leg.joints[j].desired_position_;
      leg_poser.desired_configuration_ = default_configuration;
    }
    progress = leg_poser.transitionConfiguration(time_to_start);
  }
  executing_transition_ = (progress != 0 && progress != PROGRESS_COMPLETE);
  return progress;
}

bool Robot::legsBearingLoad() {  // model.cpp:78
  double body_height_estimate = 0.0;
  for (int i = 0; i < leg_count_; ++i) body_height_estimate += legs[i].current_tip_pose_.position_[2];
  return -(body_height_estimate / leg_count_) > HALF_BODY_DEPTH;
}

int Robot::executeSequence(SequenceSelection sequence) {  // pose_controller.cpp:145
  if (reset_transition_sequence_ && sequence == START_UP) {
    reset_transition_sequence_ = false;
    first_sequence_execution_ = true;
    transition_step_ = 0;
    for (int i = 0; i < leg_count_; ++i) {
      Leg& leg = legs[i];
      leg.poser.transition_poses_.clear();
      leg.poser.transition_poses_.push_back(leg.current_tip_pose_);
    }
  }

  int progress = 0;
  int normalised_progress = 0;
  int next_transition_step = 0, transition_step_target = 0, total_progress = 0;
  bool execute_horizontal_transition = false, execute_vertical_transition = false;
  if (sequence == START_UP) {
    execute_horizontal_transition = !(transition_step_ % 2);
    execute_vertical_transition = transition_step_ % 2;
    next_transition_step = transition_step_ + 1;
    transition_step_target = transition_step_count_;
    total_progress = transition_step_ * 100 / std::max(transition_step_count_, 1);
  } else if (sequence == SHUT_DOWN) {
    execute_horizontal_transition = transition_step_ % 2;
    execute_vertical_transition = !(transition_step_ % 2);
    next_transition_step = transition_step_ - 1;
    transition_step_target = 0;
    total_progress = 100 - transition_step_ * 100 / std::max(transition_step_count_, 1);
  }

  bool final_transition;
  bool sequence_complete = false;
  if (first_sequence_execution_) final_transition = (horizontal_transition_complete_ || vertical_transition_complete_);
  else final_transition = (next_transition_step == transition_step_target);

  double safety_factor = (first_sequence_execution_ ? SAFETY_FACTOR / (transition_step_ + 1) : 0.0);

  if (execute_horizontal_transition) {
    if (set_target_) {
      set_target_ = false;
      for (int i = 0; i < leg_count_; ++i) {
        Leg& leg = legs[i];
        LegStepper& leg_stepper = leg.stepper;
        LegPoser& leg_poser = leg.poser;
        leg_poser.leg_completed_step_ = false;
        Vec3 target_tip_position;
        if (int(leg_poser.transition_poses_.size()) > next_transition_step && next_transition_step >= 0) {
          target_tip_position = leg_poser.transition_poses_[next_transition_step].position_;
        } else {
          Vec3 default_tip_position = leg_stepper.default_tip_pose_.position_;
          target_tip_position = current_pose_.inverseTransformVector(default_tip_position);
        }
        target_tip_position[2] = leg.current_tip_pose_.position_[2];
        Quat target_tip_rotation = leg_stepper.target_tip_pose_.rotation_;
        leg_poser.target_tip_pose_ = Pose(target_tip_position, target_tip_rotation);
      }
    }

    bool direct_step = !legsBearingLoad();
    for (int i = 0; i < leg_count_; ++i) {
      Leg& leg = legs[i];
      LegPoser& leg_poser = leg.poser;
      if (!leg_poser.leg_completed_step_) {
        if (leg.group_ == current_group_ || direct_step) {
          Pose target_tip_pose = leg_poser.target_tip_pose_;
          bool apply_delta = (sequence == START_UP && final_transition);
          double step_height = direct_step ? 0.0 : params_.swing_height;
          double time_to_step = HORIZONTAL_TRANSITION_TIME / params_.step_frequency;
          time_to_step *= (first_sequence_execution_ ? 2.0 : 1.0);
          progress = leg_poser.stepToPosition(target_tip_pose, Pose::Identity(), step_height, time_to_step, apply_delta);
          leg.setDesiredTipPose(leg_poser.current_tip_pose_);
          double limit_proximity = leg.applyIK();
          bool exceeded_workspace = limit_proximity < safety_factor;
          if (first_sequence_execution_ && exceeded_workspace) {
            leg_poser.target_tip_pose_ = leg_poser.current_tip_pose_;
            progress = leg_poser.resetStepToPosition();
            proximity_alert_ = true;
          }
          if (progress == PROGRESS_COMPLETE) {
            leg_poser.leg_completed_step_ = true;
            legs_completed_step_++;
            if (first_sequence_execution_) {
              bool reached_target = !exceeded_workspace;
              Pose target_tip_pose2 = leg_poser.target_tip_pose_;
              Pose current_tip_pose = leg_poser.current_tip_pose_;
              Pose transition_pose = (reached_target ? target_tip_pose2 : current_tip_pose);
              leg_poser.transition_poses_.push_back(transition_pose);
            }
          }
        } else {
          legs_completed_step_++;
          leg_poser.leg_completed_step_ = true;
        }
      }
    }

    if (direct_step) normalised_progress = progress / std::max(transition_step_count_, 1);
    else normalised_progress = (progress / 2 + (current_group_ == 0 ? 0 : 50)) / std::max(transition_step_count_, 1);

    if (legs_completed_step_ == leg_count_) {
      set_target_ = true;
      legs_completed_step_ = 0;
      if (current_group_ == 1 || direct_step) {
        current_group_ = 0;
        transition_step_ = next_transition_step;
        horizontal_transition_complete_ = !proximity_alert_;
        sequence_complete = final_transition;
        proximity_alert_ = false;
      } else if (current_group_ == 0) {
        current_group_ = 1;
      }
    }
  }

  if (execute_vertical_transition) {
    if (set_target_) {
      set_target_ = false;
      for (int i = 0; i < leg_count_; ++i) {
        Leg& leg = legs[i];
        LegStepper& leg_stepper = leg.stepper;
        LegPoser& leg_poser = leg.poser;
        Vec3 target_tip_position;
        if (int(leg_poser.transition_poses_.size()) > next_transition_step && next_transition_step >= 0) {
          target_tip_position = leg_poser.transition_poses_[next_transition_step].position_;
        } else {
          Vec3 default_tip_position = leg_stepper.default_tip_pose_.position_;
          target_tip_position = current_pose_.inverseTransformVector(default_tip_position);
        }
        target_tip_position[0] = leg.current_tip_pose_.position_[0];
        target_tip_position[1] = leg.current_tip_pose_.position_[1];
        Quat target_tip_rotation = leg_stepper.target_tip_pose_.rotation_;
        leg_poser.target_tip_pose_ = Pose(target_tip_position, target_tip_rotation);
      }
    }

    bool all_legs_within_workspace = true;
    for (int i = 0; i < leg_count_; ++i) {
      Leg& leg = legs[i];
      LegPoser& leg_poser = leg.poser;
      Pose target_tip_pose = leg_poser.target_tip_pose_;
      bool apply_delta = (sequence == START_UP && final_transition);
      double time_to_step = VERTICAL_TRANSITION_TIME / params_.step_frequency;
      time_to_step *= (first_sequence_execution_ ? 2.0 : 1.0);
      progress = leg_poser.stepToPosition(target_tip_pose, Pose::Identity(), 0.0, time_to_step, apply_delta);
      leg.setDesiredTipPose(leg_poser.current_tip_pose_, false);
      double limit_proximity = leg.applyIK();
      all_legs_within_workspace = all_legs_within_workspace && !(limit_proximity < safety_factor);
    }

    if ((!all_legs_within_workspace && first_sequence_execution_) || progress == PROGRESS_COMPLETE) {
      for (int i = 0; i < leg_count_; ++i) {
        LegPoser& leg_poser = legs[i].poser;
        progress = leg_poser.resetStepToPosition();
        if (first_sequence_execution_) {
          bool reached_target = all_legs_within_workspace;
          Pose target_tip_pose = leg_poser.target_tip_pose_;
          Pose current_tip_pose = leg_poser.current_tip_pose_;
          Pose transition_pose = (reached_target ? target_tip_pose : current_tip_pose);
          leg_poser.transition_poses_.push_back(transition_pose);
        }
      }
      vertical_transition_complete_ = all_legs_within_workspace;
      transition_step_ = next_transition_step;
      sequence_complete = final_transition;
      set_target_ = true;
    }
    normalised_progress = progress / std::max(transition_step_count_, 1);
  }

  if (first_sequence_execution_) {
    transition_step_count_ = transition_step_;
    transition_step_target = transition_step_;
  }
  (void)transition_step_target;
  if (transition_step_ > TRANSITION_STEP_THRESHOLD) sequence_failed_ = true;  // ROS_FATAL + ros::shutdown() in the reference

  if (sequence_complete) {
    set_target_ = true;
    vertical_transition_complete_ = false;
    horizontal_transition_complete_ = false;
    first_sequence_execution_ = false;
    return PROGRESS_COMPLETE;
  }
  total_progress = std::min(total_progress + normalised_progress, PROGRESS_COMPLETE - 1);
  return (first_sequence_execution_ ? -1 : total_progress);
}

int Robot::stepToNewStance() {  // pose_controller.cpp:520 (tripod leg coordination)
  int progress = 0;
  int leg_count = leg_count_;
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    if (leg.group_ == current_group_) {
      LegStepper& leg_stepper = leg.stepper;
      LegPoser& leg_poser = leg.poser;
      double step_height = params_.swing_height;
      double step_time = 1.0 / params_.step_frequency;
      Pose target_tip_pose = leg_stepper.default_tip_pose_;
      progress = leg_poser.stepToPosition(target_tip_pose, current_pose_, step_height, step_time);
      leg.setDesiredTipPose(leg_poser.current_tip_pose_);
      leg.applyIK();
      legs_completed_step_ += int(progress == PROGRESS_COMPLETE);
    }
  }
  progress = progress / 2 + current_group_ * 50;
  current_group_ = legs_completed_step_ / (leg_count / 2);
  if (legs_completed_step_ == leg_count) {
    legs_completed_step_ = 0;
    current_group_ = 0;
  }
  reset_transition_sequence_ = true;
  return progress;
}

int Robot::packLegs(double time_to_pack) {  // pose_controller.cpp:597 (one pack step: Joint::packed_positions_ has one entry)
  int progress = 0;
  transition_step_ = 0;  // reset for the start-up / shut-down sequences (:618)
  int number_pack_steps = 1;
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    LegPoser& leg_poser = leg.poser;
    if (!executing_transition_) {
      std::vector<double> packed_configuration(leg.joint_count_, UNASSIGNED_VALUE);
      for (int j = 1; j <= leg.joint_count_; ++j) packed_configuration[j - 1] = leg.joints[j].packed_position_;
      leg_poser.desired_configuration_ = packed_configuration;
    }
    progress = leg_poser.transitionConfiguration(time_to_pack);
  }
  executing_transition_ = (progress != 0 && progress != PROGRESS_COMPLETE);
  if (progress == PROGRESS_COMPLETE && pack_step_ < number_pack_steps - 1) {
    executing_transition_ = false;
    pack_step_++;
    progress = 0;
  }
  return progress;
}

int Robot::unpackLegs(double time_to_unpack) {  // pose_controller.cpp:661
  int progress = 0;
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    LegPoser& leg_poser = leg.poser;
    if (!executing_transition_) {
      std::vector<double> unpacked_configuration(leg.joint_count_, UNASSIGNED_VALUE);
      for (int j = 1; j <= leg.joint_count_; ++j)
        unpacked_configuration[j - 1] = (pack_step_ > 0) ? leg.joints[j].packed_position_ : leg.joints[j].unpacked_position_;
      leg_poser.desired_configuration_ = unpacked_configuration;
    }
    progress = leg_poser.transitionConfiguration(time_to_unpack);
  }
  executing_transition_ = (progress != 0 && progress != PROGRESS_COMPLETE);
  if (progress == PROGRESS_COMPLETE && pack_step_ != 0) {
    executing_transition_ = false;
    pack_step_--;
    progress = 0;
  }
  return progress;
}

void Robot::updateCurrentPose(RobotState robot_state) {  // pose_controller.cpp:811
  Pose new_pose = Pose::Identity();
  updateWalkPlanePose();
  new_pose = new_pose.addPose(walk_plane_pose_);
  default_pose_model_ = walk_plane_pose_;  // model_->setDefaultPose
  if (params_.manual_posing) {
    updateManualPose();
    new_pose = new_pose.addPose(manual_pose_);
  }
  if (params_.inclination_posing) {
    updateInclinationPose();
    new_pose = new_pose.addPose(inclination_pose_);
  }
  if (params_.imu_posing && robot_state == RUNNING) {
    updateIMUPose();
    new_pose = new_pose.addPose(imu_pose_);
  } else if (params_.auto_posing) {
    updateAutoPose();
    new_pose = new_pose.addPose(auto_pose_);
  }
  if (params_.gravity_aligned_tips && legs[0].joint_count_ <= 3) {
    updateTipAlignPose();
    new_pose = new_pose.addPose(tip_align_pose_);
  }
  current_pose_ = new_pose;
}

void Robot::updateManualPose() {  // pose_controller.cpp:863
  double time_delta = params_.time_delta;
  Vec3 current_position = manual_pose_.position_;
  Vec3 current_rotation = quaternionToEulerAngles(manual_pose_.rotation_, true);
  Vec3 default_position = default_pose_.position_;
  Vec3 default_rotation = quaternionToEulerAngles(default_pose_.rotation_, true);
  Vec3 max_position(params_.max_translation[0], params_.max_translation[1], params_.max_translation[2]);
  Vec3 max_rotation(params_.max_rotation[0], params_.max_rotation[1], params_.max_rotation[2]);

  Vec3 translation_limit(0, 0, 0), rotation_limit(0, 0, 0), translation_velocity(0, 0, 0), rotation_velocity(0, 0, 0);
  Vec3 desired_position(0, 0, 0), desired_rotation(0, 0, 0);

  bool reset_translation[3] = {false, false, false};
  bool reset_rotation[3] = {false, false, false};
  switch (pose_reset_mode_) {
    case Z_AND_YAW_RESET:
      reset_translation[2] = true;
      reset_rotation[2] = true;
      break;
    case X_AND_Y_RESET:
      reset_translation[0] = true;
      reset_translation[1] = true;
      break;
    case PITCH_AND_ROLL_RESET:
      reset_rotation[0] = true;
      reset_rotation[1] = true;
      break;
    case ALL_RESET:
      reset_translation[0] = reset_translation[1] = reset_translation[2] = true;
      reset_rotation[0] = reset_rotation[1] = reset_rotation[2] = true;
      break;
    case IMMEDIATE_ALL_RESET:
      manual_pose_ = default_pose_;
      return;
    case NO_RESET:
    default:
      break;
  }

  for (int i = 0; i < 3; i++) {
    if (reset_translation[i]) {
      double diff = current_position[i] - default_position[i];
      if (diff < 0) translation_velocity_input_[i] = 1.0;
      else if (diff > 0) translation_velocity_input_[i] = -1.0;
    }
    if (reset_rotation[i]) {
      double diff = current_rotation[i] - default_rotation[i];
      if (diff < 0) rotation_velocity_input_[i] = 1.0;
      else if (diff > 0) rotation_velocity_input_[i] = -1.0;
    }
    translation_velocity[i] = translation_velocity_input_[i] * params_.max_translation_velocity;
    rotation_velocity[i] = rotation_velocity_input_[i] * params_.max_rotation_velocity;
    desired_position[i] = current_position[i] + translation_velocity[i] * time_delta;
    desired_rotation[i] = current_rotation[i] + rotation_velocity[i] * time_delta;

    translation_limit[i] = sign(translation_velocity[i]) * max_position[i];
    if (reset_translation[i] && default_position[i] < max_position[i] && default_position[i] > -max_position[i])
      translation_limit[i] = default_position[i];
    bool positive_translation_velocity = sign(translation_velocity[i]) > 0;
    bool exceeds_positive_translation_limit = positive_translation_velocity && desired_position[i] > translation_limit[i];
    bool exceeds_negative_translation_limit = !positive_translation_velocity && desired_position[i] < translation_limit[i];
    if (exceeds_positive_translation_limit || exceeds_negative_translation_limit)
      translation_velocity[i] = (translation_limit[i] - current_position[i]) / time_delta;

    rotation_limit[i] = sign(rotation_velocity[i]) * max_rotation[i];
    if (reset_rotation[i] && default_rotation[i] < max_rotation[i] && default_rotation[i] > -max_rotation[i])
      rotation_limit[i] = default_rotation[i];
    bool positive_rotation_velocity = sign(rotation_velocity[i]) > 0;
    bool exceeds_positive_rotation_limit = positive_rotation_velocity && desired_rotation[i] > rotation_limit[i];
    bool exceeds_negative_rotation_limit = !positive_rotation_velocity && desired_rotation[i] < rotation_limit[i];
    if (exceeds_positive_rotation_limit || exceeds_negative_rotation_limit)
      rotation_velocity[i] = (rotation_limit[i] - current_rotation[i]) / time_delta;

    desired_position[i] = current_position[i] + translation_velocity[i] * time_delta;
    desired_rotation[i] = current_rotation[i] + rotation_velocity[i] * time_delta;
  }
  manual_pose_.position_ = desired_position;
  manual_pose_.rotation_ = correctRotation(eulerAnglesToQuaternion(desired_rotation, true), Quat::Identity());
}

void Robot::updateTipAlignPose() {  // pose_controller.cpp:1024
  for (int li = 0; li < leg_count_; ++li) {
    Leg& leg = legs[li];
    LegStepper& leg_stepper = leg.stepper;
    double swing_progress = leg_stepper.swing_progress_;
    if (swing_progress != -1.0) {
      Vec3 walk_plane_normal = leg_stepper.walk_plane_normal_;
      Quat walk_plane_rotation = fromTwoVectors(UnitZ(), walk_plane_normal);
      Vec3 tip_position = leg.tipPoseRobotFrame().position_;
      Vec3 joint_position = leg.jointPoseRobotFrame(leg.joint_count_).position_;
      Vec3 tip_to_joint = joint_position - tip_position;
      double link_length = (tip_position - joint_position).norm();
      Vec3 a = walk_plane_rotation.transformVector(tip_to_joint);
      Vec3 b = link_length * walk_plane_normal;
      Vec3 rejection = a - (a.dot(b) / b.dot(b)) * b;
      Vec3 translation_to_alignment = -rejection;
      a = tip_align_pose_.position_;
      b = walk_plane_normal;
      rejection = a - (a.dot(b) / b.dot(b)) * b;
      Vec3 current_walk_plane_aligned_translation = rejection;
      Vec3 target_translation = current_walk_plane_aligned_translation + translation_to_alignment;
      Vec3 limit(params_.max_translation[0], params_.max_translation[1], params_.max_translation[2]);
      target_translation = clampedVecBuggy(target_translation, limit);
      double c = smoothStep(swing_progress);
      if (swing_progress < 0.5) {
        c = smoothStep(c * 2.0);
        tip_align_pose_ = origin_tip_align_pose_.interpolate(c, Pose::Identity());
      } else if (swing_progress >= 0.5) {
        c = smoothStep((c - 0.5) * 2.0);
        tip_align_pose_ = Pose::Identity().interpolate(c, Pose(target_translation, Quat::Identity()));
      }
      if (swing_progress == 1.0) origin_tip_align_pose_ = tip_align_pose_;
    }
  }
}

void Robot::updateWalkPlanePose() {  // pose_controller.cpp:1092
  Vec3 walk_plane(0, 0, 0);
  Vec3 walk_plane_normal = UnitZ();
  double c = 0.0;
  for (int li = 0; li < leg_count_; ++li) {
    LegStepper& leg_stepper = legs[li].stepper;
    double swing_progress_scaler = std::max(1.0, double(params_.swing_phase) / params_.phase_offset);
    double swing_progress = leg_stepper.swing_progress_ * swing_progress_scaler;
    if (swing_progress >= 0 && swing_progress <= 1.0) {
      c = smoothStep(swing_progress);
      walk_plane = leg_stepper.walk_plane_;
      walk_plane_normal = leg_stepper.walk_plane_normal_;
    }
  }
  Pose new_walk_plane_pose;
  new_walk_plane_pose.rotation_ = fromTwoVectors(UnitZ(), walk_plane_normal);
  new_walk_plane_pose.rotation_ = correctRotation(new_walk_plane_pose.rotation_, Quat::Identity());
  Vec3 body_clearance(0, 0, params_.body_clearance);
  new_walk_plane_pose.position_ = new_walk_plane_pose.rotation_.transformVector(body_clearance);
  new_walk_plane_pose.position_[2] += walk_plane[2];
  walk_plane_pose_ = origin_walk_plane_pose_.interpolate(c, new_walk_plane_pose);
  if (c == 1.0) origin_walk_plane_pose_ = walk_plane_pose_;
}

void Robot::updateAutoPose() {  // pose_controller.cpp:1134
  LegStepper& leg_stepper = legs[auto_pose_reference_leg_].stepper;
  auto_pose_ = Pose::Identity();
  bool zero_body_velocity = leg_stepper.stride_vector_.norm() == 0;
  if (walk_state_ == STARTING || walk_state_ == MOVING) {
    auto_posing_state_ = POSING;
  } else if ((zero_body_velocity && walk_state_ == STOPPING) || walk_state_ == STOPPED) {
    auto_posing_state_ = STOP_POSING;
  }
  int master_phase;
  bool sync_with_step_cycle = (pose_frequency_ == -1.0);
  if (sync_with_step_cycle) {
    master_phase = leg_stepper.phase_;
  } else {
    master_phase = pose_phase_;
    pose_phase_ = (pose_phase_ + 1) % pose_phase_length_;
  }
  int auto_posers_complete = 0;
  for (AutoPoser& auto_poser : auto_posers_) {
    Pose updated_pose = auto_poser.updatePose(master_phase);
    auto_posers_complete += int(!auto_poser.allow_posing_);
    auto_pose_ = auto_pose_.addPose(updated_pose);
  }
  if (auto_posers_complete == int(auto_posers_.size())) auto_posing_state_ = POSING_COMPLETE;
  for (int li = 0; li < leg_count_; ++li) legs[li].poser.updateAutoPose(master_phase);
}

void Robot::updateIMUPose() {  // pose_controller.cpp:1191
  Quat current_rotation = correctRotation(getImuData().orientation, Quat::Identity());
  Quat target_rotation = correctRotation(manual_pose_.rotation_, Quat::Identity());
  Quat rotation_error = (current_rotation * target_rotation.inverse()).normalized();
  double kp = params_.rotation_pid_p, ki = params_.rotation_pid_i, kd = params_.rotation_pid_d;
  rotation_position_error_ = quaternionToEulerAngles(rotation_error);
  rotation_position_error_[2] = 0.0;
  if (rotation_position_error_.norm() < IMU_POSING_DEADBAND) return;
  rotation_absement_error_ += rotation_position_error_ * params_.time_delta;
  double smoothing_factor = 0.15;
  rotation_velocity_error_ =
      smoothing_factor * -getImuData().angular_velocity + (1 - smoothing_factor) * rotation_velocity_error_;
  Vec3 rotation_correction =
      -(kd * rotation_velocity_error_ + kp * rotation_position_error_ + ki * rotation_absement_error_);
  double max_roll = params_.max_rotation[0];
  double max_pitch = params_.max_rotation[1];
  rotation_correction[0] = clamped(rotation_correction[0], -max_roll, max_roll);
  rotation_correction[1] = clamped(rotation_correction[1], -max_pitch, max_pitch);
  rotation_correction[2] = quaternionToEulerAngles(target_rotation)[2];
  if (rotation_correction.norm() > STABILITY_THRESHOLD) imu_unstable_ = true;  // reference: ROS_FATAL + shutdown
  imu_pose_.rotation_ = eulerAnglesToQuaternion(rotation_correction);
  imu_pose_.rotation_ = correctRotation(imu_pose_.rotation_, target_rotation);
}

void Robot::updateInclinationPose() {  // pose_controller.cpp:1240
  Quat compensation_combined = (manual_pose_.rotation_ * auto_pose_.rotation_).normalized();
  Quat compensation_removed = (getImuData().orientation * compensation_combined.inverse()).normalized();
  Vec3 euler = quaternionToEulerAngles(compensation_removed);
  double body_height = params_.body_clearance;
  double longitudinal_correction = -body_height * std::tan(euler[1]);
  double lateral_correction = body_height * std::tan(euler[0]);
  double max_translation_x = params_.max_translation[0];
  double max_translation_y = params_.max_translation[1];
  longitudinal_correction = clamped(longitudinal_correction, -max_translation_x, max_translation_x);
  lateral_correction = clamped(lateral_correction, -max_translation_y, max_translation_y);
  inclination_pose_.position_[0] = longitudinal_correction;
  inclination_pose_.position_[1] = lateral_correction;
}

Pose AutoPoser::updatePose(int phase) {  // pose_controller.cpp:1338
  Pose return_pose = Pose::Identity();
  int start_phase = start_phase_ * robot->normaliser_;
  int end_phase = end_phase_ * robot->normaliser_;
  if (start_phase > end_phase) {
    end_phase += robot->pose_phase_length_;
    if (phase < start_phase) phase += robot->pose_phase_length_;
  }
  PosingState state = robot->auto_posing_state_;
  bool sync_with_step_cycle = (robot->pose_frequency_ == -1.0);
  start_check_ = !sync_with_step_cycle || (!start_check_ && state == POSING && phase == start_phase);
  end_check_first = (end_check_first || (state == STOP_POSING && phase == start_phase));
  end_check_second = (end_check_second || (state == STOP_POSING && phase == end_phase && end_check_first));
  if (!allow_posing_ && start_check_) {
    allow_posing_ = true;
    end_check_first = end_check_second = false;
  } else if (allow_posing_ && sync_with_step_cycle && end_check_first && end_check_second) {
    allow_posing_ = false;
    start_check_ = false;
  }
  if (phase >= start_phase && phase < end_phase && allow_posing_) {
    int iteration = phase - start_phase + 1;
    int num_iterations = end_phase - start_phase;
    Vec3 zero(0.0, 0.0, 0.0);
    Vec3 position_control_nodes[5] = {zero, zero, zero, zero, zero};
    Vec3 rotation_control_nodes[5] = {zero, zero, zero, zero, zero};
    bool first_half = iteration <= num_iterations / 2;
    Vec3 gravity_direction = robot->estimateGravity().normalized();
    Vec3 rot_amp(roll_amplitude_, pitch_amplitude_, yaw_amplitude_);
    Vec3 pos_amp = gravity_amplitude_ != 0.0 ? gravity_direction * gravity_amplitude_
                                             : Vec3(x_amplitude_, y_amplitude_, z_amplitude_);
    if (first_half) {
      rotation_control_nodes[3] = rot_amp;
      rotation_control_nodes[4] = rot_amp;
      position_control_nodes[3] = pos_amp;
      position_control_nodes[4] = pos_amp;
    } else {
      rotation_control_nodes[0] = rot_amp;
      rotation_control_nodes[1] = rot_amp;
      position_control_nodes[0] = pos_amp;
      position_control_nodes[1] = pos_amp;
    }
    double delta_t = 1.0 / (num_iterations / 2.0);
    int offset = static_cast<int>((first_half ? 0 : num_iterations / 2.0));
    double time_input = (iteration - offset) * delta_t;
    Vec3 position = quarticBezier(position_control_nodes, time_input);
    Vec3 rotation = quarticBezier(rotation_control_nodes, time_input);
    return_pose = Pose(position, eulerAnglesToQuaternion(rotation));
  }
  return return_pose;
}

int LegPoser::transitionConfiguration(double transition_time) {  // pose_controller.cpp:1476
  if (desired_configuration_.size() == 0) return PROGRESS_COMPLETE;
  if (first_iteration_) {
    origin_configuration_.clear();
    for (int j = 1; j <= leg_->joint_count_; ++j) origin_configuration_.push_back(leg_->joints[j].desired_position_);
    first_iteration_ = false;
    master_iteration_count_ = 0;
  }
  int num_iterations = std::max(1, int(roundToInt(transition_time / robot->params_.time_delta)));
  double delta_t = 1.0 / num_iterations;
  master_iteration_count_++;
  for (int j = 1; j <= leg_->joint_count_; ++j) {
    Joint& joint = leg_->joints[j];
    double control_nodes[4];
    control_nodes[0] = origin_configuration_[j - 1];
    control_nodes[1] = origin_configuration_[j - 1];
    control_nodes[2] = desired_configuration_[j - 1];
    control_nodes[3] = desired_configuration_[j - 1];
    joint.prev_desired_position_ = joint.desired_position_;
    joint.desired_position_ = cubicBezier(control_nodes, master_iteration_count_ * delta_t);
  }
  leg_->applyFK();
  int progress = int((double(master_iteration_count_ - 1) / double(num_iterations)) * PROGRESS_COMPLETE);
  progress = clampedInt(progress, 1, PROGRESS_COMPLETE);
  if (master_iteration_count_ >= num_iterations) {
    first_iteration_ = true;
    return PROGRESS_COMPLETE;
  }
  return progress;
}

int LegPoser::stepToPosition(const Pose& target_tip_pose, const Pose& target_pose, double lift_height,
                             double time_to_step, bool apply_delta) {  // pose_controller.cpp:1571
  if (first_iteration_) {
    origin_tip_pose_ = leg_->current_tip_pose_;
    master_iteration_count_ = 0;
    first_iteration_ = false;
  }
  Pose desired_tip_pose = target_tip_pose;
  if (desired_tip_pose == Pose::Undefined()) {
    desired_tip_pose = origin_tip_pose_;
    desired_tip_pose.rotation_ = UndefinedRotation();
  }
  Vec3 position_delta = origin_tip_pose_.position_ - target_pose.inverseTransformVector(desired_tip_pose.position_);
  bool transition_position = position_delta.norm() > TIP_TOLERANCE;
  bool transition_rotation = false;
  if (!isApproxQuat(desired_tip_pose.rotation_, UndefinedRotation())) {
    Vec3 origin_tip_direction = origin_tip_pose_.rotation_.transformVector(UnitX());
    Vec3 desired_tip_direction = desired_tip_pose.rotation_.transformVector(UnitX());
    double angle;
    Vec3 axis;
    angleAxisFromQuat(fromTwoVectors(origin_tip_direction, desired_tip_direction), &angle, &axis);
    transition_rotation = angle > JOINT_TOLERANCE;
  }
  if (!transition_position && !transition_rotation && lift_height == 0.0) {
    first_iteration_ = true;
    current_tip_pose_ = origin_tip_pose_;
    return PROGRESS_COMPLETE;
  }
  bool manually_manipulated = (leg_->leg_state_ == MANUAL || leg_->leg_state_ == WALKING_TO_MANUAL);
  if (apply_delta && !manually_manipulated) desired_tip_pose.position_ += leg_->admittance_delta_;

  master_iteration_count_++;
  int num_iterations = std::max(1, int(roundToInt(time_to_step / robot->params_.time_delta)));
  double delta_t = 1.0 / num_iterations;
  double completion_ratio = (double(master_iteration_count_ - 1) / double(num_iterations));
  Pose desired_pose = Pose::Identity().interpolate(smoothStep(completion_ratio), target_pose);

  Quat new_tip_rotation = UndefinedRotation();
  if (!isApproxQuat(desired_tip_pose.rotation_, UndefinedRotation())) {
    Vec3 origin_tip_direction = origin_tip_pose_.rotation_.transformVector(UnitX());
    Vec3 desired_tip_direction = desired_tip_pose.rotation_.transformVector(UnitX());
    Vec3 new_tip_direction = om::interpolate(origin_tip_direction, desired_tip_direction, smoothStep(completion_ratio));
    new_tip_rotation = fromTwoVectors(UnitX(), new_tip_direction.normalized());
  }

  double time_input;
  Vec3 new_tip_position = origin_tip_pose_.position_;
  if (desired_tip_pose.position_ != UndefinedPosition()) {
    int half_swing_iteration = num_iterations / 2;
    Vec3 control_nodes_primary[5], control_nodes_secondary[5];
    Vec3 origin_to_target = origin_tip_pose_.position_ - desired_tip_pose.position_;
    control_nodes_primary[0] = origin_tip_pose_.position_;
    control_nodes_primary[1] = origin_tip_pose_.position_;
    control_nodes_primary[2] = origin_tip_pose_.position_;
    control_nodes_primary[3] = desired_tip_pose.position_ + 0.75 * origin_to_target;
    control_nodes_primary[4] = desired_tip_pose.position_ + 0.5 * origin_to_target;
    control_nodes_primary[2][2] += lift_height;
    control_nodes_primary[3][2] += lift_height;
    control_nodes_primary[4][2] += lift_height;
    control_nodes_secondary[0] = desired_tip_pose.position_ + 0.5 * origin_to_target;
    control_nodes_secondary[1] = desired_tip_pose.position_ + 0.25 * origin_to_target;
    control_nodes_secondary[2] = desired_tip_pose.position_;
    control_nodes_secondary[3] = desired_tip_pose.position_;
    control_nodes_secondary[4] = desired_tip_pose.position_;
    control_nodes_secondary[0][2] += lift_height;
    control_nodes_secondary[1][2] += lift_height;
    control_nodes_secondary[2][2] += lift_height;
    int swing_iteration_count = (master_iteration_count_ + (num_iterations - 1)) % (num_iterations) + 1;
    if (swing_iteration_count <= half_swing_iteration) {
      time_input = swing_iteration_count * delta_t * 2.0;
      new_tip_position = quarticBezier(control_nodes_primary, time_input);
    } else {
      time_input = (swing_iteration_count - half_swing_iteration) * delta_t * 2.0;
      new_tip_position = quarticBezier(control_nodes_secondary, time_input);
    }
  }
  if (leg_->leg_state_ != MANUAL) {
    current_tip_pose_.position_ = desired_pose.inverseTransformVector(new_tip_position);
    current_tip_pose_.rotation_ = new_tip_rotation;
  }
  if (master_iteration_count_ >= num_iterations) {
    first_iteration_ = true;
    return PROGRESS_COMPLETE;
  }
  return int(completion_ratio * PROGRESS_COMPLETE);
}

void LegPoser::updateAutoPose(int phase) {  // pose_controller.cpp:1716
  int start_phase = pose_negation_phase_start_ * robot->normaliser_;
  int end_phase = pose_negation_phase_end_ * robot->normaliser_;
  int negation_phase = phase;
  if (start_phase == 0) start_phase = robot->pose_phase_length_;
  if (end_phase == 0) end_phase = robot->pose_phase_length_;
  if (start_phase > end_phase) {
    end_phase += robot->pose_phase_length_;
    if (negation_phase < start_phase) negation_phase += robot->pose_phase_length_;
  }
  StepState step_state = leg_->stepper.step_state_;
  if (step_state != FORCE_STANCE && step_state != FORCE_STOP && negation_phase == start_phase) negate_auto_pose_ = true;
  if (negation_phase < start_phase || negation_phase > end_phase) negate_auto_pose_ = false;
  auto_pose_ = robot->auto_pose_;
  if (negate_auto_pose_) {
    int iteration = negation_phase - start_phase + 1;
    int num_iterations = end_phase - start_phase;
    bool first_half = iteration <= num_iterations / 2;
    double control_input = 1.0;
    if (negation_transition_ratio_ > 0.0) {
      if (first_half) control_input = std::min(1.0, iteration / (num_iterations * negation_transition_ratio_));
      else control_input = std::min(1.0, (num_iterations - iteration) / (num_iterations * negation_transition_ratio_));
    }
    control_input = smoothStep(control_input);
    Pose negation = Pose::Identity().interpolate(control_input, auto_pose_);
    auto_pose_ = auto_pose_.removePose(negation);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// AdmittanceController (admittance_controller.cpp)
// ---------------------------------------------------------------------------------------------------------------
void Robot::updateAdmittance() {  // admittance_controller.cpp:22
  for (int li = 0; li < leg_count_; ++li) {
    Leg& leg = legs[li];
    Vec3 admittance_delta(0, 0, 0);
    bool use_calculated_tip_force = params_.use_joint_effort;
    Vec3 tip_force = use_calculated_tip_force ? leg.tip_force_calculated_ : leg.tip_force_measured_;
    tip_force *= params_.force_gain;
    for (int i = 0; i < 3; ++i) {
      double force_input = std::max(tip_force[i], 0.0);
      double damping = params_.virtual_damping_ratio;
      double stiffness = params_.virtual_stiffness;  // global value, not the per-leg dynamic stiffness (trap 3)
      double mass = params_.virtual_mass;
      double step_time = params_.integrator_step_time;
      double* x = leg.admittance_state_;
      double virtual_damping = damping * 2 * std::sqrt(mass * stiffness);
      // boost::numeric::odeint::integrate_const(runge_kutta4, sys, x, 0.0, step_time, step_time/30): 30 classic RK4
      // steps (Boost.Odeint unpinned; algorithm as published — SURVEY.md §8c).
      auto sys = [&](const double s[2], double dxdt[2]) {
        dxdt[0] = s[1];
        dxdt[1] = -force_input / mass - virtual_damping / mass * s[1] - stiffness / mass * s[0];
      };
      const double dt = step_time / 30;
      for (int step = 0; step < 30; ++step) {
        double k1[2], k2[2], k3[2], k4[2], tmp[2];
        sys(x, k1);
        tmp[0] = x[0] + dt * 0.5 * k1[0]; tmp[1] = x[1] + dt * 0.5 * k1[1];
        sys(tmp, k2);
        tmp[0] = x[0] + dt * 0.5 * k2[0]; tmp[1] = x[1] + dt * 0.5 * k2[1];
        sys(tmp, k3);
        tmp[0] = x[0] + dt * k3[0]; tmp[1] = x[1] + dt * k3[1];
        sys(tmp, k4);
        x[0] += dt / 6.0 * (k1[0] + 2.0 * k2[0] + 2.0 * k3[0] + k4[0]);
        x[1] += dt / 6.0 * (k1[1] + 2.0 * k2[1] + 2.0 * k3[1] + k4[1]);
      }
      double delta = clamped(-x[0], -0.2, 0.2);
      double delta_direction = delta / std::fabs(delta);
      if (std::fabs(delta) > ADMITTANCE_DEADBAND)
        admittance_delta[i] = delta_direction * (std::fabs(delta) - ADMITTANCE_DEADBAND) / (1 - ADMITTANCE_DEADBAND);
    }
    leg.setAdmittanceDelta(admittance_delta);
  }
}

void Robot::updateStiffness() {  // admittance_controller.cpp:96 (walker overload)
  for (int li = 0; li < leg_count_; ++li) legs[li].virtual_stiffness_ = params_.virtual_stiffness;
  for (int li = 0; li < leg_count_; ++li) {
    Leg& leg = legs[li];
    LegStepper& leg_stepper = leg.stepper;
    if (leg_stepper.step_state_ == SWING) {
      double z_diff = leg_stepper.current_tip_pose_.position_[2] - leg_stepper.default_tip_pose_.position_[2];
      double step_reference = 0;
      step_reference += std::fabs(z_diff / params_.swing_height);
      int leg_id = leg.id_number_;
      Leg& adjacent_leg_1 = legs[mod(leg_id - 1, leg_count_)];
      Leg& adjacent_leg_2 = legs[mod(leg_id + 1, leg_count_)];
      double virtual_stiffness = params_.virtual_stiffness;
      double swing_stiffness = virtual_stiffness * (step_reference * (params_.swing_stiffness_scaler - 1) + 1);
      double load_stiffness = virtual_stiffness * (step_reference * (params_.load_stiffness_scaler - 1));
      double current_stiffness_1 = adjacent_leg_1.virtual_stiffness_;
      double current_stiffness_2 = adjacent_leg_2.virtual_stiffness_;
      leg.virtual_stiffness_ = swing_stiffness;
      adjacent_leg_1.virtual_stiffness_ = current_stiffness_1 + load_stiffness;
      adjacent_leg_2.virtual_stiffness_ = current_stiffness_2 + load_stiffness;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// StateController harness (state_controller.cpp:127-447, 1098-1136)
// ---------------------------------------------------------------------------------------------------------------
void Robot::stateInit() {  // StateController::init :127 + initModel(use_default=true) (state_controller.h:64)
  walkerInit();
  poserInit();
  robot_state_ = UNKNOWN;
  initLegs(true);
}

void Robot::requestRobotState(RobotState input_state) {  // robotStateCallback :1098
  if (input_state != robot_state_ && !transition_state_flag_) {
    new_robot_state_ = input_state;
    if (new_robot_state_ > robot_state_) {
      new_robot_state_ = static_cast<RobotState>(robot_state_ + 1);
      transition_state_flag_ = true;
    } else if (new_robot_state_ < robot_state_) {
      new_robot_state_ = static_cast<RobotState>(robot_state_ - 1);
      transition_state_flag_ = true;
    }
  }
}

void Robot::setBodyVelocityInput(double vx, double vy, double wz) {  // bodyVelocityInputCallback :1127
  if (robot_state_ == RUNNING) {
    linear_velocity_input_[0] = vx * params_.body_velocity_scaler;
    linear_velocity_input_[1] = vy * params_.body_velocity_scaler;
    angular_velocity_input_ = wz * params_.body_velocity_scaler;
    double n = std::sqrt(linear_velocity_input_[0] * linear_velocity_input_[0] +
                         linear_velocity_input_[1] * linear_velocity_input_[1]);
    if (params_.velocity_input_mode == SHC_VELOCITY_THROTTLE && n > 1.0) {
      // min(1.0, norm) * input.normalized()
      double s = std::min(1.0, n);
      linear_velocity_input_[0] = s * (linear_velocity_input_[0] / n);
      linear_velocity_input_[1] = s * (linear_velocity_input_[1] / n);
    }
  }
}

void Robot::loop() {  // StateController::loop :162
  if (robot_state_ != UNKNOWN) {
    updateCurrentPose(robot_state_);
    pose_state_ = auto_posing_state_;  // walker_->setPoseState(poser_->getAutoPoseState())
    if (params_.admittance_control) {
      if (walk_state_ != STOPPED && params_.dynamic_stiffness) updateStiffness();
      updateAdmittance();
    }
  }
  if (transition_state_flag_) transitionRobotState();
  if (robot_state_ == RUNNING) runningState();
}

void Robot::transitionRobotState() {  // :196 — direct (start_up_sequence: false) transitions only
  if (robot_state_ == UNKNOWN) {
    // default.yaml has start_up_sequence false: whatever the joint guess says, the state becomes PACKED (:226-256).
    robot_state_ = PACKED;
    new_robot_state_ = robot_state_;
  } else if (robot_state_ == PACKED && new_robot_state_ == READY) {
    int progress = directStartup();
    if (progress == PROGRESS_COMPLETE) {
      robot_state_ = READY;
      updateDefaultConfiguration();
      generateWorkspaces();
      generateWalkspace();
    }
  } else if (robot_state_ == READY && new_robot_state_ == RUNNING) {
    robot_state_ = RUNNING;
  } else if (robot_state_ == RUNNING && new_robot_state_ != RUNNING) {
    transition_state_flag_ = false;
  }
  if (robot_state_ == new_robot_state_) transition_state_flag_ = false;
}

void Robot::runningState() {  // :379 (no gait change / leg toggle / planner / cruise / parameter adjust in the harness)
  bool update_tip_position = true;
  if (transition_state_flag_) {
    transitionRobotState();
    update_tip_position = false;
  }
  update_tip_position = update_tip_position || walk_state_ != STOPPED;
  if (update_tip_position) {
    updateWalk(linear_velocity_input_, angular_velocity_input_);
    // updateManual x2 are no-ops: no leg is in MANUAL state in the harness (walk_controller.cpp:661,718)
    updateStance();
    updateModel();
  }
}

int Robot::startUp() {
  int loops = 0;
  while (robot_state_ != READY && loops < 100000) {
    requestRobotState(RUNNING);  // the remote keeps publishing the desired state between cycles (main.cpp:130)
    loop();
    ++loops;
  }
  // READY -> RUNNING is a pure state change (state_controller.cpp:277-281); the harness applies it between cycles so
  // that cycle 0 of a rollout is the first full loop() in RUNNING state.
  robot_state_ = RUNNING;
  new_robot_state_ = RUNNING;
  transition_state_flag_ = false;
  return loops;
}

}  // namespace shc_oracle
