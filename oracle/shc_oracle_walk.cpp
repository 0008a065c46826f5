// shc_oracle_walk.cpp — TEST INFRASTRUCTURE ONLY.  Restates /root/reference/src/walk_controller.cpp in IEEE
// double for the parity oracle.  PINNED to the reference's own code (oracle/_ref, tests/test_reference_pin.py; see shc_oracle.hpp).
#include "shc_oracle.hpp"

namespace shc_oracle {

// ---------------------------------------------------------------------------------------------------------------
// WalkController
// ---------------------------------------------------------------------------------------------------------------
void Robot::walkerInit() {  // walk_controller.cpp:22
  time_delta_ = params_.time_delta;
  walk_state_ = STOPPED;
  walk_plane_ = Vec3(0, 0, 0);
  walk_plane_normal_ = UnitZ();
  odometry_ideal_ = Pose::Identity();
  for (int i = 0; i < leg_count_; ++i) {
    Leg& leg = legs[i];
    double x_position = params_.stance_x[i];
    double y_position = params_.stance_y[i];
    Quat identity_tip_rotation = UndefinedRotation();
    if (leg.joint_count_ > 3 && params_.gravity_aligned_tips) {
      identity_tip_rotation = fromTwoVectors(UnitX(), -UnitZ());
      identity_tip_rotation = correctRotation(identity_tip_rotation, Quat::Identity());
    }
    Pose identity_tip_pose(Vec3(x_position, y_position, 0.0), identity_tip_rotation);
    leg.stepper.construct(this, &leg, identity_tip_pose);
  }
  desired_linear_velocity_[0] = desired_linear_velocity_[1] = 0;
  desired_angular_velocity_ = 0;
  generateStepCycle();
}

void Robot::generateWalkspace() {  // walk_controller.cpp:57
  walkspace_.clear();
  for (int li = 0; li < leg_count_; ++li) {
    int leg_count = leg_count_;
    Leg& leg = legs[li];
    Leg& adjacent_leg_1 = legs[mod(leg.id_number_ + 1, leg_count)];
    Leg& adjacent_leg_2 = legs[mod(leg.id_number_ - 1, leg_count)];
    Vec3 default_tip_position = leg.stepper.default_tip_pose_.position_;
    Vec3 adjacent_1_tip_position = adjacent_leg_1.stepper.default_tip_pose_.position_;
    Vec3 adjacent_2_tip_position = adjacent_leg_2.stepper.default_tip_pose_.position_;

    double distance_to_adjacent_leg_1 = (default_tip_position - adjacent_1_tip_position).norm() / 2.0;
    double distance_to_adjacent_leg_2 = (default_tip_position - adjacent_2_tip_position).norm() / 2.0;
    double bearing_to_adjacent_leg_1 = radiansToDegrees(std::atan2(adjacent_1_tip_position[1] - default_tip_position[1],
                                                                   adjacent_1_tip_position[0] - default_tip_position[0]));
    double bearing_to_adjacent_leg_2 = radiansToDegrees(std::atan2(adjacent_2_tip_position[1] - default_tip_position[1],
                                                                   adjacent_2_tip_position[0] - default_tip_position[0]));
    for (int bearing = 0; bearing <= 360; bearing += BEARING_STEP) {
      int bearing_diff_1 = std::abs(mod(static_cast<int>(bearing_to_adjacent_leg_1), 360) - bearing);
      int bearing_diff_2 = std::abs(mod(static_cast<int>(bearing_to_adjacent_leg_2), 360) - bearing);
      double distance_to_overlap_1 = UNASSIGNED_VALUE;
      double distance_to_overlap_2 = UNASSIGNED_VALUE;
      if ((bearing_diff_1 < 90 || bearing_diff_1 > 270) && distance_to_adjacent_leg_1 > 0.0)
        distance_to_overlap_1 = distance_to_adjacent_leg_1 / std::cos(degreesToRadians(bearing_diff_1));
      if ((bearing_diff_2 < 90 || bearing_diff_2 > 270) && distance_to_adjacent_leg_2 > 0.0)
        distance_to_overlap_2 = distance_to_adjacent_leg_2 / std::cos(degreesToRadians(bearing_diff_2));
      bool overlapping = params_.overlapping_walkspaces;
      double min_distance = overlapping ? MAX_WORKSPACE_RADIUS : std::min(distance_to_overlap_1, distance_to_overlap_2);
      min_distance = std::min(min_distance, MAX_WORKSPACE_RADIUS);
      if (walkspace_.find(bearing) != walkspace_.end() && min_distance < walkspace_[bearing]) {
        walkspace_[bearing] = min_distance;
      } else {
        walkspace_.insert(LimitMap::value_type(bearing, min_distance));
      }
    }
  }

  for (int li = 0; li < leg_count_; ++li) {
    Leg& leg = legs[li];
    LegStepper& leg_stepper = leg.stepper;
    Pose current_pose = current_pose_;
    Vec3 identity_tip_position = current_pose.inverseTransformVector(leg_stepper.identity_tip_pose_.position_);
    Vec3 default_tip_position = current_pose.inverseTransformVector(leg_stepper.default_tip_pose_.position_);
    Vec3 default_shift = default_tip_position - identity_tip_position;
    double target_workplane_height = default_shift[2];
    LimitMap workplane = leg.getWorkplane(target_workplane_height);
    if (workplane.empty()) continue;

    for (auto walkspace_it = walkspace_.begin(); walkspace_it != walkspace_.end(); ++walkspace_it) {
      int bearing = walkspace_it->first;
      double radius = walkspace_it->second;
      if (default_shift.norm() == 0.0) {
        radius = workplane.at(bearing);
      } else {
        Vec3 new_point = UnitX() * MAX_WORKSPACE_RADIUS;
        new_point = angleAxisRotate(degreesToRadians(bearing), UnitZ(), new_point);
        new_point = setPrecision(new_point, 3);
        for (auto workplane_it = workplane.begin(); workplane_it != workplane.end(); ++workplane_it) {
          int bearing_1 = workplane_it->first;
          double radius_1 = workplane_it->second;
          Vec3 point_1 = UnitX() * radius_1;
          point_1 = angleAxisRotate(degreesToRadians(bearing_1), UnitZ(), point_1);
          point_1 -= default_shift;
          point_1[2] = 0.0;
          point_1 = setPrecision(point_1, 3);
          if (bearing_1 == workplane.rbegin()->first) {
            radius = 0.0;
            break;
          }
          int bearing_2 = std::next(workplane_it)->first;
          double radius_2 = std::next(workplane_it)->second;
          Vec3 point_2 = UnitX() * radius_2;
          point_2 = angleAxisRotate(degreesToRadians(bearing_2), UnitZ(), point_2);
          point_2 -= default_shift;
          point_2[2] = 0.0;
          point_2 = setPrecision(point_2, 3);
          if (point_1.cross(new_point).norm() == 0.0) {
            radius = point_1.norm();
            break;
          } else if (point_2.cross(new_point).norm() == 0.0) {
            radius = point_2.norm();
            break;
          } else if (point_1.cross(new_point).dot(point_1.cross(point_2)) >= 0.0 &&
                     point_2.cross(new_point).dot(point_2.cross(point_1)) >= 0.0) {
            double dx = point_2[0] - point_1[0];
            double dy = point_2[1] - point_1[1];
            Vec3 normal_1 = Vec3(dy, -dx, 0.0).normalized();
            Vec3 normal_2 = Vec3(-dy, dx, 0.0).normalized();
            bool same_direction_as_new_point = getProjection(new_point, normal_1).dot(normal_1) >= 0.0;
            Vec3 normal = (same_direction_as_new_point ? normal_1 : normal_2);
            Vec3 new_point_projection = getProjection(new_point, normal);
            Vec3 point_1_projection = getProjection(point_1, normal);
            double ratio = point_1_projection.norm() / new_point_projection.norm();
            radius = ratio * MAX_WORKSPACE_RADIUS;
            break;
          }
        }
      }
      int opposite_bearing = mod(bearing + 180, 360);
      if (radius < walkspace_.at(bearing)) {
        walkspace_[bearing] = radius;
        walkspace_[opposite_bearing] = radius;
      }
    }
  }
  walkspace_[360] = walkspace_[0];
  generateLimits();
}

void Robot::generateLimits() {  // walk_controller.cpp:231 (set_limits form, step = step_)
  const StepCycle& step = step_;
  int base_step_period = params_.stance_phase + params_.swing_phase;
  int normaliser = step.period_ / base_step_period;
  int base_step_offset = int(params_.phase_offset * normaliser);

  max_linear_speed_.clear();
  max_linear_acceleration_.clear();
  max_angular_speed_.clear();
  max_angular_acceleration_.clear();

  int max_stance_extension = 0;
  for (int li = 0; li < leg_count_; ++li) {
    int multiplier = params_.offset_multiplier[li];
    LegStepper& leg_stepper = legs[li].stepper;
    int step_offset = (base_step_offset * multiplier) % step.period_;
    leg_stepper.phase_offset_ = step_offset;
    if (step_offset > step.swing_start_ && step_offset < step.swing_end_)
      max_stance_extension = std::max(max_stance_extension, step.swing_end_ - step_offset);
  }
  double time_to_max_stride = (max_stance_extension + step.stance_period_ + step.swing_period_) * time_delta_;

  for (auto it = walkspace_.begin(); it != walkspace_.end(); ++it) {
    double walkspace_radius = it->second;
    double on_ground_ratio = double(step.stance_period_) / step.period_;
    double max_speed = (walkspace_radius * 2.0) / (on_ground_ratio / step.frequency_);
    double max_acceleration = max_speed / time_to_max_stride;

    double stance_overshoot = 0;
    for (int li = 0; li < leg_count_; ++li) {
      LegStepper& leg_stepper = legs[li].stepper;
      double step_offset = leg_stepper.phase_offset_;
      double t = step_offset * time_delta_;
      double time_to_swing_end = time_to_max_stride - t;
      double v0 = max_acceleration * time_to_swing_end;
      double stride_length = v0 * (on_ground_ratio / step.frequency_);
      double d0 = -stride_length / 2.0;
      double d1 = d0 + v0 * t + 0.5 * max_acceleration * sqr(t);
      double d2 = max_speed * (step.stance_period_ * time_delta_ - t);
      stance_overshoot = std::max(stance_overshoot, d1 + d2 - walkspace_radius);
    }
    double swing_overshoot = 0.5 * max_speed * step.swing_period_ / (2.0 * step.period_ * step.frequency_);
    double scaled_walkspace_radius = (walkspace_radius / (walkspace_radius + stance_overshoot + swing_overshoot)) * walkspace_radius;

    LegStepper& reference_leg_stepper = legs[0].stepper;
    double x_position = reference_leg_stepper.default_tip_pose_.position_[0];
    double y_position = reference_leg_stepper.default_tip_pose_.position_[1];
    double stance_radius = std::sqrt(x_position * x_position + y_position * y_position);

    double max_linear_speed = (scaled_walkspace_radius * 2.0) / (on_ground_ratio / step.frequency_);
    double max_linear_acceleration = max_linear_speed / time_to_max_stride;
    double max_angular_speed = max_linear_speed / stance_radius;
    double max_angular_acceleration = max_angular_speed / time_to_max_stride;

    if (walkspace_radius == 0.0) {
      max_linear_speed = 0.0;
      max_linear_acceleration = UNASSIGNED_VALUE;
      max_angular_speed = 0.0;
      max_angular_acceleration = UNASSIGNED_VALUE;
    }
    max_linear_speed_.insert(LimitMap::value_type(it->first, max_linear_speed));
    max_linear_acceleration_.insert(LimitMap::value_type(it->first, max_linear_acceleration));
    max_angular_speed_.insert(LimitMap::value_type(it->first, max_angular_speed));
    max_angular_acceleration_.insert(LimitMap::value_type(it->first, max_angular_acceleration));
  }
}

StepCycle Robot::generateStepCycle(bool set_step_cycle) {  // walk_controller.cpp:365
  StepCycle step;
  step.stance_end_ = static_cast<int>(params_.stance_phase * 0.5);
  step.swing_start_ = step.stance_end_;
  step.swing_end_ = step.swing_start_ + params_.swing_phase;
  step.stance_start_ = step.swing_end_;

  int base_step_period = params_.stance_phase + params_.swing_phase;
  double swing_ratio = double(params_.swing_phase) / double(base_step_period);
  double raw_step_period = ((1.0 / params_.step_frequency) / time_delta_) / swing_ratio;
  step.period_ = roundToEvenInt(raw_step_period / base_step_period) * base_step_period;

  step.frequency_ = 1.0 / (step.period_ * time_delta_);
  int normaliser = step.period_ / base_step_period;
  step.stance_end_ *= normaliser;
  step.swing_start_ *= normaliser;
  step.swing_end_ *= normaliser;
  step.stance_start_ *= normaliser;

  step.stance_period_ = mod(step.stance_end_ - step.stance_start_, step.period_);
  step.swing_period_ = step.swing_end_ - step.swing_start_;
  assert(step.stance_period_ % 2 == 0);
  assert(step.swing_period_ % 2 == 0);

  if (set_step_cycle) step_ = step;  // (the MOVING-state updatePhase branch is only reachable on a live gait change)
  return step;
}

double Robot::getLimit(const double lin[2], double ang, const LimitMap& limit) {  // walk_controller.cpp:414
  double min_limit = UNASSIGNED_VALUE;
  for (int li = 0; li < leg_count_; ++li) {
    LegStepper& leg_stepper = legs[li].stepper;
    Vec3 tip_position = leg_stepper.current_tip_pose_.position_;
    double rotation_normal[2] = {-tip_position[1], tip_position[0]};
    double stride_vector[2] = {lin[0] + ang * rotation_normal[0], lin[1] + ang * rotation_normal[1]};
    int bearing = mod(roundToInt(radiansToDegrees(std::atan2(stride_vector[1], stride_vector[0]))), 360);
    int upper_bound = limit.lower_bound(bearing)->first;
    int lower_bound = mod(upper_bound - BEARING_STEP, 360);
    bearing += (bearing < lower_bound) ? 360 : 0;
    upper_bound += (upper_bound < lower_bound) ? 360 : 0;
    double control_input = (bearing - lower_bound) / (upper_bound - lower_bound);  // int / int (trap 1)
    double limit_interpolation = interpolate(limit.at(lower_bound), limit.at(mod(upper_bound, 360)), control_input);
    min_limit = std::min(min_limit, limit_interpolation);
  }
  return min_limit;
}

void Robot::updateWalk(const double linear_velocity_input[2], double angular_velocity_input) {  // :440
  double new_linear_velocity[2];
  double new_angular_velocity = 0;
  const double input_norm =
      std::sqrt(linear_velocity_input[0] * linear_velocity_input[0] + linear_velocity_input[1] * linear_velocity_input[1]);

  double max_linear_speed = getLimit(linear_velocity_input, angular_velocity_input, max_linear_speed_);
  double max_angular_speed = getLimit(linear_velocity_input, angular_velocity_input, max_angular_speed_);
  double max_linear_acceleration = getLimit(linear_velocity_input, angular_velocity_input, max_linear_acceleration_);
  double max_angular_acceleration = getLimit(linear_velocity_input, angular_velocity_input, max_angular_acceleration_);

  if (walk_state_ != STOPPING) {
    if (params_.velocity_input_mode == SHC_VELOCITY_THROTTLE) {
      // clamped(vector, 1.0): value.norm() > magnitude ? value * (magnitude / value.norm()) : value
      double cl[2] = {linear_velocity_input[0], linear_velocity_input[1]};
      if (input_norm > 1.0) {
        cl[0] = linear_velocity_input[0] * (1.0 / input_norm);
        cl[1] = linear_velocity_input[1] * (1.0 / input_norm);
      }
      new_linear_velocity[0] = cl[0] * max_linear_speed;
      new_linear_velocity[1] = cl[1] * max_linear_speed;
      new_angular_velocity = clamped(angular_velocity_input, -1.0, 1.0) * max_angular_speed;
      new_linear_velocity[0] *= (1.0 - std::fabs(angular_velocity_input));
      new_linear_velocity[1] *= (1.0 - std::fabs(angular_velocity_input));
    } else {  // "real"
      double cl[2] = {linear_velocity_input[0], linear_velocity_input[1]};
      if (input_norm > max_linear_speed) {
        cl[0] = linear_velocity_input[0] * (max_linear_speed / input_norm);
        cl[1] = linear_velocity_input[1] * (max_linear_speed / input_norm);
      }
      new_linear_velocity[0] = cl[0];
      new_linear_velocity[1] = cl[1];
      new_angular_velocity = clamped(angular_velocity_input, -max_angular_speed, max_angular_speed);
      double scale = (max_angular_speed != 0.0 ? (1.0 - std::fabs(new_angular_velocity / max_angular_speed)) : 0.0);
      new_linear_velocity[0] *= scale;
      new_linear_velocity[1] *= scale;
    }
  } else {
    new_linear_velocity[0] = new_linear_velocity[1] = 0.0;
    new_angular_velocity = 0.0;
  }

  bool has_velocity_command = input_norm || angular_velocity_input;

  for (int li = 0; li < leg_count_; ++li)
    if (legs[li].leg_state_ != WALKING) return;

  double linear_acceleration[2] = {new_linear_velocity[0] - desired_linear_velocity_[0],
                                   new_linear_velocity[1] - desired_linear_velocity_[1]};
  double la_norm = std::sqrt(linear_acceleration[0] * linear_acceleration[0] + linear_acceleration[1] * linear_acceleration[1]);
  if (la_norm < max_linear_acceleration * time_delta_) {
    desired_linear_velocity_[0] += linear_acceleration[0];
    desired_linear_velocity_[1] += linear_acceleration[1];
  } else {
    // linear_acceleration.normalized() * max_linear_acceleration * time_delta_
    double n2 = linear_acceleration[0] * linear_acceleration[0] + linear_acceleration[1] * linear_acceleration[1];
    double nx = linear_acceleration[0], ny = linear_acceleration[1];
    if (n2 > 0.0) {
      nx /= std::sqrt(n2);
      ny /= std::sqrt(n2);
    }
    desired_linear_velocity_[0] += nx * max_linear_acceleration * time_delta_;
    desired_linear_velocity_[1] += ny * max_linear_acceleration * time_delta_;
  }

  double angular_acceleration = new_angular_velocity - desired_angular_velocity_;
  if (std::fabs(angular_acceleration) < max_angular_acceleration * time_delta_) {
    desired_angular_velocity_ += angular_acceleration;
  } else {
    desired_angular_velocity_ += sign(angular_acceleration) * max_angular_acceleration * time_delta_;
  }

  int leg_count = leg_count_;
  if (walk_state_ == STOPPED && has_velocity_command) {
    walk_state_ = STARTING;
    for (int li = 0; li < leg_count_; ++li) {
      LegStepper& leg_stepper = legs[li].stepper;
      leg_stepper.at_correct_phase_ = false;
      leg_stepper.completed_first_step_ = false;
      leg_stepper.step_state_ = STANCE;
      leg_stepper.phase_ = leg_stepper.phase_offset_;
      leg_stepper.updateStepState();
    }
    return;
  } else if (walk_state_ == STARTING && legs_at_correct_phase_ == leg_count && legs_completed_first_step_ == leg_count) {
    legs_at_correct_phase_ = 0;
    legs_completed_first_step_ = 0;
    walk_state_ = MOVING;
  } else if (walk_state_ == MOVING && !has_velocity_command) {
    walk_state_ = STOPPING;
  } else if (walk_state_ == STOPPING && legs_at_correct_phase_ == leg_count && pose_state_ == POSING_COMPLETE) {
    legs_at_correct_phase_ = 0;
    walk_state_ = STOPPED;
  }

  for (int li = 0; li < leg_count_; ++li) {
    Leg& leg = legs[li];
    LegStepper& leg_stepper = leg.stepper;
    if (walk_state_ == STARTING) {
      if (legs_at_correct_phase_ == leg_count) {
        if (leg_stepper.phase_ == step_.swing_end_ && !leg_stepper.completed_first_step_) {
          leg_stepper.completed_first_step_ = true;
          legs_completed_first_step_++;
        }
      }
      if (!leg_stepper.at_correct_phase_) {
        if (leg_stepper.phase_offset_ > step_.swing_start_ && leg_stepper.phase_offset_ < step_.swing_end_ &&
            leg_stepper.phase_ != step_.swing_end_) {
          leg_stepper.step_state_ = FORCE_STANCE;
        } else {
          legs_at_correct_phase_++;
          leg_stepper.at_correct_phase_ = true;
        }
      }
    } else if (walk_state_ == MOVING) {
      leg_stepper.at_correct_phase_ = false;
    } else if (walk_state_ == STOPPING) {
      bool zero_body_velocity = leg_stepper.stride_vector_.norm() == 0;
      Vec3 walk_plane_normal = leg_stepper.walk_plane_normal_;
      Vec3 error = (leg_stepper.current_tip_pose_.position_ - leg_stepper.target_tip_pose_.position_);
      error = getRejection(error, walk_plane_normal);
      bool at_target_tip_position = (error.norm() < TIP_TOLERANCE);
      if (zero_body_velocity && !leg_stepper.at_correct_phase_ && leg_stepper.phase_ == step_.swing_end_) {
        if (at_target_tip_position || return_to_default_attempted_) {
          return_to_default_attempted_ = false;
          leg_stepper.updateDefaultTipPosition();
          leg_stepper.step_state_ = FORCE_STOP;
          leg_stepper.at_correct_phase_ = true;
          legs_at_correct_phase_++;
        } else {
          return_to_default_attempted_ = true;
        }
      }
    } else if (walk_state_ == STOPPED) {
      leg_stepper.step_state_ = FORCE_STOP;
      leg_stepper.phase_ = 0;
    }

    if (leg.leg_state_ == WALKING) {
      leg_stepper.updateTipPosition();
      leg_stepper.updateTipRotation();
      leg_stepper.iteratePhase();
    }
  }
  updateWalkPlane();
  odometry_ideal_ = odometry_ideal_.addPose(calculateOdometry(time_delta_));
}

void Robot::updateWalkPlane() {  // walk_controller.cpp:748
  if (leg_count_ >= 3) {
    MatX A(leg_count_, 3), B(leg_count_, 1);
    for (int li = 0; li < leg_count_; ++li) {
      const Vec3& p = legs[li].stepper.default_tip_pose_.position_;
      A(li, 0) = p[0];
      A(li, 1) = p[1];
      A(li, 2) = 1.0;
      B(li, 0) = p[2];
    }
    // A is up to 8x3; MatX holds <= 36 entries, so form A^T A and A^T explicitly.
    MatX At = A.transpose();
    MatX pseudo_inverse_A = (At * A).inverse() * At;
    MatX wp = pseudo_inverse_A * B;
    walk_plane_ = Vec3(wp(0, 0), wp(1, 0), wp(2, 0));
    walk_plane_normal_ = Vec3(-walk_plane_[0], -walk_plane_[1], 1.0).normalized();
  } else {
    walk_plane_ = Vec3(0, 0, 0);
    walk_plane_normal_ = UnitZ();
  }
}

Pose Robot::calculateOdometry(double time_period) {  // walk_controller.cpp:783
  Vec3 desired_linear_velocity(desired_linear_velocity_[0], desired_linear_velocity_[1], 0);
  Vec3 position_delta = desired_linear_velocity * time_period;
  Quat rotation_delta = quatFromAngleAxis(desired_angular_velocity_ * time_period, UnitZ());
  return Pose(position_delta, rotation_delta);
}

// ---------------------------------------------------------------------------------------------------------------
// LegStepper
// ---------------------------------------------------------------------------------------------------------------
void LegStepper::construct(Robot* r, Leg* leg, const Pose& identity_tip_pose) {  // walk_controller.cpp:795
  *this = LegStepper();
  robot = r;
  leg_ = leg;
  identity_tip_pose_ = identity_tip_pose;
  default_tip_pose_ = identity_tip_pose;
  current_tip_pose_ = default_tip_pose_;
  origin_tip_pose_ = current_tip_pose_;
  target_tip_pose_ = default_tip_pose_;
  walk_plane_ = Vec3(0, 0, 0);
  walk_plane_normal_ = UnitZ();
  stride_vector_ = Vec3(0, 0, 0);
  current_tip_velocity_ = Vec3(0, 0, 0);
  swing_origin_tip_position_ = default_tip_pose_.position_;
  stance_origin_tip_position_ = default_tip_pose_.position_;
  swing_clearance_ = Vec3(0.0, 0.0, r->params_.swing_height);
}

void LegStepper::iteratePhase() {  // walk_controller.cpp:871
  const StepCycle step = robot->step_;
  phase_ = (phase_ + 1) % (step.period_);
  updateStepState();
  step_progress_ = double(phase_) / step.period_;
  if (step_state_ == SWING) {
    swing_progress_ = double(phase_ - step.swing_start_ + 1) / double(step.swing_end_ - step.swing_start_);
    swing_progress_ = clamped(swing_progress_, 0.0, 1.0);
    stance_progress_ = -1.0;
  } else if (step_state_ == STANCE) {
    stance_progress_ = double(mod(phase_ + (step.period_ - step.stance_start_), step.period_) + 1) /
                       double(mod(step.stance_end_ - step.stance_start_, step.period_));
    stance_progress_ = clamped(stance_progress_, 0.0, 1.0);
    swing_progress_ = -1.0;
  } else if (step_state_ == FORCE_STOP) {
    stance_progress_ = 0.0;
    swing_progress_ = -1.0;
  }
}

void LegStepper::updateStepState() {  // walk_controller.cpp:901
  const StepCycle step = robot->step_;
  if (step_state_ == FORCE_STOP) {
    return;
  } else if (phase_ >= step.swing_start_ && phase_ < step.swing_end_ && step_state_ != FORCE_STANCE) {
    step_state_ = SWING;
  } else if (phase_ < step.stance_end_ || phase_ >= step.stance_start_) {
    step_state_ = STANCE;
  }
}

void LegStepper::updateStride() {  // walk_controller.cpp:921
  walk_plane_ = robot->walk_plane_;
  walk_plane_normal_ = robot->walk_plane_normal_;
  Vec3 stride_vector_linear(robot->desired_linear_velocity_[0], robot->desired_linear_velocity_[1], 0.0);
  Vec3 radius = getRejection(current_tip_pose_.position_, UnitZ());
  Vec3 angular_velocity = robot->desired_angular_velocity_ * UnitZ();
  Vec3 stride_vector_angular = angular_velocity.cross(radius);
  stride_vector_ = stride_vector_linear + stride_vector_angular;
  const StepCycle step = robot->step_;
  double on_ground_ratio = double(step.stance_period_) / step.period_;
  stride_vector_ *= (on_ground_ratio / step.frequency_);
  swing_clearance_ = robot->params_.swing_height * walk_plane_normal_.normalized();
}

Vec3 LegStepper::calculateStanceSpanChange() {  // walk_controller.cpp:949
  Vec3 default_shift = default_tip_pose_.position_ - identity_tip_pose_.position_;
  double target_workplane_height = setPrecision(default_shift[2], 3);
  Workspace workspace = leg_->workspace_;
  Workspace::iterator upper_bound_it = workspace.upper_bound(target_workplane_height);
  double stance_span_modifier = robot->params_.stance_span_modifier;
  bool positive_y_axis = (UnitY().dot(identity_tip_pose_.position_) > 0.0);
  int bearing = (positive_y_axis ^ (stance_span_modifier > 0.0)) ? 270 : 90;
  stance_span_modifier *= (positive_y_axis ? 1.0 : -1.0);
  double radius = 0.0;
  if (workspace.size() == 1) {
    // The reference dereferences prev(upper_bound) before this test; with a one-plane workspace those values are
    // never used (walk_controller.cpp:957-974), so they are not formed here.
    radius = workspace.at(0.0).at(bearing);
  } else {
    Workspace::iterator lower_bound_it = std::prev(upper_bound_it);
    double upper_workplane_height = setPrecision(upper_bound_it->first, 3);
    double lower_workplane_height = setPrecision(lower_bound_it->first, 3);
    LimitMap upper_workplane = upper_bound_it->second;
    LimitMap lower_workplane = lower_bound_it->second;
    double i = (target_workplane_height - lower_workplane_height) / (upper_workplane_height - lower_workplane_height);
    radius = lower_workplane.at(bearing) * (1.0 - i) + upper_workplane.at(bearing) * i;
  }
  return Vec3(0.0, radius * stance_span_modifier, 0.0);
}

void LegStepper::updateDefaultTipPosition() {  // walk_controller.cpp:984
  if (external_default_.defined_) {  // externally requested default, transformed by the robot's movement since the request
    Pose new_default_tip_pose = external_default_.pose_.removePose(external_default_.transform_);
    default_tip_pose_ = new_default_tip_pose;
    return;
  }
  Vec3 identity_tip_position = identity_tip_pose_.position_;
  identity_tip_position += calculateStanceSpanChange();
  identity_tip_position = robot->default_pose_model_.transformVector(identity_tip_position);
  Vec3 identity_to_stance_origin = stance_origin_tip_position_ - identity_tip_position;
  Vec3 projection_to_walk_plane = getProjection(identity_to_stance_origin, walk_plane_normal_);
  Pose new_default_tip_pose(identity_tip_position + projection_to_walk_plane, UndefinedRotation());
  default_tip_pose_ = new_default_tip_pose;
}

void LegStepper::updateTipPosition() {  // walk_controller.cpp:1018
  const shc_config& params = robot->params_;
  bool rough_terrain_mode = params.rough_terrain_mode;
  bool force_normal_touchdown = params.force_normal_touchdown;
  double time_delta = robot->time_delta_;
  const StepCycle step = robot->step_;

  bool standard_stance_period = (step_state_ == SWING || completed_first_step_);
  int modified_stance_start = standard_stance_period ? step.stance_start_ : phase_offset_;
  int modified_stance_period = mod(step.stance_end_ - modified_stance_start, step.period_);
  if (step.stance_end_ == modified_stance_start) modified_stance_period = step.period_;

  int swing_iterations = int((double(step.swing_period_) / step.period_) / (step.frequency_ * time_delta));
  swing_iterations = roundToEvenInt(swing_iterations);
  swing_delta_t_ = 1.0 / (swing_iterations / 2.0);

  int stance_iterations = int((double(modified_stance_period) / step.period_) / (step.frequency_ * time_delta));
  stance_delta_t_ = 1.0 / stance_iterations;

  target_tip_pose_.position_ = default_tip_pose_.position_ + 0.5 * stride_vector_;

  if (step_state_ == SWING) {
    updateStride();
    int iteration = phase_ - step.swing_start_ + 1;
    bool first_half = iteration <= swing_iterations / 2;
    if (iteration == 1) {
      swing_origin_tip_position_ = current_tip_pose_.position_;
      swing_origin_tip_velocity_ = current_tip_velocity_;
      if (rough_terrain_mode) updateDefaultTipPosition();
    }
    // Update the target to the externally requested pose or to meet the step surface (walk_controller.cpp:1065-1107)
    if (rough_terrain_mode) {
      if (external_target_.defined_) {
        target_tip_pose_ = external_target_.pose_.removePose(external_target_.transform_);
        swing_clearance_ = swing_clearance_.normalized() * external_target_.swing_clearance_;
        if (external_target_.odom_ideal_frame_) {  // add lead to compensate for the moving target
          double time_to_swing_end = (swing_iterations - iteration) * time_delta;
          Vec3 target_lead = robot->calculateOdometry(time_to_swing_end).position_;
          target_tip_pose_.position_ -= target_lead;
        }
      } else if (touchdown_detection_) {
        Pose step_plane_pose = leg_->step_plane_pose_;
        if (step_plane_pose != Pose::Undefined()) {  // proactive: the step plane is known
          Vec3 step_plane_position = step_plane_pose.position_ - leg_->current_tip_pose_.position_;
          Vec3 target_tip_position = current_tip_pose_.position_ + step_plane_position;
          Vec3 difference = target_tip_position - target_tip_pose_.position_;
          target_tip_pose_.position_ += getProjection(difference, walk_plane_normal_);
        } else {  // reactive: reach down by the step depth and rely on contact detection
          target_tip_pose_.position_ -= params.step_depth * UnitZ();
        }
      }
    }
    bool ground_contact = (leg_->step_plane_pose_ != Pose::Undefined() && rough_terrain_mode);
    generatePrimarySwingControlNodes();
    generateSecondarySwingControlNodes(!first_half && ground_contact);
    if (force_normal_touchdown && !ground_contact) forceNormalTouchdown();

    Vec3 delta_pos(0, 0, 0);
    double time_input = 0;
    if (first_half) {
      time_input = swing_delta_t_ * iteration;
      delta_pos = swing_delta_t_ * quarticBezierDot(swing_1_nodes_, time_input);
    } else {
      time_input = swing_delta_t_ * (iteration - swing_iterations / 2);
      delta_pos = swing_delta_t_ * quarticBezierDot(swing_2_nodes_, time_input);
    }
    current_tip_pose_.position_ += delta_pos;
    current_tip_velocity_ = delta_pos / time_delta;
  } else if (step_state_ == STANCE || step_state_ == FORCE_STANCE) {
    updateStride();
    int iteration = mod(phase_ + (step.period_ - modified_stance_start), step.period_) + 1;
    if (iteration == 1) {
      stance_origin_tip_position_ = current_tip_pose_.position_;
      external_target_.defined_ = false;  // reset external target after every swing period (:1159)
      if (rough_terrain_mode) updateDefaultTipPosition();
    }
    double stride_scaler = double(modified_stance_period) / (mod(step.stance_end_ - step.stance_start_, step.period_));
    generateStanceControlNodes(stride_scaler);
    double time_input = iteration * stance_delta_t_;
    Vec3 delta_pos = stance_delta_t_ * quarticBezierDot(stance_nodes_, time_input);
    current_tip_pose_.position_ += delta_pos;
    current_tip_velocity_ = delta_pos / time_delta;
  }
}

void LegStepper::updateTipRotation() {  // walk_controller.cpp:1193
  if (leg_->joint_count_ > 3 && (stance_progress_ >= 0.0 || swing_progress_ >= 0.5)) {
    if (robot->params_.gravity_aligned_tips && isApproxQuat(target_tip_pose_.rotation_, UndefinedRotation())) {
      target_tip_pose_.rotation_ = fromTwoVectors(UnitX(), robot->estimateGravity());
    }
    if (isApproxQuat(target_tip_pose_.rotation_, UndefinedRotation())) {
      current_tip_pose_.rotation_ = target_tip_pose_.rotation_;
    } else {
      current_tip_pose_.rotation_ = correctRotation(target_tip_pose_.rotation_, origin_tip_pose_.rotation_);
      if (swing_progress_ >= 0.5) {
        double c = smoothStep(std::min(1.0, 2.0 * (swing_progress_ - 0.5)));
        Vec3 origin_tip_direction = origin_tip_pose_.rotation_.transformVector(UnitX());
        Vec3 target_tip_direction = target_tip_pose_.rotation_.transformVector(UnitX());
        Vec3 new_tip_direction = interpolate(origin_tip_direction, target_tip_direction, c);
        Quat new_tip_rotation = fromTwoVectors(UnitX(), new_tip_direction.normalized());
        current_tip_pose_.rotation_ = correctRotation(new_tip_rotation, current_tip_pose_.rotation_);
      }
    }
  } else {
    origin_tip_pose_.rotation_ = leg_->current_tip_pose_.rotation_;
    current_tip_pose_.rotation_ = UndefinedRotation();
  }
}

void LegStepper::generatePrimarySwingControlNodes() {  // walk_controller.cpp:1238
  Vec3 mid_tip_position = (swing_origin_tip_position_ + target_tip_pose_.position_) / 2.0;
  mid_tip_position[2] = std::max(swing_origin_tip_position_[2], target_tip_pose_.position_[2]);
  mid_tip_position += swing_clearance_;
  double mid_lateral_shift = robot->params_.swing_width;
  bool positive_y_axis = (UnitY().dot(identity_tip_pose_.position_) > 0.0);
  mid_tip_position[1] += positive_y_axis ? mid_lateral_shift : -mid_lateral_shift;
  Vec3 stance_node_seperation = 0.25 * swing_origin_tip_velocity_ * (robot->time_delta_ / swing_delta_t_);
  swing_1_nodes_[0] = swing_origin_tip_position_;
  swing_1_nodes_[1] = swing_origin_tip_position_ + stance_node_seperation;
  swing_1_nodes_[2] = swing_origin_tip_position_ + 2.0 * stance_node_seperation;
  swing_1_nodes_[3] = (mid_tip_position + swing_1_nodes_[2]) / 2.0;
  swing_1_nodes_[3][2] = mid_tip_position[2];
  swing_1_nodes_[4] = mid_tip_position;
}

void LegStepper::generateSecondarySwingControlNodes(bool ground_contact) {  // walk_controller.cpp:1265
  Vec3 final_tip_velocity = -stride_vector_ * (stance_delta_t_ / robot->time_delta_);
  Vec3 stance_node_seperation = 0.25 * final_tip_velocity * (robot->time_delta_ / swing_delta_t_);
  swing_2_nodes_[0] = swing_1_nodes_[4];
  swing_2_nodes_[1] = swing_1_nodes_[4] - (swing_1_nodes_[3] - swing_1_nodes_[4]);
  swing_2_nodes_[2] = target_tip_pose_.position_ - 2.0 * stance_node_seperation;
  swing_2_nodes_[3] = target_tip_pose_.position_ - stance_node_seperation;
  swing_2_nodes_[4] = target_tip_pose_.position_;
  if (ground_contact) {
    swing_2_nodes_[0] = current_tip_pose_.position_ + 0.0 * stance_node_seperation;
    swing_2_nodes_[1] = current_tip_pose_.position_ + 1.0 * stance_node_seperation;
    swing_2_nodes_[2] = current_tip_pose_.position_ + 2.0 * stance_node_seperation;
    swing_2_nodes_[3] = current_tip_pose_.position_ + 3.0 * stance_node_seperation;
    swing_2_nodes_[4] = current_tip_pose_.position_ + 4.0 * stance_node_seperation;
  }
}

void LegStepper::generateStanceControlNodes(double stride_scaler) {  // walk_controller.cpp:1295
  Vec3 stance_node_seperation = -stride_vector_ * stride_scaler * 0.25;
  stance_nodes_[0] = stance_origin_tip_position_ + 0.0 * stance_node_seperation;
  stance_nodes_[1] = stance_origin_tip_position_ + 1.0 * stance_node_seperation;
  stance_nodes_[2] = stance_origin_tip_position_ + 2.0 * stance_node_seperation;
  stance_nodes_[3] = stance_origin_tip_position_ + 3.0 * stance_node_seperation;
  stance_nodes_[4] = stance_origin_tip_position_ + 4.0 * stance_node_seperation;
}

void LegStepper::forceNormalTouchdown() {  // walk_controller.cpp:1314
  Vec3 final_tip_velocity = -stride_vector_ * (stance_delta_t_ / robot->time_delta_);
  Vec3 stance_node_seperation = 0.25 * final_tip_velocity * (robot->time_delta_ / swing_delta_t_);
  Vec3 bezier_target = target_tip_pose_.position_;
  Vec3 bezier_origin = target_tip_pose_.position_ - 4.0 * stance_node_seperation;
  bezier_origin[2] = std::max(swing_origin_tip_position_[2], target_tip_pose_.position_[2]);
  bezier_origin += swing_clearance_;
  swing_1_nodes_[4] = bezier_origin;
  swing_2_nodes_[0] = bezier_origin;
  swing_2_nodes_[2] = bezier_target - 2.0 * stance_node_seperation;
  swing_1_nodes_[3] = swing_2_nodes_[0] - (swing_2_nodes_[2] - bezier_origin) / 2.0;
  swing_2_nodes_[1] = swing_2_nodes_[0] + (swing_2_nodes_[2] - bezier_origin) / 2.0;
}

}  // namespace shc_oracle
