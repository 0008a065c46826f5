// shc_oracle_model.cpp — TEST INFRASTRUCTURE ONLY.  Restates /root/reference/src/model.cpp (+ model.h inline
// chain helpers) in IEEE double for the parity oracle.  PINNED to the reference's own code (oracle/_ref, tests/test_reference_pin.py; see shc_oracle.hpp).
#include "shc_oracle.hpp"

namespace shc_oracle {

// ---------------------------------------------------------------------------------------------------------------
// Model
// ---------------------------------------------------------------------------------------------------------------
Robot::Robot(const shc_config& cfg) : params_(cfg) {
  // Model::Model (model.cpp:16-27)
  leg_count_ = cfg.leg_count;
  time_delta_ = cfg.time_delta;
  current_pose_ = Pose::Identity();
  default_pose_model_ = Pose::Identity();
  imu_data_.orientation = UndefinedRotation();
  imu_data_.linear_acceleration = Vec3(0, 0, 0);
  imu_data_.angular_velocity = Vec3(0, 0, 0);
  // Model::generate (model.cpp:44-62)
  for (int i = 0; i < leg_count_; ++i) legs[i].construct(this, i);
  // PoseController::PoseController (pose_controller.cpp:14-26) -> resetAllPosing (pose_controller.h:123-136)
  manual_pose_ = auto_pose_ = imu_pose_ = inclination_pose_ = default_pose_ = tip_align_pose_ = Pose::Identity();
  origin_tip_align_pose_ = tip_align_pose_;
  walk_plane_pose_ = Pose::Identity();
  origin_walk_plane_pose_ = walk_plane_pose_;
}

void Robot::initLegs(bool use_default) {  // model.cpp:66
  for (int i = 0; i < leg_count_; ++i) legs[i].init(use_default);
}

void Robot::updateDefaultConfiguration() {  // model.cpp:108
  for (int i = 0; i < leg_count_; ++i) legs[i].updateDefaultConfiguration();
}

void Robot::generateWorkspaces() {  // model.cpp:120
  // The reference copies the whole model and searches on the copy; a copied leg only ever reads the (identical)
  // body pose and parameters of its model, so copying each leg and keeping the parent pointer is equivalent.
  for (int i = 0; i < leg_count_; ++i) {
    Leg search_leg = legs[i];  // Leg copy-ctor + generate(reference_leg) (model.cpp:192-282)
    search_leg.workspace_.clear();
    search_leg.stepper.leg_ = &search_leg;
    search_leg.poser.leg_ = &search_leg;
    search_leg.init(true);  // search_model->initLegs(true) (model.cpp:126)
    legs[i].workspace_ = search_leg.generateWorkspace();
  }
}

void Robot::updateModel() {  // model.cpp:142
  for (int i = 0; i < leg_count_; ++i) {
    legs[i].setDesiredTipPose();
    legs[i].last_ik_result_ = legs[i].applyIK();
  }
}

Vec3 Robot::estimateGravity() {  // model.cpp:156 (reads imu_data_ directly, not getImuData())
  Vec3 euler = quaternionToEulerAngles(imu_data_.orientation);
  Vec3 gravity(0, 0, GRAVITY_ACCELERATION);
  gravity = angleAxisRotate(-euler[1], UnitY(), gravity);
  gravity = angleAxisRotate(-euler[0], UnitX(), gravity);
  return gravity;
}

ImuData Robot::getImuData() const {  // model.h:132
  ImuData d = imu_data_;
  if (isApproxQuat(imu_data_.orientation, UndefinedRotation())) d.orientation = Quat::Identity();
  return d;
}

void Robot::setImuData(const Quat& q, const Vec3& acc, const Vec3& gyro) {  // model.h:146
  imu_data_.orientation = q.normalized();
  imu_data_.linear_acceleration = acc;
  imu_data_.angular_velocity = gyro;
}

// ---------------------------------------------------------------------------------------------------------------
// Leg
// ---------------------------------------------------------------------------------------------------------------
void Leg::construct(Robot* r, int id) {  // Leg::Leg (model.cpp:169-188) + Leg::generate (model.cpp:221-239)
  robot = r;
  id_number_ = id;
  const shc_config& p = r->params_;
  joint_count_ = p.joint_count;
  leg_state_ = WALKING;
  admittance_delta_ = Vec3(0, 0, 0);
  admittance_state_[0] = admittance_state_[1] = 0.0;
  desired_tip_pose_ = Pose::Undefined();
  current_tip_pose_ = Pose::Undefined();
  step_plane_pose_ = Pose::Undefined();
  group_ = id % 2;
  for (int i = 0; i <= joint_count_; ++i) {  // Link::Link (model.cpp:992)
    links[i].d = p.link_d[id][i];
    links[i].theta = p.link_theta[id][i];
    links[i].r = p.link_r[id][i];
    links[i].alpha = p.link_alpha[id][i];
  }
  for (int i = 1; i <= joint_count_; ++i) {  // Joint::Joint (model.cpp:1026-1073)
    Joint& j = joints[i];
    j.min_position_ = p.joint_min[id][i - 1];
    j.max_position_ = p.joint_max[id][i - 1];
    j.offset_ = p.joint_offset[id][i - 1];
    j.max_angular_speed_ = p.joint_max_vel[id][i - 1];
    j.packed_position_ = p.joint_packed[id][i - 1];
    j.unpacked_position_ = p.joint_unpacked[id][i - 1];
    j.default_position_ = clamped(0.0, j.min_position_, j.max_position_);
    const Link& ref = links[i - 1];
    j.current_transform_ = createDHMatrix(ref.d, ref.theta, ref.r, ref.alpha);
  }
  const Link& last = links[joint_count_];  // Tip::Tip (model.cpp:1119)
  tip_transform_ = createDHMatrix(last.d, last.theta, last.r, last.alpha);
  stepper.robot = r;
  stepper.leg_ = this;
  poser.robot = r;
  poser.leg_ = this;
}

Mat4 Leg::jointTransformFrom(int joint_id, int target_joint_id) const {  // model.h:594
  int next_joint = joint_id - 1;  // reference_link_->actuating_joint_
  if (target_joint_id == next_joint) return joints[joint_id].current_transform_;
  return jointTransformFrom(next_joint, target_joint_id) * joints[joint_id].current_transform_;
}

Mat4 Leg::tipTransformFrom(int target_joint_id) const {  // model.h:674
  int next_joint = joint_count_;
  if (target_joint_id == next_joint) return tip_transform_;
  return jointTransformFrom(next_joint, target_joint_id) * tip_transform_;
}

Pose Leg::jointPoseRobotFrame(int joint_id, const Pose& p) const { return p.transform(jointTransformFrom(joint_id, 0)); }
Pose Leg::jointPoseJointFrame(int joint_id, const Pose& p) const {
  return p.transform(jointTransformFrom(joint_id, 0).inverse());
}
Pose Leg::tipPoseRobotFrame(const Pose& p) const { return p.transform(tipTransformFrom(0)); }

void Leg::init(bool use_default_joint_positions) {  // model.cpp:286
  for (int i = 1; i <= joint_count_; ++i) {
    Joint& j = joints[i];
    if (use_default_joint_positions) {
      j.current_position_ = j.default_position_;
      j.current_velocity_ = j.default_velocity_;
      j.current_effort_ = j.default_effort_;
    }
    j.desired_position_ = j.current_position_;
    j.desired_velocity_ = j.current_velocity_;
    j.desired_effort_ = j.current_effort_;
    j.prev_desired_position_ = j.desired_position_;
  }
  applyFK();
  desired_tip_pose_ = current_tip_pose_;
}

Workspace Leg::generateWorkspace() {  // model.cpp:309-510
  const shc_config& params = robot->params_;
  bool workspace_generation_complete = false;
  bool simple_workspace = !params.rough_terrain_mode;

  LimitMap max_workplane, min_workplane;
  for (int bearing = 0; bearing <= 360; bearing += BEARING_STEP) {
    max_workplane[bearing] = MAX_WORKSPACE_RADIUS;
    min_workplane[bearing] = 0.0;
  }
  workspace_.clear();

  Pose current_pose = robot->current_pose_;
  Vec3 identity_tip_position = current_pose.inverseTransformVector(stepper.identity_tip_pose_.position_);

  if ((identity_tip_position - current_tip_pose_.position_).norm() > IK_TOLERANCE) {
    workspace_[0.0] = min_workplane;
    return workspace_;
  }
  if (simple_workspace) workspace_[0.0] = max_workplane;

  bool found_lower_limit = simple_workspace;
  bool found_upper_limit = simple_workspace;
  double max_plane_height = simple_workspace ? 0.0 : MAX_WORKSPACE_RADIUS;
  double min_plane_height = simple_workspace ? 0.0 : -MAX_WORKSPACE_RADIUS;
  double search_height_delta = MAX_WORKSPACE_RADIUS / WORKSPACE_LAYERS;
  (void)max_plane_height;

  double search_height = 0.0;
  int search_bearing = 0;
  bool within_limits = true;
  int iteration = 1;
  Vec3 origin_tip_position, target_tip_position;
  double distance_from_origin;
  int number_iterations = 1;

  while (true) {
    Pose pose = robot->current_pose_;
    Vec3 identity_tip = pose.inverseTransformVector(stepper.identity_tip_pose_.position_);
    identity_tip[2] += search_height;

    if (iteration == 1) {
      within_limits = true;
      init(true);
      if (!found_lower_limit || !found_upper_limit) {
        number_iterations = roundToInt(MAX_WORKSPACE_RADIUS / MAX_POSITION_DELTA);
        origin_tip_position = identity_tip;
        Vec3 search_limit = (found_lower_limit ? MAX_WORKSPACE_RADIUS : -MAX_WORKSPACE_RADIUS) * UnitZ();
        target_tip_position = identity_tip + search_limit;
      } else if (search_bearing == 0) {
        number_iterations = roundToInt(search_height_delta / MAX_POSITION_DELTA);
        number_iterations = std::max(1, number_iterations);
        origin_tip_position = current_tip_pose_.position_;
        target_tip_position = identity_tip;
      } else {
        number_iterations = roundToInt(MAX_WORKSPACE_RADIUS / MAX_POSITION_DELTA);
        origin_tip_position = identity_tip;
        target_tip_position = origin_tip_position;
        target_tip_position[0] += MAX_WORKSPACE_RADIUS * std::cos(degreesToRadians(search_bearing));
        target_tip_position[1] += MAX_WORKSPACE_RADIUS * std::sin(degreesToRadians(search_bearing));
      }
    }

    double i = double(iteration) / number_iterations;
    Vec3 desired_tip_position = origin_tip_position * (1.0 - i) + target_tip_position * i;
    setDesiredTipPose(Pose(desired_tip_position, UndefinedRotation()));
    double ik_result = applyIK(true);
    distance_from_origin = (current_tip_pose_.position_ - identity_tip).norm();

    within_limits = within_limits && ik_result != 0.0;

    if (within_limits && iteration < number_iterations) {
      iteration++;
    } else {
      iteration = 1;
      if (!found_lower_limit) {
        found_lower_limit = true;
        min_plane_height = -distance_from_origin;
        workspace_[min_plane_height] = min_workplane;
        continue;
      } else if (!found_upper_limit) {
        found_upper_limit = true;
        max_plane_height = distance_from_origin;
        search_height_delta = (max_plane_height - min_plane_height) / WORKSPACE_LAYERS;
        int upper_levels = int(std::fabs(max_plane_height) / search_height_delta);
        search_height = upper_levels * search_height_delta;
        workspace_[max_plane_height] = min_workplane;
        workspace_.insert(Workspace::value_type(search_height, max_workplane));
        continue;
      } else if (search_bearing == 0) {
        updateDefaultConfiguration();
      } else {
        workspace_.at(search_height)[search_bearing] = distance_from_origin;
      }

      if (search_bearing + BEARING_STEP <= 360) {
        search_bearing += BEARING_STEP;
      } else {
        search_bearing = 0;
        workspace_.at(search_height)[0] = workspace_.at(search_height).at(360);
        search_height -= search_height_delta;
        if (search_height >= min_plane_height) {
          workspace_.insert(Workspace::value_type(search_height, max_workplane));
        } else {
          workspace_generation_complete = true;
        }
      }
    }
    if (workspace_generation_complete) return workspace_;
  }
}

LimitMap Leg::getWorkplane(double height) {  // model.cpp:514
  bool within_workspace = (height >= workspace_.begin()->first && height <= workspace_.rbegin()->first);
  if (!within_workspace) {
    return LimitMap();
  } else if (workspace_.size() == 1) {
    return workspace_.at(0.0);
  }
  Workspace::iterator upper_bound_it = workspace_.upper_bound(height);
  Workspace::iterator lower_bound_it = std::prev(upper_bound_it);
  double upper_workplane_height = setPrecision(upper_bound_it->first, 3);
  double lower_workplane_height = setPrecision(lower_bound_it->first, 3);
  LimitMap upper_workplane = upper_bound_it->second;
  LimitMap lower_workplane = lower_bound_it->second;
  double i = (height - lower_workplane_height) / (upper_workplane_height - lower_workplane_height);
  LimitMap workplane = upper_workplane;
  for (auto it = workplane.begin(); it != workplane.end(); ++it) {
    int bearing = it->first;
    double radius = lower_workplane.at(bearing) * (1.0 - i) + upper_workplane.at(bearing) * i;
    workplane[bearing] = within_workspace ? radius : 0.0;
  }
  return workplane;
}

void Leg::updateDefaultConfiguration() {  // model.cpp:593
  for (int i = 1; i <= joint_count_; ++i) joints[i].default_position_ = joints[i].desired_position_;
}

void Leg::setDesiredTipPose(const Pose& tip_pose, bool apply_delta) {  // model.cpp:653
  apply_delta = apply_delta && !(leg_state_ == MANUAL || leg_state_ == WALKING_TO_MANUAL);
  bool use_poser_tip_pose = (Pose::Undefined() == tip_pose);
  desired_tip_pose_ = use_poser_tip_pose ? poser.current_tip_pose_ : tip_pose;
  desired_tip_pose_.position_ += (apply_delta ? admittance_delta_ : Vec3(0, 0, 0));
}

void Leg::calculateTipForce() {  // model.cpp:667
  const int D = joint_count_;
  Vec3 pe = tipTransformFrom(1).col3(3);
  Vec3 z0(0, 0, 1), p0(0, 0, 0);
  MatX jacobian(6, D);
  Vec3 lin = z0.cross(pe - p0);
  for (int r = 0; r < 3; ++r) {
    jacobian(r, 0) = lin[r];
    jacobian(3 + r, 0) = z0[r];
  }
  MatX joint_torques(D, 1);
  joint_torques(0, 0) = joints[1].current_effort_;
  for (int i = 1; i < D; ++i) {
    Mat4 t = jointTransformFrom(i + 1, 1);
    Vec3 l = t.col3(2).cross(pe - t.col3(3));
    Vec3 a = t.col3(2);
    for (int r = 0; r < 3; ++r) {
      jacobian(r, i) = l[r];
      jacobian(3 + r, i) = a[r];
    }
    joint_torques(i, 0) = joints[i + 1].current_effort_;
  }
  MatX identity = MatX::Identity(D);
  MatX transformation = jacobian * ((jacobian.transpose() * jacobian + identity * sqr(DLS_COEFFICIENT)).inverse());
  MatX raw_tip_force_leg_frame = transformation * joint_torques;
  Quat rotation = jointPoseJointFrame(1).rotation_;
  Vec3 raw_tip_force = rotation.transformVector(
      Vec3(raw_tip_force_leg_frame(0, 0), raw_tip_force_leg_frame(1, 0), raw_tip_force_leg_frame(2, 0)));
  double s = 0.15;
  double gain = robot->params_.force_gain;
  tip_force_calculated_[0] = s * raw_tip_force[0] * gain + (1 - s) * tip_force_calculated_[0];
  tip_force_calculated_[1] = s * raw_tip_force[1] * gain + (1 - s) * tip_force_calculated_[1];
  tip_force_calculated_[2] = s * raw_tip_force[2] * gain + (1 - s) * tip_force_calculated_[2];
}

void Leg::solveIK(const double delta_in[6], bool solve_rotation, double* out) {  // model.cpp:726
  const int D = joint_count_;
  Vec3 pe = tipTransformFrom(1).col3(3);
  Vec3 z0(0, 0, 1), p0(0, 0, 0);
  MatX jacobian(6, D);
  Vec3 lin = z0.cross(pe - p0);
  Vec3 ang = solve_rotation ? z0 : Vec3(0, 0, 0);
  for (int r = 0; r < 3; ++r) {
    jacobian(r, 0) = lin[r];
    jacobian(3 + r, 0) = ang[r];
  }
  for (int i = 1; i < D; ++i) {
    Mat4 t = jointTransformFrom(i + 1, 1);
    Vec3 l = t.col3(2).cross(pe - t.col3(3));
    Vec3 a = solve_rotation ? t.col3(2) : Vec3(0, 0, 0);
    for (int r = 0; r < 3; ++r) {
      jacobian(r, i) = l[r];
      jacobian(3 + r, i) = a[r];
    }
  }
  MatX identity = MatX::Identity(6);
  MatX j = jacobian;
  MatX jacobian_inverse = j.transpose() * ((j * j.transpose() + identity * sqr(DLS_COEFFICIENT)).inverse());

  double position_limit_cost = 0.0, velocity_limit_cost = 0.0;
  MatX position_cost_gradient(D, 1), velocity_cost_gradient(D, 1), combined_cost_gradient(D, 1);
  for (int i = 0; i < D; ++i) {
    const Joint& joint = joints[i + 1];
    double joint_position_range = joint.max_position_ - joint.min_position_;
    double position_range_centre = joint.min_position_ + joint_position_range / 2.0;
    if (joint_position_range != 0.0) {
      position_limit_cost += sqr(std::fabs(JOINT_LIMIT_COST_WEIGHT * (joint.desired_position_ - position_range_centre) /
                                           joint_position_range));
      position_cost_gradient(i, 0) = -sqr(JOINT_LIMIT_COST_WEIGHT) * (joint.desired_position_ - position_range_centre) /
                                     sqr(joint_position_range);
    }
    double joint_velocity_range = 2 * joint.max_angular_speed_;
    double velocity_range_centre = 0.0;
    velocity_limit_cost += sqr(std::fabs(JOINT_LIMIT_COST_WEIGHT * (joint.desired_velocity_ - velocity_range_centre) /
                                         joint_velocity_range));
    velocity_cost_gradient(i, 0) = -sqr(JOINT_LIMIT_COST_WEIGHT) * (joint.desired_velocity_ - velocity_range_centre) /
                                   sqr(joint_velocity_range);
  }
  position_cost_gradient = position_cost_gradient * (position_limit_cost == 0.0 ? 0.0 : 1.0 / std::sqrt(position_limit_cost));
  velocity_cost_gradient = velocity_cost_gradient * (velocity_limit_cost == 0.0 ? 0.0 : 1.0 / std::sqrt(velocity_limit_cost));
  combined_cost_gradient = position_cost_gradient * (1.0 - 0.75) + velocity_cost_gradient * 0.75;

  MatX delta(6, 1);
  for (int r = 0; r < 6; ++r) delta(r, 0) = delta_in[r];
  MatX identityD = MatX::Identity(D);
  MatX result = jacobian_inverse * delta + (identityD - jacobian_inverse * j) * combined_cost_gradient;
  for (int i = 0; i < D; ++i) out[i] = result(i, 0);
}

double Leg::updateJointPositions(const double* delta, bool simulation) {  // model.cpp:799
  const shc_config& params = robot->params_;
  double min_limit_proximity = 1.0;
  for (int index = 0; index < joint_count_; ++index) {
    Joint& joint = joints[index + 1];
    joint.desired_velocity_ = delta[index] / robot->time_delta_;
    if (params.clamp_joint_velocities && !simulation) {
      if (std::fabs(joint.desired_velocity_) > joint.max_angular_speed_) {
        double max_velocity = joint.max_angular_speed_;
        joint.desired_velocity_ = clamped(joint.desired_velocity_, -max_velocity, max_velocity);
      }
    }
    joint.prev_desired_position_ = joint.desired_position_;
    joint.desired_position_ = joint.prev_desired_position_ + joint.desired_velocity_ * robot->time_delta_;
    if (params.clamp_joint_positions) {
      if (joint.desired_position_ < joint.min_position_) joint.desired_position_ = joint.min_position_;
      else if (joint.desired_position_ > joint.max_position_) joint.desired_position_ = joint.max_position_;
    }
    double min_diff = std::fabs(joint.min_position_ - joint.desired_position_);
    double max_diff = std::fabs(joint.max_position_ - joint.desired_position_);
    double half_joint_range = (joint.max_position_ - joint.min_position_) / 2.0;
    double limit_proximity = half_joint_range != 0 ? std::min(min_diff, max_diff) / half_joint_range : 1.0;
    min_limit_proximity = std::min(limit_proximity, min_limit_proximity);
  }
  return min_limit_proximity;
}

void Leg::touchdownDetection() {  // model.cpp:712
  if (tip_force_measured_.norm() > robot->params_.touchdown_threshold && step_plane_pose_ == Pose::Undefined()) {
    step_plane_pose_ = current_tip_pose_;
  } else if (tip_force_measured_.norm() < robot->params_.liftoff_threshold) {
    step_plane_pose_ = Pose::Undefined();
  }
}

double Leg::applyIK(bool simulation) {  // model.cpp:861
  Pose leg_frame_desired_tip_pose = jointPoseJointFrame(1, desired_tip_pose_);
  Pose leg_frame_current_tip_pose = jointPoseJointFrame(1, current_tip_pose_);
  Vec3 position_delta = leg_frame_desired_tip_pose.position_ - leg_frame_current_tip_pose.position_;

  double delta[6] = {position_delta[0], position_delta[1], position_delta[2], 0, 0, 0};
  double joint_position_delta[SHC_MAX_DOF];
  solveIK(delta, false, joint_position_delta);

  bool rotation_constrained = !isApproxQuat(desired_tip_pose_.rotation_, UndefinedRotation());
  if (rotation_constrained) {
    updateJointPositions(joint_position_delta, true);
    applyFK();
    Vec3 desired_tip_direction = leg_frame_desired_tip_pose.rotation_.transformVector(UnitX());
    Vec3 current_tip_direction = leg_frame_current_tip_pose.rotation_.transformVector(UnitX());
    Quat difference = fromTwoVectors(current_tip_direction, desired_tip_direction);
    double angle;
    Vec3 axis;
    angleAxisFromQuat(difference.normalized(), &angle, &axis);
    Vec3 rotation_delta = axis * angle;
    double d2[6] = {0, 0, 0, rotation_delta[0], rotation_delta[1], rotation_delta[2]};
    solveIK(d2, true, joint_position_delta);
  }

  double ik_success = updateJointPositions(joint_position_delta, simulation);
  applyFK();

  for (int i = 0; i < 3; ++i) {
    Vec3 position_error = current_tip_pose_.position_ - desired_tip_pose_.position_;
    if (std::fabs(position_error[i]) > IK_TOLERANCE) ik_success = 0.0;
  }

  if (rotation_constrained && !ik_success) {
    desired_tip_pose_.rotation_ = UndefinedRotation();
    ik_success = applyIK(simulation);
  }

  calculateTipForce();
  return ik_success;
}

Pose Leg::applyFK(bool set_current, bool use_actual) {  // model.cpp:945
  for (int i = 2; i <= joint_count_; ++i) {
    const Link& ref = links[i - 1];
    double joint_angle = use_actual ? joints[i - 1].current_position_ : joints[i - 1].desired_position_;
    joints[i].current_transform_ = createDHMatrix(ref.d, ref.theta + joint_angle, ref.r, ref.alpha);
  }
  const Link& ref = links[joint_count_];
  double joint_angle = use_actual ? joints[joint_count_].current_position_ : joints[joint_count_].desired_position_;
  tip_transform_ = createDHMatrix(ref.d, ref.theta + joint_angle, ref.r, ref.alpha);

  Pose tip_pose = tipPoseRobotFrame();
  if (set_current && !use_actual) {
    if (current_tip_pose_ != Pose::Undefined()) {
      current_tip_velocity_ = (tip_pose.position_ - current_tip_pose_.position_) / robot->time_delta_;
    }
    current_tip_pose_ = tip_pose;
  }
  return tip_pose;
}

}  // namespace shc_oracle
