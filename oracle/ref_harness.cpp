// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// extern "C" harness around the REFERENCE'S OWN control code.  oracle/Makefile.ref compiles the reference's unmodified
// translation units where they lie (/root/reference/src/{state_controller,model,walk_controller,pose_controller,
// admittance_controller,debug_visualiser}.cpp) against the stand-in headers of oracle/shim/ (ROS, tf2, Eigen and Boost are
// absent from the image) and links them with this file into oracle/_ref/libshc_ref.so.  Nothing of the reference is
// copied: this file only
//   * fills the in-process parameter table the reference's StateController::initParameters reads (from the same
//     shc_config the CUDA engine and the restated oracle are configured with),
//   * delivers inputs by calling the reference's own subscriber callbacks (what ros::spinOnce() would do),
//   * calls StateController::loop() and the reference's publishers, and
//   * reads the controllers' members back into shc_robot_state / shc_startup records for comparison.
// It is used by tests/ and by tests/golden/make_ref_golden.py to pin the oracle; the product never
// loads it.  What it cannot pin is third-party arithmetic: Eigen's and Boost.Odeint's kernels are the stand-ins'.
#include <cstring>
#include <map>
#include <memory>
#include <chrono>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "ros/ros.h"
#include "msg_common.h"
#include "tf2_ros/tf2_common.h"
#include "dynamic_reconfigure/server.h"
#include "boost/numeric/odeint.hpp"
#include "Eigen/Geometry"
#include "syropod_highlevel_controller/msgs_generated.h"

// the harness reads the controllers' private members (access specifiers do not change the layout)
#define private public
#define protected public
#include "syropod_highlevel_controller/state_controller.h"
#include "syropod_highlevel_controller/admittance_controller.h"
#undef private
#undef protected

#include "../include/shc_config.h"
#include "../include/shc_msgs.h"
#include "../include/shc_state.h"

namespace {

std::vector<std::string> legNames(int L) {
  if (L == 6) return {"AR", "BR", "CR", "CL", "BL", "AL"};
  if (L == 8) return {"AR", "BR", "CR", "DR", "DL", "CL", "BL", "AL"};
  std::vector<std::string> v;
  for (int i = 0; i < L; ++i) v.push_back(std::string("L") + char('A' + i));
  return v;
}
std::vector<std::string> jointNames(int D) {
  static const char* n[] = {"coxa", "femur", "tibia", "tarsus", "pes", "j6"};
  std::vector<std::string> v;
  for (int i = 0; i < D; ++i) v.push_back(n[i]);
  return v;
}

void fillGaitParams(const shc_config& c, const std::string& gait_name);

std::map<std::string, double> adjustable(double v) { return {{"default", v}, {"min", -1e9}, {"max", 1e9}, {"step", 0.0}}; }

// The reference's rosparam tree (config/default.yaml, gait.yaml, auto_pose.yaml layout) from an shc_config.
void fillParams(const shc_config& c) {
  shc_shim::ParamTable& p = shc_shim::params();
  p.clear();
  const std::string P = "syropod/parameters/";
  const int L = c.leg_count, D = c.joint_count;
  const std::vector<std::string> legs = legNames(L), joints = jointNames(D);
  p.d[P + "time_delta"] = c.time_delta;
  p.b[P + "imu_posing"] = c.imu_posing;
  p.b[P + "auto_posing"] = c.auto_posing;
  p.b[P + "rough_terrain_mode"] = c.rough_terrain_mode;
  p.b[P + "manual_posing"] = c.manual_posing;
  p.b[P + "inclination_posing"] = c.inclination_posing;
  p.b[P + "admittance_control"] = c.admittance_control;
  p.b[P + "individual_control_interface"] = true;
  p.b[P + "combined_control_interface"] = true;
  p.s[P + "syropod_type"] = "harness";
  p.vs[P + "leg_id"] = legs;
  p.vs[P + "joint_id"] = joints;
  std::vector<std::string> links = {"base"};
  links.insert(links.end(), joints.begin(), joints.end());
  p.vs[P + "link_id"] = links;
  std::map<std::string, int> dof;
  for (const auto& n : legs) dof[n] = D;
  p.mi[P + "leg_DOF"] = dof;
  p.b[P + "clamp_joint_positions"] = c.clamp_joint_positions;
  p.b[P + "clamp_joint_velocities"] = c.clamp_joint_velocities;
  p.b[P + "ignore_IK_warnings"] = true;
  for (int l = 0; l < L; ++l) {
    p.md[P + legs[l] + "_base_link_parameters"] = {{"d", c.link_d[l][0]}, {"theta", c.link_theta[l][0]}, {"r", c.link_r[l][0]}, {"alpha", c.link_alpha[l][0]}};
    for (int j = 0; j < D; ++j) {
      p.md[P + legs[l] + "_" + joints[j] + "_link_parameters"] = {{"d", c.link_d[l][j + 1]}, {"theta", c.link_theta[l][j + 1]}, {"r", c.link_r[l][j + 1]}, {"alpha", c.link_alpha[l][j + 1]}};
      p.md[P + legs[l] + "_" + joints[j] + "_joint_parameters"] = {{"min", c.joint_min[l][j]}, {"max", c.joint_max[l][j]}, {"offset", c.joint_offset[l][j]}, {"packed", c.joint_packed[l][j]}, {"unpacked", c.joint_unpacked[l][j]}, {"max_vel", c.joint_max_vel[l][j]}};
    }
    p.md[P + legs[l] + "_stance_position"] = {{"x", c.stance_x[l]}, {"y", c.stance_y[l]}};
  }
  p.s[P + "gait_type"] = "tripod_gait";
  p.d[P + "body_clearance"] = c.body_clearance;
  p.md[P + "step_frequency"] = adjustable(c.step_frequency);
  p.md[P + "swing_height"] = adjustable(c.swing_height);
  p.md[P + "swing_width"] = adjustable(c.swing_width);
  p.md[P + "step_depth"] = adjustable(c.step_depth);
  p.md[P + "stance_span_modifier"] = adjustable(c.stance_span_modifier);
  p.s[P + "velocity_input_mode"] = c.velocity_input_mode == SHC_VELOCITY_REAL ? "real" : "throttle";
  p.d[P + "body_velocity_scaler"] = c.body_velocity_scaler;
  p.b[P + "force_cruise_velocity"] = true;
  p.md[P + "linear_cruise_velocity"] = {{"x", 1.0}, {"y", 0.0}};
  p.d[P + "angular_cruise_velocity"] = 0.5;
  p.d[P + "cruise_control_time_limit"] = 0.0;
  p.b[P + "overlapping_walkspaces"] = c.overlapping_walkspaces;
  p.b[P + "force_normal_touchdown"] = c.force_normal_touchdown;
  p.b[P + "gravity_aligned_tips"] = c.gravity_aligned_tips;
  p.d[P + "touchdown_threshold"] = c.touchdown_threshold;
  p.d[P + "liftoff_threshold"] = c.liftoff_threshold;
  p.s[P + "auto_pose_type"] = "auto";
  p.b[P + "start_up_sequence"] = false;
  p.d[P + "time_to_start"] = c.time_to_start;
  p.md[P + "rotation_pid_gains"] = {{"p", c.rotation_pid_p}, {"i", c.rotation_pid_i}, {"d", c.rotation_pid_d}};
  p.md[P + "max_translation"] = {{"x", c.max_translation[0]}, {"y", c.max_translation[1]}, {"z", c.max_translation[2]}};
  p.md[P + "max_rotation"] = {{"roll", c.max_rotation[0]}, {"pitch", c.max_rotation[1]}, {"yaw", c.max_rotation[2]}};
  p.d[P + "max_translation_velocity"] = c.max_translation_velocity;
  p.d[P + "max_rotation_velocity"] = c.max_rotation_velocity;
  p.s[P + "leg_manipulation_mode"] = "tip_control";
  p.b[P + "dynamic_stiffness"] = c.dynamic_stiffness;
  p.b[P + "use_joint_effort"] = c.use_joint_effort;
  p.d[P + "integrator_step_time"] = c.integrator_step_time;
  p.md[P + "virtual_mass"] = adjustable(c.virtual_mass);
  p.md[P + "virtual_stiffness"] = adjustable(c.virtual_stiffness);
  p.md[P + "virtual_damping_ratio"] = adjustable(c.virtual_damping_ratio);
  p.md[P + "force_gain"] = adjustable(c.force_gain);
  p.d[P + "load_stiffness_scaler"] = c.load_stiffness_scaler;
  p.d[P + "swing_stiffness_scaler"] = c.swing_stiffness_scaler;
  p.b[P + "debug_rviz"] = false;
  p.s[P + "console_verbosity"] = "error";
  for (const char* k : {"debug_move_to_joint_position", "debug_step_to_position", "debug_swing_trajectory", "debug_stance_trajectory",
                        "debug_execute_sequence", "debug_workspace_calculations", "debug_ik"})
    p.b[P + k] = false;
  fillGaitParams(c, "tripod_gait");
}

// Gait and auto-pose parameter sets as gait.yaml / auto_pose.yaml file them: under the gait's name.  The harness uses the
// four names of the reference's GaitDesignation merely as slots ("tripod_gait" at creation, another one per gait change).
void fillGaitParams(const shc_config& c, const std::string& gait_name) {
  shc_shim::ParamTable& p = shc_shim::params();
  const int L = c.leg_count;
  const std::vector<std::string> legs = legNames(L);
  const std::string G = "syropod/gait_parameters/" + gait_name + "/";
  p.i[G + "stance_phase"] = c.stance_phase;
  p.i[G + "swing_phase"] = c.swing_phase;
  p.i[G + "phase_offset"] = c.phase_offset;
  std::map<std::string, int> mult;
  for (int l = 0; l < L; ++l) mult[legs[l]] = c.offset_multiplier[l];
  p.mi[G + "offset_multiplier"] = mult;
  const std::string A = "syropod/auto_pose_parameters/" + gait_name + "_pose/";
  p.d[A + "pose_frequency"] = c.pose_frequency;
  p.i[A + "pose_phase_length"] = c.pose_phase_length;
  const int K = c.auto_poser_count;
  p.vi[A + "pose_phase_starts"] = std::vector<int>(c.pose_phase_starts, c.pose_phase_starts + K);
  p.vi[A + "pose_phase_ends"] = std::vector<int>(c.pose_phase_ends, c.pose_phase_ends + K);
  std::map<std::string, int> ns, ne;
  std::map<std::string, double> nr;
  for (int l = 0; l < L; ++l) {
    ns[legs[l]] = c.pose_negation_phase_starts[l];
    ne[legs[l]] = c.pose_negation_phase_ends[l];
    nr[legs[l]] = c.negation_transition_ratio[l];
  }
  p.mi[A + "pose_negation_phase_starts"] = ns;
  p.mi[A + "pose_negation_phase_ends"] = ne;
  p.md[A + "negation_transition_ratio"] = nr;
  p.vd[A + "x_amplitudes"] = std::vector<double>(c.x_amplitudes, c.x_amplitudes + K);
  p.vd[A + "y_amplitudes"] = std::vector<double>(c.y_amplitudes, c.y_amplitudes + K);
  p.vd[A + "z_amplitudes"] = std::vector<double>(c.z_amplitudes, c.z_amplitudes + K);
  p.vd[A + "gravity_amplitudes"] = std::vector<double>(c.gravity_amplitudes, c.gravity_amplitudes + K);
  p.vd[A + "roll_amplitudes"] = std::vector<double>(c.roll_amplitudes, c.roll_amplitudes + K);
  p.vd[A + "pitch_amplitudes"] = std::vector<double>(c.pitch_amplitudes, c.pitch_amplitudes + K);
  p.vd[A + "yaw_amplitudes"] = std::vector<double>(c.yaw_amplitudes, c.yaw_amplitudes + K);
}

bool g_transition_through_loop = false;

struct RefRobot {
  shc_config cfg;
  std::unique_ptr<StateController> sc;
  std::vector<std::string> legs;
  int startup_loops = 0;
  int id = 0;  // distinguishes the time stamps (hence the tf answers) of requests to different robots of one process
};

void put3(double* o, const Eigen::Vector3d& v) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
void putQuat(double* o, const Eigen::Quaterniond& q) { o[0] = q.w(); o[1] = q.x(); o[2] = q.y(); o[3] = q.z(); }
void putPose(double* o, const Pose& p) { put3(o, p.position_); putQuat(o + 3, p.rotation_); }

std_msgs::Int8 int8(int v) { std_msgs::Int8 m; m.data = int8_t(v); return m; }

}  // namespace

extern "C" {

// StateController() + what main.cpp does before its loop (main.cpp:98-99: init(), initModel(use_default_joint_positions =
// true, no joint states have arrived)), then START is pressed (robot state RUNNING requested) every loop until the direct
// start-up has brought the robot to READY (state_controller.cpp:254-281).  As in the restated oracle's harness the robot is
// then put in RUNNING state directly, so that cycle 0 is the first full loop() in RUNNING state.
void* shc_ref_create(const shc_config* cfg) {
  static int next_id = 0;
  RefRobot* r = new RefRobot();
  r->id = next_id++;
  r->cfg = *cfg;
  r->legs = legNames(cfg->leg_count);
  fillParams(*cfg);
  shc_shim::runtime() = shc_shim::Runtime();
  shc_shim::tf_table().clear();
  r->sc.reset(new StateController());
  StateController& sc = *r->sc;
  sc.systemStateCallback(int8(OPERATIONAL));
  sc.init();
  sc.initModel(true);
  // Leg::virtual_stiffness_ has no initialiser in the reference (model.h:516; first written by updateStiffness,
  // admittance_controller.cpp:103): zero until then, as in the restated oracle and the engine's state record.
  for (auto& lp : *sc.model_->getLegContainer()) lp.second->virtual_stiffness_ = 0.0;
  // StateController::linear_velocity_input_ has no initialiser either (state_controller.h:366; angular_velocity_input_ has):
  // "no command received yet" is zero
  sc.linear_velocity_input_ = Eigen::Vector2d::Zero();
  // PoseController::pose_reset_mode_ has no initialiser and no constructor writes it (pose_controller.h:269): in the node the
  // remote's pose_reset_mode topic sets it before it matters; here "nothing requested" is NO_RESET.  (Left to the heap, a
  // rollout depended on what had run in the process before.)
  sc.poser_->setPoseResetMode(NO_RESET);
  int loops = 0;
  while (sc.robot_state_ != READY && loops < 100000) {
    sc.robotStateCallback(int8(RUNNING));
    sc.loop();
    ++loops;
  }
  r->startup_loops = loops;
  if (g_transition_through_loop) {
    // READY -> RUNNING the reference's own way (state_controller.cpp:283-288): START pressed again, and the loop() that makes
    // the transition also runs its first runningState() — with a zero velocity input, since bodyVelocityInputCallback
    // ignores commands before RUNNING (state_controller.cpp:1129)
    sc.robotStateCallback(int8(RUNNING));
    sc.loop();
    return r;
  }
  sc.robot_state_ = RUNNING;
  sc.new_robot_state_ = RUNNING;
  sc.transition_state_flag_ = false;
  return r;
}

// 1: the robots created from now on enter RUNNING through robotStateCallback + loop() (one control cycle with a zero command
// has then already run); 0 (default): RUNNING is set directly, so that cycle 0 is the caller's.  tests/test_reference_pin.py
// shows the two to be the same state one zero-command cycle apart.
void shc_ref_transition_through_loop(int on) { g_transition_through_loop = on != 0; }

// The joint commands of every loop() of the reference's direct start-up for a robot whose joints are at q_init [L][D] when it
// begins: a joint_states message through jointStatesCallback, then init() / initModel(false) as main.cpp:66-99 does once all
// joint positions have arrived.  out [max_loops][L][D], one row per loop() from the first that moves the joints; returns
// the number of rows.
int shc_ref_startup_trajectory(const shc_config* cfg, const double* q_init, int max_loops, double* out) {
  fillParams(*cfg);
  shc_shim::runtime() = shc_shim::Runtime();
  StateController sc;
  const int L = cfg->leg_count, D = cfg->joint_count;
  sc.systemStateCallback(int8(OPERATIONAL));
  bool use_defaults = true;
  if (q_init) {
    sensor_msgs::JointState js;
    int k = 0;
    for (auto& lp : *sc.model_->getLegContainer())
      for (auto& jp : *lp.second->getJointContainer()) {
        js.name.push_back(jp.second->id_name_);
        js.position.push_back(q_init[k++] + jp.second->offset_);
      }
    sc.jointStatesCallback(js);
    use_defaults = !sc.jointPositionsInitialised();
  }
  sc.init();
  sc.initModel(use_defaults);
  sc.linear_velocity_input_ = Eigen::Vector2d::Zero();
  sc.poser_->setPoseResetMode(NO_RESET);
  int rows = 0, loops = 0;
  while (sc.robot_state_ != READY && loops < 100000) {
    sc.robotStateCallback(int8(RUNNING));
    const bool moving = sc.robot_state_ == PACKED;  // the first loop() only leaves UNKNOWN
    sc.loop();
    ++loops;
    if (moving && rows < max_loops) {
      int k = 0;
      for (auto& lp : *sc.model_->getLegContainer())
        for (auto& jp : *lp.second->getJointContainer()) out[size_t(rows) * L * D + k++] = jp.second->desired_position_;
      ++rows;
    }
  }
  return rows;
}

void shc_ref_destroy(void* h) { delete static_cast<RefRobot*>(h); }
int shc_ref_startup_loops(void* h) { return static_cast<RefRobot*>(h)->startup_loops; }
long shc_ref_assert_failures(char* first, int cap) {
  const shc_shim::Runtime& rt = shc_shim::runtime();
  if (first && cap > 0) {
    std::strncpy(first, rt.first_assert.c_str(), size_t(cap) - 1);
    first[cap - 1] = 0;
  }
  return rt.assert_failures;
}
int shc_ref_shutdown_requested(void) { return shc_shim::runtime().shutdown_requested ? 1 : 0; }

// PoseController::setPoseResetMode — the setter the engine's shc_set_pose_reset_mode and the oracle's input stand for —
// and, separately, the subscriber callback around it, which ignores requests while an IMMEDIATE_ALL_RESET is pending
// (state_controller.cpp:1193-1203).
void shc_ref_set_pose_reset_mode(void* h, int mode) { static_cast<RefRobot*>(h)->sc->poser_->setPoseResetMode(PoseResetMode(mode)); }
void shc_ref_pose_reset_callback(void* h, int mode) { static_cast<RefRobot*>(h)->sc->poseResetCallback(int8(mode)); }

// joint_states message with efforts only for the robot's joints in their current positions (jointStatesCallback)
void shc_ref_set_joint_efforts(void* h, const double* efforts) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  sensor_msgs::JointState js;
  int l = 0;
  for (auto& lp : *sc.model_->getLegContainer()) {
    int j = 0;
    for (auto& jp : *lp.second->getJointContainer()) {
      js.name.push_back(jp.second->id_name_);
      js.position.push_back(jp.second->current_position_ + jp.second->offset_);
      js.effort.push_back(efforts[(l * r->cfg.joint_count) + j]);
      ++j;
    }
    ++l;
  }
  sc.jointStatesCallback(js);
}

// Gait switch (gaitSelectionCallback, then StateController::changeGait inside the following loop() calls: it zeroes the
// velocity input until the walker has STOPPED, then re-reads the gait and auto-pose parameter sets, regenerates the step
// cycle, the limit maps and phase offsets and the auto posers — state_controller.cpp:513-540).  new_cfg supplies the new
// gait / auto-pose block (everything else must be unchanged).  Returns the slot name's index.
int shc_ref_select_gait(void* h, const shc_config* new_cfg) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  static const char* names[] = {"wave_gait", "amble_gait", "ripple_gait", "tripod_gait"};  // GaitDesignation order
  int slot = (int(sc.gait_selection_) + 1) % 4;
  if (slot < 0) slot = 0;
  fillGaitParams(*new_cfg, names[slot]);
  r->cfg = *new_cfg;
  sc.gaitSelectionCallback(int8(slot));
  return slot;
}
int shc_ref_gait_change_pending(void* h) { return static_cast<RefRobot*>(h)->sc->gait_change_flag_ ? 1 : 0; }

// Adjustable parameters (step_frequency, swing_height, swing_width, step_depth, stance_span_modifier, virtual_mass,
// virtual_stiffness, virtual_damping_ratio, force_gain): one dynamic_reconfigure request carrying new_cfg's values through
// dynamicParameterCallback (state_controller.cpp:1465-1548), which picks the FIRST parameter that differs;
// StateController::adjustParameter applies it inside the following loop() (state_controller.cpp:451-508).  Returns 1 if a
// parameter differed.
int shc_ref_adjust_parameter(void* h, const shc_config* new_cfg) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  syropod_highlevel_controller::DynamicConfig c;
  c.step_frequency = new_cfg->step_frequency;
  c.swing_height = new_cfg->swing_height;
  c.swing_width = new_cfg->swing_width;
  c.step_depth = new_cfg->step_depth;
  c.stance_span_modifier = new_cfg->stance_span_modifier;
  c.virtual_mass = new_cfg->virtual_mass;
  c.virtual_stiffness = new_cfg->virtual_stiffness;
  c.virtual_damping_ratio = new_cfg->virtual_damping_ratio;
  c.force_gain = new_cfg->force_gain;
  const Parameters& p = sc.params_;
  const bool differs = c.step_frequency != p.step_frequency.current_value || c.swing_height != p.swing_height.current_value ||
                       c.swing_width != p.swing_width.current_value || c.step_depth != p.step_depth.current_value ||
                       c.stance_span_modifier != p.stance_span_modifier.current_value || c.virtual_mass != p.virtual_mass.current_value ||
                       c.virtual_stiffness != p.virtual_stiffness.current_value ||
                       c.virtual_damping_ratio != p.virtual_damping_ratio.current_value || c.force_gain != p.force_gain.current_value;
  if (!differs) return 0;
  sc.dynamicParameterCallback(c, 0);
  r->cfg = *new_cfg;
  return 1;
}
int shc_ref_parameter_adjust_pending(void* h) { return static_cast<RefRobot*>(h)->sc->parameter_adjust_flag_ ? 1 : 0; }

// One control cycle: inputs through the reference's callbacks, then loop().
// cmd [3]; imu [10] (quat wxyz, gyro xyz, accel xyz) or NULL; tip_force [L][3] or NULL; manual [6] or NULL;
// step_plane [L][3] or NULL.
void shc_ref_step(void* h, const double* cmd, const double* imu, const double* tip_force, const double* manual,
                  const double* step_plane) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  const int L = r->cfg.leg_count;
  shc_shim::runtime().now += r->cfg.time_delta;
  if (imu) {
    sensor_msgs::Imu m;
    m.orientation.w = imu[0]; m.orientation.x = imu[1]; m.orientation.y = imu[2]; m.orientation.z = imu[3];
    m.angular_velocity.x = imu[4]; m.angular_velocity.y = imu[5]; m.angular_velocity.z = imu[6];
    m.linear_acceleration.x = imu[7]; m.linear_acceleration.y = imu[8]; m.linear_acceleration.z = imu[9];
    sc.imuCallback(m);
  }
  if (tip_force) {
    syropod_highlevel_controller::TipState ts;
    for (int l = 0; l < L; ++l) {
      ts.name.push_back(r->legs[l] + "_tip");
      geometry_msgs::Wrench w;
      w.force.x = tip_force[3 * l]; w.force.y = tip_force[3 * l + 1]; w.force.z = tip_force[3 * l + 2];
      ts.wrench.push_back(w);
    }
    sc.tipStatesCallback(ts);
  }
  if (step_plane) {
    syropod_highlevel_controller::TipState ts;
    for (int l = 0; l < L; ++l) {
      ts.name.push_back(r->legs[l] + "_tip");
      geometry_msgs::Vector3 v;
      v.x = step_plane[3 * l]; v.y = step_plane[3 * l + 1]; v.z = step_plane[3 * l + 2];
      ts.step_plane.push_back(v);
    }
    sc.tipStatesCallback(ts);
  }
  if (manual) {
    geometry_msgs::Twist t;
    t.linear.x = manual[0]; t.linear.y = manual[1]; t.linear.z = manual[2];
    t.angular.x = manual[3]; t.angular.y = manual[4]; t.angular.z = manual[5];
    sc.bodyPoseInputCallback(t);
  }
  geometry_msgs::Twist v;
  v.linear.x = cmd[0]; v.linear.y = cmd[1]; v.angular.z = cmd[2];
  sc.bodyVelocityInputCallback(v);
  sc.loop();
}

// One loop() of a stepping / joint-space sequence, called on the reference's PoseController as StateController::
// transitionRobotState does (state_controller.cpp:286-350): kind 0 = stepToNewStance, 1 = packLegs(time), 2 = unpackLegs(time),
// 3 / 4 = executeSequence(START_UP / SHUT_DOWN).  Returns the progress value; -2 if the reference asked for a shutdown
// (pose_controller.cpp:441: the generated sequence could not be executed).
int shc_ref_sequence_step(void* h, int kind, double time) {
  RefRobot* r = static_cast<RefRobot*>(h);
  PoseController& p = *r->sc->poser_;
  int progress;
  switch (kind) {
    case 0: progress = p.stepToNewStance(); break;
    case 1: progress = p.packLegs(time); break;
    case 2: progress = p.unpackLegs(time); break;
    case 3: progress = p.executeSequence(START_UP); break;
    default: progress = p.executeSequence(SHUT_DOWN); break;
  }
  if (shc_shim::runtime().shutdown_requested) progress = -2;
  return progress;
}

// Overwrites the desired joint positions / velocities [L][D] and re-runs the forward kinematics (Leg::applyFK), the way the
// oracle's state import does: lets a test take single loops "from identical joint state" on legs whose redundant joints
// drift apart between any two double-precision builds over thousands of closed-loop IK iterations.
void shc_ref_set_joint_state(void* h, const double* position, const double* velocity) {
  RefRobot* r = static_cast<RefRobot*>(h);
  int k = 0;
  for (auto& lp : *r->sc->model_->getLegContainer()) {
    for (auto& jp : lp.second->joint_container_) {
      jp.second->desired_position_ = position[k];
      jp.second->desired_velocity_ = velocity[k];
      ++k;
    }
    lp.second->applyFK();
  }
}

// CPU baseline timing (bench.py --impl reference / cpu_baseline): `cycles` control cycles of n robots with per-cycle
// commands cmd_seq [cycles][n][3], the robots partitioned once over n_threads std::threads (robots are independent);
// returns wall seconds.  Each cycle is what ros::spinOnce() + StateController::loop() do for a velocity command.
double shc_ref_batch_run_seq(void** handles, int n, const double* cmd_seq, int cycles, int n_threads) {
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      StateController& sc = *static_cast<RefRobot*>(handles[i])->sc;
      for (int c = 0; c < cycles; ++c) {
        const double* cmd = cmd_seq + (size_t(c) * n + i) * 3;
        geometry_msgs::Twist v;
        v.linear.x = cmd[0]; v.linear.y = cmd[1]; v.angular.z = cmd[2];
        sc.bodyVelocityInputCallback(v);
        sc.loop();
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, int((long long)n * t / n_threads), int((long long)n * (t + 1) / n_threads));
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void shc_ref_get_joints(void* h, double* out) {
  RefRobot* r = static_cast<RefRobot*>(h);
  int k = 0;
  for (auto& lp : *r->sc->model_->getLegContainer())
    for (auto& jp : *lp.second->getJointContainer()) out[k++] = jp.second->desired_position_;
}

void shc_ref_get_state(void* h, shc_robot_state* s) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  WalkController& w = *sc.walker_;
  PoseController& p = *sc.poser_;
  Model& m = *sc.model_;
  std::memset(s, 0, sizeof(*s));
  s->desired_linear_velocity[0] = w.desired_linear_velocity_[0];
  s->desired_linear_velocity[1] = w.desired_linear_velocity_[1];
  s->desired_angular_velocity = w.desired_angular_velocity_;
  s->walk_state = w.walk_state_;
  s->legs_at_correct_phase = w.legs_at_correct_phase_;
  s->legs_completed_first_step = w.legs_completed_first_step_;
  s->return_to_default_attempted = w.return_to_default_attempted_;
  s->pose_state = w.pose_state_;
  put3(s->walk_plane, w.walk_plane_);
  put3(s->walk_plane_normal, w.walk_plane_normal_);
  putPose(s->odometry_ideal, w.odometry_ideal_);
  putPose(s->walk_plane_pose, p.walk_plane_pose_);
  putPose(s->origin_walk_plane_pose, p.origin_walk_plane_pose_);
  putPose(s->manual_pose, p.manual_pose_);
  putPose(s->imu_pose, p.imu_pose_);
  putPose(s->inclination_pose, p.inclination_pose_);
  putPose(s->auto_pose, p.auto_pose_);
  put3(s->rotation_absement_error, p.rotation_absement_error_);
  put3(s->rotation_position_error, p.rotation_position_error_);
  put3(s->rotation_velocity_error, p.rotation_velocity_error_);
  putPose(s->tip_align_pose, p.tip_align_pose_);
  putPose(s->origin_tip_align_pose, p.origin_tip_align_pose_);
  s->auto_posing_state = p.auto_posing_state_;
  s->pose_phase = p.pose_phase_;
  size_t k = 0;
  for (auto& ap : p.auto_poser_container_) {
    if (k >= SHC_MAX_AUTO_POSERS) break;
    s->auto_poser_flags[k++] = (ap->start_check_ ? 1 : 0) | (ap->end_check_.first ? 2 : 0) | (ap->end_check_.second ? 4 : 0) |
                               (ap->allow_posing_ ? 8 : 0);
  }
  putPose(s->current_pose, m.current_pose_);
  int i = 0;
  for (auto& lp : *m.getLegContainer()) {
    Leg& leg = *lp.second;
    LegStepper& st = *leg.leg_stepper_;
    shc_leg_state& o = s->legs[i++];
    int j = 0;
    for (auto& jp : leg.joint_container_) {
      o.joint_position[j] = jp.second->desired_position_;
      o.joint_velocity[j] = jp.second->desired_velocity_;
      ++j;
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
    o.negate_auto_pose = leg.leg_poser_->negate_auto_pose_;
    putQuat(o.tip_rotation, st.current_tip_pose_.rotation_);
    putQuat(o.origin_tip_rotation, st.origin_tip_pose_.rotation_);
    putQuat(o.target_tip_rotation, st.target_tip_pose_.rotation_);
    o.step_plane_defined = leg.step_plane_pose_ != Pose::Undefined();
    if (o.step_plane_defined) put3(o.step_plane_position, leg.step_plane_pose_.position_);
    o.touchdown_detection = st.touchdown_detection_;
    // An ExternalTarget the reference has never been given (empty frame id) is indeterminate memory there (Pose()
    // initialises nothing, pose.h:21; swing_clearance_ has no initialiser, walk_controller.h:41): reported as the identity
    // record.
    const Pose identity = Pose::Identity();
    const bool tgt = !st.external_target_.frame_id_.empty(), dft = !st.external_default_.frame_id_.empty();
    putPose(o.external_target_pose, tgt ? st.external_target_.pose_ : identity);
    putPose(o.external_target_transform, tgt ? st.external_target_.transform_ : identity);
    o.external_target_clearance = tgt ? st.external_target_.swing_clearance_ : 0.0;
    o.external_target_defined = st.external_target_.defined_;
    o.external_target_odom_frame = tgt && st.external_target_.frame_id_ == "odom_ideal";
    putPose(o.external_default_pose, dft ? st.external_default_.pose_ : identity);
    putPose(o.external_default_transform, dft ? st.external_default_.transform_ : identity);
    o.external_default_defined = st.external_default_.defined_;
    put3(o.model_tip_position, leg.current_tip_pose_.position_);
    put3(o.desired_tip_position, leg.desired_tip_pose_.position_);
  }
}

void shc_ref_get_startup(void* h, shc_startup* out) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  WalkController& w = *sc.walker_;
  std::memset(out, 0, sizeof(*out));
  int i = 0;
  for (auto& lp : *sc.model_->getLegContainer()) {
    Leg& leg = *lp.second;
    int j = 0;
    for (auto& jp : leg.joint_container_) out->default_joint[i][j++] = jp.second->default_position_;
    if (!leg.workspace_.empty()) {
      const Workplane& wp = leg.workspace_.begin()->second;
      for (int b = 0; b < SHC_N_BEARINGS; ++b) out->workspace[i][b] = wp.count(b * 45) ? wp.at(b * 45) : 0.0;
    }
    out->phase_offsets[i] = leg.leg_stepper_->phase_offset_;
    ++i;
  }
  for (int b = 0; b < SHC_N_BEARINGS; ++b) {
    const int k = b * 45;
    out->walkspace[b] = w.walkspace_.count(k) ? w.walkspace_.at(k) : 0.0;
    out->max_linear_speed[b] = w.max_linear_speed_.count(k) ? w.max_linear_speed_.at(k) : 0.0;
    out->max_angular_speed[b] = w.max_angular_speed_.count(k) ? w.max_angular_speed_.at(k) : 0.0;
    out->max_linear_acceleration[b] = w.max_linear_acceleration_.count(k) ? w.max_linear_acceleration_.at(k) : 0.0;
    out->max_angular_acceleration[b] = w.max_angular_acceleration_.count(k) ? w.max_angular_acceleration_.at(k) : 0.0;
  }
  out->step_frequency = w.step_.frequency_;
  out->period = w.step_.period_;
  out->swing_period = w.step_.swing_period_;
  out->stance_period = w.step_.stance_period_;
  out->stance_end = w.step_.stance_end_;
  out->swing_start = w.step_.swing_start_;
  out->swing_end = w.step_.swing_end_;
  out->stance_start = w.step_.stance_start_;
  out->pose_phase_length = sc.poser_->pose_phase_length_;
  out->pose_normaliser = sc.poser_->normaliser_;
  out->auto_pose_reference_leg = sc.poser_->auto_pose_reference_leg_ ? sc.poser_->auto_pose_reference_leg_->getIDNumber() : 0;
  out->startup_loops = r->startup_loops;
}

// Externally requested swing targets and default tip poses: one TargetTipPose message through targetTipPoseCallback
// (state_controller.cpp:1706-1768), and — standing for the tf tree generateExternalTargetTransforms looks up every loop
// (:703-750) — the robot's movement since each request, filed under the request's time stamp.  Per leg: *_defined [L],
// *_pose / *_transform [L][7] (position, quaternion wxyz), clearance [L], odom_frame [L] (frame "odom_ideal" or "walk_plane").
void shc_ref_request_tip_targets(void* h, const int* target_defined, const double* target_pose, const double* target_transform,
                                 const double* clearance, const int* odom_frame, const int* default_defined,
                                 const double* default_pose, const double* default_transform) {
  RefRobot* r = static_cast<RefRobot*>(h);
  const int L = r->cfg.leg_count;
  syropod_highlevel_controller::TargetTipPose msg;
  auto toMsg = [](const double* p) {
    geometry_msgs::Pose m;
    m.position.x = p[0]; m.position.y = p[1]; m.position.z = p[2];
    m.orientation.w = p[3]; m.orientation.x = p[4]; m.orientation.y = p[5]; m.orientation.z = p[6];
    return m;
  };
  auto file = [](const std::string& frame, double stamp, const double* t) {
    geometry_msgs::TransformStamped ts;
    ts.transform.translation.x = t[0]; ts.transform.translation.y = t[1]; ts.transform.translation.z = t[2];
    ts.transform.rotation.w = t[3]; ts.transform.rotation.x = t[4]; ts.transform.rotation.y = t[5]; ts.transform.rotation.z = t[6];
    shc_shim::tf_table()[frame + "@" + std::to_string(stamp) + "<-walk_plane"] = ts;
  };
  const geometry_msgs::Pose undefined = Pose::Undefined().toPoseMessage();
  for (int l = 0; l < L; ++l) {
    msg.name.push_back(r->legs[l]);
    geometry_msgs::PoseStamped t, d;
    t.pose = undefined;
    d.pose = undefined;
    if (target_defined[l]) {
      t.pose = toMsg(target_pose + 7 * l);
      t.header.frame_id = odom_frame[l] ? "odom_ideal" : "walk_plane";
      t.header.stamp = ros::Time(1000.0 + 100.0 * r->id + l);
      file(t.header.frame_id, t.header.stamp.toSec(), target_transform + 7 * l);
    }
    if (default_defined[l]) {
      d.pose = toMsg(default_pose + 7 * l);
      d.header.frame_id = "walk_plane";
      d.header.stamp = ros::Time(1050.0 + 100.0 * r->id + l);
      file(d.header.frame_id, d.header.stamp.toSec(), default_transform + 7 * l);
    }
    msg.target.push_back(t);
    msg.stance.push_back(d);
    msg.swing_clearance.push_back(clearance[l]);
  }
  r->sc->targetTipPoseCallback(msg);
}

// The reference's publishers (state_controller.cpp:777-1047) run as main.cpp runs them after loop(); what they published is
// read off the stand-in message bus / tf broadcaster into the same records the engine's shc_pack_messages fills.
// measured [L][D]: joint positions a joint_states message reports (NULL: the desired positions), through
// jointStatesCallback.
void shc_ref_get_messages(void* h, const double* measured, shc_joint_state_msg* js, shc_leg_state_msg* legs, shc_body_msg* body) {
  RefRobot* r = static_cast<RefRobot*>(h);
  StateController& sc = *r->sc;
  const int D = r->cfg.joint_count;
  {
    sensor_msgs::JointState in;
    int k = 0;
    for (auto& lp : *sc.model_->getLegContainer())
      for (auto& jp : *lp.second->getJointContainer()) {
        in.name.push_back(jp.second->id_name_);
        in.position.push_back((measured ? measured[k] : jp.second->desired_position_) + jp.second->offset_);
        ++k;
      }
    sc.jointStatesCallback(in);
  }
  sc.publishLegState();
  sc.publishVelocity();
  sc.publishPose();
  sc.publishRotationPoseError();
  sc.publishFrameTransforms();
  sc.publishDesiredJointState();
  shc_shim::Bus& bus = shc_shim::bus();
  auto pose7 = [](double* o, const geometry_msgs::Pose& p) {
    o[0] = p.position.x; o[1] = p.position.y; o[2] = p.position.z;
    o[3] = p.orientation.w; o[4] = p.orientation.x; o[5] = p.orientation.y; o[6] = p.orientation.z;
  };
  auto tf7 = [](double* o, const std::string& child) {
    const geometry_msgs::Transform& t = shc_shim::tf_sent().at(child).transform;
    o[0] = t.translation.x; o[1] = t.translation.y; o[2] = t.translation.z;
    o[3] = t.rotation.w; o[4] = t.rotation.x; o[5] = t.rotation.y; o[6] = t.rotation.z;
  };
  std::memset(js, 0, sizeof(*js));
  const sensor_msgs::JointState& out = *std::static_pointer_cast<sensor_msgs::JointState>(bus.last.at("desired_joint_states"));
  for (size_t k = 0; k < out.position.size(); ++k) {
    js->position[k] = out.position[k];
    js->velocity[k] = out.velocity[k];
    js->effort[k] = out.effort[k];
  }
  int l = 0;
  for (auto& lp : *sc.model_->getLegContainer()) {
    Leg& leg = *lp.second;
    shc_leg_state_msg& m = legs[l];
    std::memset(&m, 0, sizeof(m));
    const syropod_highlevel_controller::LegState& ls =
        *std::static_pointer_cast<syropod_highlevel_controller::LegState>(bus.last.at("shc/" + leg.getIDName() + "/state"));
    pose7(m.walker_tip_pose, ls.walker_tip_pose.pose);
    pose7(m.target_tip_pose, ls.target_tip_pose.pose);
    pose7(m.poser_tip_pose, ls.poser_tip_pose.pose);
    pose7(m.model_tip_pose, ls.model_tip_pose.pose);
    pose7(m.actual_tip_pose, ls.actual_tip_pose.pose);
    m.model_tip_velocity[0] = ls.model_tip_velocity.twist.linear.x;
    m.model_tip_velocity[1] = ls.model_tip_velocity.twist.linear.y;
    m.model_tip_velocity[2] = ls.model_tip_velocity.twist.linear.z;
    int j = 0;
    for (auto& jp : leg.joint_container_) {
      m.joint_positions[j] = ls.joint_positions[j];
      m.joint_velocities[j] = ls.joint_velocities[j];
      m.joint_efforts[j] = ls.joint_efforts[j];
      js->position_command[l * D + j] = std::static_pointer_cast<std_msgs::Float64>(bus.last.at(jp.second->id_name_ + "/command"))->data;
      tf7(m.joint_transform[j], jp.second->id_name_);
      ++j;
    }
    tf7(m.tip_transform, leg.getTip()->id_name_);
    m.stance_progress = ls.stance_progress;
    m.swing_progress = ls.swing_progress;
    m.time_to_swing_end = ls.time_to_swing_end;
    pose7(m.pose_delta, ls.pose_delta);
    pose7(m.auto_pose, ls.auto_pose);
    m.tip_force[0] = ls.tip_force.x; m.tip_force[1] = ls.tip_force.y; m.tip_force[2] = ls.tip_force.z;
    m.admittance_delta[0] = ls.admittance_delta.x; m.admittance_delta[1] = ls.admittance_delta.y; m.admittance_delta[2] = ls.admittance_delta.z;
    m.virtual_stiffness = ls.virtual_stiffness;
    ++l;
  }
  std::memset(body, 0, sizeof(*body));
  const geometry_msgs::Twist& v = *std::static_pointer_cast<geometry_msgs::Twist>(bus.last.at("shc/velocity"));
  body->velocity[0] = v.linear.x; body->velocity[1] = v.linear.y; body->velocity[2] = v.linear.z;
  body->velocity[3] = v.angular.x; body->velocity[4] = v.angular.y; body->velocity[5] = v.angular.z;
  const geometry_msgs::Twist& bp = *std::static_pointer_cast<geometry_msgs::Twist>(bus.last.at("shc/pose"));
  body->pose[0] = bp.linear.x; body->pose[1] = bp.linear.y; body->pose[2] = bp.linear.z;
  body->pose[3] = bp.angular.x; body->pose[4] = bp.angular.y; body->pose[5] = bp.angular.z;
  const std_msgs::Float32MultiArray& e = *std::static_pointer_cast<std_msgs::Float32MultiArray>(bus.last.at("shc/rotation_pose_error"));
  for (int k = 0; k < 9; ++k) body->rotation_pose_error[k] = e.data[k];  // float32 on the wire
  tf7(body->odom_ideal_to_base_link, "base_link");
  tf7(body->base_link_to_walk_plane, "walk_plane");
}

// Layered / simple workspace of one leg as the reference generated it in its start-up: heights [P], radii [P][9].
int shc_ref_workspace(void* h, int leg_index, int max_planes, double* heights, double* radii) {
  RefRobot* r = static_cast<RefRobot*>(h);
  Leg& leg = *r->sc->model_->getLegByIDNumber(leg_index);
  int n = 0;
  for (auto& pl : leg.workspace_) {
    if (n < max_planes) {
      heights[n] = pl.first;
      for (int b = 0; b < SHC_N_BEARINGS; ++b) radii[n * SHC_N_BEARINGS + b] = pl.second.count(b * 45) ? pl.second.at(b * 45) : 0.0;
    }
    ++n;
  }
  return n;
}

}  // extern "C"
