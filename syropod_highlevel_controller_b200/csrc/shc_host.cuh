// shc_host.cuh — host side of the engine: configuration checks, the constants block, and the engine's own restatement
// of the reference's one-off start-up computations whose results the per-cycle path treats as constants:
//   WalkController::generateStepCycle   walk_controller.cpp:365      generateLimits      walk_controller.cpp:231
//   PoseController::directStartup       pose_controller.cpp:463      (LegPoser::stepToPosition :1571,
//                                                                     LegPoser::transitionConfiguration :1476)
//   Model::generateWorkspaces           model.cpp:120 / Leg::generateWorkspace :309 (simple, one-plane workspace)
//   WalkController::generateWalkspace   walk_controller.cpp:57
// This runs once per engine on the CPU in double (it is start-up, not the hot path; SURVEY.md §8b "Non-cycle methods
// still called ... Host implementations; results uploaded as constants") and reuses the same chain / DLS routines the
// kernels use.  It shares no code with oracle/.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "../../include/shc_state.h"
#include "shc_cycle.cuh"
#include "shc_startup.cuh"

namespace shc {

inline bool validate_config(const shc_config& c, std::string& err, bool& unsupported) {
  unsupported = false;
  auto bad = [&](const char* m) { err = m; return false; };
  if (c.leg_count < 1 || c.leg_count > SHC_MAX_LEGS) return bad("leg_count must be 1..8");
  if (c.joint_count < 1 || c.joint_count > SHC_MAX_DOF) return bad("joint_count must be 1..5");
  if (!(c.time_delta > 0.0)) return bad("time_delta must be > 0");
  if (c.stance_phase <= 0 || c.swing_phase <= 0 || c.phase_offset <= 0) return bad("gait phases must be > 0");
  if (!(c.step_frequency > 0.0)) return bad("step_frequency must be > 0");
  if (c.auto_poser_count < 0 || c.auto_poser_count > SHC_MAX_AUTO_POSERS) return bad("auto_poser_count out of range");
  if (c.rough_terrain_mode && c.stance_span_modifier != 0.0) {
    unsupported = true;  // calculateStanceSpanChange would interpolate the layered workspace every default-tip update
    return bad("rough_terrain_mode with a stance_span_modifier is not built (stance span change over the layered workspace)");
  }
  if (c.rough_terrain_mode && !(c.liftoff_threshold <= c.touchdown_threshold)) return bad("liftoff_threshold must not exceed touchdown_threshold");
  if (c.joint_count != 3 && c.joint_count != 4 && c.joint_count != 5) { unsupported = true; return bad("kernels are instantiated for 3, 4 and 5 joints per leg"); }
  return true;
}

template <class R> void fill_static_consts(const shc_config& c, RealConsts<R>& k) {
  std::memset(&k, 0, sizeof(k));
  const double w = 0.1;  // JOINT_LIMIT_COST_WEIGHT (model.h:20)
  for (int l = 0; l < c.leg_count; ++l) {
    double th = c.link_theta[l][0], al = c.link_alpha[l][0], r = c.link_r[l][0], d = c.link_d[l][0];
    double m[9] = {cos(th), -sin(th) * cos(al), sin(th) * sin(al), sin(th), cos(th) * cos(al), -cos(th) * sin(al), 0.0, sin(al), cos(al)};
    for (int i = 0; i < 9; ++i) k.leg[l].t1r[i] = R(m[i]);
    k.leg[l].t1p[0] = R(r * cos(th)); k.leg[l].t1p[1] = R(r * sin(th)); k.leg[l].t1p[2] = R(d);
    for (int j = 0; j < c.joint_count; ++j) {
      k.leg[l].dh_d[j] = R(c.link_d[l][j + 1]);
      k.leg[l].dh_theta[j] = R(c.link_theta[l][j + 1]);
      k.leg[l].dh_r[j] = R(c.link_r[l][j + 1]);
      k.leg[l].dh_ca[j] = R(cos(c.link_alpha[l][j + 1]));
      k.leg[l].dh_sa[j] = R(sin(c.link_alpha[l][j + 1]));
      double lo = c.joint_min[l][j], hi = c.joint_max[l][j], vm = c.joint_max_vel[l][j];
      double range = hi - lo;
      k.leg[l].jmin[j] = R(lo); k.leg[l].jmax[j] = R(hi); k.leg[l].vmax[j] = R(vm);
      k.leg[l].joffset[j] = R(c.joint_offset[l][j]);
      k.leg[l].jcentre[j] = R(lo + range / 2.0);
      k.leg[l].jcost_pos[j] = R(range != 0.0 ? w / range : 0.0);
      k.leg[l].jgrad_pos[j] = R(range != 0.0 ? -(w * w) / (range * range) : 0.0);
      double vr = 2 * vm;
      k.leg[l].jcost_vel[j] = R(w / vr);
      k.leg[l].jgrad_vel[j] = R(-(w * w) / (vr * vr));
    }
    k.leg[l].identity_x = R(c.stance_x[l]);
    k.leg[l].identity_y = R(c.stance_y[l]);
    k.leg[l].ysign = R(c.stance_y[l] > 0.0 ? 1.0 : -1.0);
    k.leg[l].neg_ratio = R(c.negation_transition_ratio[l]);
  }
  k.dt = R(c.time_delta);
  k.inv_dt = R(1.0 / c.time_delta);
  k.swing_height = R(c.swing_height);
  k.swing_width = R(c.swing_width);
  k.body_clearance = R(c.body_clearance);
  k.lambda2 = R(0.02 * 0.02);
  k.swing_progress_scaler = R(std::max(1.0, double(c.swing_phase) / c.phase_offset));
  {
    const double edges[8] = {44.5, 89.5, 134.5, 179.5, 0.5, 45.5, 90.5, 135.5};
    for (int i = 0; i < 8; ++i) {
      k.sec_cos[i] = R(cos(edges[i] / 360.0 * 2.0 * kPi));
      k.sec_sin[i] = R(sin(edges[i] / 360.0 * 2.0 * kPi));
    }
  }
  for (int i = 0; i < 3; ++i) {
    k.max_translation[i] = R(c.max_translation[i]);
    k.max_rotation[i] = R(c.max_rotation[i]);
  }
  k.max_translation_velocity = R(c.max_translation_velocity);
  k.max_rotation_velocity = R(c.max_rotation_velocity);
  k.pid_p = R(c.rotation_pid_p); k.pid_i = R(c.rotation_pid_i); k.pid_d = R(c.rotation_pid_d);
  k.force_gain = R(c.force_gain);
  k.virtual_stiffness = R(c.virtual_stiffness);
  k.swing_stiffness_scaler = R(c.swing_stiffness_scaler);
  k.load_stiffness_scaler = R(c.load_stiffness_scaler);
  k.body_velocity_scaler = R(c.body_velocity_scaler);
  k.step_depth = R(c.step_depth);
  k.touchdown_threshold = R(c.touchdown_threshold);
  k.liftoff_threshold = R(c.liftoff_threshold);
  if (c.gravity_aligned_tips && c.joint_count > 3) {
    // WalkController::init (walk_controller.cpp:36-41): the identity tip rotation points the tip's x axis down; every leg's
    // target_tip_pose_.rotation_ keeps it (nothing redefines it without rough-terrain targets)
    Q4<double> q = correct_rotation(from_two_vectors(V3<double>{1.0, 0.0, 0.0}, V3<double>{0.0, 0.0, -1.0}), qidentity<double>());
    k.tip_target_rot[0] = R(q.w); k.tip_target_rot[1] = R(q.x); k.tip_target_rot[2] = R(q.y); k.tip_target_rot[3] = R(q.z);
  }
  for (int a = 0; a < c.auto_poser_count; ++a) {
    k.ap_pos[a][0] = R(c.x_amplitudes[a]); k.ap_pos[a][1] = R(c.y_amplitudes[a]); k.ap_pos[a][2] = R(c.z_amplitudes[a]);
    k.ap_rot[a][0] = R(c.roll_amplitudes[a]); k.ap_rot[a][1] = R(c.pitch_amplitudes[a]); k.ap_rot[a][2] = R(c.yaw_amplitudes[a]);
    k.ap_gravity[a] = R(c.gravity_amplitudes[a]);
  }
  // AdmittanceController::updateAdmittance (admittance_controller.cpp:33-52): 30 classic RK4 steps of the linear ODE
  // x' = A x + b F, A = [[0,1],[-k/m,-c/m]], b = (0,-1/m), h = step_time/30, collapse to x <- P x + q F.
  {
    double m_ = c.virtual_mass, ks = c.virtual_stiffness, z = c.virtual_damping_ratio;
    double cdamp = z * 2 * sqrt(m_ * ks);
    double h = c.integrator_step_time / 30;
    double A[2][2] = {{0.0, 1.0}, {-ks / m_, -cdamp / m_}};
    double b[2] = {0.0, -1.0 / m_};
    auto mul = [](const double X[2][2], const double Y[2][2], double Z[2][2]) {
      double t[2][2];
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) t[i][j] = X[i][0] * Y[0][j] + X[i][1] * Y[1][j];
      std::memcpy(Z, t, sizeof(t));
    };
    double A2[2][2], A3[2][2], A4[2][2];
    mul(A, A, A2); mul(A2, A, A3); mul(A3, A, A4);
    double Ph[2][2], Bh[2][2];  // one RK4 step: x <- Ph x + Bh b F
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        double I = i == j ? 1.0 : 0.0;
        Ph[i][j] = I + h * A[i][j] + h * h / 2 * A2[i][j] + h * h * h / 6 * A3[i][j] + h * h * h * h / 24 * A4[i][j];
        Bh[i][j] = h * (I + h / 2 * A[i][j] + h * h / 6 * A2[i][j] + h * h * h / 24 * A3[i][j]);
      }
    double qh[2] = {Bh[0][0] * b[0] + Bh[0][1] * b[1], Bh[1][0] * b[0] + Bh[1][1] * b[1]};
    double P[2][2] = {{1, 0}, {0, 1}}, q[2] = {0, 0};
    for (int s = 0; s < 30; ++s) {
      double nq[2] = {Ph[0][0] * q[0] + Ph[0][1] * q[1] + qh[0], Ph[1][0] * q[0] + Ph[1][1] * q[1] + qh[1]};
      q[0] = nq[0]; q[1] = nq[1];
      mul(Ph, P, P);
    }
    if (m_ > 0 && ks >= 0) {
      k.adm_P[0] = R(P[0][0]); k.adm_P[1] = R(P[0][1]); k.adm_P[2] = R(P[1][0]); k.adm_P[3] = R(P[1][1]);
      k.adm_q[0] = R(q[0]); k.adm_q[1] = R(q[1]);
    }
  }
}

// WalkController::generateStepCycle (walk_controller.cpp:365) and the phase offsets of generateLimits (:237-278)
inline void compute_step_cycle(const shc_config& c, shc_startup& su) {
  int stance_end = int(c.stance_phase * 0.5);
  int swing_start = stance_end;
  int swing_end = swing_start + c.swing_phase;
  int stance_start = swing_end;
  int base = c.stance_phase + c.swing_phase;
  double swing_ratio = double(c.swing_phase) / double(base);
  double raw = ((1.0 / c.step_frequency) / c.time_delta) / swing_ratio;
  su.period = round_to_even_int(raw / base) * base;
  su.step_frequency = 1.0 / (su.period * c.time_delta);
  int normaliser = su.period / base;
  su.stance_end = stance_end * normaliser;
  su.swing_start = swing_start * normaliser;
  su.swing_end = swing_end * normaliser;
  su.stance_start = stance_start * normaliser;
  su.stance_period = imod(su.stance_end - su.stance_start, su.period);
  su.swing_period = su.swing_end - su.swing_start;
  int base_offset = int(c.phase_offset * normaliser);
  for (int l = 0; l < c.leg_count; ++l) su.phase_offsets[l] = (base_offset * c.offset_multiplier[l]) % su.period;
  // PoseController::setAutoPoseParams (pose_controller.cpp:44-63)
  int base_len;
  double raw_len;
  if (c.pose_frequency == -1.0) {
    base_len = c.stance_phase + c.swing_phase;
    double sr = double(c.swing_phase) / base_len;
    raw_len = ((1.0 / c.step_frequency) / c.time_delta) / sr;
  } else {
    base_len = c.pose_phase_length;
    raw_len = ((1.0 / c.pose_frequency) / c.time_delta);
  }
  if (base_len > 0) {
    su.pose_phase_length = round_to_even_int(raw_len / base_len) * base_len;
    su.pose_normaliser = su.pose_phase_length / base_len;
  }
  su.auto_pose_reference_leg = 0;
  for (int l = 0; l < c.leg_count; ++l)
    if (c.offset_multiplier[l] == 0) su.auto_pose_reference_leg = l;
  // PoseController::directStartup (:463) finishes its stepToPosition after round(time_to_start / dt) calls; the loop()
  // that requested the transition comes on top (state_controller.cpp:254-281)
  su.startup_loops = std::max(1, round_to_int(c.time_to_start / c.time_delta)) + 1;
}

// One Leg::applyIK(simulation) on host joint arrays; returns applyIK's value and the new tip (base_link frame).
template <int D>
double host_apply_ik(const RealConsts<double>& ck, const shc_config& c, int leg, double* q, double* qd, V3<double> desired,
                     bool simulation, V3<double>* tip_robot) {
  return apply_ik_full<D>(ck, leg, q, qd, desired, c.clamp_joint_positions != 0, c.clamp_joint_velocities != 0 && !simulation, tip_robot);
}

template <int D> V3<double> host_fk(const RealConsts<double>& ck, int leg, const double* q) { return leg_fk<D>(ck, leg, q); }

// directStartup + generateWorkspaces + generateWalkspace + generateLimits.
template <int D> void compute_startup(const shc_config& c, const RealConsts<double>& ck, shc_startup& su) {
  const int L = c.leg_count;
  const double dt = c.time_delta;
  const double pi = kPi;
  const StartupParams sp = startup_params(c);

  for (int l = 0; l < L; ++l) {
    // ---- PoseController::directStartup (:463) from the default joint state (model.cpp:1038) ----
    double q[D];
    startup_desired_configuration<D>(ck, sp, l, q);
    for (int j = 0; j < D; ++j) {
      const double q0 = clamp_(0.0, c.joint_min[l][j], c.joint_max[l][j]);  // Joint::default_position_: where a new robot's joints are
      su.default_joint[l][j] = transition_configuration(q0, q[j], sp.startup_iterations, sp.startup_iterations);
    }
  }

  // ---- Leg::generateWorkspace (model.cpp:309): the simple workspace (one plane at height 0, 8 bearings), or in
  // rough-terrain mode the layered workspace, of which the start-up only reads the workplane at the height of the default
  // tips, 0 (Leg::getWorkplane, model.cpp:514: interpolation between the two planes that bound the height) ----
  for (int l = 0; l < L; ++l) {
    if (!c.rough_terrain_mode) {
      double h, radii[SHC_N_BEARINGS];
      workspace_sweep_leg<D>(ck, sp, l, su.default_joint[l], false, 1, 0, 1, &h, radii);
      for (int b = 0; b < SHC_N_BEARINGS; ++b) su.workspace[l][b] = radii[b];
      continue;
    }
    constexpr int kMaxPlanes = 24;
    double heights[kMaxPlanes], radii[kMaxPlanes * SHC_N_BEARINGS];
    const int np = std::min(workspace_sweep_leg<D>(ck, sp, l, su.default_joint[l], true, kMaxPlanes, 0, 1, heights, radii), kMaxPlanes);
    // the reference keeps the planes in a std::map keyed by height: a later plane of the same height replaces an earlier one
    int order[kMaxPlanes], no = 0;
    for (int p = 0; p < np; ++p) {
      int dup = -1;
      for (int k = 0; k < no; ++k)
        if (heights[order[k]] == heights[p]) dup = k;
      if (dup >= 0) order[dup] = p;
      else order[no++] = p;
    }
    std::sort(order, order + no, [&](int a, int b) { return heights[a] < heights[b]; });
    for (int b = 0; b < SHC_N_BEARINGS; ++b) su.workspace[l][b] = 0.0;
    const double height = 0.0;
    if (no == 0 || !(height >= heights[order[0]] && height <= heights[order[no - 1]])) continue;  // outside: empty workplane
    if (no == 1) {
      for (int b = 0; b < SHC_N_BEARINGS; ++b) su.workspace[l][b] = radii[order[0] * SHC_N_BEARINGS + b];
      continue;
    }
    int up = 0;
    while (up < no && !(heights[order[up]] > height)) ++up;  // std::map::upper_bound
    if (up == 0 || up == no) continue;  // (the reference dereferences end() here; cannot happen for height 0 inside the range)
    auto prec3 = [](double v) { return round_to_int(v * 1000.0) / 1000.0; };  // setPrecision(v, 3) (standard_includes.h:142)
    const double hu = prec3(heights[order[up]]), hl = prec3(heights[order[up - 1]]);
    const double i = (height - hl) / (hu - hl);
    for (int b = 0; b < SHC_N_BEARINGS; ++b)
      su.workspace[l][b] = radii[order[up - 1] * SHC_N_BEARINGS + b] * (1.0 - i) + radii[order[up] * SHC_N_BEARINGS + b] * i;
  }

  // ---- WalkController::generateWalkspace (walk_controller.cpp:57) ----
  double ws[SHC_N_BEARINGS];
  bool have[SHC_N_BEARINGS] = {false};
  for (int l = 0; l < L; ++l) {
    int a1 = imod(l + 1, L), a2 = imod(l - 1, L);
    double dx1 = c.stance_x[a1] - c.stance_x[l], dy1 = c.stance_y[a1] - c.stance_y[l];
    double dx2 = c.stance_x[a2] - c.stance_x[l], dy2 = c.stance_y[a2] - c.stance_y[l];
    double dist1 = sqrt(dx1 * dx1 + dy1 * dy1 + 0.0) / 2.0, dist2 = sqrt(dx2 * dx2 + dy2 * dy2 + 0.0) / 2.0;
    double b1 = (atan2(dy1, dx1) / (2.0 * pi)) * 360.0, b2 = (atan2(dy2, dx2) / (2.0 * pi)) * 360.0;
    for (int bi = 0; bi < SHC_N_BEARINGS; ++bi) {
      int bearing = bi * 45;
      int diff1 = abs(imod(int(b1), 360) - bearing), diff2 = abs(imod(int(b2), 360) - bearing);
      double o1 = 2147483647.0, o2 = 2147483647.0;
      if ((diff1 < 90 || diff1 > 270) && dist1 > 0.0) o1 = dist1 / cos(diff1 / 360.0 * 2.0 * pi);
      if ((diff2 < 90 || diff2 > 270) && dist2 > 0.0) o2 = dist2 / cos(diff2 / 360.0 * 2.0 * pi);
      double md = c.overlapping_walkspaces ? 1.0 : std::min(o1, o2);
      md = std::min(md, 1.0);
      if (have[bi] && md < ws[bi]) ws[bi] = md;
      else if (!have[bi]) { ws[bi] = md; have[bi] = true; }
    }
  }
  for (int l = 0; l < L; ++l) {
    // default tip == identity tip at start-up, so the target workplane is the one plane at height 0 and the radius
    // is the workplane radius itself (walk_controller.cpp:137-140)
    for (int bi = 0; bi < SHC_N_BEARINGS; ++bi) {
      double radius = su.workspace[l][bi];
      int opp = imod(bi * 45 + 180, 360) / 45;
      if (radius < ws[bi]) {
        ws[bi] = radius;
        ws[opp] = radius;
      }
    }
  }
  ws[8] = ws[0];
  for (int bi = 0; bi < SHC_N_BEARINGS; ++bi) su.walkspace[bi] = ws[bi];

  // ---- WalkController::generateLimits (walk_controller.cpp:231) ----
  int max_ext = 0;
  for (int l = 0; l < L; ++l) {
    int off = su.phase_offsets[l];
    if (off > su.swing_start && off < su.swing_end) max_ext = std::max(max_ext, su.swing_end - off);
  }
  double t_max = (max_ext + su.stance_period + su.swing_period) * dt;
  double stance_radius = sqrt(c.stance_x[0] * c.stance_x[0] + c.stance_y[0] * c.stance_y[0]);
  for (int bi = 0; bi < SHC_N_BEARINGS; ++bi) {
    double wr = ws[bi];
    double ogr = double(su.stance_period) / su.period;
    double max_speed = (wr * 2.0) / (ogr / su.step_frequency);
    double max_acc = max_speed / t_max;
    double overshoot = 0;
    for (int l = 0; l < L; ++l) {
      double off = su.phase_offsets[l];
      double t = off * dt;
      double tse = t_max - t;
      double v0 = max_acc * tse;
      double sl = v0 * (ogr / su.step_frequency);
      double d0 = -sl / 2.0;
      double d1 = d0 + v0 * t + 0.5 * max_acc * (t * t);
      double d2 = max_speed * (su.stance_period * dt - t);
      overshoot = std::max(overshoot, d1 + d2 - wr);
    }
    double swing_overshoot = 0.5 * max_speed * su.swing_period / (2.0 * su.period * su.step_frequency);
    double scaled = (wr / (wr + overshoot + swing_overshoot)) * wr;
    double mls = (scaled * 2.0) / (ogr / su.step_frequency);
    double mla = mls / t_max;
    double mas = mls / stance_radius;
    double maa = mas / t_max;
    if (wr == 0.0) { mls = 0.0; mla = 2147483647.0; mas = 0.0; maa = 2147483647.0; }
    su.max_linear_speed[bi] = mls;
    su.max_linear_acceleration[bi] = mla;
    su.max_angular_speed[bi] = mas;
    su.max_angular_acceleration[bi] = maa;
  }
}

// Constants that depend on the start-up results (timing, limit tables, stance-span shift).
template <class R> void fill_startup_consts(const shc_config& c, const shc_startup& su, RealConsts<R>& k, IntConsts& ci) {
  const double dt = c.time_delta;
  ci.period = su.period; ci.swing_period = su.swing_period; ci.stance_period = su.stance_period;
  ci.stance_end = su.stance_end; ci.swing_start = su.swing_start; ci.swing_end = su.swing_end; ci.stance_start = su.stance_start;
  // LegStepper::updateTipPosition (walk_controller.cpp:1035-1041), evaluated in double exactly as written
  int swing_iterations = int((double(su.swing_period) / su.period) / (su.step_frequency * dt));
  swing_iterations = round_to_even_int(swing_iterations);
  ci.swing_iterations = swing_iterations;
  {
    // updateWalkPlanePose (pose_controller.cpp:1102-1106) accepts a leg when (n / swing_period) * scaler <= 1.0; both
    // operations are monotonic in n, so the test is an integer window ending at the largest n that still passes
    const double scaler = std::max(1.0, double(c.swing_phase) / c.phase_offset);
    int nmax = -1;
    for (int n = 0; n <= su.swing_period; ++n) {
      volatile double progress = (double)n / (double)su.swing_period;
      volatile double scaled = progress * scaler;
      if (scaled <= 1.0) nmax = n;
    }
    ci.swing_ref_max = nmax;
  }
  k.swing_dt = R(1.0 / (swing_iterations / 2.0));
  auto stance_iter = [&](int mod_period) { return int((double(mod_period) / su.period) / (su.step_frequency * dt)); };
  int std_period = imod(su.stance_end - su.stance_start, su.period);
  if (su.stance_end == su.stance_start) std_period = su.period;
  k.stance_dt_std = R(1.0 / stance_iter(std_period));
  for (int l = 0; l < c.leg_count; ++l) {
    ci.phase_offset[l] = su.phase_offsets[l];
    ci.mod_stance_start[l] = su.phase_offsets[l];
    int mp = imod(su.stance_end - su.phase_offsets[l], su.period);
    if (su.stance_end == su.phase_offsets[l]) mp = su.period;
    k.leg[l].stance_dt_mod = R(1.0 / stance_iter(mp));
    k.leg[l].stride_scaler_mod = R(double(mp) / imod(su.stance_end - su.stance_start, su.period));
    // LegStepper::calculateStanceSpanChange (:949) with the one-plane workspace
    double ssm = c.stance_span_modifier;
    bool pos_y = c.stance_y[l] > 0.0;
    int bearing = (pos_y ^ (ssm > 0.0)) ? 270 : 90;
    ssm *= (pos_y ? 1.0 : -1.0);
    k.leg[l].span_dy = R(su.workspace[l][bearing / 45] * ssm);
  }
  double ogr = double(su.stance_period) / su.period;
  k.stride_scale = R(ogr / su.step_frequency);
  k.inv_swing_period = R(1.0 / su.swing_period);
  k.inv_stance_period = R(1.0 / su.stance_period);
  for (int b = 0; b < SHC_N_BEARINGS; ++b) {
    k.limits[0][b] = R(su.max_linear_speed[b]);
    k.limits[1][b] = R(su.max_angular_speed[b]);
    k.limits[2][b] = R(su.max_linear_acceleration[b]);
    k.limits[3][b] = R(su.max_angular_acceleration[b]);
  }
  ci.pose_phase_length = su.pose_phase_length;
  ci.pose_normaliser = su.pose_normaliser;
  ci.auto_ref_leg = su.auto_pose_reference_leg;
}

// State at the end of the direct start-up (READY -> RUNNING): every robot of a new engine starts here.
template <int D> void initial_state(const shc_config& c, const RealConsts<double>& ck, const shc_startup& su, shc_robot_state& s) {
  std::memset(&s, 0, sizeof(s));
  s.walk_state = WALK_STOPPED;
  s.pose_state = POSE_COMPLETE;
  s.auto_posing_state = POSE_COMPLETE;
  s.walk_plane_normal[2] = 1.0;
  auto ident = [](double* p) { for (int i = 0; i < 7; ++i) p[i] = 0.0; p[3] = 1.0; };
  ident(s.odometry_ideal); ident(s.walk_plane_pose); ident(s.origin_walk_plane_pose); ident(s.manual_pose);
  ident(s.imu_pose); ident(s.inclination_pose); ident(s.auto_pose); ident(s.current_pose);
  ident(s.tip_align_pose); ident(s.origin_tip_align_pose);
  s.walk_plane_pose[2] = c.body_clearance;
  s.origin_walk_plane_pose[2] = c.body_clearance;
  s.current_pose[2] = c.body_clearance;
  for (int l = 0; l < c.leg_count; ++l) {
    shc_leg_state& g = s.legs[l];
    for (int j = 0; j < D; ++j) g.joint_position[j] = su.default_joint[l][j];
    double id[3] = {c.stance_x[l], c.stance_y[l], 0.0};
    for (int k = 0; k < 3; ++k) {
      g.tip_position[k] = id[k];
      g.swing_origin_position[k] = id[k];
      g.stance_origin_position[k] = id[k];
      g.default_tip_position[k] = id[k];
      g.target_tip_position[k] = id[k];
    }
    if (c.gravity_aligned_tips && D > 3)  // LegStepper(): current = origin = target = identity tip pose (walk_controller.cpp:795)
      for (int k = 0; k < 4; ++k) g.tip_rotation[k] = g.origin_tip_rotation[k] = g.target_tip_rotation[k] = ck.tip_target_rot[k];
    g.walk_plane_normal[2] = 1.0;
    g.swing_progress = -1.0;
    g.stance_progress = -1.0;
    g.phase = 0;
    g.step_state = STEP_STANCE;
    V3<double> tip = host_fk<D>(ck, l, g.joint_position);
    g.model_tip_position[0] = tip.x; g.model_tip_position[1] = tip.y; g.model_tip_position[2] = tip.z;
    g.desired_tip_position[0] = tip.x; g.desired_tip_position[1] = tip.y; g.desired_tip_position[2] = tip.z;
    g.ik_result = 1.0;
  }
  if (c.auto_posing && c.pose_frequency != -1.0) {
    // An auto poser with its own cycle keeps cycling while the robot starts up (updateCurrentPose runs in every loop(),
    // pose_controller.cpp:811): replay those loops with the routines the kernel uses — walk state STOPPED, legs in STANCE.
    IntConsts ci;
    std::memset(&ci, 0, sizeof(ci));
    ci.n_posers = c.auto_poser_count;
    ci.pose_sync = 0;
    ci.pose_phase_length = su.pose_phase_length;
    ci.pose_normaliser = su.pose_normaliser;
    for (int a = 0; a < c.auto_poser_count; ++a) { ci.ap_start[a] = c.pose_phase_starts[a]; ci.ap_end[a] = c.pose_phase_ends[a]; }
    for (int l = 0; l < c.leg_count; ++l) { ci.neg_start[l] = c.pose_negation_phase_starts[l]; ci.neg_end[l] = c.pose_negation_phase_ends[l]; }
    int auto_state = POSE_COMPLETE, pflags = 0, pose_phase = 0;
    bool negate[SHC_MAX_LEGS] = {false};
    PoseT<double> ap = pose_identity<double>();
    for (int it = 0; it + 1 < su.startup_loops; ++it) {  // the first loop() finds the robot state UNKNOWN: no pose update (:165)
      auto_state = POSE_STOP_POSING;  // walk state STOPPED (:1146)
      const int master = pose_phase;
      pose_phase = (pose_phase + 1) % su.pose_phase_length;
      ap = auto_posers_update<double>(ci, ck, master, Q4<double>{0.0, 0.0, 0.0, 0.0}, auto_state, pflags);
      for (int l = 0; l < c.leg_count; ++l) leg_auto_pose<double>(ci, ck.leg[l].neg_ratio, l, master, STEP_STANCE, ap, negate[l]);
    }
    s.auto_posing_state = auto_state;
    s.pose_state = auto_state;
    s.pose_phase = pose_phase;
    for (int a = 0; a < c.auto_poser_count; ++a) s.auto_poser_flags[a] = (pflags >> (4 * a)) & 15;
    for (int l = 0; l < c.leg_count; ++l) s.legs[l].negate_auto_pose = negate[l];
    s.auto_pose[0] = ap.p.x; s.auto_pose[1] = ap.p.y; s.auto_pose[2] = ap.p.z;
    s.auto_pose[3] = ap.q.w; s.auto_pose[4] = ap.q.x; s.auto_pose[5] = ap.q.y; s.auto_pose[6] = ap.q.z;
  } else if (c.auto_posing) {
    // The reference runs updateAutoPose during the ~300 start-up cycles (walk state STOPPED, master phase 0).  The
    // latches reach a fixed point after one evaluation (pose_controller.cpp:1359-1371, 1743-1751).
    const int len = su.pose_phase_length, norm = su.pose_normaliser;
    const bool sync = c.pose_frequency == -1.0;
    s.auto_posing_state = POSE_COMPLETE;
    for (int a = 0; a < c.auto_poser_count; ++a) {
      int phase = 0, sp = c.pose_phase_starts[a] * norm, ep = c.pose_phase_ends[a] * norm;
      if (sp > ep) { ep += len; if (phase < sp) phase += len; }
      bool start_check = !sync, end1 = (phase == sp), end2 = (phase == ep && end1), allow = false;
      if (!allow && start_check) { allow = true; end1 = end2 = false; }
      s.auto_poser_flags[a] = (start_check ? 1 : 0) | (end1 ? 2 : 0) | (end2 ? 4 : 0) | (allow ? 8 : 0);
    }
    for (int l = 0; l < c.leg_count; ++l) {
      int sp = c.pose_negation_phase_starts[l] * norm, ep = c.pose_negation_phase_ends[l] * norm, np_ = 0;
      if (sp == 0) sp = len;
      if (ep == 0) ep = len;
      if (sp > ep) { ep += len; if (np_ < sp) np_ += len; }
      bool negate = (np_ == sp);
      if (np_ < sp || np_ > ep) negate = false;
      s.legs[l].negate_auto_pose = negate;
    }
  }
}

}  // namespace shc
