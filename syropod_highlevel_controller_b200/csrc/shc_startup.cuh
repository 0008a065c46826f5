// shc_startup.cuh — the reference's start-up computations as host/device routines over ONE leg, shared by the host
// start-up of an engine (csrc/shc_host.cuh: constants of a new engine) and by the batched device kernels of SURVEY.md §8(f)
// (csrc/shc_engine.cu: shc_direct_startup, shc_generate_workspaces):
//   PoseController::directStartup      pose_controller.cpp:463  (LegPoser::stepToPosition :1571, transitionConfiguration :1476)
//   Leg::generateWorkspace             model.cpp:309-510        (IK search :390-421: up to 500 Leg::applyIK(true) steps)
// They reuse the chain / DLS routines of the control cycle (shc_cycle.cuh); nothing here is shared with oracle/.
#pragma once
#include "shc_cycle.cuh"

namespace shc {

// The handful of scalar parameters the start-up needs beyond the constants block.
struct StartupParams {
  double body_clearance, dt;
  int startup_iterations;      // max(1, roundToInt(time_to_start / dt))        (pose_controller.cpp:1583)
  int clamp_joint_positions;   // clamp_joint_velocities is irrelevant: every start-up IK step is a simulation step
};
inline StartupParams startup_params(const shc_config& c) {
  StartupParams p;
  p.body_clearance = c.body_clearance;
  p.dt = c.time_delta;
  p.startup_iterations = max_(1, round_to_int(c.time_to_start / c.time_delta));
  p.clamp_joint_positions = c.clamp_joint_positions;
  return p;
}

template <int D> SHC_HD V3<double> leg_fk(const RealConsts<double>& ck, int leg, const double* q) {
  Chain<double, D> ch;
  leg_chain<double, D>(ck.leg[leg], q, ch);
  return t1_rotate(ck.leg[leg], ch.tip) + V3<double>{ck.leg[leg].t1p[0], ck.leg[leg].t1p[1], ck.leg[leg].t1p[2]};
}

// One Leg::applyIK (model.cpp:861, position only) on joint arrays; returns applyIK's value and the new tip (base_link).
template <int D>
SHC_HD double apply_ik_full(const RealConsts<double>& ck, int leg, double* q, double* qd, V3<double> desired, bool clamp_positions,
                            bool clamp_velocities, V3<double>* tip_robot) {
  const LegConsts<double>& lc = ck.leg[leg];
  Chain<double, D> ch;
  leg_chain<double, D>(lc, q, ch);
  V3<double> des_leg;
  apply_ik_step<double, D>(ck, lc, ch, q, qd, desired, clamp_positions, clamp_velocities, &des_leg);
  Chain<double, D> ch2;
  leg_chain<double, D>(lc, q, ch2);
  if (tip_robot) *tip_robot = t1_rotate(lc, ch2.tip) + V3<double>{lc.t1p[0], lc.t1p[1], lc.t1p[2]};
  return ik_result_value<double, D>(lc, ch2, q, des_leg);
}

// PoseController::directStartup (:463), first half, for one leg: the desired configuration is found on a COPY of the leg
// that starts from the default joint positions (test_leg.init(true), :475) — LegPoser::stepToPosition (:1571, lift height
// 0) replayed with Leg::applyIK(true) every iteration while the body rises to its clearance.  It does not depend on where
// the real leg's joints are.
template <int D>
SHC_HD void startup_desired_configuration(const RealConsts<double>& ck, const StartupParams& sp, int l, double* q_out) {
  // Body pose during the start-up cycles: walk-plane pose (0,0,clearance) with identity rotation; manual, inclination and
  // auto poses are identities while STOPPED with no inputs (pose_controller.cpp:811-859).
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  const bool clamp = sp.clamp_joint_positions != 0;
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    q[j] = clamp_(0.0, ck.leg[l].jmin[j], ck.leg[l].jmax[j]);  // Joint::default_position_ (model.cpp:1038)
    qd[j] = 0.0;
  }
  const V3<double> origin = leg_fk<D>(ck, l, q);
  const V3<double> target{ck.leg[l].identity_x, ck.leg[l].identity_y, 0.0};  // default tip pose, walk-plane frame
  const V3<double> pdelta = origin - pose_inverse_transform(body, target);
  const int num = sp.startup_iterations;
  if (norm(pdelta) > 0.01) {  // TIP_TOLERANCE; lift_height is 0
    const double delta_t = 1.0 / num;
    const int half = num / 2;
    const V3<double> o2t = origin - target;
    const V3<double> n1[5] = {origin, origin, origin, target + o2t * 0.75, target + o2t * 0.5};
    const V3<double> n2[5] = {target + o2t * 0.5, target + o2t * 0.25, target, target, target};
    for (int count = 1; count <= num; ++count) {
      const double ratio = double(count - 1) / double(num);
      const PoseT<double> dpose = pose_interpolate(pose_identity<double>(), smooth_step(ratio), body);
      const int sic = (count + (num - 1)) % num + 1;
      V3<double> tip;
      if (sic <= half) tip = quartic_bezier(n1, sic * delta_t * 2.0);
      else tip = quartic_bezier(n2, (sic - half) * delta_t * 2.0);
      apply_ik_full<D>(ck, l, q, qd, pose_inverse_transform(dpose, tip), clamp, false, nullptr);
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) q_out[j] = q[j];
}

// LegPoser::transitionConfiguration (:1476): iteration `it` (1 .. num) of the joint-space move from the configuration the
// leg was in when the start-up began to the desired one — cubic Bezier with nodes (o, o, d, d) at t = it * (1 / num).
SHC_HD double transition_configuration(double origin, double desired, int it, int num) {
  const double t = it * (1.0 / num), s = 1.0 - t;
  return origin * (s * s * s) + origin * (3.0 * t * s * s) + desired * (3.0 * t * t * s) + desired * (t * t * t);
}

// Leg::generateWorkspace, bearing-0 pass of one workplane (model.cpp:372-380, 450-453): track from the tip of the default
// configuration qdef to the workplane origin (identity tip raised by `height`) in roundToInt(height_delta / 0.002) steps
// and make the configuration reached the default of the plane's searches (updateDefaultConfiguration).
template <int D>
SHC_HD void workspace_origin_pass(const RealConsts<double>& ck, const StartupParams& sp, int l, double height, double height_delta,
                                  double* qdef) {
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  V3<double> identity_tip = pose_inverse_transform(body, V3<double>{ck.leg[l].identity_x, ck.leg[l].identity_y, 0.0});
  const V3<double> cur = leg_fk<D>(ck, l, qdef);
  identity_tip.z += height;
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) { q[j] = qdef[j]; qd[j] = 0.0; }
  const int n = max_(1, round_to_int(height_delta / 0.002));  // MAX_POSITION_DELTA
  bool within = true;
  for (int it = 1; it <= n && within; ++it) {
    const double i = double(it) / n;
    within = apply_ik_full<D>(ck, l, q, qd, cur * (1.0 - i) + identity_tip * i, sp.clamp_joint_positions != 0, false, nullptr) != 0.0;
  }
#pragma unroll
  for (int j = 0; j < D; ++j) qdef[j] = q[j];
}

// A straight-line search of Leg::generateWorkspace (model.cpp:381-421) from the configuration qdef: the tip is walked from
// `origin` towards `origin + dir` (|dir| = MAX_WORKSPACE_RADIUS = 1 m) in 500 steps of 2 mm until Leg::applyIK(true) reports
// failure; returns the distance of the last tip reached from `reference`.
template <int D>
SHC_HD double workspace_line_search(const RealConsts<double>& ck, const StartupParams& sp, int l, V3<double> origin, V3<double> dir,
                                    V3<double> reference, const double* qdef) {
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) { q[j] = qdef[j]; qd[j] = 0.0; }
  const int n = round_to_int(1.0 / 0.002);
  const V3<double> tg = origin + dir;
  V3<double> tip = origin;
  bool within = true;
  for (int it = 1; it <= n && within; ++it) {
    const double i = double(it) / n;
    within = apply_ik_full<D>(ck, l, q, qd, origin * (1.0 - i) + tg * i, sp.clamp_joint_positions != 0, false, &tip) != 0.0;
  }
  return norm(tip - reference);
}

// Leg::generateWorkspace (model.cpp:309-510) for one leg from its default configuration qdef0.  full = false: the simple
// workspace (one plane at height 0 — what the engine's constants use); full = true: the layered workspace the reference
// builds in rough-terrain mode (lower / upper reach, then WORKSPACE_LAYERS = 10 planes between them, the configuration of a
// plane's origin seeding the next plane).  The eight bearing searches of a plane are independent: lane `lane` of `nlanes`
// cooperating lanes runs bearings lane+1, lane+1+nlanes, ...; everything sequential (origin passes, reach searches) is
// computed redundantly by every lane, so the lanes never communicate.  Planes are written in generation order: heights[p],
// radii[p][0..8] (bearing 45 * k; [0] = [8]); returns the number of planes (every lane returns the same).
template <int D>
SHC_HD int workspace_sweep_leg(const RealConsts<double>& ck, const StartupParams& sp, int l, const double* qdef0, bool full,
                               int max_planes, int lane, int nlanes, double* heights, double* radii) {
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  const V3<double> identity0 = pose_inverse_transform(body, V3<double>{ck.leg[l].identity_x, ck.leg[l].identity_y, 0.0});
  double qdef[D];
#pragma unroll
  for (int j = 0; j < D; ++j) qdef[j] = qdef0[j];
  int np = 0;
  auto put_plane = [&](double h, double fill) {
    if (np < max_planes) {
      heights[np] = h;
      for (int b = 0; b < SHC_N_BEARINGS; ++b) radii[np * SHC_N_BEARINGS + b] = fill;
    }
    return np++;
  };
  if (norm(identity0 - leg_fk<D>(ck, l, qdef)) > 0.005) {  // IK_TOLERANCE (model.cpp:330)
    put_plane(0.0, 0.0);
    return np;
  }
  double min_h = 0.0, delta = 0.1, search_h = 0.0;  // MAX_WORKSPACE_RADIUS / WORKSPACE_LAYERS
  if (full) {
    // lower, then upper reach of the leg straight below / above its identity tip (model.cpp:355-371, 423-441)
    const double down = workspace_line_search<D>(ck, sp, l, identity0, V3<double>{0.0, 0.0, -1.0}, identity0, qdef);
    min_h = -down;
    put_plane(min_h, 0.0);
    const double up = workspace_line_search<D>(ck, sp, l, identity0, V3<double>{0.0, 0.0, 1.0}, identity0, qdef);
    const double max_h = up;
    delta = (max_h - min_h) / 10;
    const int upper_levels = int(abs_(max_h) / delta);
    search_h = upper_levels * delta;
    put_plane(max_h, 0.0);
  }
  while (true) {
    const int p = put_plane(search_h, 1.0);  // max_workplane: MAX_WORKSPACE_RADIUS everywhere until searched
    V3<double> origin = identity0;
    origin.z += search_h;
    workspace_origin_pass<D>(ck, sp, l, search_h, delta, qdef);  // bearing 0 + updateDefaultConfiguration
    for (int b = 1 + lane; b <= 8; b += nlanes) {
      const double rad = (45 * b) / 360.0 * 2.0 * kPi;
      const double r = workspace_line_search<D>(ck, sp, l, origin, V3<double>{1.0 * cos_(rad), 1.0 * sin_(rad), 0.0}, origin, qdef);
      if (p < max_planes) {
        radii[p * SHC_N_BEARINGS + b] = r;
        if (b == 8) radii[p * SHC_N_BEARINGS] = r;
      }
    }
    search_h -= delta;
    if (!full || !(search_h >= min_h)) break;
  }
  return np;
}

}  // namespace shc
