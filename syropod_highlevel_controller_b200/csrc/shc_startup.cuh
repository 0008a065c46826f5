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

// PoseController::directStartup (:463) for one leg from the joint angles q0: LegPoser::stepToPosition (:1571, lift height
// 0) replayed with Leg::applyIK(true) every iteration while the body rises to its clearance, then
// LegPoser::transitionConfiguration (:1476), whose last sample (cubic Bezier (o, o, d, d) at t = num * (1 / num)) is the
// default configuration.  q_out may alias q0.
template <int D>
SHC_HD void direct_startup_leg(const RealConsts<double>& ck, const StartupParams& sp, int l, const double* q0, double* q_out) {
  // Body pose during the start-up cycles: walk-plane pose (0,0,clearance) with identity rotation; manual, inclination and
  // auto poses are identities while STOPPED with no inputs (pose_controller.cpp:811-859).
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  const bool clamp = sp.clamp_joint_positions != 0;
  double q[D], qd[D], qs[D];
#pragma unroll
  for (int j = 0; j < D; ++j) { q[j] = q0[j]; qs[j] = q0[j]; qd[j] = 0.0; }
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
  const double t = num * (1.0 / num), s = 1.0 - t;
#pragma unroll
  for (int j = 0; j < D; ++j) q_out[j] = qs[j] * (s * s * s) + qs[j] * (3.0 * t * s * s) + q[j] * (3.0 * t * t * s) + q[j] * (t * t * t);
}

// Leg::generateWorkspace, bearing-0 pass of one workplane (model.cpp:372-380, 450-453): track from the tip of the default
// configuration qdef to the workplane origin (identity tip raised by `height`) in roundToInt(height_delta / 0.002) steps
// and make the configuration reached the default of the plane's searches (updateDefaultConfiguration).  Returns false when
// the leg's tip is not at its identity position to IK_TOLERANCE (model.cpp:330: the workspace is then empty); only the
// plane at height 0 makes that test.
template <int D>
SHC_HD bool workspace_origin_pass(const RealConsts<double>& ck, const StartupParams& sp, int l, double height, double height_delta,
                                  double* qdef) {
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  V3<double> identity_tip = pose_inverse_transform(body, V3<double>{ck.leg[l].identity_x, ck.leg[l].identity_y, 0.0});
  const V3<double> cur = leg_fk<D>(ck, l, qdef);
  if (height == 0.0 && norm(identity_tip - cur) > 0.005) return false;
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
  return true;
}

// Leg::generateWorkspace, one bearing of one workplane (model.cpp:381-421): from the plane's default configuration walk
// the tip outwards along `bearing_deg` in 2 mm steps (up to MAX_WORKSPACE_RADIUS = 1 m, 500 Leg::applyIK(true) steps) until
// applyIK reports failure; the radius is the distance of the last tip reached from the plane's origin.
template <int D>
SHC_HD double workspace_bearing_search(const RealConsts<double>& ck, const StartupParams& sp, int l, double height, int bearing_deg,
                                       const double* qdef) {
  const PoseT<double> body{{0.0, 0.0, sp.body_clearance}, qidentity<double>()};
  V3<double> identity_tip = pose_inverse_transform(body, V3<double>{ck.leg[l].identity_x, ck.leg[l].identity_y, 0.0});
  identity_tip.z += height;
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) { q[j] = qdef[j]; qd[j] = 0.0; }
  const int n = round_to_int(1.0 / 0.002);
  V3<double> tg = identity_tip;
  const double rad = bearing_deg / 360.0 * 2.0 * kPi;
  tg.x += 1.0 * cos_(rad);
  tg.y += 1.0 * sin_(rad);
  V3<double> tip = identity_tip;
  bool within = true;
  for (int it = 1; it <= n && within; ++it) {
    const double i = double(it) / n;
    within = apply_ik_full<D>(ck, l, q, qd, identity_tip * (1.0 - i) + tg * i, sp.clamp_joint_positions != 0, false, &tip) != 0.0;
  }
  return norm(tip - identity_tip);
}

}  // namespace shc
