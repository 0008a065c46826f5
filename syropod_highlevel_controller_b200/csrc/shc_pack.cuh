// shc_pack.cuh — host-side glue shared by the engine (shc_engine.cu) and by the test-only host emulator of the cycle
// code (tests/cpp/shc_emu.cu): plane layout of an engine, state record <-> planes packing, start-up constants.  Everything
// here is plain host C++ over `E`, any struct with the members {cfg, su, c, precision, n, n_pad}.
#pragma once
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../include/shc_b200.h"
#include "shc_host.cuh"

namespace shc {

inline void layout(const shc_config& cfg, int n, IntConsts& ci) {
  ci.L = cfg.leg_count;
  ci.D = cfg.joint_count;
  ci.n_robots = n;
  ci.n_pad = (n + 31) / 32 * 32;
  const int D = cfg.joint_count;
  const bool imu = cfg.imu_posing || cfg.inclination_posing;
  const bool adm = cfg.admittance_control || cfg.use_joint_effort;
  int s = RS_COUNT;
  ci.offS_imu = s;
  if (imu) s += IMU_COUNT;
  ci.offS_auto = s;
  if (cfg.auto_posing) s += AUTO_COUNT;
  ci.tip_mode = !cfg.gravity_aligned_tips ? TIP_NONE : D <= 3 ? TIP_ALIGN_POSE : TIP_ROTATION;
  ci.offS_tip = s;
  if (ci.tip_mode == TIP_ALIGN_POSE) s += TA_COUNT;
  const LegOff lo(D);
  ci.frontS_leg = adm ? -lo.ADM_X : 0;
  ci.offS_leg = s + ci.frontS_leg;  // plane of leg 0's first joint position; the staged admittance planes sit in front
  ci.strideS_leg = lo.COUNT + (adm ? ADM_COUNT : 0);
  ci.tipS_leg = ci.strideS_leg - ci.frontS_leg;  // behind the appended admittance planes
  if (ci.tip_mode == TIP_ROTATION) ci.strideS_leg += TR_COUNT;
  ci.rough_terrain = cfg.rough_terrain_mode;
  ci.roughS_leg = ci.strideS_leg - ci.frontS_leg;
  if (ci.rough_terrain) ci.strideS_leg += RT_COUNT;
  ci.nS = s + ci.strideS_leg * cfg.leg_count;
  ci.offD_leg = RD_COUNT;
  ci.strideD_leg = LD_COUNT;
  ci.nD = RD_COUNT + LD_COUNT * cfg.leg_count;
  int i = RI_COUNT;
  ci.offI_auto = i;
  if (cfg.auto_posing) i += AI_COUNT;
  ci.offI_leg = i;
  ci.strideI_leg = LI_COUNT;
  ci.nI = i + LI_COUNT * cfg.leg_count;
  ci.manual_posing = cfg.manual_posing;
  ci.auto_posing = cfg.auto_posing;
  ci.inclination_posing = cfg.inclination_posing;
  ci.imu_posing = cfg.imu_posing;
  ci.admittance_control = cfg.admittance_control;
  ci.dynamic_stiffness = cfg.admittance_control && cfg.dynamic_stiffness;
  ci.use_joint_effort = cfg.use_joint_effort;
  ci.clamp_joint_positions = cfg.clamp_joint_positions;
  ci.clamp_joint_velocities = cfg.clamp_joint_velocities;
  ci.velocity_input_mode = cfg.velocity_input_mode;
  ci.force_normal_touchdown = cfg.force_normal_touchdown;
  ci.n_posers = cfg.auto_poser_count;
  ci.pose_sync = cfg.pose_frequency == -1.0;
  for (int a = 0; a < cfg.auto_poser_count; ++a) {
    ci.ap_start[a] = cfg.pose_phase_starts[a];
    ci.ap_end[a] = cfg.pose_phase_ends[a];
  }
  for (int l = 0; l < cfg.leg_count; ++l) {
    ci.neg_start[l] = cfg.pose_negation_phase_starts[l];
    ci.neg_end[l] = cfg.pose_negation_phase_ends[l];
  }
}

struct HostPlanes {
  std::vector<double> s;  // storage planes widened to double
  std::vector<double> d;
  std::vector<int> i;
};

inline int prog_num(double progress, int den) {
  if (progress < 0.0) return -1;
  int n = (int)(progress * den + 0.5);
  return std::min(std::max(n, 0), den);
}

// Packs records in[0..n) into lanes [0, n_fill): lanes past n repeat the last record (padding robots of the last tile
// run through the kernel like any other lane, so they must hold a valid state).
template <class E> void pack(const E* e, const shc_robot_state* in, size_t n, HostPlanes& h, size_t n_fill) {
  const IntConsts& ci = e->c.i;
  const size_t np = ci.n_pad;
  const int D = ci.D, L = ci.L;
  const LegOff lo(D);
  const bool imu = e->cfg.imu_posing || e->cfg.inclination_posing;
  const bool adm = e->cfg.admittance_control || e->cfg.use_joint_effort;
  // tile-major planes: [tile][plane][32 lanes] (shc_layout.h)
  auto S = [&](int plane, size_t r) -> double& { return h.s[((r >> 5) * ci.nS + plane) * 32 + (r & 31)]; };
  auto Dd = [&](int plane, size_t r) -> double& { return h.d[((r >> 5) * ci.nD + plane) * 32 + (r & 31)]; };
  auto I = [&](int plane, size_t r) -> int& { return h.i[((r >> 5) * ci.nI + plane) * 32 + (r & 31)]; };
  for (size_t r = 0; r < n_fill; ++r) {
    const shc_robot_state& s = in[r < n ? r : n - 1];
    S(RS_VEL, r) = s.desired_linear_velocity[0];
    S(RS_VEL + 1, r) = s.desired_linear_velocity[1];
    S(RS_ANGVEL, r) = s.desired_angular_velocity;
    for (int k = 0; k < 3; ++k) {
      S(RS_WPL + k, r) = s.walk_plane[k];
      S(RS_WPN + k, r) = s.walk_plane_normal[k];
      Dd(RD_ODOMP + k, r) = s.odometry_ideal[k];
    }
    for (int k = 0; k < 4; ++k) S(RS_ODOMQ + k, r) = s.odometry_ideal[3 + k];
    for (int k = 0; k < 7; ++k) {
      S(RS_WPP + k, r) = s.walk_plane_pose[k];
      S(RS_OWPP + k, r) = s.origin_walk_plane_pose[k];
      S(RS_MAN + k, r) = s.manual_pose[k];
    }
    if (imu) {
      for (int k = 0; k < 4; ++k) S(ci.offS_imu + IMU_Q + k, r) = s.imu_pose[3 + k];
      for (int k = 0; k < 3; ++k) {
        S(ci.offS_imu + IMU_ABS + k, r) = s.rotation_absement_error[k];
        S(ci.offS_imu + IMU_VEL + k, r) = s.rotation_velocity_error[k];
        S(ci.offS_imu + IMU_POS + k, r) = s.rotation_position_error[k];
      }
      S(ci.offS_imu + IMU_INCL, r) = s.inclination_pose[0];
      S(ci.offS_imu + IMU_INCL + 1, r) = s.inclination_pose[1];
    }
    if (e->cfg.auto_posing) {
      for (int k = 0; k < 7; ++k) S(ci.offS_auto + AUTO_POSE + k, r) = s.auto_pose[k];
      int pf = 0;
      for (int a = 0; a < e->cfg.auto_poser_count; ++a) pf |= (s.auto_poser_flags[a] & 15) << (4 * a);
      I(ci.offI_auto + AI_FLAGS, r) = pf;
      I(ci.offI_auto + AI_PHASE, r) = s.pose_phase;
    }
    if (ci.tip_mode == TIP_ALIGN_POSE)
      for (int k = 0; k < 7; ++k) {
        S(ci.offS_tip + TA_POSE + k, r) = s.tip_align_pose[k];
        S(ci.offS_tip + TA_ORIGIN + k, r) = s.origin_tip_align_pose[k];
      }
    I(RI_BITS, r) = (s.walk_state & 3) | ((s.legs_at_correct_phase & 15) << 2) | ((s.legs_completed_first_step & 15) << 6) |
                    ((s.return_to_default_attempted & 1) << 10) | ((s.pose_state & 3) << 11) | ((s.auto_posing_state & 3) << 13) |
                    (1 << RB_PLANE_CHANGED) |  // a state written from outside: the first cycle reads the legs' saved planes
                    (1u << RB_PLANE_STALE);    // ... and fits the walk plane to the default tips whatever they are
    {
      const double* m = s.manual_pose;
      const bool ident = m[0] == 0.0 && m[1] == 0.0 && m[2] == 0.0 && m[3] == 1.0 && m[4] == 0.0 && m[5] == 0.0 && m[6] == 0.0;
      if (ident) I(RI_BITS, r) |= 1 << RB_MANUAL_IDENTITY;
    }
    for (int l = 0; l < L; ++l) {
      const shc_leg_state& g = s.legs[l];
      const int sb = ci.offS_leg + l * ci.strideS_leg;
      for (int j = 0; j < D; ++j) {
        S(sb + lo.Q + j, r) = g.joint_position[j];
        S(sb + lo.QD + j, r) = g.joint_velocity[j];
      }
      for (int k = 0; k < 3; ++k) {
        S(sb + lo.TIPVEL + k, r) = g.tip_velocity[k];
        S(sb + lo.SWO_P + k, r) = g.swing_origin_position[k];
        S(sb + lo.SWO_V + k, r) = g.swing_origin_velocity[k];
        S(sb + lo.STO_P + k, r) = g.stance_origin_position[k];
        S(sb + lo.DEF + k, r) = g.default_tip_position[k];
        S(sb + lo.TGT + k, r) = g.target_tip_position[k];
        S(sb + lo.STRIDE + k, r) = g.stride_vector[k];
        S(sb + lo.WP + k, r) = g.walk_plane[k];
        S(sb + lo.WPN + k, r) = g.walk_plane_normal[k];
        Dd(ci.offD_leg + l * ci.strideD_leg + LD_TIP + k, r) = g.tip_position[k];
      }
      if (adm) {
        S(sb + lo.ADM_X, r) = g.admittance_state[0];
        S(sb + lo.ADM_X + 1, r) = g.admittance_state[1];
        for (int k = 0; k < 3; ++k) {
          S(sb + lo.ADM_DELTA + k, r) = g.admittance_delta[k];
          S(sb + lo.ADM_FORCE + k, r) = g.tip_force_calculated[k];
        }
        S(sb + lo.STIFF, r) = g.virtual_stiffness;
      }
      if (ci.tip_mode == TIP_ROTATION)
        for (int k = 0; k < 4; ++k) {
          S(sb + ci.tipS_leg + TR_CUR + k, r) = g.tip_rotation[k];
          S(sb + ci.tipS_leg + TR_ORIGIN + k, r) = g.origin_tip_rotation[k];
        }
      if (ci.rough_terrain) {
        for (int k = 0; k < 3; ++k) S(sb + ci.roughS_leg + RT_STEP_PLANE + k, r) = g.step_plane_position[k];
        for (int k = 0; k < 7; ++k) {
          S(sb + ci.roughS_leg + RT_ET_POSE + k, r) = g.external_target_pose[k];
          S(sb + ci.roughS_leg + RT_ET_TF + k, r) = g.external_target_transform[k];
          S(sb + ci.roughS_leg + RT_ED_POSE + k, r) = g.external_default_pose[k];
          S(sb + ci.roughS_leg + RT_ED_TF + k, r) = g.external_default_transform[k];
        }
        S(sb + ci.roughS_leg + RT_ET_CLR, r) = g.external_target_clearance;
      }
      const int ib = ci.offI_leg + l * ci.strideI_leg;
      I(ib + LI_BITS, r) = (g.phase & 0xffff) | ((g.step_state & 3) << 16) | ((g.at_correct_phase ? 1 : 0) << 18) |
                           ((g.completed_first_step ? 1 : 0) << 19) | ((g.negate_auto_pose ? 1 : 0) << 20) |
                           ((ci.rough_terrain && g.step_plane_defined ? 1 : 0) << LB_STEP_PLANE) |
                           ((ci.rough_terrain && g.touchdown_detection ? 1 : 0) << LB_TOUCHDOWN) |
                           ((ci.rough_terrain && g.external_target_defined ? 1 : 0) << LB_EXT_TARGET) |
                           ((ci.rough_terrain && g.external_target_odom_frame ? 1 : 0) << LB_EXT_ODOM) |
                           ((ci.rough_terrain && g.external_default_defined ? 1 : 0) << LB_EXT_DEFAULT);
      int sn = prog_num(g.swing_progress, ci.swing_period), tn = prog_num(g.stance_progress, ci.stance_period);
      I(ib + LI_PROG, r) = (sn & 0xffff) | ((tn & 0xffff) << 16);
    }
  }
}

// Unpacks the records of the robots [r0, r0 + n) of `h` (robot indices relative to the first tile held in h) into out[0..n).
template <int D, class E> void unpack(const E* e, const HostPlanes& h, shc_robot_state* out, size_t n, size_t r0 = 0) {
  const IntConsts& ci = e->c.i;
  const RealConsts<double>& ck = e->c.d;
  const size_t np = ci.n_pad;
  const int L = ci.L;
  const bool imu = e->cfg.imu_posing || e->cfg.inclination_posing;
  const bool adm = e->cfg.admittance_control || e->cfg.use_joint_effort;
  auto S = [&](int plane, size_t r) { return h.s[((r >> 5) * ci.nS + plane) * 32 + (r & 31)]; };
  auto Dd = [&](int plane, size_t r) { return h.d[((r >> 5) * ci.nD + plane) * 32 + (r & 31)]; };
  auto I = [&](int plane, size_t r) { return h.i[((r >> 5) * ci.nI + plane) * 32 + (r & 31)]; };
  for (size_t r = r0; r < r0 + n; ++r) {
    shc_robot_state& s = out[r - r0];
    std::memset(&s, 0, sizeof(s));
    s.desired_linear_velocity[0] = S(RS_VEL, r);
    s.desired_linear_velocity[1] = S(RS_VEL + 1, r);
    s.desired_angular_velocity = S(RS_ANGVEL, r);
    for (int k = 0; k < 3; ++k) {
      s.walk_plane[k] = S(RS_WPL + k, r);
      s.walk_plane_normal[k] = S(RS_WPN + k, r);
      s.odometry_ideal[k] = Dd(RD_ODOMP + k, r);
    }
    for (int k = 0; k < 4; ++k) s.odometry_ideal[3 + k] = S(RS_ODOMQ + k, r);
    for (int k = 0; k < 7; ++k) {
      s.walk_plane_pose[k] = S(RS_WPP + k, r);
      s.origin_walk_plane_pose[k] = S(RS_OWPP + k, r);
      s.manual_pose[k] = S(RS_MAN + k, r);
    }
    auto ident = [](double* p) { for (int i = 0; i < 7; ++i) p[i] = 0.0; p[3] = 1.0; };
    ident(s.imu_pose); ident(s.inclination_pose); ident(s.auto_pose); ident(s.tip_align_pose); ident(s.origin_tip_align_pose);
    if (ci.tip_mode == TIP_ALIGN_POSE)
      for (int k = 0; k < 7; ++k) {
        s.tip_align_pose[k] = S(ci.offS_tip + TA_POSE + k, r);
        s.origin_tip_align_pose[k] = S(ci.offS_tip + TA_ORIGIN + k, r);
      }
    if (!e->cfg.manual_posing) ident(s.manual_pose);
    if (imu) {
      for (int k = 0; k < 4; ++k) s.imu_pose[3 + k] = S(ci.offS_imu + IMU_Q + k, r);
      for (int k = 0; k < 3; ++k) {
        s.rotation_absement_error[k] = S(ci.offS_imu + IMU_ABS + k, r);
        s.rotation_velocity_error[k] = S(ci.offS_imu + IMU_VEL + k, r);
        s.rotation_position_error[k] = S(ci.offS_imu + IMU_POS + k, r);
      }
      s.inclination_pose[0] = S(ci.offS_imu + IMU_INCL, r);
      s.inclination_pose[1] = S(ci.offS_imu + IMU_INCL + 1, r);
    }
    if (e->cfg.auto_posing) {
      for (int k = 0; k < 7; ++k) s.auto_pose[k] = S(ci.offS_auto + AUTO_POSE + k, r);
      int pf = I(ci.offI_auto + AI_FLAGS, r);
      for (int a = 0; a < e->cfg.auto_poser_count; ++a) s.auto_poser_flags[a] = (pf >> (4 * a)) & 15;
      s.pose_phase = I(ci.offI_auto + AI_PHASE, r);
    }
    int rb = I(RI_BITS, r);
    s.walk_state = rb & 3;
    s.legs_at_correct_phase = (rb >> 2) & 15;
    s.legs_completed_first_step = (rb >> 6) & 15;
    s.return_to_default_attempted = (rb >> 10) & 1;
    s.pose_state = (rb >> 11) & 3;
    s.auto_posing_state = (rb >> 13) & 3;
    s.status_flags = (rb >> RB_STATUS_SHIFT) & RB_STATUS_MASK;
    // Model::current_pose_ is recomputed every cycle from the stored sub-poses (pose_controller.cpp:811-859)
    {
      PoseT<double> p = pose_identity<double>();
      auto rd = [](const double* a) { return PoseT<double>{{a[0], a[1], a[2]}, {a[3], a[4], a[5], a[6]}}; };
      p = pose_add(p, rd(s.walk_plane_pose));
      if (e->cfg.manual_posing) p = pose_add(p, rd(s.manual_pose));
      if (e->cfg.inclination_posing) p = pose_add(p, rd(s.inclination_pose));
      if (e->cfg.imu_posing) p = pose_add(p, rd(s.imu_pose));
      else if (e->cfg.auto_posing) p = pose_add(p, rd(s.auto_pose));
      if (ci.tip_mode == TIP_ALIGN_POSE) p = pose_add(p, rd(s.tip_align_pose));
      s.current_pose[0] = p.p.x; s.current_pose[1] = p.p.y; s.current_pose[2] = p.p.z;
      s.current_pose[3] = p.q.w; s.current_pose[4] = p.q.x; s.current_pose[5] = p.q.y; s.current_pose[6] = p.q.z;
    }
    for (int l = 0; l < L; ++l) {
      shc_leg_state& g = s.legs[l];
      const int sb = ci.offS_leg + l * ci.strideS_leg;
      for (int j = 0; j < D; ++j) {
        g.joint_position[j] = S(sb + LegS<D>::Q + j, r);
        g.joint_velocity[j] = S(sb + LegS<D>::QD + j, r);
      }
      using LS = LegS<D>;
      for (int k = 0; k < 3; ++k) {
        g.tip_velocity[k] = S(sb + LS::TIPVEL + k, r);
        g.swing_origin_position[k] = S(sb + LS::SWO_P + k, r);
        g.swing_origin_velocity[k] = S(sb + LS::SWO_V + k, r);
        g.stance_origin_position[k] = S(sb + LS::STO_P + k, r);
        g.default_tip_position[k] = S(sb + LS::DEF + k, r);
        g.target_tip_position[k] = S(sb + LS::TGT + k, r);
        g.stride_vector[k] = S(sb + LS::STRIDE + k, r);
        g.walk_plane[k] = S(sb + LS::WP + k, r);
        g.walk_plane_normal[k] = S(sb + LS::WPN + k, r);
        g.tip_position[k] = Dd(ci.offD_leg + l * ci.strideD_leg + LD_TIP + k, r);
      }
      if (adm) {
        g.admittance_state[0] = S(sb + LS::ADM_X, r);
        g.admittance_state[1] = S(sb + LS::ADM_X + 1, r);
        for (int k = 0; k < 3; ++k) {
          g.admittance_delta[k] = S(sb + LS::ADM_DELTA + k, r);
          g.tip_force_calculated[k] = S(sb + LS::ADM_FORCE + k, r);
        }
        g.virtual_stiffness = S(sb + LS::STIFF, r);
      }
      if (ci.tip_mode == TIP_ROTATION)
        for (int k = 0; k < 4; ++k) {
          g.tip_rotation[k] = S(sb + ci.tipS_leg + TR_CUR + k, r);
          g.origin_tip_rotation[k] = S(sb + ci.tipS_leg + TR_ORIGIN + k, r);
          g.target_tip_rotation[k] = ck.tip_target_rot[k];
        }
      const int ib = ci.offI_leg + l * ci.strideI_leg;
      int b = I(ib + LI_BITS, r), pg = I(ib + LI_PROG, r);
      g.phase = b & 0xffff;
      g.step_state = (b >> 16) & 3;
      g.at_correct_phase = (b >> 18) & 1;
      g.completed_first_step = (b >> 19) & 1;
      g.negate_auto_pose = (b >> 20) & 1;
      if (ci.rough_terrain) {
        g.step_plane_defined = (b >> LB_STEP_PLANE) & 1;
        g.touchdown_detection = (b >> LB_TOUCHDOWN) & 1;
        if (g.step_plane_defined)
          for (int k = 0; k < 3; ++k) g.step_plane_position[k] = S(sb + ci.roughS_leg + RT_STEP_PLANE + k, r);
        g.external_target_defined = (b >> LB_EXT_TARGET) & 1;
        g.external_target_odom_frame = (b >> LB_EXT_ODOM) & 1;
        g.external_default_defined = (b >> LB_EXT_DEFAULT) & 1;
        for (int k = 0; k < 7; ++k) {
          g.external_target_pose[k] = S(sb + ci.roughS_leg + RT_ET_POSE + k, r);
          g.external_target_transform[k] = S(sb + ci.roughS_leg + RT_ET_TF + k, r);
          g.external_default_pose[k] = S(sb + ci.roughS_leg + RT_ED_POSE + k, r);
          g.external_default_transform[k] = S(sb + ci.roughS_leg + RT_ED_TF + k, r);
        }
        g.external_target_clearance = S(sb + ci.roughS_leg + RT_ET_CLR, r);
      }
      int sn = (int)(short)(pg & 0xffff), tn = (int)(short)((pg >> 16) & 0xffff);
      g.swing_progress = sn < 0 ? -1.0 : (double)sn / (double)ci.swing_period;
      g.stance_progress = tn < 0 ? -1.0 : (double)tn / (double)ci.stance_period;
      V3<double> tip = host_fk<D>(ck, l, g.joint_position);  // Leg::current_tip_pose_ = FK(joint positions)
      g.model_tip_position[0] = tip.x; g.model_tip_position[1] = tip.y; g.model_tip_position[2] = tip.z;
      {  // Leg::desired_tip_pose_ = poser tip pose (+ admittance delta); the per-leg auto pose is not stored
        PoseT<double> cp{{s.current_pose[0], s.current_pose[1], s.current_pose[2]},
                         {s.current_pose[3], s.current_pose[4], s.current_pose[5], s.current_pose[6]}};
        V3<double> des = pose_inverse_transform(cp, V3<double>{g.tip_position[0], g.tip_position[1], g.tip_position[2]});
        g.desired_tip_position[0] = des.x + g.admittance_delta[0];
        g.desired_tip_position[1] = des.y + g.admittance_delta[1];
        g.desired_tip_position[2] = des.z + g.admittance_delta[2];
        // Leg::applyIK's return value (model.cpp:845-856, 916-929)
        Chain<double, D> ch;
        leg_chain<double, D>(ck.leg[l], g.joint_position, ch);
        V3<double> des_leg = t1_rotate_inv(ck.leg[l], V3<double>{g.desired_tip_position[0], g.desired_tip_position[1], g.desired_tip_position[2]} -
                                                        V3<double>{ck.leg[l].t1p[0], ck.leg[l].t1p[1], ck.leg[l].t1p[2]});
        g.ik_result = ik_result_value<double, D>(ck.leg[l], ch, g.joint_position, des_leg);
      }
    }
  }
}

template <class F> inline int dispatch_D_raw(int D, F&& f) {
  switch (D) {
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
  }
  return SHC_E_UNSUPPORTED;
}

inline bool engine_full(const shc_config& cfg) {
  return cfg.auto_posing || cfg.admittance_control || cfg.imu_posing || cfg.inclination_posing || cfg.use_joint_effort ||
         cfg.gravity_aligned_tips || cfg.rough_terrain_mode;
}
// Kernel instantiation of an engine: 0 walking only, 1 every optional stage, 2 those plus the extended paths: tip orientation
// (gravity_aligned_tips: tip-align posing on legs of at most three joints, tip-rotation IK beyond) and rough-terrain mode
inline int engine_mode(const shc_config& cfg) { return (cfg.gravity_aligned_tips || cfg.rough_terrain_mode) ? 2 : engine_full(cfg) ? 1 : 0; }

// Everything shc_create refuses: invalid configurations and reference features outside the built scope.
inline bool check_supported(const shc_config& cfg, std::string& err, bool& unsupported) {
  if (!validate_config(cfg, err, unsupported)) return false;
  if (cfg.auto_posing && cfg.pose_frequency != -1.0 && !(cfg.pose_frequency > 0.0 && cfg.pose_phase_length > 0)) {
    err = "auto posing with its own cycle needs pose_frequency > 0 and pose_phase_length > 0 (or pose_frequency = -1 to follow the step cycle)";
    return false;
  }
  return true;
}

// The packed integer state (shc_layout.h) keeps the phase in 16 bits and the progress numerators in signed 16 bits: a
// step cycle that does not fit (step_frequency below ~0.003 Hz at 50 Hz control rate; the reference allows 0.001) is
// refused instead of silently aliasing phases.
inline bool check_step_cycle(const shc_startup& su, std::string& err) {
  if (su.period < 1 || su.period > 65535 || su.swing_period < 1 || su.swing_period > 32767 || su.stance_period < 0 ||
      su.stance_period > 32767) {
    err = "step cycle does not fit the packed phase / progress words (period <= 65535, swing and stance period <= 32767): "
          "raise step_frequency or time_delta";
    return false;
  }
  return true;
}

// Constants block + start-up results of an engine (host arithmetic only).  `startup` = null: the engine's own restatement
// of the reference's start-up path (shc_host.cuh).
// A poser with its own cycle sways the body while the reference computes its workspaces: the tip check of model.cpp:330
// fails, the workspace and every speed limit come out zero (tests/parity_cases.py::auto_posing_own_cycle shows it on the
// oracle).  The engine does not re-derive that accident: such a configuration needs explicit start-up constants.
inline bool own_startup_supported(const shc_config& cfg, std::string& err) {
  if (cfg.auto_posing && cfg.pose_frequency != -1.0) {
    err = "auto posing with its own pose_frequency: pass explicit shc_startup constants (the engine's own start-up models a body "
          "at rest; the reference's start-up under a free-running poser yields an empty workspace)";
    return false;
  }
  return true;
}

template <class E> int core_init(E* e, const shc_config& cfg, const shc_startup* startup, int n_robots, int precision) {
  e->cfg = cfg;
  e->precision = precision;
  e->n = n_robots;
  std::memset(&e->c, 0, sizeof(e->c));
  layout(cfg, n_robots, e->c.i);
  e->n_pad = e->c.i.n_pad;
  fill_static_consts<double>(cfg, e->c.d);
  fill_static_consts<float>(cfg, e->c.f);
  std::memset(&e->su, 0, sizeof(e->su));
  if (startup) {
    e->su = *startup;
  } else {
    compute_step_cycle(cfg, e->su);
    dispatch_D_raw(cfg.joint_count, [&](auto dtag) -> int { compute_startup<decltype(dtag)::value>(e->cfg, e->c.d, e->su); return 0; });
  }
  IntConsts tmp = e->c.i;
  fill_startup_consts<double>(cfg, e->su, e->c.d, e->c.i);
  fill_startup_consts<float>(cfg, e->su, e->c.f, tmp);
  return SHC_OK;
}

// One tile (32 robots) of a new engine: every robot in the post-start-up state.  `h` is sized for one tile.
template <class E> void initial_tile(const E* e, HostPlanes& h) {
  const IntConsts& ci = e->c.i;
  shc_robot_state init;
  dispatch_D_raw(e->cfg.joint_count, [&](auto dtag) -> int { initial_state<decltype(dtag)::value>(e->cfg, e->c.d, e->su, init); return 0; });
  h.s.assign((size_t)ci.nS * 32, 0.0);
  h.d.assign((size_t)ci.nD * 32, 0.0);
  h.i.assign((size_t)ci.nI * 32, 0);
  std::vector<shc_robot_state> tile_states(32, init);
  pack(e, tile_states.data(), 32, h, 32);
}

// Planes of a new engine: the initial tile, replicated over the batch (host side; the engine replicates on the device).
template <class E> void initial_planes(const E* e, HostPlanes& h) {
  const IntConsts& ci = e->c.i;
  const size_t n_tiles = ci.n_pad / 32;
  HostPlanes t;
  initial_tile(e, t);
  h.s.resize(t.s.size() * n_tiles);
  h.d.resize(t.d.size() * n_tiles);
  h.i.resize(t.i.size() * n_tiles);
  for (size_t k = 0; k < n_tiles; ++k) {
    std::copy(t.s.begin(), t.s.end(), h.s.begin() + k * t.s.size());
    std::copy(t.d.begin(), t.d.end(), h.d.begin() + k * t.d.size());
    std::copy(t.i.begin(), t.i.end(), h.i.begin() + k * t.i.size());
  }
}

}  // namespace shc
