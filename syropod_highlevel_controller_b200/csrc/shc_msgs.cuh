// shc_msgs.cuh — packs the reference's output wire formats (include/shc_msgs.h) for one (robot, leg) from the state planes.
// Host/device: the kernel of shc_pack_messages (csrc/shc_engine.cu) runs it with one thread per leg; the CPU test-suite
// runs the same routine on the host emulator's planes against the oracle's restatement of the publishers.
//   publishDesiredJointState  state_controller.cpp:777-805      publishLegState          :809-893
//   publishVelocity / Pose / RotationPoseError  :897-961        publishFrameTransforms   :963-1047
#pragma once
#include "../../include/shc_msgs.h"
#include "shc_cycle.cuh"
#include "shc_startup.cuh"

namespace shc {

template <class R> SHC_HD void put_pose(double* o, V3<R> p, Q4<R> q) {
  o[0] = (double)p.x; o[1] = (double)p.y; o[2] = (double)p.z;
  o[3] = (double)q.w; o[4] = (double)q.x; o[5] = (double)q.y; o[6] = (double)q.z;
}

// Frames of one leg in the base_link frame: origin and orientation of every joint frame (before the joint's own rotation,
// i.e. Joint::getPoseRobotFrame, model.h:604) and of the tip frame (Tip::getPoseRobotFrame, :684).  Orientations are the
// normalised quaternions of the chain's rotation matrices, as Pose::transform builds them (pose.h:135-146).
template <int D> struct LegFrames {
  V3<double> joint_p[D], tip_p;
  Q4<double> joint_q[D], tip_q;
};
template <int D> SHC_HD void leg_frames(const LegConsts<double>& lc, const double* q, LegFrames<D>& f) {
  V3<double> ax{1, 0, 0}, ay{0, 1, 0}, az{0, 0, 1}, ap{0, 0, 0};
  auto emit = [&](V3<double>& p_out, Q4<double>& q_out) {
    // robot frame: R = T1r * [ax ay az] (columns), p = T1r * ap + t1p
    const V3<double> cx = t1_rotate(lc, ax), cy = t1_rotate(lc, ay), cz = t1_rotate(lc, az);
    const double m[3][3] = {{cx.x, cy.x, cz.x}, {cx.y, cy.y, cz.y}, {cx.z, cy.z, cz.z}};
    q_out = qnormalized(matrix_to_quat<double>(m));
    p_out = t1_rotate(lc, ap) + V3<double>{lc.t1p[0], lc.t1p[1], lc.t1p[2]};
  };
#pragma unroll
  for (int j = 0; j < D; ++j) {
    emit(f.joint_p[j], f.joint_q[j]);
    double s, c;
    sincos_(lc.dh_theta[j] + q[j], &s, &c);
    const double ca = lc.dh_ca[j], sa = lc.dh_sa[j];
    const V3<double> nx = ax * c + ay * s, u = ay * c - ax * s;
    const V3<double> ny = u * ca + az * sa, nz = az * ca - u * sa;
    ap = ap + nx * lc.dh_r[j] + az * lc.dh_d[j];
    ax = nx; ay = ny; az = nz;
  }
  emit(f.tip_p, f.tip_q);
}

// Model::current_pose_ from the stored sub-poses (pose_controller.cpp:811-859).
template <class S> struct PlaneReader {
  const S* s;
  const double* d;
  const int* i;
  SHC_HD double S_(int plane) const { return (double)s[plane * 32]; }
  SHC_HD double D_(int plane) const { return d[plane * 32]; }
  SHC_HD int I_(int plane) const { return i[plane * 32]; }
  SHC_HD PoseT<double> pose(int plane) const {
    return {{S_(plane), S_(plane + 1), S_(plane + 2)}, {S_(plane + 3), S_(plane + 4), S_(plane + 5), S_(plane + 6)}};
  }
};
template <class S> SHC_HD PlaneReader<S> plane_reader(const IntConsts& ci, Planes<S> pl, int r) {
  const size_t tile = (size_t)(r >> 5);
  const int lane = r & 31;
  return {pl.s + tile * (size_t)(ci.nS * 32) + lane, pl.d + tile * (size_t)(ci.nD * 32) + lane, pl.i + tile * (size_t)(ci.nI * 32) + lane};
}
template <class S> SHC_HD PoseT<double> current_pose_of(const IntConsts& ci, const PlaneReader<S>& rd) {
  PoseT<double> p = pose_add(pose_identity<double>(), rd.pose(RS_WPP));
  if (ci.manual_posing) p = pose_add(p, rd.pose(RS_MAN));
  if (ci.inclination_posing) p = pose_add(p, PoseT<double>{{rd.S_(ci.offS_imu + IMU_INCL), rd.S_(ci.offS_imu + IMU_INCL + 1), 0.0}, qidentity<double>()});
  if (ci.imu_posing) {
    const int b = ci.offS_imu + IMU_Q;
    p = pose_add(p, PoseT<double>{{0.0, 0.0, 0.0}, {rd.S_(b), rd.S_(b + 1), rd.S_(b + 2), rd.S_(b + 3)}});
  } else if (ci.auto_posing) {
    p = pose_add(p, rd.pose(ci.offS_auto + AUTO_POSE));
  }
  if (ci.tip_mode == TIP_ALIGN_POSE) p = pose_add(p, rd.pose(ci.offS_tip + TA_POSE));
  return p;
}

// One leg of one robot: its LegState record and its slice of the JointState record.  `measured` = the leg's measured joint
// positions [D] (jointStatesCallback) or null.
template <class S, int D>
SHC_HD void pack_leg_message(const Consts& c, Planes<S> pl, int r, int l, const float* measured, shc_leg_state_msg& m, shc_joint_state_msg* js) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  const RealConsts<double>& ck = c.d;
  const LegConsts<double>& lc = ck.leg[l];
  const PlaneReader<S> rd = plane_reader(ci, pl, r);
  const int sb = ci.offS_leg + l * ci.strideS_leg, db = ci.offD_leg + l * ci.strideD_leg, ib = ci.offI_leg + l * ci.strideI_leg;
  const bool adm = ci.admittance_control || ci.use_joint_effort;
  const Q4<double> undefined{0.0, 0.0, 0.0, 0.0};  // UNDEFINED_ROTATION: the engine carries no tip rotations
  const V3<double> tip{rd.D_(db + LD_TIP), rd.D_(db + LD_TIP + 1), rd.D_(db + LD_TIP + 2)};
  put_pose(m.walker_tip_pose, tip, undefined);
  put_pose(m.target_tip_pose, V3<double>{rd.S_(sb + LS::TGT), rd.S_(sb + LS::TGT + 1), rd.S_(sb + LS::TGT + 2)}, undefined);
  const PoseT<double> cur = current_pose_of(ci, rd);
  put_pose(m.poser_tip_pose, pose_inverse_transform(cur, tip), undefined);  // updateStance (pose_controller.cpp:110)
  double q[D], qd[D], qm[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    q[j] = rd.S_(sb + LS::Q + j);
    qd[j] = rd.S_(sb + LS::QD + j);
    qm[j] = measured ? (double)measured[j] : q[j];
  }
  LegFrames<D> f;
  leg_frames<D>(lc, q, f);
  put_pose(m.model_tip_pose, f.tip_p, f.tip_q);
  put_pose(m.tip_transform, f.tip_p, f.tip_q);
#pragma unroll
  for (int j = 0; j < SHC_MAX_DOF; ++j) {
    if (j < D) {
      double s, co;
      sincos_(0.5 * q[j], &s, &co);
      put_pose(m.joint_transform[j], f.joint_p[j], qmul(f.joint_q[j], Q4<double>{co, 0.0, 0.0, s}));  // * AngleAxis(q, UnitZ)
    } else {
      for (int k = 0; k < 7; ++k) m.joint_transform[j][k] = 0.0;
    }
  }
  {
    LegFrames<D> fa;
    leg_frames<D>(lc, qm, fa);
    put_pose(m.actual_tip_pose, fa.tip_p, fa.tip_q);
    // publishLegState re-runs applyFK() on the unchanged desired joint positions (state_controller.cpp:842) before it reads
    // Leg::current_tip_velocity_ (:846-848), which that call has just set to (tip - tip) / time_delta: the reference's wire
    // value is zero, and so is this one (checked against the reference's own publishers, tests/test_reference_pin.py).
    m.model_tip_velocity[0] = 0.0; m.model_tip_velocity[1] = 0.0; m.model_tip_velocity[2] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < SHC_MAX_DOF; ++j) {
    m.joint_positions[j] = j < D ? q[j] : 0.0;
    m.joint_velocities[j] = j < D ? qd[j] : 0.0;
    m.joint_efforts[j] = 0.0;
    if (js && j < D) {
      js->position[l * D + j] = q[j];
      js->velocity[l * D + j] = qd[j];
      js->effort[l * D + j] = 0.0;
      js->position_command[l * D + j] = q[j] + lc.joffset[j];
    }
  }
  const int prog = rd.I_(ib + LI_PROG);
  const int sn = (int)(short)(prog & 0xffff), tn = (int)(short)((prog >> 16) & 0xffff);
  m.swing_progress = sn < 0 ? -1.0 : (double)sn / (double)ci.swing_period;
  m.stance_progress = tn < 0 ? -1.0 : (double)tn / (double)ci.stance_period;
  const double frequency = 1.0 / (ci.period * ck.dt);  // StepCycle::frequency_ (walk_controller.cpp:374)
  const double swing_time = (double(ci.swing_period) / ci.period) / frequency;
  const double stance_time = (double(ci.stance_period) / ci.period) / frequency;
  m.time_to_swing_end = m.stance_progress >= 0.0 ? stance_time * (1.0 - m.stance_progress) + swing_time : swing_time * (1.0 - m.swing_progress);
  {
    const double T = m.time_to_swing_end;  // calculateOdometry (walk_controller.cpp:783)
    double s, co;
    sincos_(0.5 * (rd.S_(RS_ANGVEL) * T), &s, &co);
    put_pose(m.pose_delta, V3<double>{rd.S_(RS_VEL) * T, rd.S_(RS_VEL + 1) * T, 0.0 * T}, Q4<double>{co, 0.0, 0.0, s});
  }
  {
    const PoseT<double> ap = ci.auto_posing ? rd.pose(ci.offS_auto + AUTO_POSE) : pose_identity<double>();
    put_pose(m.auto_pose, ap.p, ap.q);
  }
  for (int k = 0; k < 3; ++k) {
    m.tip_force[k] = adm ? rd.S_(sb + LS::ADM_FORCE + k) * ck.force_gain : 0.0;
    m.admittance_delta[k] = adm ? rd.S_(sb + LS::ADM_DELTA + k) : 0.0;
  }
  m.virtual_stiffness = adm ? rd.S_(sb + LS::STIFF) : 0.0;
}

template <class S> SHC_HD void pack_body_message(const Consts& c, Planes<S> pl, int r, shc_body_msg& m) {
  const IntConsts& ci = c.i;
  const PlaneReader<S> rd = plane_reader(ci, pl, r);
  m.velocity[0] = rd.S_(RS_VEL); m.velocity[1] = rd.S_(RS_VEL + 1); m.velocity[2] = 0.0;
  m.velocity[3] = 0.0; m.velocity[4] = 0.0; m.velocity[5] = rd.S_(RS_ANGVEL);
  const PoseT<double> cur = current_pose_of(ci, rd);
  const V3<double> e = quat_to_euler_nc(cur.q, false);
  m.pose[0] = cur.p.x; m.pose[1] = cur.p.y; m.pose[2] = cur.p.z; m.pose[3] = e.x; m.pose[4] = e.y; m.pose[5] = e.z;
  const bool imu = ci.imu_posing || ci.inclination_posing;
  for (int k = 0; k < 3; ++k) {
    m.rotation_pose_error[k] = imu ? rd.S_(ci.offS_imu + IMU_ABS + k) : 0.0;
    m.rotation_pose_error[3 + k] = imu ? rd.S_(ci.offS_imu + IMU_POS + k) : 0.0;
    m.rotation_pose_error[6 + k] = imu ? rd.S_(ci.offS_imu + IMU_VEL + k) : 0.0;
  }
  const PoseT<double> odom{{rd.D_(RD_ODOMP), rd.D_(RD_ODOMP + 1), rd.D_(RD_ODOMP + 2)},
                           {rd.S_(RS_ODOMQ), rd.S_(RS_ODOMQ + 1), rd.S_(RS_ODOMQ + 2), rd.S_(RS_ODOMQ + 3)}};
  const PoseT<double> o2b = pose_add(odom, cur);
  put_pose(m.odom_ideal_to_base_link, o2b.p, o2b.q);
  const PoseT<double> inv = pose_inverse(cur);
  put_pose(m.base_link_to_walk_plane, inv.p, inv.q);
}

}  // namespace shc
