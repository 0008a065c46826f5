// shc_sequence.cuh — stepping and joint-space sequences of the PoseController for one robot of the batch (SURVEY.md §8(f)
// rank 2), host/device: the kernels of shc_step_to_new_stance / shc_transition_step (csrc/shc_engine.cu) run them with one
// thread per robot / per joint, the CPU test-suite runs the same routines on the host emulator's planes.
//   PoseController::stepToNewStance     pose_controller.cpp:520   (tripod leg coordination: one leg group steps at a time)
//   LegPoser::stepToPosition            :1571                     (dual quartic Bezier tip path + body pose blend)
//   PoseController::packLegs/unpackLegs :597 / :661               (LegPoser::transitionConfiguration :1476, joint space)
// followed, as in the reference, by Leg::setDesiredTipPose (model.cpp:653) and one Leg::applyIK (model.cpp:861) per stepping
// leg per loop().  The sequences work on the engine's own state planes (joint positions / velocities) and keep their
// bookkeeping (LegPoser::first_iteration_, master_iteration_count_, origin_tip_pose_; PoseController::legs_completed_step_,
// current_group_) in a small plane-major side buffer, so every robot follows its own course through the reference's
// control flow — the early "already there" exit of stepToPosition, group switches — without any host round trip.
#pragma once
#include "shc_msgs.cuh"

namespace shc {

// Per-robot sequence state, plane-major (index [field][robot], stride n_pad): device memory for the engine, host memory
// for the emulator.
struct SeqBuffers {
  double* origin;  // [L][3][n_pad]  LegPoser::origin_tip_pose_.position_ (base_link frame)
  int* count;      // [L][n_pad]     LegPoser::master_iteration_count_; -1 while first_iteration_ is set
  int* robot;      // [2][n_pad]     PoseController::legs_completed_step_, current_group_
  size_t n_pad;
};
inline size_t seq_origin_count(int L, size_t n_pad) { return (size_t)L * 3 * n_pad; }
inline size_t seq_count_count(int L, size_t n_pad) { return (size_t)L * n_pad; }
inline size_t seq_robot_count(size_t n_pad) { return 2 * n_pad; }

struct NewStanceParams {
  double lift_height;   // swing_height.current_value (pose_controller.cpp:531)
  int num_iterations;   // max(1, roundToInt((1 / step_frequency) / time_delta)) (:532, :1612)
  int apply_delta;      // stepToPosition's apply_delta defaults to true (pose_controller.h:503)
};

// One loop() of PoseController::stepToNewStance for robot r.  joints_out: [N][L][D] or null.  Returns the reference's
// progress value (0..100; 100 is never returned by the reference formula before the wrap-around, see :540-548).
template <class S, int D>
SHC_HD int step_to_new_stance_robot(const Consts& c, Planes<S> pl, const SeqBuffers& sq, const NewStanceParams& np, int r,
                                    float* joints_out) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  const RealConsts<double>& ck = c.d;
  const int L = ci.L;
  const PlaneReader<S> rd = plane_reader(ci, pl, r);
  const PoseT<double> target_pose = current_pose_of(ci, rd);  // model_->getCurrentPose()
  const bool adm = ci.admittance_control || ci.use_joint_effort;
  int completed = sq.robot[r], group = sq.robot[sq.n_pad + r];
  int progress = 0;
  for (int l = 0; l < L; ++l) {
    if ((l & 1) != group) continue;  // Leg::group_ = id % 2 (model.cpp:187)
    const LegConsts<double>& lc = ck.leg[l];
    S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
    double q[D], qd[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      q[j] = (double)sl[(LS::Q + j) * 32];
      qd[j] = (double)sl[(LS::QD + j) * 32];
    }
    Chain<double, D> ch;
    leg_chain<double, D>(lc, q, ch);
    const V3<double> t1p{lc.t1p[0], lc.t1p[1], lc.t1p[2]};
    V3<double> adelta{0.0, 0.0, 0.0};
    if (adm) adelta = {(double)sl[(LS::ADM_DELTA) * 32], (double)sl[(LS::ADM_DELTA + 1) * 32], (double)sl[(LS::ADM_DELTA + 2) * 32]};
    // ---- LegPoser::stepToPosition(default tip pose, current pose, swing height, 1 / step frequency) ----
    int count = sq.count[(size_t)l * sq.n_pad + r];
    double* op = sq.origin + (size_t)l * 3 * sq.n_pad + r;
    V3<double> origin;
    if (count < 0) {  // first_iteration_: origin = Leg::current_tip_pose_ (the model's tip, FK of the joint positions)
      origin = t1_rotate(lc, ch.tip) + t1p;
      op[0] = origin.x; op[sq.n_pad] = origin.y; op[2 * sq.n_pad] = origin.z;
      count = 0;
    } else {
      origin = {op[0], op[sq.n_pad], op[2 * sq.n_pad]};
    }
    V3<double> desired_tip{(double)sl[(LS::DEF) * 32], (double)sl[(LS::DEF + 1) * 32], (double)sl[(LS::DEF + 2) * 32]};
    const V3<double> position_delta = origin - pose_inverse_transform(target_pose, desired_tip);
    V3<double> poser_tip;
    int leg_progress;
    if (!(norm(position_delta) > 0.01) && np.lift_height == 0.0) {  // TIP_TOLERANCE: nothing to do (:1600)
      count = -1;
      poser_tip = origin;
      leg_progress = 100;
    } else {
      if (np.apply_delta) desired_tip = desired_tip + adelta;
      ++count;
      const int num = np.num_iterations;
      const double delta_t = 1.0 / num;
      const double completion_ratio = double(count - 1) / double(num);
      const PoseT<double> desired_pose = pose_interpolate(pose_identity<double>(), smooth_step(completion_ratio), target_pose);
      const int half = num / 2;
      const V3<double> o2t = origin - desired_tip;
      const V3<double> lift{0.0, 0.0, np.lift_height};
      const V3<double> n1[5] = {origin, origin, origin + lift, desired_tip + o2t * 0.75 + lift, desired_tip + o2t * 0.5 + lift};
      const V3<double> n2[5] = {desired_tip + o2t * 0.5 + lift, desired_tip + o2t * 0.25 + lift, desired_tip + lift, desired_tip,
                                desired_tip};
      const int sic = (count + (num - 1)) % num + 1;
      const V3<double> new_tip = sic <= half ? quartic_bezier(n1, sic * delta_t * 2.0) : quartic_bezier(n2, (sic - half) * delta_t * 2.0);
      poser_tip = pose_inverse_transform(desired_pose, new_tip);
      if (count >= num) {
        count = -1;
        leg_progress = 100;
      } else {
        leg_progress = int(completion_ratio * 100);
      }
    }
    sq.count[(size_t)l * sq.n_pad + r] = count;
    // ---- leg->setDesiredTipPose(poser tip pose) (adds the admittance delta again, model.cpp:661) + leg->applyIK() ----
    V3<double> des_leg;
    apply_ik_step<double, D>(ck, lc, ch, q, qd, poser_tip + adelta, ci.clamp_joint_positions != 0, ci.clamp_joint_velocities != 0, &des_leg);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      sl[(LS::Q + j) * 32] = S(q[j]);
      sl[(LS::QD + j) * 32] = S(qd[j]);
    }
    progress = leg_progress;
    completed += leg_progress == 100 ? 1 : 0;
  }
  if (joints_out) {
    for (int l = 0; l < L; ++l) {
      const S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
#pragma unroll
      for (int j = 0; j < D; ++j) joints_out[((size_t)r * L + l) * D + j] = (float)((double)sl[(LS::Q + j) * 32] + ck.leg[l].joffset[j]);
    }
  }
  progress = progress / 2 + group * 50;
  group = completed / (L / 2);
  if (completed == L) {
    completed = 0;
    group = 0;
  }
  sq.robot[r] = completed;
  sq.robot[sq.n_pad + r] = group;
  return progress;
}

// LegPoser::transitionConfiguration (:1476) for joint j of leg l of robot r: iteration `it` of `num` from the configuration
// latched when the transition began (origin [N][L][D]) to `desired` ([kMaxLegs][kMaxDof]).  The joint position goes into
// the state planes every iteration (the reference sets Joint::desired_position_ and leaves desired_velocity_ alone).
template <class S, int D>
SHC_HD void transition_joint(const Consts& c, Planes<S> pl, const double* origin, const double* desired, int it, int num, long long i,
                             float* joints_out) {
  const int L = c.i.L;
  const int j = (int)(i % D), l = (int)((i / D) % L), r = (int)(i / ((long long)D * L));
  const double q = transition_configuration(origin[i], desired[l * kMaxDof + j], it, num);
  if (joints_out) joints_out[i] = (float)(q + c.d.leg[l].joffset[j]);
  S* sl = pl.s + ((size_t)(r >> 5) * c.i.nS + c.i.offS_leg + l * c.i.strideS_leg) * 32 + (r & 31);
  sl[(LegS<D>::Q + j) * 32] = S(q);
}
// The joint positions the state planes hold, as the origin of a transition ([N][L][D] doubles).
template <class S, int D> SHC_HD void latch_joint(const Consts& c, Planes<S> pl, long long i, double* origin) {
  const int L = c.i.L;
  const int j = (int)(i % D), l = (int)((i / D) % L), r = (int)(i / ((long long)D * L));
  const S* sl = pl.s + ((size_t)(r >> 5) * c.i.nS + c.i.offS_leg + l * c.i.strideS_leg) * 32 + (r & 31);
  origin[i] = (double)sl[(LegS<D>::Q + j) * 32];
}

}  // namespace shc
