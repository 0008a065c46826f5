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
constexpr int kMaxTransitionPoses = 24;  // TRANSITION_STEP_THRESHOLD (20) + the initial pose + slack (pose_controller.h:24)
struct SeqBuffers {
  double* origin;  // [L][3][n_pad]  LegPoser::origin_tip_pose_.position_ (base_link frame)
  int* count;      // [L][n_pad]     LegPoser::master_iteration_count_; -1 while first_iteration_ is set
  int* robot;      // [5][n_pad]     PoseController::legs_completed_step_, current_group_, transition_step_,
                   //                transition_step_count_, flags (SEQ_* bits)
  size_t n_pad;
  // executeSequence only (allocated on its first use)
  double* target;  // [L][3][n_pad]  LegPoser::target_tip_pose_.position_
  double* poses;   // [kMaxTransitionPoses][L][3][n_pad]  LegPoser::transition_poses_ (positions)
  int* leg;        // [2][L][n_pad]  LegPoser::leg_completed_step_, number of transition poses
};
enum : int { SEQ_SET_TARGET = 1, SEQ_PROXIMITY_ALERT = 2, SEQ_H_COMPLETE = 4, SEQ_V_COMPLETE = 8, SEQ_FIRST_EXECUTION = 16,
             SEQ_RESET_SEQUENCE = 32, SEQ_FAILED = 64,
             // batch bookkeeping, not reference state: the robot's last completed sequence.  The reference's state machine stops
             // calling executeSequence once it has returned 100 (state_controller.cpp:314-350); in a batch the robots complete
             // at different loops, so a robot that is through its start-up (shut-down) ignores further start-up (shut-down) loops
             SEQ_DONE_START_UP = 128, SEQ_DONE_SHUT_DOWN = 256 };
constexpr int kSeqInitialFlags = SEQ_SET_TARGET | SEQ_FIRST_EXECUTION | SEQ_RESET_SEQUENCE;  // pose_controller.h:299-304
inline size_t seq_origin_count(int L, size_t n_pad) { return (size_t)L * 3 * n_pad; }
inline size_t seq_count_count(int L, size_t n_pad) { return (size_t)L * n_pad; }
inline size_t seq_robot_count(size_t n_pad) { return 5 * n_pad; }
inline size_t seq_poses_count(int L, size_t n_pad) { return (size_t)kMaxTransitionPoses * L * 3 * n_pad; }
inline size_t seq_leg_count(int L, size_t n_pad) { return (size_t)2 * L * n_pad; }

struct NewStanceParams {
  double lift_height;   // swing_height.current_value (pose_controller.cpp:531)
  int num_iterations;   // max(1, roundToInt((1 / step_frequency) / time_delta)) (:532, :1612)
  int apply_delta;      // stepToPosition's apply_delta defaults to true (pose_controller.h:503)
};

// LegPoser::stepToPosition (pose_controller.cpp:1571) for the tip POSITION (the target rotations of these sequences are the
// stepper's, undefined without gravity_aligned_tips): one iteration towards desired_tip under the body pose target_pose.
// count / origin are the leg's LegPoser::master_iteration_count_ (-1 = first_iteration_) and origin_tip_pose_; model_tip =
// Leg::current_tip_pose_.position_.  Returns the progress (0..100) and the poser's new tip position.
SHC_HD int step_to_position(V3<double> desired_tip, const PoseT<double>& target_pose, double lift_height, int num, V3<double> delta,
                            bool apply_delta, V3<double> model_tip, int& count, V3<double>& origin, V3<double>* poser_tip,
                            bool* already_there) {
  *already_there = false;
  if (count < 0) {
    origin = model_tip;
    count = 0;
  }
  const V3<double> position_delta = origin - pose_inverse_transform(target_pose, desired_tip);
  if (!(norm(position_delta) > 0.01) && lift_height == 0.0) {  // TIP_TOLERANCE: nothing to do (:1600)
    count = -1;
    *poser_tip = origin;
    *already_there = true;  // current_tip_pose_ = origin_tip_pose_: the poser's pose now carries the MODEL's tip rotation
    return 100;
  }
  if (apply_delta) desired_tip = desired_tip + delta;
  ++count;
  const double delta_t = 1.0 / num;
  const double completion_ratio = double(count - 1) / double(num);
  const PoseT<double> desired_pose = pose_interpolate(pose_identity<double>(), smooth_step(completion_ratio), target_pose);
  const int half = num / 2;
  const V3<double> o2t = origin - desired_tip;
  const V3<double> lift{0.0, 0.0, lift_height};
  const V3<double> n1[5] = {origin, origin, origin + lift, desired_tip + o2t * 0.75 + lift, desired_tip + o2t * 0.5 + lift};
  const V3<double> n2[5] = {desired_tip + o2t * 0.5 + lift, desired_tip + o2t * 0.25 + lift, desired_tip + lift, desired_tip, desired_tip};
  const int sic = (count + (num - 1)) % num + 1;
  const V3<double> new_tip = sic <= half ? quartic_bezier(n1, sic * delta_t * 2.0) : quartic_bezier(n2, (sic - half) * delta_t * 2.0);
  *poser_tip = pose_inverse_transform(desired_pose, new_tip);
  if (count >= num) {
    count = -1;
    return 100;
  }
  return int(completion_ratio * 100);
}

// One leg of a stepping sequence: LegPoser::stepToPosition, Leg::setDesiredTipPose (adds the admittance delta when
// desired_delta is set, model.cpp:661) and Leg::applyIK on the joint state in the planes.  Returns stepToPosition's progress;
// *limit_proximity = applyIK's return value, *poser_tip = LegPoser::current_tip_pose_.position_.
template <class S, int D>
SHC_HD int sequence_step_leg(const Consts& c, Planes<S> pl, const SeqBuffers& sq, int r, int l, V3<double> desired_tip,
                             const PoseT<double>& target_pose, double lift_height, int num, bool apply_delta, bool desired_delta,
                             V3<double>* poser_tip, double* limit_proximity) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  const RealConsts<double>& ck = c.d;
  const LegConsts<double>& lc = ck.leg[l];
  const bool adm = ci.admittance_control || ci.use_joint_effort;
  S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    q[j] = (double)sl[(LS::Q + j) * 32];
    qd[j] = (double)sl[(LS::QD + j) * 32];
  }
  Chain<double, D> ch;
  leg_chain<double, D>(lc, q, ch);
  const V3<double> model_tip = t1_rotate(lc, ch.tip) + V3<double>{lc.t1p[0], lc.t1p[1], lc.t1p[2]};
  V3<double> adelta{0.0, 0.0, 0.0};
  if (adm) adelta = {(double)sl[(LS::ADM_DELTA) * 32], (double)sl[(LS::ADM_DELTA + 1) * 32], (double)sl[(LS::ADM_DELTA + 2) * 32]};
  int count = sq.count[(size_t)l * sq.n_pad + r];
  double* op = sq.origin + (size_t)l * 3 * sq.n_pad + r;
  V3<double> origin{op[0], op[sq.n_pad], op[2 * sq.n_pad]};
  bool already_there;
  const int progress = step_to_position(desired_tip, target_pose, lift_height, num, adelta, apply_delta, model_tip, count, origin, poser_tip,
                                        &already_there);
  sq.count[(size_t)l * sq.n_pad + r] = count;
  op[0] = origin.x; op[sq.n_pad] = origin.y; op[2 * sq.n_pad] = origin.z;
  V3<double> des_leg;
  const V3<double> desired = desired_delta ? *poser_tip + adelta : *poser_tip;
  if (already_there) {
    // stepToPosition's "already there" exit hands back origin_tip_pose_ = Leg::current_tip_pose_ whole (:1602), so the desired
    // pose of this applyIK has a DEFINED rotation — the tip's own: the reference runs its rotation-constrained branch
    // (model.cpp:880-900: the position step without the velocity clamp, then a rotation step for a zero rotation)
    const V3<double> cx = t1_rotate(lc, ch.tipx), cy = t1_rotate(lc, ch.tipy), cz = t1_rotate(lc, ch.tipz);
    const double m[3][3] = {{cx.x, cy.x, cz.x}, {cx.y, cy.y, cz.y}, {cx.z, cy.z, cz.z}};
    bool ok;
    int force_updates;
    apply_ik_pose<double, D>(ck, lc, ch, q, qd, desired, qnormalized(matrix_to_quat(m)), ci.clamp_joint_positions != 0,
                             ci.clamp_joint_velocities != 0, &des_leg, &ok, &force_updates);
  } else {
    apply_ik_step<double, D>(ck, lc, ch, q, qd, desired, ci.clamp_joint_positions != 0, ci.clamp_joint_velocities != 0, &des_leg);
    if (limit_proximity) leg_chain<double, D>(lc, q, ch);
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    sl[(LS::Q + j) * 32] = S(q[j]);
    sl[(LS::QD + j) * 32] = S(qd[j]);
  }
  if (limit_proximity) *limit_proximity = ik_result_value<double, D>(lc, ch, q, des_leg);  // `ch` = the chain at the new joints
  return progress;
}

template <class S, int D> SHC_HD void sequence_write_joints(const Consts& c, Planes<S> pl, int r, float* joints_out) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  for (int l = 0; l < ci.L; ++l) {
    const S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
#pragma unroll
    for (int j = 0; j < D; ++j) joints_out[((size_t)r * ci.L + l) * D + j] = (float)((double)sl[(LS::Q + j) * 32] + c.d.leg[l].joffset[j]);
  }
}

// One loop() of PoseController::stepToNewStance for robot r.  joints_out: [N][L][D] or null.  Returns the reference's
// progress value (0..100; 100 is never returned by the reference formula before the wrap-around, see :540-548).
template <class S, int D>
SHC_HD int step_to_new_stance_robot(const Consts& c, Planes<S> pl, const SeqBuffers& sq, const NewStanceParams& np, int r,
                                    float* joints_out) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  const int L = ci.L;
  const PlaneReader<S> rd = plane_reader(ci, pl, r);
  const PoseT<double> target_pose = current_pose_of(ci, rd);  // model_->getCurrentPose()
  int completed = sq.robot[r], group = sq.robot[sq.n_pad + r];
  int progress = 0;
  for (int l = 0; l < L; ++l) {
    if ((l & 1) != group) continue;  // Leg::group_ = id % 2 (model.cpp:187)
    const S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
    const V3<double> default_tip{(double)sl[(LS::DEF) * 32], (double)sl[(LS::DEF + 1) * 32], (double)sl[(LS::DEF + 2) * 32]};
    V3<double> poser_tip;
    progress = sequence_step_leg<S, D>(c, pl, sq, r, l, default_tip, target_pose, np.lift_height, np.num_iterations, np.apply_delta != 0, true,
                                       &poser_tip, nullptr);
    completed += progress == 100 ? 1 : 0;
  }
  if (joints_out) sequence_write_joints<S, D>(c, pl, r, joints_out);
  progress = progress / 2 + group * 50;
  group = completed / (L / 2);
  if (completed == L) {
    completed = 0;
    group = 0;
  }
  sq.robot[r] = completed;
  sq.robot[sq.n_pad + r] = group;
  sq.robot[4 * sq.n_pad + r] |= SEQ_RESET_SEQUENCE;  // reset_transition_sequence_ = true (:554)
  return progress;
}

// One loop() of PoseController::executeSequence (pose_controller.cpp:145-461) for robot r: the start-up (shut_down = false)
// or shut-down sequence — alternating horizontal transitions (the two leg groups step in turn, or all legs at once while
// the body does not bear load) and vertical transitions (all legs, the body rises / sinks), towards the transition poses
// the first start-up recorded (its own course: safety factor on the joint-limit proximity, early stops on a proximity
// alert).  Returns the reference's value: -1 while the first execution generates the sequence, else 0..100; -2 once the
// sequence has needed more than TRANSITION_STEP_THRESHOLD steps (where the reference shuts the controller down).
struct ExecuteSequenceParams {
  double lift_height;  // swing_height.current_value
  double step_frequency, time_delta;
};
template <class S, int D>
SHC_HD int execute_sequence_robot(const Consts& c, Planes<S> pl, const SeqBuffers& sq, const ExecuteSequenceParams& ep, bool shut_down, int r,
                                  float* joints_out) {
  using LS = LegS<D>;
  const IntConsts& ci = c.i;
  const RealConsts<double>& ck = c.d;
  const int L = ci.L;
  const size_t np = sq.n_pad;
  int completed = sq.robot[r], group = sq.robot[np + r], tstep = sq.robot[2 * np + r], tcount = sq.robot[3 * np + r];
  int flags = sq.robot[4 * np + r];
  if (flags & (shut_down ? SEQ_DONE_SHUT_DOWN : SEQ_DONE_START_UP)) {
    if (joints_out) sequence_write_joints<S, D>(c, pl, r, joints_out);
    return 100;
  }
  flags &= ~(SEQ_DONE_START_UP | SEQ_DONE_SHUT_DOWN);
  auto pose_ptr = [&](int k, int l) { return sq.poses + ((size_t)(k * L + l) * 3) * np + r; };
  auto n_poses = [&](int l) -> int& { return sq.leg[(size_t)(L + l) * np + r]; };
  auto leg_done = [&](int l) -> int& { return sq.leg[(size_t)l * np + r]; };
  auto add_pose = [&](int l, V3<double> p) {
    int& n = n_poses(l);
    if (n < kMaxTransitionPoses) {
      double* d = pose_ptr(n, l);
      d[0] = p.x; d[np] = p.y; d[2 * np] = p.z;
      ++n;
    }
  };
  // Leg::current_tip_pose_.position_ of every leg as this loop() finds it
  V3<double> model_tip[kMaxLegs];
  double height_sum = 0.0;
  for (int l = 0; l < L; ++l) {
    const S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
    double q[D];
#pragma unroll
    for (int j = 0; j < D; ++j) q[j] = (double)sl[(LS::Q + j) * 32];
    Chain<double, D> ch;
    leg_chain<double, D>(ck.leg[l], q, ch);
    model_tip[l] = t1_rotate(ck.leg[l], ch.tip) + V3<double>{ck.leg[l].t1p[0], ck.leg[l].t1p[1], ck.leg[l].t1p[2]};
    height_sum += model_tip[l].z;
  }
  if ((flags & SEQ_RESET_SEQUENCE) && !shut_down) {  // (:149-161)
    flags = (flags & ~SEQ_RESET_SEQUENCE) | SEQ_FIRST_EXECUTION;
    tstep = 0;
    for (int l = 0; l < L; ++l) {
      n_poses(l) = 0;
      add_pose(l, model_tip[l]);
    }
  }
  const bool first = (flags & SEQ_FIRST_EXECUTION) != 0;
  int progress = 0, normalised_progress = 0, next_step, total_progress, step_target;
  bool horizontal, vertical;
  if (!shut_down) {
    horizontal = !(tstep % 2);
    vertical = tstep % 2;
    next_step = tstep + 1;
    step_target = tcount;
    total_progress = tstep * 100 / max_(tcount, 1);
  } else {
    horizontal = tstep % 2;
    vertical = !(tstep % 2);
    next_step = tstep - 1;
    step_target = 0;
    total_progress = 100 - tstep * 100 / max_(tcount, 1);
  }
  const bool final_transition = first ? ((flags & (SEQ_H_COMPLETE | SEQ_V_COMPLETE)) != 0) : (next_step == step_target);
  bool sequence_complete = false;
  const double safety_factor = first ? 0.15 / (tstep + 1) : 0.0;  // SAFETY_FACTOR (pose_controller.h:20)
  const PlaneReader<S> rd = plane_reader(ci, pl, r);
  const PoseT<double> current_pose = current_pose_of(ci, rd);
  const PoseT<double> identity = pose_identity<double>();
  auto target_of = [&](int l) {  // the transition pose of the next step, else the default stance tip under the body pose
    if (next_step >= 0 && n_poses(l) > next_step) {
      const double* d = pose_ptr(next_step, l);
      return V3<double>{d[0], d[np], d[2 * np]};
    }
    const S* sl = pl.s + ((size_t)(r >> 5) * ci.nS + ci.offS_leg + l * ci.strideS_leg) * 32 + (r & 31);
    return pose_inverse_transform(current_pose, V3<double>{(double)sl[(LS::DEF) * 32], (double)sl[(LS::DEF + 1) * 32], (double)sl[(LS::DEF + 2) * 32]});
  };
  auto target_ptr = [&](int l) { return sq.target + (size_t)l * 3 * np + r; };
  const bool apply_delta = !shut_down && final_transition;
  const double time_scale = first ? 2.0 : 1.0;

  if (horizontal) {
    if (flags & SEQ_SET_TARGET) {
      flags &= ~SEQ_SET_TARGET;
      for (int l = 0; l < L; ++l) {
        leg_done(l) = 0;
        V3<double> t = target_of(l);
        t.z = model_tip[l].z;  // maintain the height
        double* d = target_ptr(l);
        d[0] = t.x; d[np] = t.y; d[2 * np] = t.z;
      }
    }
    const bool direct_step = !(-(height_sum / L) > 0.05);  // !Model::legsBearingLoad() (model.cpp:78, HALF_BODY_DEPTH)
    const int num = max_(1, round_to_int(((1.0 / ep.step_frequency) * time_scale) / ep.time_delta));  // HORIZONTAL_TRANSITION_TIME = 1
    for (int l = 0; l < L; ++l) {
      if (leg_done(l)) continue;
      if ((l & 1) == group || direct_step) {
        double* d = target_ptr(l);
        V3<double> poser_tip;
        double limit_proximity;
        progress = sequence_step_leg<S, D>(c, pl, sq, r, l, V3<double>{d[0], d[np], d[2 * np]}, identity, direct_step ? 0.0 : ep.lift_height, num,
                                           apply_delta, true, &poser_tip, &limit_proximity);
        const bool exceeded_workspace = limit_proximity < safety_factor;
        if (first && exceeded_workspace) {  // stop this leg's transition early
          d[0] = poser_tip.x; d[np] = poser_tip.y; d[2 * np] = poser_tip.z;
          sq.count[(size_t)l * np + r] = -1;  // resetStepToPosition
          progress = 100;
          flags |= SEQ_PROXIMITY_ALERT;
        }
        if (progress == 100) {
          leg_done(l) = 1;
          ++completed;
          if (first) add_pose(l, exceeded_workspace ? poser_tip : V3<double>{d[0], d[np], d[2 * np]});
        }
      } else {
        ++completed;
        leg_done(l) = 1;
      }
    }
    if (direct_step) normalised_progress = progress / max_(tcount, 1);
    else normalised_progress = (progress / 2 + (group == 0 ? 0 : 50)) / max_(tcount, 1);
    if (completed == L) {
      flags |= SEQ_SET_TARGET;
      completed = 0;
      if (group == 1 || direct_step) {
        group = 0;
        tstep = next_step;
        flags = (flags & ~SEQ_H_COMPLETE) | ((flags & SEQ_PROXIMITY_ALERT) ? 0 : SEQ_H_COMPLETE);
        sequence_complete = final_transition;
        flags &= ~SEQ_PROXIMITY_ALERT;
      } else if (group == 0) {
        group = 1;
      }
    }
  }

  if (vertical) {
    if (flags & SEQ_SET_TARGET) {
      flags &= ~SEQ_SET_TARGET;
      for (int l = 0; l < L; ++l) {
        V3<double> t = target_of(l);
        t.x = model_tip[l].x;  // maintain the horizontal position
        t.y = model_tip[l].y;
        double* d = target_ptr(l);
        d[0] = t.x; d[np] = t.y; d[2 * np] = t.z;
      }
    }
    const int num = max_(1, round_to_int(((3.0 / ep.step_frequency) * time_scale) / ep.time_delta));  // VERTICAL_TRANSITION_TIME = 3
    bool all_within = true;
    V3<double> poser_tips[kMaxLegs];
    for (int l = 0; l < L; ++l) {
      double* d = target_ptr(l);
      double limit_proximity;
      progress = sequence_step_leg<S, D>(c, pl, sq, r, l, V3<double>{d[0], d[np], d[2 * np]}, identity, 0.0, num, apply_delta, false, &poser_tips[l],
                                         &limit_proximity);
      all_within = all_within && !(limit_proximity < safety_factor);
    }
    if ((!all_within && first) || progress == 100) {
      for (int l = 0; l < L; ++l) {
        sq.count[(size_t)l * np + r] = -1;  // resetStepToPosition
        progress = 100;
        if (first) {
          const double* d = target_ptr(l);
          add_pose(l, all_within ? V3<double>{d[0], d[np], d[2 * np]} : poser_tips[l]);
        }
      }
      flags = (flags & ~SEQ_V_COMPLETE) | (all_within ? SEQ_V_COMPLETE : 0);
      tstep = next_step;
      sequence_complete = final_transition;
      flags |= SEQ_SET_TARGET;
    }
    normalised_progress = progress / max_(tcount, 1);
  }

  if (first) tcount = tstep;
  if (tstep > 20) flags |= SEQ_FAILED;  // TRANSITION_STEP_THRESHOLD: the reference logs FATAL and shuts down (:438-442)
  int result;
  if (sequence_complete) {
    flags = (flags | SEQ_SET_TARGET) & ~(SEQ_V_COMPLETE | SEQ_H_COMPLETE | SEQ_FIRST_EXECUTION);
    flags |= shut_down ? SEQ_DONE_SHUT_DOWN : SEQ_DONE_START_UP;
    result = 100;
  } else {
    total_progress = min_(total_progress + normalised_progress, 99);
    result = (flags & SEQ_FIRST_EXECUTION) ? -1 : total_progress;
  }
  sq.robot[r] = completed;
  sq.robot[np + r] = group;
  sq.robot[2 * np + r] = tstep;
  sq.robot[3 * np + r] = tcount;
  sq.robot[4 * np + r] = flags;
  if (joints_out) sequence_write_joints<S, D>(c, pl, r, joints_out);
  return (flags & SEQ_FAILED) ? -2 : result;
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
