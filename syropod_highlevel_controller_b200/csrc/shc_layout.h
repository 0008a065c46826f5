// shc_layout.h — device state layout: struct-of-arrays planes, robot index fastest (DESIGN.md "Data layout in HBM").
//
//   storage planes  S[nS][n_pad]   S = float (mixed) or double (f64): everything that is not an open-loop accumulator
//   double planes   D[nD][n_pad]   open-loop integrated accumulators: stepper tip position, odometry position
//   int planes      I[nI][n_pad]   packed integer state
//
// A warp of 32 consecutive robots reads/writes 32 consecutive words of a plane: one fully used 128-byte line (256 B for
// double planes).  Field -> reference member mapping is the one of include/shc_state.h.
#pragma once

namespace shc {

// robot-level storage planes (always present)
enum : int {
  RS_VEL = 0,    // desired_linear_velocity_ (2)
  RS_ANGVEL = 2, // desired_angular_velocity_
  RS_WPL = 3,    // WalkController::walk_plane_ (3)
  RS_WPN = 6,    // WalkController::walk_plane_normal_ (3)
  RS_ODOMQ = 9,  // odometry_ideal_.rotation_ (4: w x y z)
  RS_WPP = 13,   // walk_plane_pose_ (7: p, q)
  RS_OWPP = 20,  // origin_walk_plane_pose_ (7)
  RS_MAN = 27,   // manual_pose_ (7)
  RS_COUNT = 34
};
// optional IMU block (imu_posing || inclination_posing), relative to offS_imu
enum : int { IMU_Q = 0 /*imu_pose_.rotation_ (4)*/, IMU_ABS = 4 /*absement (3)*/, IMU_VEL = 7 /*velocity err (3)*/,
             IMU_INCL = 10 /*inclination_pose_ x,y (2)*/, IMU_POS = 12 /*rotation_position_error_ (3): only read back by
             publishRotationPoseError*/, IMU_COUNT = 15 };
// optional auto-pose block (auto_posing), relative to offS_auto
enum : int { AUTO_POSE = 0 /*auto_pose_ (7)*/, AUTO_COUNT = 7 };
// optional tip-align block (gravity_aligned_tips on legs of at most three joints), relative to offS_tip
enum : int { TA_POSE = 0 /*tip_align_pose_ (7)*/, TA_ORIGIN = 7 /*origin_tip_align_pose_ (7)*/, TA_COUNT = 14 };
// optional per-leg tip-rotation block (gravity_aligned_tips on legs of more than three joints), relative to
// offS_leg + leg * strideS_leg + tipS_leg: quaternions w x y z, all zero = UNDEFINED_ROTATION
enum : int { TR_CUR = 0 /*LegStepper::current_tip_pose_.rotation_*/, TR_ORIGIN = 4 /*origin_tip_pose_.rotation_*/, TR_COUNT = 8 };
// optional per-leg rough-terrain block (rough_terrain_mode), relative to offS_leg + leg * strideS_leg + roughS_leg
enum : int { RT_STEP_PLANE = 0 /*Leg::step_plane_pose_.position_ (3), base_link frame*/,
             RT_ET_POSE = 3 /*external_target_.pose_ (7)*/, RT_ET_TF = 10 /*external_target_.transform_ (7)*/,
             RT_ET_CLR = 17 /*external_target_.swing_clearance_*/, RT_ED_POSE = 18 /*external_default_.pose_ (7)*/,
             RT_ED_TF = 25 /*external_default_.transform_ (7)*/, RT_COUNT = 32 };
// tip_mode of an engine
enum : int { TIP_NONE = 0, TIP_ALIGN_POSE = 1, TIP_ROTATION = 2 };

// per-leg storage planes, relative to offS_leg + leg * strideS_leg (the plane of the leg's first joint position).
// The planes every cycle reads come first and are contiguous: [front .. STAGED) is what one TMA bulk copy stages into
// shared memory per leg (shc_cycle.cuh); the rest is write-mostly and goes to HBM with plain coalesced stores.
template <int D> struct LegS {
  enum : int {
    // optional admittance planes IN FRONT of the joint planes (negative offsets; present when admittance_control ||
    // use_joint_effort): they are read every cycle, so they sit inside the staged range
    ADM_X = -5,          // admittance_state_ (2)
    ADM_FORCE = -3,      // tip_force_calculated_ (3)
    Q = 0, QD = D,
    DEF = 2 * D,         // default_tip_pose_.position_
    STRIDE = 2 * D + 3,  // stride_vector_
    SWO_P = 2 * D + 6,   // swing_origin_tip_position_
    SWO_V = 2 * D + 9,   // swing_origin_tip_velocity_
    STAGED = 2 * D + 12, // end of the staged range
    TIPVEL = 2 * D + 12, // LegStepper::current_tip_velocity_
    STO_P = 2 * D + 15,  // stance_origin_tip_position_
    TGT = 2 * D + 18,    // target_tip_pose_.position_
    WP = 2 * D + 21,     // walk_plane_ (saved)
    WPN = 2 * D + 24,    // walk_plane_normal_ (saved)
    COUNT = 2 * D + 27,
    ADM_DELTA = COUNT,   // admittance_delta_ (3), appended when the admittance block is present
    STIFF = COUNT + 3    // virtual_stiffness_ (dynamic stiffness, admittance_controller.cpp:96), same block
  };
};
// the same offsets for host code that knows D only at run time (pack)
struct LegOff {
  int Q, QD, DEF, STRIDE, SWO_P, SWO_V, STAGED, TIPVEL, STO_P, TGT, WP, WPN, COUNT, ADM_X, ADM_FORCE, ADM_DELTA, STIFF;
  explicit LegOff(int D)
      : Q(0), QD(D), DEF(2 * D), STRIDE(2 * D + 3), SWO_P(2 * D + 6), SWO_V(2 * D + 9), STAGED(2 * D + 12), TIPVEL(2 * D + 12),
        STO_P(2 * D + 15), TGT(2 * D + 18), WP(2 * D + 21), WPN(2 * D + 24), COUNT(2 * D + 27), ADM_X(-5), ADM_FORCE(-3),
        ADM_DELTA(2 * D + 27), STIFF(2 * D + 30) {}
};
enum : int { ADM_COUNT = 9 };  // planes of the optional per-leg admittance block (5 in front + 4 appended)

// double planes
enum : int { RD_ODOMP = 0 /*odometry_ideal_.position_ (3)*/, RD_COUNT = 3 };
enum : int { LD_TIP = 0 /*LegStepper::current_tip_pose_.position_ (3)*/, LD_COUNT = 3 };

// int planes
enum : int { RI_BITS = 0, RI_COUNT = 1 };
enum : int { AI_FLAGS = 0 /*4 latch bits per AutoPoser*/, AI_PHASE = 1 /*pose_phase_*/, AI_COUNT = 2 };
enum : int { LI_BITS = 0, LI_PROG = 1, LI_COUNT = 2 };

// RI_BITS: walk_state[0:2) legs_at_correct_phase[2:6) legs_completed_first_step[6:10) return_to_default_attempted[10]
//          pose_state[11:13) auto_posing_state[13:15) walk plane changed by the last cycle / unknown [15]  status flags [16:28)
//          manual pose is exactly the identity [30]  walk plane not known to be the least-squares plane of the stored
//          default tips (state written from outside) [31]
// The last three are bookkeeping of the engine, not reference state: they let a cycle skip work whose result is already
// in HBM (the saved plane copies of the legs, the plane fit over unchanged default tips, the identity manual pose).
enum : int { RB_PLANE_CHANGED = 15, RB_STATUS_SHIFT = 16, RB_STATUS_MASK = 0xfff, RB_MANUAL_IDENTITY = 30, RB_PLANE_STALE = 31 };
// LI_BITS: phase[0:16) step_state[16:18) at_correct_phase[18] completed_first_step[19] negate_auto_pose[20]
//          the leg's saved walk plane (WP / WPN planes) equals the walker's plane of the previous cycle's start [21]
//          rough-terrain mode: step_plane_pose_ is defined [22]  touchdown_detection_ [23]  external_target_.defined_ [24]
//          external target in the odom_ideal frame [25]  external_default_.defined_ [26]
enum : int { LB_PLANE_SAVED = 21, LB_STEP_PLANE = 22, LB_TOUCHDOWN = 23, LB_EXT_TARGET = 24, LB_EXT_ODOM = 25, LB_EXT_DEFAULT = 26 };
// LI_PROG: swing progress numerator (int16, -1 = "-1.0") | stance progress numerator (int16) << 16
//          progress = numerator / swing_period (resp. stance_period): walk_controller.cpp:878-896 divides two ints
//          converted to double, so keeping the numerator makes the value exact in every precision.

enum : int { WALK_STARTING = 0, WALK_MOVING = 1, WALK_STOPPING = 2, WALK_STOPPED = 3 };   // parameters_and_states.h:103
enum : int { STEP_SWING = 0, STEP_STANCE = 1, STEP_FORCE_STANCE = 2, STEP_FORCE_STOP = 3 }; // :116
enum : int { POSE_POSING = 0, POSE_STOP_POSING = 1, POSE_COMPLETE = 2 };                    // :128

}  // namespace shc
