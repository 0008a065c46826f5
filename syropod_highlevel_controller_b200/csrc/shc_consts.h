// shc_consts.h — batch-wide constants of one engine, passed to every kernel as a __grid_constant__ parameter
// (constant bank, uniform access).  Everything here is either a reference parameter (include/shc_config.h) or a
// value the reference's start-up path produces once (include/shc_config.h: shc_startup), pre-digested on the host in
// IEEE double exactly as the reference computes it (e.g. the FP-derived iteration counts of
// walk_controller.cpp:1035-1041 — SURVEY.md §7 "FP-derived integers").
#pragma once
#include "../../include/shc_config.h"

namespace shc {

constexpr int kMaxLegs = SHC_MAX_LEGS;
constexpr int kMaxDof = SHC_MAX_DOF;
constexpr int kMaxPosers = SHC_MAX_AUTO_POSERS;

// Per-leg constants in one precision, array-of-structs so that one uniform base offset per leg addresses them all.
template <class R> struct LegConsts {
  R t1r[9];  // rotation of the constant base transform T1 = DH(link 0), row-major (model.cpp:224, A.1)
  R t1p[3];  // translation of T1
  R dh_d[kMaxDof], dh_theta[kMaxDof], dh_r[kMaxDof];  // links 1..D
  R dh_ca[kMaxDof], dh_sa[kMaxDof];                   // cos/sin(alpha) are constants
  R jmin[kMaxDof], jmax[kMaxDof], vmax[kMaxDof];
  R joffset[kMaxDof];    // added to the published joint command only (state_controller.cpp:795)
  R jcentre[kMaxDof];    // min + range/2                                   (model.cpp:771)
  R jcost_pos[kMaxDof];  // w / range (0 when range == 0)                   (model.cpp:775)
  R jgrad_pos[kMaxDof];  // -w^2 / range^2 (0 when range == 0)              (model.cpp:777)
  R jcost_vel[kMaxDof];  // w / (2 vmax)                                    (model.cpp:784)
  R jgrad_vel[kMaxDof];  // -w^2 / (2 vmax)^2                               (model.cpp:786)
  R identity_x, identity_y;  // identity tip position (walk_controller.cpp:34-42)
  R ysign;                   // +1 when identity y > 0 else -1 (walk_controller.cpp:1244)
  R span_dy;                 // calculateStanceSpanChange().y with the simple one-plane workspace (:949)
  R stance_dt_mod;           // 1/stance_iterations with modified_stance_start = phase offset (:1025-1041)
  R stride_scaler_mod;       // modified_stance_period / stance_period (:1167)
  R neg_ratio;               // negation_transition_ratio (pose_controller.cpp:1763)
};

// Real-valued constants in one precision.
template <class R> struct RealConsts {
  LegConsts<R> leg[kMaxLegs];
  // robot-wide --------------------------------------------------------------------------------------------------
  R dt, inv_dt;
  R swing_dt;          // swing_delta_t_ (:1037)
  R stance_dt_std;     // stance_delta_t_ for the standard stance period
  R stride_scale;      // on_ground_ratio / frequency (:940-941)
  R swing_height, swing_width, body_clearance;
  R lambda2;           // DLS_COEFFICIENT^2 (model.h:19)
  R limits[4][SHC_N_BEARINGS];  // max linear speed, angular speed, linear accel, angular accel (walk_controller.cpp:231)
  R swing_progress_scaler;      // max(1, swing_phase / phase_offset) (pose_controller.cpp:1102)
  // cos/sin of the limit-map bucket edges (walk_controller.cpp:426-431): 44.5, 89.5, 134.5, 179.5, 0.5, 45.5, 90.5, 135.5 deg
  R sec_cos[8], sec_sin[8];
  R inv_swing_period, inv_stance_period;
  R max_translation[3], max_rotation[3], max_translation_velocity, max_rotation_velocity;
  R pid_p, pid_i, pid_d;
  R adm_P[4], adm_q[2];  // 30 RK4 steps of the virtual mass-spring-damper as one affine map x <- P x + q F (A.6)
  R force_gain;
  R virtual_stiffness, swing_stiffness_scaler, load_stiffness_scaler;  // updateStiffness (admittance_controller.cpp:96)
  R body_velocity_scaler;  // bodyVelocityInputCallback (state_controller.cpp:1131)
  R step_depth;            // rough-terrain mode: reactive reach below the default target (walk_controller.cpp:1099)
  R touchdown_threshold, liftoff_threshold;  // Leg::touchdownDetection (model.cpp:712)
  R tip_target_rot[4];     // identity / target tip rotation w x y z with gravity_aligned_tips (walk_controller.cpp:36-41)
  // auto posers (pose_controller.cpp:1338)
  R ap_pos[kMaxPosers][3], ap_rot[kMaxPosers][3], ap_gravity[kMaxPosers];
};

struct IntConsts {
  int L, D;
  int n_robots, n_pad;  // plane stride (n_pad is a multiple of 32)
  // StepCycle (walk_controller.h:23)
  int period, swing_period, stance_period, stance_end, swing_start, swing_end, stance_start;
  int swing_iterations;        // even-rounded (:1035-1036)
  int swing_ref_max;           // largest swing-progress numerator n with (n / swing_period) * swing_progress_scaler <= 1
  int phase_offset[kMaxLegs];
  int mod_stance_start[kMaxLegs];  // = phase offset
  // flags
  int manual_posing, auto_posing, inclination_posing, imu_posing, admittance_control, use_joint_effort, dynamic_stiffness;
  int rough_terrain;           // rough_terrain_mode: default-tip updates, touchdown detection, target shifting
  int tip_mode;                // TIP_NONE / TIP_ALIGN_POSE (D <= 3) / TIP_ROTATION (D > 3): gravity_aligned_tips
  int clamp_joint_positions, clamp_joint_velocities, velocity_input_mode, force_normal_touchdown;
  // auto posing
  int n_posers, pose_phase_length, pose_normaliser, pose_sync, auto_ref_leg;
  int ap_start[kMaxPosers], ap_end[kMaxPosers];
  int neg_start[kMaxLegs], neg_end[kMaxLegs];
  // plane indices (see shc_layout.h)
  int nS, nD, nI;              // number of storage / double / int planes
  int offS_imu, offS_auto, offS_leg, strideS_leg;
  int offS_tip;                // robot-level tip-align block (TIP_ALIGN_POSE)
  int roughS_leg;              // per-leg rough-terrain block, relative to the leg's first joint plane (rough_terrain)
  int tipS_leg;                // per-leg tip-rotation block, relative to the leg's first joint plane (TIP_ROTATION)
  int frontS_leg;              // staged admittance planes in front of each leg's joint planes (0 or 5)
  int smem_per_warp;           // bytes of dynamic shared memory per warp (two staging slots + barriers)
  int offD_leg, strideD_leg;
  int offI_leg, strideI_leg, offI_auto;
};

struct Consts {
  IntConsts i;
  RealConsts<double> d;
  RealConsts<float> f;
};

template <class R> struct ConstSel;
template <> struct ConstSel<double> {
  static __host__ __device__ __forceinline__ const RealConsts<double>& get(const Consts& c) { return c.d; }
};
template <> struct ConstSel<float> {
  static __host__ __device__ __forceinline__ const RealConsts<float>& get(const Consts& c) { return c.f; }
};

}  // namespace shc
