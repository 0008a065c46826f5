// shc_cycle.cuh — one fused control cycle for one robot (device code shared by the control_cycle kernels).
//
// Replaces, in the reference's call order (state_controller.cpp:162-193, 379-447):
//   PoseController::updateCurrentPose      pose_controller.cpp:811   (walk-plane :1092, manual :863, inclination :1240,
//                                                                     IMU PID :1191, auto :1134/:1338/:1716)
//   AdmittanceController::updateAdmittance admittance_controller.cpp:22
//   WalkController::updateWalk             walk_controller.cpp:440   (getLimit :414, LegStepper::updateTipPosition :1018,
//                                                                     iteratePhase :871, updateWalkPlane :748, odometry :783)
//   PoseController::updateStance           pose_controller.cpp:110
//   Model::updateModel                     model.cpp:142             (setDesiredTipPose :653, applyIK :861, solveIK :726,
//                                                                     updateJointPositions :799, applyFK :945,
//                                                                     calculateTipForce :667)
//
// Precision policy P: S = storage planes, T = trajectory arithmetic, K = pose/kinematics arithmetic.  Branch decisions
// the reference takes on doubles (limit-map bucket, walk-plane progress window, stop tolerance) are always taken in
// double from the double accumulators, so the fp32 instantiation follows the same control flow as the reference.
#pragma once
#include "shc_consts.h"
#include "shc_layout.h"
#include "shc_math.cuh"
#include <cstring>

namespace shc {

template <class S_, class T_, class K_> struct Prec {
  using S = S_;
  using T = T_;
  using K = K_;
};
using PrecF64 = Prec<double, double, double>;

struct StepIO {
  const float* cmd;        // [N][3]
  const float* imu;        // [N][10] or null
  const float* tip_force;  // [N][L][3] or null
  const float* manual;     // [N][6] or null
  const float* efforts;    // [N][L][D] or null: measured joint efforts (jointStatesCallback, state_controller.cpp:1565)
  const float* step_planes;  // [N][L][3] or null: TipState.step_plane of the tip range sensors (x, y slopes, z range;
                             // z >= SHC_RANGE_UNASSIGNED = no reading), tipStatesCallback (state_controller.cpp:1650-1675)
  float* joints_out;       // [N][L][D] (this shard's joint commands)
  // Fused all-gather (multi-GPU): when n_gather > 0 the tile is ALSO stored into every rank's gather buffer
  // gather[p][gather_offset + ...], p = 0 .. n_gather-1 — peer-mapped device pointers (CUDA IPC), i.e. plain stores that
  // travel over NVLink / NVSwitch while the other tiles are still being computed.  The own rank's buffer is one of them.
  float* gather[8];
  float* gather_mc;         // multicast mapping of the gather buffers (NVLS) or null: one store reaches every rank
  long long gather_offset;  // elements: buffer index * world * N * L * D + rank * N * L * D
  int n_gather;
  int tile_begin, tile_end;  // tiles (32 robots each) of this launch: a step may be issued as several tile ranges so that
                             // the D2H copy of one range overlaps the arithmetic of the next (shc_step_host)
  int* flags_out;          // [N] or null
  int pose_reset_mode;
#ifdef SHC_TRACE
  unsigned long long* trace;  // kernel-tuning builds only (-DSHC_TRACE): [tiles][32] globaltimer stamps written by lane 0
#endif
};
// Kernel-tuning builds (-DSHC_TRACE, tools/dev_trace.py): per-tile timeline of the control cycle.  Compiles to nothing otherwise.
#ifdef SHC_TRACE
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SHC_STAMP(k) do { if (io.trace && lane == 0) io.trace[(size_t)tile_idx * 32 + (k)] = gtimer(); } while (0)
#else
#define SHC_STAMP(k) do { } while (0)
#endif

// ---- TMA bulk copies + mbarrier (sm_90+; SASS: UBLKCP.S.G / SYNCS) ------------------------------------------------
// Each warp owns a two-slot staging ring in shared memory.  Lane 0 asks the TMA unit for the whole contiguous chunk of
// a leg's every-cycle planes (one cp.async.bulk per plane set) two legs ahead of the arithmetic; completion is
// signalled on a warp-private mbarrier by byte count.  The dependent chain of a robot therefore never waits on HBM
// inside the leg loop: its loads are shared-memory reads of data that landed while the previous legs were computed.
//
// -DSHC_EMU (tests/cpp/shc_emu.cpp only, compiled by plain g++ — never the product library): Cycle::run is compiled for the
// host, where the staging primitives below degenerate to a memcpy per lane and the lanes of a tile run one after the other.  That lets the
// CPU test-suite drive the very source the kernel is built from against the oracle; it is not a fallback of the engine.
#if defined(SHC_EMU)
#define SHC_CYCLE_FN __host__ __device__
#else
#define SHC_CYCLE_FN __device__
#endif
#if defined(__CUDACC__)
#define SHC_LANE0(lane) ((lane) == 0)
#define SHC_SYNCWARP() __syncwarp()
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk store shared -> global (the destination may be peer-mapped memory of another GPU): asynchronous, one request
// for the whole byte range instead of one 128-byte store per warp instruction.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void store_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_sys_add(int* p, int v) {
  asm volatile("red.release.sys.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int load_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// NVSwitch multicast (NVLS): a store to the multicast mapping of a symmetric buffer lands in every rank's copy.
__device__ __forceinline__ void multimem_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_f32(float* mc, float v) {
  asm volatile("multimem.st.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_red_release_add(int* mc, int v) {
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
// cp.async (LDGSTS): per-lane asynchronous global -> shared copies that need no registers while in flight
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
#else
// host pass (SHC_EMU): every lane stages its own copy synchronously
#define SHC_LANE0(lane) (true)
#define SHC_SYNCWARP() do { } while (0)
inline void mbar_init(uint64_t*, unsigned) {}
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_expect_tx(uint64_t*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t*) { std::memcpy(dst, src, bytes); }
inline void mbar_wait(uint64_t*, unsigned) {}
inline void cp_async_16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
inline void cp_async_4(void* dst, const void* src) { std::memcpy(dst, src, 4); }
inline void cp_async_wait_all() {}
#endif
// The Euler decomposition (three atan2, ~700 instructions) as ONE out-of-line copy per kernel: the pose stages call it from
// up to five places, and inlining every call bloats the instruction footprint of a kernel that is fetch-bound already.
#if defined(__CUDA_ARCH__)
template <class R> __device__ __noinline__ V3<R> quat_to_euler_nc(Q4<R> q, bool intrinsic) { return quat_to_euler(q, intrinsic); }
#else
template <class R> inline V3<R> quat_to_euler_nc(Q4<R> q, bool intrinsic) { return quat_to_euler(q, intrinsic); }
#endif

template <class S> struct Planes {
  S* s;
  double* d;
  int* i;
};

// DH chain of one leg in the leg (joint-1) frame.  Frame i (i = 1..D-1) is the frame of joint i+1, i.e. the product
// T2..T(i+1) of A.1; `tip` is T2..TD*Ttip.  Composition uses the sparse DH form
//   x' = x c + y s ; u = y c - x s ; y' = u ca + z sa ; z' = z ca - u sa ; p' = p + r x' + d z
template <class K, int D> struct Chain {
  V3<K> z[D];  // joint axes in the leg frame (z[0] = (0,0,1))
  V3<K> p[D];  // joint origins in the leg frame (p[0] = 0)
  V3<K> tip;   // tip position in the leg frame
  V3<K> tipx;  // x axis of the tip frame in the leg frame (last-link direction)
  V3<K> tipy, tipz;  // its other two axes (only the tip-orientation path reads them)
};

template <class K, int D, class QT>
SHC_HD void leg_chain(const LegConsts<K>& lc, const QT* q, Chain<K, D>& ch) {
  V3<K> ax{K(1), K(0), K(0)}, ay{K(0), K(1), K(0)}, az{K(0), K(0), K(1)}, ap{K(0), K(0), K(0)};
#pragma unroll
  for (int j = 0; j < D; ++j) {
    ch.z[j] = az;
    ch.p[j] = ap;
    K s, c;
    sincos_(lc.dh_theta[j] + K(q[j]), &s, &c);
    const K ca = lc.dh_ca[j], sa = lc.dh_sa[j];
    V3<K> nx = ax * c + ay * s;
    V3<K> u = ay * c - ax * s;
    V3<K> ny = u * ca + az * sa;
    V3<K> nz = az * ca - u * sa;
    ap = ap + nx * lc.dh_r[j] + az * lc.dh_d[j];
    ax = nx; ay = ny; az = nz;
  }
  ch.tip = ap;
  ch.tipx = ax;
  ch.tipy = ay;
  ch.tipz = az;
}

template <class K> SHC_HD V3<K> t1_rotate(const LegConsts<K>& lc, V3<K> v) {
  const K* r = lc.t1r;
  return {r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z};
}
template <class K> SHC_HD V3<K> t1_rotate_inv(const LegConsts<K>& lc, V3<K> v) {
  const K* r = lc.t1r;
  return {r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z};
}

// Solve the symmetric positive definite 3x3 system A x = b (A = Jp Jp^T + lambda^2 I) by Cholesky.
template <class K> SHC_HD V3<K> spd3_solve(K a00, K a01, K a02, K a11, K a12, K a22, V3<K> b) {
  // only the reciprocals of the Cholesky diagonal are needed
  K i00 = rsqrt_(a00);
  K l10 = a01 * i00, l20 = a02 * i00;
  K i11 = rsqrt_(a11 - l10 * l10);
  K l21 = (a12 - l20 * l10) * i11;
  K i22 = rsqrt_(a22 - l20 * l20 - l21 * l21);
  K y0 = b.x * i00;
  K y1 = (b.y - l10 * y0) * i11;
  K y2 = (b.z - l20 * y0 - l21 * y1) * i22;
  K x2 = y2 * i22;
  K x1 = (y1 - l21 * x2) * i11;
  K x0 = (y0 - l10 * x1 - l20 * x2) * i00;
  return {x0, x1, x2};
}

// Solve the symmetric positive definite DxD system A x = b in place (Cholesky), D <= 5.
template <class K, int D> SHC_HD void spdN_solve(K A[D][D], K b[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      K sum = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) sum -= A[i][k] * A[j][k];
      if (i == j) A[i][i] = sqrt_(sum);
      else A[i][j] = sum / A[j][j];
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    K sum = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) sum -= A[i][k] * b[k];
    b[i] = sum / A[i][i];
  }
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    K sum = b[i];
#pragma unroll
    for (int k = i + 1; k < D; ++k) sum -= A[k][i] * b[k];
    b[i] = sum / A[i][i];
  }
}

// Leg::calculateTipForce (model.cpp:667-708): F_leg = Jp (J^T J + l^2 I_D)^-1 tau with the full 6xD Jacobian
// (angular rows = joint axes), rotated by the rotation of T1^-1, before the low-pass filter.
template <class K, int D>
SHC_HD V3<K> raw_tip_force(const RealConsts<K>& ck, const LegConsts<K>& lc, const Chain<K, D>& ch, const K* tau) {
  V3<K> Jp[D];
#pragma unroll
  for (int j = 0; j < D; ++j) Jp[j] = cross(ch.z[j], ch.tip - ch.p[j]);
  K A[D][D];
  K x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x[i] = tau[i];
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = dot(Jp[i], Jp[j]) + dot(ch.z[i], ch.z[j]) + (i == j ? ck.lambda2 : K(0));
  }
  spdN_solve<K, D>(A, x);
  V3<K> f{K(0), K(0), K(0)};
#pragma unroll
  for (int j = 0; j < D; ++j) f = f + Jp[j] * x[j];
  return t1_rotate_inv(lc, f);
}

// Leg::solveIK (position rows only: the angular rows of J are zero when solve_rotation is false, so
// J^T (J J^T + l^2 I6)^-1 [dp;0] = Jp^T (Jp Jp^T + l^2 I3)^-1 dp and J^+ J = Jp^T (..)^-1 Jp — model.cpp:726-795).
template <class K, int D>
SHC_HD void solve_ik(const RealConsts<K>& ck, const LegConsts<K>& lc, const Chain<K, D>& ch, V3<K> dp, const K* q,
                                         const K* qd, K* dq) {
  V3<K> J[D];
#pragma unroll
  for (int j = 0; j < D; ++j) J[j] = cross(ch.z[j], ch.tip - ch.p[j]);
  K a00 = ck.lambda2, a01 = K(0), a02 = K(0), a11 = ck.lambda2, a12 = K(0), a22 = ck.lambda2;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    a00 += J[j].x * J[j].x; a01 += J[j].x * J[j].y; a02 += J[j].x * J[j].z;
    a11 += J[j].y * J[j].y; a12 += J[j].y * J[j].z; a22 += J[j].z * J[j].z;
  }
  // joint limit cost gradient (model.cpp:759-790)
  K g[D];
  K pos_cost = K(0), vel_cost = K(0);
  K gp[D], gv[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    K e = q[j] - lc.jcentre[j];
    K cp = lc.jcost_pos[j] * e;
    pos_cost += cp * cp;
    gp[j] = lc.jgrad_pos[j] * e;
    K cv = lc.jcost_vel[j] * qd[j];
    vel_cost += cv * cv;
    gv[j] = lc.jgrad_vel[j] * qd[j];
  }
  K sp = pos_cost == K(0) ? K(0) : rsqrt_(pos_cost);
  K sv = vel_cost == K(0) ? K(0) : rsqrt_(vel_cost);
  V3<K> Jg{K(0), K(0), K(0)};
#pragma unroll
  for (int j = 0; j < D; ++j) {
    g[j] = K(0.25) * (gp[j] * sp) + K(0.75) * (gv[j] * sv);
    Jg = Jg + J[j] * g[j];
  }
  // dq = Jp^T A^-1 (dp - Jp g) + g
  V3<K> x = spd3_solve(a00, a01, a02, a11, a12, a22, dp - Jg);
#pragma unroll
  for (int j = 0; j < D; ++j) dq[j] = dot(J[j], x) + g[j];
}

// Leg::applyIK without the trailing FK (model.cpp:861-904): position delta in the leg frame from the chain at the
// current joint angles, one DLS step, Leg::updateJointPositions (:799).  q/qd are updated in place; returns the
// SHC_FLAG_*_CLAMP bits.  `des_leg_out` receives the desired tip position in the leg frame.
template <class K, int D>
SHC_HD int apply_ik_step(const RealConsts<K>& ck, const LegConsts<K>& lc, const Chain<K, D>& ch, K* q, K* qd, V3<K> desired_robot,
                         bool clamp_positions, bool clamp_velocities, V3<K>* des_leg_out) {
  V3<K> des_leg = t1_rotate_inv(lc, desired_robot - V3<K>{lc.t1p[0], lc.t1p[1], lc.t1p[2]});
  *des_leg_out = des_leg;
  K dq[D];
  solve_ik<K, D>(ck, lc, ch, des_leg - ch.tip, q, qd, dq);
  int status = 0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    K v = dq[j] * ck.inv_dt;
    if (clamp_velocities && abs_(v) > lc.vmax[j]) {
      v = clamp_(v, -lc.vmax[j], lc.vmax[j]);
      status |= 4;
    }
    K nq = q[j] + v * ck.dt;
    if (clamp_positions) {
      if (nq < lc.jmin[j]) { nq = lc.jmin[j]; status |= 2; }
      else if (nq > lc.jmax[j]) { nq = lc.jmax[j]; status |= 2; }
    }
    q[j] = nq;
    qd[j] = v;
  }
  return status;
}

// Leg::updateJointPositions (model.cpp:799): q/qd from a joint position delta; returns the SHC_FLAG_*_CLAMP bits.
template <class K, int D>
SHC_HD int update_joint_positions(const RealConsts<K>& ck, const LegConsts<K>& lc, const K* dq, K* q, K* qd, bool clamp_positions,
                                  bool clamp_velocities) {
  int status = 0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    K v = dq[j] * ck.inv_dt;
    if (clamp_velocities && abs_(v) > lc.vmax[j]) {
      v = clamp_(v, -lc.vmax[j], lc.vmax[j]);
      status |= 4;
    }
    K nq = q[j] + v * ck.dt;
    if (clamp_positions) {
      if (nq < lc.jmin[j]) { nq = lc.jmin[j]; status |= 2; }
      else if (nq > lc.jmax[j]) { nq = lc.jmax[j]; status |= 2; }
    }
    q[j] = nq;
    qd[j] = v;
  }
  return status;
}

// Leg::solveIK with solve_rotation = true and delta = [0; drot] (model.cpp:726-795): the full 6 x D Jacobian (linear rows
// z_j x (tip - p_j), angular rows z_j), dq = J^T (J J^T + l^2 I6)^-1 (delta - J g) + g with the joint-limit cost gradient g.
template <class K, int D>
SHC_HD void solve_ik_rotation(const RealConsts<K>& ck, const LegConsts<K>& lc, const Chain<K, D>& ch, V3<K> drot, const K* q, const K* qd,
                              K* dq) {
  K J[6][D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const V3<K> jp = cross(ch.z[j], ch.tip - ch.p[j]);
    J[0][j] = jp.x; J[1][j] = jp.y; J[2][j] = jp.z;
    J[3][j] = ch.z[j].x; J[4][j] = ch.z[j].y; J[5][j] = ch.z[j].z;
  }
  K g[D];
  {
    K pos_cost = K(0), vel_cost = K(0), gp[D], gv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      K e = q[j] - lc.jcentre[j];
      K cp = lc.jcost_pos[j] * e;
      pos_cost += cp * cp;
      gp[j] = lc.jgrad_pos[j] * e;
      K cv = lc.jcost_vel[j] * qd[j];
      vel_cost += cv * cv;
      gv[j] = lc.jgrad_vel[j] * qd[j];
    }
    K sp = pos_cost == K(0) ? K(0) : rsqrt_(pos_cost);
    K sv = vel_cost == K(0) ? K(0) : rsqrt_(vel_cost);
#pragma unroll
    for (int j = 0; j < D; ++j) g[j] = K(0.25) * (gp[j] * sp) + K(0.75) * (gv[j] * sv);
  }
  K A[6][6], b[6];
  const K delta[6] = {K(0), K(0), K(0), drot.x, drot.y, drot.z};
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    K jg = K(0);
#pragma unroll
    for (int j = 0; j < D; ++j) jg += J[a][j] * g[j];
    b[a] = delta[a] - jg;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      K s = a == c ? ck.lambda2 : K(0);
#pragma unroll
      for (int j = 0; j < D; ++j) s += J[a][j] * J[c][j];
      A[a][c] = s;
    }
  }
  spdN_solve<K, 6>(A, b);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    K s = g[j];
#pragma unroll
    for (int a = 0; a < 6; ++a) s += J[a][j] * b[a];
    dq[j] = s;
  }
}

template <class K, int D> SHC_HD bool ik_within_tolerance(const LegConsts<K>& lc, const Chain<K, D>& ch, V3<K> des_leg) {
  V3<K> er = t1_rotate(lc, ch.tip - des_leg);  // current - desired tip position in the base_link frame
  return !(abs_(er.x) > K(0.005) || abs_(er.y) > K(0.005) || abs_(er.z) > K(0.005));
}

// Leg::applyIK complete (model.cpp:861-941) for a desired tip POSE: the position step; when the desired rotation is defined
// (gravity_aligned_tips on legs of more than three joints) a second, rotation-only step on the updated chain that turns
// the last link towards the desired direction; the IK tolerance check; and, when a rotation-constrained attempt misses the
// position tolerance or ends on a joint limit, one more position-only step (the reference recurses with the rotation dropped, :932-936).
// `ch` is the chain at q on entry and at the new q on return.  *force_updates = how often calculateTipForce ran (the
// recursion runs it at both levels).  Returns the clamp status bits; *ok = within IK_TOLERANCE.
template <class K, int D>
SHC_HD int apply_ik_pose(const RealConsts<K>& ck, const LegConsts<K>& lc, Chain<K, D>& ch, K* q, K* qd, V3<K> desired_robot,
                         Q4<K> desired_rot, bool clamp_positions, bool clamp_velocities, V3<K>* des_leg_out, bool* ok, int* force_updates) {
  const V3<K> des_leg = t1_rotate_inv(lc, desired_robot - V3<K>{lc.t1p[0], lc.t1p[1], lc.t1p[2]});
  *des_leg_out = des_leg;
  *force_updates = 1;
  K dq[D];
  int status = 0;
  const bool rotation_constrained = !(desired_rot.w == K(0) && desired_rot.x == K(0) && desired_rot.y == K(0) && desired_rot.z == K(0));
  solve_ik<K, D>(ck, lc, ch, des_leg - ch.tip, q, qd, dq);
  if (rotation_constrained) {
    const V3<K> current_dir = ch.tipx;  // the leg-frame tip direction the cycle started with (:866, :889)
    status |= update_joint_positions<K, D>(ck, lc, dq, q, qd, clamp_positions, false);  // updateJointPositions(delta, true)
    leg_chain<K, D>(lc, q, ch);
    const V3<K> desired_dir = t1_rotate_inv(lc, qrot(qnormalized(desired_rot), V3<K>{K(1), K(0), K(0)}));
    const V3<K> drot = rotation_vector(qnormalized(from_two_vectors(current_dir, desired_dir)));
    solve_ik_rotation<K, D>(ck, lc, ch, drot, q, qd, dq);
  }
  status |= update_joint_positions<K, D>(ck, lc, dq, q, qd, clamp_positions, clamp_velocities);
  leg_chain<K, D>(lc, q, ch);
  *ok = ik_within_tolerance<K, D>(lc, ch, des_leg);
  // "!ik_success" (:932) tests the DOUBLE applyIK would return: zero on a missed tolerance, but also when the smallest
  // joint-limit proximity is exactly zero, i.e. when the update left a joint sitting on one of its limits
  bool on_limit = false;
#pragma unroll
  for (int j = 0; j < D; ++j) on_limit = on_limit || (lc.jmax[j] != lc.jmin[j] && (q[j] == lc.jmin[j] || q[j] == lc.jmax[j]));
  if (rotation_constrained && (!*ok || on_limit)) {
    solve_ik<K, D>(ck, lc, ch, des_leg - ch.tip, q, qd, dq);
    status |= update_joint_positions<K, D>(ck, lc, dq, q, qd, clamp_positions, clamp_velocities);
    leg_chain<K, D>(lc, q, ch);
    *ok = ik_within_tolerance<K, D>(lc, ch, des_leg);
    *force_updates = 2;
  }
  return status;
}

// Return value of Leg::applyIK (model.cpp:845-856, 916-929): the smallest joint-limit proximity, or 0 when the tip
// deviates from the desired position by more than IK_TOLERANCE on a base_link axis.  `ch2` is the chain at the NEW q.
template <class K, int D>
SHC_HD K ik_result_value(const LegConsts<K>& lc, const Chain<K, D>& ch2, const K* q, V3<K> des_leg) {
  K prox = K(1);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    K min_diff = abs_(lc.jmin[j] - q[j]);
    K max_diff = abs_(lc.jmax[j] - q[j]);
    K half = (lc.jmax[j] - lc.jmin[j]) / K(2);
    K lp = half != K(0) ? min_(min_diff, max_diff) / half : K(1);
    prox = min_(lp, prox);
  }
  V3<K> er = t1_rotate(lc, ch2.tip - des_leg);
  if (abs_(er.x) > K(0.005) || abs_(er.y) > K(0.005) || abs_(er.z) > K(0.005)) prox = K(0);
  return prox;
}

// Model::estimateGravity (model.cpp:156) from the raw IMU orientation (all-zero quaternion when no IMU data: the
// rotation matrix of the zero quaternion is the identity).
template <class K> SHC_HD V3<K> estimate_gravity(Q4<K> imu_raw) {
  V3<K> e = quat_to_euler_nc(imu_raw, false);
  K s, c;
  sincos_(-e.y, &s, &c);  // rotate (0,0,g) about Y by -pitch
  const K g = K(-9.81);
  V3<K> v{s * g, K(0), c * g};
  sincos_(-e.x, &s, &c);  // then about X by -roll
  return {v.x, c * v.y - s * v.z, s * v.y + c * v.z};
}

// The AutoPoser loop of PoseController::updateAutoPose (pose_controller.cpp:1163-1178, AutoPoser::updatePose :1338): every
// poser advances its start / end latches (4 bits each in `pflags`) and contributes its quartic-Bezier sway; returns the
// summed auto pose and sets auto_state to POSE_COMPLETE when no poser is posing any more.
template <class K>
SHC_HD PoseT<K> auto_posers_update(const IntConsts& ci, const RealConsts<K>& ck, int master_phase, Q4<K> imu_raw, int& auto_state,
                                   int& pflags) {
  PoseT<K> auto_pose = pose_identity<K>();
  int complete = 0;
  V3<K> grav_dir{K(0), K(0), K(-1)};
  bool grav_done = false;
  for (int a = 0; a < ci.n_posers; ++a) {
    int f = (pflags >> (4 * a)) & 15;
    bool start_check = f & 1, end1 = f & 2, end2 = f & 4, allow = f & 8;
    int phase = master_phase;
    int start_phase = ci.ap_start[a] * ci.pose_normaliser;
    int end_phase = ci.ap_end[a] * ci.pose_normaliser;
    if (start_phase > end_phase) {
      end_phase += ci.pose_phase_length;
      if (phase < start_phase) phase += ci.pose_phase_length;
    }
    start_check = !ci.pose_sync || (!start_check && auto_state == POSE_POSING && phase == start_phase);
    end1 = end1 || (auto_state == POSE_STOP_POSING && phase == start_phase);
    end2 = end2 || (auto_state == POSE_STOP_POSING && phase == end_phase && end1);
    if (!allow && start_check) {
      allow = true;
      end1 = end2 = false;
    } else if (allow && ci.pose_sync && end1 && end2) {
      allow = false;
      start_check = false;
    }
    PoseT<K> upd = pose_identity<K>();
    if (phase >= start_phase && phase < end_phase && allow) {
      int iteration = phase - start_phase + 1;
      int num_iterations = end_phase - start_phase;
      bool first_half = iteration <= num_iterations / 2;
      V3<K> rot_amp{ck.ap_rot[a][0], ck.ap_rot[a][1], ck.ap_rot[a][2]};
      V3<K> pos_amp{ck.ap_pos[a][0], ck.ap_pos[a][1], ck.ap_pos[a][2]};
      if (ck.ap_gravity[a] != K(0)) {
        if (!grav_done) { grav_dir = normalized(estimate_gravity(imu_raw)); grav_done = true; }
        pos_amp = grav_dir * ck.ap_gravity[a];
      }
      K delta_t = K(1) / (K(num_iterations) / K(2));
      int offset = (int)(first_half ? 0.0 : num_iterations / 2.0);
      K t = K(iteration - offset) * delta_t;
      K s = K(1) - t;
      // quartic Bezier with nodes {0,0,0,A,A} (first half) or {A,A,0,0,0} (second half)
      K w = first_half ? (K(4) * t * t * t * s + t * t * t * t) : (s * s * s * s + K(4) * t * s * s * s);
      upd.p = pos_amp * w;
      upd.q = euler_to_quat(rot_amp * w, false);
    }
    complete += allow ? 0 : 1;
    auto_pose = pose_add(auto_pose, upd);
    f = (start_check ? 1 : 0) | (end1 ? 2 : 0) | (end2 ? 4 : 0) | (allow ? 8 : 0);
    pflags = (pflags & ~(15 << (4 * a))) | (f << (4 * a));
  }
  if (complete == ci.n_posers) auto_state = POSE_COMPLETE;
  return auto_pose;
}

// LegPoser::updateAutoPose (pose_controller.cpp:1716): the leg's own auto pose = the body's with the leg's negation
// window blended out; `negate` is the leg's latch, `step_state` the state the cycle found the leg in.
template <class K>
SHC_HD PoseT<K> leg_auto_pose(const IntConsts& ci, K neg_ratio, int l, int master_phase, int step_state, const PoseT<K>& auto_pose,
                              bool& negate) {
  int start_phase = ci.neg_start[l] * ci.pose_normaliser;
  int end_phase = ci.neg_end[l] * ci.pose_normaliser;
  int negation_phase = master_phase;
  if (start_phase == 0) start_phase = ci.pose_phase_length;
  if (end_phase == 0) end_phase = ci.pose_phase_length;
  if (start_phase > end_phase) {
    end_phase += ci.pose_phase_length;
    if (negation_phase < start_phase) negation_phase += ci.pose_phase_length;
  }
  if (step_state != STEP_FORCE_STANCE && step_state != STEP_FORCE_STOP && negation_phase == start_phase) negate = true;
  if (negation_phase < start_phase || negation_phase > end_phase) negate = false;
  if (!negate) return auto_pose;
  int iteration = negation_phase - start_phase + 1;
  int num_iterations = end_phase - start_phase;
  bool first_half = iteration <= num_iterations / 2;
  K ctrl = K(1);
  if (neg_ratio > K(0)) {
    if (first_half) ctrl = min_(K(1), K(iteration) / (K(num_iterations) * neg_ratio));
    else ctrl = min_(K(1), K(num_iterations - iteration) / (K(num_iterations) * neg_ratio));
  }
  ctrl = smooth_step(ctrl);
  PoseT<K> negation = pose_interpolate(pose_identity<K>(), ctrl, auto_pose);
  return pose_remove(auto_pose, negation);
}

// ---------------------------------------------------------------------------------------------------------------------
// FULL = false compiles the walking-only engine (no auto / IMU / inclination posing, no admittance): the optional stages
// and the registers they keep alive across the leg loop disappear at compile time.
#ifndef SHC_SLOTS
#define SHC_SLOTS 1
#endif
// MODE: 0 walking only, 1 = FULL, 2 = EXT = FULL + the extended paths, each behind its own run-time flag: the tip-orientation
// path of gravity_aligned_tips (updateTipAlignPose on legs of at most three joints; updateTipRotation and the rotation
// branch of Leg::applyIK beyond) and rough-terrain mode (default-tip updates, touchdown detection, target shifting).
template <class P, int D, int MODE> struct Cycle {
  static constexpr bool FULL = MODE >= 1;
  static constexpr bool EXT = MODE == 2;
  static constexpr int kSlots = SHC_SLOTS;
  using S = typename P::S;
  using T = typename P::T;
  using K = typename P::K;
  using LS = LegS<D>;

  static SHC_CYCLE_FN __forceinline__ V3<K> ld3K(const S* sp, int plane) {
    return {K(sp[(plane) * 32]), K(sp[(plane + 1) * 32]), K(sp[(plane + 2) * 32])};
  }
  static SHC_CYCLE_FN __forceinline__ V3<T> ld3T(const S* sp, int plane) {
    return {T(sp[(plane) * 32]), T(sp[(plane + 1) * 32]), T(sp[(plane + 2) * 32])};
  }
  template <class R> static SHC_CYCLE_FN __forceinline__ void st3(S* sp, int plane, V3<R> v) {
    sp[(plane) * 32] = S(v.x);
    sp[(plane + 1) * 32] = S(v.y);
    sp[(plane + 2) * 32] = S(v.z);
  }
  static SHC_CYCLE_FN __forceinline__ PoseT<K> ldPose(const S* sp, int plane) {
    PoseT<K> p;
    p.p = ld3K(sp, plane);
    p.q = {K(sp[(plane + 3) * 32]), K(sp[(plane + 4) * 32]), K(sp[(plane + 5) * 32]),
           K(sp[(plane + 6) * 32])};
    return p;
  }
  static SHC_CYCLE_FN __forceinline__ void stPose(S* sp, int plane, PoseT<K> p) {
    st3(sp, plane, p.p);
    sp[(plane + 3) * 32] = S(p.q.w);
    sp[(plane + 4) * 32] = S(p.q.x);
    sp[(plane + 5) * 32] = S(p.q.y);
    sp[(plane + 6) * 32] = S(p.q.z);
  }

  // PoseController::updateManualPose (pose_controller.cpp:863-1003); default_pose_ is the identity (no manually
  // manipulated legs in the batched engine, so calculateDefaultPose never moves it).
  static SHC_CYCLE_FN __forceinline__ PoseT<K> manual_pose_update(const RealConsts<K>& ck, PoseT<K> man, const float* in6,
                                                                int reset_mode) {
    if (reset_mode == 5) return pose_identity<K>();  // IMMEDIATE_ALL_RESET
    V3<K> cur_rot = quat_to_euler_nc(man.q, true);
    K tin[3] = {K(0), K(0), K(0)}, rin[3] = {K(0), K(0), K(0)};
    if (in6) {
      tin[0] = K(in6[0]); tin[1] = K(in6[1]); tin[2] = K(in6[2]);
      rin[0] = K(in6[3]); rin[1] = K(in6[4]); rin[2] = K(in6[5]);
    }
    bool rt[3] = {false, false, false}, rr[3] = {false, false, false};
    if (reset_mode == 1) { rt[2] = true; rr[2] = true; }
    else if (reset_mode == 2) { rt[0] = rt[1] = true; }
    else if (reset_mode == 3) { rr[0] = rr[1] = true; }
    else if (reset_mode == 4) { rt[0] = rt[1] = rt[2] = true; rr[0] = rr[1] = rr[2] = true; }
    K cp[3] = {man.p.x, man.p.y, man.p.z};
    K cr[3] = {cur_rot.x, cur_rot.y, cur_rot.z};
    K dpv[3], drv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (rt[i]) {
        if (cp[i] < K(0)) tin[i] = K(1);
        else if (cp[i] > K(0)) tin[i] = K(-1);
      }
      if (rr[i]) {
        if (cr[i] < K(0)) rin[i] = K(1);
        else if (cr[i] > K(0)) rin[i] = K(-1);
      }
      K tv = tin[i] * ck.max_translation_velocity;
      K rv = rin[i] * ck.max_rotation_velocity;
      K dpos = cp[i] + tv * ck.dt;
      K drot = cr[i] + rv * ck.dt;
      K tlim = sign_(tv) * ck.max_translation[i];
      if (rt[i] && K(0) < ck.max_translation[i] && K(0) > -ck.max_translation[i]) tlim = K(0);
      bool ptv = sign_(tv) > K(0);
      if ((ptv && dpos > tlim) || (!ptv && dpos < tlim)) tv = (tlim - cp[i]) / ck.dt;
      K rlim = sign_(rv) * ck.max_rotation[i];
      if (rr[i] && K(0) < ck.max_rotation[i] && K(0) > -ck.max_rotation[i]) rlim = K(0);
      bool prv = sign_(rv) > K(0);
      if ((prv && drot > rlim) || (!prv && drot < rlim)) rv = (rlim - cr[i]) / ck.dt;
      dpv[i] = cp[i] + tv * ck.dt;
      drv[i] = cr[i] + rv * ck.dt;
    }
    PoseT<K> out;
    out.p = {dpv[0], dpv[1], dpv[2]};
    out.q = correct_rotation(euler_to_quat(V3<K>{drv[0], drv[1], drv[2]}, true), qidentity<K>());
    return out;
  }

  // Bytes of one staging slot: the staged storage planes of a leg, its double planes and its int planes.
  static __host__ __device__ __forceinline__ int slot_s_bytes(int frontS) { return (LS::STAGED + frontS) * 32 * (int)sizeof(S); }
  static __host__ __device__ __forceinline__ int slot_bytes(int frontS) {
    return slot_s_bytes(frontS) + LD_COUNT * 32 * 8 + LI_COUNT * 32 * 4;
  }
  // Dynamic shared memory of one warp: [kSlots staging slots][joint tile: 32 robots x L*D floats][2 mbarriers][scratch],
  // 128-B granular.  kSlots = 1: leg l+1's planes are requested as soon as leg l has copied its own out of the slot (a leg
  // body is several DRAM latencies long, so one slot already hides HBM); the warp then needs ~9 KB and the SM holds as
  // many warps as the register budget allows.  kSlots = 2 requests two legs ahead.
  // bytes of the per-warp scratch planes behind the barriers: the body velocity (3 x T) of the tile's robots, written by
  // the robot-level stage and read by every leg (keeps six registers out of the leg loop without an L2 round trip per leg)
  static __host__ __device__ __forceinline__ int scratch_bytes() { return 3 * 32 * (int)sizeof(T); }
  static __host__ __device__ __forceinline__ int smem_per_warp(int frontS, int L) {
    return (kSlots * slot_bytes(frontS) + 32 * L * D * 4 + 16 + scratch_bytes() + 127) / 128 * 128;
  }

  // One warp = one tile of 32 robots (lane = robot).  `wsm` is the warp's shared memory (see smem_per_warp): the joint
  // commands are staged in its joint tile and written out by the whole warp as coalesced 128-byte lines by the caller.
  // Lanes past n_robots run on the (initialised) padding robots of the last tile so that the warp-collective staging
  // stays convergent; they read the inputs of the last real robot and their outputs are dropped by the caller.
  static SHC_CYCLE_FN void run(const Consts& c, Planes<S> pl, int tile_idx, int lane, const StepIO& io, unsigned char* __restrict__ wsm) {
    const IntConsts& ci = c.i;
    const RealConsts<T>& ct = ConstSel<T>::get(c);
    const RealConsts<K>& ck = ConstSel<K>::get(c);
    const RealConsts<double>& cd = c.d;
    const int L = ci.L;
    const bool f_auto = FULL && ci.auto_posing, f_incl = FULL && ci.inclination_posing, f_imu = FULL && ci.imu_posing;
    const bool f_adm = FULL && ci.admittance_control, f_effort = FULL && ci.use_joint_effort;
    const bool f_tipalign = EXT && D <= 3 && ci.tip_mode == TIP_ALIGN_POSE, f_tiprot = EXT && D > 3 && ci.tip_mode == TIP_ROTATION;
    const bool f_rough = EXT && ci.rough_terrain;
    const int front = FULL ? ci.frontS_leg : 0;
    // tile-major planes [tile][plane][32 lanes]: every field of this robot is at a compile-time offset from these bases
    const size_t tile = (size_t)tile_idx;
    const int r_real = tile_idx * 32 + lane;
    const bool live = r_real < ci.n_robots;
    const int r = live ? r_real : ci.n_robots - 1;  // index for the per-robot INPUT arrays only
    S* __restrict__ sp = pl.s + tile * (size_t)(ci.nS * 32) + lane;
    double* __restrict__ dp = pl.d + tile * (size_t)(ci.nD * 32) + lane;
    int* __restrict__ ip = pl.i + tile * (size_t)(ci.nI * 32) + lane;

    SHC_STAMP(0);
    // ---- staging ring ------------------------------------------------------------------------------------------------
    const int sS_bytes = slot_s_bytes(front), s_bytes = slot_bytes(front);
    float* __restrict__ stage = reinterpret_cast<float*>(wsm + kSlots * s_bytes) + lane * (L * D);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + kSlots * s_bytes + 32 * L * D * 4);
    auto bar_of = [&](int s) { return bars + s; };
    T* __restrict__ scratch = reinterpret_cast<T*>(wsm + kSlots * s_bytes + 32 * L * D * 4 + 16) + lane;
    const S* tileS = pl.s + tile * (size_t)(ci.nS * 32);
    const double* tileD = pl.d + tile * (size_t)(ci.nD * 32);
    const int* tileI = pl.i + tile * (size_t)(ci.nI * 32);
    auto issue_leg = [&](int l) {  // lane 0: ask the TMA unit for leg l's every-cycle planes
      unsigned char* slot = wsm + (l % kSlots) * s_bytes;
      uint64_t* bar = bar_of(l % kSlots);
      mbar_expect_tx(bar, (unsigned)s_bytes);
      bulk_g2s(slot, tileS + (ci.offS_leg + l * ci.strideS_leg - front) * 32, (unsigned)sS_bytes, bar);
      bulk_g2s(slot + sS_bytes, tileD + (ci.offD_leg + l * ci.strideD_leg) * 32, LD_COUNT * 32 * 8, bar);
      bulk_g2s(slot + sS_bytes + LD_COUNT * 32 * 8, tileI + (ci.offI_leg + l * ci.strideI_leg) * 32, LI_COUNT * 32 * 4, bar);
    };
    if (SHC_LANE0(lane)) {
      mbar_init(bar_of(0), 1);
      mbar_init(bar_of(1), 1);
      fence_mbar_init();
    }
    SHC_SYNCWARP();

    // Measured tip forces ([N][L][3] floats, this robot's L*3 values contiguous): copied asynchronously (cp.async, no
    // registers in flight) into the lane's own stretch of the joint tile, BEHIND the joints: leg l's force sits at float
    // L*(D-3) + 3l of the lane's L*D floats, which the joint commands of the legs 0..l (written at the end of each leg) never
    // reach before leg l has read it.  The robot-level stage below hides the latency.
    const bool stage_forces = FULL && f_adm && !f_effort && io.tip_force != nullptr;
    if (stage_forces) {
      const float* fsrc = io.tip_force + (size_t)r * (L * 3);
      float* fdst = stage + L * (D - 3);
      const bool wide = ((L * 12) & 15) == 0 && ((L * D * 4) & 15) == 0 && ((L * (D - 3) * 4) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(io.tip_force) & 15) == 0;
      if (wide) {
        for (int k = 0; k < (L * 3) / 4; ++k) cp_async_16(fdst + 4 * k, fsrc + 4 * k);
      } else {
        for (int k = 0; k < L * 3; ++k) cp_async_4(fdst + k, fsrc + k);
      }
    }

    // ---- every unconditional robot-level load, issued back to back (one exposed HBM latency for the whole stage) ------
    int rbits = ip[(RI_BITS) * 32];
    int progs[kMaxLegs];
    double tips_x[kMaxLegs], tips_y[kMaxLegs];
#pragma unroll
    for (int l = 0; l < kMaxLegs; ++l) {
      progs[l] = -1;
      tips_x[l] = tips_y[l] = 0.0;
      if (l < L) {
        progs[l] = ip[(ci.offI_leg + l * ci.strideI_leg + LI_PROG) * 32];
        tips_x[l] = dp[(ci.offD_leg + l * ci.strideD_leg + LD_TIP) * 32];
        tips_y[l] = dp[(ci.offD_leg + l * ci.strideD_leg + LD_TIP + 1) * 32];
      }
    }
    PoseT<K> owpp = ldPose(sp, RS_OWPP);
    T dvx = T(sp[(RS_VEL) * 32]), dvy = T(sp[(RS_VEL + 1) * 32]), dw = T(sp[(RS_ANGVEL) * 32]);
    const K walker_wp_z = K(sp[(RS_WPL + 2) * 32]);
    const V3<K> walker_wpn_k = ld3K(sp, RS_WPN);
    // the manual pose is the identity on almost every robot (state bit): only the others pay for its seven planes
    PoseT<K> man = pose_identity<K>();
    bool man_identity = true;
    if (ci.manual_posing && !((rbits >> RB_MANUAL_IDENTITY) & 1)) {
      man = ldPose(sp, RS_MAN);
      man_identity = man.p.x == K(0) && man.p.y == K(0) && man.p.z == K(0) && man.q.w == K(1) && man.q.x == K(0) &&
                     man.q.y == K(0) && man.q.z == K(0);
    }

    // ---- inputs: bodyVelocityInputCallback (state_controller.cpp:1127-1136) --------------------------------------
    double vin_x = (double)io.cmd[3 * (size_t)r + 0] * cd.body_velocity_scaler;
    double vin_y = (double)io.cmd[3 * (size_t)r + 1] * cd.body_velocity_scaler;
    double win = (double)io.cmd[3 * (size_t)r + 2] * cd.body_velocity_scaler;
    double in_norm = sqrt(vin_x * vin_x + vin_y * vin_y);
    if (ci.velocity_input_mode == SHC_VELOCITY_THROTTLE && in_norm > 1.0) {
      double s = fmin(1.0, in_norm);
      vin_x = s * (vin_x / in_norm);
      vin_y = s * (vin_y / in_norm);
      in_norm = sqrt(vin_x * vin_x + vin_y * vin_y);
    }

    Q4<K> imu_raw{K(0), K(0), K(0), K(0)};  // Model::imu_data_.orientation (UNDEFINED until set)
    V3<K> gyro{K(0), K(0), K(0)};
    if (FULL && io.imu) {
      const float* m = io.imu + 10 * (size_t)r;
      imu_raw = qnormalized(Q4<K>{K(m[0]), K(m[1]), K(m[2]), K(m[3])});  // Model::setImuData (model.h:146)
      gyro = {K(m[4]), K(m[5]), K(m[6])};
    }
    // Model::getImuData (model.h:132): undefined orientation reads as identity
    Q4<K> imu_q = (imu_raw.w == K(0) && imu_raw.x == K(0) && imu_raw.y == K(0) && imu_raw.z == K(0)) ? qidentity<K>() : imu_raw;

    // The first two legs' transfers start once the robot-level loads are back (rbits is the first of them; the
    // comparison only creates the dependency): the opening HBM burst of a wave of warps is then the 8 KB the first stage
    // needs, and the legs' planes stream in behind it while that stage computes (-3 % per launch, measured).
    if (SHC_LANE0(lane) && rbits != 0x7fffffff) {
      issue_leg(0);
      if (kSlots > 1 && L > 1) issue_leg(1);
    }
    SHC_STAMP(2);
    int walk_state = rbits & 3;
    // AdmittanceController::updateStiffness runs at the top of loop() when the walker is not STOPPED (state_controller.cpp:175)
    const bool f_stiff = FULL && ci.dynamic_stiffness && walk_state != WALK_STOPPED;
    int legs_at_correct = (rbits >> 2) & 15;
    int legs_completed = (rbits >> 6) & 15;
    int rtd = (rbits >> 10) & 1;
    int auto_state = (rbits >> 13) & 3;
    int status = 0;

    // =================================================================================================================
    // 1. PoseController::updateCurrentPose
    // =================================================================================================================
    // updateWalkPlanePose (pose_controller.cpp:1092): the last leg (id order) whose scaled swing progress is in [0,1]
    double c_in = 0.0;
    int ref_leg = -1;
    // scaled progress = (n / swing_period) * scaler is monotonic in the integer numerator n, so "0 <= progress <= 1" is
    // the integer window 0 <= n <= swing_ref_max (found on the host with the same two double operations)
    int ref_num = 0;
#pragma unroll
    for (int l = 0; l < kMaxLegs; ++l) {
      int swing_num = (int)(short)(progs[l] & 0xffff);  // -1 for the legs past L
      if (swing_num >= 0 && swing_num <= ci.swing_ref_max) {
        ref_num = swing_num;
        ref_leg = l;
      }
    }
    if (ref_leg >= 0) c_in = smooth_step(((double)ref_num / (double)ci.swing_period) * cd.swing_progress_scaler);
    K wp_z = K(0);
    V3<K> wpn_ref{K(0), K(0), K(1)};
    if (ref_leg >= 0) {
      // A leg with a valid swing progress ran its tip update in the previous cycle and saved the walker's plane of that
      // cycle's start (walk_controller.cpp:1049).  Unless the previous cycle's updateWalkPlane changed the plane
      // (RB_PLANE_CHANGED, also set for states written from outside), that is the walker's plane held now: no
      // dependent HBM round trip for the leg's saved copy.
      if ((rbits >> RB_PLANE_CHANGED) & 1) {
        int base = ci.offS_leg + ref_leg * ci.strideS_leg;
        wp_z = K(sp[(base + LS::WP + 2) * 32]);
        wpn_ref = ld3K(sp, base + LS::WPN);
      } else {
        wp_z = walker_wp_z;
        wpn_ref = walker_wpn_k;
      }
    }
    PoseT<K> wpp = owpp;
    if (c_in != 0.0) {
      PoseT<K> new_wpp;
      if (wpn_ref.x == K(0) && wpn_ref.y == K(0) && wpn_ref.z == K(1)) {
        // FromTwoVectors(z, z) is exactly the identity and rotating by it is exact: same bits as the general path
        new_wpp.q = qidentity<K>();
        new_wpp.p = V3<K>{K(0), K(0), ck.body_clearance};
      } else {
        new_wpp.q = correct_rotation(from_two_vectors(V3<K>{K(0), K(0), K(1)}, wpn_ref), qidentity<K>());
        new_wpp.p = qrot(new_wpp.q, V3<K>{K(0), K(0), ck.body_clearance});
      }
      new_wpp.p.z += wp_z;
      wpp = pose_interpolate(owpp, K(c_in), new_wpp);
      stPose(sp, RS_WPP, wpp);
      if (c_in == 1.0) stPose(sp, RS_OWPP, wpp);
    } else {
      // c = 0: interpolate(origin, 0, new) returns the origin pose bit for bit (scale0 = 1, scale1 = 0)
      stPose(sp, RS_WPP, wpp);
    }

    PoseT<K> cur_pose = pose_add(pose_identity<K>(), wpp);
    if (ci.manual_posing) {
      if (!(man_identity && io.manual == nullptr && io.pose_reset_mode == 0)) {
        // (with an identity pose, no input and no reset the update returns the identity again, exactly)
        man = manual_pose_update(ck, man, io.manual ? io.manual + 6 * (size_t)r : nullptr, io.pose_reset_mode);
        stPose(sp, RS_MAN, man);
        man_identity = man.p.x == K(0) && man.p.y == K(0) && man.p.z == K(0) && man.q.w == K(1) && man.q.x == K(0) &&
                       man.q.y == K(0) && man.q.z == K(0);
        cur_pose = pose_add(cur_pose, man);
      }
    }
    PoseT<K> auto_pose = pose_identity<K>();
    if (f_auto) auto_pose = ldPose(sp, ci.offS_auto + AUTO_POSE);
    // With an identity manual pose and no auto pose (the usual case) the rotation the inclination stage removes and the
    // target of the IMU stage are both the identity: "IMU rotation with the pose removed" and "rotation error" are then the
    // same quaternion up to sign, their rotation matrices are identical bit for bit (every entry is a product of two
    // components), and one Euler decomposition serves both stages.
    const bool plain_rot = FULL && man_identity &&
                           (!f_auto || (auto_pose.q.w == K(1) && auto_pose.q.x == K(0) && auto_pose.q.y == K(0) && auto_pose.q.z == K(0)));
    V3<K> e_imu{K(0), K(0), K(0)};
    if (FULL && plain_rot && (f_incl || f_imu)) e_imu = quat_to_euler_nc(qnormalized(imu_q), false);
    if (f_incl) {  // updateInclinationPose (:1240)
      V3<K> e = e_imu;
      if (!plain_rot) {
        Q4<K> comb = qnormalized(qmul(man.q, auto_pose.q));
        Q4<K> removed = qnormalized(qmul(imu_q, qinverse(comb)));
        e = quat_to_euler_nc(removed, false);
      }
      K lon = -ck.body_clearance * tan_(e.y);
      K lat = ck.body_clearance * tan_(e.x);
      lon = clamp_(lon, -ck.max_translation[0], ck.max_translation[0]);
      lat = clamp_(lat, -ck.max_translation[1], ck.max_translation[1]);
      sp[(ci.offS_imu + IMU_INCL) * 32] = S(lon);
      sp[(ci.offS_imu + IMU_INCL + 1) * 32] = S(lat);
      PoseT<K> incl = pose_identity<K>();
      incl.p = {lon, lat, K(0)};
      cur_pose = pose_add(cur_pose, incl);
    }
    int master_phase = 0;
    bool run_auto = false;
    if (f_imu) {  // updateIMUPose (:1191); robot_state is RUNNING in every engine cycle
      Q4<K> target_rotation = correct_rotation(man.q, qidentity<K>());
      V3<K> pe = e_imu;
      if (!plain_rot) {
        Q4<K> current_rotation = correct_rotation(imu_q, qidentity<K>());
        Q4<K> rot_err = qnormalized(qmul(current_rotation, qinverse(target_rotation)));
        pe = quat_to_euler_nc(rot_err, false);
      }
      pe.z = K(0);
      const int b = ci.offS_imu;
      st3(sp, b + IMU_POS, pe);  // rotation_position_error_ (:1210)
      Q4<K> imu_pose_q;
      // IMU_POSING_DEADBAND is 0.0: "norm < 0" never holds, the PID always runs (pose_controller.h:25)
      V3<K> abs_err = ld3K(sp, b + IMU_ABS) + pe * ck.dt;
      V3<K> vel_err = (-gyro) * K(0.15) + ld3K(sp, b + IMU_VEL) * (K(1) - K(0.15));
      st3(sp, b + IMU_ABS, abs_err);
      st3(sp, b + IMU_VEL, vel_err);
      V3<K> corr = -(vel_err * ck.pid_d + pe * ck.pid_p + abs_err * ck.pid_i);
      corr.x = clamp_(corr.x, -ck.max_rotation[0], ck.max_rotation[0]);
      corr.y = clamp_(corr.y, -ck.max_rotation[1], ck.max_rotation[1]);
      corr.z = man_identity ? K(0) : quat_to_euler_nc(target_rotation, false).z;  // yaw of the identity is 0
      if (norm(corr) > K(100)) status |= 8;
      imu_pose_q = correct_rotation(euler_to_quat(corr, false), target_rotation);
      sp[(b + IMU_Q) * 32] = S(imu_pose_q.w);
      sp[(b + IMU_Q + 1) * 32] = S(imu_pose_q.x);
      sp[(b + IMU_Q + 2) * 32] = S(imu_pose_q.y);
      sp[(b + IMU_Q + 3) * 32] = S(imu_pose_q.z);
      PoseT<K> imu_pose = pose_identity<K>();
      imu_pose.q = imu_pose_q;
      cur_pose = pose_add(cur_pose, imu_pose);
    } else if (f_auto) {  // updateAutoPose (:1134)
      run_auto = true;
      const int rb = ci.offI_leg + ci.auto_ref_leg * ci.strideI_leg;
      int ref_bits = ip[(rb + LI_BITS) * 32];
      // zero_body_velocity of the reference leg: stride_vector_.norm() == 0
      V3<T> ref_stride = ld3T(sp, ci.offS_leg + ci.auto_ref_leg * ci.strideS_leg + LS::STRIDE);
      bool zero_body_velocity = (ref_stride.x * ref_stride.x + ref_stride.y * ref_stride.y + ref_stride.z * ref_stride.z) == T(0);
      if (walk_state == WALK_STARTING || walk_state == WALK_MOVING) auto_state = POSE_POSING;
      else if ((zero_body_velocity && walk_state == WALK_STOPPING) || walk_state == WALK_STOPPED) auto_state = POSE_STOP_POSING;
      int pose_phase = ip[(ci.offI_auto + AI_PHASE) * 32];
      if (ci.pose_sync) {
        master_phase = ref_bits & 0xffff;
      } else {
        master_phase = pose_phase;
        pose_phase = (pose_phase + 1) % ci.pose_phase_length;
        ip[(ci.offI_auto + AI_PHASE) * 32] = pose_phase;
      }
      int pflags = ip[(ci.offI_auto + AI_FLAGS) * 32];
      auto_pose = auto_posers_update<K>(ci, ck, master_phase, imu_raw, auto_state, pflags);
      ip[(ci.offI_auto + AI_FLAGS) * 32] = pflags;
      stPose(sp, ci.offS_auto + AUTO_POSE, auto_pose);
      cur_pose = pose_add(cur_pose, auto_pose);
    }
    if (f_tipalign) {
      // updateTipAlignPose (pose_controller.cpp:1024, EXPERIMENTAL in the reference): every leg with a swing progress, in leg
      // order, moves the one tip_align_pose_ - a body translation that brings the leg's last joint over its tip along the
      // walk-plane normal during the second half of the swing and back to zero during the first half of the next one.
      PoseT<K> tap = ldPose(sp, ci.offS_tip + TA_POSE);
      PoseT<K> otap = ldPose(sp, ci.offS_tip + TA_ORIGIN);
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const int n_swing = (int)(short)(ip[(ci.offI_leg + l * ci.strideI_leg + LI_PROG) * 32] & 0xffff);
        if (n_swing < 0) continue;  // swing_progress == -1
        const K swing_progress = K(n_swing) / K(ci.swing_period);
        const S* __restrict__ sl = sp + (ci.offS_leg + l * ci.strideS_leg) * 32;
        const LegConsts<K>& lk = ck.leg[l];
        const V3<K> wpn = ld3K(sl, LS::WPN);
        const Q4<K> wp_rot = from_two_vectors(V3<K>{K(0), K(0), K(1)}, wpn);
        K ql[D];
#pragma unroll
        for (int j = 0; j < D; ++j) ql[j] = K(sl[(LS::Q + j) * 32]);
        Chain<K, D> chl;
        leg_chain<K, D>(lk, ql, chl);  // the model as the previous cycle's applyFK left it
        const V3<K> tip_to_joint = t1_rotate(lk, chl.p[D - 1] - chl.tip);
        const K link_length = norm(tip_to_joint);
        V3<K> a = qrot(wp_rot, tip_to_joint);
        V3<K> b = wpn * link_length;
        const V3<K> to_alignment = -(a - b * (dot(a, b) / dot(b, b)));
        a = tap.p;
        b = wpn;
        const V3<K> aligned_now = a - b * (dot(a, b) / dot(b, b));
        V3<K> target = aligned_now + to_alignment;
        // clamped(vector, limit) of standard_includes.h:134 bounds every axis above by limit[1]: kept
        target.x = max_(-ck.max_translation[0], min_(target.x, ck.max_translation[1]));
        target.y = max_(-ck.max_translation[1], min_(target.y, ck.max_translation[1]));
        target.z = max_(-ck.max_translation[2], min_(target.z, ck.max_translation[1]));
        K c = smooth_step(swing_progress);
        if (swing_progress < K(0.5)) {
          c = smooth_step(c * K(2));
          tap = pose_interpolate(otap, c, pose_identity<K>());
        } else {
          c = smooth_step((c - K(0.5)) * K(2));
          tap = pose_interpolate(pose_identity<K>(), c, PoseT<K>{target, qidentity<K>()});
        }
        if (swing_progress == K(1)) otap = tap;
      }
      stPose(sp, ci.offS_tip + TA_POSE, tap);
      stPose(sp, ci.offS_tip + TA_ORIGIN, otap);
      cur_pose = pose_add(cur_pose, tap);
    }
    const int pose_state = auto_state;  // walker_->setPoseState(poser_->getAutoPoseState())

    // =================================================================================================================
    // 2. WalkController::updateWalk — robot-level part (walk_controller.cpp:440-564)
    // =================================================================================================================
    // odometry_ideal_ (updated right after the body velocity below: it depends on nothing else): the loads fly while the
    // limits and the velocity are worked out
    const Q4<T> odom_q{T(sp[(RS_ODOMQ) * 32]), T(sp[(RS_ODOMQ + 1) * 32]), T(sp[(RS_ODOMQ + 2) * 32]), T(sp[(RS_ODOMQ + 3) * 32])};
    const double odom_x = dp[(RD_ODOMP) * 32], odom_y = dp[(RD_ODOMP + 1) * 32], odom_z = dp[(RD_ODOMP + 2) * 32];
    double lim[4] = {2147483647.0, 2147483647.0, 2147483647.0, 2147483647.0};
#pragma unroll
    for (int l = 0; l < kMaxLegs; ++l) {  // getLimit (:414): the four calls share the bearing of each leg
      if (l >= L) break;
      const double tx = tips_x[l], ty = tips_y[l];
      double sx = vin_x + win * (-ty), sy = vin_y + win * tx;
      // bucket = mod(roundToInt(deg(atan2(sy, sx))), 360) / 45 (int / int floors to the 45-degree bucket, trap 1),
      // decided without atan2: the rounding puts the bucket edges at 44.5, 89.5, 134.5, 179.5 degrees on the upper half
      // plane (closed below: theta >= edge) and, mirrored through the origin, at -135.5, -90.5, -45.5, -0.5 on the lower
      // one (open: theta > edge).  theta >= phi <=> sin(theta - phi) >= 0, and a point of the lower half plane has passed
      // the mirrored edge exactly when its own sine against the upper edge is negative: one set of four tests serves
      // both halves, with no divergence between lanes.
      int passed_hi = 0, passed_lo = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double t = sy * cd.sec_cos[k] - sx * cd.sec_sin[k];
        passed_hi += t >= 0.0 ? 1 : 0;
        passed_lo += t < 0.0 ? 1 : 0;
      }
      int bucket = sy >= 0.0 ? passed_hi : ((4 + passed_lo) & 7);
      if (sx == 0.0 && sy == 0.0) bucket = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) lim[k] = fmin(lim[k], cd.limits[k][bucket]);
    }
    T max_lin_speed = T(lim[0]), max_ang_speed = T(lim[1]), max_lin_acc = T(lim[2]), max_ang_acc = T(lim[3]);

    T nvx, nvy, nw;
    if (walk_state != WALK_STOPPING) {
      if (ci.velocity_input_mode == SHC_VELOCITY_THROTTLE) {
        T cx = T(vin_x), cy = T(vin_y);
        if (in_norm > 1.0) {
          cx = T(vin_x * (1.0 / in_norm));
          cy = T(vin_y * (1.0 / in_norm));
        }
        T k = T(1) - T(fabs(win));
        nvx = cx * max_lin_speed * k;
        nvy = cy * max_lin_speed * k;
        nw = T(fmax(-1.0, fmin(win, 1.0))) * max_ang_speed;
      } else {
        T cx = T(vin_x), cy = T(vin_y);
        if (in_norm > (double)max_lin_speed) {
          cx = T(vin_x * ((double)max_lin_speed / in_norm));
          cy = T(vin_y * ((double)max_lin_speed / in_norm));
        }
        nw = clamp_(T(win), -max_ang_speed, max_ang_speed);
        T k = max_ang_speed != T(0) ? (T(1) - abs_(nw / max_ang_speed)) : T(0);
        nvx = cx * k;
        nvy = cy * k;
      }
    } else {
      nvx = nvy = nw = T(0);
    }
    const bool has_cmd = (in_norm != 0.0) || (win != 0.0);

    {
      T ax = nvx - dvx, ay = nvy - dvy;
      T an = sqrt_(ax * ax + ay * ay);
      if (an < max_lin_acc * ct.dt) {
        dvx += ax;
        dvy += ay;
      } else {
        T n2 = ax * ax + ay * ay;
        T ux = ax, uy = ay;
        if (n2 > T(0)) {
          T inv = T(1) / sqrt_(n2);
          ux *= inv;
          uy *= inv;
        }
        dvx += ux * max_lin_acc * ct.dt;
        dvy += uy * max_lin_acc * ct.dt;
      }
      T aa = nw - dw;
      if (abs_(aa) < max_ang_acc * ct.dt) dw += aa;
      else dw += sign_(aa) * max_ang_acc * ct.dt;
    }
    sp[(RS_VEL) * 32] = S(dvx);
    sp[(RS_VEL + 1) * 32] = S(dvy);
    sp[(RS_ANGVEL) * 32] = S(dw);
    // the values the legs (updateStride) and the odometry read back: the stored ones
    dvx = T(S(dvx)); dvy = T(S(dvy)); dw = T(S(dw));
    scratch[0] = dvx;
    scratch[32] = dvy;
    scratch[64] = dw;

    bool starting_now = false;  // STOPPED -> STARTING returns before any tip update (trap 7)
    if (walk_state == WALK_STOPPED && has_cmd) {
      walk_state = WALK_STARTING;
      starting_now = true;
    } else if (walk_state == WALK_STARTING && legs_at_correct == L && legs_completed == L) {
      legs_at_correct = 0;
      legs_completed = 0;
      walk_state = WALK_MOVING;
    } else if (walk_state == WALK_MOVING && !has_cmd) {
      walk_state = WALK_STOPPING;
    } else if (walk_state == WALK_STOPPING && legs_at_correct == L && pose_state == POSE_COMPLETE) {
      legs_at_correct = 0;
      walk_state = WALK_STOPPED;
    }

    // calculateOdometry (:783): odometry_ideal_ = odometry_ideal_.addPose(velocity * dt); it depends on the body velocity
    // only, so it is done here, while its planes (loaded with the other robot-level planes) are still at hand
    if (!starting_now) {
      V3<T> dpos = qrot(odom_q, V3<T>{dvx * ct.dt, dvy * ct.dt, T(0)});
      dp[(RD_ODOMP) * 32] = odom_x + (double)dpos.x;
      dp[(RD_ODOMP + 1) * 32] = odom_y + (double)dpos.y;
      dp[(RD_ODOMP + 2) * 32] = odom_z + (double)dpos.z;
      Q4<T> nq = qmul(odom_q, q_axis_z(dw * ct.dt));
      sp[(RS_ODOMQ) * 32] = S(nq.w);
      sp[(RS_ODOMQ + 1) * 32] = S(nq.x);
      sp[(RS_ODOMQ + 2) * 32] = S(nq.y);
      sp[(RS_ODOMQ + 3) * 32] = S(nq.z);
    }
    // flat walk plane (the only one a batch without rough-terrain inputs ever holds): the swing clearance is then
    // (0, 0, swing_height) and the legs need not read the plane back
    const bool plane_flat = walker_wpn_k.x == K(0) && walker_wpn_k.y == K(0) && walker_wpn_k.z == K(1);
    const bool plane_changed_prev = (rbits >> RB_PLANE_CHANGED) & 1;
    bool def_changed = false;


    // =================================================================================================================
    // 3. per leg: walk state machine + LegStepper + updateStance + Leg::applyIK
    // =================================================================================================================
    SHC_STAMP(3);
    if (stage_forces) cp_async_wait_all();
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      // per-leg plane bases in HBM (stores, rare loads): every field of this leg is at an immediate offset from them
      S* __restrict__ sl = sp + (ci.offS_leg + l * ci.strideS_leg) * 32;
      double* __restrict__ dl = dp + (ci.offD_leg + l * ci.strideD_leg) * 32;
      int* __restrict__ il = ip + (ci.offI_leg + l * ci.strideI_leg) * 32;
      const LegConsts<K>& lk = ck.leg[l];
      const LegConsts<T>& lt = ct.leg[l];
      // the staged copy of this leg's every-cycle planes: wait for the TMA transfer issued two legs ago
      const unsigned char* slot = wsm + (l % kSlots) * s_bytes;
      SHC_STAMP(4 + 3 * l);
      mbar_wait(bar_of(l % kSlots), (unsigned)((l / kSlots) & 1));
      SHC_STAMP(5 + 3 * l);
      const S* __restrict__ ss = reinterpret_cast<const S*>(slot) + front * 32 + lane;
      const double* __restrict__ sd = reinterpret_cast<const double*>(slot + sS_bytes) + lane;
      const int* __restrict__ si = reinterpret_cast<const int*>(slot + sS_bytes + LD_COUNT * 32 * 8) + lane;
      int bits = si[(LI_BITS) * 32];
      int prog = si[(LI_PROG) * 32];
      double tipx = sd[(LD_TIP) * 32];
      double tipy = sd[(LD_TIP + 1) * 32];
      double tipz = sd[(LD_TIP + 2) * 32];
      V3<T> def = ld3T(ss, LS::DEF);
      V3<T> stride = ld3T(ss, LS::STRIDE);
      V3<T> tgt;
      K q[D], qd[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        q[j] = K(ss[(LS::Q + j) * 32]);
        qd[j] = K(ss[(LS::QD + j) * 32]);
      }
      K adm_x0 = K(0), adm_x1 = K(0);
      V3<K> adm_force{K(0), K(0), K(0)};
      if (f_adm || f_effort) {
        adm_x0 = K(ss[(LS::ADM_X) * 32]);
        adm_x1 = K(ss[(LS::ADM_X + 1) * 32]);
        if (f_effort) adm_force = ld3K(ss, LS::ADM_FORCE);
        else if (stage_forces) {
          const float* f = stage + L * (D - 3) + 3 * l;
          adm_force = {K(f[0]), K(f[1]), K(f[2])};
        }
      }
      // tip-rotation state (gravity_aligned_tips, D > 3) and the chain at the joint angles the cycle starts with
      Q4<K> tr_cur{K(0), K(0), K(0), K(0)}, tr_origin{K(0), K(0), K(0), K(0)};
      Chain<K, D> ch;
      bool step_plane_defined = false, touchdown_detection = false;
      bool ext_target = false, ext_odom = false, ext_default = false;  // external_target_ / external_default_ (rough terrain)
      V3<K> step_plane{K(0), K(0), K(0)}, model_tip{K(0), K(0), K(0)};
      if constexpr (EXT) {
        leg_chain<K, D>(lk, q, ch);
        if (f_tiprot) {
          const S* __restrict__ tp = sl + ci.tipS_leg * 32;
          tr_cur = {K(tp[(TR_CUR) * 32]), K(tp[(TR_CUR + 1) * 32]), K(tp[(TR_CUR + 2) * 32]), K(tp[(TR_CUR + 3) * 32])};
          tr_origin = {K(tp[(TR_ORIGIN) * 32]), K(tp[(TR_ORIGIN + 1) * 32]), K(tp[(TR_ORIGIN + 2) * 32]), K(tp[(TR_ORIGIN + 3) * 32])};
        }
        if (f_rough) {
          // Leg::current_tip_pose_.position_ (base_link frame) as the previous cycle's applyFK left it
          model_tip = t1_rotate(lk, ch.tip) + V3<K>{lk.t1p[0], lk.t1p[1], lk.t1p[2]};
          step_plane_defined = (bits >> LB_STEP_PLANE) & 1;
          touchdown_detection = (bits >> LB_TOUCHDOWN) & 1;
          ext_target = (bits >> LB_EXT_TARGET) & 1;
          ext_odom = (bits >> LB_EXT_ODOM) & 1;
          ext_default = (bits >> LB_EXT_DEFAULT) & 1;
          S* __restrict__ rp = sl + ci.roughS_leg * 32;
          if (step_plane_defined) step_plane = ld3K(rp, RT_STEP_PLANE);
          if (io.tip_force) {
            // tipStatesCallback with wrench values (state_controller.cpp:1636-1648), delivered before this loop(): touchdown
            // detection is on from now on, and Leg::touchdownDetection (model.cpp:712) defines / forgets the step plane
            touchdown_detection = true;
            const float* f = io.tip_force + ((size_t)r * L + l) * 3;
            const K fn = norm(V3<K>{K(f[0]), K(f[1]), K(f[2])});
            if (fn > ck.touchdown_threshold && !step_plane_defined) {
              step_plane = model_tip;
              step_plane_defined = true;
              st3(rp, RT_STEP_PLANE, step_plane);
            } else if (fn < ck.liftoff_threshold) {
              step_plane_defined = false;
            }
          }
          if (io.step_planes) {
            // range sensor readings (:1650-1675): the step plane sits `range` along the tip's x axis; its orientation is
            // never read (walk_controller.cpp:1085-1110 uses the position and whether the pose is defined)
            touchdown_detection = true;
            const float range = io.step_planes[((size_t)r * L + l) * 3 + 2];
            if (range < SHC_RANGE_UNASSIGNED) {
              step_plane = model_tip + t1_rotate(lk, ch.tipx) * K(range);
              step_plane_defined = true;
              st3(rp, RT_STEP_PLANE, step_plane);
            } else {
              step_plane_defined = false;
            }
          }
        }
      }
      // the one HBM read the swing branch may need (tip velocity at the first swing iteration), issued early
      const bool swing_begins = (bits & 0xffff) == ci.swing_start;
      V3<T> tipvel_prev{T(0), T(0), T(0)};
      if (swing_begins) tipvel_prev = ld3T(sl, LS::TIPVEL);
      int phase = bits & 0xffff;
      int step_state = (bits >> 16) & 3;
      bool at_correct = (bits >> 18) & 1;
      bool completed = (bits >> 19) & 1;
      bool negate = (bits >> 20) & 1;
      bool plane_saved = (bits >> LB_PLANE_SAVED) & 1;
      if (f_stiff) {
        // updateStiffness (admittance_controller.cpp:96) looks at every leg as the cycle found it: a swinging leg's own
        // stiffness drops and its neighbours' rise with |tip z - default z| / swing height.  The per-leg reference is
        // parked in the leg's stiffness plane and combined across legs behind the loop.
        const K z_diff = K(tipz) - K(def.z);
        sl[(LS::STIFF) * 32] = S(step_state == STEP_SWING ? abs_(z_diff / ck.swing_height) : K(-1));
      }
      int swing_num = (int)(short)(prog & 0xffff);
      int stance_num = (int)(short)((prog >> 16) & 0xffff);

      // LegPoser::updateAutoPose (pose_controller.cpp:1716) — uses the step state of the previous cycle
      PoseT<K> leg_auto = auto_pose;
      if (run_auto) {
        leg_auto = leg_auto_pose<K>(ci, lk.neg_ratio, l, master_phase, step_state, auto_pose, negate);
      } else if (f_auto) {
        leg_auto = pose_identity<K>();  // LegPoser::auto_pose_ is only refreshed by updateAutoPose; IMU posing keeps identity
      }

      if (starting_now) {
        // walk_controller.cpp:535-545
        at_correct = false;
        completed = false;
        step_state = STEP_STANCE;
        phase = ci.phase_offset[l];
        if (phase >= ci.swing_start && phase < ci.swing_end) step_state = STEP_SWING;
        else if (phase < ci.stance_end || phase >= ci.stance_start) step_state = STEP_STANCE;
      } else {
        // ---- walk state machine for this leg (walk_controller.cpp:573-632) ----
        if (walk_state == WALK_STARTING) {
          if (legs_at_correct == L) {
            if (phase == ci.swing_end && !completed) {
              completed = true;
              legs_completed++;
            }
          }
          if (!at_correct) {
            if (ci.phase_offset[l] > ci.swing_start && ci.phase_offset[l] < ci.swing_end && phase != ci.swing_end) {
              step_state = STEP_FORCE_STANCE;
            } else {
              legs_at_correct++;
              at_correct = true;
            }
          }
        } else if (walk_state == WALK_MOVING) {
          at_correct = false;
        } else if (walk_state == WALK_STOPPING) {
          bool zero_body_velocity = (stride.x * stride.x + stride.y * stride.y + stride.z * stride.z) == T(0);
          if (zero_body_velocity && !at_correct && phase == ci.swing_end) {
            V3<double> wpn_l = cvt<double>(ld3T(sl, LS::WPN));
            tgt = ld3T(sl, LS::TGT);  // last cycle's target (rare path: straight from HBM)
            V3<double> err{tipx - (double)tgt.x, tipy - (double)tgt.y, tipz - (double)tgt.z};
            err = rejection(err, wpn_l);
            bool at_target = norm(err) < 0.01;  // TIP_TOLERANCE (pose_controller.h:19)
            if (at_target || rtd) {
              rtd = 0;
              // LegStepper::updateDefaultTipPosition (:984)
              V3<K> idt{lk.identity_x, lk.identity_y + lk.span_dy, K(0)};
              idt = pose_transform(ldPose(sp, RS_WPP), idt);  // Model::default_pose_ = walk_plane_pose_ (pose_controller.cpp:819)
              V3<K> sto = ld3K(sl, LS::STO_P);
              V3<K> proj = projection(sto - idt, cvt<K>(wpn_l));
              def = cvt<T>(idt + proj);
              if constexpr (EXT) {
                if (ext_default) {  // an externally requested default, moved back by the robot's motion since the request (:988-992)
                  const S* __restrict__ rp = sl + ci.roughS_leg * 32;
                  def = cvt<T>(pose_transform(ldPose(rp, RT_ED_POSE), -ld3K(rp, RT_ED_TF)));
                }
              }
              st3(sl, LS::DEF, def);
              def_changed = true;
              step_state = STEP_FORCE_STOP;
              at_correct = true;
              legs_at_correct++;
            } else {
              rtd = 1;
            }
          }
        } else {  // STOPPED
          step_state = STEP_FORCE_STOP;
          phase = 0;
        }

        // LegStepper::updateDefaultTipPosition (:984) without an external default: the identity tip under the model's default
        // body pose (= this cycle's walk-plane pose), moved along the leg's walk-plane normal to the height the stance began at
        auto update_default_tip = [&](V3<K> stance_origin, V3<K> normal) {
          V3<K> idt{lk.identity_x, lk.identity_y + lk.span_dy, K(0)};
          idt = pose_transform(ldPose(sp, RS_WPP), idt);
          def = cvt<T>(idt + projection(stance_origin - idt, normal));
          if constexpr (EXT) {
            if (ext_default) {  // externally requested default (:988-992): pose_.removePose(transform_)
              const S* __restrict__ rp = sl + ci.roughS_leg * 32;
              def = cvt<T>(pose_transform(ldPose(rp, RT_ED_POSE), -ld3K(rp, RT_ED_TF)));
            }
          }
          st3(sl, LS::DEF, def);
          def_changed = true;
        };
        // ---- LegStepper::updateTipPosition (walk_controller.cpp:1018) ----
        const bool standard = (step_state == STEP_SWING || completed);
        const T stance_dt = standard ? ct.stance_dt_std : lt.stance_dt_mod;
        tgt = def + stride * T(0.5);  // uses the previous cycle's stride (trap 5)
        st3(sl, LS::TGT, tgt);
        if (step_state != STEP_FORCE_STOP) {
          // updateStride (:921); the body velocity and the walker's plane were written / are still held in this
          // robot's planes (L1 hits), which keeps them out of the loop's live registers
          const T bvx = scratch[0], bvy = scratch[32], bw = scratch[64];
          stride = V3<T>{bvx - bw * T(tipy), bvy + bw * T(tipx), T(0)} * ct.stride_scale;
          st3(sl, LS::STRIDE, stride);
          // walk_plane_ / walk_plane_normal_ of the leg = the walker's plane (:1049).  The copy in HBM already holds
          // it unless the plane changed in the previous cycle or the leg sat out while it did.
          if (!plane_saved || plane_changed_prev) {
            st3(sl, LS::WP, ld3T(sp, RS_WPL));
            st3(sl, LS::WPN, ld3T(sp, RS_WPN));
          }
          plane_saved = true;
          // tip += dt_u * B'(u) for a quartic Bezier B.  Only the four node differences enter B', and for all three
          // curves of the reference they collapse to three vectors:
          //   stance (generateStanceControlNodes :1295)         differences (sep, sep, sep, sep)
          //   swing, first half (generatePrimary... :1238)      (sep1, sep1, n3 - n2, n4 - n3)
          //   swing, second half (generateSecondary... :1267)   (m1 - m0, m2 - m1, sep2, sep2), with m1 - m0 = n4 - n3: the
          //                                                     first-half list read backwards, i.e. with (s, t) swapped
          // so every lane evaluates ONE curve  X (4 s^3 + 12 s^2 t) + Y 12 s t^2 + Z 4 t^3  on its own (X, Y, Z, s, t):
          // the swing / stance and first / second half branches only select operands.
          V3<T> X, Y, Z;
          T bs, bt, bdt;
          if (step_state == STEP_SWING) {
            int iteration = phase - ci.swing_start + 1;
            const int half = ci.swing_iterations / 2;
            const bool first_half = iteration <= half;
            V3<T> swo_p, swo_v;
            if (iteration == 1) {
              swo_p = V3<T>{T(tipx), T(tipy), T(tipz)};
              swo_v = swing_begins ? tipvel_prev : ld3T(sl, LS::TIPVEL);
              st3(sl, LS::SWO_P, swo_p);
              st3(sl, LS::SWO_V, swo_v);
            } else {
              swo_p = ld3T(ss, LS::SWO_P);
              swo_v = ld3T(ss, LS::SWO_V);
            }
            bool ground_contact = false;
            T ext_clearance = T(-1);  // >= 0: an external target scales the swing clearance (:1071)
            if constexpr (EXT) {
              if (f_rough) {
                // the leg's walk_plane_normal_ is the walker's as of this cycle's start (updateStride above)
                const V3<K> leg_wpn = ld3K(sp, RS_WPN);
                if (iteration == 1) update_default_tip(ld3K(sl, LS::STO_P), leg_wpn);  // (:1058-1061)
                // Target set to the externally requested pose, moved back by the robot's motion since the request (the caller
                // refreshes `transform` as the reference does from the tf tree) and, in the odometry frame, led by the distance
                // the body will cover until the swing ends (:1068-1078) — else moved to meet the step surface (:1081-1100)
                if (ext_target) {
                  const S* __restrict__ rp = sl + ci.roughS_leg * 32;
                  V3<K> t = pose_transform(ldPose(rp, RT_ET_POSE), -ld3K(rp, RT_ET_TF));
                  ext_clearance = T(rp[(RT_ET_CLR) * 32]);
                  if (ext_odom) {
                    const K time_to_swing_end = K(ci.swing_iterations - iteration) * ck.dt;
                    t.x -= K(bvx) * time_to_swing_end;
                    t.y -= K(bvy) * time_to_swing_end;
                  }
                  tgt = cvt<T>(t);
                  st3(sl, LS::TGT, tgt);
                } else if (touchdown_detection) {
                  if (step_plane_defined) {  // proactive: the step plane is known
                    const V3<K> target_tip_position = V3<K>{K(tipx), K(tipy), K(tipz)} + (step_plane - model_tip);
                    tgt = tgt + cvt<T>(projection(target_tip_position - cvt<K>(tgt), leg_wpn));
                  } else {  // reactive: reach down by the step depth and rely on contact detection
                    tgt.z -= ct.step_depth;
                  }
                  st3(sl, LS::TGT, tgt);
                }
                ground_contact = step_plane_defined;
              }
            }
            // Control nodes relative to the swing origin (the origin cancels in the differences).
            V3<T> clr{T(0), T(0), ct.swing_height};
            if (!plane_flat) clr = normalized(ld3T(sp, RS_WPN)) * ct.swing_height;
            if constexpr (EXT) {
              if (ext_clearance >= T(0)) clr = normalized(clr) * ext_clearance;
            }
            V3<T> tr = tgt - swo_p;
            V3<T> mid{tr.x * T(0.5) + clr.x, tr.y * T(0.5) + clr.y + lt.ysign * ct.swing_width, max_(T(0), tr.z) + clr.z};
            V3<T> sep1 = swo_v * (T(0.25) * (ct.dt / ct.swing_dt));
            V3<T> n2 = sep1 * T(2);
            V3<T> n3 = (mid + n2) * T(0.5);
            n3.z = mid.z;
            V3<T> ftv = -stride * (stance_dt / ct.dt);
            V3<T> sep2 = ftv * (T(0.25) * (ct.dt / ct.swing_dt));
            V3<T> m2 = tr - sep2 * T(2);
            V3<T> m1;
            if (ci.force_normal_touchdown && !ground_contact) {  // forceNormalTouchdown (:1314): n4 = m0 = bo, n3 = bo - h, m1 = bo + h
              V3<T> bo = tr - sep2 * T(4);
              bo.z = max_(T(0), tr.z);
              bo = bo + clr;
              Z = (m2 - bo) * T(0.5);
              n3 = bo - Z;
              m1 = bo + Z;
            } else {  // n4 = m0 = mid, m1 = n4 - (n3 - n4)
              Z = mid - n3;
              m1 = mid + Z;
            }
            X = first_half ? sep1 : sep2;
            Y = first_half ? n3 - n2 : m2 - m1;
            // ground contact in the second half (generateSecondarySwingControlNodes(true), :1282-1289): the nodes restart at
            // the tip, one stance node separation apart - the tip only keeps its touchdown velocity
            if (!first_half && ground_contact) { Y = sep2; Z = sep2; }
            const T u = ct.swing_dt * T(first_half ? iteration : iteration - half);
            bs = first_half ? T(1) - u : u;
            bt = first_half ? u : T(1) - u;
            bdt = ct.swing_dt;
          } else {  // STANCE / FORCE_STANCE
            int mod_start = standard ? ci.stance_start : ci.phase_offset[l];
            int iteration = phase - mod_start;  // mod(phase + (period - start), period) + 1, both in [0, period)
            iteration += iteration < 0 ? ci.period + 1 : 1;
            if (iteration == 1) {
              st3(sl, LS::STO_P, V3<T>{T(tipx), T(tipy), T(tipz)});
              ext_target = false;  // external_target_.defined_ = false after every swing period (:1159)
              if constexpr (EXT) {
                if (f_rough) update_default_tip(V3<K>{K(tipx), K(tipy), K(tipz)}, ld3K(sp, RS_WPN));  // (:1160-1163)
              }
            }
            T scaler = standard ? T(1) : lt.stride_scaler_mod;
            X = -stride * scaler * T(0.25);
            Y = X;
            Z = X;
            bt = T(iteration) * stance_dt;
            bs = T(1) - bt;
            bdt = stance_dt;
          }
          const T bss = bs * bs, btt = bt * bt;
          const T w01 = T(4) * bss * (bs + T(3) * bt), w2 = T(12) * bs * btt, w3 = T(4) * btt * bt;
          const V3<T> delta = (X * w01 + Y * w2 + Z * w3) * bdt;
          tipx += (double)delta.x;
          tipy += (double)delta.y;
          tipz += (double)delta.z;
          dl[(LD_TIP) * 32] = tipx;
          dl[(LD_TIP + 1) * 32] = tipy;
          dl[(LD_TIP + 2) * 32] = tipz;
          st3(sl, LS::TIPVEL, delta * ct.inv_dt);
        } else if (plane_changed_prev) {
          plane_saved = false;  // the leg sits this cycle out: its saved plane is now older than the walker's
        }

        if (f_tiprot) {
          // ---- LegStepper::updateTipRotation (:1193), on the progress values the cycle found ----
          const K swing_progress = swing_num < 0 ? K(-1) : K(swing_num) / K(ci.swing_period);
          if (stance_num >= 0 || swing_progress >= K(0.5)) {
            // the target rotation is the identity tip rotation for good (nothing un-defines it without rough-terrain targets)
            const Q4<K> target_rot{ck.tip_target_rot[0], ck.tip_target_rot[1], ck.tip_target_rot[2], ck.tip_target_rot[3]};
            tr_cur = correct_rotation(target_rot, tr_origin);
            if (swing_progress >= K(0.5)) {
              const K c = smooth_step(min_(K(1), K(2) * (swing_progress - K(0.5))));
              const V3<K> ux{K(1), K(0), K(0)};
              const V3<K> origin_dir = qrot(tr_origin, ux), target_dir = qrot(target_rot, ux);
              const V3<K> new_dir = origin_dir * (K(1) - c) + target_dir * c;
              tr_cur = correct_rotation(from_two_vectors(ux, normalized(new_dir)), tr_cur);
            }
          } else {
            // origin = Leg::current_tip_pose_.rotation_: the tip frame of the last applyFK in the base_link frame
            const V3<K> cx = t1_rotate(lk, ch.tipx), cy = t1_rotate(lk, ch.tipy), cz = t1_rotate(lk, ch.tipz);
            const K m[3][3] = {{cx.x, cy.x, cz.x}, {cx.y, cy.y, cz.y}, {cx.z, cy.z, cz.z}};
            tr_origin = qnormalized(matrix_to_quat(m));
            tr_cur = {K(0), K(0), K(0), K(0)};
            S* __restrict__ tp = sl + ci.tipS_leg * 32;
            tp[(TR_ORIGIN) * 32] = S(tr_origin.w); tp[(TR_ORIGIN + 1) * 32] = S(tr_origin.x);
            tp[(TR_ORIGIN + 2) * 32] = S(tr_origin.y); tp[(TR_ORIGIN + 3) * 32] = S(tr_origin.z);
          }
          S* __restrict__ tp = sl + ci.tipS_leg * 32;
          tp[(TR_CUR) * 32] = S(tr_cur.w); tp[(TR_CUR + 1) * 32] = S(tr_cur.x);
          tp[(TR_CUR + 2) * 32] = S(tr_cur.y); tp[(TR_CUR + 3) * 32] = S(tr_cur.z);
        }

        // ---- LegStepper::iteratePhase (:871) + updateStepState (:901) ----
        phase = phase + 1 == ci.period ? 0 : phase + 1;  // (phase + 1) % period with phase in [0, period)
        if (step_state != STEP_FORCE_STOP) {
          if (phase >= ci.swing_start && phase < ci.swing_end && step_state != STEP_FORCE_STANCE) step_state = STEP_SWING;
          else if (phase < ci.stance_end || phase >= ci.stance_start) step_state = STEP_STANCE;
        }
        if (step_state == STEP_SWING) {
          swing_num = min_(max_(phase - ci.swing_start + 1, 0), ci.swing_period);
          stance_num = -1;
        } else if (step_state == STEP_STANCE) {
          int since = phase - ci.stance_start;
          since += since < 0 ? ci.period + 1 : 1;
          stance_num = min_(max_(since, 0), ci.stance_period);
          swing_num = -1;
        } else if (step_state == STEP_FORCE_STOP) {
          stance_num = 0;
          swing_num = -1;
        }
      }
      // every lane has read what it needs from this slot: hand it back to the TMA unit for leg l + 2
      SHC_SYNCWARP();
      if (SHC_LANE0(lane) && l + kSlots < L) {
        fence_proxy_async_smem();
        issue_leg(l + kSlots);
      }
      SHC_STAMP(6 + 3 * l);
      bits = (phase & 0xffff) | (step_state << 16) | ((at_correct ? 1 : 0) << 18) | ((completed ? 1 : 0) << 19) |
             ((negate ? 1 : 0) << 20) | ((plane_saved ? 1 : 0) << LB_PLANE_SAVED) | ((step_plane_defined ? 1 : 0) << LB_STEP_PLANE) |
             ((touchdown_detection ? 1 : 0) << LB_TOUCHDOWN) | ((ext_target ? 1 : 0) << LB_EXT_TARGET) |
             ((ext_odom ? 1 : 0) << LB_EXT_ODOM) | ((ext_default ? 1 : 0) << LB_EXT_DEFAULT);
      prog = (swing_num & 0xffff) | ((stance_num & 0xffff) << 16);
      il[(LI_BITS) * 32] = bits;
      il[(LI_PROG) * 32] = prog;

      // ---- PoseController::updateStance (pose_controller.cpp:110) ----
      PoseT<K> leg_pose = cur_pose;
      if (f_auto) {
        leg_pose = pose_remove(leg_pose, auto_pose);
        leg_pose = pose_add(leg_pose, leg_auto);
      }
      V3<K> desired = pose_inverse_transform(leg_pose, V3<K>{K(tipx), K(tipy), K(tipz)});

      // ---- joint state + chain at the previous joint angles ----
      if constexpr (!EXT) leg_chain<K, D>(lk, q, ch);

      // ---- AdmittanceController::updateAdmittance (admittance_controller.cpp:22) ----
      if (f_adm) {
        K x0 = adm_x0, x1 = adm_x1;
        V3<K> force = adm_force * ck.force_gain;
        K fa[3] = {force.x, force.y, force.z};
        K da[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // the same 2-state is integrated once per axis (trap 2)
          K F = max_(fa[a], K(0));
          K n0 = ck.adm_P[0] * x0 + ck.adm_P[1] * x1 + ck.adm_q[0] * F;
          K n1 = ck.adm_P[2] * x0 + ck.adm_P[3] * x1 + ck.adm_q[1] * F;
          x0 = n0;
          x1 = n1;
          K dl = clamp_(-x0, K(-0.2), K(0.2));
          // delta_direction * (|delta| - deadband) / (1 - deadband) with deadband 0 (trap 8): delta / |delta| is exactly
          // +-1, so the product is delta itself for every non-zero delta
          da[a] = dl != K(0) ? dl : K(0);
        }
        sl[(LS::ADM_X) * 32] = S(x0);
        sl[(LS::ADM_X + 1) * 32] = S(x1);
        // Leg::setAdmittanceDelta (model.h:365): projection onto the tip frame x axis (base_link frame)
        V3<K> dirx = t1_rotate(lk, ch.tipx);
        V3<K> adelta = projection(V3<K>{da[0], da[1], da[2]}, dirx);
        st3(sl, LS::ADM_DELTA, adelta);
        desired = desired + adelta;  // Leg::setDesiredTipPose (model.cpp:653)
      }

      // ---- Leg::applyIK (model.cpp:861): one DLS step ----
      V3<K> des_leg;
      if (f_tiprot) {
        // the poser's tip rotation (updateStance, pose_controller.cpp:129) is the desired one: position step, rotation step,
        // and the position-only retry when the rotation-constrained attempt misses the tolerance
        const Q4<K> desired_rot = qmul(qinverse(leg_pose.q), tr_cur);
        bool ok;
        int force_updates;
        status |= apply_ik_pose<K, D>(ck, lk, ch, q, qd, desired, desired_rot, ci.clamp_joint_positions != 0,
                                      ci.clamp_joint_velocities != 0, &des_leg, &ok, &force_updates);
        if (!ok) status |= 1;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          sl[(LS::Q + j) * 32] = S(q[j]);
          sl[(LS::QD + j) * 32] = S(qd[j]);
          stage[l * D + j] = (float)(q[j] + lk.joffset[j]);
        }
        if (f_effort) {  // calculateTipForce runs at both levels of the retry (model.cpp:932-938)
          K tau[D];
#pragma unroll
          for (int j = 0; j < D; ++j) tau[j] = io.efforts ? K(io.efforts[((size_t)r * L + l) * D + j]) : K(0);
          const V3<K> raw = raw_tip_force<K, D>(ck, lk, ch, tau);
          V3<K> f = adm_force;
          for (int k = 0; k < force_updates; ++k) f = raw * (K(0.15) * ck.force_gain) + f * (K(1) - K(0.15));
          st3(sl, LS::ADM_FORCE, f);
        }
        continue;
      }
      status |= apply_ik_step<K, D>(ck, lk, ch, q, qd, desired, ci.clamp_joint_positions != 0, ci.clamp_joint_velocities != 0,
                                    &des_leg);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        sl[(LS::Q + j) * 32] = S(q[j]);
        sl[(LS::QD + j) * 32] = S(qd[j]);
        stage[l * D + j] = (float)(q[j] + lk.joffset[j]);  // state_controller.cpp:795
      }
      if (io.flags_out || f_effort) {
        // applyFK at the new joint angles (model.cpp:904) for the IK tolerance check (:916-929)
        Chain<K, D> ch2;
        leg_chain<K, D>(lk, q, ch2);
        V3<K> er = t1_rotate(lk, ch2.tip - des_leg);  // current - desired tip position in the base_link frame
        if (abs_(er.x) > K(0.005) || abs_(er.y) > K(0.005) || abs_(er.z) > K(0.005)) status |= 1;
        if (f_effort) {  // calculateTipForce (model.cpp:938)
          K tau[D];
#pragma unroll
          for (int j = 0; j < D; ++j) tau[j] = io.efforts ? K(io.efforts[((size_t)r * L + l) * D + j]) : K(0);
          V3<K> raw = raw_tip_force<K, D>(ck, lk, ch2, tau);
          V3<K> f = raw * (K(0.15) * ck.force_gain) + adm_force * (K(1) - K(0.15));
          st3(sl, LS::ADM_FORCE, f);
        }
      }
    }

    // =================================================================================================================
    // 4. updateWalkPlane (:748) + odometry (:783)
    // =================================================================================================================
    if (f_stiff) {
      // reset every leg to the global stiffness, then in leg order: a swinging leg takes its swing stiffness and adds
      // its load term to whatever its two neighbours hold at that point (a later swinging neighbour overwrites it)
      K kk[kMaxLegs], ref[kMaxLegs];
#pragma unroll
      for (int l = 0; l < kMaxLegs; ++l) {
        kk[l] = ck.virtual_stiffness;
        ref[l] = K(-1);
        if (l < L) ref[l] = K(sp[(ci.offS_leg + l * ci.strideS_leg + LS::STIFF) * 32]);
      }
#pragma unroll
      for (int l = 0; l < kMaxLegs; ++l) {
        if (l < L && ref[l] >= K(0)) {
          const K swing_k = ck.virtual_stiffness * (ref[l] * (ck.swing_stiffness_scaler - K(1)) + K(1));
          const K load_k = ck.virtual_stiffness * (ref[l] * (ck.load_stiffness_scaler - K(1)));
          const int a1 = l == 0 ? L - 1 : l - 1, a2 = l == L - 1 ? 0 : l + 1;
          // same order as the reference: both neighbours are read, then the three writes (so that legs coinciding on
          // robots of one or two legs end with the value the reference ends with)
          K c1 = K(0), c2 = K(0);
#pragma unroll
          for (int j = 0; j < kMaxLegs; ++j) {
            if (j == a1) c1 = kk[j];
            if (j == a2) c2 = kk[j];
          }
          kk[l] = swing_k;
#pragma unroll
          for (int j = 0; j < kMaxLegs; ++j) {
            if (j == a1) kk[j] = c1 + load_k;
            if (j == a2) kk[j] = c2 + load_k;
          }
        }
      }
#pragma unroll
      for (int l = 0; l < kMaxLegs; ++l)
        if (l < L) sp[(ci.offS_leg + l * ci.strideS_leg + LS::STIFF) * 32] = S(kk[l]);
    }

    // The plane is the least-squares fit of the legs' default tips, which only move when a stopping leg re-seats its
    // default tip (def_changed) — otherwise the fit of the previous cycle, already in HBM, is this cycle's result bit for
    // bit.  RB_PLANE_STALE (states written from outside) forces one fit.  A cycle without updateWalkPlane (trap 7)
    // leaves everything as it was.
    bool plane_changed = plane_changed_prev;
    bool plane_stale = (rbits >> RB_PLANE_STALE) & 1;
    if (!starting_now) {
      plane_changed = false;
      if (def_changed || plane_stale) {
        V3<double> new_wpl{0.0, 0.0, 0.0}, new_wpn{0.0, 0.0, 1.0};
        if (L >= 3) {
          // normal equations over the (possibly updated) default tips: A = [x y 1], b = z
          double sxx = 0, sxy = 0, sx1 = 0, syy = 0, sy1 = 0, sxz = 0, syz = 0, sz1 = 0;
#pragma unroll 1
          for (int l = 0; l < L; ++l) {
            const S* __restrict__ sl = sp + (ci.offS_leg + l * ci.strideS_leg) * 32;
            double x = (double)sl[(LS::DEF) * 32], y = (double)sl[(LS::DEF + 1) * 32], z = (double)sl[(LS::DEF + 2) * 32];
            sxx += x * x; sxy += x * y; sx1 += x; syy += y * y; sy1 += y; sxz += x * z; syz += y * z; sz1 += z;
          }
          // solve (A^T A) w = A^T b with A^T A = [[sxx sxy sx1],[sxy syy sy1],[sx1 sy1 L]]
          double n = (double)L;
          double c00 = syy * n - sy1 * sy1, c01 = sx1 * sy1 - sxy * n, c02 = sxy * sy1 - syy * sx1;
          double c11 = sxx * n - sx1 * sx1, c12 = sxy * sx1 - sxx * sy1;
          double c22 = sxx * syy - sxy * sxy;
          double det = sxx * c00 + sxy * c01 + sx1 * c02;
          double id = 1.0 / det;
          double a = (c00 * sxz + c01 * syz + c02 * sz1) * id;
          double b = (c01 * sxz + c11 * syz + c12 * sz1) * id;
          double cc = (c02 * sxz + c12 * syz + c22 * sz1) * id;
          V3<double> nrm = normalized(V3<double>{-a, -b, 1.0});
          new_wpl = V3<double>{a, b, cc};
          new_wpn = nrm;
        }
        // stored values as the next cycle will read them; "changed" compares the stored BITS: the first fit turns the
        // initial normal (0, 0, 1) into normalized(-0, -0, 1) = (-0, -0, 1), and the legs' saved copies must follow it
        // (they are exact copies in the reference, walk_controller.cpp:1049) so that a state record does not depend on
        // which cycles a leg happened to save its plane in
        const V3<S> ol = {sp[(RS_WPL) * 32], sp[(RS_WPL + 1) * 32], sp[(RS_WPL + 2) * 32]};
        const V3<S> on = {sp[(RS_WPN) * 32], sp[(RS_WPN + 1) * 32], sp[(RS_WPN + 2) * 32]};
        auto same = [](S a, S b) { return a == b && copysign(S(1), a) == copysign(S(1), b); };
        plane_changed = !(same(ol.x, S(new_wpl.x)) && same(ol.y, S(new_wpl.y)) && same(ol.z, S(new_wpl.z)) &&
                          same(on.x, S(new_wpn.x)) && same(on.y, S(new_wpn.y)) && same(on.z, S(new_wpn.z)));
        st3(sp, RS_WPL, new_wpl);
        st3(sp, RS_WPN, new_wpn);
        plane_stale = false;
      }
    }

    rbits = (walk_state & 3) | ((legs_at_correct & 15) << 2) | ((legs_completed & 15) << 6) | ((rtd & 1) << 10) |
            ((pose_state & 3) << 11) | ((auto_state & 3) << 13) | ((plane_changed ? 1 : 0) << RB_PLANE_CHANGED) |
            ((status & RB_STATUS_MASK) << RB_STATUS_SHIFT) | ((man_identity ? 1 : 0) << RB_MANUAL_IDENTITY) |
            (int)((plane_stale ? 1u : 0u) << RB_PLANE_STALE);
    ip[(RI_BITS) * 32] = rbits;
    if (io.flags_out && live) io.flags_out[r] = status;
    SHC_STAMP(30);
  }
};

}  // namespace shc
