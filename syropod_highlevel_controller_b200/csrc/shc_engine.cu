// shc_engine.cu — kernels + C-ABI of libshc_b200.so (include/shc_b200.h).
//
// Kernels (sm_100a):
//   control_cycle_kernel<P, D>  one fused control cycle, one thread per robot (DESIGN.md "Kernels")
//   apply_ik_kernel<D>          stand-alone batched Leg::applyIK (model.cpp:861) in double
// There is no CPU fallback: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/shc_b200.h"
#include "shc_msgs.cuh"
#include "shc_sequence.cuh"
#include "shc_pack.cuh"

namespace shc {

using PrecMixed = Prec<float, float, double>;  // fp32 state + trajectory; fp64 pose/kinematics (see DESIGN.md "Precision")

// One warp per block: warps never synchronise with each other, and 32-thread blocks let the block scheduler refill an SM
// warp by warp (measured 7 % faster than 128-thread blocks at the same register budget: profiles/).
#ifndef SHC_MIN_BLOCKS
#define SHC_MIN_BLOCKS 16
#endif
#ifndef SHC_BLOCK
#define SHC_BLOCK 32
#endif

// One warp per tile of 32 robots (lane = robot); warps are independent (no block-level synchronisation), each with
// its own slice of the dynamic shared memory: two TMA staging slots, the joint-command tile and two mbarriers.
// The full engine (auto / IMU / inclination posing, admittance) keeps more state live across the leg loop: it trades
// occupancy for registers (SHC_MIN_BLOCKS_FULL warps per SM) instead of spilling.
#ifndef SHC_MIN_BLOCKS_FULL
#define SHC_MIN_BLOCKS_FULL 12
#endif
// the tip-orientation path (6 x 6 solve, second DLS step and retry per leg) is not a throughput configuration: 8 warps per SM
#ifndef SHC_MIN_BLOCKS_TIP
#define SHC_MIN_BLOCKS_TIP 8
#endif
#ifdef SHC_MAXNREG
#define SHC_KERNEL_BOUNDS __maxnreg__(SHC_MAXNREG)
#else
#define SHC_KERNEL_BOUNDS __launch_bounds__(SHC_BLOCK, MODE == 2 ? SHC_MIN_BLOCKS_TIP : MODE == 1 ? SHC_MIN_BLOCKS_FULL : SHC_MIN_BLOCKS)
#endif
template <class P, int D, int MODE>
__global__ void SHC_KERNEL_BOUNDS control_cycle_kernel(const __grid_constant__ Consts c, Planes<typename P::S> pl, StepIO io) {
  extern __shared__ __align__(128) unsigned char shc_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = io.tile_begin + blockIdx.x * (SHC_BLOCK / 32) + warp;
  const int tile_first = tile * 32;
  if (tile >= io.tile_end || tile_first >= c.i.n_robots) return;  // whole warp
  using CY = Cycle<P, D, MODE>;
  const int front = CY::FULL ? c.i.frontS_leg : 0;
  unsigned char* wsm = shc_smem + (size_t)warp * c.i.smem_per_warp;
  CY::run(c, pl, tile, lane, io, wsm);
  __syncwarp();
  // The tile of joint commands leaves shared memory as TMA bulk stores: one request to the local buffer and, in the
  // fused all-gather, one per rank straight into that rank's gather buffer over NVLink (peer-mapped memory) — the
  // exchange overlaps the computation of the other tiles instead of running as a separate collective afterwards.
  const int LD = c.i.L * D;
  const int robots = min(32, c.i.n_robots - tile_first);
  const int valid = robots * LD;
  const size_t base = (size_t)tile_first * LD;
  const float* src = reinterpret_cast<const float*>(wsm + CY::kSlots * CY::slot_bytes(front));
  bool any_bulk = false;
  auto put = [&](float* dst) {
    // whole tiles go as one bulk store when the destination is 16-byte aligned (always, unless a shard's robot count makes
    // its slice of a gather buffer start off a 16-byte boundary); the batch's tail tile goes as plain coalesced stores
    if (robots == 32 && (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
      if (lane == 0) bulk_s2g(dst, src, (unsigned)(valid * 4));
      any_bulk = true;
    } else {
      for (int i = lane; i < valid; i += 32) dst[i] = src[i];
    }
  };
  if (lane == 0) fence_proxy_async_smem();  // the lanes' generic-proxy writes of the tile, ordered by the __syncwarp above
  if (io.joints_out) put(io.joints_out + base);
  if (io.gather_mc) {
    // NVSwitch multicast: ONE store per 16 bytes leaves this GPU and the switch replicates it into every rank's gather
    // buffer (this rank's included) — egress per rank and cycle is the shard itself, not (world - 1) copies of it
    float* dst = io.gather_mc + io.gather_offset + base;
    if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0 && (valid & 3) == 0) {
      const float4* src4 = reinterpret_cast<const float4*>(src);
      for (int i = lane; i < valid / 4; i += 32) multimem_st_v4(dst + 4 * i, src4[i]);
    } else {
      for (int i = lane; i < valid; i += 32) multimem_st_f32(dst + i, src[i]);
    }
  }
  for (int p = 0; p < io.n_gather; ++p) put(io.gather[p] + io.gather_offset + base);
  if (any_bulk && lane == 0) bulk_commit_and_wait_read();  // shared memory may go away once the TMA unit has read it
#ifdef SHC_TRACE
  if (io.trace && lane == 0) io.trace[(size_t)tile * 32 + 31] = gtimer();
#endif
}

// "This rank's shard of one more cycle has landed everywhere": launched on a side stream behind the control-cycle kernel
// (its stores are performed when it completes; the system-scope release below waits for the posted NVLink writes to
// drain), it bumps this rank's landed counter in every rank's flag array — through the multicast address when there is
// one (multimem.red, the same path the data took), else with one st.release.sys per peer.  Counters are monotonic: the
// value is the number of cycles of this source rank that have landed.
struct SignalArgs {
  int* flag[8];
  int* mc_flag;
  int n;
  int count;                  // cycles this signal accounts for
  unsigned long long* trace;  // tuning (SHC_GATHER_TRACE): globaltimer stamps of this launch, else null
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void gather_signal_kernel(SignalArgs a) {
  if (a.trace && threadIdx.x == 0) a.trace[1] = global_ns();
  if (a.mc_flag) {
    if (threadIdx.x == 0) multimem_red_release_add(a.mc_flag, a.count);
  } else if (threadIdx.x < a.n) {
    red_release_sys_add(a.flag[threadIdx.x], a.count);
  }
  if (a.trace && threadIdx.x == 0) a.trace[2] = global_ns();
}
__global__ void gather_stamp_kernel(unsigned long long* slot) { *slot = global_ns(); }
// Device-side wait on the local landed counters (ld.acquire.sys spin, one lane per source rank): the stream continues
// once every source has landed `need[p]` cycles here.  Gives up after ~4 s (a peer died) and records it in *err instead of
// hanging the GPU.
struct WaitArgs {
  const int* flags;
  int need[8];
  int n;
  int* err;
  unsigned long long* trace;
};
__global__ void gather_wait_kernel(WaitArgs a) {
  if (a.trace && threadIdx.x == 0) a.trace[0] = global_ns();
  if (threadIdx.x >= a.n) return;
  const int* f = a.flags + threadIdx.x;
  const long long t0 = clock64();
  while (load_acquire_sys(f) - a.need[threadIdx.x] < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 8000000000ll) {
      if (a.err) *a.err = 1 + threadIdx.x;
      break;
    }
  }
  if (a.trace) atomicMax(a.trace + 1, global_ns());
}

template <int D>
__global__ void __launch_bounds__(128) apply_ik_kernel(const __grid_constant__ Consts c, int n, const int* __restrict__ leg_id,
                                                       double* __restrict__ q_io, double* __restrict__ qd_io,
                                                       const double* __restrict__ desired, int simulation,
                                                       double* __restrict__ tip_out, double* __restrict__ result) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RealConsts<double>& ck = c.d;
  const int leg = leg_id[i];
  double q[D], qd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    q[j] = q_io[(size_t)i * D + j];
    qd[j] = qd_io[(size_t)i * D + j];
  }
  Chain<double, D> ch;
  const LegConsts<double>& lc = ck.leg[leg];
  leg_chain<double, D>(lc, q, ch);
  V3<double> des{desired[3 * (size_t)i], desired[3 * (size_t)i + 1], desired[3 * (size_t)i + 2]};
  V3<double> des_leg;
  apply_ik_step<double, D>(ck, lc, ch, q, qd, des, c.i.clamp_joint_positions != 0, c.i.clamp_joint_velocities != 0 && !simulation,
                           &des_leg);
  Chain<double, D> ch2;
  leg_chain<double, D>(lc, q, ch2);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    q_io[(size_t)i * D + j] = q[j];
    qd_io[(size_t)i * D + j] = qd[j];
  }
  if (tip_out) {
    V3<double> t = t1_rotate(lc, ch2.tip) + V3<double>{lc.t1p[0], lc.t1p[1], lc.t1p[2]};
    tip_out[3 * (size_t)i] = t.x;
    tip_out[3 * (size_t)i + 1] = t.y;
    tip_out[3 * (size_t)i + 2] = t.z;
  }
  if (result) result[i] = ik_result_value<double, D>(lc, ch2, q, des_leg);
}

// ---- output wire formats -----------------------------------------------------------------------
// One thread per (robot, leg) of the range: the LegState record and the leg's slice of the JointState record; the thread of
// leg 0 adds the robot's body record.  Replaces the per-leg / per-joint host loops of the reference's publishers.
template <class S, int D>
__global__ void __launch_bounds__(128) pack_messages_kernel(const __grid_constant__ Consts c, Planes<S> pl, int first, int count,
                                                            const float* __restrict__ measured, shc_joint_state_msg* js,
                                                            shc_leg_state_msg* legs, shc_body_msg* body) {
  const int L = c.i.L;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)count * L) return;
  const int k = (int)(i / L), l = (int)(i % L), r = first + k;
  shc_leg_state_msg m;
  pack_leg_message<S, D>(c, pl, r, l, measured ? measured + ((size_t)r * L + l) * D : nullptr, m, js ? &js[k] : nullptr);
  if (legs) legs[(size_t)k * L + l] = m;
  if (l == 0 && body) pack_body_message<S>(c, pl, r, body[k]);
}

// ---- start-up on the device ---------------------------------------------------------------------
// Copies tile 0 of a plane set over the tiles 1 .. n_tiles-1 (a new batch is n_tiles copies of one initial tile).
template <class W> __global__ void replicate_tile_kernel(W* planes, size_t tile_words, size_t n_tiles) {
  const size_t total = tile_words * (n_tiles - 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    planes[tile_words + i] = planes[i % tile_words];
}

// One cycle of PoseController::directStartup (pose_controller.cpp:463) for the batch = LegPoser::transitionConfiguration
// (:1476) of every joint of every robot: iteration `it` of the joint-space move from the robot's own origin configuration
// (its joint angles when the start-up began) to the batch-wide desired configuration.  One thread per joint.  The joint
// commands go to joints_out; the last iteration also seats the configuration in the state planes (velocities zero).
template <class S, int D>
__global__ void __launch_bounds__(256) transition_configuration_kernel(const __grid_constant__ Consts c, Planes<S> pl, const double* __restrict__ origin,
                                                                       const double* __restrict__ desired, int it, int num,
                                                                       float* __restrict__ joints_out, int write_state) {
  const int L = c.i.L;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)c.i.n_robots * L * D) return;
  const int j = (int)(i % D), l = (int)((i / D) % L), r = (int)(i / ((long long)D * L));
  const LegConsts<double>& lc = c.d.leg[l];
  const double o = origin ? origin[i] : clamp_(0.0, lc.jmin[j], lc.jmax[j]);  // Joint::default_position_ (model.cpp:1038)
  const double q = transition_configuration(o, desired[l * kMaxDof + j], it, num);
  if (joints_out) joints_out[i] = (float)(q + lc.joffset[j]);
  if (write_state) {
    S* sp_ = pl.s + ((size_t)(r >> 5) * c.i.nS + c.i.offS_leg + l * c.i.strideS_leg) * 32 + (r & 31);
    sp_[(LegS<D>::Q + j) * 32] = S(q);
    sp_[(LegS<D>::QD + j) * 32] = S(0);
  }
}

// Leg::generateWorkspace (model.cpp:309-510) for every leg: one block per leg, eight lanes = the eight bearings of a plane.
template <int D>
__global__ void __launch_bounds__(32) workspace_sweep_kernel(const __grid_constant__ Consts c, StartupParams sp, const double* __restrict__ qdef,
                                                             int full, int max_planes, double* heights, double* radii, int* n_planes) {
  const int l = blockIdx.x, lane = threadIdx.x;
  if (lane >= 8) return;
  const int np = workspace_sweep_leg<D>(c.d, sp, l, qdef + l * kMaxDof, full != 0, max_planes, lane, 8, heights + (size_t)l * max_planes,
                                        radii + (size_t)l * max_planes * SHC_N_BEARINGS);
  if (lane == 0) n_planes[l] = np;
}


// PoseController::stepToNewStance (pose_controller.cpp:520) for every robot: one thread per robot (csrc/shc_sequence.cuh).
template <class S, int D>
__global__ void __launch_bounds__(128) step_to_new_stance_kernel(const __grid_constant__ Consts c, Planes<S> pl, SeqBuffers sq, NewStanceParams np,
                                                                 float* __restrict__ joints_out, int* __restrict__ progress_out, int* min_progress) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.i.n_robots) return;
  const int p = step_to_new_stance_robot<S, D>(c, pl, sq, np, r, joints_out);
  if (progress_out) progress_out[r] = p;
  atomicMin(min_progress, p);
}
template <class S, int D>
__global__ void __launch_bounds__(128) execute_sequence_kernel(const __grid_constant__ Consts c, Planes<S> pl, SeqBuffers sq, ExecuteSequenceParams ep,
                                                               int shut_down, float* __restrict__ joints_out, int* __restrict__ progress_out,
                                                               int* min_progress) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.i.n_robots) return;
  const int p = execute_sequence_robot<S, D>(c, pl, sq, ep, shut_down != 0, r, joints_out);
  if (progress_out) progress_out[r] = p;
  atomicMin(min_progress, p);
}
template <class S, int D>
__global__ void __launch_bounds__(256) transition_joint_kernel(const __grid_constant__ Consts c, Planes<S> pl, const double* __restrict__ origin,
                                                               const double* __restrict__ desired, int it, int num, float* __restrict__ joints_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)c.i.n_robots * c.i.L * D) return;
  transition_joint<S, D>(c, pl, origin, desired, it, num, i, joints_out);
}
template <class S, int D>
__global__ void __launch_bounds__(256) latch_joint_kernel(const __grid_constant__ Consts c, Planes<S> pl, double* __restrict__ origin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)c.i.n_robots * c.i.L * D) return;
  latch_joint<S, D>(c, pl, i, origin);
}
__global__ void fill_int_kernel(int* p, size_t n, int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace shc

using namespace shc;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
template <class F> static int dispatch_D(int D, F&& f) {
  int rc = dispatch_D_raw(D, f);
  if (rc == SHC_E_UNSUPPORTED && D != 3 && D != 4 && D != 5) return fail(SHC_E_UNSUPPORTED, "unsupported joint count");
  return rc;
}
#define CUDA_TRY(x)                                                                                   \
  do {                                                                                                \
    cudaError_t err__ = (x);                                                                          \
    if (err__ != cudaSuccess) return fail(SHC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// Fused gather: cycle k writes buffer k % B of every rank (B = 24); see shc_gather_step for the protocol.  Reuse checks and
// landed signals both come every 8th cycle; with 24 buffers a reuse check asks for a count the peers signalled a whole
// signalling period earlier (need = cycle - 15 against a counter that reads cycle - 8 even if this cycle's signal is still
// in flight), so it never waits on a signal that has only just been issued.
constexpr int kGatherBuffers = 24;
constexpr int kGatherWaitEvery = 8;   // reuse check every 8th cycle, for the next 8 cycles
constexpr size_t kMaxCachedGraphs = 16;
constexpr int kHostChunks = 8;  // tile ranges of one shc_step_host call (kernel k+1 overlaps the D2H of range k)

struct GraphKey {
  int k;
  const float *cmd, *imu, *force;
  float* out;
  bool operator<(const GraphKey& o) const {
    if (k != o.k) return k < o.k;
    if (cmd != o.cmd) return cmd < o.cmd;
    if (imu != o.imu) return imu < o.imu;
    if (force != o.force) return force < o.force;
    return out < o.out;
  }
};

struct shc_engine {
  shc_config cfg;
  shc_startup su;
  Consts c;
  int device = 0;
  int precision = SHC_PRECISION_F64;
  int n = 0, n_pad = 0;
  int options = 0;
  int pose_reset_mode = 0;
  size_t s_elem = 8;  // bytes per storage word
  size_t smem_block = 0;  // dynamic shared memory per block of the control-cycle kernel
#ifdef SHC_TRACE
  unsigned long long* trace = nullptr;
#endif
  void* s_planes = nullptr;
  double* d_planes = nullptr;
  int* i_planes = nullptr;
  int* d_flags = nullptr;
  const float* d_efforts = nullptr;
  const float* d_step_planes = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy = nullptr;                       // D2H of shc_step_host, behind the tile-range kernels
  cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // pinned staging for shc_step_host
  float *h_cmd = nullptr, *h_imu = nullptr, *h_force = nullptr, *h_manual = nullptr, *h_out = nullptr;
  float *d_cmd = nullptr, *d_imu = nullptr, *d_force = nullptr, *d_manual = nullptr, *d_out = nullptr;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  // multi-GPU: NCCL communicator (opaque) + side stream / events for the overlapped per-cycle all-gather
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  bool done_valid[2] = {false, false};
  // fused all-gather over peer memory: every rank's gather buffer ([kGatherBuffers][world][N][L][D] floats + one landed
  // counter per source rank), mapped here either through CUDA IPC (shc_gather_alloc / shc_gather_open_peer: cudaMalloc'ed
  // and owned by the engines) or handed in by the caller (shc_gather_attach: e.g. torch symmetric memory, which also
  // provides the NVSwitch multicast mapping of the same buffers)
  float* gather_own = nullptr;
  float* gather_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* gather_mc = nullptr;
  bool gather_owned = false;
  bool gather_opened[8] = {false, false, false, false, false, false, false, false};
  // direct start-up on the device (shc_startup_begin / shc_startup_step)
  double* startup_origin = nullptr;   // [N][L][D] joint angles when the start-up began
  double* startup_desired = nullptr;  // [kMaxLegs][kMaxDof] desired configuration
  bool startup_has_origin = false;
  int startup_iteration = 0, startup_num = 0;
  // stepping / joint-space sequences on the device (shc_step_to_new_stance, shc_transition_*, shc_pack_legs / shc_unpack_legs)
  double* seq_origin = nullptr;   // [L][3][n_pad] LegPoser::origin_tip_pose_.position_
  int* seq_count = nullptr;       // [L][n_pad] LegPoser::master_iteration_count_ (-1: first_iteration_)
  int* seq_robot = nullptr;       // [2][n_pad] legs_completed_step_, current_group_
  int* seq_min = nullptr;         // device word: smallest progress of the last stepToNewStance / executeSequence launch
  double* seq_target = nullptr;   // executeSequence: [L][3][n_pad] LegPoser::target_tip_pose_.position_
  double* seq_poses = nullptr;    //                  [kMaxTransitionPoses][L][3][n_pad] LegPoser::transition_poses_
  int* seq_leg = nullptr;         //                  [2][L][n_pad] leg_completed_step_, number of transition poses
  double* tr_origin = nullptr;    // [N][L][D] joint positions when the transition began
  double* tr_desired = nullptr;   // [kMaxLegs][kMaxDof]
  int tr_iteration = 0, tr_num = 0;
  bool tr_executing = false;      // PoseController::executing_transition_ (pack / unpack)
  int* gather_err = nullptr;  // mapped pinned word: set by a device-side wait that gave up
  cudaStream_t signal = nullptr;  // high-priority stream of the landed-signal kernels
  unsigned long long* gather_trace = nullptr;  // tuning (SHC_GATHER_TRACE): mapped pinned stamps [4096][4]: kernel end, signal
  long long gather_trace_waits = 0;            // start, signal end, -; waits from row 2048: start, end
  cudaEvent_t ev_kernel[kGatherBuffers] = {};  // "cycle's kernel done", for the side-stream signal kernels
  long long gather_cycle = 0;   // cycles issued (same on every rank)
  long long gather_waited = 0;  // every source is known to have landed at least this many cycles here
  long long gather_signalled = 0;  // cycles whose landed signal has been issued (<= gather_cycle)
  int gather_signal_every = 8;     // a landed signal every so many cycles inside a call (and at every shc_gather_sync)
  int gather_signal_memop = 0;     // 1: the signal is a stream memory operation (cuStreamWriteValue32) instead of a kernel
};

// NCCL is resolved at run time from the libnccl already loaded in the process (torch's), so libshc_b200.so has no
// link-time dependency on it and single-GPU users never touch it.
namespace {
struct NcclUid { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (int (*)(NcclUid*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(h, "ncclCommInitRank");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
      api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
      api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.CommDestroy;
    }
  }
  return api;
}
}  // namespace

// ---- state record <-> planes (host side; get/set are not on the hot path) -----------------------------------------
namespace {
int download(shc_engine* e, HostPlanes& h) {
  const IntConsts& ci = e->c.i;
  const size_t np = ci.n_pad;
  h.s.resize((size_t)ci.nS * np);
  h.d.resize((size_t)ci.nD * np);
  h.i.resize((size_t)ci.nI * np);
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->precision == SHC_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(h.s.data(), e->s_planes, h.s.size() * 8, cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(h.s.size());
    CUDA_TRY(cudaMemcpy(tmp.data(), e->s_planes, tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < tmp.size(); ++k) h.s[k] = tmp[k];
  }
  CUDA_TRY(cudaMemcpy(h.d.data(), e->d_planes, h.d.size() * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(h.i.data(), e->i_planes, h.i.size() * 4, cudaMemcpyDeviceToHost));
  return SHC_OK;
}

// The tiles [t0, t1) only: a tile's planes are contiguous in each plane set (tile-major layout), so a robot range costs
// three small copies however large the batch is.
int download_tiles(shc_engine* e, size_t t0, size_t t1, HostPlanes& h) {
  const IntConsts& ci = e->c.i;
  const size_t nt = t1 - t0;
  h.s.resize((size_t)ci.nS * 32 * nt);
  h.d.resize((size_t)ci.nD * 32 * nt);
  h.i.resize((size_t)ci.nI * 32 * nt);
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->precision == SHC_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(h.s.data(), (const double*)e->s_planes + t0 * ci.nS * 32, h.s.size() * 8, cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(h.s.size());
    CUDA_TRY(cudaMemcpy(tmp.data(), (const float*)e->s_planes + t0 * ci.nS * 32, tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < tmp.size(); ++k) h.s[k] = tmp[k];
  }
  CUDA_TRY(cudaMemcpy(h.d.data(), e->d_planes + t0 * ci.nD * 32, h.d.size() * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(h.i.data(), e->i_planes + t0 * ci.nI * 32, h.i.size() * 4, cudaMemcpyDeviceToHost));
  return SHC_OK;
}

int upload(shc_engine* e, const HostPlanes& h) {
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->precision == SHC_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(e->s_planes, h.s.data(), h.s.size() * 8, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> tmp(h.s.size());
    for (size_t k = 0; k < tmp.size(); ++k) tmp[k] = (float)h.s[k];
    CUDA_TRY(cudaMemcpy(e->s_planes, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(e->d_planes, h.d.data(), h.d.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->i_planes, h.i.data(), h.i.size() * 4, cudaMemcpyHostToDevice));
  return SHC_OK;
}

// The tiles [t0, t1) back to the device (h holds exactly those tiles).
int upload_tiles(shc_engine* e, size_t t0, size_t t1, const HostPlanes& h) {
  const IntConsts& ci = e->c.i;
  (void)t1;
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->precision == SHC_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy((double*)e->s_planes + t0 * ci.nS * 32, h.s.data(), h.s.size() * 8, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> tmp(h.s.size());
    for (size_t k = 0; k < tmp.size(); ++k) tmp[k] = (float)h.s[k];
    CUDA_TRY(cudaMemcpy((float*)e->s_planes + t0 * ci.nS * 32, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(e->d_planes + t0 * ci.nD * 32, h.d.data(), h.d.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->i_planes + t0 * ci.nI * 32, h.i.data(), h.i.size() * 4, cudaMemcpyHostToDevice));
  return SHC_OK;
}

}  // namespace


// Every robot of the batch into the post-start-up state: one packed tile goes up, the device replicates it.
static int reset_to_initial_state(shc_engine* e) {
  const IntConsts& ci = e->c.i;
  const size_t n_tiles = ci.n_pad / 32;
  HostPlanes t;
  initial_tile(e, t);
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->precision == SHC_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(e->s_planes, t.s.data(), t.s.size() * 8, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> tmp(t.s.begin(), t.s.end());
    CUDA_TRY(cudaMemcpy(e->s_planes, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(e->d_planes, t.d.data(), t.d.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->i_planes, t.i.data(), t.i.size() * 4, cudaMemcpyHostToDevice));
  if (n_tiles > 1) {
    const int blocks = 148 * 8;
    if (e->precision == SHC_PRECISION_F64) replicate_tile_kernel<double><<<blocks, 256>>>((double*)e->s_planes, t.s.size(), n_tiles);
    else replicate_tile_kernel<float><<<blocks, 256>>>((float*)e->s_planes, t.s.size(), n_tiles);
    replicate_tile_kernel<double><<<blocks, 256>>>(e->d_planes, t.d.size(), n_tiles);
    replicate_tile_kernel<int><<<blocks, 256>>>(e->i_planes, t.i.size(), n_tiles);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaDeviceSynchronize());
  return SHC_OK;
}

// Calls f(kernel pointer, Planes) for the instantiation this engine runs.
template <class F> static int with_cycle_kernel(shc_engine* e, F&& f) {
  const int mode = engine_mode(e->cfg);
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      Planes<double> pl{(double*)e->s_planes, e->d_planes, e->i_planes};
      return mode == 2 ? f(control_cycle_kernel<PrecF64, D, 2>, pl)
             : mode == 1 ? f(control_cycle_kernel<PrecF64, D, 1>, pl) : f(control_cycle_kernel<PrecF64, D, 0>, pl);
    }
    Planes<float> pl{(float*)e->s_planes, e->d_planes, e->i_planes};
    return mode == 2 ? f(control_cycle_kernel<PrecMixed, D, 2>, pl)
           : mode == 1 ? f(control_cycle_kernel<PrecMixed, D, 1>, pl) : f(control_cycle_kernel<PrecMixed, D, 0>, pl);
  });
}

// Shared memory of the control-cycle kernel is dynamic (it depends on L, D, the precision and the admittance block):
// size it, opt in above 48 KB and ask for the full shared-memory carve-out once per engine.
static int configure_cycle_kernel(shc_engine* e) {
  const bool full = engine_full(e->cfg);
  const int front = full ? e->c.i.frontS_leg : 0;
  dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    e->c.i.smem_per_warp = e->precision == SHC_PRECISION_F64 ? Cycle<PrecF64, D, 0>::smem_per_warp(front, e->cfg.leg_count)
                                                            : Cycle<PrecMixed, D, 0>::smem_per_warp(front, e->cfg.leg_count);
    return 0;
  });
  e->smem_block = (size_t)e->c.i.smem_per_warp * (SHC_BLOCK / 32);
  if (const char* pad = getenv("SHC_SMEM_PAD")) e->smem_block += (size_t)atoi(pad);  // kernel tuning: lowers the occupancy
  return with_cycle_kernel(e, [&](auto kernel, auto) -> int {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_block));
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return SHC_OK;
  });
}

// Launches the control cycle for the tiles [io.tile_begin, tile_end) (tile = 32 robots).
static int launch_cycle(shc_engine* e, const StepIO& io, cudaStream_t st) {
  const int threads = SHC_BLOCK;
  const int tiles = io.tile_end - io.tile_begin;
  if (tiles <= 0) return SHC_OK;
  const int blocks = (tiles + threads / 32 - 1) / (threads / 32);
  return with_cycle_kernel(e, [&](auto kernel, auto pl) -> int {
    kernel<<<blocks, threads, e->smem_block, st>>>(e->c, pl, io);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("control_cycle launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
}

extern "C" {

const char* shc_last_error(void) { return g_err.c_str(); }
size_t shc_sizeof_config(void) { return sizeof(shc_config); }
size_t shc_sizeof_startup(void) { return sizeof(shc_startup); }
size_t shc_sizeof_robot_state(void) { return sizeof(shc_robot_state); }

int shc_create(const shc_config* cfg, const shc_startup* startup, int n_robots, int device, int precision, shc_engine** out) {
  if (!cfg || !out || n_robots < 1) return fail(SHC_E_INVALID, "shc_create: bad arguments");
  if (precision != SHC_PRECISION_F64 && precision != SHC_PRECISION_MIXED) return fail(SHC_E_INVALID, "unknown precision");
  std::string err;
  bool unsupported = false;
  if (!check_supported(*cfg, err, unsupported)) return fail(unsupported ? SHC_E_UNSUPPORTED : SHC_E_INVALID, err);
  if (!startup && !own_startup_supported(*cfg, err)) return fail(SHC_E_UNSUPPORTED, err);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    return fail(SHC_E_CUDA, "no CUDA device: the SHC engine has no CPU fallback");
  if (device < 0 || device >= count) return fail(SHC_E_INVALID, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));

  shc_engine* e = new shc_engine();
  e->device = device;
  core_init(e, *cfg, startup, n_robots, precision);
  if (!check_step_cycle(e->su, err)) { delete e; return fail(SHC_E_UNSUPPORTED, err); }

  e->s_elem = precision == SHC_PRECISION_F64 ? 8 : 4;
  {
    int rc = configure_cycle_kernel(e);
    if (rc != SHC_OK) { std::string m = g_err; shc_destroy(e); return fail(rc, m); }
  }
  const IntConsts& ci = e->c.i;
  const size_t np = ci.n_pad;
  auto cleanup = [&](int code, const std::string& m) { shc_destroy(e); return fail(code, m); };
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(SHC_E_CUDA, "stream create failed");
  if (cudaMalloc(&e->s_planes, (size_t)ci.nS * np * e->s_elem) != cudaSuccess ||
      cudaMalloc((void**)&e->d_planes, (size_t)ci.nD * np * 8) != cudaSuccess ||
      cudaMalloc((void**)&e->i_planes, (size_t)ci.nI * np * 4) != cudaSuccess ||
      cudaMalloc((void**)&e->d_flags, np * 4) != cudaSuccess)
    return cleanup(SHC_E_CUDA, "device allocation failed");
  if (cudaMemset(e->s_planes, 0, (size_t)ci.nS * np * e->s_elem) != cudaSuccess ||
      cudaMemset(e->d_planes, 0, (size_t)ci.nD * np * 8) != cudaSuccess ||
      cudaMemset(e->i_planes, 0, (size_t)ci.nI * np * 4) != cudaSuccess || cudaMemset(e->d_flags, 0, np * 4) != cudaSuccess)
    return cleanup(SHC_E_CUDA, "device memset failed");

  // every robot starts in the post-start-up state
  {
    int rc = reset_to_initial_state(e);
    if (rc != SHC_OK) { std::string m = g_err; return cleanup(rc, m); }
  }
  *out = e;
  return SHC_OK;
}

// Start-up constants only (host arithmetic, no device needed): the engine's restatement of generateStepCycle,
// directStartup, generateWorkspaces, generateWalkspace and generateLimits for `cfg`.
int shc_compute_startup(const shc_config* cfg, shc_startup* out) {
  if (!cfg || !out) return fail(SHC_E_INVALID, "null argument");
  std::string err;
  bool unsupported = false;
  if (!validate_config(*cfg, err, unsupported)) return fail(unsupported ? SHC_E_UNSUPPORTED : SHC_E_INVALID, err);
  RealConsts<double>* ck = new RealConsts<double>();
  fill_static_consts<double>(*cfg, *ck);
  std::memset(out, 0, sizeof(*out));
  compute_step_cycle(*cfg, *out);
  dispatch_D(cfg->joint_count, [&](auto dtag) -> int { compute_startup<decltype(dtag)::value>(*cfg, *ck, *out); return 0; });
  delete ck;
  return SHC_OK;
}

// Host evaluation (double) of the same Leg::applyIK routine the kernels run, for CPU-side unit tests of the shared
// kinematics code and for the start-up sweeps.  q/qd [D] in/out, desired/tip in the base_link frame.
int shc_host_apply_ik(const shc_config* cfg, int leg, double* q, double* qd, const double* desired, int simulation,
                      double* tip_out, double* ik_result) {
  if (!cfg || !q || !qd || !desired) return fail(SHC_E_INVALID, "null argument");
  std::string err;
  bool unsupported = false;
  if (!validate_config(*cfg, err, unsupported)) return fail(unsupported ? SHC_E_UNSUPPORTED : SHC_E_INVALID, err);
  if (leg < 0 || leg >= cfg->leg_count) return fail(SHC_E_INVALID, "leg out of range");
  RealConsts<double>* ck = new RealConsts<double>();
  fill_static_consts<double>(*cfg, *ck);
  int rc = dispatch_D(cfg->joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    V3<double> tip;
    double res = host_apply_ik<D>(*ck, *cfg, leg, q, qd, V3<double>{desired[0], desired[1], desired[2]}, simulation != 0, &tip);
    if (tip_out) { tip_out[0] = tip.x; tip_out[1] = tip.y; tip_out[2] = tip.z; }
    if (ik_result) *ik_result = res;
    return SHC_OK;
  });
  delete ck;
  return rc;
}

void shc_destroy(shc_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  if (e->nccl_comm && nccl().ok) nccl().CommDestroy(e->nccl_comm);
  for (int b = 0; b < 2; ++b) {
    if (e->ev_ready[b]) cudaEventDestroy(e->ev_ready[b]);
    if (e->ev_done[b]) cudaEventDestroy(e->ev_done[b]);
  }
  if (e->side) cudaStreamDestroy(e->side);
  if (e->signal) cudaStreamDestroy(e->signal);
  if (e->copy) cudaStreamDestroy(e->copy);
  for (int p = 0; p < 8; ++p)
    if (e->gather_opened[p]) cudaIpcCloseMemHandle(e->gather_peer[p]);
  if (e->gather_owned) cudaFree(e->gather_own);
  if (e->gather_err) cudaFreeHost(e->gather_err);
  if (e->gather_trace) cudaFreeHost(e->gather_trace);
  for (int b = 0; b < kGatherBuffers; ++b) {
    if (e->ev_kernel[b]) cudaEventDestroy(e->ev_kernel[b]);
  }
  for (auto ev : e->ev_chunk)
    if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->s_planes);
  cudaFree(e->d_planes);
  cudaFree(e->i_planes);
  cudaFree(e->d_flags);
  cudaFree(e->startup_origin);
  cudaFree(e->startup_desired);
  cudaFree(e->seq_origin);
  cudaFree(e->seq_count);
  cudaFree(e->seq_robot);
  cudaFree(e->seq_min);
  cudaFree(e->seq_target);
  cudaFree(e->seq_poses);
  cudaFree(e->seq_leg);
  cudaFree(e->tr_origin);
  cudaFree(e->tr_desired);
  cudaFree(e->d_cmd); cudaFree(e->d_imu); cudaFree(e->d_force); cudaFree(e->d_manual); cudaFree(e->d_out);
  cudaFreeHost(e->h_cmd); cudaFreeHost(e->h_imu); cudaFreeHost(e->h_force); cudaFreeHost(e->h_manual); cudaFreeHost(e->h_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int shc_get_startup(const shc_engine* e, shc_startup* out) {
  if (!e || !out) return fail(SHC_E_INVALID, "null argument");
  *out = e->su;
  return SHC_OK;
}
int shc_n_robots(const shc_engine* e) { return e ? e->n : 0; }
int shc_options(const shc_engine* e) { return e ? e->options : 0; }
// Captured rollouts bake the kernel arguments in (options, pose reset mode, efforts pointer): changing any of them drops
// the cached graphs, so that the next shc_rollout captures the new values.
static void drop_graphs(shc_engine* e) {
  if (e->graphs.empty()) return;
  cudaSetDevice(e->device);
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
}
int shc_set_options(shc_engine* e, int options) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (options != e->options) drop_graphs(e);
  e->options = options;
  return SHC_OK;
}
int shc_set_pose_reset_mode(shc_engine* e, int mode) {
  if (!e || mode < 0 || mode > 5) return fail(SHC_E_INVALID, "bad pose reset mode");
  if (mode != e->pose_reset_mode) drop_graphs(e);
  e->pose_reset_mode = mode;
  return SHC_OK;
}
int shc_set_joint_efforts(shc_engine* e, const float* efforts_dev) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (efforts_dev != e->d_efforts) drop_graphs(e);
  e->d_efforts = efforts_dev;
  return SHC_OK;
}

int shc_set_tip_step_planes(shc_engine* e, const float* step_planes_dev) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (step_planes_dev != e->d_step_planes) drop_graphs(e);
  e->d_step_planes = step_planes_dev;
  return SHC_OK;
}

int shc_get_state(shc_engine* e, shc_robot_state* out, size_t n_records) {
  if (!e || !out || n_records != (size_t)e->n) return fail(SHC_E_INVALID, "shc_get_state: need n_robots records");
  CUDA_TRY(cudaSetDevice(e->device));
  HostPlanes h;
  int rc = download(e, h);
  if (rc != SHC_OK) return rc;
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int { unpack<decltype(dtag)::value>(e, h, out, n_records); return SHC_OK; });
}

// Records of the robots [first, first + count) only (the getters a publisher calls for one robot need not move the batch).
int shc_get_state_range(shc_engine* e, size_t first, size_t count, shc_robot_state* out) {
  if (!e || !out || count < 1 || first + count > (size_t)e->n) return fail(SHC_E_INVALID, "shc_get_state_range: range outside the batch");
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t t0 = first / 32, t1 = (first + count + 31) / 32;
  HostPlanes h;
  int rc = download_tiles(e, t0, t1, h);
  if (rc != SHC_OK) return rc;
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    unpack<decltype(dtag)::value>(e, h, out, count, first - t0 * 32);
    return SHC_OK;
  });
}

// The records of the robots [first, first + count) replaced (e.g. one robot's external swing target, its manual pose): only
// the tiles of 32 robots that hold them travel — read, re-packed with the new records, written back.
int shc_set_state_range(shc_engine* e, size_t first, size_t count, const shc_robot_state* in) {
  if (!e || !in || count < 1 || first + count > (size_t)e->n) return fail(SHC_E_INVALID, "shc_set_state_range: range outside the batch");
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t t0 = first / 32, t1 = (first + count + 31) / 32, lanes = (t1 - t0) * 32;
  HostPlanes h;
  int rc = download_tiles(e, t0, t1, h);
  if (rc != SHC_OK) return rc;
  std::vector<shc_robot_state> recs(lanes);
  rc = dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    unpack<decltype(dtag)::value>(e, h, recs.data(), lanes, 0);  // (padding lanes of the last tile hold valid robots too)
    return SHC_OK;
  });
  if (rc != SHC_OK) return rc;
  for (size_t k = 0; k < count; ++k) recs[first - t0 * 32 + k] = in[k];
  pack(e, recs.data(), lanes, h, lanes);
  return upload_tiles(e, t0, t1, h);
}

// WalkController::setLinearSpeedLimitMap / setAngularSpeedLimitMap / setLinearAccelerationLimitMap /
// setAngularAccelerationLimitMap (walk_controller.h:126-141): replaces the four limit tables (9 bearings each, 0..360 in
// steps of 45 degrees) that getLimit (walk_controller.cpp:414) reads; NULL keeps a table.  Takes effect from the next cycle.
int shc_clone_reconfigured(shc_engine* src, const shc_config* cfg, const shc_startup* startup, int flags, shc_engine** out) {
  if (!src || !cfg || !out) return fail(SHC_E_INVALID, "shc_clone_reconfigured: bad arguments");
  if (cfg->leg_count != src->cfg.leg_count || cfg->joint_count != src->cfg.joint_count)
    return fail(SHC_E_INVALID, "shc_clone_reconfigured: the new configuration must describe the same model (leg and joint counts)");
  shc_engine* e2 = nullptr;
  int rc;
  shc_startup su;
  if (flags & SHC_RECONF_KEEP_POSE_CYCLE) {
    if (startup) su = *startup;
    else if ((rc = shc_compute_startup(cfg, &su)) != SHC_OK) return rc;
    su.pose_phase_length = src->su.pose_phase_length;  // PoseController::setAutoPoseParams is not re-run (adjustParameter)
    su.pose_normaliser = src->su.pose_normaliser;
    startup = &su;
  }
  rc = shc_create(cfg, startup, src->n, src->device, src->precision, &e2);
  if (rc != SHC_OK) return rc;
  e2->options = src->options;
  e2->pose_reset_mode = src->pose_reset_mode;
  // A new step cycle (gait or step frequency) is only taken over by robots at rest: the reference re-phases walking legs
  // (LegStepper::updatePhase) and defers per robot, neither of which a batch-wide switch reproduces.
  const shc_startup &a = src->su, &b = e2->su;
  bool cycle_changed = a.period != b.period || a.swing_period != b.swing_period || a.stance_period != b.stance_period ||
                       a.stance_end != b.stance_end || a.swing_start != b.swing_start || a.swing_end != b.swing_end ||
                       a.stance_start != b.stance_start;
  for (int l = 0; l < cfg->leg_count; ++l) cycle_changed = cycle_changed || a.phase_offsets[l] != b.phase_offsets[l];
  const size_t n = (size_t)src->n, chunk = 8192;  // whole tiles of 32 robots per piece
  std::vector<shc_robot_state> buf(std::min(n, chunk));
  for (size_t first = 0; first < n && rc == SHC_OK; first += chunk) {
    const size_t count = std::min(chunk, n - first);
    rc = shc_get_state_range(src, first, count, buf.data());
    if (rc == SHC_OK && cycle_changed)
      for (size_t i = 0; i < count; ++i)
        if (buf[i].walk_state != 3 /* STOPPED */) {
          rc = fail(SHC_E_UNSUPPORTED, "shc_clone_reconfigured: the step cycle changes (gait or step frequency) and robot " +
                                           std::to_string(first + i) + " is not STOPPED: stop the batch first");
          break;
        }
    if (rc == SHC_OK) rc = shc_set_state_range(e2, first, count, buf.data());
  }
  if (rc != SHC_OK) {
    shc_destroy(e2);
    return rc;
  }
  *out = e2;
  return SHC_OK;
}

int shc_set_limit_maps(shc_engine* e, const double* max_linear_speed, const double* max_angular_speed,
                       const double* max_linear_acceleration, const double* max_angular_acceleration) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  const double* src[4] = {max_linear_speed, max_angular_speed, max_linear_acceleration, max_angular_acceleration};
  double* dst[4] = {e->su.max_linear_speed, e->su.max_angular_speed, e->su.max_linear_acceleration, e->su.max_angular_acceleration};
  for (int k = 0; k < 4; ++k) {
    if (!src[k]) continue;
    for (int b = 0; b < SHC_N_BEARINGS; ++b) {
      dst[k][b] = src[k][b];
      e->c.d.limits[k][b] = src[k][b];
      e->c.f.limits[k][b] = (float)src[k][b];
    }
  }
  drop_graphs(e);  // captured rollouts carry the constants block
  return SHC_OK;
}

int shc_set_state(shc_engine* e, const shc_robot_state* in, size_t n_records) {
  if (!e || !in || n_records != (size_t)e->n) return fail(SHC_E_INVALID, "shc_set_state: need n_robots records");
  CUDA_TRY(cudaSetDevice(e->device));
  const IntConsts& ci = e->c.i;
  HostPlanes h;
  h.s.assign((size_t)ci.nS * ci.n_pad, 0.0);
  h.d.assign((size_t)ci.nD * ci.n_pad, 0.0);
  h.i.assign((size_t)ci.nI * ci.n_pad, 0);
  pack(e, in, n_records, h, ci.n_pad);
  return upload(e, h);
}

static StepIO make_io(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual, float* joints_out) {
  StepIO io;
  io.cmd = cmd; io.imu = imu; io.tip_force = tip_force; io.manual = manual; io.efforts = e->d_efforts; io.step_planes = e->d_step_planes;
  io.joints_out = joints_out;
  io.tile_begin = 0;
  io.tile_end = (e->n + 31) / 32;
  io.n_gather = 0;
  io.gather_offset = 0;
  io.gather_mc = nullptr;
  for (auto& g : io.gather) g = nullptr;

  io.flags_out = (e->options & SHC_OPT_STATUS_FLAGS) ? e->d_flags : nullptr;
  io.pose_reset_mode = e->pose_reset_mode;
#ifdef SHC_TRACE
  io.trace = e->trace;
#endif
  return io;
}

int shc_step(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual, float* joints_out,
             void* stream) {
  if (!e || !cmd || !joints_out) return fail(SHC_E_INVALID, "shc_step: cmd and joints_out are required");
  CUDA_TRY(cudaSetDevice(e->device));
  return launch_cycle(e, make_io(e, cmd, imu, tip_force, manual, joints_out), stream ? (cudaStream_t)stream : e->stream);
}

static int ensure_staging(shc_engine* e, bool imu, bool force, bool manual) {
  const size_t n = e->n, L = e->cfg.leg_count, D = e->cfg.joint_count;
  if (!e->h_cmd) {
    CUDA_TRY(cudaMallocHost((void**)&e->h_cmd, n * 3 * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_cmd, n * 3 * 4));
    CUDA_TRY(cudaMallocHost((void**)&e->h_out, n * L * D * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_out, n * L * D * 4));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->copy, cudaStreamNonBlocking));
    for (int k = 0; k < kHostChunks; ++k) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_chunk[k], cudaEventDisableTiming));
  }
  if (imu && !e->h_imu) {
    CUDA_TRY(cudaMallocHost((void**)&e->h_imu, n * 10 * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_imu, n * 10 * 4));
  }
  if (force && !e->h_force) {
    CUDA_TRY(cudaMallocHost((void**)&e->h_force, n * L * 3 * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_force, n * L * 3 * 4));
  }
  if (manual && !e->h_manual) {
    CUDA_TRY(cudaMallocHost((void**)&e->h_manual, n * 6 * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_manual, n * 6 * 4));
  }
  return SHC_OK;
}

// Page-locked (cudaMallocHost / cudaHostRegister) host memory can be the source / target of the DMA directly.
static bool host_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// One control cycle with HOST buffers.  Inputs go up in one copy each (straight from the caller's buffer when it is
// page-locked, through the engine's pinned staging buffer otherwise); the batch is then issued as kHostChunks tile
// ranges on the engine's stream, and each range's joint angles start their way down on a second stream as soon as its
// kernel has finished, so that all but the first range's arithmetic hides behind the D2H transfer (the longer of the two).
int shc_step_host(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual,
                  float* joints_out) {
  if (!e || !cmd || !joints_out) return fail(SHC_E_INVALID, "shc_step_host: cmd and joints_out are required");
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = ensure_staging(e, imu != nullptr, tip_force != nullptr, manual != nullptr);
  if (rc != SHC_OK) return rc;
  const size_t n = e->n, L = e->cfg.leg_count, D = e->cfg.joint_count;
  cudaStream_t st = e->stream;
  auto upload = [&](const float* src, float* staging, float* dev, size_t bytes) -> int {
    const float* from = src;
    if (!host_pinned(src)) {
      std::memcpy(staging, src, bytes);
      from = staging;
    }
    CUDA_TRY(cudaMemcpyAsync(dev, from, bytes, cudaMemcpyHostToDevice, st));
    return SHC_OK;
  };
  const bool out_pinned = host_pinned(joints_out);
  // Page-locked output: the kernel's TMA bulk stores can target the caller's buffer directly (mapped host memory, same
  // address under UVA), so the joint angles cross PCIe as posted writes while the rest of the batch is still being
  // computed — one launch, no D2H copy.  SHC_HOST_ZEROCOPY=0 keeps the tile-range / copy-engine path below.
  static const bool zero_copy = [] { const char* v = getenv("SHC_HOST_ZEROCOPY"); return !(v && v[0] == '0'); }();
  // Page-locked inputs, same idea in the other direction for the SMALL records: each tile's warp reads its 32 robots'
  // velocity commands (12 B per robot; manual-pose inputs, 24 B) straight from the caller's buffer over PCIe in its
  // prologue instead of waiting behind a host-to-device copy of the whole batch: +5.7 % end to end on the hexapod shard
  // (4.98e8 -> 5.26e8 steps/s, profiles/r2ae_zero_copy_inputs.log).  The IMU and tip-force records (40 + 12 L bytes per
  // robot) stay on the copy engine: read in place they cost the octopod 7 % (SHC_HOST_ZEROCOPY_IN=2 reads everything
  // in place, =0 copies everything).
  static const int zero_copy_in = [] { const char* v = getenv("SHC_HOST_ZEROCOPY_IN"); return v ? atoi(v) : 1; }();
  float* alias = nullptr;
  const bool direct_out = out_pinned && zero_copy && cudaHostGetDevicePointer((void**)&alias, joints_out, 0) == cudaSuccess && alias;
  if (!direct_out) cudaGetLastError();
  auto input = [&](const float* src, float* staging, float* dev, size_t bytes, bool small, const float** use) -> int {
    *use = nullptr;
    if (!src) return SHC_OK;
    if (direct_out && zero_copy_in >= (small ? 1 : 2) && host_pinned(src)) {
      float* a = nullptr;
      if (cudaHostGetDevicePointer((void**)&a, const_cast<float*>(src), 0) == cudaSuccess && a) {
        *use = a;
        return SHC_OK;
      }
      cudaGetLastError();
    }
    *use = dev;
    return upload(src, staging, dev, bytes);
  };
  const float *cmd_d, *imu_d, *force_d, *manual_d;
  if ((rc = input(cmd, e->h_cmd, e->d_cmd, n * 3 * 4, true, &cmd_d)) != SHC_OK) return rc;
  if ((rc = input(imu, e->h_imu, e->d_imu, n * 10 * 4, false, &imu_d)) != SHC_OK) return rc;
  if ((rc = input(tip_force, e->h_force, e->d_force, n * L * 3 * 4, false, &force_d)) != SHC_OK) return rc;
  if ((rc = input(manual, e->h_manual, e->d_manual, n * 6 * 4, true, &manual_d)) != SHC_OK) return rc;

  if (direct_out) {
    StepIO io = make_io(e, cmd_d, imu_d, force_d, manual_d, alias);
    if ((rc = launch_cycle(e, io, st)) != SHC_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return SHC_OK;
  }
  float* down = out_pinned ? joints_out : e->h_out;
  StepIO io = make_io(e, e->d_cmd, imu ? e->d_imu : nullptr, tip_force ? e->d_force : nullptr, manual ? e->d_manual : nullptr, e->d_out);
  const int tiles = (int)((n + 31) / 32);
  const int chunks = tiles >= 64 * kHostChunks ? kHostChunks : 1;
  for (int k = 0; k < chunks; ++k) {
    io.tile_begin = (int)((long long)tiles * k / chunks);
    io.tile_end = (int)((long long)tiles * (k + 1) / chunks);
    if ((rc = launch_cycle(e, io, st)) != SHC_OK) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_chunk[k], st));
    CUDA_TRY(cudaStreamWaitEvent(e->copy, e->ev_chunk[k], 0));
    const size_t r0 = (size_t)io.tile_begin * 32, r1 = std::min((size_t)io.tile_end * 32, n);
    CUDA_TRY(cudaMemcpyAsync(down + r0 * L * D, e->d_out + r0 * L * D, (r1 - r0) * L * D * 4, cudaMemcpyDeviceToHost, e->copy));
  }
  CUDA_TRY(cudaStreamSynchronize(e->copy));
  if (!out_pinned) std::memcpy(joints_out, e->h_out, n * L * D * 4);
  return SHC_OK;
}

int shc_rollout(shc_engine* e, int k_cycles, const float* cmd_seq, const float* imu_seq, const float* force_seq, float* joints_out,
                void* stream) {
  if (!e || !cmd_seq || !joints_out || k_cycles < 1) return fail(SHC_E_INVALID, "shc_rollout: bad arguments");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  GraphKey key{k_cycles, cmd_seq, imu_seq, force_seq, joints_out};
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    const size_t n = e->n, L = e->cfg.leg_count;
    // capture on the engine's own stream (the legacy default stream cannot be captured), launch on the caller's
    cudaStream_t cap = e->stream;
    CUDA_TRY(cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
    int rc = SHC_OK;
    for (int k = 0; k < k_cycles && rc == SHC_OK; ++k)
      rc = shc_step(e, cmd_seq + (size_t)k * n * 3, imu_seq ? imu_seq + (size_t)k * n * 10 : nullptr,
                    force_seq ? force_seq + (size_t)k * n * L * 3 : nullptr, nullptr, joints_out, cap);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(cap, &graph);
    if (rc != SHC_OK || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      if (rc != SHC_OK) return rc;
      return fail(SHC_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
    }
    cudaGraphExec_t exec;
    cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(SHC_E_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ie));
    // a caller sliding a window over a long command buffer presents a new pointer set every call: keep the cache small
    if (e->graphs.size() >= kMaxCachedGraphs) drop_graphs(e);
    it = e->graphs.emplace(key, exec).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second, st));
  return SHC_OK;
}

const int* shc_status_flags_device(const shc_engine* e) {
  return (e && (e->options & SHC_OPT_STATUS_FLAGS)) ? e->d_flags : nullptr;
}

int shc_get_status_flags(shc_engine* e, int* host_out) {
  if (!e || !host_out) return fail(SHC_E_INVALID, "null argument");
  if (!(e->options & SHC_OPT_STATUS_FLAGS)) return fail(SHC_E_INVALID, "status flags are not enabled (shc_set_options)");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(host_out, e->d_flags, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
  return SHC_OK;
}

int shc_apply_ik(shc_engine* e, int n_legs, const int* leg_id, double* q, double* qd, const double* desired_tip, int simulation,
                 double* tip_out, double* ik_result, void* stream) {
  if (!e || n_legs < 1 || !leg_id || !q || !qd || !desired_tip) return fail(SHC_E_INVALID, "shc_apply_ik: bad arguments");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  const int threads = 128, blocks = (n_legs + threads - 1) / threads;
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    apply_ik_kernel<D><<<blocks, threads, 0, st>>>(e->c, n_legs, leg_id, q, qd, desired_tip, simulation, tip_out, ik_result);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("apply_ik launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
}

// Output wire formats (include/shc_msgs.h) of the robots [first, first + count) as of the last cycle, packed by one kernel.
// Outputs may be device memory or page-locked host memory (written in place over PCIe: no copy); any of them may be NULL.
// measured_joint_positions_dev (float [N][L][D], may be NULL): the measured joint states for LegState.actual_tip_pose.
int shc_pack_messages(shc_engine* e, size_t first, size_t count, const float* measured_joint_positions_dev,
                      shc_joint_state_msg* joint_state_out, shc_leg_state_msg* leg_state_out, shc_body_msg* body_out, void* stream) {
  if (!e || count < 1 || first + count > (size_t)e->n) return fail(SHC_E_INVALID, "shc_pack_messages: range outside the batch");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  auto alias = [](void* p) -> void* {  // page-locked host memory: the device-side alias of the same buffer
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return p; }
    if (a.type == cudaMemoryTypeHost) {
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess && d) return d;
      cudaGetLastError();
    }
    return p;
  };
  auto* js = (shc_joint_state_msg*)alias(joint_state_out);
  auto* legs = (shc_leg_state_msg*)alias(leg_state_out);
  auto* body = (shc_body_msg*)alias(body_out);
  const long long total = (long long)count * e->cfg.leg_count;
  const int threads = 128, blocks = (int)((total + threads - 1) / threads);
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      Planes<double> pl{(double*)e->s_planes, e->d_planes, e->i_planes};
      pack_messages_kernel<double, D><<<blocks, threads, 0, st>>>(e->c, pl, (int)first, (int)count, measured_joint_positions_dev, js, legs, body);
    } else {
      Planes<float> pl{(float*)e->s_planes, e->d_planes, e->i_planes};
      pack_messages_kernel<float, D><<<blocks, threads, 0, st>>>(e->c, pl, (int)first, (int)count, measured_joint_positions_dev, js, legs, body);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("pack_messages launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
}

// PoseController::directStartup (pose_controller.cpp:463) for the whole batch on the device.
//   shc_startup_begin  the start-up begins: every robot's origin configuration is latched (joint_positions_dev, double
//                      [N][L][D]: the measured joint states of jointStatesCallback; NULL = the default joint positions), the
//                      desired configuration of every leg is computed (test-leg IK replay, csrc/shc_startup.cuh) and the
//                      robots' stepper / walker / poser state is reset to the post-start-up state.
//   shc_startup_step   one loop() of the start-up: LegPoser::transitionConfiguration for every joint of the batch (one
//                      kernel); joints_out_dev (float [N][L][D]) gets the joint commands to publish.  Returns the reference's
//                      progress value (1..99 while moving, 100 = PROGRESS_COMPLETE: the robots are READY and their state
//                      planes hold the configuration reached), or a negative SHC_E_* code.
//   shc_direct_startup begin + the last iteration only: for callers that do not need the trajectory.
int shc_startup_begin(shc_engine* e, const double* joint_positions_dev) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = reset_to_initial_state(e);
  if (rc != SHC_OK) return rc;
  const size_t count = (size_t)e->n * e->cfg.leg_count * e->cfg.joint_count;
  if (joint_positions_dev) {
    if (!e->startup_origin) CUDA_TRY(cudaMalloc((void**)&e->startup_origin, count * 8));
    CUDA_TRY(cudaMemcpy(e->startup_origin, joint_positions_dev, count * 8, cudaMemcpyDeviceToDevice));
    e->startup_has_origin = true;
  } else {
    e->startup_has_origin = false;
  }
  double desired[kMaxLegs][kMaxDof] = {};
  const StartupParams sp = startup_params(e->cfg);
  dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    for (int l = 0; l < e->cfg.leg_count; ++l) startup_desired_configuration<decltype(dtag)::value>(e->c.d, sp, l, desired[l]);
    return SHC_OK;
  });
  if (!e->startup_desired) CUDA_TRY(cudaMalloc((void**)&e->startup_desired, sizeof(desired)));
  CUDA_TRY(cudaMemcpy(e->startup_desired, desired, sizeof(desired), cudaMemcpyHostToDevice));
  e->startup_iteration = 0;
  e->startup_num = sp.startup_iterations;
  return SHC_OK;
}

static int startup_launch(shc_engine* e, int it, float* joints_out_dev, cudaStream_t st) {
  const long long total = (long long)e->n * e->cfg.leg_count * e->cfg.joint_count;
  const int threads = 256, blocks = (int)((total + threads - 1) / threads);
  const double* origin = e->startup_has_origin ? e->startup_origin : nullptr;
  const int last = it >= e->startup_num;
  return dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      Planes<double> pl{(double*)e->s_planes, e->d_planes, e->i_planes};
      transition_configuration_kernel<double, D><<<blocks, threads, 0, st>>>(e->c, pl, origin, e->startup_desired, it, e->startup_num, joints_out_dev, last);
    } else {
      Planes<float> pl{(float*)e->s_planes, e->d_planes, e->i_planes};
      transition_configuration_kernel<float, D><<<blocks, threads, 0, st>>>(e->c, pl, origin, e->startup_desired, it, e->startup_num, joints_out_dev, last);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("transition_configuration launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
}

int shc_startup_step(shc_engine* e, float* joints_out_dev, void* stream) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (e->startup_num <= 0) return fail(SHC_E_INVALID, "shc_startup_begin has not been called");
  if (e->startup_iteration >= e->startup_num) return 100;  // PROGRESS_COMPLETE
  CUDA_TRY(cudaSetDevice(e->device));
  const int it = ++e->startup_iteration;
  int rc = startup_launch(e, it, joints_out_dev, stream ? (cudaStream_t)stream : e->stream);
  if (rc != SHC_OK) return rc;
  if (it >= e->startup_num) return 100;
  const int progress = int((double(it - 1) / double(e->startup_num)) * 100);
  return progress < 1 ? 1 : progress;
}

int shc_direct_startup(shc_engine* e, const double* joint_positions_dev, float* joints_out_dev, void* stream) {
  int rc = shc_startup_begin(e, joint_positions_dev);
  if (rc != SHC_OK) return rc;
  e->startup_iteration = e->startup_num;
  return startup_launch(e, e->startup_num, joints_out_dev, stream ? (cudaStream_t)stream : e->stream);
}

// ---- stepping and joint-space sequences (device routines in csrc/shc_sequence.cuh) -----------------
static int seq_alloc(shc_engine* e) {
  if (e->seq_origin) return SHC_OK;
  const int L = e->cfg.leg_count;
  const size_t np = e->n_pad;
  CUDA_TRY(cudaMalloc((void**)&e->seq_origin, seq_origin_count(L, np) * 8));
  CUDA_TRY(cudaMalloc((void**)&e->seq_count, seq_count_count(L, np) * 4));
  CUDA_TRY(cudaMalloc((void**)&e->seq_robot, seq_robot_count(np) * 4));
  CUDA_TRY(cudaMalloc((void**)&e->seq_min, 4));
  return shc_sequence_reset(e);
}

// Forgets any stepping sequence in progress: every LegPoser back to first_iteration_, group 0, no completed steps.
int shc_sequence_reset(shc_engine* e) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  CUDA_TRY(cudaSetDevice(e->device));
  if (!e->seq_origin) return seq_alloc(e);
  const size_t nc = seq_count_count(e->cfg.leg_count, e->n_pad);
  fill_int_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, e->stream>>>(e->seq_count, nc, -1);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(e->seq_robot, 0, seq_robot_count(e->n_pad) * 4, e->stream));
  fill_int_kernel<<<(unsigned)((e->n_pad + 255) / 256), 256, 0, e->stream>>>(e->seq_robot + 4 * (size_t)e->n_pad, (size_t)e->n_pad, kSeqInitialFlags);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(e->seq_origin, 0, seq_origin_count(e->cfg.leg_count, e->n_pad) * 8, e->stream));
  if (e->seq_leg) CUDA_TRY(cudaMemsetAsync(e->seq_leg, 0, seq_leg_count(e->cfg.leg_count, e->n_pad) * 4, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SHC_OK;
}

// One loop() of PoseController::stepToNewStance (pose_controller.cpp:520) for every robot of the batch: the legs of each
// robot's current group step to their default tip positions (LegPoser::stepToPosition :1571 with the swing height as lift
// and one step period as duration, under the robot's current body pose), followed by Leg::applyIK; the groups alternate as
// each robot's own legs complete.  joints_out_dev float [N][L][D] or NULL; progress_out_dev int [N] or NULL (each robot's
// return value).  Returns the smallest progress over the batch (>= 0; blocking on `stream`) or a negative SHC_E_* code.
int shc_step_to_new_stance(shc_engine* e, float* joints_out_dev, int* progress_out_dev, void* stream) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (e->c.i.tip_mode == TIP_ROTATION) return fail(SHC_E_UNSUPPORTED, "stepping sequences with tip-rotation targets (gravity_aligned_tips, D > 3) are not built");
  if (e->cfg.leg_count % 2 != 0) return fail(SHC_E_UNSUPPORTED, "stepToNewStance coordinates two leg groups: leg_count must be even");
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = seq_alloc(e);
  if (rc != SHC_OK) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  NewStanceParams np;
  np.lift_height = e->cfg.swing_height;
  np.num_iterations = std::max(1, round_to_int((1.0 / e->cfg.step_frequency) / e->cfg.time_delta));
  np.apply_delta = 1;
  SeqBuffers sq{e->seq_origin, e->seq_count, e->seq_robot, (size_t)e->n_pad, e->seq_target, e->seq_poses, e->seq_leg};
  const int init = 1000;
  CUDA_TRY(cudaMemcpyAsync(e->seq_min, &init, 4, cudaMemcpyHostToDevice, st));
  const int threads = 128, blocks = (e->n + threads - 1) / threads;
  rc = dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      Planes<double> pl{(double*)e->s_planes, e->d_planes, e->i_planes};
      step_to_new_stance_kernel<double, D><<<blocks, threads, 0, st>>>(e->c, pl, sq, np, joints_out_dev, progress_out_dev, e->seq_min);
    } else {
      Planes<float> pl{(float*)e->s_planes, e->d_planes, e->i_planes};
      step_to_new_stance_kernel<float, D><<<blocks, threads, 0, st>>>(e->c, pl, sq, np, joints_out_dev, progress_out_dev, e->seq_min);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("step_to_new_stance launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
  if (rc != SHC_OK) return rc;
  int progress = 0;
  CUDA_TRY(cudaMemcpyAsync(&progress, e->seq_min, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return progress;
}

// One loop() of PoseController::executeSequence (pose_controller.cpp:145-461) for every robot: the start-up (shut_down = 0)
// or shut-down sequence, each robot on its own course (csrc/shc_sequence.cuh).  joints_out_dev float [N][L][D] /
// progress_out_dev int [N] (each robot's return value: -1 while its first start-up generates the sequence, 0..100, -2 once
// it has failed: SHC_SEQ_GENERATING / SHC_SEQ_FAILED), either may be NULL.  *min_progress_out = the smallest value over the
// batch (blocking on `stream`); the return value is SHC_OK or an SHC_E_* code.
int shc_execute_sequence(shc_engine* e, int shut_down, float* joints_out_dev, int* progress_out_dev, int* min_progress_out, void* stream) {
  if (!e || !min_progress_out) return fail(SHC_E_INVALID, "shc_execute_sequence: bad arguments");
  if (e->c.i.tip_mode == TIP_ROTATION) return fail(SHC_E_UNSUPPORTED, "stepping sequences with tip-rotation targets (gravity_aligned_tips, D > 3) are not built");
  if (e->cfg.leg_count % 2 != 0) return fail(SHC_E_UNSUPPORTED, "executeSequence coordinates two leg groups: leg_count must be even");
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = seq_alloc(e);
  if (rc != SHC_OK) return rc;
  const int L = e->cfg.leg_count;
  if (!e->seq_poses) {
    CUDA_TRY(cudaMalloc((void**)&e->seq_target, seq_origin_count(L, e->n_pad) * 8));
    CUDA_TRY(cudaMalloc((void**)&e->seq_poses, seq_poses_count(L, e->n_pad) * 8));
    CUDA_TRY(cudaMalloc((void**)&e->seq_leg, seq_leg_count(L, e->n_pad) * 4));
    CUDA_TRY(cudaMemset(e->seq_target, 0, seq_origin_count(L, e->n_pad) * 8));
    CUDA_TRY(cudaMemset(e->seq_poses, 0, seq_poses_count(L, e->n_pad) * 8));
    CUDA_TRY(cudaMemset(e->seq_leg, 0, seq_leg_count(L, e->n_pad) * 4));
  }
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  ExecuteSequenceParams ep{e->cfg.swing_height, e->cfg.step_frequency, e->cfg.time_delta};
  SeqBuffers sq{e->seq_origin, e->seq_count, e->seq_robot, (size_t)e->n_pad, e->seq_target, e->seq_poses, e->seq_leg};
  const int init = 1000;
  CUDA_TRY(cudaMemcpyAsync(e->seq_min, &init, 4, cudaMemcpyHostToDevice, st));
  const int threads = 128, blocks = (e->n + threads - 1) / threads;
  rc = dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      Planes<double> pl{(double*)e->s_planes, e->d_planes, e->i_planes};
      execute_sequence_kernel<double, D><<<blocks, threads, 0, st>>>(e->c, pl, sq, ep, shut_down, joints_out_dev, progress_out_dev, e->seq_min);
    } else {
      Planes<float> pl{(float*)e->s_planes, e->d_planes, e->i_planes};
      execute_sequence_kernel<float, D><<<blocks, threads, 0, st>>>(e->c, pl, sq, ep, shut_down, joints_out_dev, progress_out_dev, e->seq_min);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("execute_sequence launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
  if (rc != SHC_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(min_progress_out, e->seq_min, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return SHC_OK;
}

// PoseController::transitionConfiguration (pose_controller.cpp:703) for the whole batch: from the joint positions the state
// planes hold to desired_configuration (host doubles [L][D]) along LegPoser::transitionConfiguration's cubic Bezier (:1476)
// in max(1, roundToInt(transition_time / time_delta)) loops.
int shc_transition_begin(shc_engine* e, const double* desired_configuration, double transition_time) {
  if (!e || !desired_configuration || !(transition_time >= 0.0)) return fail(SHC_E_INVALID, "shc_transition_begin: bad arguments");
  CUDA_TRY(cudaSetDevice(e->device));
  const int L = e->cfg.leg_count, D = e->cfg.joint_count;
  const long long total = (long long)e->n * L * D;
  if (!e->tr_origin) CUDA_TRY(cudaMalloc((void**)&e->tr_origin, (size_t)total * 8));
  if (!e->tr_desired) CUDA_TRY(cudaMalloc((void**)&e->tr_desired, sizeof(double) * kMaxLegs * kMaxDof));
  double desired[kMaxLegs][kMaxDof] = {};
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < D; ++j) desired[l][j] = desired_configuration[l * D + j];
  CUDA_TRY(cudaMemcpyAsync(e->tr_desired, desired, sizeof(desired), cudaMemcpyHostToDevice, e->stream));
  const int threads = 256, blocks = (int)((total + threads - 1) / threads);
  int rc = dispatch_D(D, [&](auto dtag) -> int {
    constexpr int DD = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) latch_joint_kernel<double, DD><<<blocks, threads, 0, e->stream>>>(e->c, Planes<double>{(double*)e->s_planes, e->d_planes, e->i_planes}, e->tr_origin);
    else latch_joint_kernel<float, DD><<<blocks, threads, 0, e->stream>>>(e->c, Planes<float>{(float*)e->s_planes, e->d_planes, e->i_planes}, e->tr_origin);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("latch_joint launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
  if (rc != SHC_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));  // `desired` is a stack array
  e->tr_iteration = 0;
  e->tr_num = std::max(1, round_to_int(transition_time / e->cfg.time_delta));
  return SHC_OK;
}

// One loop() of the transition: every joint of the batch moves one iteration on (one kernel); joints_out_dev (float
// [N][L][D]) gets the joint commands.  Returns the reference's progress (1..99, 100 = PROGRESS_COMPLETE) or a negative code.
int shc_transition_step(shc_engine* e, float* joints_out_dev, void* stream) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (e->tr_num <= 0) return fail(SHC_E_INVALID, "shc_transition_begin has not been called");
  if (e->tr_iteration >= e->tr_num) return 100;
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  const int it = ++e->tr_iteration;
  const long long total = (long long)e->n * e->cfg.leg_count * e->cfg.joint_count;
  const int threads = 256, blocks = (int)((total + threads - 1) / threads);
  int rc = dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64)
      transition_joint_kernel<double, D><<<blocks, threads, 0, st>>>(e->c, Planes<double>{(double*)e->s_planes, e->d_planes, e->i_planes}, e->tr_origin, e->tr_desired, it, e->tr_num, joints_out_dev);
    else
      transition_joint_kernel<float, D><<<blocks, threads, 0, st>>>(e->c, Planes<float>{(float*)e->s_planes, e->d_planes, e->i_planes}, e->tr_origin, e->tr_desired, it, e->tr_num, joints_out_dev);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("transition_joint launch: ") + cudaGetErrorString(err));
    return SHC_OK;
  });
  if (rc != SHC_OK) return rc;
  if (it >= e->tr_num) return 100;
  const int progress = int((double(it - 1) / double(e->tr_num)) * 100);
  return progress < 1 ? 1 : progress;
}

// One loop() of PoseController::packLegs (pose_controller.cpp:597) / unpackLegs (:661): the first call of a sequence latches
// the batch's joint positions and targets the packed / unpacked joint positions of the configuration (one pack step),
// every call moves the transition one iteration on.  Returns the progress (100 = complete, the next call starts anew).
static int pack_unpack(shc_engine* e, bool pack, double time, float* joints_out_dev, void* stream) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  if (pack && e->seq_robot)  // packLegs: transition_step_ = 0 (pose_controller.cpp:618)
    CUDA_TRY(cudaMemsetAsync(e->seq_robot + 2 * (size_t)e->n_pad, 0, (size_t)e->n_pad * 4, stream ? (cudaStream_t)stream : e->stream));
  if (!e->tr_executing) {
    double desired[kMaxLegs * kMaxDof];
    const int L = e->cfg.leg_count, D = e->cfg.joint_count;
    for (int l = 0; l < L; ++l)
      for (int j = 0; j < D; ++j) desired[l * D + j] = pack ? e->cfg.joint_packed[l][j] : e->cfg.joint_unpacked[l][j];
    int rc = shc_transition_begin(e, desired, time);
    if (rc != SHC_OK) return rc;
  }
  const int progress = shc_transition_step(e, joints_out_dev, stream);
  if (progress < 0) return progress;
  e->tr_executing = progress != 0 && progress != 100;
  return progress;
}
int shc_pack_legs(shc_engine* e, double time_to_pack, float* joints_out_dev, void* stream) { return pack_unpack(e, true, time_to_pack, joints_out_dev, stream); }
int shc_unpack_legs(shc_engine* e, double time_to_unpack, float* joints_out_dev, void* stream) { return pack_unpack(e, false, time_to_unpack, joints_out_dev, stream); }

// Host-buffer form of one loop() of a sequence, for callers without device memory of their own (the C++ facade): kind
// SHC_SEQ_NEW_STANCE / SHC_SEQ_PACK / SHC_SEQ_UNPACK / SHC_SEQ_DIRECT_STARTUP (the direct start-up begins on its first call
// from the default joint positions); joints_out [N][L][D] and progress_out [N] are HOST arrays (either may be NULL).
// Returns the smallest progress over the batch or a negative SHC_E_* code.
int shc_sequence_step_host(shc_engine* e, int kind, double time, float* joints_out, int* progress_out) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  int rc = ensure_staging(e, false, false, false);
  if (rc != SHC_OK) return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t n = e->n, L = e->cfg.leg_count, D = e->cfg.joint_count;
  int progress;
  int* d_progress = nullptr;
  if (kind == SHC_SEQ_NEW_STANCE) {
    if (progress_out) CUDA_TRY(cudaMalloc((void**)&d_progress, n * 4));
    progress = shc_step_to_new_stance(e, e->d_out, d_progress, e->stream);
  } else if (kind == SHC_SEQ_PACK) {
    progress = shc_pack_legs(e, time, e->d_out, e->stream);
  } else if (kind == SHC_SEQ_UNPACK) {
    progress = shc_unpack_legs(e, time, e->d_out, e->stream);
  } else if (kind == SHC_SEQ_START_UP || kind == SHC_SEQ_SHUT_DOWN) {
    if (progress_out) CUDA_TRY(cudaMalloc((void**)&d_progress, n * 4));
    rc = shc_execute_sequence(e, kind == SHC_SEQ_SHUT_DOWN, e->d_out, d_progress, &progress, e->stream);
    if (rc != SHC_OK) { cudaFree(d_progress); return rc; }
  } else if (kind == SHC_SEQ_DIRECT_STARTUP) {
    if (e->startup_num <= 0 || e->startup_iteration >= e->startup_num) {
      if ((rc = shc_startup_begin(e, nullptr)) != SHC_OK) return rc;
    }
    progress = shc_startup_step(e, e->d_out, e->stream);
  } else {
    return fail(SHC_E_INVALID, "shc_sequence_step_host: unknown sequence");
  }
  const bool exec_seq = kind == SHC_SEQ_START_UP || kind == SHC_SEQ_SHUT_DOWN;  // (their -1 / -2 are progress values)
  if (progress < 0 && !exec_seq) { cudaFree(d_progress); return progress; }
  if (joints_out) CUDA_TRY(cudaMemcpyAsync(joints_out, e->d_out, n * L * D * 4, cudaMemcpyDeviceToHost, e->stream));
  if (progress_out) {
    if (d_progress) CUDA_TRY(cudaMemcpyAsync(progress_out, d_progress, n * 4, cudaMemcpyDeviceToHost, e->stream));
    else for (size_t r = 0; r < n; ++r) progress_out[r] = progress;
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  cudaFree(d_progress);
  return progress;
}

// Model::generateWorkspaces (model.cpp:120) / Leg::generateWorkspace (:309-510) on the device: one block per leg, the eight
// bearing searches of a workplane (up to 500 Leg::applyIK(true) steps each) on eight lanes.  full = 0: the simple workspace
// (one plane at height 0, what the engine's limit tables are built from); full = 1: the layered workspace of rough-terrain
// mode.  Host outputs: heights [L][max_planes], radii [L][max_planes][9] (bearings 0..360 step 45), n_planes [L] (planes
// generated; more than max_planes means the tail was dropped).  The legs start from the engine's default configuration.
int shc_generate_workspaces(shc_engine* e, int full, int max_planes, double* heights_out, double* radii_out, int* n_planes_out) {
  if (!e || max_planes < 1 || !heights_out || !radii_out || !n_planes_out) return fail(SHC_E_INVALID, "shc_generate_workspaces: bad arguments");
  CUDA_TRY(cudaSetDevice(e->device));
  const int L = e->cfg.leg_count;
  double *d_q = nullptr, *d_h = nullptr, *d_r = nullptr;
  int* d_n = nullptr;
  const size_t nh = (size_t)L * max_planes, nr = nh * SHC_N_BEARINGS;
  auto cleanup = [&] { cudaFree(d_q); cudaFree(d_h); cudaFree(d_r); cudaFree(d_n); };
  if (cudaMalloc((void**)&d_q, sizeof(e->su.default_joint)) != cudaSuccess || cudaMalloc((void**)&d_h, nh * 8) != cudaSuccess ||
      cudaMalloc((void**)&d_r, nr * 8) != cudaSuccess || cudaMalloc((void**)&d_n, L * 4) != cudaSuccess) {
    cleanup();
    return fail(SHC_E_CUDA, "shc_generate_workspaces: device allocation failed");
  }
  cudaMemcpy(d_q, e->su.default_joint, sizeof(e->su.default_joint), cudaMemcpyHostToDevice);
  cudaMemset(d_h, 0, nh * 8);
  cudaMemset(d_r, 0, nr * 8);
  const StartupParams sp = startup_params(e->cfg);
  int rc = dispatch_D(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    workspace_sweep_kernel<D><<<L, 32, 0, e->stream>>>(e->c, sp, d_q, full, max_planes, d_h, d_r, d_n);
    return SHC_OK;
  });
  cudaError_t err = cudaStreamSynchronize(e->stream);
  if (err == cudaSuccess) err = cudaMemcpy(heights_out, d_h, nh * 8, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess) err = cudaMemcpy(radii_out, d_r, nr * 8, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess) err = cudaMemcpy(n_planes_out, d_n, L * 4, cudaMemcpyDeviceToHost);
  cleanup();
  if (err != cudaSuccess) return fail(SHC_E_CUDA, std::string("shc_generate_workspaces: ") + cudaGetErrorString(err));
  return rc;
}

// The same sweep on the host (double, the very routine the kernel runs: csrc/shc_startup.cuh), for callers without a device
// at hand and for the CPU tests.  `startup` NULL: the engine's own default configuration for `cfg`.
int shc_host_generate_workspaces(const shc_config* cfg, const shc_startup* startup, int full, int max_planes, double* heights_out,
                                 double* radii_out, int* n_planes_out) {
  if (!cfg || max_planes < 1 || !heights_out || !radii_out || !n_planes_out) return fail(SHC_E_INVALID, "shc_host_generate_workspaces: bad arguments");
  std::string err;
  bool unsupported = false;
  if (!check_supported(*cfg, err, unsupported)) return fail(unsupported ? SHC_E_UNSUPPORTED : SHC_E_INVALID, err);
  shc_startup su;
  if (startup) su = *startup;
  else {
    int rc = shc_compute_startup(cfg, &su);
    if (rc != SHC_OK) return rc;
  }
  RealConsts<double>* ck = new RealConsts<double>();
  fill_static_consts<double>(*cfg, *ck);
  const StartupParams sp = startup_params(*cfg);
  int rc = dispatch_D(cfg->joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    for (int l = 0; l < cfg->leg_count; ++l) {
      for (size_t k = 0; k < (size_t)max_planes; ++k) heights_out[(size_t)l * max_planes + k] = 0.0;
      for (size_t k = 0; k < (size_t)max_planes * SHC_N_BEARINGS; ++k) radii_out[(size_t)l * max_planes * SHC_N_BEARINGS + k] = 0.0;
      n_planes_out[l] = workspace_sweep_leg<D>(*ck, sp, l, su.default_joint[l], full != 0, max_planes, 0, 1, heights_out + (size_t)l * max_planes,
                                               radii_out + (size_t)l * max_planes * SHC_N_BEARINGS);
    }
    return SHC_OK;
  });
  delete ck;
  return rc;
}

int shc_nccl_unique_id(void* out128) {
  if (!out128) return fail(SHC_E_INVALID, "null argument");
  if (!nccl().ok) return fail(SHC_E_UNSUPPORTED, "libnccl.so.2 is not available in this process");
  int rc = nccl().GetUniqueId((NcclUid*)out128);
  if (rc != 0) return fail(SHC_E_CUDA, std::string("ncclGetUniqueId: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "error"));
  return SHC_OK;
}

int shc_nccl_init(shc_engine* e, const void* uid128, int rank, int world_size) {
  if (!e || !uid128 || rank < 0 || rank >= world_size) return fail(SHC_E_INVALID, "shc_nccl_init: bad arguments");
  if (!nccl().ok) return fail(SHC_E_UNSUPPORTED, "libnccl.so.2 is not available in this process");
  CUDA_TRY(cudaSetDevice(e->device));
  NcclUid uid;
  std::memcpy(&uid, uid128, sizeof(uid));
  int rc = nccl().CommInitRank(&e->nccl_comm, world_size, uid, rank);
  if (rc != 0) return fail(SHC_E_CUDA, std::string("ncclCommInitRank: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "error"));
  e->rank = rank;
  e->world = world_size;
  CUDA_TRY(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_ready[b], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_done[b], cudaEventDisableTiming));
  }
  return SHC_OK;
}

int shc_allgather_joints(shc_engine* e, const float* local, float* full, void* stream) {
  if (!e || !local || !full) return fail(SHC_E_INVALID, "shc_allgather_joints: bad arguments");
  if (!e->nccl_comm) return fail(SHC_E_INVALID, "shc_nccl_init has not been called");
  CUDA_TRY(cudaSetDevice(e->device));
  size_t count = (size_t)e->n * e->cfg.leg_count * e->cfg.joint_count;
  int rc = nccl().AllGather(local, full, count, /*ncclFloat32*/ 7, e->nccl_comm, stream ? (cudaStream_t)stream : e->side);
  if (rc != 0) return fail(SHC_E_CUDA, std::string("ncclAllGather: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "error"));
  return SHC_OK;
}

int shc_rollout_allgather(shc_engine* e, int k_cycles, const float* cmd_seq, float* local2, float* full2, void* stream) {
  if (!e || !cmd_seq || !local2 || !full2 || k_cycles < 1) return fail(SHC_E_INVALID, "shc_rollout_allgather: bad arguments");
  if (!e->nccl_comm) return fail(SHC_E_INVALID, "shc_nccl_init has not been called");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  const size_t n = e->n, per_rank = n * e->cfg.leg_count * e->cfg.joint_count;
  for (int k = 0; k < k_cycles; ++k) {
    const int b = k & 1;
    if (e->done_valid[b]) CUDA_TRY(cudaStreamWaitEvent(st, e->ev_done[b], 0));  // buffer b's previous gather has drained
    int rc = shc_step(e, cmd_seq + (size_t)k * n * 3, nullptr, nullptr, nullptr, local2 + b * per_rank, st);
    if (rc != SHC_OK) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_ready[b], st));
    CUDA_TRY(cudaStreamWaitEvent(e->side, e->ev_ready[b], 0));
    rc = shc_allgather_joints(e, local2 + b * per_rank, full2 + (size_t)b * per_rank * e->world, e->side);
    if (rc != SHC_OK) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_done[b], e->side));
    e->done_valid[b] = true;
  }
  // join: the caller's stream continues only after the last gathers
  for (int b = 0; b < 2; ++b)
    if (e->done_valid[b]) CUDA_TRY(cudaStreamWaitEvent(st, e->ev_done[b], 0));
  return SHC_OK;
}

// ---- fused all-gather over peer memory (NVLink / NVSwitch) ---------------------------------------------------------------
// The control-cycle kernel stores each finished tile of joint commands into ALL ranks' gather buffers.  Two ways to get the
// buffers mapped:
//   * shc_gather_attach: the caller owns symmetric buffers (one per rank, same size) and passes every rank's mapping of
//     them plus, when the fabric offers it, their NVSwitch MULTICAST mapping — then one multimem.st per 16 bytes leaves
//     this GPU and the switch replicates it (egress = the shard, not world-1 copies of it);
//   * shc_gather_alloc / shc_gather_open_peer: the engines cudaMalloc the buffers and exchange CUDA-IPC handles; the
//     kernel then issues one TMA bulk store per peer (unicast).
// cuStreamWriteValue32 of the driver already loaded in the process (no link-time dependency on libcuda)
typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
static StreamWriteValue32Fn stream_write_value32() {
  static StreamWriteValue32Fn fn = [] {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libcuda.so.1", RTLD_NOW);
    StreamWriteValue32Fn f = h ? (StreamWriteValue32Fn)dlsym(h, "cuStreamWriteValue32_v2") : nullptr;
    if (!f && h) f = (StreamWriteValue32Fn)dlsym(h, "cuStreamWriteValue32");
    return f;
  }();
  return fn;
}
static size_t gather_data_bytes(const shc_engine* e, int world) {
  const size_t per_rank = (size_t)e->n * e->cfg.leg_count * e->cfg.joint_count;
  return ((size_t)kGatherBuffers * world * per_rank * 4 + 255) / 256 * 256;
}
static int* gather_flags_of(const shc_engine* e, int p) {
  return reinterpret_cast<int*>(reinterpret_cast<char*>(e->gather_peer[p]) + gather_data_bytes(e, e->world));
}
static int gather_common_init(shc_engine* e) {
  if (!e->signal) {
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&e->signal, cudaStreamNonBlocking, hi));
  }
  for (int b = 0; b < kGatherBuffers; ++b)
    if (!e->ev_kernel[b]) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_kernel[b], cudaEventDisableTiming));
  if (!e->gather_err) {
    CUDA_TRY(cudaHostAlloc((void**)&e->gather_err, 4, cudaHostAllocMapped));
    *e->gather_err = 0;
  }
  e->gather_cycle = 0;
  e->gather_waited = 0;
  e->gather_signalled = 0;
  // tuning switches: SHC_GATHER_SIGNAL_EVERY=k (1 <= k <= kGatherWaitEvery), SHC_GATHER_SIGNAL=memop
  if (const char* v = getenv("SHC_GATHER_SIGNAL_EVERY")) e->gather_signal_every = std::min(std::max(atoi(v), 1), kGatherWaitEvery);
  if (const char* v = getenv("SHC_GATHER_SIGNAL")) e->gather_signal_memop = !strcmp(v, "memop") && stream_write_value32() != nullptr;
  if (getenv("SHC_GATHER_TRACE") && !e->gather_trace) {
    CUDA_TRY(cudaHostAlloc((void**)&e->gather_trace, 4096 * 4 * 8, cudaHostAllocMapped));
    std::memset(e->gather_trace, 0, 4096 * 4 * 8);
  }
  return SHC_OK;
}

size_t shc_gather_bytes(const shc_engine* e, int world_size) {
  if (!e || world_size < 1) return 0;
  return gather_data_bytes(e, world_size) + 256;
}
int shc_gather_buffers(void) { return kGatherBuffers; }

int shc_gather_attach(shc_engine* e, int rank, int world_size, void* const* peer_buffers, void* multicast_buffer) {
  if (!e || !peer_buffers || rank < 0 || rank >= world_size) return fail(SHC_E_INVALID, "shc_gather_attach: bad arguments");
  if (world_size > 8) return fail(SHC_E_UNSUPPORTED, "fused gather supports up to 8 ranks (one node)");
  if (e->gather_own) return fail(SHC_E_INVALID, "a gather buffer is already attached");
  for (int p = 0; p < world_size; ++p)
    if (!peer_buffers[p]) return fail(SHC_E_INVALID, "shc_gather_attach: null peer buffer");
  CUDA_TRY(cudaSetDevice(e->device));
  e->rank = rank;
  e->world = world_size;
  for (int p = 0; p < world_size; ++p) e->gather_peer[p] = (float*)peer_buffers[p];
  e->gather_own = e->gather_peer[rank];
  e->gather_mc = (float*)multicast_buffer;
  e->gather_owned = false;
  // landed counters start at zero; the caller runs a barrier between this call and the first cycle
  CUDA_TRY(cudaMemset(reinterpret_cast<char*>(e->gather_own) + gather_data_bytes(e, world_size), 0, 256));
  CUDA_TRY(cudaDeviceSynchronize());
  return gather_common_init(e);
}

int shc_gather_alloc(shc_engine* e, void* handle64_out, float** buffer_out) {
  if (!e || !handle64_out) return fail(SHC_E_INVALID, "shc_gather_alloc: bad arguments");
  if (!e->nccl_comm) return fail(SHC_E_INVALID, "shc_nccl_init has not been called");
  if (e->world > 8) return fail(SHC_E_UNSUPPORTED, "fused gather supports up to 8 ranks (one node)");
  CUDA_TRY(cudaSetDevice(e->device));
  if (!e->gather_own) {
    const size_t bytes = gather_data_bytes(e, e->world) + 256;
    CUDA_TRY(cudaMalloc((void**)&e->gather_own, bytes));
    e->gather_owned = true;
    CUDA_TRY(cudaMemset(e->gather_own, 0, bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = gather_common_init(e);
    if (rc != SHC_OK) return rc;
  }
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, e->gather_own));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  std::memcpy(handle64_out, &h, 64);
  e->gather_peer[e->rank] = e->gather_own;
  if (buffer_out) *buffer_out = e->gather_own;
  return SHC_OK;
}

int shc_gather_open_peer(shc_engine* e, int peer_rank, const void* handle64) {
  if (!e || !handle64 || peer_rank < 0 || peer_rank >= e->world || peer_rank >= 8) return fail(SHC_E_INVALID, "shc_gather_open_peer: bad arguments");
  if (peer_rank == e->rank) return SHC_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  e->gather_peer[peer_rank] = (float*)p;
  e->gather_opened[peer_rank] = true;
  return SHC_OK;
}

static int gather_ready(shc_engine* e, const char* who) {
  if (!e->gather_own) return fail(SHC_E_INVALID, std::string(who) + ": no gather buffer (shc_gather_attach / shc_gather_alloc)");
  for (int p = 0; p < e->world; ++p)
    if (!e->gather_peer[p]) return fail(SHC_E_INVALID, std::string(who) + ": a peer's gather buffer is not mapped");
  return SHC_OK;
}

// The stream waits until every source rank has landed `need` cycles in this rank's buffer.
static int gather_wait(shc_engine* e, long long need, cudaStream_t st) {
  if (need <= e->gather_waited) return SHC_OK;
  WaitArgs wa;
  wa.flags = gather_flags_of(e, e->rank);
  wa.n = e->world;
  for (int p = 0; p < 8; ++p) wa.need[p] = (int)need;
  wa.err = nullptr;
  if (cudaHostGetDevicePointer((void**)&wa.err, e->gather_err, 0) != cudaSuccess) { cudaGetLastError(); wa.err = nullptr; }
  wa.trace = nullptr;
  if (e->gather_trace && e->gather_trace_waits < 2048) {
    e->gather_trace[(2048 + e->gather_trace_waits) * 4 + 2] = (unsigned long long)need;
    wa.trace = e->gather_trace + (2048 + e->gather_trace_waits++) * 4;
  }
  gather_wait_kernel<<<1, 32, 0, st>>>(wa);
  CUDA_TRY(cudaGetLastError());
  e->gather_waited = need;
  return SHC_OK;
}

// "Every cycle issued so far has landed everywhere": an event behind the last kernel, then on the high-priority side
// stream either a one-warp kernel whose system-scope release waits for the posted NVLink writes to drain and bumps this
// rank's counter on every rank (through the multicast address when there is one: the path the data took), or
// (SHC_GATHER_SIGNAL=memop) one stream memory operation per rank that writes the new count once the kernel has completed.
// Sparse inside a call (every gather_signal_every cycles: the release holds up the SM it runs on while the next cycle's
// warps on that SM keep storing) and always issued by shc_gather_sync.
static int gather_signal(shc_engine* e, cudaStream_t st) {
  const long long upto = e->gather_cycle;
  if (upto <= e->gather_signalled) return SHC_OK;
  const int slot = (int)((upto - 1) % kGatherBuffers);
  CUDA_TRY(cudaEventRecord(e->ev_kernel[slot], st));
  CUDA_TRY(cudaStreamWaitEvent(e->signal, e->ev_kernel[slot], 0));
  if (e->gather_signal_memop) {
    for (int p = 0; p < e->world; ++p) {
      const int rc = stream_write_value32()(e->signal, (unsigned long long)(uintptr_t)(gather_flags_of(e, p) + e->rank), (unsigned)upto, 0);
      if (rc != 0) return fail(SHC_E_CUDA, "cuStreamWriteValue32 failed (" + std::to_string(rc) + ")");
    }
  } else {
    SignalArgs sa;
    sa.trace = (e->gather_trace && upto - 1 < 2048) ? e->gather_trace + (upto - 1) * 4 : nullptr;
    sa.n = 0;
    sa.count = (int)(upto - e->gather_signalled);
    sa.mc_flag = nullptr;
    if (e->gather_mc) {
      sa.mc_flag = reinterpret_cast<int*>(reinterpret_cast<char*>(e->gather_mc) + gather_data_bytes(e, e->world)) + e->rank;
    } else {
      for (int p = 0; p < e->world; ++p) sa.flag[sa.n++] = gather_flags_of(e, p) + e->rank;  // own counter included
    }
    gather_signal_kernel<<<1, 32, 0, e->signal>>>(sa);
    CUDA_TRY(cudaGetLastError());
  }
  e->gather_signalled = upto;
  return SHC_OK;
}

// One control cycle whose joint commands land in buffer (cycle % shc_gather_buffers()) of EVERY rank, stored from inside
// the kernel.  Per-cycle protocol (no collective, nothing on the critical path but the kernel itself):
//   * reuse: buffer b was last written by cycle c - B.  A peer's consumers of that data were ordered on its stream
//     before its kernel c - B + 1, so "every source has landed c - B + 2 cycles here" proves they are done.  Checked
//     every kGatherWaitEvery cycles for the cycles up to the next check, by a one-warp spin kernel on counters that are
//     normally long past the mark (peers run within a cycle or two of each other);
//   * landed signal: an event behind the kernel, then on the high-priority side stream a one-warp kernel whose
//     system-scope release waits for this cycle's posted NVLink writes to drain and bumps this rank's counter on every
//     rank — concurrent with the next cycle's kernel.
int shc_gather_step(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual, void* stream) {
  if (!e || !cmd) return fail(SHC_E_INVALID, "shc_gather_step: cmd is required");
  int rc = gather_ready(e, "shc_gather_step");
  if (rc != SHC_OK) return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  const size_t per_rank = (size_t)e->n * e->cfg.leg_count * e->cfg.joint_count;
  const long long cyc = e->gather_cycle;
  const int b = (int)(cyc % kGatherBuffers);
  if (cyc % kGatherWaitEvery == 0) {
    const long long need = cyc + kGatherWaitEvery - 1 - kGatherBuffers + 2;  // for the cycles cyc .. cyc + W - 1
    if (need > 0 && (rc = gather_wait(e, need, st)) != SHC_OK) return rc;
  }
  // tuning switches (never set in production): SHC_GATHER_TUNE=nostore keeps the protocol but stores only locally,
  // =nosignal stores everywhere but raises no landed signal (the waits then pass on the stale counters' timeout... so it
  // also skips the waits): they separate the cost of the stores from the cost of the signalling
  static const int tune = [] { const char* v = getenv("SHC_GATHER_TUNE"); return !v ? 0 : !strcmp(v, "nostore") ? 1 : !strcmp(v, "nosignal") ? 2 : 0; }();
  float* own_slot = e->gather_own + ((size_t)b * e->world + e->rank) * per_rank;
  const bool mc = e->gather_mc && tune != 1;
  StepIO io = make_io(e, cmd, imu, tip_force, manual, mc ? nullptr : own_slot);
  io.gather_offset = (long long)(((size_t)b * e->world + e->rank) * per_rank);
  io.gather_mc = mc ? e->gather_mc : nullptr;
  io.n_gather = 0;
  if (!e->gather_mc && tune != 1)
    for (int p = 0; p < e->world; ++p)
      if (p != e->rank) io.gather[io.n_gather++] = e->gather_peer[p];
  if ((rc = launch_cycle(e, io, st)) != SHC_OK) return rc;
  e->gather_cycle = cyc + 1;
  if (tune == 2) { e->gather_waited = e->gather_signalled = cyc + 1; return SHC_OK; }
  if (e->gather_trace && cyc < 2048) gather_stamp_kernel<<<1, 1, 0, st>>>(e->gather_trace + cyc * 4);
  if ((cyc + 1) % e->gather_signal_every == 0) return gather_signal(e, st);
  return SHC_OK;
}

// `stream` continues once every rank's shard of every cycle issued so far has landed in this rank's buffer.
int shc_gather_sync(shc_engine* e, int* last_buffer_out, void* stream) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  int rc = gather_ready(e, "shc_gather_sync");
  if (rc != SHC_OK) return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (last_buffer_out) *last_buffer_out = e->gather_cycle > 0 ? (int)((e->gather_cycle - 1) % kGatherBuffers) : -1;
  if ((rc = gather_signal(e, st)) != SHC_OK) return rc;
  return gather_wait(e, e->gather_cycle, st);
}

// Tuning (SHC_GATHER_TRACE=1): host pointer to the stamp table, [4096][4] u64 nanoseconds (rows 0..2047 per cycle: kernel
// end, signal start, signal end; rows 2048.. per wait: start, end, cycles needed).
unsigned long long* shc_gather_trace(shc_engine* e) { return e ? e->gather_trace : nullptr; }

// 0 while every device-side wait has completed; SHC_E_CUDA once one gave up (a peer stopped signalling).
int shc_gather_status(shc_engine* e) {
  if (!e || !e->gather_err) return SHC_OK;
  if (*(volatile int*)e->gather_err != 0)
    return fail(SHC_E_CUDA, "fused gather: a device-side wait for a peer's landed counter timed out (source rank " +
                                std::to_string(*(volatile int*)e->gather_err - 1) + ")");
  return SHC_OK;
}

// k cycles + the final wait.  Every rank must make the same sequence of calls.  A failure leaves the engine's cycle
// counter where the last successful cycle put it: the caller can tell how far every rank got and resume or tear down.
int shc_rollout_gather_fused(shc_engine* e, int k_cycles, const float* cmd_seq, int* last_buffer_out, void* stream) {
  if (!e || !cmd_seq || k_cycles < 1) return fail(SHC_E_INVALID, "shc_rollout_gather_fused: bad arguments");
  const size_t n = e->n;
  for (int k = 0; k < k_cycles; ++k) {
    int rc = shc_gather_step(e, cmd_seq + (size_t)k * n * 3, nullptr, nullptr, nullptr, stream);
    if (rc != SHC_OK) return rc;
  }
  return shc_gather_sync(e, last_buffer_out, stream);
}

#ifdef SHC_TRACE
int shc_debug_trace(shc_engine* e, unsigned long long* dev_buf) { e->trace = dev_buf; return SHC_OK; }
#endif
int shc_synchronize(shc_engine* e) {
  if (!e) return fail(SHC_E_INVALID, "null engine");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SHC_OK;
}
void* shc_stream(shc_engine* e) { return e ? (void*)e->stream : nullptr; }

size_t shc_bytes_per_step_device(const shc_engine* e) {
  if (!e) return 0;
  const IntConsts& ci = e->c.i;
  size_t state = (size_t)ci.nS * e->s_elem + (size_t)ci.nD * 8 + (size_t)ci.nI * 4;
  size_t in = 3 * 4 + (e->cfg.imu_posing || e->cfg.inclination_posing ? 10 * 4 : 0) +
              (e->cfg.admittance_control && !e->cfg.use_joint_effort ? (size_t)ci.L * 3 * 4 : 0);
  size_t outb = (size_t)ci.L * ci.D * 4;
  return 2 * state + in + outb;
}

size_t shc_bytes_per_step_algorithmic(const shc_engine* e) {
  // B_alg = 4 * [ 2 * (L * (2D + 33 + e_leg) + 38 + e_robot) + in + out ]
  if (!e) return 0;
  const int L = e->cfg.leg_count, D = e->cfg.joint_count;
  const bool imu = e->cfg.imu_posing || e->cfg.inclination_posing;
  int e_leg = e->cfg.admittance_control ? 8 : 0;
  int e_robot = imu ? 18 : 0;
  int in = 3 + (imu ? 10 : 0) + (e->cfg.admittance_control ? 3 * L : 0);
  int outw = L * D;
  return 4 * (size_t)(2 * (L * (2 * D + 33 + e_leg) + 38 + e_robot) + in + outw);
}

}  // extern "C"
