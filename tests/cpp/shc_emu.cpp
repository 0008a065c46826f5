// shc_emu.cpp — TEST INFRASTRUCTURE ONLY: runs the control-cycle source the CUDA kernel is built from (csrc/shc_cycle.cuh,
// Cycle<P, D, FULL>::run) on the host, lane after lane, over host-resident planes.  Built by tests/emu.py with plain g++
// into tests/cpp/_build/libshc_emu.so and loaded only by the CPU test-suite: it lets `pytest -m "not gpu"` check every
// branch of the cycle code against the oracle where no GPU exists.  It is not part of libshc_b200.so, the product has no
// CPU path, and nothing outside tests/ may load it.  Device-only pieces (TMA staging ring, mbarriers, the bounded sincos /
// rsqrt sequences) are replaced by their host equivalents in the headers themselves (#if defined(__CUDA_ARCH__)), so the
// emulator agrees with the device to rounding (1e-11 on one cycle), not bit for bit.
#define __host__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define SHC_EMU 1
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../syropod_highlevel_controller_b200/csrc/shc_msgs.cuh"
#include "../../syropod_highlevel_controller_b200/csrc/shc_sequence.cuh"
#include "../../syropod_highlevel_controller_b200/csrc/shc_pack.cuh"

using namespace shc;
using PrecMixed = Prec<float, float, double>;

struct shc_emu {
  shc_config cfg;
  shc_startup su;
  Consts c;
  int precision = 0, n = 0, n_pad = 0;
  int options = 0, pose_reset_mode = 0;
  std::vector<double> s64;
  std::vector<float> s32;
  std::vector<double> d;
  std::vector<int> i;
  std::vector<int> flags;
  std::vector<float> efforts, step_planes;
  bool have_efforts = false, have_step_planes = false;
  // sequences (csrc/shc_sequence.cuh on host planes)
  std::vector<double> seq_origin, tr_origin;
  std::vector<int> seq_count, seq_robot, seq_leg;
  std::vector<double> seq_target, seq_poses;
  double tr_desired[kMaxLegs * kMaxDof] = {};
  int tr_iteration = 0, tr_num = 0;
  bool tr_executing = false;
};

static thread_local std::string g_err;
static int fail(int code, const std::string& m) { g_err = m; return code; }

static void to_host(const shc_emu* e, HostPlanes& h) {
  if (e->precision == SHC_PRECISION_F64) h.s = e->s64;
  else h.s.assign(e->s32.begin(), e->s32.end());
  h.d = e->d;
  h.i = e->i;
}
static void from_host(shc_emu* e, const HostPlanes& h) {
  if (e->precision == SHC_PRECISION_F64) e->s64 = h.s;
  else { e->s32.resize(h.s.size()); for (size_t k = 0; k < h.s.size(); ++k) e->s32[k] = (float)h.s[k]; }
  e->d = h.d;
  e->i = h.i;
}

template <class P, int D, int MODE>
static void step_all(shc_emu* e, const StepIO& io_in) {
  using CY = Cycle<P, D, MODE>;
  using S = typename P::S;
  const IntConsts& ci = e->c.i;
  Planes<S> pl;
  if constexpr (sizeof(S) == 8) pl.s = (S*)e->s64.data(); else pl.s = (S*)e->s32.data();
  pl.d = e->d.data();
  pl.i = e->i.data();
  const int front = CY::FULL ? ci.frontS_leg : 0;
  std::vector<unsigned char> raw(ci.smem_per_warp + 256);
  unsigned char* wsm = (unsigned char*)(((uintptr_t)raw.data() + 127) / 128 * 128);
  const int LD = ci.L * D;
  const int tiles = (e->n + 31) / 32;
  for (int tile = 0; tile < tiles; ++tile) {
    for (int lane = 0; lane < 32; ++lane) CY::run(e->c, pl, tile, lane, io_in, wsm);
    const float* src = reinterpret_cast<const float*>(wsm + CY::kSlots * CY::slot_bytes(front));
    const int robots = std::min(32, e->n - tile * 32);
    for (int k = 0; k < robots * LD; ++k) io_in.joints_out[(size_t)tile * 32 * LD + k] = src[k];
  }
}

extern "C" {
const char* shc_emu_last_error(void) { return g_err.c_str(); }

int shc_emu_create(const shc_config* cfg, const shc_startup* startup, int n_robots, int precision, shc_emu** out) {
  if (!cfg || !out || n_robots < 1) return fail(SHC_E_INVALID, "bad arguments");
  std::string err;
  bool unsupported = false;
  if (!check_supported(*cfg, err, unsupported)) return fail(unsupported ? SHC_E_UNSUPPORTED : SHC_E_INVALID, err);
  if (!startup && !own_startup_supported(*cfg, err)) return fail(SHC_E_UNSUPPORTED, err);
  shc_emu* e = new shc_emu();
  core_init(e, *cfg, startup, n_robots, precision);
  if (!check_step_cycle(e->su, err)) { delete e; return fail(SHC_E_UNSUPPORTED, err); }
  const bool full = engine_full(e->cfg);
  const int front = full ? e->c.i.frontS_leg : 0;
  dispatch_D_raw(cfg->joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    e->c.i.smem_per_warp = precision == SHC_PRECISION_F64 ? Cycle<PrecF64, D, 0>::smem_per_warp(front, cfg->leg_count)
                                                          : Cycle<PrecMixed, D, 0>::smem_per_warp(front, cfg->leg_count);
    return 0;
  });
  HostPlanes h;
  initial_planes(e, h);
  from_host(e, h);
  e->flags.assign(e->n_pad, 0);
  *out = e;
  return SHC_OK;
}
void shc_emu_destroy(shc_emu* e) { delete e; }
int shc_emu_get_startup(const shc_emu* e, shc_startup* out) { *out = e->su; return SHC_OK; }
int shc_emu_set_options(shc_emu* e, int o) { e->options = o; return SHC_OK; }
int shc_emu_set_pose_reset_mode(shc_emu* e, int m) {
  if (m < 0 || m > 5) return fail(SHC_E_INVALID, "bad pose reset mode");
  e->pose_reset_mode = m;
  return SHC_OK;
}
int shc_emu_set_joint_efforts(shc_emu* e, const float* eff) {
  e->have_efforts = eff != nullptr;
  if (eff) e->efforts.assign(eff, eff + (size_t)e->n * e->cfg.leg_count * e->cfg.joint_count);
  return SHC_OK;
}
int shc_emu_set_tip_step_planes(shc_emu* e, const float* sp) {
  e->have_step_planes = sp != nullptr;
  if (sp) e->step_planes.assign(sp, sp + (size_t)e->n * e->cfg.leg_count * 3);
  return SHC_OK;
}
int shc_emu_get_status_flags(shc_emu* e, int* out) {
  if (!(e->options & SHC_OPT_STATUS_FLAGS)) return fail(SHC_E_INVALID, "status flags are not enabled");
  for (int r = 0; r < e->n; ++r) out[r] = e->flags[r];
  return SHC_OK;
}
int shc_emu_get_state(shc_emu* e, shc_robot_state* out, size_t n) {
  if (n != (size_t)e->n) return fail(SHC_E_INVALID, "need n_robots records");
  HostPlanes h;
  to_host(e, h);
  return dispatch_D_raw(e->cfg.joint_count, [&](auto dtag) -> int { unpack<decltype(dtag)::value>(e, h, out, n); return SHC_OK; });
}
int shc_emu_set_state(shc_emu* e, const shc_robot_state* in, size_t n) {
  if (n != (size_t)e->n) return fail(SHC_E_INVALID, "need n_robots records");
  const IntConsts& ci = e->c.i;
  HostPlanes h;
  h.s.assign((size_t)ci.nS * ci.n_pad, 0.0);
  h.d.assign((size_t)ci.nD * ci.n_pad, 0.0);
  h.i.assign((size_t)ci.nI * ci.n_pad, 0);
  pack(e, in, n, h, ci.n_pad);
  from_host(e, h);
  return SHC_OK;
}
int shc_emu_step(shc_emu* e, const float* cmd, const float* imu, const float* tip_force, const float* manual, float* joints_out) {
  if (!e || !cmd || !joints_out) return fail(SHC_E_INVALID, "cmd and joints_out are required");
  StepIO io;
  std::memset(&io, 0, sizeof(io));
  io.cmd = cmd; io.imu = imu; io.tip_force = tip_force; io.manual = manual;
  io.efforts = e->have_efforts ? e->efforts.data() : nullptr;
  io.step_planes = e->have_step_planes ? e->step_planes.data() : nullptr;
  io.joints_out = joints_out;
  io.tile_begin = 0;
  io.tile_end = (e->n + 31) / 32;
  io.flags_out = (e->options & SHC_OPT_STATUS_FLAGS) ? e->flags.data() : nullptr;
  io.pose_reset_mode = e->pose_reset_mode;
  const int mode = engine_mode(e->cfg);
  return dispatch_D_raw(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    if (e->precision == SHC_PRECISION_F64) {
      if (mode == 2) step_all<PrecF64, D, 2>(e, io); else if (mode == 1) step_all<PrecF64, D, 1>(e, io); else step_all<PrecF64, D, 0>(e, io);
    } else {
      if (mode == 2) step_all<PrecMixed, D, 2>(e, io); else if (mode == 1) step_all<PrecMixed, D, 1>(e, io); else step_all<PrecMixed, D, 0>(e, io);
    }
    return SHC_OK;
  });
}

// The sequence routines of csrc/shc_sequence.cuh (shc_step_to_new_stance, shc_pack_legs / shc_unpack_legs) on the emulator's
// planes, robot after robot.  kind 0 = stepToNewStance, 1 = packLegs(time), 2 = unpackLegs(time), 3 / 4 = executeSequence
// (START_UP / SHUT_DOWN); joints_out [n][L][D],
// progress_out [n] (pack / unpack: the same value for every robot).
int shc_emu_sequence_reset(shc_emu* e) {
  e->seq_origin.assign(seq_origin_count(e->cfg.leg_count, e->n_pad), 0.0);
  e->seq_count.assign(seq_count_count(e->cfg.leg_count, e->n_pad), -1);
  e->seq_robot.assign(seq_robot_count(e->n_pad), 0);
  for (int r = 0; r < e->n_pad; ++r) e->seq_robot[4 * (size_t)e->n_pad + r] = kSeqInitialFlags;
  e->seq_target.assign(seq_origin_count(e->cfg.leg_count, e->n_pad), 0.0);
  e->seq_poses.assign(seq_poses_count(e->cfg.leg_count, e->n_pad), 0.0);
  e->seq_leg.assign(seq_leg_count(e->cfg.leg_count, e->n_pad), 0);
  return SHC_OK;
}
int shc_emu_sequence_step(shc_emu* e, int kind, double time, float* joints_out, int* progress_out) {
  if (!e || !joints_out || !progress_out) return fail(SHC_E_INVALID, "bad arguments");
  const int L = e->cfg.leg_count;
  return dispatch_D_raw(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    auto go = [&](auto pl) -> int {
      using S = typename std::remove_pointer<decltype(pl.s)>::type;
      if (e->seq_count.empty()) shc_emu_sequence_reset(e);
      if (kind == 3 || kind == 4) {
        ExecuteSequenceParams ep{e->cfg.swing_height, e->cfg.step_frequency, e->cfg.time_delta};
        SeqBuffers sq{e->seq_origin.data(), e->seq_count.data(), e->seq_robot.data(), (size_t)e->n_pad, e->seq_target.data(),
                      e->seq_poses.data(), e->seq_leg.data()};
        for (int r = 0; r < e->n; ++r) progress_out[r] = execute_sequence_robot<S, D>(e->c, pl, sq, ep, kind == 4, r, joints_out);
        return SHC_OK;
      }
      if (kind == 0) {
        NewStanceParams np;
        np.lift_height = e->cfg.swing_height;
        np.num_iterations = std::max(1, round_to_int((1.0 / e->cfg.step_frequency) / e->cfg.time_delta));
        np.apply_delta = 1;
        SeqBuffers sq{e->seq_origin.data(), e->seq_count.data(), e->seq_robot.data(), (size_t)e->n_pad, e->seq_target.data(),
                      e->seq_poses.data(), e->seq_leg.data()};
        for (int r = 0; r < e->n; ++r) progress_out[r] = step_to_new_stance_robot<S, D>(e->c, pl, sq, np, r, joints_out);
        return SHC_OK;
      }
      const long long total = (long long)e->n * L * D;
      if (kind == 1)  // packLegs: transition_step_ = 0 (pose_controller.cpp:618)
        for (int r = 0; r < e->n_pad; ++r) e->seq_robot[2 * (size_t)e->n_pad + r] = 0;
      if (!e->tr_executing) {
        e->tr_origin.resize((size_t)total);
        for (long long i = 0; i < total; ++i) latch_joint<S, D>(e->c, pl, i, e->tr_origin.data());
        for (int l = 0; l < L; ++l)
          for (int j = 0; j < D; ++j) e->tr_desired[l * kMaxDof + j] = kind == 1 ? e->cfg.joint_packed[l][j] : e->cfg.joint_unpacked[l][j];
        e->tr_iteration = 0;
        e->tr_num = std::max(1, round_to_int(time / e->cfg.time_delta));
      }
      int progress = 100;
      if (e->tr_iteration < e->tr_num) {
        const int it = ++e->tr_iteration;
        for (long long i = 0; i < total; ++i) transition_joint<S, D>(e->c, pl, e->tr_origin.data(), e->tr_desired, it, e->tr_num, i, joints_out);
        progress = it >= e->tr_num ? 100 : std::max(1, int((double(it - 1) / double(e->tr_num)) * 100));
      }
      e->tr_executing = progress != 0 && progress != 100;
      for (int r = 0; r < e->n; ++r) progress_out[r] = progress;
      return SHC_OK;
    };
    if (e->precision == SHC_PRECISION_F64) return go(Planes<double>{e->s64.data(), e->d.data(), e->i.data()});
    return go(Planes<float>{e->s32.data(), e->d.data(), e->i.data()});
  });
}

// The message packer of shc_pack_messages (csrc/shc_msgs.cuh) on the emulator's planes: records of robot `r`.
int shc_emu_pack_messages(shc_emu* e, int r, const float* measured, shc_joint_state_msg* js, shc_leg_state_msg* legs, shc_body_msg* body) {
  if (!e || r < 0 || r >= e->n) return fail(SHC_E_INVALID, "robot out of range");
  const int L = e->cfg.leg_count;
  return dispatch_D_raw(e->cfg.joint_count, [&](auto dtag) -> int {
    constexpr int D = decltype(dtag)::value;
    auto go = [&](auto pl) {
      using S = typename std::remove_pointer<decltype(pl.s)>::type;
      for (int l = 0; l < L; ++l) pack_leg_message<S, D>(e->c, pl, r, l, measured ? measured + ((size_t)r * L + l) * D : nullptr, legs[l], js);
      pack_body_message<S>(e->c, pl, r, *body);
    };
    if (e->precision == SHC_PRECISION_F64) go(Planes<double>{e->s64.data(), e->d.data(), e->i.data()});
    else go(Planes<float>{e->s32.data(), e->d.data(), e->i.data()});
    return SHC_OK;
  });
}
}
