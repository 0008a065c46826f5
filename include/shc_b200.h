/*
 * shc_b200.h — C-ABI of the batched, B200-native SHC control-cycle engine (libshc_b200.so).
 *
 * The reference (csiro-robotics/syropod_highlevel_controller v0.5.11) has no plugin / FFI interface: it is one ROS
 * executable (CMakeLists.txt:116-117,146).  The only seam is the C++ class surface StateController drives each
 * cycle.  Every entry point below therefore cites the reference call(s) it replaces; the C++ facade in
 * include/shc_facade.hpp re-exposes them under the reference's own class and method names.
 *
 * One engine = N independent robots of one morphology (L legs x D joints) on one CUDA device, state resident in HBM.
 * All functions return 0 on success and a negative SHC_E_* code on failure; nothing throws.  An engine handle is not
 * re-entrant; different handles may be used from different threads (as the reference: single-threaded per controller,
 * main.cpp:106-132).  There is NO CPU fallback: without a CUDA device shc_create fails with SHC_E_CUDA.
 */
#ifndef SHC_B200_H
#define SHC_B200_H

#include <stddef.h>

#include "shc_config.h"
#include "shc_msgs.h"
#include "shc_state.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct shc_engine shc_engine;

enum {
  SHC_OK = 0,
  SHC_E_INVALID = -1,     /* bad argument / unsupported configuration (message via shc_last_error) */
  SHC_E_CUDA = -2,        /* CUDA runtime error or no device */
  SHC_E_UNSUPPORTED = -3  /* reference feature outside the hot-path scope (rough terrain, manual legs, ...) */
};

/* Arithmetic of the device path. */
enum {
  SHC_PRECISION_F64 = 0,   /* state and arithmetic in double: tracks the reference's Eigen double path to ~1e-12 */
  SHC_PRECISION_MIXED = 1  /* fp32 state/IO, fp64 only for the open-loop accumulators and branch decisions */
};

/* Per-robot status word written by shc_step when flags are enabled (shc_set_options).  Replaces the reference's
 * rosconsole warnings / ROS_FATAL (model.cpp:852,916-929; pose_controller.cpp:1228-1232). */
enum {
  SHC_FLAG_IK_DEVIATION = 1,   /* some leg: |FK(q) - desired| > IK_TOLERANCE on an axis (applyIK would return 0) */
  SHC_FLAG_POSITION_CLAMP = 2, /* some joint clamped to its position limit */
  SHC_FLAG_VELOCITY_CLAMP = 4, /* some joint clamped to its velocity limit */
  SHC_FLAG_IMU_UNSTABLE = 8    /* IMU PID correction beyond STABILITY_THRESHOLD */
};

enum { SHC_OPT_STATUS_FLAGS = 1 }; /* compute the status word each cycle (costs one extra FK per leg) */

/* Creates an engine for n_robots robots in the state the reference reaches at the end of its direct start-up
 * (StateController::init + transitionRobotState PACKED->READY->RUNNING, state_controller.cpp:127-281).
 * `startup` may be NULL: the engine then runs its own host-side restatement of directStartup / generateWorkspaces /
 * generateWalkspace / generateLimits once (pose_controller.cpp:463, model.cpp:120, walk_controller.cpp:57,231);
 * otherwise the given constants are used as they are.  `device` is the CUDA ordinal. */
int shc_create(const shc_config* cfg, const shc_startup* startup, int n_robots, int device, int precision,
               shc_engine** out);
void shc_destroy(shc_engine* e);

/* The start-up constants alone (host arithmetic only; needs no device): what shc_create computes when `startup` is
 * NULL.  Mirrors WalkController::generateStepCycle / generateLimits / generateWalkspace, Model::generateWorkspaces and
 * PoseController::directStartup of the reference. */
int shc_compute_startup(const shc_config* cfg, shc_startup* out);

/* Host (double) evaluation of the Leg::applyIK routine shared with the kernels: one leg, q/qd [D] in/out. */
int shc_host_apply_ik(const shc_config* cfg, int leg, double* q, double* qd, const double* desired, int simulation,
                      double* tip_out, double* ik_result);

/* Last error text of this thread (static storage). */
const char* shc_last_error(void);

/* Start-up constants in use (WalkController::getWalkspace / limit maps / getStepCycle). */
int shc_get_startup(const shc_engine* e, shc_startup* out);
int shc_n_robots(const shc_engine* e);
int shc_options(const shc_engine* e);
int shc_set_options(shc_engine* e, int options);
/* PoseController::setPoseResetMode (pose_controller.h:105); applies to every robot of the batch. */
int shc_set_pose_reset_mode(shc_engine* e, int mode);

/* Whole-batch state exchange in the array-of-structs record of shc_state.h (n_robots records).
 * Replaces the reference's getters over Model/Leg/LegStepper/PoseController members (state_controller.cpp:809-1078). */
int shc_get_state(shc_engine* e, shc_robot_state* out, size_t n_records);
int shc_set_state(shc_engine* e, const shc_robot_state* in, size_t n_records);
/* Records of the robots [first, first + count) only: three small device->host copies however large the batch is (what the
 * facade's per-robot getters use). */
int shc_get_state_range(shc_engine* e, size_t first, size_t count, shc_robot_state* out);
/* The records of the robots [first, first + count) replaced; only the tiles of 32 robots that hold them travel. */
int shc_set_state_range(shc_engine* e, size_t first, size_t count, const shc_robot_state* in);
/* Gait switch / parameter change for the whole batch = what StateController::changeGait does once the walker has STOPPED
 * (state_controller.cpp:513-540: initGaitParameters, WalkController::generateStepCycle + generateLimits — step cycle, the
 * four limit maps, the legs' phase offsets — and, with auto posing, initAutoPoseParameters + PoseController::
 * setAutoPoseParams), and what a change of a constants-only adjustable parameter does (swing height / width, step depth,
 * admittance parameters: state_controller.cpp:451-508 without the step-frequency branch, which re-phases walking legs).
 * A NEW engine is created for `cfg` (same leg / joint counts, batch size, device and precision as `src`; `startup` as for
 * shc_create) and the state records of `src` are carried over chunk by chunk, as the reference keeps its state across
 * changeGait; options and the pose reset mode follow.  `src` is left untouched: the caller destroys it (and re-attaches
 * gather / NCCL buffers and input latches to the new engine).  Timing, as in the reference: a gait change takes the place of
 * one cycle (the loop() in which changeGait runs updates no tips, state_controller.cpp:391-395, 427) once every robot has
 * STOPPED — that loop() of the reference still runs its pose and admittance stages, so with IMU posing or admittance
 * control active the reference's IMU PID and admittance filters are one step ahead of the engine's after the switch (4.5e-4
 * rad of imu_pose in the octopod test configuration); without them the switch is exact; walk parameters act in the loop() that applies them (switch before that cycle); admittance parameters one loop
 * later (updateAdmittance has run when runningState() applies them: switch after that cycle); step_frequency with the batch
 * at rest (re-phasing walking legs, LegStepper::updatePhase, is not supported; nor is the reference's deferral of the new
 * step cycle while its signed velocity test fails, state_controller.cpp:489-491 — that decision is the caller's, with
 * shc_set_limit_maps for the interim speed limits): SHC_E_UNSUPPORTED if the step cycle changes while a robot is not
 * STOPPED.  Checked against the reference's own changeGait / adjustParameter
 * (tests/test_emu_parity.py on the host, tests/test_gpu_properties.py on the B200). */
enum { SHC_RECONF_KEEP_POSE_CYCLE = 1 }; /* adjustParameter semantics: the auto-pose cycle (pose phase length and normaliser,
                                          * PoseController::setAutoPoseParams) is NOT regenerated — the reference re-runs it
                                          * in changeGait only, so a step-frequency change leaves a synchronised auto-pose
                                          * cycle at its old length (state_controller.cpp:451-508 vs :513-540) */
int shc_clone_reconfigured(shc_engine* src, const shc_config* cfg, const shc_startup* startup, int flags, shc_engine** out);

/* WalkController::set{LinearSpeed,AngularSpeed,LinearAcceleration,AngularAcceleration}LimitMap (walk_controller.h:126-141):
 * replaces the limit tables getLimit (walk_controller.cpp:414) reads, 9 values each (bearings 0..360 step 45); NULL keeps a
 * table.  From the next cycle on. */
int shc_set_limit_maps(shc_engine* e, const double* max_linear_speed, const double* max_angular_speed,
                       const double* max_linear_acceleration, const double* max_angular_acceleration);

/* ONE control cycle for every robot = StateController::loop() in RUNNING state (state_controller.cpp:162-193):
 *   PoseController::updateCurrentPose -> WalkController::setPoseState -> [AdmittanceController::updateStiffness,
 *   updateAdmittance] -> runningState(): WalkController::updateWalk -> PoseController::updateStance ->
 *   Model::updateModel (Leg::setDesiredTipPose + Leg::applyIK per leg).
 * All pointers are DEVICE pointers, float32:
 *   cmd        [N][3]  body velocity command (vx, vy, wz) as published on the reference's velocity topic; the
 *                      bodyVelocityInputCallback scaling/clamp (state_controller.cpp:1127-1136) is applied inside
 *   imu        [N][10] or NULL: orientation quaternion (w,x,y,z), angular velocity (3), linear acceleration (3)
 *                      (imuCallback -> Model::setImuData, state_controller.cpp:1551-1562)
 *   tip_force  [N][L][3] or NULL: measured tip force (tipStatesCallback, state_controller.cpp:1618-1646)
 *   manual     [N][6]  or NULL: manual pose velocity input, translation xyz + rotation rpy (setManualPoseInput :1148)
 *   joints_out [N][L][D] desired joint positions + offset, the order of Leg::generateDesiredJointStateMsg
 *                      (model.cpp:605-617) / publishDesiredJointState (state_controller.cpp:777-805)
 * `stream` is a cudaStream_t (NULL = the engine's own stream).  Asynchronous. */
int shc_step(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual,
             float* joints_out, void* stream);

/* Same cycle with HOST buffers: pinned staging, host->device copy of the inputs, kernel, device->host copy of the joint
 * angles, then a stream synchronize.  This is the call the C++ facade makes per cycle. */
int shc_step_host(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual,
                  float* joints_out);

/* k_cycles cycles with device-resident per-cycle commands cmd_seq [k][N][3] (imu_seq / force_seq likewise or NULL);
 * joints_out receives the last cycle.  Launches are captured once into a CUDA graph per (k, pointer set). */
int shc_rollout(shc_engine* e, int k_cycles, const float* cmd_seq, const float* imu_seq, const float* force_seq,
                float* joints_out, void* stream);

/* Measured joint efforts for Leg::calculateTipForce (jointStatesCallback, state_controller.cpp:1565-1590): device
 * pointer float [N][L][D], latched until changed; NULL = all zero.  Only read when use_joint_effort is set. */
int shc_set_joint_efforts(shc_engine* e, const float* efforts_dev);

/* Tip range-sensor readings for rough-terrain mode (TipState.step_plane, tipStatesCallback state_controller.cpp:1650-1675):
 * device pointer float [N][L][3] = (x slope, y slope, z range along the tip's x axis), latched until changed; NULL = no range
 * sensors.  A range >= SHC_RANGE_UNASSIGNED (shc_config.h) means "no reading" (the reference's UNASSIGNED_VALUE): the leg's step plane is
 * forgotten.  Only read when rough_terrain_mode is set. */
int shc_set_tip_step_planes(shc_engine* e, const float* step_planes_dev);

/* Output wire formats (records in shc_msgs.h): JointState, LegState, velocity / pose / rotation-error
 * and frame-transform records of the robots [first, first + count) as of the last cycle, packed by ONE kernel from the state
 * planes (one thread per leg) — the per-leg / per-joint host loops of state_controller.cpp:777-1047 disappear.  Outputs:
 * joint_state_out [count], leg_state_out [count][L], body_out [count]; device memory or page-locked host memory (written
 * in place, no copy); any may be NULL.  measured_joint_positions_dev: float [N][L][D] or NULL (LegState.actual_tip_pose).
 * Asynchronous on `stream` (NULL = the engine's). */
int shc_pack_messages(shc_engine* e, size_t first, size_t count, const float* measured_joint_positions_dev,
                      shc_joint_state_msg* joint_state_out, shc_leg_state_msg* leg_state_out, shc_body_msg* body_out, void* stream);

/* Start-up on the device.
 *   shc_startup_begin         PoseController::directStartup (pose_controller.cpp:463) begins for the whole batch: latches
 *                             every robot's origin configuration (joint_positions_dev, double [N][L][D]: the measured joint
 *                             states; NULL = the default joint positions), computes the desired configuration of every leg
 *                             and resets the robots to the post-start-up stepper / walker / poser state.
 *   shc_startup_step          one loop() of the start-up: LegPoser::transitionConfiguration (:1476) for every joint of the
 *                             batch (one kernel); joints_out_dev (float [N][L][D]) gets the joint commands.  Returns the
 *                             reference's progress (1..99, 100 = PROGRESS_COMPLETE: READY) or a negative SHC_E_* code.
 *   shc_direct_startup        begin + the final iteration only (callers that do not need the trajectory).
 *   shc_generate_workspaces   Leg::generateWorkspace (model.cpp:309-510) for every leg, the eight bearing searches of a
 *                             workplane (<= 500 Leg::applyIK(true) steps each) on eight lanes; full = 0 the simple one-plane
 *                             workspace, 1 the layered workspace of rough-terrain mode.  HOST outputs heights [L][max_planes],
 *                             radii [L][max_planes][9], n_planes [L].
 *   shc_host_generate_workspaces  the same routine on the host (no device needed). */
/* Stepping and joint-space sequences on the device, each call = one loop() for the whole batch:
 *   shc_step_to_new_stance    PoseController::stepToNewStance (pose_controller.cpp:520): the legs of each robot's current
 *                             group step to their default tip positions (LegPoser::stepToPosition :1571: dual quartic
 *                             Bezier, swing height as lift, one step period, body pose blended in) + Leg::applyIK
 *                             (model.cpp:861); the two leg groups alternate per robot as its own legs complete.
 *                             joints_out_dev float [N][L][D] / progress_out_dev int [N] (each robot's return value), either
 *                             may be NULL.  Returns the smallest progress over the batch (blocking on `stream`).
 *   shc_sequence_reset        forgets a stepping sequence in progress.
 *   shc_transition_begin/step PoseController::transitionConfiguration (:703) / LegPoser::transitionConfiguration (:1476): the
 *                             batch moves from the joint positions its state holds to desired_configuration (HOST doubles
 *                             [L][D]) in max(1, roundToInt(time / time_delta)) loops; step returns the progress 1..100.
 *   shc_pack_legs / shc_unpack_legs  one loop() of PoseController::packLegs (:597) / unpackLegs (:661) towards
 *                             shc_config.joint_packed / joint_unpacked; return the progress (100 = complete).
 *   shc_execute_sequence      PoseController::executeSequence (:145-461): the start-up (shut_down = 0) / shut-down sequence —
 *                             alternating horizontal (leg groups step in turn, or all at once while the body bears no
 *                             load) and vertical (body rises / sinks) transitions towards the transition poses the first
 *                             start-up records, with its safety factor on the joint-limit proximity; every robot follows
 *                             its own course on the device.  progress_out_dev int [N]: -1 (SHC_SEQ_GENERATING) while a
 *                             robot's first start-up generates its sequence, 0..100, -2 (SHC_SEQ_FAILED) once it needed
 *                             more than 20 transition steps (where the reference shuts down).  *min_progress_out = the
 *                             smallest value over the batch (blocking on `stream`); the return value is SHC_OK or an error. */
int shc_step_to_new_stance(shc_engine* e, float* joints_out_dev, int* progress_out_dev, void* stream);
int shc_sequence_reset(shc_engine* e);
enum { SHC_SEQ_GENERATING = -1, SHC_SEQ_FAILED = -2 };
int shc_execute_sequence(shc_engine* e, int shut_down, float* joints_out_dev, int* progress_out_dev, int* min_progress_out, void* stream);
int shc_transition_begin(shc_engine* e, const double* desired_configuration, double transition_time);
int shc_transition_step(shc_engine* e, float* joints_out_dev, void* stream);
int shc_pack_legs(shc_engine* e, double time_to_pack, float* joints_out_dev, void* stream);
int shc_unpack_legs(shc_engine* e, double time_to_unpack, float* joints_out_dev, void* stream);
/* Host-buffer form of one loop() of a sequence (the C++ facade uses it): joints_out [N][L][D] / progress_out [N] are HOST
 * arrays, either may be NULL; SHC_SEQ_DIRECT_STARTUP begins the direct start-up on its first call (default joint positions).
 * Returns the smallest progress over the batch or a negative SHC_E_* code. */
enum { SHC_SEQ_NEW_STANCE = 0, SHC_SEQ_PACK = 1, SHC_SEQ_UNPACK = 2, SHC_SEQ_DIRECT_STARTUP = 3, SHC_SEQ_START_UP = 4, SHC_SEQ_SHUT_DOWN = 5 };
int shc_sequence_step_host(shc_engine* e, int kind, double time, float* joints_out, int* progress_out);
int shc_startup_begin(shc_engine* e, const double* joint_positions_dev);
int shc_startup_step(shc_engine* e, float* joints_out_dev, void* stream);
int shc_direct_startup(shc_engine* e, const double* joint_positions_dev, float* joints_out_dev, void* stream);
int shc_generate_workspaces(shc_engine* e, int full, int max_planes, double* heights_out, double* radii_out, int* n_planes_out);
int shc_host_generate_workspaces(const shc_config* cfg, const shc_startup* startup, int full, int max_planes, double* heights_out,
                                 double* radii_out, int* n_planes_out);

/* Multi-GPU: robots are independent, so the batch is sharded across ranks with no data-path
 * collective; the one exchange is an all-gather of the joint angles per control cycle (BASELINE configs[4]).  NCCL is
 * resolved at run time from the libnccl already loaded in the process.
 *   shc_nccl_unique_id   rank 0 creates the 128-byte id, the caller broadcasts it (e.g. torch.distributed)
 *   shc_nccl_init        every rank joins (one communicator per engine)
 *   shc_allgather_joints local [n][L][D] -> full [world*n][L][D] on `stream` (NULL = the engine's side stream)
 *   shc_rollout_allgather k cycles; cycle t's gather runs on a side stream, double buffered (local2 = 2 x [n][L][D],
 *                        full2 = 2 x [world*n][L][D]), overlapping cycle t+1's kernel; `stream` resumes after the last
 *                        gather. */
int shc_nccl_unique_id(void* out128);
int shc_nccl_init(shc_engine* e, const void* uid128, int rank, int world_size);
int shc_allgather_joints(shc_engine* e, const float* local, float* full, void* stream);
int shc_rollout_allgather(shc_engine* e, int k_cycles, const float* cmd_seq, float* local2, float* full2, void* stream);

/* Fused all-gather over peer memory (one node, <= 8 ranks; SURVEY.md 8(e) "fused variant"): the control-cycle kernel
 * stores every finished tile of joint commands into ALL ranks' gather buffers, so the exchange travels over NVLink /
 * NVSwitch while the rest of the batch is still being computed.  Cycle t lands in buffer t % shc_gather_buffers() of
 * every rank: element [buffer][source rank][robot][leg][joint], followed (256-byte aligned) by one landed counter per
 * source rank.  Two ways to map the buffers:
 *   shc_gather_attach        the caller owns one symmetric buffer per rank (shc_gather_bytes() each, e.g. torch
 *                            symmetric memory) and passes every rank's mapping of them and — when the fabric has NVLS —
 *                            their NVSwitch multicast mapping: the kernel then issues ONE multimem.st per 16 bytes and the
 *                            switch replicates it (egress per rank = its shard).  multicast_buffer NULL: unicast TMA bulk
 *                            stores, one per peer.  The caller runs a barrier between attach and the first cycle.
 *   shc_gather_alloc /       the engines cudaMalloc the buffers and exchange CUDA-IPC handles (after shc_nccl_init, which
 *   shc_gather_open_peer     fixes rank and world size); unicast TMA bulk stores.
 * Cycles (every rank makes the same sequence of calls; nothing on the per-cycle path is a collective):
 *   shc_gather_step          one control cycle (inputs as shc_step) into the next buffer.  Behind the kernel a one-warp
 *                            kernel on a high-priority side stream waits for the posted NVLink writes to drain and bumps
 *                            this rank's landed counter on every rank (multimem.red / st.release.sys), concurrently with
 *                            the next cycle; every 8th cycle a one-warp ld.acquire.sys spin kernel checks that no rank
 *                            is more than a buffer-reuse window behind (it normally passes at once).
 *   shc_gather_sync          `stream` continues once every rank's shard of every cycle issued so far has landed in this
 *                            rank's buffer; *last_buffer_out = buffer of the last cycle.
 *   shc_rollout_gather_fused k x shc_gather_step + shc_gather_sync.
 *   shc_gather_status        SHC_E_CUDA once a device-side wait gave up (a peer stopped signalling for seconds). */
size_t shc_gather_bytes(const shc_engine* e, int world_size);
int shc_gather_attach(shc_engine* e, int rank, int world_size, void* const* peer_buffers, void* multicast_buffer);
int shc_gather_alloc(shc_engine* e, void* handle64_out, float** buffer_out);
int shc_gather_open_peer(shc_engine* e, int peer_rank, const void* handle64);
int shc_gather_buffers(void);
int shc_gather_step(shc_engine* e, const float* cmd, const float* imu, const float* tip_force, const float* manual, void* stream);
int shc_gather_sync(shc_engine* e, int* last_buffer_out, void* stream);
int shc_gather_status(shc_engine* e);
int shc_rollout_gather_fused(shc_engine* e, int k_cycles, const float* cmd_seq, int* last_buffer_out, void* stream);

/* The engine's own CUDA stream (cudaStream_t) and a blocking wait on it. */
void* shc_stream(shc_engine* e);
int shc_synchronize(shc_engine* e);

/* Device pointer to the per-robot status words (int32 [N]) of the last cycle; NULL unless SHC_OPT_STATUS_FLAGS. */
const int* shc_status_flags_device(const shc_engine* e);
/* Blocking copy of the status words to host memory (int32 [N]). */
int shc_get_status_flags(shc_engine* e, int* host_out);

/* Stand-alone batched Leg::applyIK (model.cpp:861) on caller-provided joint state, used by the start-up sweeps
 * (workspace generation) and by the kernel unit tests.  Device pointers; n_legs rows each using the chain of leg
 * index leg_id[i]:  q, qd [n][D] (in/out, double), desired_tip [n][3] (base_link frame, double), simulation flag as
 * Leg::applyIK(simulation); ik_result [n] (double) receives applyIK's return value. */
int shc_apply_ik(shc_engine* e, int n_legs, const int* leg_id, double* q, double* qd, const double* desired_tip,
                 int simulation, double* tip_out, double* ik_result, void* stream);

/* Byte sizes for binding checks. */
size_t shc_sizeof_config(void);
size_t shc_sizeof_startup(void);
size_t shc_sizeof_robot_state(void);

/* Algorithmic bytes one control cycle moves per robot for this engine's configuration (state read + write, inputs,
 * outputs), as laid out on the device; and the figure of SURVEY.md §8(d) (4-byte words) for the same configuration. */
size_t shc_bytes_per_step_device(const shc_engine* e);
size_t shc_bytes_per_step_algorithmic(const shc_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* SHC_B200_H */
