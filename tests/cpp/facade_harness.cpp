// facade_harness.cpp — drives the batched engine through the C++ façade (include/shc_facade.hpp) with the call
// sequence of the reference's StateController::loop() / runningState() (state_controller.cpp:162-193, 379-447).
// ROS is absent from the image, so state_controller.cpp cannot be built as the node it is (oracle/_ref compiles it against
// stand-in headers for the parity pin only); this harness is the stand-in caller of the facade.
//
//   facade_harness <config.bin> <startup.bin> <n_robots> <cycles> <cmd.bin [cycles][n][3] f32> <out.bin>
// writes, per cycle, the desired joint positions [n][L][D] (f64, from Joint::desired_position_) followed by each
// robot's walk state (f64) to out.bin.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "shc_facade.hpp"

using namespace shc_b200;

template <class T> static bool read_blob(const char* path, T* out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  size_t n = std::fread(out, 1, sizeof(T), f);
  std::fclose(f);
  return n == sizeof(T);
}

// One robot's slice of StateController, holding the same members (state_controller.h:300-370).
struct StateControllerLike {
  std::shared_ptr<Model> model_;
  std::shared_ptr<WalkController> walker_;
  std::shared_ptr<PoseController> poser_;
  std::shared_ptr<AdmittanceController> admittance_;
  RobotState robot_state_ = RUNNING;
  Vector2d linear_velocity_input_;
  double angular_velocity_input_ = 0.0;
  bool admittance_control = false, dynamic_stiffness = false;
  int primary_leg_selection_ = -1, secondary_leg_selection_ = -1;
  Vector3d primary_tip_velocity_input_, secondary_tip_velocity_input_;
  Pose primary_pose_input_, secondary_pose_input_;

  void runningState() {  // state_controller.cpp:379-447 (no gait change / leg toggle / planner / cruise in the harness)
    walker_->updateWalk(linear_velocity_input_, angular_velocity_input_);
    walker_->updateManual(primary_leg_selection_, primary_tip_velocity_input_, secondary_leg_selection_, secondary_tip_velocity_input_);
    walker_->updateManual(primary_leg_selection_, primary_pose_input_, secondary_leg_selection_, secondary_pose_input_);
    poser_->updateStance();
    model_->updateModel();
  }
  void loop() {  // state_controller.cpp:162-193
    if (robot_state_ != UNKNOWN) {
      poser_->updateCurrentPose(robot_state_);
      walker_->setPoseState(poser_->getAutoPoseState());
      if (admittance_control) {
        if (walker_->getWalkState() != STOPPED && dynamic_stiffness) admittance_->updateStiffness(walker_);
        admittance_->updateAdmittance();
      }
    }
    if (robot_state_ == RUNNING) runningState();
  }
};

int main(int argc, char** argv) {
  if (argc != 7) {
    std::fprintf(stderr, "usage: %s config.bin startup.bin n_robots cycles cmd.bin out.bin\n", argv[0]);
    return 2;
  }
  Parameters params;
  shc_startup startup;
  if (!read_blob(argv[1], &params.cfg) || !read_blob(argv[2], &startup)) {
    std::fprintf(stderr, "cannot read config/startup blobs\n");
    return 2;
  }
  const int n = std::atoi(argv[3]), cycles = std::atoi(argv[4]);
  const int L = params.cfg.leg_count, D = params.cfg.joint_count;
  std::vector<float> cmd(size_t(cycles) * n * 3);
  {
    FILE* f = std::fopen(argv[5], "rb");
    if (!f || std::fread(cmd.data(), sizeof(float), cmd.size(), f) != cmd.size()) {
      std::fprintf(stderr, "cannot read commands\n");
      return 2;
    }
    std::fclose(f);
  }
  try {
    Batch batch(params, n, 0, SHC_PRECISION_F64, &startup);
    std::vector<StateControllerLike> sc(n);
    for (int r = 0; r < n; ++r) {
      Controllers& c = batch.robot(r);
      sc[r].model_ = c.model_; sc[r].walker_ = c.walker_; sc[r].poser_ = c.poser_; sc[r].admittance_ = c.admittance_;
      sc[r].admittance_control = params.cfg.admittance_control != 0;
      sc[r].dynamic_stiffness = params.cfg.dynamic_stiffness != 0;
    }
    FILE* out = std::fopen(argv[6], "wb");
    std::vector<double> row(size_t(n) * L * D + n);
    for (int c = 0; c < cycles; ++c) {
      for (int r = 0; r < n; ++r) {  // bodyVelocityInputCallback (state_controller.cpp:1127) then loop()
        const float* m = &cmd[(size_t(c) * n + r) * 3];
        sc[r].linear_velocity_input_ = Vector2d(m[0], m[1]);
        sc[r].angular_velocity_input_ = m[2];
        sc[r].loop();
      }
      for (int r = 0; r < n; ++r) {  // publishDesiredJointState (state_controller.cpp:777-805)
        for (int l = 0; l < L; ++l)
          for (int j = 0; j < D; ++j)
            row[(size_t(r) * L + l) * D + j] = sc[r].model_->getLegByIDNumber(l)->getJointByIDNumber(j + 1)->desired_position_;
        row[size_t(n) * L * D + r] = double(sc[r].walker_->getWalkState());
      }
      std::fwrite(row.data(), sizeof(double), row.size(), out);
    }
    // Sequences through the facade (pose_controller.cpp:520 stepToNewStance, :597 packLegs): every robot's StateController
    // calls its own poser_ each loop, as adjustParameter / the PACKED transition do; rows = joint commands + each robot's progress.
    const long steps_before = batch.cycles();
    const int num = std::max(1, int((1.0 / params.cfg.step_frequency) / params.cfg.time_delta + 0.5));
    int seq_loops = 0;
    for (int phase = 0; phase < 2; ++phase) {
      for (int k = 0; k < (phase == 0 ? 2 * num : 100000); ++k) {
        int min_progress = 1000;
        for (int r = 0; r < n; ++r) {
          const int p = phase == 0 ? sc[r].poser_->stepToNewStance() : sc[r].poser_->packLegs(2.0);
          row[size_t(n) * L * D + r] = double(p);
          min_progress = std::min(min_progress, p);
        }
        const std::vector<float>& j = batch.desiredJointPositions();
        for (size_t i = 0; i < size_t(n) * L * D; ++i) row[i] = j[i];
        std::fwrite(row.data(), sizeof(double), row.size(), out);
        ++seq_loops;
        if (phase == 1 && min_progress == 100) break;
      }
    }
    std::fclose(out);
    std::printf("facade harness: %d robots x %d cycles, %ld engine steps, %d sequence loops in %ld device steps\n", n, cycles,
                steps_before, seq_loops, batch.cycles() - steps_before);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
