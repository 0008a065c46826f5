"""workspace generation (Leg::generateWorkspace, model.cpp:309-510) and the direct start-up
(PoseController::directStartup, pose_controller.cpp:463).  The routines live in csrc/shc_startup.cuh and are compiled for
the host (engine constants, shc_host_generate_workspaces) and for the device (shc_generate_workspaces, shc_startup_step).
CPU part: the host build against the oracle.  GPU part: the device kernels against the host build and the oracle."""
import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config


def _sorted(h, r, n):
    order = np.argsort(h[:n], kind="stable")
    return h[:n][order], r[:n][order]


@pytest.mark.parametrize("robot", ["hexapod", "octopod"])
@pytest.mark.parametrize("full", [False, True], ids=["simple", "layered"])
def test_host_workspace_sweep_matches_oracle(shc_lib, oracle, robot, full):
    from syropod_highlevel_controller_b200 import engine

    cfg = hexapod_config("tripod_gait") if robot == "hexapod" else octopod_config("tripod_gait")
    h, r, n = engine.host_generate_workspaces(cfg, full=full, max_planes=16)
    for l in range(cfg.leg_count):
        ho, ro = oracle.workspace(cfg, l, full)
        assert n[l] == len(ho) == (12 if full else 1)
        hs, rs = _sorted(h[l], r[l], n[l])
        assert np.abs(hs - ho).max() <= 1e-7 and np.abs(rs - ro).max() <= 1e-7
        if full:  # reach below and above the identity tip, empty limit planes, non-trivial layers in between
            assert ho[0] < -0.02 and ho[-1] > 0.02 and ro[0].max() == 0.0 and ro[-1].max() == 0.0 and ro[1:-1].max() > 0.02
        assert np.array_equal(rs[:, 0], rs[:, 8])  # bearing 0 = bearing 360 (model.cpp:466)
    if not full:  # the engine's own limit tables are built from exactly this plane
        su = engine.compute_startup(cfg)
        assert np.array_equal(np.array(su.workspace)[: cfg.leg_count], r[:, 0, :])


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["hexapod", "octopod"])
def test_device_workspace_sweep(shc_lib, oracle, robot):
    """One block per leg, eight lanes = eight bearings, up to 500 Leg::applyIK(true) steps each: the device sweep gives the
    host sweep's planes (same routine, device arithmetic) and the oracle's to 1e-7."""
    from syropod_highlevel_controller_b200 import engine
    from syropod_highlevel_controller_b200.engine import Engine

    cfg = hexapod_config("tripod_gait") if robot == "hexapod" else octopod_config("tripod_gait")
    eng = Engine(cfg, 32, precision="f64")
    for full in (False, True):
        hd, rd, nd = eng.generate_workspaces(full=full, max_planes=16)
        hh, rh, nh = engine.host_generate_workspaces(cfg, full=full, max_planes=16)
        assert np.array_equal(nd, nh)
        worst = 0.0
        for l in range(cfg.leg_count):
            ho, ro = oracle.workspace(cfg, l, full)
            hs, rs = _sorted(hd[l], rd[l], nd[l])
            worst = max(worst, np.abs(hs - ho).max(), np.abs(rs - ro).max())
            assert np.abs(hd[l] - hh[l]).max() <= 1e-7 and np.abs(rd[l] - rh[l]).max() <= 1e-7
        print(f"[workspace] {robot} {'layered' if full else 'simple'}: device vs oracle {worst:.2e} m")
        assert worst <= 1e-7
    eng.close()


@pytest.mark.gpu
def test_device_direct_startup(shc_lib, oracle):
    """shc_startup_begin / shc_startup_step: every loop() of the direct start-up for a batch whose robots begin at their OWN
    joint angles gives, robot by robot, the joint commands of the oracle's start-up from those angles; at PROGRESS_COMPLETE
    the batch is in the state a new engine starts in (joints to 1e-12: the cubic's last sample keeps ~1e-16 of the origin)."""
    import torch
    from syropod_highlevel_controller_b200.engine import Engine

    for cfg in (hexapod_config("tripod_gait"), octopod_config("tripod_gait", 0.01)):
        L, D = cfg.leg_count, cfg.joint_count
        n = 70
        rng = np.random.default_rng(4)
        lo = np.array([[cfg.joint_min[l][j] for j in range(D)] for l in range(L)])
        hi = np.array([[cfg.joint_max[l][j] for j in range(D)] for l in range(L)])
        q0 = lo + (hi - lo) * rng.uniform(0.05, 0.95, size=(n, L, D))
        eng = Engine(cfg, n, precision="f64")
        fresh = eng.get_state()
        offs = np.array([[cfg.joint_offset[l][j] for j in range(D)] for l in range(L)])
        eng.startup_begin(torch.from_numpy(q0).cuda())
        rows = []
        while True:
            p = eng.startup_step()
            rows.append(eng.joints.cpu().numpy().astype(np.float64) - offs)
            assert 1 <= p <= 100
            if p == 100:
                break
        rows = np.stack(rows)  # [loops, n, L, D]
        for r in (0, 1, 33, n - 1):
            want = oracle.startup_trajectory(cfg, q0[r])
            assert want.shape[0] == rows.shape[0] == round(cfg.time_to_start / cfg.time_delta)
            assert np.abs(rows[:, r] - want).max() < 2e-7  # float32 joint commands
        done = eng.get_state()
        for r in (0, n - 1):
            a, b = done[r], fresh[r]
            for l in range(L):
                assert np.abs(np.array(a.legs[l].joint_position[:D]) - np.array(b.legs[l].joint_position[:D])).max() < 1e-12
                assert list(a.legs[l].tip_position) == list(b.legs[l].tip_position)
            assert a.walk_state == b.walk_state == 3
        # and the one-shot form lands in the same place
        eng.direct_startup(torch.from_numpy(q0).cuda())
        torch.cuda.synchronize()
        assert bytes(eng.get_state()) == bytes(done)
        eng.close()
