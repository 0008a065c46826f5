"""GPU: size-independent properties at BASELINE.json's full batch sizes (where the oracle would take minutes), the
host/graph entry points, the stand-alone IK kernel and the status flags."""
import ctypes as C
import os

import numpy as np
import pytest

from syropod_highlevel_controller_b200.config import ShcConfig, hexapod_config, octopod_config
from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

pytestmark = pytest.mark.gpu


def _engine(cfg, n, precision="f64", startup=None):
    from syropod_highlevel_controller_b200.engine import Engine

    return Engine(cfg, n, precision=precision, startup=startup)


@pytest.mark.parametrize("precision", ["f64", "mixed"])
def test_batch_invariance_at_65536(shc_lib, precision):
    """BASELINE configs[2] size: a robot's result depends only on its own state and inputs, not on the batch size or its
    position in the batch (the property that makes sharding across GPUs exact).  Bit-identical."""
    import torch

    cfg = hexapod_config("tripod_gait")
    n_big, ids = 65536, np.array([0, 1, 31, 32, 4095, 4096, 20000, 65535])
    big = _engine(cfg, n_big, precision)
    small = _engine(cfg, len(ids), precision, startup=big.startup())
    cs = CommandStream(n_big, min_len=20, max_len=60)
    for c in range(150):
        cmd = cs.next()
        jb = big.step(torch.from_numpy(cmd).cuda())
        js = small.step(torch.from_numpy(cmd[ids]).cuda())
        assert torch.equal(jb[torch.from_numpy(ids).cuda()], js), c
    sb, ss = big.get_state(), small.get_state()
    for k, r in enumerate(ids):
        assert bytes(sb[int(r)]) == bytes(ss[k])
    big.close(); small.close()


def test_octopod_262144_runs_and_matches_small_batch(shc_lib):
    """BASELINE configs[3] size: 262144 octopods with admittance + IMU + inclination posing."""
    import torch

    cfg = octopod_config("tripod_gait")
    n_big, ids = 262144, np.array([0, 77, 131071, 262143])
    big = _engine(cfg, n_big, "f64")
    small = _engine(cfg, len(ids), "f64", startup=big.startup())
    cs, ims, fs = CommandStream(n_big, min_len=10, max_len=40), ImuStream(n_big), ForceStream(n_big, 8)
    sel = torch.from_numpy(ids).cuda()
    for c in range(40):
        cmd, imu, force = cs.next(), ims.next(cfg.time_delta), fs.next()
        jb = big.step(torch.from_numpy(cmd).cuda(), torch.from_numpy(imu).cuda(), torch.from_numpy(force).cuda())
        js = small.step(torch.from_numpy(cmd[ids]).cuda(), torch.from_numpy(imu[ids]).cuda(), torch.from_numpy(force[ids]).cuda())
        assert torch.equal(jb[sel], js), c
    assert torch.isfinite(jb).all()
    big.close(); small.close()


def test_standstill_and_symmetry_at_scale(shc_lib):
    """Zero command: the walk state stays STOPPED and stepper tips never move; a common forward command keeps tripod
    legs {0,2,4} / {1,3,5} half a period apart on every robot."""
    import torch

    cfg = hexapod_config("tripod_gait")
    n = 32768
    eng = _engine(cfg, n, "f64")
    s0 = eng.get_state()
    zero = torch.zeros((n, 3), device="cuda")
    j0 = None
    for c in range(30):
        j = eng.step(zero)
        j0 = j.clone() if j0 is None else j0
    s1 = eng.get_state()
    for r in (0, 1, 12345, n - 1):
        assert s1[r].walk_state == 3
        for l in range(6):
            assert list(s1[r].legs[l].tip_position) == list(s0[r].legs[l].tip_position)
    assert (j - j0).abs().max() < 6e-3  # joints only move inside the reference's stand-still limit cycle
    fwd = torch.tensor([[0.5, 0.0, 0.0]], device="cuda").repeat(n, 1)
    for c in range(300):
        eng.step(fwd)
    s2 = eng.get_state()
    su = eng.startup()
    for r in (0, 7, 9999, n - 1):
        ph = [s2[r].legs[l].phase for l in range(6)]
        assert ph[0] == ph[2] == ph[4] and ph[1] == ph[3] == ph[5] and (ph[1] - ph[0]) % su.period == su.period // 2
        assert s2[r].walk_state == 1
    eng.close()


def test_state_round_trip_is_idempotent(shc_lib):
    import torch

    for cfg, prec in ((hexapod_config("wave_gait"), "f64"), (octopod_config("tripod_gait"), "mixed")):
        n = 300
        eng = _engine(cfg, n, prec)
        cs = CommandStream(n, min_len=20, max_len=60)
        for c in range(120):
            eng.step(torch.from_numpy(cs.next()).cuda())
        a = eng.get_state()
        eng.set_state(a)
        b = eng.get_state()
        assert bytes(a) == bytes(b)
        eng.close()


def test_state_range_and_limit_maps(shc_lib):
    """shc_get_state_range returns the bytes shc_get_state returns for the same robots (tile boundaries included),
    shc_set_state_range writes ragged robot ranges through their tiles, and
    shc_set_limit_maps (WalkController::set*LimitMap) takes effect from the next cycle: with the linear-speed table halved,
    the body velocity of a full-throttle robot settles at half the original limit."""
    import torch

    cfg = hexapod_config("tripod_gait")
    n = 1000
    eng = _engine(cfg, n, "f64")
    cs = CommandStream(n, min_len=20, max_len=60)
    for c in range(90):
        eng.step(torch.from_numpy(cs.next()).cuda())
    full = eng.get_state()
    for first, count in ((0, 1), (31, 2), (32, 32), (500, 77), (999, 1), (0, 1000)):
        part = eng.get_state_range(first, count)
        assert bytes(part) == bytes(full)[first * C.sizeof(full[0]):(first + count) * C.sizeof(full[0])], (first, count)
    # shc_set_state_range: a second engine takes the first one's state in ragged pieces and ends up in the same state
    twin = _engine(cfg, n, "f64")
    for first, count in ((0, 5), (5, 60), (65, 31), (96, 904)):
        twin.set_state_range(first, eng.get_state_range(first, count))
    cmd = torch.from_numpy(cs.next()).cuda()
    assert torch.equal(eng.step(cmd).clone(), twin.step(cmd).clone())
    st_a, st_b = eng.get_state(), twin.get_state()
    assert bytes(st_a) == bytes(st_b)
    twin.close()
    su = eng.startup()
    lim = np.array(list(su.max_linear_speed))
    fwd = torch.tensor([[1.0, 0.0, 0.0]], device="cuda").repeat(n, 1)
    for c in range(400):
        eng.step(fwd)
    v_full = np.array([eng.get_state_range(r, 1)[0].desired_linear_velocity[0] for r in (0, 499, 999)])
    assert np.allclose(v_full, lim[0], rtol=1e-9)
    eng.set_limit_maps(max_linear_speed=0.5 * lim)
    assert list(eng.startup().max_linear_speed) == list(0.5 * lim)
    for c in range(400):
        eng.step(fwd)
    v_half = np.array([eng.get_state_range(r, 1)[0].desired_linear_velocity[0] for r in (0, 499, 999)])
    assert np.allclose(v_half, 0.5 * lim[0], rtol=1e-9)
    eng.close()


def test_host_entry_point_and_graph_rollout_match_step(shc_lib):
    """shc_step_host (pinned staging + H2D + kernel + D2H) and shc_rollout (CUDA graph) give the device path's bits."""
    import torch

    cfg = octopod_config("tripod_gait")
    n, k = 1000, 24
    a, b, c_ = _engine(cfg, n, "f64"), None, None
    b = _engine(cfg, n, "f64", startup=a.startup())
    c_ = _engine(cfg, n, "f64", startup=a.startup())
    cs, ims, fs = CommandStream(n, min_len=5, max_len=20), ImuStream(n), ForceStream(n, 8)
    cmds = np.stack([cs.next() for _ in range(k)])
    imus = np.stack([ims.next(cfg.time_delta) for _ in range(k)])
    forces = np.stack([fs.next() for _ in range(k)])
    for i in range(k):
        ja = a.step(torch.from_numpy(cmds[i]).cuda(), torch.from_numpy(imus[i]).cuda(), torch.from_numpy(forces[i]).cuda())
        jb = b.step_host(cmds[i], imus[i], forces[i])
        assert np.array_equal(ja.cpu().numpy(), jb), i
    jc = c_.rollout(torch.from_numpy(cmds).cuda(), torch.from_numpy(imus).cuda(), torch.from_numpy(forces).cuda())
    torch.cuda.synchronize()
    assert torch.equal(jc, ja)
    assert bytes(a.get_state()) == bytes(c_.get_state())
    for e in (a, b, c_):
        e.close()


def test_apply_ik_kernel_matches_oracle(shc_lib, oracle):
    """Stand-alone batched Leg::applyIK (model.cpp:861) — the §8(f) workspace-sweep building block."""
    L_ = oracle.lib()
    dp = C.POINTER(C.c_double)
    L_.shc_oracle_apply_ik.restype = C.c_double
    L_.shc_oracle_apply_ik.argtypes = [C.POINTER(ShcConfig), C.c_int, dp, dp, dp, C.c_int, dp]
    rng = np.random.default_rng(9)
    for cfg in (hexapod_config(), octopod_config()):
        D, L = cfg.joint_count, cfg.leg_count
        eng = _engine(cfg, 1, "f64")
        m = 600
        leg = rng.integers(0, L, m)
        q = np.zeros((m, D)); qd = rng.normal(0, 1.0, (m, D)); des = np.zeros((m, 3))
        for i in range(m):
            lo = np.array([cfg.joint_min[leg[i]][j] for j in range(D)]); hi = np.array([cfg.joint_max[leg[i]][j] for j in range(D)])
            q[i] = lo + (hi - lo) * rng.uniform(0.02, 0.98, D)
            des[i] = oracle.fk(cfg, int(leg[i]), q[i]) + rng.normal(0, 0.006, 3)
        for sim in (True, False):
            qg, qdg, tipg, resg = (t.cpu().numpy() for t in eng.apply_ik(leg, q, qd, des, simulation=sim))
            for i in range(m):
                qo, qdo, tipo = q[i].copy(), qd[i].copy(), np.empty(3)
                ro = L_.shc_oracle_apply_ik(C.byref(cfg), int(leg[i]), qo.ctypes.data_as(dp), qdo.ctypes.data_as(dp),
                                            np.ascontiguousarray(des[i]).ctypes.data_as(dp), int(sim), tipo.ctypes.data_as(dp))
                assert np.abs(qg[i] - qo).max() < 1e-11 and np.abs(tipg[i] - tipo).max() < 1e-12
                assert resg[i] == pytest.approx(ro, abs=1e-11)
        eng.close()


def test_status_flags_match_oracle_and_do_not_change_results(shc_lib, oracle):
    import torch
    from syropod_highlevel_controller_b200.engine import FLAG_IK_DEVIATION, OPT_STATUS_FLAGS

    cfg = hexapod_config("tripod_gait")
    n = 256
    ob = oracle.OracleBatch(cfg, n)
    plain = _engine(cfg, n, "f64", startup=ob.startup())
    flagged = _engine(cfg, n, "f64", startup=ob.startup())
    flagged.set_options(OPT_STATUS_FLAGS)
    cs = CommandStream(n, min_len=40, max_len=150)
    seen_dev = 0
    for c in range(500):
        cmd = cs.next()
        cmd[: n // 4] *= np.array([1.0, 1.0, 1.0], dtype=np.float32)
        ja = plain.step(torch.from_numpy(cmd).cuda())
        jb = flagged.step(torch.from_numpy(cmd).cuda())
        ob.step(cmd.astype(np.float64), threads=8)
        assert torch.equal(ja, jb)
        if c % 10 == 9:
            fl = flagged.status_flags()
            st = ob.get_state()
            for r in range(n):
                dev = any(max(abs(st[r].legs[l].model_tip_position[k] - st[r].legs[l].desired_tip_position[k]) for k in range(3)) > 0.005
                          for l in range(6))
                margin = min(abs(max(abs(st[r].legs[l].model_tip_position[k] - st[r].legs[l].desired_tip_position[k]) for k in range(3)) - 0.005)
                             for l in range(6))
                if margin > 1e-7:  # away from the 5 mm threshold itself
                    assert bool(fl[r] & FLAG_IK_DEVIATION) == dev, (c, r)
                seen_dev += dev
    assert seen_dev > 0  # the random full-throttle streams do push some legs past IK_TOLERANCE
    for e in (plain, flagged):
        e.close()
    ob.close()


@pytest.mark.gpu
def test_fused_gather_two_gpus():
    """N > 1: the kernel's peer-memory stores deliver every rank's joint angles to every rank, bit for
    bit what a plain all_gather of the per-rank results gives.  Needs two GPUs on the box; runs under torchrun."""
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "dev_p2p_check.py"), "2065"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("-> OK") >= 2 and "MISMATCH" not in out.stdout


def test_host_entry_point_tile_ranges_and_pinned_buffers(shc_lib):
    """shc_step_host issues a large batch as 8 tile ranges (D2H of one overlapping the kernel of the next) and moves
    page-locked caller buffers by DMA in place: same bits as one device-resident launch, pinned or pageable, and the tail
    tile (n not a multiple of 32) is handled."""
    import torch

    cfg = hexapod_config("tripod_gait")
    n, k = 20011, 12
    a = _engine(cfg, n, "f64")
    b = _engine(cfg, n, "f64", startup=a.startup())
    c_ = _engine(cfg, n, "f64", startup=a.startup())
    cs = CommandStream(n, min_len=3, max_len=8)
    pin_cmd, pin_out = b.pinned_host(n, 3), b.pinned_host(n, 6, 3)
    for i in range(k):
        cmd = cs.next()
        ja = a.step(torch.from_numpy(cmd).cuda()).cpu().numpy()
        pin_cmd[:] = cmd
        jb = b.step_host(pin_cmd, out=pin_out)
        jc = c_.step_host(cmd)
        assert jb is pin_out
        assert np.array_equal(ja, jb), i
        assert np.array_equal(ja, jc), i
    assert bytes(a.get_state()) == bytes(b.get_state()) == bytes(c_.get_state())
    for e in (a, b, c_):
        e.close()


def test_clone_reconfigured_gait_change_against_the_reference(shc_lib):
    """shc_clone_reconfigured (Engine.reconfigured): a batch-wide gait switch on the B200 against the reference's own
    StateController::changeGait (oracle/_ref), walking before and after — the case of tests/test_emu_parity.py through the
    C-ABI.  Also: a model mismatch is refused and the source engine stays usable."""
    import test_emu_parity as T
    from backends import Backend
    from syropod_highlevel_controller_b200.config import octopod_config
    from syropod_highlevel_controller_b200.engine import ShcError

    def switch(stepper, cfg, keep_pose_cycle):
        new = stepper.__class__.__new__(stepper.__class__)
        new.torch, new.n = stepper.torch, stepper.n
        new.eng = stepper.eng.reconfigured(cfg, keep_pose_cycle=keep_pose_cycle)  # the engine's own constants for the new parameters
        with pytest.raises(ShcError):  # another model is refused, the source engine stays usable
            stepper.eng.reconfigured(octopod_config() if stepper.eng.L == 6 else hexapod_config())
        return new

    T._gait_change_case(Backend("gpu"), switch, n=6)
    # a new step cycle is refused while a robot is walking (the reference would re-phase its legs); constants-only changes pass
    import torch

    eng = _engine(hexapod_config("tripod_gait"), 64)
    fwd = torch.tensor([[0.6, 0.0, 0.1]], device="cuda").repeat(64, 1)
    for c in range(150):
        eng.step(fwd)
    with pytest.raises(ShcError, match="not STOPPED"):
        eng.reconfigured(hexapod_config("wave_gait"))
    with pytest.raises(ShcError, match="not STOPPED"):
        eng.reconfigured(hexapod_config("tripod_gait", step_frequency=1.4))
    eng.reconfigured(hexapod_config("tripod_gait", swing_height=0.03)).close()
    eng.close()
    # constants-only parameters and a step-frequency change at rest, against the reference's adjustParameter
    for model, base, change, at_rest, lag, first_cmd in T.PARAMETER_CHANGES[:1] + T.PARAMETER_CHANGES[3:7]:
        T._parameter_change_case(Backend("gpu"), switch, model, base, change, at_rest, lag, first_cmd, n=4)
