#!/usr/bin/env python
"""bench.py — control-cycle steps/sec of the batched SHC hot path on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 5                    # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                        # N > 1, one rank per GPU over NCCL
    python bench.py --impl reference --gpus 1 --steps 50 --warmup 5   # the reference path on the host cores

One "step" = one full control cycle (pose + admittance + walk + Bezier tip trajectories + DLS IK) for
every robot of the batch.  Workload: BASELINE.json configs[4]'s per-GPU shard — 131072 hexapods (6 legs x 3 DOF) per
GPU, tripod gait, per-robot synthetic command streams; weak scaling (8 GPUs = the 1,048,576-robot job), one NCCL
all-gather of the joint angles per cycle when N > 1.  State (272 MB per GPU in the f64 parity mode) is larger than
the 126 MB L2, so every step streams it from HBM.

Prints ONE JSON line (rank 0).  `value` = robots x steps / max-over-ranks device time with inputs resident in HBM;
`e2e` = the same metric through the C-ABI host entry point (pinned staging, H2D of the commands, kernel, D2H of the
joint angles every step); `roofline` = algorithmic bytes (2260 B per hexapod step) / kernel time against
the measured HBM copy bandwidth; `cpu_baseline` / `--impl reference` = the reference's OWN control code on this box's host
cores (oracle/_ref: its unmodified sources compiled against stand-in ROS / Eigen / Boost headers, kind "reference"), with
the restated oracle's figure (kind "port", faster: static storage, no shared_ptr / std::map) reported beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "control-cycle steps/sec (6-leg x 3-DOF IK+Bezier)"
UNIT = "steps/s"
ROBOTS_PER_GPU = 131072


def measured_peak():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json when present (the sustained figure if the file
    distinguishes burst / sustained: the kernel is timed inside a loop of back-to-back launches), else the fallback
    B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)

        def find(obj, want, path=""):
            if isinstance(obj, dict):
                for k, v in obj.items():
                    kl = (path + "/" + k).lower()
                    if isinstance(v, (int, float)) and all(w in kl for w in want):
                        return float(v)
                for k, v in obj.items():
                    r = find(v, want, path + "/" + k)
                    if r:
                        return r
            return None

        for want in (("hbm", "sustain"), ("hbm_gbs",), ("hbm", "gb"), ("hbm",), ("copy", "gb")):
            v = find(d, want)
            if v and v > 100.0:
                return v, "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(precision, workload):
    """dram__bytes_read + dram__bytes_write per launch of the control-cycle kernel.  NOT measured in this run: it is read
    from the committed `ncu --set full` capture of the same command (profiles/ncu_summary.json), and labelled so."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(path) as f:
            d = json.load(f)
        key = precision if workload == "hexapod" else f"{workload}_{precision}"
        v = d.get("traffic_bytes_per_launch", {}).get(key)
        return v, (f"static: ncu --set full capture {d.get('capture', '?')} (profiles/ncu_summary.json), not measured in this run"
                   if v is not None else None)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def nvlink_counters(index):
    """NVLink data bytes (tx, rx, raw tx, raw rx) of one GPU summed over its links, from the NVML throughput counters (KiB,
    monotonic); None when the driver does not expose them.  Read on both sides of the timed region at N > 1: the difference
    per cycle is the driver-side evidence of what the fused gather puts on the wire."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ids = [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX,
               pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX]
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(i, 0xFFFFFFFF) for i in ids])  # scope UINT_MAX = all links
        if all(v.nvmlReturn == 0 for v in vals):
            return [int(v.value.ullVal) * 1024 for v in vals]
        out = [0, 0, 0, 0]  # some drivers only answer per link
        seen = False
        for link in range(18):
            vals = pynvml.nvmlDeviceGetFieldValues(h, [(i, link) for i in ids])
            for k, v in enumerate(vals):
                if v.nvmlReturn == 0:
                    out[k] += int(v.value.ullVal) * 1024
                    seen = True
        return out if seen else None
    except Exception:
        return None


def _oracle_rate(cfg, robots, cycles, warm, threads):
    """steps/s of the CPU oracle over `robots` robots: per-robot command streams generated BEFORE the timed region, the
    whole rollout inside the C library (std::threads spawned once, robots partitioned over them)."""
    from oracle import oracle_py as O
    from syropod_highlevel_controller_b200.streams import CommandStream

    ob = O.OracleBatch(cfg, robots)
    cs = CommandStream(robots)
    seq = np.stack([cs.next() for _ in range(warm + cycles)]).astype(np.float64)
    if warm:
        ob.run_seq(seq[:warm], threads)
    secs = ob.run_seq(seq[warm:], threads)
    ob.close()
    return robots * cycles / secs, secs


def _reference_rate(cfg, robots, cycles, warm, threads):
    """steps/s of the REFERENCE'S OWN control code (oracle/_ref/libshc_ref.so: one StateController per robot, each through
    its own direct start-up, untimed): bodyVelocityInputCallback + StateController::loop() per robot and cycle, the whole
    rollout inside the C library (std::threads spawned once, robots partitioned over them)."""
    from oracle import ref_py as R
    from syropod_highlevel_controller_b200.streams import CommandStream

    bots = [R.RefRobot(cfg) for _ in range(robots)]
    cs = CommandStream(robots)
    seq = np.stack([cs.next() for _ in range(warm + cycles)]).astype(np.float64)
    if warm:
        R.batch_run_seq(bots, seq[:warm], threads)
    secs = R.batch_run_seq(bots, seq[warm:], threads)
    single_secs = R.batch_run_seq(bots[:64], np.ascontiguousarray(seq[warm:, :64]), 1)
    for b in bots:
        b.close()
    return robots * cycles / secs, secs, 64 * cycles / single_secs


def _reference_available():
    try:
        from oracle import ref_py as R

        return R.available()
    except Exception:
        return False


REFERENCE_NOTE = ("the reference's own unmodified sources (state_controller / model / walk_controller / pose_controller / "
                  "admittance_controller .cpp) compiled with g++ -O2 against stand-in ROS / Eigen / Boost headers (oracle/shim; "
                  "the stand-in Eigen has no expression templates or SIMD); port_value = the restated oracle (static storage, "
                  "g++ -O3 -march=native), the faster and therefore more conservative CPU figure")


def cpu_baseline(cfg, threads, robots=16384, cycles=100):
    """The reference path on the host cores (bounded sample of the same workload), all cores and one core: the reference's
    own code when oracle/_ref is here (kind "reference"), and the restated oracle (kind "port" / port_value)."""
    value, _ = _oracle_rate(cfg, robots, cycles, 20, threads)
    single, _ = _oracle_rate(cfg, 2048, 100, 20, 1)
    port_sample = (f"{robots} hexapods x {cycles} cycles of the synthetic per-robot command streams (20 warm-up cycles), "
                   f"{threads} std::thread(s) over robots inside the C library; single thread: 2048 robots x 100 cycles")
    if not _reference_available():
        return {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": single,
                "sample": port_sample + "; restated oracle, g++ -O3 -march=native (oracle/_ref is not on this machine)"}
    n_ref = 768
    rv, _, rs = _reference_rate(cfg, n_ref, cycles, 20, threads)
    return {"value": rv, "unit": UNIT, "cores": threads, "kind": "reference", "single_thread_value": rs,
            "port_value": value, "port_single_thread_value": single,
            "sample": f"{n_ref} hexapods (one StateController each) x {cycles} cycles of the synthetic per-robot command streams "
                      f"(20 warm-up cycles), {threads} std::thread(s) over robots; {REFERENCE_NOTE}; port sample: {port_sample}"}


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, same metric / config: the reference's OWN code when
    oracle/_ref is on this machine (kind "reference"; a 1024-robot sample, since every robot is a StateController that has
    to run its own start-up), else the restated oracle (kind "port").  The port is timed either way and reported as
    port_value, on the full per-GPU shard (131072 robots, ~5 GB of oracle state) when the host has the memory for it.  The
    timed region is the in-C loop only: streams are generated beforehand and the worker threads live for the whole rollout."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from syropod_highlevel_controller_b200.config import hexapod_config

    cfg = hexapod_config(args.gait)
    threads = os.cpu_count() or 1
    n = args.robots_per_gpu
    try:
        avail_kb = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1])
        while n > 1024 and n * 48 > 0.5 * avail_kb:  # ~40 KB of oracle state per robot + the command sequence
            n //= 2
    except Exception:
        n = min(n, 16384)
    port_value, port_secs = _oracle_rate(cfg, n, args.steps, args.warmup, threads)
    port_single, _ = _oracle_rate(cfg, 2048, max(args.steps, 20), args.warmup, 1)
    same = n == args.robots_per_gpu
    port_sample = (f"each step = one control cycle over {'the full ' if same else 'a '}{n}-robot "
                   f"{'per-GPU shard' if same else f'sample of the {args.robots_per_gpu}-robot per-GPU shard'}, restated oracle, "
                   f"{threads} threads inside the C library, command streams pre-generated")
    if _reference_available():
        # the reference's own code: one StateController per robot (start-up ~25 ms each, untimed), a bounded sample
        n_ref = 1024
        value, secs, single = _reference_rate(cfg, n_ref, max(args.steps, 20), args.warmup, threads)
        secs = secs * args.steps / max(args.steps, 20)
        kind, same = "reference", False
        sample = (f"each step = one control cycle (bodyVelocityInputCallback + StateController::loop()) over a {n_ref}-robot sample of "
                  f"the {args.robots_per_gpu}-robot per-GPU shard, {threads} threads over robots inside the C library, command "
                  f"streams pre-generated; {REFERENCE_NOTE}; port sample: {port_sample}")
        extra = {"port_value": port_value, "port_single_thread_value": port_single, "port_ms_per_step": port_secs / args.steps * 1e3}
    else:
        value, secs, single, kind, sample, extra = port_value, port_secs, port_single, "port", port_sample, {}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config5 shard: {args.robots_per_gpu} hexapods/GPU, {args.gait}, default.yaml", "sample": sample,
                       "same_config": same},
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                                  "single_thread_value": single}, **extra),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f64", choices=["f64", "mixed"], help="f64 = parity mode (default)")
    ap.add_argument("--robots-per-gpu", type=int, default=ROBOTS_PER_GPU)
    ap.add_argument("--gait", default="tripod_gait")
    ap.add_argument("--workload", default="hexapod", choices=["hexapod", "octopod"],
                    help="hexapod = the headline workload (configs[4] shard); octopod = configs[3]: 8 legs x 5 DOF with "
                         "admittance + IMU + inclination posing, 262144 robots (secondary line, same JSON shape)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-path", default="auto", choices=["auto", "symm", "symm-unicast", "ipc"],
                    help="fused gather: auto = torch symmetric memory with NVSwitch multicast when available, else CUDA IPC")
    ap.add_argument("--no-verify-gather", action="store_true", help="N > 1: skip the bit-exact check of the gathered shards")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: fused = the kernel stores joint angles into every rank's buffer over NVLink (peer memory); "
                         "nccl = library-issued ncclAllGather per cycle on a side stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from syropod_highlevel_controller_b200.config import hexapod_config, octopod_config
    from syropod_highlevel_controller_b200.engine import Engine
    from syropod_highlevel_controller_b200.parallel import shard_robots
    from syropod_highlevel_controller_b200.streams import CommandStream, ForceStream, ImuStream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SHC engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    octo = args.workload == "octopod"
    n = args.robots_per_gpu if not octo or args.robots_per_gpu != ROBOTS_PER_GPU else 262144
    shard = shard_robots(n * world, rank, world)
    cfg = octopod_config(args.gait) if octo else hexapod_config(args.gait)
    L, D = cfg.leg_count, cfg.joint_count
    K, W = args.steps, args.warmup

    eng = Engine(cfg, n, device=local_rank, precision=args.precision)
    # synthetic per-robot command streams, resident in HBM before the timed region
    cs = CommandStream(n, robot_offset=shard.offset)
    pre = 300  # untimed pre-roll so that the batch is in its steady mix of walk states (STARTING/MOVING/STOPPING/STOPPED)
    cmd_host = np.stack([cs.next() for _ in range(pre + W + K)])
    cmd_dev = torch.from_numpy(cmd_host).to(dev)
    # configs[3]: IMU orientation / gyro and tip forces are per-cycle inputs too (a short cycle of distinct frames)
    imu_dev = force_dev = None
    if octo:
        ims, fs = ImuStream(n, robot_offset=shard.offset), ForceStream(n, L, robot_offset=shard.offset)
        imu_dev = torch.from_numpy(np.stack([ims.next(cfg.time_delta) for _ in range(8)])).to(dev)
        force_dev = torch.from_numpy(np.stack([fs.next() for _ in range(8)])).to(dev)
    # N > 1: the per-cycle all-gather of the joint angles.  fused (default): the kernel stores every tile into every rank's
    # buffer over NVLink (NVSwitch multicast when the fabric has it); nccl: library-issued ncclAllGather on a side stream
    gather_mode = None
    if world > 1:
        if args.gather == "fused":
            try:
                eng.init_gather_fused(rank, world, mode=args.gather_path)
                gather_mode = eng.gather_mode
            except RuntimeError as ex:  # raised on every rank together: no peer-memory path on this box
                if rank == 0:
                    print(f"bench.py: {ex}; falling back to --gather nccl", file=sys.stderr)
                args.gather = "nccl"
        if args.gather != "fused":
            eng.init_nccl(rank, world)
            local2 = torch.empty((2, n, L, D), dtype=torch.float32, device=dev)
            full2 = torch.empty((2, n * world, L, D), dtype=torch.float32, device=dev)
            gather_mode = "NCCL all-gather per cycle on a side stream"

    def inputs(i):
        return cmd_dev[i], None if imu_dev is None else imu_dev[i & 7], None if force_dev is None else force_dev[i & 7]

    def run(lo, hi):
        if world == 1:
            if octo:  # per-cycle IMU / tip-force frames: one C-ABI call per cycle
                for i in range(lo, hi):
                    eng.step(*inputs(i))
            else:     # commands only: the whole rollout is one C-ABI call (shc_rollout: the k launches as one CUDA graph)
                eng.rollout(cmd_dev[lo:hi])
        elif args.gather == "fused":
            if octo:  # per-cycle IMU / tip-force frames: one C-ABI call per cycle, then the wait for the last one
                for i in range(lo, hi):
                    eng.gather_step(*inputs(i))
                eng.gather_sync()
            else:     # commands only: the whole rollout is one C-ABI call (k x shc_gather_step + shc_gather_sync)
                eng.rollout_gather_fused(cmd_dev[lo:hi])
        else:
            if octo:
                raise SystemExit("bench.py: --workload octopod with N > 1 needs --gather fused (the NCCL rollout carries commands only)")
            eng.rollout_allgather(cmd_dev[lo:hi], local2, full2)

    run(0, pre + W)
    if world == 1 and not octo:
        # shc_rollout instantiates a CUDA graph per (k, buffers): do that for the timed call's arguments outside the timed
        # region (K more untimed cycles on the same commands; the timed region then replays them from the state reached)
        run(pre + W, pre + W + K)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    nvl0 = nvlink_counters(local_rank) if world > 1 else None  # (a slow host call: before the ranks are aligned)
    if world > 1:
        dist.barrier()
        # the host-side barrier above lets the ranks' processes leave up to a millisecond apart; a stream-ordered 1-element
        # all-reduce makes the GPUs themselves start the timed region together (a rank's region cannot end before the
        # slowest rank's last shard has landed, so start skew would be billed to the fastest rank)
        align = torch.zeros(1, device=dev)
        dist.all_reduce(align)
    ev0.record()
    run(pre + W, pre + W + K)
    ev1.record()
    torch.cuda.synchronize()
    nvl1 = nvlink_counters(local_rank) if world > 1 else None
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = n * world * K / (ms_total * 1e-3)

    # N > 1: what landed in this rank's gather buffer for the last cycle must be, bit for bit, what a plain NCCL all_gather
    # of every rank's own shard gives (driver-side correctness evidence of the fused exchange)
    gather_ok = None
    if world > 1 and args.gather == "fused" and not args.no_verify_gather:
        b = eng.gather_sync()
        torch.cuda.synchronize()
        eng.gather_status()
        dist.barrier()
        got = eng.gather[b]  # [world, n, L, D]
        own = got[rank].clone()
        want = torch.empty_like(got)
        dist.all_gather_into_tensor(want.view(world * n, L, D), own)
        okt = torch.tensor([1 if torch.equal(got, want) and bool(torch.isfinite(got).all()) and float(got.abs().max()) > 0 else 0],
                           device=dev, dtype=torch.int32)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_ok = bool(int(okt.item()))

    # Average launch duration of the control-cycle kernel for the roofline.  N = 1: the timed region above IS K launches of
    # that kernel on this stream and nothing else, so its CUDA-event time / K is the figure (launch gaps included).  N > 1: the
    # region also holds the exchange, so K more launches without it are timed the same way.
    if world == 1:
        kernel_ms, kernel_ms_src = ms / K, "timed region (K launches of the kernel, CUDA events on the launching stream)"
    else:
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if octo:
            k0.record()
            for i in range(pre + W, pre + W + K):
                eng.step(*inputs(i))
            k1.record()
        else:  # commands only: the K launches as one CUDA graph (no host launch gaps between them)
            eng.rollout(cmd_dev[pre + W:pre + W + K])  # capture + first replay, untimed
            torch.cuda.synchronize()
            k0.record()
            eng.rollout(cmd_dev[pre + W:pre + W + K])
            k1.record()
        torch.cuda.synchronize()
        kernel_ms, kernel_ms_src = k0.elapsed_time(k1) / K, "K launches without the exchange, after the timed region"
    peak, peak_src = measured_peak()
    b_alg = eng.bytes_per_step_algorithmic
    achieved = n * b_alg / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.precision, args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "control_cycle_kernel", "kernel_ms": kernel_ms, "kernel_ms_source": kernel_ms_src,
                "algorithmic_bytes_per_step": b_alg, "device_bytes_per_step": eng.bytes_per_step_device,
                "device_bytes_GBps": n * eng.bytes_per_step_device / (kernel_ms * 1e-3) / 1e9, "peak_source": peak_src,
                "nominal_peak": 8000.0, "frac_of_nominal": achieved / 8000.0}  # also quote the nominal ~8 TB/s

    # end to end through the C-ABI host entry point: every step moves that step's commands up from page-locked host
    # memory and reads the joint angles back into page-locked host memory, inside the timed region
    e2e_steps = max(3, min(K, 20))
    host_cmd = eng.pinned_host(e2e_steps, n, 3)
    host_cmd[:] = cmd_host[pre + W: pre + W + e2e_steps]
    host_out = eng.pinned_host(n, L, D)
    host_imu = host_force = None
    if octo:
        host_imu, host_force = eng.pinned_host(n, 10), eng.pinned_host(n, L, 3)
        host_imu[:] = imu_dev[0].cpu().numpy()
        host_force[:] = force_dev[0].cpu().numpy()
    eng.step_host(host_cmd[0], host_imu, host_force, out=host_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.step_host(host_cmd[i], host_imu, host_force, out=host_out)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": n * world * e2e_steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": n * (3 + (10 + 3 * L if octo else 0)) * 4,
           "d2h_bytes_per_step": n * L * D * 4, "steps": e2e_steps,
           "api": "shc_step_host (C-ABI, page-locked host buffers: every step the kernel's warps read their robots' velocity "
                  "commands from the caller's buffer over PCIe (IMU / tip-force records go up by a host-to-device copy first), "
                  "one kernel launch whose TMA bulk stores write each tile's joint angles straight into the caller's "
                  "page-locked buffer over PCIe, stream sync)",
           "gpu_launches_per_step": 1}

    io_bytes = (3 + (10 + 3 * L if octo else 0) + L * D) * 4
    state_mb = n * (eng.bytes_per_step_device - io_bytes) / 2 / 1e6
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": (f"configs[3]: {n} octopods/GPU (8 legs x 5 DOF), admittance + IMU + inclination posing, {args.gait}, "
                                    f"per-robot command / IMU / tip-force streams" if octo else
                                    f"config5 shard: {n} hexapods/GPU (6 legs x 3 DOF), {args.gait}, default.yaml parameters, "
                                    f"per-robot splitmix64 command streams") +
                                   (f"; all-gather of the joint angles every cycle: {gather_mode}" if world > 1 else ""),
                       "robots_per_gpu": n, "robots_total": n * world, "precision": args.precision,
                       "l2": f"state {state_mb:.0f} MB per GPU > 126 MB L2 (inputs larger than L2, no flush)",
                       "pre_roll_cycles": pre,
                       "call": ("one C-ABI call per cycle (shc_step)" if octo and world == 1 else
                                "one C-ABI call for the K cycles (shc_rollout: K launches replayed as one CUDA graph, instantiated by an "
                                "untimed call on the same arguments)" if world == 1 else
                                "one C-ABI call for the K cycles (shc_rollout_gather_fused)" if args.gather == "fused" and not octo else
                                "per-cycle C-ABI calls")},
            "gpu_launches": K, "e2e": e2e, "roofline": roofline, "clocks": clocks}

    if gather_ok is not None:
        line["gather_ok"] = gather_ok
        # control cycle per step; landed signal every 8th cycle and at the sync; reuse checks every 8th cycle; final wait
        line["gpu_launches"] = K + (K + 7) // 8 + 1 + (K + 7) // 8 + 1
    if nvl0 is not None and nvl1 is not None:
        d = [(b - a) / K for a, b in zip(nvl0, nvl1)]
        line["nvlink"] = {"tx_bytes_per_step": d[0], "rx_bytes_per_step": d[1], "raw_tx_bytes_per_step": d[2],
                          "raw_rx_bytes_per_step": d[3], "shard_bytes": n * L * D * 4,
                          "source": "NVML NVLINK_THROUGHPUT_DATA/RAW counters of rank 0's GPU, all links, over the timed region"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(cfg, os.cpu_count() or 1)
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU measurement
            line["cpu_baseline"] = {"error": str(ex)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
