#!/usr/bin/env python
"""bench.py -- env-steps/sec (step + teacher reward + render) of the batched XWorld2D engine.

Workload (BASELINE.json `metric`: "... at 64k envs"; configs[2]): XWorld2D walls.json rules,
11x11 spanning-tree maze, 4 goals, 30 blocks, 84x84x3 uint8 observations, 65536 envs PER GPU
(weak scaling), i.i.d. uniform actions, auto-reset on (max_steps = 2*H*W = 242), synthetic icons.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this engine
    python bench.py --impl reference [...]                         # CPU arm: the oracle port on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU, NCCL

One JSON line on rank 0 (contract in the task description): value = whole-job env-steps/s with
inputs resident in HBM; e2e = the same metric through the host-buffer C ABI (xw_step_hd: host actions
in, host reward + game_over out, frames left in HBM for a co-located learner) with the copies inside
the timed region; roofline = the render kernel's algorithmic bytes / its CUDA-event time vs the
measured HBM peak; cpu_baseline = the oracle port timed on this box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ENVS_PER_GPU = 65536
# The default ("c3") is the configuration BASELINE.json's metric is quoted on (64k envs, 11x11 maze, 84x84 frames);
# --workload c2 / c4 time the other two XWorld2D configurations of BASELINE.json with the same contract;
# --workload c5 the SimpleRace step kernel at 1,048,576 envs (BASELINE configs[4], render off).
WORKLOADS = {
    "c3": dict(cfg=dict(height=11, width=11, n_goals=4, n_blocks=30, rules=1, out_h=84, out_w=84, max_steps=242,
                        auto_reset=1, seed=1234, simulator_seed=1),
               name="XWorld2D walls.json rules, 11x11 maze, 84x84x3 u8 obs, 65536 envs/GPU (BASELINE configs[2])",
               map="11x11", side=84, rules="walls.json", max_steps=242, envs=65536),
    "c2": dict(cfg=dict(height=7, width=7, n_goals=4, n_blocks=12, rules=0, out_h=84, out_w=84, auto_reset=1, seed=1234,
                        simulator_seed=1),
               name="XWorld2D navigation2d.json rules, 7x7 map, 84x84x3 u8 obs (BASELINE configs[1] at bench scale)",
               map="7x7", side=84, rules="navigation2d.json", max_steps=0, envs=65536),
    "c4": dict(cfg=dict(height=15, width=15, n_goals=4, n_blocks=56, rules=0, out_h=128, out_w=128, auto_reset=1, seed=1234,
                        simulator_seed=1),
               name="XWorld2D navigation2d.json rules, 15x15 map, 128x128x3 u8 obs, 32768 envs/GPU (BASELINE configs[3])",
               map="15x15", side=128, rules="navigation2d.json", max_steps=0, envs=32768),
    # BASELINE configs[2]'s map in the only 84x84 view the reference itself produces for it: the first-person
    # view with --visible_radius 7 (xworld_simulator.cpp:62-68), six actions
    "fpv": dict(cfg=dict(height=11, width=11, n_goals=4, n_blocks=30, rules=1, visible_radius=7, max_steps=242, auto_reset=1,
                         seed=1234, simulator_seed=1),
                name="XWorld2D walls.json rules, 11x11 maze, first-person view visible_radius=7 -> 84x84x3 u8 obs, 65536 envs/GPU",
                map="11x11", side=84, rules="walls.json", max_steps=242, envs=65536),
}
WORKLOAD = WORKLOADS["c3"]["cfg"]
WORKLOAD_NAME = WORKLOADS["c3"]["name"]
# 3*84*84 frame write + 121 grid bytes + 28 bytes of action/reward/game_over/agent state
BYTES_PER_ENV_STEP = 3 * 84 * 84 + 11 * 11 + 28
SIDE = 84


def select_workload(key):
    global WORKLOAD, WORKLOAD_NAME, BYTES_PER_ENV_STEP, SIDE
    w = WORKLOADS[key]
    WORKLOAD, WORKLOAD_NAME, SIDE = w["cfg"], w["name"], w["side"]
    BYTES_PER_ENV_STEP = 3 * SIDE * SIDE + w["cfg"]["height"] * w["cfg"]["width"] + 28
    return w


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.stamps, self.proc, self.gpu = [], [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        """Starts nvidia-smi (20 ms period) and waits for its first row, so that it is already sampling when the timed
        region begins (the region of the default run is ~60 ms: shorter than nvidia-smi's start-up)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            end = time.perf_counter() + 3.0
            while not self.rows and time.perf_counter() < end:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.stamps.append(time.perf_counter())
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        rows, where = self.rows, "whole run"
        if self.t0 is not None and self.t1 is not None:
            # a row printed at time t describes the ~20 ms before it: rows stamped inside [t0, t1 + one period]
            inside = [r for r, t in zip(self.rows, self.stamps) if self.t0 <= t <= self.t1 + 0.03]
            if inside:
                rows, where = inside, "timed region"
            elif self.rows:  # region shorter than a period: the rows on either side of it
                k = min(range(len(self.rows)), key=lambda i: abs(self.stamps[i] - self.t1))
                rows, where = self.rows[max(0, k - 1):k + 2], "rows adjacent to the timed region"
        self.rows = rows
        self.where = where
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": getattr(self, "where", "whole run")}


def cpu_reference_arm(steps, warmup, sample_envs, threads):
    """The reference's CPU path for this workload = the oracle port (the reference cannot link OpenCV
    here): step + teacher + the full canvas/resize render pipeline, all host threads."""
    import numpy as np
    import oracle
    from xworld_b200 import _abi
    from xworld_b200.catalog import Catalog
    cfg = _abi.default_config(**WORKLOAD)
    cat = Catalog.synthetic(seed=0)
    orc = oracle.Oracle(cfg, cat, sample_envs, threads=threads)
    orc.reset()
    rng = np.random.RandomState(0)
    acts = [rng.randint(0, 6 if cfg.visible_radius > 0 else 4, sample_envs).astype(np.int32) for _ in range(8)]
    for s in range(warmup):
        orc.step(acts[s % 8], render=True)
    t0 = time.perf_counter()
    for s in range(steps):
        orc.step(acts[s % 8], render=True)
    dt = time.perf_counter() - t0
    return sample_envs * steps / dt, dt


RACE_BYTES_PER_ENV_STEP = 60  # pos_x, pos_y, angle, steps read + written (32) + action 4 + reward 4 + game_over 4 + state 16
RACE_NAME = "SimpleRace straight track, easy, 2 actions, fp32 step kernel, 1,048,576 envs/GPU, render off (BASELINE configs[4])"


def race_cpu_arm(sample_envs, steps):
    """The C port of simple_race_simulator.cpp (bit-equal to the compiled reference, tests/test_oracle_pins.py), one thread."""
    import numpy as np
    import oracle
    from xworld_b200 import _abi
    L = oracle.lib()
    cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_RACE)
    envs = (oracle.XoRace * sample_envs)()
    for o in envs:
        L.xo_race_reset(C.byref(cfg), C.byref(o))
    rng = np.random.RandomState(0)
    acts = np.ascontiguousarray(rng.randint(0, 2, (steps, sample_envs)), np.int32)
    L.xo_race_batch.restype = C.c_double
    L.xo_race_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    t0 = time.perf_counter()
    L.xo_race_batch(C.byref(cfg), envs, sample_envs, acts.ctypes.data, steps)
    dt = time.perf_counter() - t0
    return sample_envs * steps / dt, dt


def measure_race(n, steps, warmup, rank, world, local_rank, dist, dev, sampler=None):
    """SimpleRace at n envs per GPU: K timed xw_step launches with inputs in HBM, then the same through xw_step_hd.
    Returns (ms of this rank, ms max over ranks, launches, total env-steps, per-rank records, e2e seconds max over ranks)."""
    import torch
    from xworld_b200 import _abi
    from xworld_b200.sharding import gather_throughput
    from xworld_b200.simulator import Simulator
    cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_RACE, auto_reset=1)
    sim = Simulator("simple_race", cfg, None, n, local_rank)
    lib, h = sim._lib, sim._h
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 + rank)
    acts = [torch.randint(0, 2, (n,), dtype=torch.int32, device=dev, generator=gen) for _ in range(8)]
    reward = torch.zeros(n, dtype=torch.float32, device=dev)
    over = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sim.reset_game()

    def step(i):
        rc = lib.xw_step(h, acts[i % 8].data_ptr(), 1, reward.data_ptr(), over.data_ptr(), None, stream)
        if rc:
            raise RuntimeError(lib.xw_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    launches0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_steps, _, per_rank = gather_throughput(n * steps, int(ms * 1e6), device=dev)
    h_act = [torch.randint(0, 2, (n,), dtype=torch.int32).pin_memory() for _ in range(4)]
    h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
    h_over = torch.zeros(n, dtype=torch.int32).pin_memory()
    for i in range(5):
        lib.xw_step_hd(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), None)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        assert lib.xw_step_hd(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), None) == 0
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    # K steps per launch (xw_step_seq: open-loop action sequences, the car in registers between the steps)
    K = 32
    reps = max(2, steps // K)
    seq_a = torch.randint(0, 2, (K, n), dtype=torch.int32, device=dev, generator=gen)
    seq_r = torch.zeros((K, n), dtype=torch.float32, device=dev)
    seq_o = torch.zeros((K, n), dtype=torch.int32, device=dev)
    lib.xw_step_seq(h, seq_a.data_ptr(), K, 1, seq_r.data_ptr(), seq_o.data_ptr(), stream)
    barrier()
    e0.record()
    for i in range(reps):
        assert lib.xw_step_seq(h, seq_a.data_ptr(), K, 1, seq_r.data_ptr(), seq_o.data_ptr(), stream) == 0
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    measure_race.multi_step = {"value": n * world * K * reps / (float(t.item()) / 1e3), "unit": "env-steps/s", "steps_per_launch": K,
                               "launches": reps, "bytes_per_env_step": 12,
                               "what": "xw_step_seq: 32 take_actions calls per launch (actions / reward / game_over [32][n] in HBM), "
                                       "the car in registers between the steps"}
    del sim
    return ms, ms_max, launches, total_steps, per_rank, float(tt.item())


def race_fields(n, steps, world, ms, ms_max, total_steps, e2e_s):
    """value / roofline / e2e of a SimpleRace measurement (the kernel is the whole step)."""
    peak, peak_src = hbm_peak()
    kernel_ms = ms / steps
    achieved = RACE_BYTES_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e9
    return {
        "value": total_steps / (ms_max / 1e3), "ms_per_step": ms_max / steps,
        "roofline": {"bound": "hbm", "kernel": "k_race_step", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": kernel_ms,
                     "kernel_share_of_step": 1.0, "algorithmic_bytes_per_launch": RACE_BYTES_PER_ENV_STEP * n,
                     "note": "one launch per step (the agent's actions arrive per step); the 63 MB of state stay in L2 between "
                             "steps, so this is an upper bound on DRAM traffic, not a bandwidth-bound kernel: launch + tail "
                             "latency sets the time at this size"},
        "multi_step": getattr(measure_race, "multi_step", None),
        "e2e": {"value": n * world * steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * n,
                "d2h_bytes_per_step": 8 * n, "steps": steps,
                "what": "xw_step_hd: pinned host actions -> H2D, k_race_step, reward + game_over D2H, stream sync"}}


def bench_race(args, rank, world, local_rank):
    """--workload c5: BASELINE configs[4].  One k_race_step launch per step; the kernel is the whole step."""
    config = {"workload": RACE_NAME, "envs_per_gpu": args.envs_per_gpu, "auto_reset": True, "actions": "iid uniform{0,1}",
              "l2": "per step the kernel touches 60 B x 1,048,576 envs = 63 MB of state (< 126 MB L2): state stays L2-resident "
                    "between steps by design -- the roofline figure is against HBM and is an upper bound on DRAM traffic",
              "bytes_per_env_step": RACE_BYTES_PER_ENV_STEP}
    if args.impl == "reference":
        if rank != 0:
            return 0
        v, dt = race_cpu_arm(65536, max(1, min(args.steps, 200)))
        print(json.dumps({"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
                          "steps": min(args.steps, 200), "warmup": 0, "ms_per_step": None, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                           "sample": "65536 envs x %d steps, C port, one thread" % min(args.steps, 200)},
                          "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    import torch
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: nothing but the warm-up may sit between it and the timed region
    ms, ms_max, launches, total_steps, per_rank, e2e_s = measure_race(n, args.steps, args.warmup, rank, world, local_rank, dist, dev,
                                                                      sampler=sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    f = race_fields(n, args.steps, world, ms, ms_max, total_steps, e2e_s)
    line = {"metric": "env_steps_per_sec", "value": f["value"], "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": f["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": f["roofline"], "gpu_launches": launches, "clocks": clocks, "per_rank": per_rank, "e2e": f["e2e"],
            "multi_step": f["multi_step"]}
    if world == 1 and not args.no_cpu_baseline:
        v, dt = race_cpu_arm(65536, 200)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                "sample": "65536 envs x 200 steps (%.1f s), the C port (bit-equal to the compiled reference), one thread" % dt}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cv2_faithful_arm(sample_envs, steps):
    """BASELINE.md §4's "faithful call sequence" leg: the reference's render pipeline with the real OpenCV, one thread --
    per frame a white canvas, per ITEM cv::warpAffine + copyTo (xitem.cpp:47-60, xmap.cpp:129-146: every wall brick, every
    frame), the identity cv::resize + HWC->CHW repack (xworld_simulator.cpp:287-307), the CHW->HWC repack + cv::resize +
    HWC->CHW repack (:508-545); the step / teacher comes from the oracle port.  None when cv2 is not importable."""
    try:
        import cv2
    except ImportError:
        return None
    import math
    import numpy as np
    import oracle
    from xworld_b200 import _abi
    from xworld_b200.catalog import Catalog
    cv2.setNumThreads(1)
    cfg = _abi.default_config(**WORKLOAD)
    cat = Catalog.synthetic(seed=0)
    orc = oracle.Oracle(cfg, cat, sample_envs, threads=1)
    orc.reset()
    H, W = cfg.height, cfg.width
    oh, ow = orc.out_h, orc.out_w
    rng = np.random.RandomState(0)
    n_act = 6 if cfg.visible_radius > 0 else 4
    rot = cv2.getRotationMatrix2D((32.0, 32.0), 90 - 1.5707963 * 180 / math.pi, 1.0)

    def frame(e):
        world = np.full((H * 64, W * 64, 3), 255, np.uint8)
        for i in range(H):
            for j in range(W):
                code = e.grid[i * W + j]
                if code == 0:
                    continue
                icon = cat.brick_icon if code == 1 else cat.agent_icon if code == 2 else e.goal_icon[code - 3]
                img = cv2.warpAffine(cat.atlas64[icon].copy(), rot, (64, 64), flags=cv2.INTER_LINEAR,
                                     borderMode=cv2.BORDER_CONSTANT, borderValue=(255, 255, 255))
                world[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64] = img
        screen = cv2.resize(world, (W * 64, H * 64), interpolation=cv2.INTER_LINEAR)
        planar = np.ascontiguousarray(screen.transpose(2, 0, 1))
        img = np.ascontiguousarray(planar.transpose(1, 2, 0))
        out = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
        return np.ascontiguousarray(out.transpose(2, 0, 1))

    t0 = time.perf_counter()
    for s in range(steps):
        orc.step(rng.randint(0, n_act, sample_envs).astype(np.int32), render=False)
        for e in orc.envs:
            frame(e)
    dt = time.perf_counter() - t0
    return sample_envs * steps / dt, dt


def cpu_legs(threads):
    """cpu_baseline: the all-core oracle port (the figure), plus the legs BASELINE.md §4 asks for."""
    import numpy as np
    import oracle
    from xworld_b200 import _abi
    from xworld_b200.catalog import Catalog
    sample, ksteps = 1024, 20
    v, dt = cpu_reference_arm(ksteps, 2, sample, threads)
    out = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
           "sample": "%d envs x %d steps of the same workload (%.1f s), oracle C port (integer restatement of the reference's "
                     "OpenCV pipeline), OpenMP over envs on all host threads" % (sample, ksteps, dt)}
    v1, dt1 = cpu_reference_arm(6, 1, 128, 1)
    out["one_core"] = {"value": v1, "unit": "env-steps/s", "cores": 1, "sample": "128 envs x 6 steps (%.1f s), the same C port, one thread" % dt1}
    cfg = _abi.default_config(**WORKLOAD)
    orc = oracle.Oracle(cfg, Catalog.synthetic(seed=0), 4096, threads=1)
    orc.reset()
    rng = np.random.RandomState(0)
    n_act = 6 if cfg.visible_radius > 0 else 4
    acts = [rng.randint(0, n_act, 4096).astype(np.int32) for _ in range(4)]
    t0 = time.perf_counter()
    for i in range(50):
        orc.step(acts[i % 4], render=False)
    dt2 = time.perf_counter() - t0
    out["step_teacher_only"] = {"value": 4096 * 50 / dt2, "unit": "env-steps/s", "cores": 1,
                                "sample": "4096 envs x 50 steps, no render: move + collision + the teacher's rules as fixed C "
                                          "(the reference runs them in embedded CPython through Boost.Python, which cannot run "
                                          "on this box: no Python 2.7 / Boost; SURVEY §6 estimates 10^2-10^3 steps/s/core for it)"}
    cv = cv2_faithful_arm(16, 2)
    out["cv2_faithful"] = None if cv is None else {
        "value": cv[0], "unit": "env-steps/s", "cores": 1,
        "sample": "16 envs x 2 steps (%.1f s): real OpenCV (cv2), the reference's call sequence incl. one warpAffine per item "
                  "per frame and the repacks (vectorised), one thread" % cv[1]}
    return out


def lib_sha256():
    import hashlib
    from xworld_b200 import _abi
    with open(_abi.LIB_PATH, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def src_sha256():
    """sha256 over the sources the library is built from (xworld_b200/csrc/* and the C header, names sorted): identifies the
    build even when nvcc is run again (its output is not byte-reproducible)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "xworld_b200", "csrc")
    files = [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(ROOT, "include", "xworld_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def traffic_for(key, n):
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the render kernel from the ncu --set full capture
    recorded in profiles/render_traffic.json -- only if it was taken on THIS build (sha256 of libxworld_b200.so, or of the sources
    it is built from when the library has been re-linked) at this size."""
    tp = os.path.join(ROOT, "profiles", "render_traffic.json")
    if not os.path.exists(tp):
        return None, "no capture on record"
    with open(tp) as f:
        rec = json.load(f)
    ent = rec.get("workloads", {}).get(key)
    if not ent or ent.get("envs") != n:
        return None, "no capture of this workload / size on record"
    if rec.get("lib_sha256") != lib_sha256() and rec.get("src_sha256") != src_sha256():
        return None, "the capture on record (%s) was taken on another build of libxworld_b200.so" % rec.get("capture")
    return ent["dram_bytes_per_launch"], rec.get("capture")


KERNEL_NAMES = {3: "k_render_sp", 1: "k_render_sb", 2: "k_render", 0: "k_render_generic", 4: "k_render_fpv_generic", 5: "k_render_fpv",
                6: "k_render_fpv_cells"}


def run_xworld(key, n, steps, warmup, rank, world, local_rank, dist, dev, sampler=None, e2e_mode="full"):
    """One XWorld2D workload on this rank's GPU: K timed steps with inputs in HBM (CUDA events, barrier on both sides, max
    over ranks), the render / step+reset kernel times from the library's own events, and the end-to-end loop through the
    host-buffer C ABI.  Returns the fields of the JSON line (rank 0) or None."""
    import numpy as np
    import torch
    from xworld_b200 import _abi
    from xworld_b200.catalog import Catalog
    from xworld_b200.sharding import gather_throughput
    from xworld_b200.simulator import Simulator
    wl = select_workload(key)
    cfg = _abi.default_config(**WORKLOAD)
    cfg.env_id_offset = rank * n  # contiguous global env ids: the union of ranks is one logical batch
    sim = Simulator("xworld", cfg, Catalog.synthetic(seed=0), n, local_rank)
    lib, h = sim._lib, sim._h
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    n_act = sim.get_num_actions()
    acts = [torch.randint(0, n_act, (n,), dtype=torch.int32, device=dev, generator=gen) for _ in range(8)]
    frames = sim.screen()
    reward = torch.zeros(n, dtype=torch.float32, device=dev)
    over = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sim.reset_game()
    dephased = None
    if wl["max_steps"] > 0:
        # Episodes of the walls.json rules end by --max_steps only: started together, all 65,536 envs would reset in the same
        # step once every 242 and never inside a short timed region.  Start them out of phase instead (env i has already
        # played (i * 97) mod 242 steps), so that every step re-generates n / 242 maps: the steady state of a long run.
        ns = ((np.arange(n, dtype=np.int64) + rank * n) * 97 % wl["max_steps"]).astype(np.int32)
        sim.set_field("num_steps", ns)
        dephased = "num_steps of env i starts at (i * 97) mod %d: ~%d of the %d envs end an episode and are re-generated in every step" % (
            wl["max_steps"], n // wl["max_steps"], n)

    def step(i):
        rc = lib.xw_step(h, acts[i % 8].data_ptr(), 1, reward.data_ptr(), over.data_ptr(), frames.data_ptr(), stream)
        if rc:
            raise RuntimeError(lib.xw_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    ep0 = int(sim.get_field("episode").astype(np.int64).sum())
    lib.xw_enable_timing(h, 1)
    lib.xw_render_ms(h, 1)
    lib.xw_step_reset_ms(h, 1)
    launches0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count() - launches0
    render_ms = lib.xw_render_ms(h, 1)
    step_reset_ms = lib.xw_step_reset_ms(h, 1)
    lib.xw_enable_timing(h, 0)
    resets = int(sim.get_field("episode").astype(np.int64).sum()) - ep0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_steps, _, per_rank = gather_throughput(n * steps, int(ms * 1e6), device=dev)
    value = total_steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer C ABI (pinned host actions in, host reward/over out)
    e2e = None
    if e2e_mode != "none":
        h_act = [torch.randint(0, n_act, (n,), dtype=torch.int32).pin_memory() for _ in range(4)]
        h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
        h_over = torch.zeros(n, dtype=torch.int32).pin_memory()
        k2 = max(10, steps)  # the same K steps as the device-resident measurement
        for i in range(5):
            lib.xw_step_hd(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), frames.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for i in range(k2):
            rc = lib.xw_step_hd(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), frames.data_ptr())
            assert rc == 0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * k2 / float(tt.item()), "unit": "env-steps/s", "h2d_bytes_per_step": 4 * n,
               "d2h_bytes_per_step": 8 * n + 4, "steps": k2,
               "what": "xw_step_hd: pinned host actions -> H2D, step+reset+render kernels, reward+game_over (+ the invalid-action "
                       "count) D2H, stream sync; frames stay in HBM (consumer = co-located learner)"}
    if e2e_mode == "full":
        # the pipelined form of the same call (xw_step_hd_async): the host still waits for every step's reward / game_over
        # before it issues the next step, but not for the frames, which a co-located learner consumes in stream order
        # (two frame buffers, alternating); everything is complete (xw_sync) inside the timed region
        frames2 = torch.empty_like(frames)
        fb = [frames, frames2]
        for i in range(5):
            lib.xw_step_hd_async(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), fb[i & 1].data_ptr())
        lib.xw_sync(h)
        barrier()
        t0 = time.perf_counter()
        for i in range(k2):
            rc = lib.xw_step_hd_async(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), fb[i & 1].data_ptr())
            assert rc == 0
        lib.xw_sync(h)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["pipelined"] = {"value": n * world * k2 / float(tt.item()), "unit": "env-steps/s", "steps": k2,
                            "what": "xw_step_hd_async: the same copies and the same host wait for reward + game_over every step; "
                                    "the frames of step t are complete in stream order (xw_wait_frames), not waited for by the host"}
        del frames2, fb
        if rank == 0 and world == 1:  # frames to the host too (PCIe-bound), reported beside it
            hf = torch.empty((n, 3, SIDE, SIDE), dtype=torch.uint8).pin_memory()
            for i in range(2):
                lib.xw_step_host(h, h_act[0].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), hf.data_ptr())
            t0 = time.perf_counter()
            k3 = 5
            for i in range(k3):
                lib.xw_step_host(h, h_act[i % 4].data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), hf.data_ptr())
            dt = time.perf_counter() - t0
            e2e["with_frames_to_host"] = {"value": n * k3 / dt, "unit": "env-steps/s",
                                          "d2h_bytes_per_step": 8 * n + n * 3 * SIDE * SIDE}
            del hf
    kernel = KERNEL_NAMES.get(sim.render_kernel(), "k_render")
    del sim
    if rank != 0:
        return None
    peak, peak_src = hbm_peak()
    achieved = BYTES_PER_ENV_STEP * n / (render_ms * 1e-3) / 1e9 if render_ms and render_ms > 0 else None
    traffic, traffic_src = traffic_for(key, n)
    return {
        "value": value, "ms_per_step": ms_max / steps, "steps": steps, "warmup": warmup,
        "config": {"workload": WORKLOAD_NAME, "map": wl["map"], "obs": "%dx%dx3 u8 (B,G,R planes)" % (SIDE, SIDE), "rules": wl["rules"],
                   "envs_per_gpu": n, "auto_reset": True, "max_steps": wl["max_steps"],
                   "actions": "iid uniform{0..%d}" % (n_act - 1), "dephased": dephased,
                   "l2": "each step writes %.2f GB of frames per GPU (>> 126 MB L2), so no L2 flush is needed" % (n * 3 * SIDE * SIDE / 1e9),
                   "bytes_per_env_step": BYTES_PER_ENV_STEP},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel_ms": render_ms, "step_reset_ms": step_reset_ms,
                     "resets_per_step": resets / float(steps),
                     "kernel_share_of_step": (render_ms / (ms / steps)) if render_ms else None,
                     "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n},
        "gpu_launches": launches, "per_rank": per_rank, "e2e": e2e,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="default: the workload's (65536; c4: 32768)")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs in the default line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    if args.workload == "c5":
        if args.envs_per_gpu <= 0:
            args.envs_per_gpu = 1 << 20
        return bench_race(args, rank, world, local_rank)
    wl = select_workload(args.workload)
    if args.envs_per_gpu <= 0:
        args.envs_per_gpu = wl["envs"]

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = 1024
        steps = max(1, args.steps)
        steps = min(steps, 40)  # bounded: ~2-3k frames/s on 8 cores
        v, dt = cpu_reference_arm(steps, min(args.warmup, 3), sample, threads)
        config = {"workload": WORKLOAD_NAME, "map": wl["map"], "obs": "%dx%dx3 u8 (B,G,R planes)" % (SIDE, SIDE), "rules": wl["rules"],
                  "envs_per_gpu": args.envs_per_gpu, "auto_reset": True, "max_steps": wl["max_steps"],
                  "bytes_per_env_step": BYTES_PER_ENV_STEP,
                  "sample": "%d envs per step (bounded sample of the %d-env workload)" % (sample, args.envs_per_gpu)}
        line = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": dt / steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                 "sample": "%d envs x %d steps, OpenMP over envs" % (sample, steps)},
                "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: nothing but the warm-up may sit between it and the timed region
    res = run_xworld(args.workload, args.envs_per_gpu, args.steps, args.warmup, rank, world, local_rank, dist, dev,
                     sampler=sampler if rank == 0 else None, e2e_mode="none" if args.no_e2e else "full")
    clocks = sampler.stop() if rank == 0 else None
    # ---- the other BASELINE configurations, a few steps each, in the same line (configs[1], configs[3] at 32,768 envs per GPU
    #      = 262,144 over 8 GPUs, the first-person view of configs[2]'s map, configs[4]); every rank runs them, rank 0 reports
    others = {}
    if args.workload == "c3" and not args.no_configs:
        for key in ("c2", "c4", "fpv"):
            r = run_xworld(key, WORKLOADS[key]["envs"], 40, 8, rank, world, local_rank, dist, dev,
                           e2e_mode="none" if args.no_e2e else "short")
            if rank == 0:
                others[key] = {"value": r["value"], "unit": "env-steps/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                               "workload": r["config"]["workload"], "envs_per_gpu": r["config"]["envs_per_gpu"],
                               "roofline": r["roofline"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"]}
        ms5, ms5_max, l5, tot5, _, e2e5 = measure_race(1 << 20, 100, 10, rank, world, local_rank, dist, dev)
        if rank == 0:
            f5 = race_fields(1 << 20, 100, world, ms5, ms5_max, tot5, e2e5)
            others["c5"] = {"value": f5["value"], "unit": "env-steps/s", "ms_per_step": f5["ms_per_step"], "steps": 100,
                            "workload": RACE_NAME, "envs_per_gpu": 1 << 20, "dtype": "f32", "roofline": f5["roofline"], "e2e": f5["e2e"],
                            "multi_step": f5["multi_step"], "gpu_launches": l5}
        select_workload(args.workload)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    line = {
        "metric": "env_steps_per_sec", "value": res["value"], "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": res["config"], "roofline": res["roofline"],
        "gpu_launches": res["gpu_launches"], "clocks": clocks, "per_rank": res["per_rank"],
    }
    if res["e2e"]:
        line["e2e"] = res["e2e"]
    if others:
        line["configs"] = others
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_legs(threads)
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
