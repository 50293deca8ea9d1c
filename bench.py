#!/usr/bin/env python
"""bench.py — throughput of the batched Ruckig calculation (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host cores

A "step" is one pass of Ruckig::calculate over one synthetic batch: `--n` 7-DoF position-interface, Time-synchronised
problems per GPU drawn from the reference bench's input distribution (bench/src/benchmark_target.rs:115-153). The
default batch is BASELINE.json configs[4] sharded over 8 GPUs (64 Mi / 8 = 8 Mi trajectories per GPU), so per-GPU work
is fixed as N grows ("weak" scaling); trajectories are independent, ranks never exchange data.

value  = trajectories/s, inputs and outputs resident in HBM, CUDA events on the library's stream, max over ranks.
e2e    = the same through the C ABI with HOST buffers (pinned): H2D of the nine input arrays, the kernel pipeline, D2H of
         status + duration — all inside the timed region; `e2e.host_path` = what the box's host path can feed (all ranks copying
         pinned memory at once).
roofline: the path is FP64-pipe bound, so `bound` is "fp64": achieved = ALGORITHMIC FP64 operations per
         trajectory (profiles/algorithmic_ops_oracle.json: the oracle compiled with a counting scalar, add/sub/mul/div/sqrt/libm
         call = 1 each, tools/count_algorithmic_ops.py) x trajectories/s; `executed` beside it = the FP64 flop the kernels
         really execute (profiles/algorithmic_ops.json, ncu: adds + muls + 2 x FMAs, divisions and square roots expanded into
         their Newton iterations, candidates evaluated that the reference skips); peak = a DFMA probe run on the same GPU
         right before the timed region (MEASURED_PEAKS.json has no FP64 entry). The HBM side is reported next to it against MEASURED_PEAKS.json's copy
         bandwidth. `latency_64k` (p50 of one 64Ki batch) and `at_time` (configs[3] rollout, HBM-write bound) ride along.
cpu_baseline: the C++ restatement of the reference (oracle/, glibc libm) on this box's host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DOF = 7
METRIC = "7-DoF trajectory calculations/s (position interface, Time sync)"
UNIT = "trajectories/s"
DELTA_TIME = 0.005
BYTES_IN = 9 * DOF * 8                 # nine mandatory f64 arrays
BYTES_OUT = (16 * DOF + 3) * 8         # compact result record of SURVEY.md §8d (the algorithmic figure); RK_RESULT_COMPACT stores 20 doubles + codes per DoF


def algorithmic_flop_per_trajectory() -> tuple[float, str]:
    """operations of the reference algorithm itself, counted by the op-counting build of the oracle."""
    p = os.path.join(ROOT, "profiles", "algorithmic_ops_oracle.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["flop_per_trajectory_7dof"]), d["definition"] + f" ({d['n']} trajectories of the bench distribution, tools/count_algorithmic_ops.py)"
    return 60e3, "a-priori static estimate of SURVEY.md §8d (lower end); the op-counting oracle has not been run"


def executed_flop_per_trajectory() -> tuple[float, str]:
    """What the kernels execute (ncu), as opposed to what the algorithm needs."""
    p = os.path.join(ROOT, "profiles", "algorithmic_ops.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["flop_per_trajectory_7dof"]), d.get("how", "ncu")
    return 0.0, "not measured"


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "_source": "fallback"}


def config(n_per_gpu: int, n_gpus: int, scaling: str = "weak", layout: str = "compact") -> dict:
    if scaling == "strong":
        what = (f"BASELINE.json configs[4] as written: {n_per_gpu * n_gpus} x 7-DoF position interface in total, sharded over {n_gpus} GPU(s) "
                f"({n_per_gpu} per GPU), Time sync, reference bench input distribution")
    else:
        what = (f"{n_per_gpu} x 7-DoF position interface per GPU, Time sync, reference bench input distribution "
                f"(BASELINE.json configs[4]: 64Mi sharded over 8 GPUs -> 8Mi per GPU)")
    return {"workload": what, "n_per_gpu": n_per_gpu, "n_total": n_per_gpu * n_gpus, "dof": DOF, "delta_time": DELTA_TIME,
            "result_layout": f"{layout}: " + ("20 doubles per (DoF, trajectory) stay on the device (t[7], control values, brake pre-trajectory, boundary "
                                              "states; every reader expands them bit-identically), SURVEY.md 8b" if layout == "compact" else
                                              "the reference's Profile verbatim, 59 doubles per (DoF, trajectory)"),
            "l2": f"inputs ({n_per_gpu * BYTES_IN / 1e9:.1f} GB per GPU) exceed the 126 MB L2, no flush needed",
            "sharding": "contiguous index ranges, no collective"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_cores(index: int) -> dict:
    """Pin this rank to the CPU cores (and thereby the NUMA node) `nvidia-smi topo -m` lists for its GPU, BEFORE any pinned host
    buffer is allocated: the staging buffers of the host-buffer path are then first-touched on the GPU's own node, and eight ranks
    do not pull their 4.2 GB per step through the other socket (round 1: end-to-end efficiency 0.52 at N = 8 with unbound ranks)."""
    import re
    info = {"bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        row = next((ln for ln in out.splitlines() if re.match(rf"^GPU{index}\s", ln)), None)
        if row is None:
            return info
        lists = [t for t in re.split(r"\s+", row.strip())[1:] if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t)]
        if not lists:
            return info
        cpus = set()
        for part in lists[0].split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info.update({"bound": True, "cpus": lists[0], "numa": lists[1] if len(lists) > 1 else None, "n_cpus": len(use)})
        else:
            info.update({"cpus": lists[0], "note": "GPU's cores = all allowed cores (single node) or none allowed; nothing to bind"})
    except Exception as e:  # no nvidia-smi, exotic topology output: run unbound
        info["error"] = str(e)[:80]
    return info


def run_reference(args):
    """The reference's CPU algorithm (C++ restatement, glibc libm) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from rsruckig_b200 import workloads
    cores = os.cpu_count() or 1
    threads = cores
    n = args.cpu_sample or 131072 * max(1, min(threads, 64))
    f = workloads.bench_inputs(n, DOF, seed=0)
    for _ in range(args.warmup):
        oracle.bench_calculate(f, n=min(n, 65536), dof=DOF, threads=threads, flavour="glibc", delta_time=DELTA_TIME)
    secs = []
    for _ in range(args.steps):
        s, _, _ = oracle.bench_calculate(f, n=n, dof=DOF, threads=threads, flavour="glibc", delta_time=DELTA_TIME)
        secs.append(s)
    total = sum(secs)
    value = n * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config(args.n, args.gpus, args.scaling, args.layout),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} trajectories of the same distribution per step; C++17 restatement of rsruckig v2.1.2 (oracle/, "
                                       f"g++ -O2 -ffp-contract=off, glibc libm), one reused calculator per thread, static partition over {threads} threads; "
                                       "the Rust reference itself cannot be built here (no cargo/rustc)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(threads_all: int) -> dict:
    """Bounded CPU sample next to the GPU number: single thread and all cores, ~10-20 s of CPU work in total."""
    from oracle import oracle
    from rsruckig_b200 import workloads
    n1 = 131072
    f1 = workloads.bench_inputs(n1, DOF, seed=0)
    oracle.bench_calculate(f1, n=16384, dof=DOF, threads=1, flavour="glibc", delta_time=DELTA_TIME)
    s1, _, _ = oracle.bench_calculate(f1, n=n1, dof=DOF, threads=1, flavour="glibc", delta_time=DELTA_TIME)
    nall = 131072 * max(1, min(threads_all, 64))
    fall = workloads.bench_inputs(nall, DOF, seed=0) if nall != n1 else f1
    sall, _, _ = oracle.bench_calculate(fall, n=nall, dof=DOF, threads=threads_all, flavour="glibc", delta_time=DELTA_TIME)
    return {"value": nall / sall, "unit": UNIT, "cores": threads_all, "kind": "port",
            "single_thread": {"value": n1 / s1, "us_per_trajectory": 1e6 * s1 / n1, "sample": f"{n1} trajectories"},
            "sample": f"{nall} trajectories of the bench distribution, C++17 restatement of rsruckig (oracle/, glibc libm), "
                      f"{threads_all} threads; reported baseline, not the target"}


def gen_inputs_device(torch, n: int, device, seed: int):
    """The reference bench distribution generated on the device (same recipe as workloads.bench_inputs)."""
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    shape = (DOF, n)
    f64 = torch.float64
    normal = lambda std: torch.empty(shape, dtype=f64, device=device).normal_(0.0, std, generator=g)
    uniform = lambda lo, hi: torch.empty(shape, dtype=f64, device=device).uniform_(lo, hi, generator=g)

    def fill_or_zero(p):
        keep = torch.empty(shape, dtype=f64, device=device).uniform_(0.0, 1.0, generator=g) < p
        return torch.where(keep, normal(0.8), torch.zeros((), dtype=f64, device=device))

    f = {"current_position": normal(4.0), "current_velocity": fill_or_zero(0.9), "current_acceleration": fill_or_zero(0.8),
         "target_position": normal(4.0), "target_velocity": fill_or_zero(0.7), "target_acceleration": fill_or_zero(0.6)}
    f["max_velocity"] = uniform(0.1, 12.0) + f["target_velocity"].abs()
    f["max_acceleration"] = uniform(0.1, 12.0) + f["target_acceleration"].abs()
    f["max_jerk"] = uniform(0.1, 12.0)
    return {k: v.contiguous() for k, v in f.items()}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path; use --impl reference for the CPU baseline)"
    if rank == 0:
        entry.build()
    binding = bind_to_gpu_cores(local) if not args.no_bind else {"bound": False, "note": "--no-bind"}
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: what NCCL prints while the communicator comes up (its version banner under
        # NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE) is sent to stderr by pointing fd 1 at fd 2 for the duration
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, _cabi as cabi
    import ctypes as C

    n = args.n
    dev = torch.device("cuda", local)
    layout = cabi.RK_RESULT_COMPACT if args.layout == "compact" else cabi.RK_RESULT_FULL
    otg = BatchRuckig(DOF, DELTA_TIME, ThrowErrorHandler, device=local, result_layout=layout)
    lib = otg._engine.lib
    stream = torch.cuda.ExternalStream(lib.rk_stream(otg.handle), device=dev)

    # FP64 peak probe on this GPU
    peak = C.c_double(0.0)
    cabi.check(lib.rk_measure_fp64_peak(otg.handle, 50.0, C.byref(peak)))
    fp64_peak_tflops = float(peak.value)

    f = gen_inputs_device(torch, n, dev, seed=rank)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    duration = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_device():
        otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE, status=status, duration=duration)

    launches0 = otg.kernel_launches()
    for _ in range(args.warmup):
        step_device()
    otg.sync()
    launches_per_step = (otg.kernel_launches() - launches0) // max(1, args.warmup)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = otg.kernel_launches()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
    otg.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gpu_launches = otg.kernel_launches() - launches1
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = n * world * args.steps / (ms_max * 1e-3)
    n_working = int((status == 0).sum().item())

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + kernels + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        # host buffers are bounded to 8Mi trajectories (4.2 GB pinned): a strong-scaling shard of 64Mi would pin 34 GB of host memory
        n_e = min(n, 8 << 20)
        otg_e = otg if n_e == n else BatchRuckig(DOF, DELTA_TIME, ThrowErrorHandler, device=local, result_layout=layout)
        hf = {k: torch.empty((v.shape[0], n_e), dtype=v.dtype, pin_memory=True) for k, v in f.items()}
        for k in hf:
            hf[k].copy_(f[k][:, :n_e])
        h_status = torch.empty(n_e, dtype=torch.int32, pin_memory=True)
        h_duration = torch.empty(n_e, dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()

        def step_host():
            otg_e.calculate(hf, n=n_e, memory_space=cabi.RK_MEM_HOST, status=h_status, duration=h_duration)

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        k_e2e = max(1, min(args.steps, 3))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            g0.record(stream)
            for _ in range(k_e2e):
                step_host()  # returns when the host outputs are filled
            g1.record(stream)
        otg.sync()
        torch.cuda.synchronize()
        dt = g0.elapsed_time(g1) * 1e-3  # device clock on the library's stream: staging copies, kernels and the read-back
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        assert bool((h_status == status[:n_e].cpu()).all()), "host-buffer and device-buffer runs disagree"
        # what the box's host path can feed: every rank copies pinned host memory to its GPU at the same time (1 GiB x 4); the
        # end-to-end rate cannot exceed what the SLOWEST rank's copy rate allows (the timing is a max over ranks)
        probe_h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        probe_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            probe_d.copy_(probe_h, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        h2d_gbs = 4 * (1 << 30) / (p0.elapsed_time(p1) * 1e-3) / 1e9
        t_sum = torch.tensor([h2d_gbs], dtype=torch.float64, device=dev)
        t_min = t_sum.clone()
        if world > 1:
            dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
            dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
        del probe_h, probe_d
        in_bytes = sum(v.numel() * v.element_size() for v in hf.values())
        host_path = {"h2d_pinned_gbs_aggregate": float(t_sum.item()), "h2d_pinned_gbs_slowest_rank": float(t_min.item()),
                     "ceiling": world * n_e / (in_bytes / (float(t_min.item()) * 1e9)), "unit": UNIT,
                     "note": "all ranks copying pinned host memory to their GPUs concurrently; ceiling = trajectories/s if the step were nothing "
                             "but the slowest rank's input copy (504 B per trajectory)"}
        # a fleet of identical robots: the same current / target states, the five limit arrays shared per DoF (RK_LIMITS_PER_DOF,
        # [dof] instead of [dof][n]): 336 instead of 504 bytes staged per trajectory. Another workload (other limits), reported apart.
        fleet = None
        if not args.no_extras:
            hs = {k: v for k, v in hf.items() if not k.startswith("max_")}
            for k in ("max_velocity", "max_acceleration", "max_jerk"):
                hs[k] = hf[k].max(dim=1).values.contiguous().pin_memory()  # the largest limit of the batch: the inputs stay valid

            def step_fleet():
                otg_e.calculate(hs, n=n_e, memory_space=cabi.RK_MEM_HOST, status=h_status, duration=h_duration, limits_layout=cabi.RK_LIMITS_PER_DOF)

            step_fleet()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                f0.record(stream)
                for _ in range(k_e2e):
                    step_fleet()
                f1.record(stream)
            otg.sync()
            torch.cuda.synchronize()
            t_f = torch.tensor([f0.elapsed_time(f1) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.barrier()
                dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
            fleet = {"value": n_e * world * k_e2e / float(t_f.item()), "unit": UNIT,
                     "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in hs.values())),
                     "working_fraction": float((h_status == 0).float().mean().item()),
                     "note": "same current / target states, max_velocity / max_acceleration / max_jerk shared per DoF (limits_layout = "
                             "RK_LIMITS_PER_DOF: [dof] arrays, the batch's largest values): a fleet of identical robots stages 336 instead of "
                             "504 bytes per trajectory; not the headline workload"}
            step_host()  # the headline workload's results back in place (the status comparison below and the readers after it)
        e2e = {"value": n_e * world * k_e2e / float(t_e.item()), "unit": UNIT, "steps": k_e2e, "n_per_gpu": n_e, "shared_limits": fleet,
               "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in hf.values())), "host_binding": binding, "host_path": host_path,
               "d2h_bytes_per_step": int(h_status.numel() * 4 + h_duration.numel() * 8),
               "note": "rk_calculate_batch with RK_MEM_HOST on pinned buffers; status + duration (12 B per trajectory) come back, the trajectories "
                       "themselves stay device-resident for rk_at_time_batch / the query helpers — a host consumer that wants every Profile pays "
                       "another 20 (compact) or 59 (full) doubles per DoF through rk_trajectories_read, which this number does not include"}

    # ---- p50 latency of one 64Ki batch (the second half of BASELINE.json's metric) and the at_time rollout of configs[3]
    latency = None
    sampling = None
    closed_loop = None
    if not args.no_extras and rank == 0:
        n64 = 65536
        f64k = {k: v[:, :n64].contiguous() for k, v in f.items()}
        otg64 = BatchRuckig(DOF, DELTA_TIME, ThrowErrorHandler, device=local, result_layout=layout)
        # the RL rollout keeps the reference's full Profile (216 MB for 64Ki x 7): at_time then reads phase starts instead of integrating them
        otg_roll = BatchRuckig(DOF, DELTA_TIME, ThrowErrorHandler, device=local, result_layout=cabi.RK_RESULT_FULL)
        otg_roll.calculate(f64k, n=n64, memory_space=cabi.RK_MEM_DEVICE)
        st64 = torch.zeros(n64, dtype=torch.int32, device=dev)
        du64 = torch.zeros(n64, dtype=torch.float64, device=dev)
        run64 = lambda: otg64.calculate(f64k, n=n64, memory_space=cabi.RK_MEM_DEVICE, status=st64, duration=du64)
        for _ in range(5):
            run64()
        otg64.sync()
        times = []
        with torch.cuda.stream(stream):
            for _ in range(50):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                run64()
                b.record(stream)
                b.synchronize()
                times.append(a.elapsed_time(b))
        times.sort()
        latency = {"batch": n64, "p50_ms": times[len(times) // 2], "p99_ms": times[-1], "runs": len(times),
                   "note": "one rk_calculate_batch on device-resident buffers, CUDA events on the library's stream"}
        # 64Ki environments x 7 DoF x K control cycles of Trajectory::at_time (p, v, a): HBM-write bound
        K = args.sample_steps
        out_bytes = 3 * 8 * DOF * n64 * K
        pos = torch.empty((K, DOF, n64), dtype=torch.float64, device=dev)
        vel = torch.empty_like(pos)
        acc = torch.empty_like(pos)
        sample = lambda: otg_roll.at_time(time_start=0.0, delta_time=DELTA_TIME, steps=K, memory_space=cabi.RK_MEM_DEVICE,
                                          position=pos, velocity=vel, acceleration=acc)
        for _ in range(3):
            sample()
        otg64.sync()
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5):
                sample()
            b.record(stream)
            b.synchronize()
        ms_s = a.elapsed_time(b) / 5
        # pure-write reference on the same buffers (torch fill): what a write-only stream reaches on this GPU
        for t_ in (pos, vel, acc):
            t_.zero_()
        torch.cuda.synchronize()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        for _ in range(3):
            for t_ in (pos, vel, acc):
                t_.zero_()
        b2.record()
        b2.synchronize()
        fill_gbs = 3 * out_bytes / (a2.elapsed_time(b2) * 1e-3) / 1e9
        peaks_ = measured_peaks()
        sampling = {"workload": f"{n64} x {DOF}-DoF trajectories (full Profile layout), {K} control cycles of Trajectory::at_time (position, velocity, acceleration)",
                    "ms_per_rollout": ms_s, "samples_per_s": n64 * K / (ms_s * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": out_bytes / (ms_s * 1e-3) / 1e9, "peak": peaks_["hbm_gbs"], "unit": "GB/s",
                                 "frac": out_bytes / (ms_s * 1e-3) / 1e9 / peaks_["hbm_gbs"], "of": peaks_["_source"],
                                 "bytes_per_launch": out_bytes, "write_only_fill_gbs": fill_gbs,
                                 "frac_of_write_only_fill": out_bytes / (ms_s * 1e-3) / 1e9 / fill_gbs, "note": "algorithmic bytes = 3 x 8 B x DoF per sample written; the 472 B profile per (DoF, trajectory) read once is amortised over K"}}
        del pos, vel, acc
        otg_roll.close()
        # closed control loops: rk_update_batch, 64Ki environments, every 10th cycle a tenth of them is
        # re-targeted mid-motion and re-solved from its current state; all advance one cycle and pass the state on
        from rsruckig_b200 import BatchRollout
        ro = BatchRollout(n64, DOF, DELTA_TIME, ThrowErrorHandler, device=local)
        fl = {k: v.clone() for k, v in f64k.items()}
        # environments must start from valid inputs (about 2 % of the raw distribution has a target velocity beyond its limit and
        # would be re-solved, and fail again, every cycle under the throwing handler): such columns are replaced by a valid one
        st0 = torch.zeros(n64, dtype=torch.int32, device=dev)
        otg64.calculate(fl, n=n64, memory_space=cabi.RK_MEM_DEVICE, status=st0)
        otg64.sync()
        okc = st0 == 0
        first_ok = int(torch.nonzero(okc)[0])
        for k in fl:
            fl[k] = torch.where(okc[None, :], fl[k], fl[k][:, first_ok:first_ok + 1]).contiguous()
        new_p = torch.empty((DOF, n64), dtype=torch.float64, device=dev)
        new_v, new_a = torch.empty_like(new_p), torch.empty_like(new_p)
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        cycles, resolved = 200, 0
        ro.update(fl, new_position=new_p, new_velocity=new_v, new_acceleration=new_a)  # first cycle solves everything
        ro.sync()
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for c in range(cycles):
                if c % 10 == 0:
                    pick = torch.rand(n64, generator=gen, device=dev) < 0.1
                    fl["target_position"].copy_(torch.where(pick[None, :], torch.randn((DOF, n64), generator=gen, device=dev, dtype=torch.float64) * 4.0,
                                                            fl["target_position"]))
                resolved += ro.update(fl, new_position=new_p, new_velocity=new_v, new_acceleration=new_a)
            b.record(stream)
            b.synchronize()
        ms_c = a.elapsed_time(b)
        closed_loop = {"workload": f"{n64} x {DOF}-DoF closed control loops, {cycles} cycles of rk_update_batch (Ruckig::update), a tenth of the "
                                   "environments re-targeted every 10th cycle",
                       "cycles_per_s": cycles / (ms_c * 1e-3), "environment_steps_per_s": cycles * n64 / (ms_c * 1e-3),
                       "recalculations": int(resolved), "ms_per_cycle": ms_c / cycles,
                       "note": "includes the torch ops that draw the new targets and one device-to-host read per cycle (the number of environments to "
                               "re-solve); all environments start from valid inputs"}
        del ro

    if rank == 0:
        flop, flop_how = algorithmic_flop_per_trajectory()
        xflop, xflop_how = executed_flop_per_trajectory()
        peaks = measured_peaks()
        per_gpu = value / world
        achieved_tflops = flop * per_gpu / 1e12
        hbm_gbs = (BYTES_IN + BYTES_OUT) * per_gpu / 1e9
        ops = {}
        ops_path = os.path.join(ROOT, "profiles", "algorithmic_ops.json")
        if os.path.exists(ops_path):
            ops = json.load(open(ops_path))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config(n, world, args.scaling, args.layout), "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step),
                "working_fraction": n_working / n, "clocks": clocks, "e2e": e2e, "latency_64k": latency, "at_time": sampling, "closed_loop": closed_loop,
                "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                             "frac": achieved_tflops / fp64_peak_tflops if fp64_peak_tflops else None,
                             "traffic": (ops.get("dram_bytes_per_trajectory") or 0) * n or None,
                             "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum over the {ops.get('kernel_launches_per_step', '?')} kernel launches of one step "
                                             "(ncu, profiles/r2_ops_per_kernel.csv), per GPU; = "
                                             f"{(ops.get('dram_bytes_per_trajectory') or 0) * per_gpu / 1e9:.0f} GB/s at the measured rate",
                             "kernels": "one step per chunk of 4Mi (DoF, trajectory) items = k1_setup_lean, 3 x (k1_quartic_roots, k1_quartic_check), "
                                        "k1_closed_cands, k1_cand_check, k1_block, k1_rescue, k_sync_lanes, 10 chain stages (k2_stage<F>; VEL/UDDU as "
                                        "k2_vel_scan_uddu + k2_polish + k2_vel_check_uddu, VEL/UDUD as k2_vel<8>), k2_tail (the last 8 stages), k_expand_compact; no single "
                                        "kernel holds more than 14 % of the step (profiles/r2_ncu_summary.md), so the roofline is quoted for the whole step",
                             "flop_per_trajectory": flop, "flop_how": flop_how,
                             "executed": {"flop_per_trajectory": xflop, "achieved": xflop * per_gpu / 1e12, "unit": "TFLOP/s",
                                          "frac": xflop * per_gpu / 1e12 / fp64_peak_tflops if fp64_peak_tflops else None, "how": xflop_how},
                             "fp64_issue_frac": (ops.get("fp64_instructions") or 0) * per_gpu / (fp64_peak_tflops * 1e12 / 2) if fp64_peak_tflops else None,
                             "peak_how": "DFMA probe (rk_measure_fp64_peak) on this GPU before the timed region, FMA = 2 flop; the kernels run "
                                         "with --fmad=false, so an add or a mul fills a DFMA issue slot with 1 flop (fp64_issue_frac = FP64 "
                                         "instructions issued / DFMA issue peak)",
                             "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                                     "of": peaks["_source"], "bytes_per_trajectory": BYTES_IN + BYTES_OUT},
                             # the other half of BASELINE.json's metric and the HBM-bound sampling kernel, inside the key the driver keeps
                             "latency_64k": latency,
                             "at_time": (sampling or {}).get("roofline")}}
        if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline_sample(os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8 << 20, help="trajectories per GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --n trajectories per GPU (the default, 8Mi = configs[4]'s shard on 8 GPUs); strong: --total trajectories split over the GPUs")
    ap.add_argument("--total", type=int, default=64 << 20, help="total trajectories of a strong-scaling run (configs[4]: 64Mi)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="trajectories per step of the reference arm (0 = 131072 per thread)")
    ap.add_argument("--layout", default="compact", choices=["compact", "full"],
                    help="device-resident result: compact record (20 doubles per DoF, SURVEY.md 8b/8d) or the reference's full Profile (59)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to the CPU cores / NUMA node of its GPU")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the 64Ki-batch latency and the at_time rollout")
    ap.add_argument("--sample-steps", type=int, default=1000, help="control cycles of the at_time rollout (configs[3])")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.scaling == "strong":
        args.n = args.total // max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
