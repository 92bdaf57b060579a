#!/usr/bin/env python3
"""bench.py — the contract benchmark (one JSON line on stdout).

A *step* is one full Monte Carlo localisation update on one scan, the cycle of
laser_handler.cpp:243-276 of the reference:

    handle_odometry -> update_importance_from_observations ->
    resample_and_cluster -> get_estimated_pose

over the workload named by BASELINE.json (default C2: 100 k particles, 2000 x
2000 map at 5 cm, 1081-beam 270 deg scan, KLD resampling; every beam is used,
num_beams = 0).  Before each step the particle set is restored to the same
initial cloud (untimed), so every step does identical work.

metric  particle*point likelihood evaluations per second over the whole update
value   device-timed, scan already resident in HBM when the step starts
e2e     the same step through the public API with HOST buffers: the scan is
        copied host->device and the estimate device->host inside the timed
        region
roofline  the sensor-model kernel (K7), timed with CUDA events inside the
        library on the launching stream, 4.02 algorithmic bytes per evaluation
        (SURVEY.md 8d), against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the CPU oracle (a restatement of the reference's single-threaded
        implementation; the reference itself needs ROS/Eigen/Boost/PCL and
        cannot be built here) timed on this host, one thread
--impl reference  times that CPU implementation on the same workload and
        prints the same line with "impl": "reference"

Multi-GPU (torchrun, one rank per GPU): particles are sharded in contiguous
global-index blocks; see quickmcl_b200/sharded.py.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from quickmcl_b200 import synthetic as syn  # noqa: E402

ALG_BYTES_PER_EVAL = 4.02  # one float32 field cell + amortised pose/weight
ODOM_OLD = (0.0, 0.0, 0.0)
ODOM_NEW = (0.25, 0.05, 0.1)

WORKLOADS = {
    # name: particles, map, beams, fov, resample, cloud, pf overrides
    "C1": dict(particles=2000, map=400, beams=360, fov=360.0, cloud="tracking",
               pf=dict(resample_type=0, particle_count_min=2000, particle_count_max=2000)),
    "C2": dict(particles=100_000, map=2000, beams=1081, fov=270.0, cloud="tracking",
               pf=dict(resample_type=2, particle_count_min=100, particle_count_max=100_000,
                       kld_epsilon=0.05, kld_z=0.95)),
    "C3": dict(particles=1_000_000, map=4000, beams=1081, fov=270.0, cloud="global",
               pf=dict(resample_type=1, particle_count_min=1_000_000,
                       particle_count_max=1_000_000, alpha_fast=0.1, alpha_slow=0.001)),
    "C4": dict(particles=16_000_000, map=4000, beams=1081, fov=270.0, cloud="global",
               pf=dict(resample_type=0, particle_count_min=16_000_000,
                       particle_count_max=16_000_000)),
    # "C4 low_variance and kld at 16 M" — the KLD variant, not a default of any N
    "C4K": dict(particles=16_000_000, map=4000, beams=1081, fov=270.0, cloud="global",
                pf=dict(resample_type=2, particle_count_min=100, particle_count_max=16_000_000,
                        kld_epsilon=0.05, kld_z=0.95)),
}
RESAMPLE_NAMES = {0: "low_variance", 1: "adaptive", 2: "kld"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_inputs(wl):
    occ, res, origin = syn.make_map(wl["map"])
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, wl["beams"], wl["fov"])
    n = wl["particles"]
    if wl["cloud"] == "tracking":
        x, y, t = syn.gaussian_cloud(pose, n)
    else:
        x, y, t = syn.uniform_cloud(occ, res, origin, n)
    return occ, res, origin, pose, scan, (x, y, t)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm
def cpu_cycle_seconds(wl, inputs, steps, warmup):
    """Times the CPU oracle's full update cycle; returns (seconds per step, evals per step,
    per-stage seconds)."""
    from oracle import binding as ob  # the one place bench.py executes oracle/
    occ, res, origin, pose, scan, (x, y, t) = inputs
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_BRUSHFIRE)  # the reference's own field
    of = ob.OracleFilter(om, seed=3)
    pfp = dict(wl["pf"])
    of.set_parameters(**pfp)
    n = len(x)
    init = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    stages = {"odometry": 0.0, "sensor": 0.0, "resample": 0.0, "estimate": 0.0}
    total = 0.0
    for it in range(warmup + steps):
        of.set_particles(init)
        of.set_rng(3, 10 + it)
        t0 = time.perf_counter()
        of.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
        t1 = time.perf_counter()
        of.update_importance_from_observations(scan)
        t2 = time.perf_counter()
        of.resample_and_cluster()
        t3 = time.perf_counter()
        of.get_estimated_pose()
        t4 = time.perf_counter()
        if it >= warmup:
            total += t4 - t0
            stages["odometry"] += t1 - t0
            stages["sensor"] += t2 - t1
            stages["resample"] += t3 - t2
            stages["estimate"] += t4 - t3
    evals = n * len(scan)
    return total / steps, evals, {k: v / steps for k, v in stages.items()}


def cpu_all_cores(wl, inputs, steps=3):
    """The same CPU cycle on every host core at once: the particle set is cut into one
    slice per thread, each slice runs its own full cycle (own filter, shared read-only
    map).  This is NOT how the reference runs (it is single-threaded, main.cpp:90, and
    slices resample independently); it bounds what the host could do if it were
    parallelised over particles.  Returns a dict for the JSON line."""
    import threading
    from oracle import binding as ob
    occ, res, origin, pose, scan, (x, y, t) = inputs
    threads_n = max(1, min(os.cpu_count() or 1, 64))
    n = len(x)
    per = n // threads_n
    if per < 64:
        return None
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_BRUSHFIRE)
    pfp = dict(wl["pf"])
    sc = per / n
    pfp["particle_count_min"] = max(1, int(pfp["particle_count_min"] * sc))
    pfp["particle_count_max"] = max(1, int(pfp["particle_count_max"] * sc))
    filters = []
    for k in range(threads_n):
        of = ob.OracleFilter(om, seed=3 + k)
        of.set_parameters(**pfp)
        sl = slice(k * per, (k + 1) * per)
        filters.append((of, ob.make_particles(x[sl], y[sl], t[sl], np.full(per, 1.0 / per))))
    barrier = threading.Barrier(threads_n + 1)

    def work(of, init):
        for it in range(steps + 1):
            of.set_particles(init)
            of.set_rng(3, 10 + it)
            if it == 1:
                barrier.wait()   # the warm-up step is over everywhere: the clock starts
            of.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
            of.update_importance_from_observations(scan)
            of.resample_and_cluster()
            of.get_estimated_pose()

    ts = [threading.Thread(target=work, args=fi) for fi in filters]
    for th in ts:
        th.start()
    barrier.wait()
    t0 = time.perf_counter()
    for th in ts:
        th.join()
    sec = (time.perf_counter() - t0) / steps
    evals = threads_n * per * len(scan)
    return {"value": evals / sec, "unit": "evals/s", "cores": threads_n,
            "sample": f"{threads_n} threads x {per} particles x {len(scan)} beams, independent "
                      f"full cycles per slice, {steps} steps after 1 warm-up",
            "note": "not reference behaviour (the reference is single-threaded); an upper bound "
                    "for a particle-parallel host"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    inputs = make_inputs(wl)
    # bounded sample: a slice of the particle set so that K+W steps end within minutes
    n_full = wl["particles"]
    budget_evals = 2.5e9  # ~15-30 s of single-thread CPU work in total
    per_step = budget_evals / max(1, args.steps + args.warmup)
    n_sample = int(min(n_full, max(1000, per_step // max(1, len(inputs[4])))))
    occ, res, origin, pose, scan, (x, y, t) = inputs
    wl_s = dict(wl)
    pf = dict(wl["pf"])
    scale = n_sample / n_full
    if scale < 1.0:
        pf["particle_count_min"] = max(1, int(pf["particle_count_min"] * scale))
        pf["particle_count_max"] = max(1, int(pf["particle_count_max"] * scale))
    wl_s["pf"] = pf
    sec, evals, stages = cpu_cycle_seconds(
        wl_s, (occ, res, origin, pose, scan, (x[:n_sample], y[:n_sample], t[:n_sample])),
        args.steps, args.warmup)
    value = evals / sec
    all_cores = cpu_all_cores(
        wl_s, (occ, res, origin, pose, scan, (x[:n_sample], y[:n_sample], t[:n_sample])))
    sample = (f"{n_sample} of {n_full} particles x {len(scan)} beams per step, full update cycle"
              if n_sample < n_full else f"full workload ({n_full} particles x {len(scan)} beams)")
    line = {
        "impl": "reference",
        "metric": "particle_point_likelihood_evals_per_sec",
        "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "ms_per_step_full_workload_extrapolated": sec * 1e3 / scale,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, wl, len(scan)),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": 1, "kind": "port",
                         "sample": sample, "cpu": cpu_model(), "host_cores": os.cpu_count(),
                         "stage_ms": {k: v * 1e3 for k, v in stages.items()},
                         "note": "CPU restatement of the reference (oracle/); the reference is "
                                 "single-threaded (main.cpp:90) and cannot be built here",
                         "all_cores": all_cores},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, wl, n_used_beams):
    return {"workload": args.workload, "particles": wl["particles"],
            "map": f"{wl['map']}x{wl['map']}@0.05m", "scan_beams": wl["beams"],
            "beams_hit": n_used_beams, "num_beams": 0,
            "resample": RESAMPLE_NAMES[wl["pf"]["resample_type"]],
            "step": "handle_odometry+update_sensor+resample_and_cluster+get_estimate",
            "l2": "flushed between steps (256 MiB write); per-step CUDA-event pairs summed",
            "host_waits": ("per value (QMCL_EAGER_SYNC)" if os.environ.get("QMCL_EAGER_SYNC", "0") not in ("", "0")
                           else "deferred read-backs (qmcl_filter_set_sync_mode default)"),
            "sharding": (f"particles, contiguous global-index blocks; exchanges via {args.transport}"
                         if args.gpus > 1 else "none")}


# ------------------------------------------------------------------ GPU arm
def run_b200(args, wl):
    import torch
    import quickmcl_b200 as q

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=120))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")

    inputs = make_inputs(wl)
    occ, res, origin, pose, scan, (x, y, t) = inputs
    n = wl["particles"]
    stream = torch.cuda.current_stream()
    gm = q.create_likelihood_map(device=local_rank)
    gm.set_stream(stream.cuda_stream)
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    pfp = q.ParticleFilterParameters(**wl["pf"])
    pfp.resample_type = q.ResampleType(wl["pf"]["resample_type"])
    if world > 1:
        from quickmcl_b200.sharded import ShardedParticleFilter
        f = ShardedParticleFilter(gm, seed=3, group=dist.group.WORLD, transport=args.transport)
    else:
        f = q.create_particle_filter(gm, seed=3)
    f.set_parameters(pfp)
    init = q.make_particles(x, y, t, np.full(n, 1.0 / n))
    d_scan = torch.from_numpy(scan).cuda()
    h_scan = torch.from_numpy(scan).pin_memory().numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    evals = n * len(scan)

    def reset(it):
        f.set_particles(init)        # untimed: same initial cloud for every step
        f.set_rng(3, 10 + it)
        flush.zero_()                # evict the map from L2

    def step_device():
        f.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
        f.update_importance_from_device_cloud(d_scan.data_ptr(), len(scan))
        f.resample_and_cluster()
        return f.get_estimated_pose()

    def step_host():
        f.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
        f.update_importance_from_observations(h_scan)
        f.resample_and_cluster()
        return f.get_estimated_pose()

    def timed_loop(step_fn, steps, warmup, it0):
        for i in range(warmup):
            reset(it0 + i)
            step_fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        f.set_timing(1)
        f.get_timing(reset=True)
        if hasattr(f, "tail_profile"):
            f.tail_profile(reset=True)
        launches0 = q.kernel_launch_count()
        h2d0, d2h0 = q.transfer_bytes()
        h2d_reset = d2h_reset = 0
        total_ms = 0.0
        est = None
        for i in range(steps):
            a0, b0 = q.transfer_bytes()
            l0 = q.kernel_launch_count()
            reset(it0 + warmup + i)
            a1, b1 = q.transfer_bytes()
            h2d_reset += a1 - a0
            d2h_reset += b1 - b0
            launches0 += q.kernel_launch_count() - l0  # exclude the untimed restore
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            est = step_fn()
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        timing = f.get_timing(reset=True)
        f.set_timing(0)
        launches = q.kernel_launch_count() - launches0
        h2d1, d2h1 = q.transfer_bytes()
        if dist:
            tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = float(tt.item())
        try:
            prof = f.tail_profile()
        except AttributeError:   # the sharded wrapper has no fused tail
            prof = None
        return dict(total_ms=total_ms, launches=launches, timing=timing, estimate=est, tail_profile=prof,
                    h2d=(h2d1 - h2d0 - h2d_reset) / steps, d2h=(d2h1 - d2h0 - d2h_reset) / steps)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev = timed_loop(step_device, args.steps, args.warmup, 0)
    e2e = timed_loop(step_host, args.steps, max(3, args.warmup // 2), 1000)
    clocks = sampler.stop() if rank == 0 else None

    # optional per-stage breakdown (untimed extra pass, stderr only)
    if args.breakdown:   # every rank takes part (the steps contain collectives)
        f.set_timing(2)
        f.get_timing(reset=True)
        for i in range(args.steps):
            reset(2000 + i)
            step_device()
        tb = f.get_timing(reset=True)
        f.set_timing(0)
        if rank == 0:
            log("stage breakdown (device ms per step, CUDA events inside the library):")
            for k, (ms, cnt) in tb.items():
                if cnt:
                    log(f"  {k:15s} {ms / args.steps:9.4f} ms  ({cnt / args.steps:.1f} scopes/step)")
            try:
                log(f"  after the last step: {f.num_clusters()} clusters")
            except Exception as e:  # noqa: BLE001 (diagnostic only)
                log(f"  cluster count unavailable: {e}")

    global_n = n
    evals_job = global_n * len(scan)
    ms_per_step = dev["total_ms"] / args.steps
    value = evals_job / (ms_per_step * 1e-3)
    e2e_ms = e2e["total_ms"] / args.steps
    sensor_ms, sensor_cnt = dev["timing"]["sensor"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else \
        "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline = None
    if sensor_cnt:
        k_ms = sensor_ms / sensor_cnt
        local_evals = (n / world) * len(scan)
        achieved = local_evals * ALG_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9
        traffic = None
        traffic_note = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "sensor_traffic.json")))
            traffic = tj.get(args.workload)
            traffic_note = tj.get("_variant")
        except (OSError, ValueError):
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "kernel": "sensor_kernel (K7)",
                    "kernel_ms": k_ms, "kernel_evals_per_s": local_evals / (k_ms * 1e-3),
                    "alg_bytes_per_eval": ALG_BYTES_PER_EVAL, "peak_kind": peak_kind,
                    "launches_timed": sensor_cnt}
        if traffic is not None and traffic_note:
            roofline["traffic_note"] = traffic_note
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 or rank == 0:
        # bounded CPU sample: a slice of the particle set, full cycle, one thread
        n_cpu = int(min(n, max(2000, 2.0e9 / 6 // len(scan))))
        wl_s = dict(wl)
        pf = dict(wl["pf"])
        sc = n_cpu / n
        if sc < 1.0:
            pf["particle_count_min"] = max(1, int(pf["particle_count_min"] * sc))
            pf["particle_count_max"] = max(1, int(pf["particle_count_max"] * sc))
        wl_s["pf"] = pf
        if not args.no_cpu_baseline:
            sec, ev, stages = cpu_cycle_seconds(
                wl_s, (occ, res, origin, pose, scan, (x[:n_cpu], y[:n_cpu], t[:n_cpu])), 5, 1)
            cpu = {"value": ev / sec, "unit": "evals/s", "cores": 1, "kind": "port",
                   "sample": f"{n_cpu} of {n} particles x {len(scan)} beams, full update cycle, "
                             f"5 steps after 1 warm-up",
                   "cpu": cpu_model(), "host_cores": os.cpu_count(),
                   "stage_ms": {k: v * 1e3 for k, v in stages.items()},
                   "all_cores": cpu_all_cores(
                       wl_s, (occ, res, origin, pose, scan, (x[:n_cpu], y[:n_cpu], t[:n_cpu])))}

    line = {
        "metric": "particle_point_likelihood_evals_per_sec",
        "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, wl, len(scan)),
        "update_latency_ms": ms_per_step,
        "e2e": {"value": evals_job / (e2e_ms * 1e-3), "unit": "evals/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": e2e_ms},
        "gpu_launches": dev["launches"],
        "gpu_launches_per_step": dev["launches"] / args.steps,
        "stage_ms": {k: ms / cnt for k, (ms, cnt) in dev["timing"].items() if cnt},
        "tail_phases_us": ({name: ns / 1e3 for name, ns in dev["tail_profile"][0]}
                           if dev.get("tail_profile") else None),
        "tail_calls": dev["tail_profile"][1] if dev.get("tail_profile") else None,
        "tail_fallbacks": dev["tail_profile"][2] if dev.get("tail_profile") else None,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "estimate": [float(v) for v in dev["estimate"][1]] if dev["estimate"] else None,
    }
    if world > 1:
        # The N = 1 line of the contract is C2; the sharded runs are C4 (strong scaling of
        # one 16 M-particle filter).  The same workload on ONE GPU, measured earlier with
        # `bench.py --workload C4` and committed, is attached so that a speed-up can be read
        # off like with like.
        try:
            ref = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_c4_1gpu_final.json")))
            if ref.get("config", {}).get("workload") == args.workload:
                line["same_workload_on_one_gpu"] = {
                    "value": ref["value"], "unit": ref["unit"], "ms_per_step": ref["ms_per_step"],
                    "source": "profiles/r1_bench_c4_1gpu_final.json (python bench.py --workload C4)"}
        except (OSError, ValueError, KeyError):
            pass
    emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else (NCCL's version
    banner, library chatter) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--breakdown", action="store_true",
                    help="extra untimed pass with per-stage CUDA events, printed to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "torch"],
                    help="sharded exchanges: the library's built-in NCCL table or torch.distributed")
    args = ap.parse_args()
    if args.workload is None:
        # N=1: the configuration the metric is quoted on (configs[1]); N>1: the sharded config
        args.workload = "C2" if args.gpus == 1 else "C4"
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_b200(args, wl)


if __name__ == "__main__":
    sys.exit(main())
