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
cpu_baseline  the reference's own particle-filter sources (map_likelihood.cpp,
        odometry.cpp, particle_filter.cpp ...), compiled into oracle/_ref from
        where they lie against stand-in Eigen / ROS / PCL headers ("kind":
        "reference"), timed on this host on one thread - the reference is
        single-threaded (main.cpp:90).  Without that library: the CPU oracle,
        a restatement of the same code ("kind": "port")
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
def ref_cycle_seconds(wl, inputs, steps, warmup):
    """Times the full update cycle of the REFERENCE'S OWN CODE (oracle/_ref: map_likelihood.cpp,
    odometry.cpp, particle_filter.cpp ... compiled from /root/reference against stand-in
    headers).  Its private random engines read a script (oracle/ref_stubs/random) that is
    filled, untimed, with the draws of the shared Philox contract - popping a word is cheaper
    than mt19937 + libstdc++'s normal_distribution, so this errs in the reference's favour.
    Only the four calls of the cycle are timed.  Returns (seconds per step, evals per step,
    per-stage seconds, notes)."""
    from oracle import binding as ob
    from oracle import ref_binding as rb
    occ, res, origin, pose, scan, (x, y, t) = inputs
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    rm = rb.RefMap(occ, res, origin, **lp)          # the reference's brushfire + table build
    rf = rb.RefFilter(rm)
    pfp = dict(wl["pf"])
    rf.set_parameters(**pfp)
    rtype = pfp.get("resample_type", 2)
    n_slots = pfp.get("particle_count_max", 5000) if rtype == 2 else pfp.get("particle_count_min", 100)
    n = len(x)
    init = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    stages = {"odometry": 0.0, "sensor": 0.0, "resample": 0.0, "estimate": 0.0}
    overruns = 0
    for it in range(warmup + steps):
        rf.set_particles(init)
        rf.set_lowpass(0.0, 0.0)
        rb.script(normals=rb.motion_script(3, 10 + it, n))
        t0 = time.perf_counter()
        rf.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
        t1 = time.perf_counter()
        rf.update_importance_from_observations(scan)
        t2 = time.perf_counter()
        if rtype == 0:
            rb.script(words=rb.low_variance_script(3, 11 + it))
        else:
            w_slow, w_fast = rf.lowpass()
            w_diff = max(0.0, 1.0 - w_fast / w_slow) if w_slow else 0.0
            rb.script(words=rb.draw_script(3, 11 + it, n_slots, w_diff)[0])
        t3 = time.perf_counter()
        ok = rf.resample_and_cluster()
        t4 = time.perf_counter()
        if not ok:
            # the reference's low-variance walk ran off the end of the set and threw
            # (particle_filter.cpp:229-235; weights sum to < 1 after the wall penalties): the
            # resampling stage ends early, the estimate is timed on a valid set of the same size
            overruns += it >= warmup
            rf.set_particles(init)
            rf.cluster()
        t5 = time.perf_counter()
        rf.get_estimated_pose()
        t6 = time.perf_counter()
        if it >= warmup:
            stages["odometry"] += t1 - t0
            stages["sensor"] += t2 - t1
            stages["resample"] += t4 - t3
            stages["estimate"] += t6 - t5
    if rb.script_state()[2]:
        raise RuntimeError("the reference drew more numbers than the script held")
    stages = {k: v / steps for k, v in stages.items()}
    notes = {"low_variance_overrun_steps": overruns} if overruns else {}
    return sum(stages.values()), n * len(scan), stages, notes


def ref_all_cores(wl, inputs, steps=3):
    """The compiled reference on every host core at once: the particle set is cut into one slice per
    thread, each slice runs its own full cycle on its own ParticleFilter (shared read-only map, a
    script of its own).  NOT how the reference runs (it is single-threaded, main.cpp:90, and slices
    resample independently); it bounds what its code could do on this host if it were parallelised
    over particles.  Runs inside the ref-worker child."""
    from oracle import binding as ob
    from oracle import ref_binding as rb
    occ, res, origin, pose, scan, (x, y, t) = inputs
    threads_n = max(1, min(os.cpu_count() or 1, 64))
    n = len(x)
    per = n // threads_n
    if per < 64:
        return None
    rm = rb.RefMap(occ, res, origin, **dict(syn.NODE_LIKELIHOOD, num_beams=0))
    pfp = sampled_workload(dict(wl, particles=n), per)["pf"]
    rtype = pfp.get("resample_type", 2)
    n_slots = pfp.get("particle_count_max", 5000) if rtype == 2 else pfp.get("particle_count_min", 100)
    jobs = []
    for k in range(threads_n):
        rf = rb.RefFilter(rm)
        rf.set_parameters(**pfp)
        sl = slice(k * per, (k + 1) * per)
        init = ob.make_particles(x[sl], y[sl], t[sl], np.full(per, 1.0 / per))
        # every draw of every step, prepared before the clock starts
        scripts = [(rb.motion_script(3 + k, 10 + it, per),
                    rb.low_variance_script(3 + k, 11 + it) if rtype == 0
                    else rb.draw_script(3 + k, 11 + it, n_slots, 0.0)[0]) for it in range(steps + 1)]
        jobs.append((rf, init, scripts))
    barrier = threading.Barrier(threads_n + 1)
    errors = []

    def work(rf, init, scripts):
        try:
            for it, (normals, words) in enumerate(scripts):
                rf.set_particles(init)
                rf.set_lowpass(0.0, 0.0)
                if it == 1:
                    barrier.wait()   # the warm-up step is over everywhere: the clock starts
                rb.script(words=words, normals=normals)
                rf.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
                rf.update_importance_from_observations(scan)
                if not rf.resample_and_cluster():
                    rf.set_particles(init)
                    rf.cluster()
                rf.get_estimated_pose()
        except BaseException as e:
            errors.append(e)
            barrier.abort()

    old = threading.stack_size(256 << 20)
    try:
        ts = [threading.Thread(target=work, args=job) for job in jobs]
        for th in ts:
            th.start()
    finally:
        threading.stack_size(old)
    barrier.wait()
    t0 = time.perf_counter()
    for th in ts:
        th.join()
    sec = (time.perf_counter() - t0) / steps
    if errors:
        raise errors[0]
    return {"value": threads_n * per * len(scan) / sec, "unit": "evals/s", "cores": threads_n, "kind": "reference",
            "sample": f"{threads_n} threads x {per} particles x {len(scan)} beams, independent full cycles per "
                      f"slice, {steps} steps after 1 warm-up",
            "note": "the reference's own compiled code, one ParticleFilter per thread; not reference behaviour "
                    "(the reference is single-threaded); an upper bound for a particle-parallel host"}


def with_deep_stack(fn, *args):
    """Runs fn on a thread with a 1 GiB stack: the reference labels clusters by recursion, one
    frame per bin (space_partitioning.cpp:101-128), and overflows the default 8 MiB stack on the
    large connected bin sets of a global cloud of >= 1e5 particles (the node was never run there).
    The stack is virtual memory; only the depth actually reached is touched."""
    box = {}

    def run():
        try:
            box["result"] = fn(*args)
        except BaseException as e:  # handed to the caller below
            box["error"] = e

    old = threading.stack_size(1 << 30)
    try:
        th = threading.Thread(target=run)
        th.start()
        th.join()
    finally:
        threading.stack_size(old)
    if "error" in box:
        raise box["error"]
    return box["result"]


MAX_REFERENCE_CLUSTERS = 65000  # space_partitioning.h:59: ClusterId is uint16_t, 65535 = invalid


def sampled_workload(wl, n_sample):
    """The workload cut down to its first n_sample particles, particle-count bounds scaled alike."""
    wl_s = dict(wl)
    pf = dict(wl["pf"])
    scale = n_sample / wl["particles"]
    if scale < 1.0:
        pf["particle_count_min"] = max(1, int(pf["particle_count_min"] * scale))
        pf["particle_count_max"] = max(1, int(pf["particle_count_max"] * scale))
    wl_s["pf"] = pf
    return wl_s


def sample_inputs(inputs, n_sample):
    occ, res, origin, pose, scan, (x, y, t) = inputs
    return occ, res, origin, pose, scan, (x[:n_sample], y[:n_sample], t[:n_sample])


def reference_can_label(wl_s, inputs):
    """The reference numbers clusters with uint16_t and treats 65535 as "none yet"
    (space_partitioning.h:59-63): on a cloud with that many clusters its recursion never ends.
    The oracle (no such limit) counts the clusters of the set before and after one cycle."""
    from oracle import binding as ob
    occ, res, origin, pose, scan, (x, y, t) = inputs
    om = ob.OracleMap(**dict(syn.NODE_LIKELIHOOD, num_beams=0))
    om.new_map(occ, res, origin, field_mode=ob.FIELD_BRUSHFIRE)
    of = ob.OracleFilter(om, seed=3)
    of.set_parameters(**wl_s["pf"])
    n = len(x)
    of.set_particles(ob.make_particles(x, y, t, np.full(n, 1.0 / n)))
    of.cluster()
    counts = [of.num_clusters()]
    of.set_rng(3, 10)
    of.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
    of.update_importance_from_observations(scan)
    of.resample_and_cluster()
    counts.append(of.num_clusters())
    return max(counts) < MAX_REFERENCE_CLUSTERS, max(counts)


def run_ref_worker(args, wl):
    """Child process of cpu_cycle(): times the compiled reference on a sample and prints one JSON
    object.  A process of its own because the reference's code can take the process down (a stack
    overflow cannot be caught) and must not take the bench line with it."""
    n_sample = args.sample or wl["particles"]
    wl_s = sampled_workload(wl, n_sample)
    inputs = sample_inputs(make_inputs(wl), n_sample)
    ok, clusters = reference_can_label(wl_s, inputs)
    if not ok:
        emit({"unusable": f"{clusters} clusters in this cloud: the reference's 16-bit cluster ids "
                          f"(space_partitioning.h:59) cannot label them, its recursion does not end"})
        return 0
    sec, ev, stages, notes = with_deep_stack(ref_cycle_seconds, wl_s, inputs, args.steps, args.warmup)
    all_cores = None
    try:
        all_cores = ref_all_cores(wl_s, inputs)
    except Exception as e:
        log(f"all-cores run of the compiled reference skipped ({type(e).__name__}: {e})")
    emit({"sec": sec, "evals": ev, "stages": stages, "notes": notes, "clusters": clusters, "all_cores": all_cores})
    return 0


def cpu_cycle(wl, wl_name, inputs, n_sample, steps, warmup, allow_ref=True):
    """The CPU baseline on the first n_sample particles of the workload: the compiled reference
    (oracle/_ref, timed in a child process) when its library is there and the cloud is one it can
    label, else the oracle port.  Returns (seconds per step, evals per step, per-stage seconds,
    kind, note, extra keys for the cpu_baseline object)."""
    wl_s = sampled_workload(wl, n_sample)
    inputs_s = sample_inputs(inputs, n_sample)
    why_port = "oracle/_ref is not built"
    if os.environ.get("QMCL_BENCH_CPU_PORT"):
        why_port = "QMCL_BENCH_CPU_PORT is set"
    elif not allow_ref:
        why_port = "N > 1: the compiled reference is timed by the N = 1 line and by --impl reference"
    else:
        try:
            from oracle import ref_binding as rb
            if rb.available():
                out = subprocess.run(
                    [sys.executable, os.path.abspath(__file__), "--impl", "ref-worker", "--workload", wl_name,
                     "--sample", str(n_sample), "--steps", str(steps), "--warmup", str(warmup)],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=900)
                res = json.loads(out.stdout.decode().strip().splitlines()[-1]) if out.returncode == 0 else None
                if res is None:
                    why_port = f"the compiled reference ended with exit code {out.returncode} on this sample"
                elif "unusable" in res:
                    why_port = res["unusable"]
                else:
                    note = ("the reference's own sources (map_likelihood.cpp, odometry.cpp, "
                            "particle_filter.cpp, space_partitioning.cpp, statistics.cpp ...) compiled into "
                            "oracle/_ref against stand-in Eigen / ROS / PCL headers, -O3 -DNDEBUG like catkin "
                            "Release; single-threaded like the node (main.cpp:90); random draws read from a "
                            "script, which is cheaper than its mt19937; run on a 1 GiB stack because its "
                            "cluster labelling recurses once per bin")
                    if res["notes"]:
                        note += f"; {res['notes']}"
                    try:  # the oracle port on the same sample, for comparison with earlier rounds
                        psec, pev, _ = cpu_cycle_seconds(wl_s, inputs_s, min(steps, 2), 1)
                        note += (f"; the oracle port (kind 'port' of earlier rounds) on the same sample: "
                                 f"{pev / psec:.4g} evals/s")
                    except Exception as e:
                        log(f"oracle port timing skipped ({type(e).__name__}: {e})")
                    extra = {"all_cores_reference": res["all_cores"]} if res.get("all_cores") else {}
                    return res["sec"], res["evals"], res["stages"], "reference", note, extra
        except Exception as e:  # the baseline must never take the bench line with it
            why_port = f"{type(e).__name__}: {e}"
    log(f"CPU baseline: timing the oracle port ({why_port})")
    sec, ev, stages = cpu_cycle_seconds(wl_s, inputs_s, steps, warmup)
    return sec, ev, stages, "port", ("CPU restatement of the reference (oracle/); the reference is "
                                     f"single-threaded (main.cpp:90); not the compiled reference: {why_port}"), {}


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
    if wl["cloud"] == "global":
        # a global cloud has about one cluster per particle until it gets dense, and the reference
        # numbers clusters with 16 bits (space_partitioning.h:59): keep the sample one it can label
        n_sample = min(n_sample, 50_000)
    occ, res, origin, pose, scan, (x, y, t) = inputs
    scale = n_sample / n_full
    wl_s = sampled_workload(wl, n_sample)
    sec, evals, stages, kind, note, extra = cpu_cycle(wl, args.workload, inputs, n_sample, args.steps, args.warmup)
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
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": 1, "kind": kind,
                         "sample": sample, "cpu": cpu_model(), "host_cores": os.cpu_count(),
                         "field": "brushfire (the reference's own field build)",
                         "stage_ms": {k: v * 1e3 for k, v in stages.items()},
                         "note": note,
                         "all_cores": all_cores, **extra},
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
            "l2": ("NOT FLUSHED (QMCL_BENCH_NO_FLUSH diagnostic run: not a bench value)"
                   if os.environ.get("QMCL_BENCH_NO_FLUSH") else
                   "flushed between steps (256 MiB write); one CUDA-event pair per step, summed "
                   "(median and p95 alongside)"),
            "host_waits": ("per value (QMCL_EAGER_SYNC)" if os.environ.get("QMCL_EAGER_SYNC", "0") not in ("", "0")
                           else ("one: the estimate (fused front + sensor + tail launches)" if args.gpus == 1
                                 else "deferred read-backs (qmcl_filter_set_sync_mode default)")),
            "sharding": (f"particles, contiguous global-index blocks; exchanges via {args.transport}"
                         if args.gpus > 1 else "none")}


# ------------------------------------------------------------------ GPU arm
class Harness:
    """One workload on one filter (single GPU or one shard of a sharded filter): builds the
    map and the filter, and times update steps with CUDA events on the filter's stream."""

    def __init__(self, q, torch, wl_name, device, stream, dist=None, transport="nccl", sharded=False):
        self.q, self.torch, self.name, self.dist = q, torch, wl_name, dist
        self.wl = wl = WORKLOADS[wl_name]
        self.stream = stream
        self.inputs = make_inputs(wl)
        occ, res, origin, pose, scan, (x, y, t) = self.inputs
        self.scan = scan
        self.n = wl["particles"]
        self.gm = q.create_likelihood_map(device=device)
        self.gm.set_stream(stream.cuda_stream)
        self.gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
        self.gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
        pfp = q.ParticleFilterParameters(**wl["pf"])
        pfp.resample_type = q.ResampleType(wl["pf"]["resample_type"])
        if sharded:
            from quickmcl_b200.sharded import ShardedParticleFilter
            self.f = ShardedParticleFilter(self.gm, seed=3, group=dist.group.WORLD, transport=transport)
            self.f.set_rebalance_threshold(0.05)   # equal blocks are restored only beyond 5 % drift
        else:
            self.f = q.create_particle_filter(self.gm, seed=3)
        self.sharded = sharded
        self.f.set_parameters(pfp)
        self.init = q.make_particles(x, y, t, np.full(self.n, 1.0 / self.n))
        self.d_scan = torch.from_numpy(scan).cuda()
        self.h_scan = torch.from_numpy(scan).pin_memory().numpy()
        self.evals = self.n * len(scan)

    def close(self):
        self.f.close()
        self.gm.close()

    def reset(self, it, flush):
        self.f.set_particles(self.init)   # untimed: same initial cloud for every step
        self.f.set_rng(3, 10 + it)
        if flush is not None:
            flush.zero_()                 # evict the map from L2

    def step(self, host):
        f = self.f
        self.cycles_run = getattr(self, "cycles_run", 0) + 1
        f.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
        if host:
            f.update_importance_from_observations(self.h_scan)
        else:
            f.update_importance_from_device_cloud(self.d_scan.data_ptr(), len(self.scan))
        f.resample_and_cluster()
        return f.get_estimated_pose()

    def timed_loop(self, host, steps, warmup, it0, flush):
        q, torch, f, dist = self.q, self.torch, self.f, self.dist if self.sharded else None
        for i in range(warmup):
            self.reset(it0 + i, flush)
            self.step(host)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        f.set_timing(1)
        f.get_timing(reset=True)
        if hasattr(f, "tail_profile") and not self.sharded:
            f.tail_profile(reset=True)
        launches = 0
        h2d = d2h = 0
        per_step = []
        est = None
        for i in range(steps):
            self.reset(it0 + warmup + i, flush)
            a0, b0 = q.transfer_bytes()
            l0 = q.kernel_launch_count()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            est = self.step(host)
            e1.record(self.stream)
            e1.synchronize()
            per_step.append(e0.elapsed_time(e1))
            launches += q.kernel_launch_count() - l0
            a1, b1 = q.transfer_bytes()
            h2d += a1 - a0
            d2h += b1 - b0
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        timing = f.get_timing(reset=True)
        f.set_timing(0)
        prof = f.tail_profile() if hasattr(f, "tail_profile") and not self.sharded else None
        total_ms = float(sum(per_step))
        if dist:
            tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = float(tt.item())
        ps = np.asarray(per_step)
        return dict(total_ms=total_ms, steps=steps, ms_per_step=total_ms / steps, median_ms=float(np.median(ps)),
                    p95_ms=float(np.percentile(ps, 95)), min_ms=float(ps.min()),
                    launches=launches, timing=timing, estimate=est, tail_profile=prof,
                    h2d=h2d / steps, d2h=d2h / steps)

    def summary(self, dev, e2e, world=1):
        """The per-workload block of the JSON line."""
        out = {"workload": self.name, "particles": self.n, "beams_hit": len(self.scan),
               "resample": RESAMPLE_NAMES[self.wl["pf"]["resample_type"]],
               "ms_per_step": dev["ms_per_step"], "median_ms": dev["median_ms"], "p95_ms": dev["p95_ms"],
               "evals_per_s": self.evals / (dev["ms_per_step"] * 1e-3),
               "gpu_launches_per_step": dev["launches"] / dev["steps"],
               "stage_ms": {k: ms / cnt for k, (ms, cnt) in dev["timing"].items() if cnt}}
        if e2e:
            out["e2e_ms_per_step"] = e2e["ms_per_step"]
        sm, sc = dev["timing"]["sensor"]
        if sc:
            k_ms = sm / sc
            local_evals = (self.n / world) * len(self.scan)
            out["k7_ms"] = k_ms
            out["k7_evals_per_s"] = local_evals / (k_ms * 1e-3)
            out["k7_roofline_frac"] = local_evals * ALG_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9 / peak_gbs()[0]
        if dev.get("tail_profile"):
            out["tail_phases_us"] = {name: ns / 1e3 for name, ns in dev["tail_profile"][0]}
            out["tail_fallbacks"] = dev["tail_profile"][2]
        return out


def peak_gbs():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def particle_checksum(particles, first):
    """Order-sensitive 64-bit checksum of a block of the particle set: slot (first + i) weighs
    the bit patterns of pose i.  Summed over blocks it identifies the whole set slot by slot."""
    u = np.ascontiguousarray(np.stack([particles["x"], particles["y"], particles["theta"]], 1)).view(np.uint32)
    h = (u[:, 0].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
         ^ u[:, 1].astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         ^ u[:, 2].astype(np.uint64) * np.uint64(0x165667B19E3779F9))
    idx = np.arange(first + 1, first + 1 + len(h), dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(h * idx, dtype=np.uint64))


def measure_field_rebuild(q, torch, stream, width=8192, reps=5):
    """configs[4]: the likelihood-field rebuild (exact EDT + Gaussian table + free-cell list) of a
    width x width map, end to end through the plugin call with a pinned host occupancy grid."""
    occ, res, origin = syn.make_map(width)
    occ = torch.from_numpy(occ).pin_memory().numpy()
    gm = q.create_likelihood_map()
    gm.set_stream(stream.cuda_stream)
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    grid = q.OccupancyGrid(occ, res, origin[0], origin[1])
    gm.new_map(grid)                      # first call allocates
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        gm.new_map(grid)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    gm.close()
    cells = width * width
    ms = 1e3 * float(np.median(ts))
    return {"workload": f"C5 field rebuild {width}x{width}", "ms_end_to_end_incl_h2d": ms,
            "h2d_bytes": cells, "algorithmic_bytes": 6 * cells,
            "roofline_frac_end_to_end": 6 * cells / (ms * 1e-3) / 1e9 / peak_gbs()[0],
            "reps": reps}


def measure_batch(q, torch, stream, n_filters=256, n_particles=2000, cycles=20):
    """configs[4]: the multi-robot batch — n_filters independent filters at QuickMCL's default
    size (C1 settings) on one map, one fused batch cycle (qmcl_batch_cycle) per step."""
    from quickmcl_b200 import batch
    occ, res, origin = syn.make_map(400)
    pf = q.ParticleFilterParameters(resample_type=q.ResampleType.low_variance,
                                    particle_count_min=n_particles, particle_count_max=n_particles)
    truth = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, truth, 360, 360.0)
    cov = np.diag([0.25, 0.25, 0.01]).astype(np.float32)
    fb = batch.FusedBatch(n_filters, q.OccupancyGrid(occ, res, origin[0], origin[1]),
                          q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)), pf,
                          seed=7, stream=stream.cuda_stream)
    for f in fb.filters:
        f.initialise(truth.astype(np.float32), cov)
    odo_new = np.tile([0.02, 0.01, 0.01], (n_filters, 1)).astype(np.float64)
    odo_old = np.zeros((n_filters, 3))
    al = np.asarray(syn.NODE_ALPHA, np.float32)
    offs = np.arange(n_filters + 1, dtype=np.uint64) * len(scan)
    packed = np.ascontiguousarray(np.tile(scan, (n_filters, 1)), np.float32)
    for _ in range(3):
        fb.cycle_packed(odo_new, odo_old, al, packed, offs)
    torch.cuda.synchronize()
    wall, devt = [], []
    l0 = q.kernel_launch_count()
    for _ in range(cycles):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        fb.cycle_packed(odo_new, odo_old, al, packed, offs)   # host scans in, host estimates out
        e1.record(stream)
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        devt.append(e0.elapsed_time(e1))
    launches = (q.kernel_launch_count() - l0) / cycles
    stats = fb.stats()
    fb.close()
    ms = 1e3 * float(np.median(wall))
    return {"workload": f"C5 batch {n_filters} filters x {n_particles} particles x {len(scan)} beams, low variance",
            "ms_per_batch_update_e2e": ms, "p95_ms": 1e3 * float(np.percentile(wall, 95)),
            "ms_per_batch_update_device": float(np.median(devt)),
            "filter_updates_per_s": n_filters / (ms * 1e-3),
            "evals_per_s": n_filters * n_particles * len(scan) / (ms * 1e-3),
            "gpu_launches_per_batch_update": launches, "cycles_fused_looped": list(stats)}


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
                                timeout=datetime.timedelta(seconds=180))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if os.environ.get("QMCL_BENCH_NO_FLUSH"):   # diagnostic only: such a line is not a bench value
        flush = None
    PARITY_IT = 5000

    # N > 1: the same workload on ONE GPU first (rank 0, in this process), so that the speed-up
    # and the parity of the NCCL path are read off like with like.
    one_gpu = None
    if world > 1 and rank == 0 and not args.no_one_gpu_anchor:
        h1 = Harness(q, torch, args.workload, local_rank, stream)
        d1 = h1.timed_loop(False, min(args.steps, 5), 2, 0, flush)
        h1.reset(PARITY_IT, flush)
        est1 = h1.step(False)
        p1 = h1.f.get_particles()
        one_gpu = {"ms_per_step": d1["ms_per_step"], "median_ms": d1["median_ms"],
                   "value": h1.evals / (d1["ms_per_step"] * 1e-3), "unit": "evals/s",
                   "steps": min(args.steps, 5), "stage_ms": {k: ms / c for k, (ms, c) in d1["timing"].items() if c},
                   "source": "measured in this run on rank 0 before the sharded run",
                   "_estimate": est1, "_checksum": particle_checksum(p1, 0), "_n": len(p1)}
        del p1
        h1.close()
        del h1
        torch.cuda.empty_cache()
    if dist:
        dist.barrier()

    h = Harness(q, torch, args.workload, local_rank, stream, dist=dist, transport=args.transport,
                sharded=world > 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev = h.timed_loop(False, args.steps, args.warmup, 0, flush)
    e2e = h.timed_loop(True, args.steps, max(3, args.warmup // 2), 1000, flush)

    # hardware parity of the sharded (NCCL) path against the one-GPU run: estimate and the
    # resampled set slot by slot (checksum of pose bits weighted by global slot index)
    parity = None
    if world > 1:
        h.reset(PARITY_IT, flush)
        est = h.step(False)
        first, count, total = h.f.shard_range()
        loc = h.f.get_particles()
        cs = torch.tensor([np.int64(np.uint64(particle_checksum(loc, first)))], dtype=torch.int64, device="cuda")
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        cs_all = int(np.uint64(np.int64(cs.item())))
        if rank == 0 and one_gpu is not None:
            e1 = one_gpu.pop("_estimate")
            d_pose = float(np.max(np.abs(np.asarray(est[1]) - np.asarray(e1[1]))))
            d_cov = float(np.max(np.abs(np.asarray(est[2]) - np.asarray(e1[2]))))
            parity = {"resampled_set_checksum_equal": cs_all == one_gpu.pop("_checksum"),
                      "particles": [int(total), one_gpu.pop("_n")],
                      "estimate_max_abs_diff": max(d_pose, d_cov), "estimate_valid": [bool(est[0]), bool(e1[0])]}
            parity["ok"] = bool(parity["resampled_set_checksum_equal"] and parity["estimate_max_abs_diff"] <= 1e-9
                                and parity["particles"][0] == parity["particles"][1])

    secondary = None
    if world > 1 and not args.no_secondary and args.workload == "C4":
        # the KLD variant of the sharded workload ("C4 low_variance and kld at 16 M"):
        # candidate draws enumerated on every rank, slot all-to-all, first-draw table min-allreduce
        hk = Harness(q, torch, "C4K", local_rank, stream, dist=dist, transport=args.transport, sharded=True)
        dk = hk.timed_loop(False, min(args.steps, 20), 3, 0, flush)
        secondary = {"C4K": hk.summary(dk, None, world)}
        secondary["C4K"]["kept_particles"] = int(hk.f.num_particles())
        hk.close()
        del hk
        torch.cuda.empty_cache()
    if world == 1 and not args.no_secondary and args.workload == "C2":
        secondary = {}
        for name, st, wu in (("C1", 50, 5), ("C3", 20, 3)):
            hs = Harness(q, torch, name, local_rank, stream)
            ds = hs.timed_loop(False, st, wu, 0, flush)
            es = hs.timed_loop(True, max(5, st // 2), 2, 1000, flush)
            secondary[name] = hs.summary(ds, es)
            hs.close()
        # the sharded workload on this one GPU: the like-for-like anchor of the N = 2/4/8 lines
        # (those runs measure it again on rank 0, see `same_workload_on_one_gpu`)
        hs = Harness(q, torch, "C4", local_rank, stream)
        secondary["C4_one_gpu"] = hs.summary(hs.timed_loop(False, 5, 2, 0, flush), None)
        hs.close()
        del hs
        torch.cuda.empty_cache()
        secondary["C5_field_rebuild"] = measure_field_rebuild(q, torch, stream)
        secondary["C5_batch"] = measure_batch(q, torch, stream)
    clocks = sampler.stop() if rank == 0 else None

    # optional per-stage breakdown (untimed extra pass, stderr only): level 2 runs the stages as
    # separate launches, i.e. the multi-launch path
    if args.breakdown:   # every rank takes part (the steps contain collectives)
        f = h.f
        f.set_timing(2)
        f.get_timing(reset=True)
        for i in range(min(args.steps, 20)):
            h.reset(2000 + i, flush)
            h.step(False)
        tb = f.get_timing(reset=True)
        f.set_timing(0)
        # every rank reports: what one rank books as a long exchange is another rank's longer stage
        rows = [f"stage breakdown of rank {rank} (device ms per step, CUDA events inside the library, "
                "multi-launch path):"]
        for k, (ms, cnt) in tb.items():
            if cnt:
                rows.append(f"  {k:15s} {ms / min(args.steps, 20):9.4f} ms  ({cnt / min(args.steps, 20):.1f} scopes/step)")
        if world > 1 and dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, rows)
            if rank == 0:
                for g in gathered:
                    log("\n".join(g))
        elif rank == 0:
            log("\n".join(rows))

    n, scan = h.n, h.scan
    evals_job = n * len(scan)
    ms_per_step = dev["ms_per_step"]
    value = evals_job / (ms_per_step * 1e-3)
    e2e_ms = e2e["ms_per_step"]
    sensor_ms, sensor_cnt = dev["timing"]["sensor"]
    peak, peak_kind = peak_gbs()
    roofline = None
    if sensor_cnt:
        k_ms = sensor_ms / sensor_cnt
        local_evals = (n / world) * len(scan)
        achieved = local_evals * ALG_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9
        traffic = traffic_note = None
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
        try:
            gp = json.load(open(os.path.join(ROOT, "profiles", "r2_probe_gather.json")))
            best = max(r["gathers_per_s"] for r in gp["results"]
                       if r["pattern"] == "random" and r["elem_bytes"] == 1 and r["table_bytes"] == 4 << 20)
            roofline["vs_measured_l2_random_gather"] = {
                "peak_gathers_per_s": best, "kernel_evals_over_peak": local_evals / (k_ms * 1e-3) / best,
                "source": "profiles/r2_probe_gather.json: independent random 1-byte gathers into a 4 MB "
                          "L2-resident table; K7 exceeds it because consecutive beams share cache lines"}
        except (OSError, ValueError, KeyError):
            pass
    if rank != 0:
        h.close()
        if dist:
            dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline:
        # bounded CPU sample: a slice of the particle set, full cycle, one thread
        occ, res, origin, pose, _, (x, y, t) = h.inputs
        n_cpu = int(min(n, max(2000, 2.0e9 / 6 // len(scan))))
        wl_s = sampled_workload(wl, n_cpu)
        sec, ev, stages, kind, note, extra = cpu_cycle(
            wl, args.workload, (occ, res, origin, pose, scan, (x, y, t)), n_cpu, 5, 1, allow_ref=(world == 1))
        cpu = {"value": ev / sec, "unit": "evals/s", "cores": 1, "kind": kind, "note": note, **extra,
               "sample": f"{n_cpu} of {n} particles x {len(scan)} beams, full update cycle, "
                         f"5 steps after 1 warm-up",
               "field": "brushfire (the reference's own field build; the GPU arm uses the exact EDT "
                        "north_star asks for — same cost per evaluation)",
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
        "update_latency_median_ms": dev["median_ms"], "update_latency_p95_ms": dev["p95_ms"],
        "e2e": {"value": evals_job / (e2e_ms * 1e-3), "unit": "evals/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": e2e_ms, "median_ms": e2e["median_ms"], "p95_ms": e2e["p95_ms"]},
        "gpu_launches": dev["launches"],
        "gpu_launches_per_step": dev["launches"] / args.steps,
        "stage_ms": {k: ms / cnt for k, (ms, cnt) in dev["timing"].items() if cnt},
        "tail_phases_us": ({name: ns / 1e3 for name, ns in dev["tail_profile"][0]}
                           if dev.get("tail_profile") else None),
        "tail_fallbacks": dev["tail_profile"][2] if dev.get("tail_profile") else None,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "estimate": [float(v) for v in dev["estimate"][1]] if dev["estimate"] else None,
    }
    if secondary is not None:
        line["secondary"] = secondary
    if world > 1:
        if one_gpu is not None:
            one_gpu = {k: v for k, v in one_gpu.items() if not k.startswith("_")}
            one_gpu["speedup_of_this_run"] = one_gpu["ms_per_step"] / ms_per_step
            line["same_workload_on_one_gpu"] = one_gpu
        line["parity_vs_one_gpu"] = parity
        if hasattr(h.f, "shard_stats"):
            # counted over every cycle this filter ran (warm-up, timed, e2e and parity passes)
            st = h.f.shard_stats()
            cycles = max(1, getattr(h, "cycles_run", 0))
            line["exchanges"] = dict(st, cycles=cycles,
                                     collectives_per_cycle=st["collectives"] / cycles,
                                     chain_rounds_per_cycle=st["chain_rounds"] / cycles)
    emit(line)
    h.close()
    if dist:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        log("bench.py: the sharded run does not reproduce the one-GPU run:", parity)
        return 1
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "ref-worker"])
    ap.add_argument("--sample", type=int, default=0, help="ref-worker only: particles of the sample")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--breakdown", action="store_true",
                    help="extra untimed pass with per-stage CUDA events, printed to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="N=1: skip the C1 / C3 / C5 measurements that ride along in `secondary`")
    ap.add_argument("--no-one-gpu-anchor", action="store_true",
                    help="N>1: skip the one-GPU run of the same workload on rank 0 (and the parity check)")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "torch"],
                    help="sharded exchanges: the library's built-in NCCL table or torch.distributed")
    args = ap.parse_args()
    if args.workload is None:
        # N=1: the configuration the metric is quoted on (configs[1]); N>1: the sharded config
        args.workload = "C2" if args.gpus == 1 else "C4"
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.impl == "ref-worker":
        return run_ref_worker(args, wl)
    return run_b200(args, wl)


if __name__ == "__main__":
    sys.exit(main())
