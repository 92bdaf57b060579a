"""C5 measurements (BASELINE.json configs[4]): likelihood-field rebuild on an
8192 x 8192 map, and a 256-filter batch at QuickMCL's default size.  Prints one
JSON object; the CPU oracle is timed on a bounded sample beside it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quickmcl_b200 as q
from quickmcl_b200 import batch, synthetic as syn
from oracle import binding as ob

out = {}
# ---- field rebuild
W = 8192
occ, res, origin = syn.make_map(W)
occ = torch.from_numpy(occ).pin_memory().numpy()   # pinned: the 67 MB host->device copy is a plain DMA
lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
gm = q.create_likelihood_map()
stream = torch.cuda.current_stream()
gm.set_stream(stream.cuda_stream)
gm.set_parameters(q.LikelihoodMapParameters(**lp))
grid = q.OccupancyGrid(occ, res, origin[0], origin[1])
gm.new_map(grid)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    gm.new_map(grid)          # includes the 67 MB host->device copy of the occupancy grid
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
cells = W * W
out["field_rebuild_8192"] = {"ms_end_to_end_incl_h2d": 1e3 * float(np.median(ts)),
                             "cells": cells, "algorithmic_bytes": 6 * cells,
                             "h2d_bytes": cells}
# oracle brushfire on a 1024^2 crop, scaled per cell
crop = np.ascontiguousarray(occ[:1024, :1024])
om = ob.OracleMap(**lp)
t0 = time.perf_counter()
om.new_map(crop, res, (0.0, 0.0), field_mode=ob.FIELD_BRUSHFIRE)
t_cpu = time.perf_counter() - t0
out["field_rebuild_8192"]["cpu_reference_brushfire_ms_extrapolated"] = 1e3 * t_cpu * cells / crop.size
out["field_rebuild_8192"]["cpu_sample"] = "1024x1024 crop, 1 thread, scaled by cell count"

# ---- 256-filter batch (C1 settings: 2000 particles, 400^2 map, 360-beam scan, low variance)
n_filters, n_particles = 256, 2000
occ1, res1, origin1 = syn.make_map(400)
lp1 = dict(syn.NODE_LIKELIHOOD, num_beams=0)
pf = q.ParticleFilterParameters(resample_type=q.ResampleType.low_variance,
                                particle_count_min=n_particles, particle_count_max=n_particles)
truth = syn.pick_truth_pose(occ1, res1, origin1)
scan = syn.make_scan(occ1, res1, origin1, truth, 360, 360.0)
cov = np.diag([0.25, 0.25, 0.01]).astype(np.float32)
# the fused batch (qmcl_batch): one map handle, four launches per cycle for all robots
fb = batch.FusedBatch(n_filters, q.OccupancyGrid(occ1, res1, origin1[0], origin1[1]),
                      q.LikelihoodMapParameters(**lp1), pf, seed=7, stream=stream.cuda_stream)
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
ts, td = [], []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    fb.cycle_packed(odo_new, odo_old, al, packed, offs)
    e1.record(stream)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
    td.append(e0.elapsed_time(e1))
ms = 1e3 * float(np.median(ts))
out["batch256_fused"] = {"ms_per_batch_update_wall": ms, "ms_per_batch_update_device": float(np.median(td)),
                         "p95_wall_ms": 1e3 * float(np.percentile(ts, 95)),
                         "filter_updates_per_s": n_filters / (ms * 1e-3),
                         "evals_per_s": n_filters * n_particles * len(scan) / (ms * 1e-3),
                         "cycles_fused_looped": fb.stats(), "launches_per_cycle": 3}
fb.close()
for lanes in ((16,) if os.environ.get("C5_QUICK") else (1, 8, 16, 32)):
    b = batch.FilterBatch(n_filters, q.OccupancyGrid(occ1, res1, origin1[0], origin1[1]),
                          q.LikelihoodMapParameters(**lp1), pf, lanes=lanes, seed=7)
    poses = np.tile(truth.astype(np.float32), (n_filters, 1))
    scans = [scan] * n_filters
    odo_new = np.tile([0.02, 0.01, 0.01], (n_filters, 1))
    odo_old = np.zeros((n_filters, 3))
    b.initialise(poses, cov)
    for _ in range(2):
        b.update(odo_new, odo_old, syn.NODE_ALPHA, scans)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        b.update(odo_new, odo_old, syn.NODE_ALPHA, scans)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(ts))
    tn = []
    for _ in range(5):
        t0 = time.perf_counter()
        b.update_native(odo_new, odo_old, syn.NODE_ALPHA, scans)
        torch.cuda.synchronize()
        tn.append(time.perf_counter() - t0)
    msn = 1e3 * float(np.median(tn))
    out[f"batch256_native_lanes{lanes}"] = {"ms_per_batch_update": msn,
                                            "filter_updates_per_s": n_filters / (msn * 1e-3),
                                            "evals_per_s": n_filters * n_particles * len(scan) / (msn * 1e-3)}
    out[f"batch256_lanes{lanes}"] = {"ms_per_batch_update": ms,
                                     "filter_updates_per_s": n_filters / (ms * 1e-3),
                                     "evals_per_s": n_filters * n_particles * len(scan) / (ms * 1e-3)}
    b.close()
# oracle: one filter, full cycle
om1 = ob.OracleMap(**lp1)
om1.new_map(occ1, res1, origin1, field_mode=ob.FIELD_BRUSHFIRE)
of = ob.OracleFilter(om1, seed=7)
of.set_parameters(resample_type=0, particle_count_min=n_particles, particle_count_max=n_particles)
of.initialise(truth.astype(np.float32), cov)
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    of.handle_odometry((0.02, 0.01, 0.01), (0, 0, 0), syn.NODE_ALPHA)
    of.update_importance_from_observations(scan)
    of.resample_and_cluster()
    of.get_estimated_pose()
    ts.append(time.perf_counter() - t0)
out["batch256_cpu_reference"] = {"ms_per_filter_update": 1e3 * float(np.median(ts)),
                                 "ms_per_batch_update_extrapolated": 256e3 * float(np.median(ts)),
                                 "cores": 1}
print(json.dumps(out, indent=1))
