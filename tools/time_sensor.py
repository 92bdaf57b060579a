"""Quick timing of the sensor update alone (K7 + K8) with device-resident inputs.
Development aid; bench.py is the contract benchmark."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import time

import numpy as np
import torch

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=100_000)
ap.add_argument("--map", type=int, default=2000)
ap.add_argument("--beams", type=int, default=1081)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cloud", default="tracking")
ap.add_argument("--path", type=int, default=0)
args = ap.parse_args()

occ, res, origin = syn.make_map(args.map)
gm = q.create_likelihood_map(device=0)
stream = torch.cuda.current_stream()
gm.set_stream(stream.cuda_stream)
gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
t0 = time.time()
gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
torch.cuda.synchronize()
print(f"map load {args.map}^2: {1e3 * (time.time() - t0):.2f} ms (first call, includes allocation)")
t0 = time.time()
gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
torch.cuda.synchronize()
print(f"map load {args.map}^2: {1e3 * (time.time() - t0):.2f} ms (second call)")
gm.set_sensor_path(args.path)
pose = syn.pick_truth_pose(occ, res, origin)
scan = syn.make_scan(occ, res, origin, pose, args.beams, 270.0)
n = args.particles
if args.cloud == "tracking":
    x, y, t = syn.gaussian_cloud(pose, n)
else:
    x, y, t = syn.uniform_cloud(occ, res, origin, n)
f = q.create_particle_filter(gm, seed=3)
f.set_particles(q.make_particles(x, y, t, np.full(n, 1.0 / n)))
d_scan = torch.from_numpy(scan).cuda()
evals = n * len(scan)
for _ in range(3):
    f.update_importance_from_device_cloud(d_scan.data_ptr(), len(scan))
times = []
for _ in range(args.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    f.update_importance_from_device_cloud(d_scan.data_ptr(), len(scan))
    e1.record(stream)
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
print(f"path={args.path} {args.cloud}: {n} particles x {len(scan)} beams = {evals:.3e} evals; "
      f"median {ms:.4f} ms -> {evals / ms / 1e9 * 1e3:.1f} G evals/s "
      f"({evals * 4.02 / ms / 1e6:.0f} GB/s algorithmic)")
