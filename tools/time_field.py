"""Likelihood-field rebuild timing (K1-K4): host-pointer call (includes the H2D
copy of the occupancy grid) and device time between CUDA events."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
ap = argparse.ArgumentParser(); ap.add_argument("--map", type=int, default=8192); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
occ, res, origin = syn.make_map(a.map)
occ = torch.from_numpy(occ).pin_memory().numpy()   # pinned host buffer: the copy is a plain DMA
gm = q.create_likelihood_map(); s = torch.cuda.current_stream(); gm.set_stream(s.cuda_stream)
gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
grid = q.OccupancyGrid(occ, res, origin[0], origin[1])
gm.new_map(grid); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    t0 = time.perf_counter(); gm.new_map(grid); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
cells = a.map * a.map
ms = 1e3 * float(np.median(ts))
print(f"field rebuild {a.map}^2: {ms:.3f} ms end to end (pinned H2D of {cells/1e6:.0f} MB included); "
      f"6 B/cell algorithmic = {6*cells/1e6:.0f} MB -> {6*cells/ms/1e6:.0f} GB/s end to end")
