"""Seeded synthetic inputs for the tests and the benchmark.

Host-side input generation only (numpy): occupancy maps with rooms, ray-cast
scans and particle clouds.  Nothing here is on the measured path.
"""
import math

import numpy as np

RES = 0.05  # metres per cell


def make_map(width, seed=1, res=RES):
    """Square `width` x `width` int8 occupancy grid.

    Outer 10-cell strip unknown (-1), a wall ring (100), axis-aligned interior
    walls with door gaps, free interior (0).  Returns (occ, res, (ox, oy)) with
    the origin chosen so that the map is centred on the world origin.
    """
    rng = np.random.default_rng(seed)
    W = int(width)
    occ = np.zeros((W, W), np.int8)
    b = 10
    th = 2  # wall thickness in cells
    lo, hi = b, W - b
    # interior walls: a jittered lattice of rooms 6-10 m across
    def wall_positions():
        pos, p = [], lo + int(rng.integers(100, 160))
        while p < hi - 100:
            pos.append(p)
            p += int(rng.integers(120, 200))
        return pos
    xs, ys = wall_positions(), wall_positions()
    door = 20
    for x in xs:
        occ[lo:hi, x:x + th] = 100
        edges = [lo] + ys + [hi]
        for a, c in zip(edges[:-1], edges[1:]):
            if c - a > 3 * door:
                d = int(rng.integers(a + door, c - 2 * door))
                occ[d:d + door, x:x + th] = 0
    for y in ys:
        occ[y:y + th, lo:hi] = 100
        edges = [lo] + xs + [hi]
        for a, c in zip(edges[:-1], edges[1:]):
            if c - a > 3 * door:
                d = int(rng.integers(a + door, c - 2 * door))
                occ[y:y + th, d:d + door] = 0
    # a few grey cells (35..65: unknown state, not an obstacle) and clutter
    n_clutter = max(4, W // 50)
    cx = rng.integers(lo + 5, hi - 8, n_clutter)
    cy = rng.integers(lo + 5, hi - 8, n_clutter)
    for x, y in zip(cx, cy):
        occ[y:y + 3, x:x + 3] = 100
    gx = rng.integers(lo + 5, hi - 5, n_clutter)
    gy = rng.integers(lo + 5, hi - 5, n_clutter)
    occ[gy, gx] = 50
    # wall ring and unknown border
    occ[lo:lo + th, lo:hi] = 100
    occ[hi - th:hi, lo:hi] = 100
    occ[lo:hi, lo:lo + th] = 100
    occ[lo:hi, hi - th:hi] = 100
    occ[:b, :] = -1
    occ[-b:, :] = -1
    occ[:, :b] = -1
    occ[:, -b:] = -1
    origin = (-W * res / 2.0, -W * res / 2.0)
    return occ, res, origin


def pick_truth_pose(occ, res, origin, seed=2, clearance=20):
    """A free pose near the map centre with `clearance` free cells around it."""
    rng = np.random.default_rng(seed)
    H, W = occ.shape
    cy, cx = H // 2, W // 2
    for radius in range(0, max(H, W) // 2, 7):
        for _ in range(64):
            x = cx + int(rng.integers(-radius, radius + 1))
            y = cy + int(rng.integers(-radius, radius + 1))
            if clearance <= x < W - clearance and clearance <= y < H - clearance:
                win = occ[y - clearance:y + clearance + 1, x - clearance:x + clearance + 1]
                if np.all(win == 0):
                    th = float(rng.uniform(-math.pi, math.pi))
                    return np.array([origin[0] + (x + 0.5) * res, origin[1] + (y + 0.5) * res, th])
    raise RuntimeError("no free pose found")


def make_scan(occ, res, origin, pose, n_beams=1081, fov_deg=270.0, max_range=14.0, noise=0.01,
              seed=2):
    """Ray-cast a scan from `pose`; returns float32 [n_hits, 2] points in the base frame."""
    rng = np.random.default_rng(seed)
    H, W = occ.shape
    if fov_deg >= 360.0:
        ang = -math.pi + np.arange(n_beams) * (2 * math.pi / n_beams)
    else:
        fov = math.radians(fov_deg)
        ang = -fov / 2 + np.arange(n_beams) * (fov / (n_beams - 1))
    world = ang + pose[2]
    dx, dy = np.cos(world), np.sin(world)
    hit_range = np.full(n_beams, np.inf)
    alive = np.ones(n_beams, bool)
    step = res / 2
    for t in np.arange(step, max_range, step):
        ix = np.floor((pose[0] + t * dx - origin[0]) / res).astype(np.int64)
        iy = np.floor((pose[1] + t * dy - origin[1]) / res).astype(np.int64)
        inside = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        occd = np.zeros(n_beams, bool)
        occd[inside] = occ[iy[inside], ix[inside]] > 65
        new = alive & occd
        hit_range[new] = t
        alive &= ~new & inside
        if not alive.any():
            break
    ok = np.isfinite(hit_range)
    r = hit_range[ok] + rng.normal(0.0, noise, ok.sum())
    pts = np.stack([r * np.cos(ang[ok]), r * np.sin(ang[ok])], axis=1)
    return pts.astype(np.float32)


def gaussian_cloud(pose, n, sigma_xy=0.5, sigma_theta=0.1, seed=3):
    """Host-made tracking cloud around `pose` (for tests that set particles explicitly)."""
    rng = np.random.default_rng(seed)
    x = (pose[0] + rng.normal(0, sigma_xy, n)).astype(np.float32)
    y = (pose[1] + rng.normal(0, sigma_xy, n)).astype(np.float32)
    t = pose[2] + rng.normal(0, sigma_theta, n)
    t = ((t + math.pi) % (2 * math.pi) - math.pi).astype(np.float32)
    return x, y, t


def uniform_cloud(occ, res, origin, n, seed=3):
    """Host-made global cloud over the whole map extent (some particles land outside
    free space on purpose, to exercise the free-space penalty)."""
    rng = np.random.default_rng(seed)
    H, W = occ.shape
    x = (origin[0] + rng.uniform(-0.2, W * res + 0.2, n)).astype(np.float32)
    y = (origin[1] + rng.uniform(-0.2, H * res + 0.2, n)).astype(np.float32)
    t = rng.uniform(-math.pi, math.pi, n).astype(np.float32)
    return x, y, t


# BASELINE.json configs (SURVEY.md §8 shorthand)
CONFIGS = {
    "C1": dict(particles=2000, map=400, beams=360, fov=360.0, resample="low_variance"),
    "C2": dict(particles=100_000, map=2000, beams=1081, fov=270.0, resample="kld"),
    "C3": dict(particles=1_000_000, map=4000, beams=1081, fov=270.0, resample="adaptive"),
    "C4": dict(particles=16_000_000, map=4000, beams=1081, fov=270.0, resample="low_variance"),
}

# the node's effective defaults
NODE_LIKELIHOOD = dict(z_hit=0.9, z_rand=0.1, max_laser_distance=14.0,
                       max_obstacle_distance=2.0, sigma_hit=0.1)
NODE_ALPHA = (0.05, 0.1, 0.02, 0.05)
