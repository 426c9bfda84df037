"""Seeded synthetic inputs: KITTI-/Waymo-shaped clouds and rotated boxes."""
from __future__ import annotations

import numpy as np

KITTI_RANGE = ((0.0, -40.0, -3.0), (70.4, 40.0, 1.0))     # fast_cpc.yaml:75
WAYMO_RANGE = ((-75.2, -75.2, -2.0), (75.2, 75.2, 4.0))   # waymo_dataset.yaml:6


def cloud_uniform(b, n, seed=0, rng_range=KITTI_RANGE):
    r = np.random.default_rng(seed)
    lo, hi = np.array(rng_range[0]), np.array(rng_range[1])
    return r.uniform(lo, hi, size=(b, n, 3)).astype(np.float32)


def cloud_ground_objects(b, n, seed=0):
    """70 % ground plane z = -1.7 + N(0,0.05), 30 % in 40 Gaussian clusters (sigma 0.6)."""
    r = np.random.default_rng(seed)
    out = np.empty((b, n, 3), np.float32)
    for i in range(b):
        ng = int(0.7 * n)
        g = r.uniform(KITTI_RANGE[0], KITTI_RANGE[1], size=(ng, 3))
        g[:, 2] = -1.7 + r.normal(0, 0.05, ng)
        centres = r.uniform((5, -30, -1.5), (65, 30, 0.0), size=(40, 3))
        which = r.integers(0, 40, n - ng)
        o = centres[which] + r.normal(0, 0.6, size=(n - ng, 3))
        pts = np.concatenate([g, o], 0)
        r.shuffle(pts)
        out[i] = pts.astype(np.float32)
    return out


def cloud_dup_padded(b, n, seed=0, unique_frac=0.75):
    """unique_frac*n unique points + exact duplicates of random rows, shuffled
    (mirrors data_processor.py:180-185; exercises FPS ties)."""
    r = np.random.default_rng(seed)
    out = np.empty((b, n, 3), np.float32)
    nu = max(1, int(unique_frac * n))
    for i in range(b):
        u = r.uniform(KITTI_RANGE[0], KITTI_RANGE[1], size=(nu, 3)).astype(np.float32)
        d = u[r.integers(0, nu, n - nu)]
        pts = np.concatenate([u, d], 0)
        r.shuffle(pts)
        out[i] = pts
    return out


def cloud_lattice(b, n, seed=0, step=0.5):
    """Points on a coarse lattice: many exactly equal distances (worst case for tie-breaking)."""
    r = np.random.default_rng(seed)
    g = r.integers(0, 24, size=(b, n, 3)).astype(np.float32) * step
    return g


def boxes_random(n, seed=0, rng_range=KITTI_RANGE):
    r = np.random.default_rng(seed)
    lo, hi = np.array(rng_range[0]), np.array(rng_range[1])
    c = r.uniform(lo, hi, size=(n, 3))
    d = r.uniform((1.5, 1.2, 1.2), (4.5, 2.2, 2.0), size=(n, 3))
    h = r.uniform(-np.pi, np.pi, size=(n, 1))
    return np.concatenate([c, d, h], 1).astype(np.float32)


def boxes_clustered(n, seed=0, centres=200, sigma=0.3, sigma_theta=0.1):
    """`centres` objects x ~n/centres jittered copies, so NMS really suppresses."""
    r = np.random.default_rng(seed)
    base = boxes_random(centres, seed + 1000)
    which = r.integers(0, centres, n)
    bx = base[which].copy()
    bx[:, 0:2] += r.normal(0, sigma, size=(n, 2))
    bx[:, 2] += r.normal(0, 0.05, size=n)
    bx[:, 3:6] *= r.uniform(0.9, 1.1, size=(n, 3))
    bx[:, 6] += r.normal(0, sigma_theta, size=n)
    return bx.astype(np.float32)


def scores_random(n, seed=0):
    return np.random.default_rng(seed + 7).uniform(0, 1, n).astype(np.float32)
