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


def voxel_scene(b, n_per, m_per, seed=0, voxel=(0.4, 0.4, 0.5), rng_range=KITTI_RANGE, crowd=True):
    """Inputs of the pointnet2_stack voxel query as the SA layers >= 1 build them: `b` frames of `n_per` source points
    (one per occupied voxel: duplicates of a voxel are dropped), their dense voxel -> row table `(b, Z, Y, X)`, and
    `m_per` query centres per frame with their voxel coordinates.  Returns dict of numpy arrays (stacked layouts)."""
    r = np.random.default_rng(seed)
    lo, hi = np.array(rng_range[0]), np.array(rng_range[1])
    vs = np.array(voxel)
    grid = np.round((hi - lo) / vs).astype(np.int64)          # x, y, z cells
    pts, cnts, new, new_cnt, coords = [], [], [], [], []
    table = -np.ones((b, grid[2], grid[1], grid[0]), dtype=np.int32)
    row0 = 0
    for f in range(b):
        p = r.uniform(lo, hi, size=(n_per, 3))
        if crowd:  # a dense patch: far more than nsample hits per centre (exercises the random replacement)
            p[: n_per // 2] = p[:1] + r.normal(0, 2.5, size=(n_per // 2, 3))
            p = np.clip(p, lo, hi - 1e-3)
        p = p.astype(np.float32)
        v = np.floor((p - lo) / vs).astype(np.int64)
        v = np.minimum(v, grid - 1)
        _, first = np.unique(v[:, 2] * grid[1] * grid[0] + v[:, 1] * grid[0] + v[:, 0], return_index=True)
        first.sort()
        p, v = p[first], v[first]
        table[f, v[:, 2], v[:, 1], v[:, 0]] = row0 + np.arange(len(p), dtype=np.int32)
        q = p[r.choice(len(p), size=m_per, replace=len(p) < m_per)] + r.normal(0, 0.2, size=(m_per, 3)).astype(np.float32)
        q = np.clip(q, lo, hi - 1e-3).astype(np.float32)
        qv = np.minimum(np.floor((q - lo) / vs).astype(np.int64), grid - 1)
        pts.append(p); cnts.append(len(p)); new.append(q); new_cnt.append(m_per)
        coords.append(np.concatenate([np.full((m_per, 1), f), qv[:, ::-1]], 1).astype(np.int32))
        row0 += len(p)
    return {"xyz": np.concatenate(pts), "xyz_batch_cnt": np.array(cnts, np.int32), "new_xyz": np.concatenate(new),
            "new_xyz_batch_cnt": np.array(new_cnt, np.int32), "new_coords": np.concatenate(coords),
            "point_indices": table}
