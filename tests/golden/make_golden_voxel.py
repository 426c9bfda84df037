"""Golden vectors for the centroid-voxelisation step, produced by the reference's OWN functions.

Run in the build container (needs /root/reference; CPU only):  python tests/golden/make_golden_voxel.py

``pcdet/utils/voxel_aggregation_utils.py`` is imported UNMODIFIED from /root/reference (its one package-relative import,
``common_utils``, is satisfied by an empty stub: the functions used here never touch it); ``scatter_point_inds`` /
``generate_voxel2pinds`` are compiled from the reference's ``common_utils.py`` source text function by function (the
module itself needs SharedArray / spconv at import).  The call-site glue of pointnet2_modules.py:1323-1355 (flip, cat
of the batch index, permute) is replayed line by line.  Outputs -> tests/golden/voxel_centroids.npz.
"""
import ast
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("TSMDET_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_functions():
    for name, path in (("pcdet", os.path.join(REF, "pcdet")), ("pcdet.utils", os.path.join(REF, "pcdet", "utils"))):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    sys.modules["pcdet.utils.common_utils"] = types.ModuleType("pcdet.utils.common_utils")
    sys.modules["pcdet.utils"].common_utils = sys.modules["pcdet.utils.common_utils"]
    vau = importlib.import_module("pcdet.utils.voxel_aggregation_utils")
    src = open(os.path.join(REF, "pcdet", "utils", "common_utils.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("scatter_point_inds", "generate_voxel2pinds"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "common_utils.py", "exec"), ns)
    return vau, ns["generate_voxel2pinds"]


def call_site(vau, new_xyz, new_features, voxel_size_tensor, range_tensor):
    """pointnet2_modules.py:1325-1355, CPU tensors."""
    batch_size, channel, num_points = new_features.shape
    voxel_idxs = vau.get_voxel_indices(new_xyz.clone().view(-1, 3).contiguous(), voxel_size=voxel_size_tensor,
                                       point_cloud_range=range_tensor)
    batch_idx = new_xyz.new_zeros(size=(batch_size, num_points))
    for i in range(batch_size):
        batch_idx[i] = batch_idx[i] + i
    batch_idx = batch_idx.view(-1, 1).long()
    voxel_idxs = torch.flip(voxel_idxs, dims=[1])
    voxel_idxs = torch.cat((batch_idx, voxel_idxs), dim=-1)
    xyz_for_voxel = new_xyz.view(-1, 3)
    xyz_for_voxel = torch.cat([batch_idx, xyz_for_voxel], dim=-1)
    features_for_voxel = new_features.permute(0, 2, 1).contiguous().view(-1, channel)
    point_for_voxel = torch.cat([xyz_for_voxel, features_for_voxel], dim=-1)
    cent, cvi, cnt, inv = vau.get_centroid_per_voxel(point_for_voxel, voxel_idxs)
    return voxel_idxs, point_for_voxel, cent, cvi, cnt, inv


def main():
    vau, gen_v2p = load_reference_functions()
    rng = np.random.default_rng(0)
    b, m, c = 3, 700, 5
    lo, hi = np.array([0.0, -40.0, -3.0]), np.array([70.4, 40.0, 1.0])
    xyz = rng.uniform(lo, hi, size=(b, m, 3)).astype(np.float32)
    # crowd a third of the points into few voxels (many points per voxel), and put exact duplicates in
    xyz[:, : m // 3] = (xyz[:, :1] + rng.normal(0, 0.15, size=(b, m // 3, 3))).astype(np.float32)
    xyz[:, m - 20:] = xyz[:, :20]
    feats = rng.normal(size=(b, c, m)).astype(np.float32)
    voxel_size = [0.4, 0.4, 0.5]
    pc_range = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
    vs_t, r_t = torch.tensor(voxel_size).float(), torch.tensor(pc_range).float()
    vidx, rows, cent, cvi, cnt, inv = call_site(vau, torch.from_numpy(xyz), torch.from_numpy(feats), vs_t, r_t)
    w = torch.from_numpy(rng.integers(1, 6, size=(b * m,)).astype(np.int64))
    cent_w, cvi_w, cnt_w, inv_w = vau.get_centroid_per_voxel(rows, vidx, num_points_in_voxel=w)
    grid = ((np.array(pc_range[3:]) - np.array(pc_range[:3])) / np.array(voxel_size)).round().astype(np.int64)[::-1]
    sp = types.SimpleNamespace(indices=cvi.int(), batch_size=b, spatial_shape=list(grid))
    v2p = gen_v2p(sp)
    np.savez_compressed(
        os.path.join(HERE, "voxel_centroids.npz"), xyz=xyz, feats=feats, voxel_size=np.array(voxel_size, np.float32),
        pc_range=np.array(pc_range, np.float32), voxel_idxs=vidx.numpy(), rows=rows.numpy(), centroids=cent.numpy(),
        centroid_voxel_idxs=cvi.numpy(), labels_count=cnt.numpy(), unique_idxs=inv.numpy(), weights=w.numpy(),
        centroids_w=cent_w.numpy(), labels_count_w=cnt_w.numpy(), spatial_shape=np.array(grid),
        v2p_nonempty=np.stack(np.nonzero(v2p.numpy() >= 0), 1), v2p_values=v2p.numpy()[v2p.numpy() >= 0])
    print("voxels", cvi.shape[0], "of", b * m, "points; max points per voxel", int(cnt.max()))


if __name__ == "__main__":
    main()
