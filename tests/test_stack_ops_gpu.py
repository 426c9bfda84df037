"""GPU parity for the pointnet2_stack ops of SURVEY.md 8 f3 (voxel query + dilated, stacked grouping, stacked FPS)
through the C ABI: vs the CPU oracle (oracle/stack_oracle.c -- incl. the restated XORWOW reservoir), vs the golden
vectors produced by the reference's own CUDA kernels, and head-to-head vs the unmodified reference extension
(oracle/_ref/pointnet2_stack_cuda.so) when it is present.  Bar: bit-exact, including every random replacement."""
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "stack_ops.npz")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ref_stack():
    from oracle import build_ref

    return build_ref.load_ref("pointnet2_stack_cuda")


def _query(sc, rng, radius, ns, stride=None, r_in=None):
    from tsmdet_b200.pointnet2_stack import pointnet2_stack_cuda as ps

    m = sc["new_coords"].shape[0]
    _, r1, r2, r3 = sc["point_indices"].shape
    idx = torch.zeros((m, ns), dtype=torch.int32, device="cuda")
    cu = torch.zeros((m, 1), dtype=torch.int32, device="cuda")
    args = (T(sc["new_xyz"]), T(sc["xyz"]), T(sc["new_coords"]), T(sc["point_indices"]))
    if stride is None:
        ps.voxel_query_wrapper(m, r1, r2, r3, ns, radius, rng[0], rng[1], rng[2], *args, idx, cu)
        return idx.cpu().numpy(), cu.cpu().numpy()[:, 0], None
    ic = torch.zeros((m, 1), dtype=torch.int32, device="cuda")
    ps.voxel_query_dilated_wrapper(m, r1, r2, r3, ns, r_in, radius, rng[0], rng[1], rng[2], stride[0], stride[1], stride[2],
                                   *args, idx, cu, ic)
    return idx.cpu().numpy(), cu.cpu().numpy()[:, 0], ic.cpu().numpy()[:, 0]


@pytest.mark.parametrize("rng,radius,ns", [((0, 0, 0), 0.4, 8), ((2, 2, 2), 0.8, 16), ((3, 8, 8), 3.2, 32),
                                           ((4, 16, 16), 6.4, 32), ((2, 4, 4), 1.6, 100)])
def test_voxel_query_vs_oracle(orc, rng, radius, ns):
    sc = synth.voxel_scene(3, 4000, 500, seed=11)
    idx, cu, _ = _query(sc, rng, radius, ns)
    widx, wcu, _ = orc.voxel_query(rng, radius, ns, sc["xyz"], sc["new_xyz"], sc["new_coords"], sc["point_indices"])
    assert np.array_equal(cu, wcu)
    assert np.array_equal(idx, widx)
    if rng[1] >= 8:
        assert int(wcu.max()) > ns, "the scene is meant to overflow nsample (random replacement path)"


@pytest.mark.parametrize("rng,stride,r_in,r_out,ns", [((2, 8, 8), (1, 1, 1), 0.8, 3.2, 16), ((4, 8, 8), (2, 2, 2), 0.0, 3.2, 32),
                                                      ((3, 6, 9), (1, 2, 3), 1.6, 6.4, 32)])
def test_voxel_query_dilated_vs_oracle(orc, rng, stride, r_in, r_out, ns):
    sc = synth.voxel_scene(2, 5000, 400, seed=12)
    idx, cu, ic = _query(sc, rng, r_out, ns, stride=stride, r_in=r_in)
    widx, wcu, wic = orc.voxel_query(rng, r_out, ns, sc["xyz"], sc["new_xyz"], sc["new_coords"], sc["point_indices"],
                                     stride=stride, former_radius=r_in)
    assert np.array_equal(cu, wcu) and np.array_equal(ic, wic) and np.array_equal(idx, widx)


def test_voxel_query_empty_and_border(orc):
    """Centres with no neighbour keep the caller's zeros with idx[0] = -1; centres on the table's border."""
    sc = synth.voxel_scene(1, 200, 64, seed=13, crowd=False)
    sc["new_xyz"][:8] += 30.0  # far from everything (their voxel coordinates stay valid)
    sc["new_coords"][8:16, 1:] = 0
    idx, cu, _ = _query(sc, (1, 1, 1), 0.5, 8)
    widx, wcu, _ = orc.voxel_query((1, 1, 1), 0.5, 8, sc["xyz"], sc["new_xyz"], sc["new_coords"], sc["point_indices"])
    assert np.array_equal(idx, widx) and np.array_equal(cu, wcu)
    assert (widx[:, 0] == -1).any()


def test_stack_ops_golden(orc):
    """The golden vectors are outputs of the reference's own CUDA kernels (tests/golden/make_golden.py)."""
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/stack_ops.npz not generated yet")
    g = np.load(GOLD)
    table = -np.ones(tuple(g["table_shape"]), dtype=np.int32)
    table[tuple(g["table_nz"].T)] = g["table_val"]
    sc = {k: g[k] for k in ("xyz", "new_xyz", "new_coords", "xyz_batch_cnt", "new_xyz_batch_cnt")}
    sc["point_indices"] = table
    for name in ("q_r4", "q_r8"):
        a = g[name + "_args"]
        idx, cu, _ = _query(sc, a[:3], float(g[name + "_radius"]), int(a[3]))
        assert np.array_equal(idx, g[name + "_idx"]) and np.array_equal(cu, g[name + "_cnt_unique"][:, 0]), name
    for name in ("d_r8", "d_s2"):
        a = g[name + "_args"]
        idx, cu, ic = _query(sc, a[:3], float(g[name + "_radii"][1]), int(a[6]), stride=a[3:6], r_in=float(g[name + "_radii"][0]))
        assert np.array_equal(idx, g[name + "_idx"]) and np.array_equal(cu, g[name + "_cnt_unique"][:, 0]), name
        assert np.array_equal(ic, g[name + "_idx_cnt"][:, 0]), name
    from tsmdet_b200.pointnet2_stack import pointnet2_utils as pu

    out = pu.grouping_operation(T(g["grp_feats"]), T(g["xyz_batch_cnt"]), T(g["grp_idx"]), T(g["new_xyz_batch_cnt"]))
    assert np.array_equal(out.cpu().numpy(), g["grp_out"])
    got = pu.stack_farthest_point_sample(T(g["sfps_xyz"]), T(g["sfps_cnt"]), T(g["sfps_npoint"]))
    assert np.array_equal(got.cpu().numpy(), g["sfps_idx"])


def test_stack_grouping_and_grad(orc):
    from tsmdet_b200.pointnet2_stack import pointnet2_utils as pu

    r = np.random.default_rng(3)
    fcnt, icnt = np.array([700, 1300, 50], np.int32), np.array([64, 200, 9], np.int32)
    for c, ns in ((5, 7), (64, 32), (130, 16)):
        feats = r.normal(size=(int(fcnt.sum()), c)).astype(np.float32)
        idx = np.concatenate([r.integers(0, fcnt[f], size=(icnt[f], ns)) for f in range(3)]).astype(np.int32)
        tf = T(feats).requires_grad_(True)
        out = pu.grouping_operation(tf, T(fcnt), T(idx), T(icnt))
        assert np.array_equal(out.detach().cpu().numpy(), orc.stack_group_points(feats, fcnt, idx, icnt))
        go = r.normal(size=out.shape).astype(np.float32)
        out.backward(T(go))
        want = orc.stack_group_points_grad(go, idx, icnt, fcnt, feats.shape[0])
        assert np.allclose(tf.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5)  # atomics: summation order differs


@pytest.mark.parametrize("sizes,npoints", [([5000, 700, 2500], [400, 100, 300]), ([16384, 9000], [1024, 512]),
                                           ([20000, 100], [256, 100]), ([1], [1])])
def test_stack_fps_vs_oracle(orc, sizes, npoints):
    """Ragged clouds: register-resident variants (<= 16 points per thread) and the global-memory one (20000 points)."""
    from tsmdet_b200.pointnet2_stack import pointnet2_utils as pu

    clouds = []
    for i, n in enumerate(sizes):
        gen = (synth.cloud_dup_padded, synth.cloud_uniform, synth.cloud_lattice)[i % 3]
        clouds.append(gen(1, n, seed=20 + i)[0])
    pts = np.concatenate(clouds).astype(np.float32)
    got = pu.stack_farthest_point_sample(T(pts), T(np.array(sizes, np.int32)), list(npoints))
    want = orc.stack_fps(pts, sizes, npoints)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want)


def test_stack_ops_vs_reference_cuda(ref_stack):
    """Head to head against the unmodified reference kernels on the same device."""
    if ref_stack is None:
        pytest.skip("oracle/_ref/pointnet2_stack_cuda.so not built")
    sc = synth.voxel_scene(2, 6000, 800, seed=31)
    m = sc["new_coords"].shape[0]
    _, r1, r2, r3 = sc["point_indices"].shape
    dev_args = (T(sc["new_xyz"]), T(sc["xyz"]), T(sc["new_coords"]), T(sc["point_indices"]))
    for rng, radius, ns in (((2, 4, 4), 1.6, 16), ((3, 8, 8), 3.2, 32)):
        idx = torch.zeros((m, ns), dtype=torch.int32, device="cuda")
        cu = torch.zeros((m, 1), dtype=torch.int32, device="cuda")
        ref_stack.voxel_query_wrapper(m, r1, r2, r3, ns, radius, rng[0], rng[1], rng[2], *dev_args, idx, cu)
        torch.cuda.synchronize()
        got_idx, got_cu, _ = _query(sc, rng, radius, ns)
        assert np.array_equal(got_idx, idx.cpu().numpy()) and np.array_equal(got_cu, cu.cpu().numpy()[:, 0])
        ic = torch.zeros((m, 1), dtype=torch.int32, device="cuda")
        idx.zero_()
        ref_stack.voxel_query_dilated_wrapper(m, r1, r2, r3, ns, 0.5 * radius, radius, rng[0], rng[1], rng[2], 1, 2, 1,
                                              *dev_args, idx, cu, ic)
        torch.cuda.synchronize()
        g2, c2, i2 = _query(sc, rng, radius, ns, stride=(1, 2, 1), r_in=0.5 * radius)
        assert np.array_equal(g2, idx.cpu().numpy()) and np.array_equal(i2, ic.cpu().numpy()[:, 0])
    # stacked FPS
    from tsmdet_b200.pointnet2_stack import pointnet2_utils as pu

    pts = np.concatenate([synth.cloud_dup_padded(1, 9000, 41)[0], synth.cloud_ground_objects(1, 4000, 42)[0]]).astype(np.float32)
    cnt, npts = np.array([9000, 4000], np.int32), np.array([700, 333], np.int32)
    temp = torch.full((pts.shape[0],), 1e10, device="cuda")
    out = torch.zeros((int(npts.sum()),), dtype=torch.int32, device="cuda")
    ref_stack.stack_farthest_point_sampling_wrapper(T(pts), temp, T(cnt), out, T(npts))
    torch.cuda.synchronize()
    got = pu.stack_farthest_point_sample(T(pts), T(cnt), T(npts))
    assert torch.equal(got, out)


def test_voxel_query_and_grouping_module(orc):
    """VoxelQueryAndGrouping(Dilated): query -> frame-local indices -> stacked grouping of xyz and features."""
    from tsmdet_b200.pointnet2_stack import voxel_query_utils as vq

    sc = synth.voxel_scene(2, 3000, 256, seed=51)
    feats = np.random.default_rng(6).normal(size=(sc["xyz"].shape[0], 12)).astype(np.float32)
    mod = vq.VoxelQueryAndGroupingDilated((2, 8, 8), (1, 1, 1), 0.4, 3.2, 16)
    gf, gx, empty, dens = mod(T(sc["new_coords"]), T(sc["xyz"]), T(sc["xyz_batch_cnt"]), T(sc["new_xyz"]),
                              T(sc["new_xyz_batch_cnt"]), T(feats), T(sc["point_indices"]))
    widx, wcu, _ = orc.voxel_query((2, 8, 8), 3.2, 16, sc["xyz"], sc["new_xyz"], sc["new_coords"], sc["point_indices"],
                                   former_radius=0.4)
    wempty = widx[:, 0] == -1
    widx = widx.copy()
    widx[wempty] = 0
    assert np.array_equal(empty.cpu().numpy(), wempty)
    live = ~wempty
    assert np.array_equal(gx.cpu().numpy()[live], sc["xyz"][widx[live]].transpose(0, 2, 1))
    assert np.array_equal(gf.cpu().numpy()[live], feats[widx[live]].transpose(0, 2, 1))
    assert np.allclose(dens.cpu().numpy()[:, 0], np.minimum(wcu / 16.0, 1.0))
    mod2 = vq.VoxelQueryAndGrouping((2, 4, 4), 1.6, 8)
    gf2, gx2, empty2, dens2 = mod2(T(sc["new_coords"]), T(sc["xyz"]), T(sc["xyz_batch_cnt"]), T(sc["new_xyz"]),
                                   T(sc["new_xyz_batch_cnt"]), T(feats), T(sc["point_indices"]))
    w2, cu2, _ = orc.voxel_query((2, 4, 4), 1.6, 8, sc["xyz"], sc["new_xyz"], sc["new_coords"], sc["point_indices"])
    live2 = w2[:, 0] != -1
    assert np.array_equal(gf2.cpu().numpy()[live2], feats[w2[live2]].transpose(0, 2, 1))
    assert np.allclose(dens2.cpu().numpy()[:, 0], cu2 / float(5 * 9 * 9))
