"""GPU parity: furthest point sampling through the C ABI vs the CPU oracle, the golden fixtures
(reference CUDA output) and -- when oracle/_ref is present -- the reference CUDA kernels themselves.
Bar: bit-exact indices."""
import os

import numpy as np
import pytest
from conftest import knob_delenv, knob_setenv
import torch

import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _dev():
    return torch.device("cuda:0")


def _fps(xyz_np, m):
    from tsmdet_b200 import pointnet2_utils as pu

    x = torch.from_numpy(xyz_np).to(_dev())
    idx = pu.farthest_point_sample(x, m)
    torch.cuda.synchronize()
    return idx.cpu().numpy()


CASES = [
    ("uniform", lambda: synth.cloud_uniform(2, 2048, 0), 256),
    ("dup", lambda: synth.cloud_dup_padded(3, 4096, 1), 1024),
    ("lattice", lambda: synth.cloud_lattice(2, 1500, 2), 700),
    ("tiny1", lambda: synth.cloud_uniform(2, 1, 3), 1),
    ("tiny7", lambda: synth.cloud_dup_padded(2, 7, 4), 7),
    ("n100", lambda: synth.cloud_dup_padded(4, 100, 5), 64),
    ("n513", lambda: synth.cloud_uniform(2, 513, 6), 200),
    ("n1000", lambda: synth.cloud_lattice(2, 1000, 7), 333),
    ("n1024", lambda: synth.cloud_dup_padded(16, 1024, 8), 512),
    ("n1025", lambda: synth.cloud_uniform(1, 1025, 9), 100),
    ("all_same", lambda: np.ones((2, 300, 3), np.float32), 50),
    ("m_gt_unique", lambda: synth.cloud_dup_padded(1, 256, 10, unique_frac=0.1), 200),
    ("n20000", lambda: synth.cloud_ground_objects(2, 20000, 11), 512),
]


@pytest.mark.parametrize("algo", ["cluster", "bucket"])
@pytest.mark.parametrize("name,gen,m", CASES, ids=[c[0] for c in CASES])
def test_fps_matches_oracle(orc, name, gen, m, algo, monkeypatch):
    """Both d-FPS kernels: the cluster kernel (fps.cu) and the spatially pruned single-CTA kernel
    (fps_bucket.cu; clouds above 16384 points stay on the cluster kernel)."""
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
    xyz = gen()
    got = _fps(xyz, m)
    want = orc.fps(xyz, m)
    assert np.array_equal(got, want), f"{name}: first mismatch at {np.argwhere(got != want)[:3]}"


@pytest.mark.parametrize("threads,pts", [(32, 4), (64, 32), (128, 8), (256, 16), (512, 8), (512, 32), (1024, 4), (1024, 16)])
def test_fps_bucket_every_launch_shape(orc, threads, pts, monkeypatch):
    """Threads x points-per-lane only changes which lane owns which (spatially sorted) point, never a pick."""
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", "bucket")
    knob_setenv(monkeypatch, "TSMDET_FPSB_T", str(threads))
    knob_setenv(monkeypatch, "TSMDET_FPSB_P", str(pts))
    cap = threads * pts
    for xyz, m in [(synth.cloud_dup_padded(2, min(cap, 4096), 20 + pts), 300),
                   (synth.cloud_lattice(3, min(cap, 2500) - 3, 30 + pts), 120),
                   (synth.cloud_ground_objects(2, cap, 40 + pts), 257)]:
        got = _fps(xyz, min(m, xyz.shape[1]))
        assert np.array_equal(got, orc.fps(xyz, min(m, xyz.shape[1])))


def test_fps_bucket_degenerate_clouds(orc, monkeypatch):
    """Grid construction edge cases: zero extent on one, two or all axes; clouds spanning huge ranges; exact
    duplicates only; initial min-distances handed in through temp."""
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", "bucket")
    rng = np.random.default_rng(7)
    flat = synth.cloud_uniform(2, 3000, 61)
    flat[:, :, 2] = 1.5
    line = synth.cloud_uniform(2, 2000, 62)
    line[:, :, 1:] = 0.25
    huge = (rng.standard_normal((2, 5000, 3)) * 1e6).astype(np.float32)
    tiny = (rng.standard_normal((2, 5000, 3)) * 1e-6).astype(np.float32)
    two = np.repeat(np.array([[[0, 0, 0], [1, 2, 3]]], np.float32), 700, axis=1)
    for name, xyz, m in [("flat", flat, 500), ("line", line, 500), ("huge", huge, 400), ("tiny", tiny, 400), ("two", two, 64)]:
        got = _fps(xyz, m)
        assert np.array_equal(got, orc.fps(xyz, m)), name
    from tsmdet_b200 import pointnet2_batch_cuda as ext

    xyz = synth.cloud_dup_padded(2, 6000, 63)
    t0 = rng.uniform(0.0, 30.0, size=(2, 6000)).astype(np.float32)
    x = torch.from_numpy(xyz).to(_dev())
    temp = torch.from_numpy(t0.copy()).to(_dev())
    idx = torch.zeros((2, 200), dtype=torch.int32, device=_dev())
    ext.farthest_point_sampling_wrapper(2, 6000, 200, x, temp, idx)
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", "cluster")
    temp2 = torch.from_numpy(t0.copy()).to(_dev())
    idx2 = torch.zeros((2, 200), dtype=torch.int32, device=_dev())
    ext.farthest_point_sampling_wrapper(2, 6000, 200, x, temp2, idx2)
    assert torch.equal(idx, idx2)
    assert torch.equal(temp.view(torch.int32), temp2.view(torch.int32))


@pytest.mark.parametrize("picks", [1, 2, 4, 8])
def test_fps_bucket_multi_pick_rounds(orc, picks, monkeypatch):
    """Rounds of up to K picks per barrier pair (fps_bucket.cu, TSMDET_FPSB_K) accept a further pick only when it is
    provably the pick the one-at-a-time algorithm would make next, so K never changes an index: duplicate-heavy
    clouds (runner-up keys equal to the candidate's), lattices (key ties across warps, resolved by reference rank),
    clouds with fewer distinct points than picks, initial min-distances from temp, and the chained bookkeeping."""
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", "bucket")
    knob_setenv(monkeypatch, "TSMDET_FPSB_K", str(picks))
    for name, xyz, m in [("dup16384", synth.cloud_dup_padded(2, 16384, 70), 2048),
                         ("obj16384", synth.cloud_ground_objects(1, 16384, 71), 4096),
                         ("lattice16384", synth.cloud_lattice(1, 16384, 72), 1500),
                         ("lattice9000", synth.cloud_lattice(2, 9000, 73, step=1.0), 900),
                         ("dup4096", synth.cloud_dup_padded(3, 4096, 74), 1024),
                         ("few_unique", synth.cloud_dup_padded(1, 12000, 75, unique_frac=0.02), 600),
                         ("all_same", np.ones((1, 10000, 3), np.float32), 40),
                         ("n129", synth.cloud_uniform(2, 129, 76), 129)]:
        got = _fps(xyz, m)
        want = orc.fps(xyz, m)
        assert np.array_equal(got, want), f"{name} K={picks}: first mismatch at {np.argwhere(got != want)[:3]}"
    from tsmdet_b200 import pointnet2_batch_cuda as ext

    rng = np.random.default_rng(77)
    xyz = synth.cloud_dup_padded(2, 16000, 78)
    t0 = rng.uniform(0.0, 30.0, size=(2, 16000)).astype(np.float32)
    x = torch.from_numpy(xyz).to(_dev())
    out = {}
    for algo in ("bucket", "cluster"):
        knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
        temp = torch.from_numpy(t0.copy()).to(_dev())
        idx = torch.zeros((2, 700), dtype=torch.int32, device=_dev())
        ext.farthest_point_sampling_wrapper(2, 16000, 700, x, temp, idx)
        out[algo] = (idx, temp)
    assert torch.equal(out["bucket"][0], out["cluster"][0])
    assert torch.equal(out["bucket"][1].view(torch.int32), out["cluster"][1].view(torch.int32))


@pytest.mark.parametrize("csize", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("threads", [128, 256, 512, 1024])
def test_fps_every_launch_shape(orc, csize, threads, monkeypatch):
    """The cluster size / block size only changes who computes what, never the result."""
    knob_setenv(monkeypatch, "TSMDET_FPS_CLUSTER", str(csize))
    knob_setenv(monkeypatch, "TSMDET_FPS_THREADS", str(threads))
    for xyz, m in [(synth.cloud_dup_padded(2, 4096, 20 + csize), 300), (synth.cloud_lattice(3, 2500, 30 + csize), 257)]:
        got = _fps(xyz, m)
        assert np.array_equal(got, orc.fps(xyz, m))


@pytest.mark.parametrize("algo", ["cluster", "bucket"])
def test_fps_kitti_full_size(orc, algo, monkeypatch):
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
    _fps_kitti_full_size(orc)


def _fps_kitti_full_size(orc):
    """BASELINE config-2 layer 1: 16384 -> 4096 on duplicate-padded and structured clouds, B=16 (oracle
    checks 2 clouds; all 16 are checked through size-independent properties)."""
    xyz = np.concatenate([synth.cloud_dup_padded(8, 16384, 40), synth.cloud_ground_objects(8, 16384, 41)], 0)
    got = _fps(xyz, 4096)
    want = orc.fps(xyz[[0, 8]], 4096)
    assert np.array_equal(got[[0, 8]], want)
    assert (got[:, 0] == 0).all() and got.min() >= 0 and got.max() < 16384
    for b in range(16):
        sel = xyz[b][got[b]]
        # distinct coordinates are never re-selected while unselected distinct points remain
        assert len(np.unique(sel, axis=0)) == len(sel)


def test_fps_large_clouds(orc):
    """Waymo-scale residency paths: 65536 (8-CTA cluster) and 180000 (16-CTA cluster / smem planes)."""
    xyz = synth.cloud_uniform(1, 65536, 50, synth.WAYMO_RANGE)
    assert np.array_equal(_fps(xyz, 600), orc.fps(xyz, 600))
    xyz = synth.cloud_uniform(1, 180000, 51, synth.WAYMO_RANGE)
    assert np.array_equal(_fps(xyz, 300), orc.fps(xyz, 300))
    xyz = synth.cloud_dup_padded(20, 16384, 52)  # more clouds than an 8-wide cluster grid holds at once
    assert np.array_equal(_fps(xyz, 128)[[0, 19]], orc.fps(xyz[[0, 19]], 128))


@pytest.mark.parametrize("algo", ["cluster", "bucket"])
def test_fps_temp_scratch_contract(orc, algo, monkeypatch):
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
    _fps_temp_scratch_contract(orc)


def _fps_temp_scratch_contract(orc):
    """temp comes in as the initial min-distance and leaves as the final one (SURVEY.md 8b ownership)."""
    from tsmdet_b200 import pointnet2_batch_cuda as ext

    xyz = synth.cloud_dup_padded(2, 3000, 60)
    x = torch.from_numpy(xyz).to(_dev())
    temp = torch.full((2, 3000), 1e10, device=_dev())
    idx = torch.zeros((2, 100), dtype=torch.int32, device=_dev())
    assert ext.farthest_point_sampling_wrapper(2, 3000, 100, x, temp, idx) == 1
    want_idx, want_temp = orc.fps(xyz, 100, return_temp=True)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(temp.cpu().numpy().view(np.uint32), want_temp.view(np.uint32))


@pytest.mark.parametrize("name,gen,n,m", [
    ("dup20000", synth.cloud_dup_padded, 20000, 700),          # 2 CTAs, many shared maxima across lanes / warps / CTAs
    ("lattice40000", synth.cloud_lattice, 40000, 500),         # 3 CTAs, equal distances between distinct points
    ("objects65536", synth.cloud_ground_objects, 65536, 900),  # 5 CTAs, largest u16-index cloud
    ("dup70001", synth.cloud_dup_padded, 70001, 600),          # 5 CTAs, indices above 65535 (slice-position mode)
    ("lattice163840", synth.cloud_lattice, 163840, 400),       # 11 CTAs (non-portable cluster size), Waymo test size
    ("objects20000", synth.cloud_ground_objects, 20000, 4096), # the reference's KITTI test shape (fast_cpc.yaml:52-56)
    ("uniform30000", synth.cloud_uniform, 30000, 2048),
], ids=["dup20000", "lattice40000", "objects65536", "dup70001", "lattice163840", "objects20000", "uniform30000"])
def test_fps_bucket_cluster_equals_cluster_kernel(name, gen, n, m, monkeypatch):
    """The pruned sampler over a CTA cluster (fps_bucket_cluster.cu, the default for 16385..240000 points) -- with one pick
    per DSMEM exchange and with rounds of several picks -- returns exactly what the brute-force cluster kernel returns,
    including the final min-distances handed back in temp."""
    from tsmdet_b200 import pointnet2_batch_cuda as ext

    xyz = torch.from_numpy(gen(2, n, 90)).to(_dev())
    outs = []
    # brute-force cluster kernel | one pick per DSMEM exchange | rounds of several picks (default list length, then 2)
    for algo, k in (("cluster", None), (None, "1"), (None, None), (None, "2")):
        if algo:
            knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
        else:
            knob_delenv(monkeypatch, "TSMDET_FPS_ALGO", raising=False)
        if k:
            knob_setenv(monkeypatch, "TSMDET_FPSC_K", k)
        else:
            knob_delenv(monkeypatch, "TSMDET_FPSC_K", raising=False)
        temp = torch.full((2, n), 1e10, device=_dev())
        idx = torch.zeros((2, m), dtype=torch.int32, device=_dev())
        ext.farthest_point_sampling_wrapper(2, n, m, xyz, temp, idx)
        torch.cuda.synchronize()
        outs.append((idx, temp))
    for i in range(1, len(outs)):
        assert torch.equal(outs[0][0], outs[i][0]), (name, i, int((outs[0][0] != outs[i][0]).sum()),
                                                     (outs[0][0] != outs[i][0]).nonzero()[:4].tolist())
        assert torch.equal(outs[0][1].view(torch.int32), outs[i][1].view(torch.int32)), (name, i)


def test_fps_weights(orc):
    from tsmdet_b200 import pointnet2_utils as pu

    for seed, (b, n, m) in enumerate([(2, 4096, 512), (3, 1000, 333), (1, 16384, 3072), (2, 37, 37)]):
        xyz = synth.cloud_dup_padded(b, n, 70 + seed)
        w = np.random.default_rng(80 + seed).uniform(0, 1, size=(b, n)).astype(np.float32) ** 2
        w[:, ::53] = 0.0
        got = pu.furthest_point_sample_weights(torch.from_numpy(xyz).to(_dev()), torch.from_numpy(w).to(_dev()), m)
        assert np.array_equal(got.cpu().numpy(), orc.fps_weights(xyz, w, m))


def test_fps_matrix_variants(orc):
    from tsmdet_b200 import pointnet2_utils as pu

    for seed, (b, n, m) in enumerate([(2, 384, 128), (1, 1500, 200), (2, 64, 64)]):
        xyz = synth.cloud_dup_padded(b, n, 90 + seed)
        x = torch.from_numpy(xyz).to(_dev())
        mat = pu.calc_dist_matrix_for_sampling(x).contiguous()
        w = np.random.default_rng(95 + seed).uniform(0, 1, size=(b, n)).astype(np.float32)
        mat_np = mat.cpu().numpy()
        assert np.array_equal(pu.furthest_point_sample_matrix(mat, m).cpu().numpy(), orc.fps_matrix(mat_np, m))
        assert np.array_equal(pu.furthest_point_sample_with_dist(mat, m).cpu().numpy(), orc.fps_matrix(mat_np, m))
        got = pu.furthest_point_sample_with_weighted_dist(mat, torch.from_numpy(w).to(_dev()), m)
        assert np.array_equal(got.cpu().numpy(), orc.fps_weighted_matrix(mat_np, w, m))


def test_fps_vs_reference_cuda(ref_pointnet2):
    """Head-to-head with the reference's own CUDA kernels (oracle/_ref) on the same device."""
    if ref_pointnet2 is None:
        pytest.skip("oracle/_ref/pointnet2_batch_cuda.so not built")
    from tsmdet_b200 import pointnet2_utils as pu

    for seed, (b, n, m, gen) in enumerate([(4, 16384, 1024, synth.cloud_dup_padded), (2, 20000, 700, synth.cloud_ground_objects),
                                           (3, 777, 300, synth.cloud_lattice)]):
        xyz = gen(b, n, 100 + seed)
        x = torch.from_numpy(xyz).to(_dev())
        temp = torch.full((b, n), 1e10, device=_dev())
        want = torch.zeros((b, m), dtype=torch.int32, device=_dev())
        ref_pointnet2.farthest_point_sampling_wrapper(b, n, m, x, temp, want)
        got = pu.farthest_point_sample(x, m)
        torch.cuda.synchronize()
        assert torch.equal(got, want)


def test_fps_golden():
    p = os.path.join(GOLD, "fps.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/fps.npz not generated yet")
    g = np.load(p)
    for name in ("uniform", "dup", "lattice", "small", "kitti"):
        xyz, idx = g[f"{name}_xyz"], g[f"{name}_idx"]
        assert np.array_equal(_fps(xyz, idx.shape[1]), idx), name


@pytest.mark.parametrize("name,gen,shortcut", [
    ("objects", lambda: synth.cloud_ground_objects(4, 16384, 70), True),     # no ties: every level takes the shortcut
    ("dup", lambda: synth.cloud_dup_padded(4, 16384, 71), True),             # duplicate points tie, same coordinates
    ("lattice", lambda: synth.cloud_lattice(3, 6000, 72), False),            # equal distances between distinct points
    ("few_unique", lambda: synth.cloud_dup_padded(2, 5000, 73, unique_frac=0.3), False),  # zero-distance picks
], ids=["objects", "dup", "lattice", "few_unique"])
@pytest.mark.parametrize("algo", ["cluster", "bucket"])
def test_fps_chained_levels_match_plain_fps(orc, name, gen, shortcut, algo, monkeypatch):
    knob_setenv(monkeypatch, "TSMDET_FPS_ALGO", algo)
    _fps_chained_levels(orc, name, gen, shortcut)


def _fps_chained_levels(orc, name, gen, shortcut):
    """Stacked samplers (4096 -> 1024 -> 512 of the previous level's centres): the chained entry point must
    return exactly what the plain sampler returns at every level, shortcut or not."""
    from tsmdet_b200 import pointnet2_utils as pu
    from tsmdet_b200.pointnet2_modules import gather_xyz

    xyz = torch.from_numpy(gen()).to(_dev())
    cur, state = xyz, None
    for m in (4096, 1024, 512):
        idx, state = pu.farthest_point_sample_chained(cur, m, state)
        want = pu.farthest_point_sample(cur, m)
        assert torch.equal(idx, want), (name, m)
        assert np.array_equal(idx[:1].cpu().numpy(), orc.fps(cur[:1].cpu().numpy(), m)), (name, m)
        cur = gather_xyz(cur, idx)
    ar = torch.arange(512, device=_dev(), dtype=torch.int32)
    # wherever the recorded run saw no tie between distinct points early enough, the prefix IS the answer
    for b in range(idx.shape[0]):
        if int(state.tie_iter[b]) >= 512 and shortcut:
            assert bool((idx[b] == ar).all())
