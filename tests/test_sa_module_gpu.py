"""GPU parity: the fused set-abstraction scale (gather + MLP + max-pool in one kernel) against the
reference's eager sequence (materialised groups, torch Conv2d/BatchNorm2d/ReLU/max_pool2d, fp32,
TF32 off).  Tolerances from BASELINE.json north_star: max-abs 1e-5 (fp32 path; relaxed to 2e-5 x
output scale), 1e-2 relative (bf16 tensor-core path)."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def D():
    return torch.device("cuda:0")


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(D())


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _randomise_bn(mod, seed):
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def _layer(precision, **kw):
    from tsmdet_b200.pointnet2_modules import PointnetSAModuleFSMSG

    torch.manual_seed(0)
    layer = PointnetSAModuleFSMSG(fused=True, precision=precision, **kw).to(D()).eval()
    _randomise_bn(layer, 1)
    return layer


CFGS = {
    "kitti_l1": dict(npoint_list=[512], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                     radii=[0.2], nsamples=[16], mlps=[[1, 16, 16, 32]]),
    "kitti_l2": dict(npoint_list=[256], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                     radii=[0.8], nsamples=[32], mlps=[[32, 64, 64, 128]]),
    "kitti_l3": dict(npoint_list=[128], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                     radii=[1.6], nsamples=[32], mlps=[[128, 128, 128, 256]]),
    "ref_layer0": dict(npoint_list=[1024], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                       radii=[0.2, 0.4, 0.8], nsamples=[32, 32, 32],
                       mlps=[[1, 16, 16, 32], [1, 16, 16, 32], [1, 32, 32, 64]], dilated_radius_group=True,
                       aggregation_mlp=[64]),
    "odd": dict(npoint_list=[100], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                radii=[1.0], nsamples=[12], mlps=[[5, 24, 40]], skip_connection=True),
}


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_sa_matches_eager(name, precision):
    from tsmdet_b200 import _lib

    cfg = CFGS[name]
    c_in = cfg["mlps"][0][0]
    n = 4096
    xyz = T(synth.cloud_ground_objects(2, n, 5))
    feats = torch.randn((2, c_in, n), generator=torch.Generator().manual_seed(3)).to(D())
    layer = _layer(precision, **cfg)
    with torch.no_grad():
        layer.fused = False
        want_xyz, want, want_idx = layer(xyz, feats)
        layer.fused = True
        try:
            got_xyz, got, got_idx = layer(xyz, feats)
        except _lib.TsmdetError as e:
            if precision == "bf16" and e.code == 1000001:
                pytest.skip("shape not supported by the tensor-core path (documented fallback: fp32)")
            raise
    torch.cuda.synchronize()
    assert torch.equal(got_idx, want_idx) and torch.equal(got_xyz, want_xyz)
    assert got.shape == want.shape
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    if precision == "fp32":
        assert err <= 2e-5 * max(scale, 1.0), f"max abs err {err} (scale {scale})"
    else:
        assert err <= 1e-2 * max(scale, 1.0), f"max abs err {err} (scale {scale})"


def test_empty_ball_outputs_relu_bias_chain():
    """Empty balls see an all-zero INPUT, so they output the ReLU(bias) chain, not zeros."""
    cfg = CFGS["kitti_l2"]
    layer = _layer("fp32", **cfg)
    xyz = T(synth.cloud_uniform(1, 2048, 9))
    feats = torch.randn((1, 32, 2048)).to(D())
    new_xyz = xyz[:, :64, :].clone().contiguous()
    new_xyz[:, 0] += 1000.0
    with torch.no_grad():
        _, fused, _ = layer(xyz, feats, new_xyz=new_xyz)
        layer.fused = False
        _, eager, _ = layer(xyz, feats, new_xyz=new_xyz)
    assert torch.allclose(fused[:, :, 0], eager[:, :, 0], atol=1e-5)
    assert float(eager[:, :, 0].abs().max()) > 0


def test_fp_module_matches_eager():
    """PointnetFPModule: three_nn -> weights -> three_interpolate -> MLP vs a plain torch restatement."""
    from tsmdet_b200.pointnet2_modules import PointnetFPModule

    torch.manual_seed(0)
    fp = PointnetFPModule(mlp=[16 + 4, 32, 32]).to(D()).eval()
    _randomise_bn(fp, 2)
    unknown = T(synth.cloud_ground_objects(2, 2000, 1))
    known = T(synth.cloud_ground_objects(2, 500, 2))
    uf = torch.randn((2, 4, 2000)).to(D())
    kf = torch.randn((2, 16, 500)).to(D())
    with torch.no_grad():
        got = fp(unknown, known, uf, kf)
        d = torch.cdist(unknown.double(), known.double())
        dist, idx = torch.topk(d, 3, dim=2, largest=False)
        w = 1.0 / (dist.float() + 1e-8)
        w = w / w.sum(2, keepdim=True)
        g = torch.gather(kf.unsqueeze(2).expand(-1, -1, 2000, -1), 3, idx.unsqueeze(1).expand(-1, 16, -1, -1))
        interp = (g * w.unsqueeze(1)).sum(-1)
        want = fp.mlp(torch.cat([interp, uf], 1).unsqueeze(-1)).squeeze(-1)
    assert got.shape == (2, 32, 2000)
    assert float((got - want).abs().max()) < 1e-3


def test_kitti_stack_runs_and_is_deterministic():
    from tsmdet_b200.pointnet2_modules import kitti_sa_stack

    torch.manual_seed(0)
    net = kitti_sa_stack(fused=True, precision="fp32").to(D()).eval()
    xyz = T(synth.cloud_ground_objects(2, 16384, 3))
    feats = torch.rand((2, 1, 16384)).to(D())
    with torch.no_grad():
        o1 = net(xyz, feats)
        o2 = net(xyz, feats)
    assert [tuple(o[1].shape) for o in o1] == [(2, 32, 4096), (2, 128, 1024), (2, 256, 512)]
    for a, b in zip(o1, o2):
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_pipeline_graph_matches_stream_and_module_paths():
    """The captured multi-stream CUDA graph, the eager multi-stream DAG and the plain module stack agree."""
    from tsmdet_b200.pipeline import SABackboneNMS

    eng = SABackboneNMS(precision="fp32", use_graph=True).to(D())
    xyz = T(synth.cloud_ground_objects(3, 16384, 5))
    feats = torch.rand((3, 1, 16384), generator=torch.Generator().manual_seed(1)).to(D())
    boxes = T(np.stack([synth.boxes_clustered(1024, 20 + i, centres=80) for i in range(3)]))
    scores = T(np.stack([synth.scores_random(1024, 30 + i) for i in range(3)]))
    r1 = {k: v.clone() for k, v in eng.forward_device(xyz, feats, boxes, scores).items()}
    r2 = {k: v.clone() for k, v in eng.forward_device(xyz, feats, boxes, scores).items()}  # replay
    eng.use_graph = False
    r3 = eng.forward_device(xyz, feats, boxes, scores)
    torch.cuda.synchronize()
    for k in r1:
        assert torch.equal(r1[k], r2[k]) and torch.equal(r1[k], r3[k]), k
    with torch.no_grad():
        outs = eng.backbone(xyz, feats)
    assert torch.equal(outs[-1][0], r1["xyz"]) and torch.equal(outs[-1][1], r1["features"])
    # NMS leg vs the single-frame reference-shaped API
    from tsmdet_b200 import iou3d_nms_utils as iu
    for f in range(3):
        k1, _ = iu.nms_gpu(boxes[f], scores[f], 0.01)
        k2, _ = iu.nms_gpu(boxes[f][k1].contiguous(), scores[f][k1].contiguous(), 0.1)
        want = k1[k2][:512]
        n = int(r1["det_num"][f])
        assert n == want.numel() and torch.equal(r1["det_idx"][f, :n], want)
        assert bool((r1["det_idx"][f, n:] == -1).all())
