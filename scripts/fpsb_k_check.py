"""Multi-pick rounds of the bucketed FPS (fps_bucket.cu, TSMDET_FPSB_K): parity against the brute-force cluster
kernel over many clouds / launch shapes / K, chained-level parity, then timing per K."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from tsmdet_b200 import pointnet2_utils as pu

dev = torch.device("cuda:0")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

def run(xyz, m, algo, T=None, P=None, K=None):
    os.environ["TSMDET_FPS_ALGO"] = algo
    for k, v in (("TSMDET_FPSB_T", T), ("TSMDET_FPSB_P", P), ("TSMDET_FPSB_K", K)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    out = pu.farthest_point_sample(xyz, m); torch.cuda.synchronize()
    return out

cases = [
    ("uniform2048", synth.cloud_uniform(2, 2048, 0), 256),
    ("dup4096", synth.cloud_dup_padded(3, 4096, 1), 1024),
    ("lattice1500", synth.cloud_lattice(2, 1500, 2), 700),
    ("tiny1", synth.cloud_uniform(2, 1, 3), 1),
    ("tiny7", synth.cloud_dup_padded(2, 7, 4), 7),
    ("n100", synth.cloud_dup_padded(4, 100, 5), 64),
    ("n513", synth.cloud_uniform(2, 513, 6), 200),
    ("n1000", synth.cloud_lattice(2, 1000, 7), 333),
    ("n1025", synth.cloud_uniform(1, 1025, 9), 100),
    ("all_same", np.ones((2, 300, 3), np.float32), 50),
    ("m_gt_unique", synth.cloud_dup_padded(1, 256, 10, unique_frac=0.1), 200),
    ("m_eq_n", synth.cloud_uniform(2, 600, 11), 600),
    ("kitti_dup", synth.cloud_dup_padded(4, 16384, 40), 4096),
    ("kitti_obj", synth.cloud_ground_objects(4, 16384, 41), 4096),
    ("kitti_uni", synth.cloud_uniform(4, 16384, 42), 4096),
    ("n16000", synth.cloud_ground_objects(2, 16000, 43), 1000),
    ("n9000", synth.cloud_dup_padded(2, 9000, 44), 1000),
    ("lattice16384", synth.cloud_lattice(2, 16384, 45), 2048),
    ("lattice4096", synth.cloud_lattice(2, 4096, 46, step=1.0), 4096),
]
bad = 0
for name, xyz_np, m in cases:
    xyz = torch.from_numpy(xyz_np).to(dev)
    want = run(xyz, m, "cluster").cpu().numpy()
    n = xyz_np.shape[1]
    shapes = [(None, None, 1), (None, None, 2), (None, None, 4), (None, None, 8)]
    for T in (32, 128, 512, 1024):
        for P in (4, 8, 16, 32):
            if T * P >= n and T * P <= 16384: shapes.append((T, P, 4))
    for (T, P, K) in shapes:
        try:
            got = run(xyz, m, "bucket", T, P, K).cpu().numpy()
        except Exception as ex:
            print("FAIL", name, T, P, K, ex, flush=True); bad += 1; continue
        if not np.array_equal(got, want):
            bad += 1
            w = np.argwhere(got != want)
            print("MISMATCH", name, T, P, K, "first", w[:3].tolist(), got[tuple(w[0])], want[tuple(w[0])], "count", len(w), flush=True)
    print("checked", name, len(shapes), "shapes", flush=True)
print("PARITY", "OK" if bad == 0 else f"{bad} BAD", flush=True)

# chained bookkeeping: chained levels equal plain FPS, and the records equal the K = 1 records
for name, xyz_np in (("obj", synth.cloud_ground_objects(4, 16384, 50)), ("dup", synth.cloud_dup_padded(4, 16384, 51)),
                     ("lattice", synth.cloud_lattice(2, 16384, 52))):
    xyz = torch.from_numpy(xyz_np).to(dev)
    recs = {}
    for K in (1, 4):
        os.environ["TSMDET_FPS_ALGO"] = "bucket"; os.environ.pop("TSMDET_FPSB_T", None); os.environ.pop("TSMDET_FPSB_P", None)
        os.environ["TSMDET_FPSB_K"] = str(K)
        i1, rec1 = pu.farthest_point_sample_chained(xyz, 4096)
        c1 = pu.gather_operation(xyz.transpose(1, 2).contiguous(), i1).transpose(1, 2).contiguous()
        i2, rec2 = pu.farthest_point_sample_chained(c1, 1024, rec1)
        torch.cuda.synchronize()
        os.environ["TSMDET_FPS_ALGO"] = "cluster"
        p1 = pu.farthest_point_sample(xyz, 4096); p2 = pu.farthest_point_sample(c1, 1024)
        recs[K] = rec1
        print("chain", name, "K", K, bool((i1 == p1).all()), bool((i2 == p2).all()), flush=True)
    try:
        a, b = recs[1], recs[4]
        print("chain records equal", name, bool((a.tie_iter == b.tie_iter).all()), bool((a.vals == b.vals).all()), a.tie_iter.tolist(), flush=True)
    except Exception as ex:
        print("chain record compare skipped:", ex, flush=True)

# timing
for (b, n, m, gen) in [(16, 16384, 4096, "obj"), (16, 16384, 4096, "uni"), (16, 16384, 4096, "dup"), (16, 8192, 2048, "obj"),
                       (16, 4096, 1024, "obj"), (16, 1024, 512, "obj"), (148, 16384, 4096, "obj")]:
    xyz_np = {"obj": synth.cloud_ground_objects, "uni": synth.cloud_uniform, "dup": synth.cloud_dup_padded}[gen](b, n, 1)
    xyz = torch.from_numpy(xyz_np).to(dev)
    for K in (1, 2, 4, 8):
        ms = timeit(lambda: run(xyz, m, "bucket", None, None, K))
        print(json.dumps(dict(b=b, n=n, m=m, gen=gen, K=K, ms=round(ms, 4), us_per_pick=round(1000 * ms / (m - 1), 4))), flush=True)
    if n == 4096:
        for (T, P) in ((512, 8), (1024, 4), (256, 16)):
            for K in (1, 4):
                ms = timeit(lambda: run(xyz, m, "bucket", T, P, K))
                print(json.dumps(dict(b=b, n=n, m=m, gen=gen, T=T, P=P, K=K, ms=round(ms, 4), us_per_pick=round(1000 * ms / (m - 1), 4))), flush=True)
