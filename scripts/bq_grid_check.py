"""Grid ball query (ball_query_grid.cu) vs the brute-force kernels: parity on many shapes, then timing."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from tsmdet_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
def bq(algo, rin, r, ns, xyz, new):
    os.environ["TSMDET_BQ_ALGO"] = algo
    out = pu.ball_query(r, ns, xyz, new) if rin is None else pu.ball_query_dilated(rin, r, ns, xyz, new)
    torch.cuda.synchronize(); return out
bad = 0
rng = np.random.default_rng(0)
cases = []
for gen in (synth.cloud_ground_objects, synth.cloud_uniform, synth.cloud_dup_padded, synth.cloud_lattice):
    for (b, n, m, rin, r, ns) in [(2, 16384, 4096, None, 0.2, 16), (2, 4096, 1024, None, 0.8, 32), (2, 1024, 512, None, 1.6, 32),
                                  (2, 16384, 4096, None, 0.8, 32), (2, 4096, 512, 0.4, 0.8, 32), (1, 20000, 4096, 0.2, 0.4, 32),
                                  (2, 3000, 700, None, 5.0, 64), (2, 2048, 130, None, 1e-3, 16), (1, 5000, 300, None, 100.0, 7),
                                  (1, 65536, 16384, None, 0.8, 32)]:
        cases.append((gen, b, n, m, rin, r, ns))
for (gen, b, n, m, rin, r, ns) in cases:
    xyz = gen(b, n, 7)
    sel = rng.permutation(n)[:m]
    new = np.ascontiguousarray(xyz[:, sel, :])
    new[:, 3] += 500.0                       # far outside the cloud
    new[:, 4] += np.float32(r * 0.7)         # off-sample centres
    new[:, 5, 2] -= 3 * r
    c0, i0 = bq("brute", rin, r, ns, T(xyz), T(new))
    c1, i1 = bq("grid", rin, r, ns, T(xyz), T(new))
    ok = bool(torch.equal(c0, c1) and torch.equal(i0, i1))
    if not ok:
        bad += 1
        w = (i0 != i1).nonzero()[:3].tolist()
        print("MISMATCH", gen.__name__, b, n, m, rin, r, ns, w, "cnt diff", int((c0 != c1).sum()), flush=True)
# degenerate clouds: flat, line, all the same point
flat = synth.cloud_uniform(2, 4000, 61); flat[:, :, 2] = 1.5
line = synth.cloud_uniform(2, 3000, 62); line[:, :, 1:] = 0.25
same = np.ones((2, 2000, 3), np.float32)
for name, xyz, r in (("flat", flat, 0.8), ("line", line, 0.5), ("same", same, 0.3)):
    new = np.ascontiguousarray(xyz[:, ::7, :])
    c0, i0 = bq("brute", None, r, 16, T(xyz), T(new)); c1, i1 = bq("grid", None, r, 16, T(xyz), T(new))
    if not (torch.equal(c0, c1) and torch.equal(i0, i1)): bad += 1; print("MISMATCH", name)
print("PARITY", "OK" if bad == 0 else f"{bad} BAD", "cases", len(cases) + 3, flush=True)

for (b, n, m, r, ns, gen) in [(16, 16384, 4096, 0.2, 16, "obj"), (16, 4096, 1024, 0.8, 32, "fps"), (16, 1024, 512, 1.6, 32, "fps"),
                              (16, 16384, 4096, 0.8, 32, "obj"), (16, 16384, 4096, 0.2, 16, "dup"), (8, 65536, 16384, 0.8, 32, "waymo")]:
    if gen == "waymo": xyz_np = synth.cloud_uniform(b, n, 1, synth.WAYMO_RANGE)
    else: xyz_np = {"obj": synth.cloud_ground_objects, "dup": synth.cloud_dup_padded, "fps": synth.cloud_ground_objects}[gen](b, 16384 if gen == "fps" else n, 1)
    xyz = T(xyz_np)
    if gen == "fps":  # realistic stacked-layer inputs: the cloud is itself an FPS sample
        i = pu.farthest_point_sample(xyz, n); xyz = torch.gather(xyz, 1, i.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    i = pu.farthest_point_sample(xyz, m); new = torch.gather(xyz, 1, i.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for algo in ("brute", "grid"):
        ms = timeit(lambda: bq(algo, None, r, ns, xyz, new))
        alg = b * (12 * n + 12 * m + 4 * m * ns + 4 * m)
        print(json.dumps(dict(b=b, n=n, m=m, r=r, ns=ns, gen=gen, algo=algo, ms=round(ms, 4), alg_GBs=round(alg / ms / 1e6, 1))), flush=True)
