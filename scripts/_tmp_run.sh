python -m pytest tests/test_sa_module_gpu.py -m gpu -x -q 2>&1 | tail -15
for op in mlp1 mlp2 fp; do python scripts/prof.py $op --precision tf32 --time --reps 5 2>&1 | grep "ms:\|Error\|error" ; done
python - <<'PY'
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np, bench, parity
from oracle import oracle as orc
from tsmdet_b200.pipeline import SABackboneNMS
orc.build()
dev = torch.device('cuda:0')
xyz, feats, boxes, scores = bench.make_inputs(4, 0)
for prec in ('tf32', 'bf16'):
    eng = SABackboneNMS(precision=prec).to(dev)
    args = [torch.from_numpy(a).to(dev) for a in (xyz, feats, boxes, scores)]
    res = eng.forward_device(*args); res = eng.forward_device(*args)
    torch.cuda.synchronize()
    m = parity.verify_step(eng, orc, xyz, feats, boxes, scores, res, prec)
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); eng.forward_device(*args); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print(prec, 'step ms', min(ts), {k: (round(v, 6) if isinstance(v, float) else v) for k, v in m.items()})
PY
