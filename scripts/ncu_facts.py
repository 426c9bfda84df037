"""Extract per-launch facts (duration, DRAM bytes, tensor-path activity, issue slots, smem wavefronts) from .ncu-rep files
into profiles/<tag>_ncu_traffic.jsonl (read by bench.py's roofline) and write the details page next to it.

    python scripts/ncu_facts.py <tag> <capture>=<file.ncu-rep> ...
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = []
for spec in sys.argv[2:]:
    cap, rep = spec.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    m = dict(zip(hdr, vals))
    f = lambda k: float(m[k].replace(",", "")) if m.get(k) not in (None, "") else None  # noqa: E731
    unit = dict(zip(hdr, rows[1]))
    dur = f("gpu__time_duration.sum")
    dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit.get("gpu__time_duration.sum", "ns"), 1e-9)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum((f(k) or 0.0) * scale.get(unit.get(k, "byte"), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    rec = {"capture": cap, "kernel": m.get("Kernel Name"), "duration_s": dur_s, "dram_bytes": dram,
           "tensor_pct": f("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
           "issue_slots_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active") or f("smsp__issue_active.avg.pct"),
           "smem_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
           "registers": f("launch__registers_per_thread"), "grid": m.get("launch__grid_size"), "block": m.get("launch__block_size")}
    out.append(rec)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{cap}.txt"), "w") as fh:
        fh.write(det)
    print(rec)
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.jsonl"), "a") as fh:
    for r in out:
        fh.write(json.dumps(r) + "\n")
