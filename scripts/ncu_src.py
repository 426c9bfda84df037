"""Summarise the SASS page of an .ncu-rep: samples per opcode, per stall reason, and the hottest instructions.

    python scripts/ncu_src.py gpurun_out/x.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op, by_stall, inst_op = Counter(), Counter(), Counter()
tot = 0
recs = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]]
    op = sass.split()[0] if sass and not sass.startswith("@") else (sass.split()[1] if len(sass.split()) > 1 else sass)
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("F2FP", "LDS", "STS", "LDG", "BAR", "SYNCS", "UTCBAR") else "")
    s = int(r[ix["# Samples"]] or 0)
    ie = int(r[ix["Instructions Executed"]] or 0)
    tot += s
    by_op[op] += s
    inst_op[op] += ie
    for st in stalls:
        v = int(r[ix[st]] or 0)
        by_stall[st] += v
    recs.append((s, r[ix["Address"]][-5:], sass[:90], {st: int(r[ix[st]] or 0) for st in stalls if int(r[ix[st]] or 0) > 0}))
print("total samples", tot, " total warp-instructions", sum(inst_op.values()))
print("by stall:", [(k, v, round(100 * v / max(tot, 1), 1)) for k, v in by_stall.most_common(10)])
print("by opcode (samples% | warp-instr%):")
ti = sum(inst_op.values())
for k, v in by_op.most_common(22):
    print(f"  {k:14s} {100*v/max(tot,1):5.1f}%  {100*inst_op[k]/max(ti,1):5.1f}%  n={inst_op[k]}")
print("hottest instructions:")
for s, a, sass, st in sorted(recs, key=lambda x: -x[0])[:top]:
    print(f"  {100*s/max(tot,1):5.1f}% {a} {sass}  {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
