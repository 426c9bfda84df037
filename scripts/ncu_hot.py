"""Top SASS instructions of an ncu report by stall samples: python scripts/ncu_hot.py <report.ncu-rep> [N]
(reads `ncu -i <rep> --page source --csv`; used here, on the CPU box, to read the captures gpurun brought back)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions {tot_inst}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
ranked = sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:top]
for i, r in sorted(ranked, key=lambda ir: ir[0]):
    st = {s[6:]: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0)}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} ({100.0 * int(r[col['# Samples']]) / max(tot, 1):4.1f}%) x{int(r[col['Instructions Executed']]):8d}  {r[col['Source']].strip()[:90]:90s} {main}")
