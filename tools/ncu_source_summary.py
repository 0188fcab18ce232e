"""Aggregate the SASS source page of an ncu report: stall reasons, opcode mix (static, executed, sampled) and shared-memory
wavefronts.  python tools/ncu_source_summary.py <report.ncu-rep> <out.csv>"""
import csv, re, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
S, OPS, OPX, OPN, WF = Counter(), Counter(), Counter(), Counter(), Counter()
def num(x):
    try: return float(x)
    except ValueError: return 0.0
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
    for s in stalls: S[s] += num(r[col[s]])
    OPS[op] += num(r[col["# Samples"]]); OPX[op] += num(r[col["Instructions Executed"]]); OPN[op] += 1
    WF[op] += num(r[col["L1 Wavefronts Shared"]])
tot = sum(S.values()) or 1
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["section", "key", "value", "pct"])
    for k, v in sorted(S.items(), key=lambda kv: -kv[1]): w.writerow(["stall", k, int(v), "%.2f" % (100 * v / tot)])
    ts, tx = sum(OPS.values()) or 1, sum(OPX.values()) or 1
    for k, v in OPS.most_common(16): w.writerow(["opcode_samples", k, int(v), "%.2f" % (100 * v / ts)])
    for k, v in OPX.most_common(16): w.writerow(["opcode_warp_instructions_executed", k, int(v), "%.2f" % (100 * v / tx)])
    for k, v in WF.most_common(6):
        if v: w.writerow(["shared_wavefronts", k, int(v), ""])
print(open(sys.argv[2]).read())
