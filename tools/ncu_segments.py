"""Splits the SASS of a profiled kernel at barriers/branches and prints the sample share of each block (ncu source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; R = rows[2:]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iN] or 0) for r in R if len(r) > iN)
print("instructions", len(R), "samples", tot)
acc = n = ex = 0; start = 0; ops = {}
for i, r in enumerate(R):
    if len(r) <= iN: continue
    s = r[iS]; acc += int(r[iN] or 0); n += 1; ex = max(ex, int(r[iE] or 0))
    parts = s.split(); op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    ops[op] = ops.get(op, 0) + 1
    if any(k in s for k in ("BRA", "BSYNC", "BSSY", "BAR.", "EXIT", "USETMAXREG")):
        if 100 * acc / tot >= thresh:
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
            print(f"{start:5d}-{i:5d} n {n:4d} maxexec {ex:9d} samples {100*acc/tot:5.1f}% | {s.strip()[:48]:48s} | " + " ".join(f"{k}:{v}" for k, v in top))
        start = i + 1; acc = n = ex = 0; ops = {}
