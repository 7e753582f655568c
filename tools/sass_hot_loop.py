"""Static view of k_physics_resident_v2's SASS for a built library: per role segment (between the USETMAXREG instructions)
the spill instructions, and for the compute role's step loop (the longest stretch between two BAR.SYNC) the instruction
count and opcode mix.  tools/sass_hot_loop.py <libdecipher_b200.so> [kernel name fragment]"""
import collections
import subprocess
import sys

lib = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "k_physics_resident_v2ILb0ELb1"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if "Function :" in l and pat in l][0]
end = [i for i, l in enumerate(sass) if "Function :" in l and i > start]
body = [l.split("*/")[1].strip() for l in sass[start:(end[0] if end else len(sass))] if "/*" in l and ";" in l]
idx = [i for i, l in enumerate(body) if "USETMAXREG" in l]
print(f"{lib}: {len(body)} instructions, register hand-over at {idx}: {[body[i][:40] for i in idx]}")
bounds = idx + [len(body)]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = body[a:b]
    print(f"  segment {a}-{b}: LDL {sum('LDL' in l for l in seg)} STL {sum('STL' in l for l in seg)}")
seg = body[idx[0]:idx[1]] if len(idx) > 1 else body[idx[0]:]
bars = [i for i, l in enumerate(seg) if "BAR.SYNC" in l]
gap = max((bars[i + 1] - bars[i], bars[i], bars[i + 1]) for i in range(len(bars) - 1))
hot = seg[gap[1]:gap[2]]
c = collections.Counter((l.split()[1] if l.startswith("@") else l.split()[0]).split(".")[0] for l in hot)
print(f"  compute step loop: {len(hot)} instructions ({16 * len(hot) / 1024:.1f} KB), FP64 {c['DFMA'] + c['DADD'] + c['DMUL']}: "
      + ", ".join(f"{k} {c[k]}" for k in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDL", "STL", "IMAD", "LOP3", "SEL", "FSEL", "ISETP")))
