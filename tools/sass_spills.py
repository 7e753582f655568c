"""Counts local-memory spill instructions inside the hot step loop of k_physics_resident_v2<false> (between the first and
the last MUFU of the kernel) for a built library: tools/sass_spills.py <libdecipher_b200.so>"""
import subprocess, sys
lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if "Function :" in l and "k_physics_resident_v2ILb0" in l][0]
end = [i for i, l in enumerate(sass) if "Function :" in l and i > start]
body = [l for l in sass[start:(end[0] if end else len(sass))] if "/*" in l and ";" in l]
mufu = [i for i, l in enumerate(body) if "MUFU.RCP64H" in l]
a, b = mufu[0] - 400, mufu[-1] + 400
hot = body[max(a, 0):b]
n_ldl = sum("LDL" in l for l in hot); n_stl = sum("STL" in l for l in hot)
print(f"{lib}: instructions {len(body)}, hot region {len(hot)}: LDL {n_ldl} STL {n_stl}; whole kernel LDL {sum('LDL' in l for l in body)} STL {sum('STL' in l for l in body)}")
