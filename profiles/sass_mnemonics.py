"""Counts the tensor-core / TMA / TMEM SASS mnemonics per kernel of the built library.

    python profiles/sass_mnemonics.py [path/to/libparakeet_b200.so] > profiles/rNN_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "parakeet.cpp_b200", "libparakeet_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UCGABAR", "HMMA", "LDGSTS", "LDSM"]
EXTRA = ["ELECT", "SYNCS", "STAS"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["instr"] += 1
        for c in COLS + EXTRA:
            if op.startswith(c):
                counts[cur][c] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
for mangled, d in zip(counts, dem):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*$", "", d)
    names[mangled] = d

tot = collections.Counter()
for k in counts:
    tot.update(counts[k])
print("# SASS mnemonics of parakeet.cpp_b200/libparakeet_b200.so (cuobjdump -sass, sm_100a)")
print("# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store (cp.async.bulk.tensor), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
print("# UCGABAR = cluster barrier, HMMA = mma.sync (long-utterance / head_dim-128 attention, TDT decode, few-row GEMM), LDGSTS = cp.async,")
print("# LDSM = ldmatrix, ELECT = elect.sync, STAS = st.async (cluster statistics exchange of the fused GEMM + LayerNorm)")
print()
print("totals: " + "  ".join(f"{c} {tot[c]}" for c in COLS + EXTRA))
print()
print(f"{'kernel':<104}{'instr':>8}" + "".join(f"{c:>9}" for c in COLS))
rows = [(names[k], counts[k]) for k in counts if any(counts[k][c] for c in COLS)]
for name, c in sorted(rows):
    print(f"{name[:103]:<104}{c['instr']:>8}" + "".join(f"{c[x]:>9}" for x in COLS))
