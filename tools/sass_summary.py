"""Per-kernel SASS instruction summary of librefid_b200.so: counts of the mnemonics that prove the tcgen05 / TMEM / TMA
path (UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMASTG = TMA
store, UTCCP = tcgen05.cp), per kernel family.  Runs without a GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "refid_b200", "librefid_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTMASTG", "UTCCP", "SYNCS", "MUFU", "HMMA", "ATOMG", "RED"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\((?!int\)|bool\)).*", "", name)  # drop the argument list, keep template casts like (int)64
        per[cur] = collections.Counter()
        per[cur]["instructions"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(2)
        per[cur]["instructions"] += 1
        for k in MN:
            if op.startswith(k):
                per[cur][k] += 1
fam = collections.OrderedDict()
for k, c in per.items():
    f = re.sub(r"<.*", "", k).replace("refid::", "")
    a = fam.setdefault(f, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
print(f"SASS summary of {os.path.relpath(so, ROOT)} (sm_100a; cuobjdump -sass), per kernel family: instantiations, then totals")
print(f"{'kernel':34s} {'inst':>4s} {'instr':>8s} " + " ".join(f"{m:>8s}" for m in MN))
tot = collections.Counter()
for f, (n, c) in fam.items():
    print(f"{f[:34]:34s} {n:4d} {c['instructions']:8d} " + " ".join(f"{c[m]:8d}" for m in MN))
    tot.update(c)
print(f"{'TOTAL':34s} {sum(n for n, _ in fam.values()):4d} {tot['instructions']:8d} " + " ".join(f"{tot[m]:8d}" for m in MN))
if "-v" in sys.argv:
    print("\nper instantiation (tensor-core kernels only):")
    for k, c in per.items():
        if c["UTCHMMA"]:
            print(f"  {k[:150]}: UTCHMMA {c['UTCHMMA']} UTMALDG {c['UTMALDG']} LDTM {c['LDTM']} UTCBAR {c['UTCBAR']} instr {c['instructions']}")
