"""Count the Blackwell-native SASS mnemonics per kernel of bifold_b200/lib/libbifold_b200.so (run here: needs cuobjdump only).
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bifold_b200", "lib", "libbifold_b200.so")
MN = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "HMMA", "MUFU.EX2", "MUFU.TANH", "SYNCS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
per, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in MN:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                per[cur][k] += 1
tot = collections.Counter()
print(f"# SASS mnemonics of {os.path.relpath(LIB)} (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store,")
print("# LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  Kernels without any of them are omitted.\n")
print(f"{'kernel':96s} " + " ".join(f"{k:>12s}" for k in MN))
for name, c in per.items():
    if not any(c[k] for k in MN[:9]):
        continue
    tot.update(c)
    print(f"{demangle(name)[:96]:96s} " + " ".join(f"{c[k]:12d}" for k in MN))
print(f"{'TOTAL':96s} " + " ".join(f"{tot[k]:12d}" for k in MN))
