"""SASS mnemonic histogram per kernel of libep_b200.so (evidence for profiles/): python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "efficient-probing_b200", "csrc", "libep_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
keys = ["UTCHMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "ACQBULK", "PREEXIT", "HMMA", "LDSM", "MUFU", "F2FP", "F2F", "STG", "LDG", "STS", "LDS", "SHFL"]
print("SASS mnemonic counts per kernel of libep_b200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,")
print("UTMAPF = TMA L2 prefetch, UBLKCP = bulk copy, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync (legacy path).\n")
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for fn, c in hist.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("ep::", "")
    cnt = lambda k: sum(v for kk, v in c.items() if kk == k or kk.startswith(k + "."))
    print(f"{name[:64]:64s} {sum(c.values()):6d} " + " ".join(f"{cnt(k):7d}" for k in keys))
