"""Opcode histogram of every kernel in libyolob200.so (cuobjdump -sass), written to profiles/.  Shows which kernels
carry the Blackwell-native instructions: UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync).  Usage: python tools/sass_summary.py [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "yolov3_tensorflow_b200", "libyolob200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
kern, hist, order = None, {}, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
KEY = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDSM", "ATOM", "RED", "MUFU")
with open(out, "w") as f:
    f.write("# SASS opcode summary of yolov3_tensorflow_b200/libyolob200.so (sm_100a) — tools/sass_summary.py\n")
    f.write("# columns: total instructions | counts of the opcode families that matter (prefix match)\n")
    tot = collections.Counter()
    for k, n in zip(order, names):
        h = hist[k]
        fam = {key: sum(v for op, v in h.items() if op.startswith(key)) for key in KEY}
        for key, v in fam.items():
            tot[key] += v
        short = re.sub(r"\(.*", "", n).replace("void yb::", "").replace("(anonymous namespace)::", "")
        f.write(f"{sum(h.values()):6d} | " + " ".join(f"{key}={v}" for key, v in fam.items() if v) + f" | {short}\n")
    f.write("# library totals: " + " ".join(f"{k}={v}" for k, v in tot.items() if v) + "\n")
    variants = collections.Counter()
    for k in order:
        for op, v in hist[k].items():
            if op.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR")):
                variants[op] += v
    f.write("# Blackwell opcode variants: " + ", ".join(f"{k} x{v}" for k, v in sorted(variants.items())) + "\n")
print(open(out).read()[-1500:])
