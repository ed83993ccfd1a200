"""SASS opcode histogram per kernel of liblns_b200 (cuobjdump -sass on the built objects): the mnemonics that prove which hardware
path a kernel uses -- UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), LDGSTS (cp.async), HMMA (legacy mma.sync), SYNCS (mbarrier), UCGABAR (cluster barrier).
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "UCGABAR", "FFMA", "STS", "LDS", "SHFL"]
print("kernel (object) : total SASS instructions | " + " ".join(OPS))
for obj in sorted(glob.glob(os.path.join(ROOT, "lns-latent-neural-pde-solver_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.Counter(), 0
    rows = []

    def flush():
        if fn and total:
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("lns::", "").replace("(anonymous namespace)::", "")
            rows.append((name, total, [counts.get(o, 0) for o in OPS]))
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            flush()
            fn, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            total += 1
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[o] += 1
    flush()
    for name, tot, cs in sorted(rows):
        if any(cs[:8]) or tot > 800:
            print(f"{name[:70]:70s} ({os.path.basename(obj)}) : {tot:6d} | " + " ".join(f"{o}={c}" for o, c in zip(OPS, cs) if c))
