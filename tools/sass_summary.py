"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): which kernels issue tcgen05.mma (UTCHMMA),
TMA loads (UTMALDG), TMEM loads / stores (LDTM / STTM), tcgen05.commit (UTCBAR), mma.sync (HMMA), MUFU ...
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stedm_b200", "libstedm_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOM", "SYNCS", "HMMA",
         "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ELECT"]


def main():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True).stdout
    demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                              text=True).stdout.splitlines()
    names = iter(demangle)
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            order.append(cur)
            counts[cur]["_arch"] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(order)} kernels, arch {archs}; instruction counts per kernel (static SASS)")
    cols = [w for w in WATCH if any(counts[k][w] for k in order)]
    print(f"{'kernel':66s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in cols))
    for k in order:
        print(f"{k[:66]:66s} {counts[k]['_total']:6d} " + " ".join(f"{counts[k][c]:7d}" for c in cols))
    tot = collections.Counter()
    for k in order:
        tot.update({c: counts[k][c] for c in cols})
    print(f"{'TOTAL':66s} {sum(counts[k]['_total'] for k in order):6d} " + " ".join(f"{tot[c]:7d}" for c in cols))


if __name__ == "__main__":
    main()
