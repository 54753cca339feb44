"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / mbarrier use (B200_PROFILING.md):
cuobjdump -sass ga-dpamsa_b200/libgadpamsa.so -> profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ga-dpamsa_b200", "libgadpamsa.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "HMMA", "IDP4A", "ACQBULK", "CCTL",
        "ATOMG", "REDG", "SHFL", "LDS", "STS"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::", "", n).split("(")[0] for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = demangle(list(counts))
    rows = []
    for (mangled, c), name in zip(counts.items(), names):
        rows.append((name, c))
    out = ["# SASS summary of libgadpamsa.so (cuobjdump -sass; sm_100a). Columns: instruction counts per kernel.",
           "%-64s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in KEYS)]
    for name, c in sorted(rows, key=lambda r: -r[1]["_total"]):
        if c["_total"] == 0:
            continue
        out.append("%-64s %7d " % (name[:64], c["_total"]) + " ".join("%8d" % c[k] for k in KEYS))
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    out.append("%-64s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[k] for k in KEYS))
    text = "\n".join(out) + "\n"
    path = os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    with open(path, "w") as fh:
        fh.write(text)
    sys.stdout.write(text if "-v" in sys.argv else "written %s (%d kernels)\n" % (path, len(rows)))


main()
