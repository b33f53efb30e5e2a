#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libdrnb200.so -> profiles/r02_sass_opcodes.txt: the evidence that the hot kernels are
Blackwell-native (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP;
legacy mma.sync would show as HMMA).

  python tools/sass_histogram.py [lib.so] [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "STSM", "LDSM",
         "FFMA", "FSETP", "LDG", "STG", "LDS", "STS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "video-seg-model-compress_b200", "drnb200", "libdrnb200.so")
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    lines = ["# SASS opcode counts per kernel of %s (cuobjdump -sass; static instruction counts)" % os.path.basename(lib),
             "# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy,",
             "# SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (none).", "",
             "%-86s %6s  %s" % ("kernel", "instrs", " ".join("%7s" % w for w in WATCH))]
    tot = collections.Counter()
    for name, c in per.items():
        d = demangle(name)
        d = re.sub(r"\(.*", "", d).replace("void drnb200::", "").replace("drnb200::", "")
        lines.append("%-86s %6d  %s" % (d[:86], sum(c.values()), " ".join("%7d" % c.get(w, 0) for w in WATCH)))
        tot.update(c)
    lines.append("%-86s %6d  %s" % ("TOTAL", sum(tot.values()), " ".join("%7d" % tot.get(w, 0) for w in WATCH)))
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines[-12:]))


if __name__ == "__main__":
    main()
