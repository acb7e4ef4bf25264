"""Per-kernel SASS opcode counts of libzett_b200.so (cuobjdump -sass): the evidence that the hot path is tcgen05 / TMEM / TMA
(UTC*MMA, LDTM, UTMALDG, UBLKCP, UTCBAR) and not a legacy tensor path (HMMA).  Writes profiles/sass_opcodes_<round>.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zett_b200", "lib", "libzett_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACCTL", "SYNCS", "ELECT", "R2UR",
         "HMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "MUFU", "FFMA", "F2FP", "ATOM", "RED", "BAR", "UCGABAR"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    out = ["SASS opcode counts per kernel, cuobjdump -sass zett_b200/lib/libzett_b200.so (round %s build)" % rnd,
           "tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR;",
           "HMMA would be a legacy mma.sync path (none expected).", ""]
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    out.append("%-72s %7s " % ("kernel", "instrs") + " ".join("%8s" % c for c in cols))
    for name, c in kernels.items():
        out.append("%-72s %7d " % (name[:72], c["_total"]) + " ".join("%8d" % c[w] for w in cols))
    path = os.path.join(ROOT, "profiles", "sass_opcodes_%s.txt" % rnd)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
