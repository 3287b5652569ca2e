"""SASS opcode histogram of liblion_b200.so (no GPU needed): which Blackwell-native instructions the shipped kernels
contain.  python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), LDGSTS = cp.async,
HMMA = mma.sync (legacy tensor path), SYNCS = mbarrier, UTMALDG/UTMASTG = tensor-map TMA (cp.async.bulk.tensor)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lion_b200", "csrc", "liblion_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "SYNCS", "ELECT",
         "R2UR", "MUFU", "ATOM", "RED", "BAR", "SHFL", "REDUX", "FENCE", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("lion::", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("liblion_b200.so: %d kernels, %d SASS instructions; sm_100a only" % (len(per), sum(total.values())))
    print("\nwhole library:")
    for op in WATCH:
        if total[op]:
            print("  %-8s %7d" % (op, total[op]))
    absent = [op for op in ("UTMALDG", "UTMASTG", "HGMMA", "QGMMA") if not total[op]]
    print("  absent: " + ", ".join(absent) + "   (operands are staged with 1-D bulk copies; no tensor-map TMA, no Hopper wgmma)")
    print("\nper kernel (tensor / async-copy / tensor-memory opcodes):")
    print("  %-52s %8s %6s %6s %7s %7s %6s %6s" % ("kernel", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "LDGSTS", "HMMA", "instrs"))
    for name, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
        if c["UTCHMMA"] or c["UBLKCP"] or c["LDGSTS"] or c["HMMA"] or sum(c.values()) > 1500:
            print("  %-52s %8d %6d %6d %7d %7d %6d %6d" % (name[:52], c["UTCHMMA"], c["LDTM"], c["UTCBAR"], c["UBLKCP"], c["LDGSTS"], c["HMMA"], sum(c.values())))


if __name__ == "__main__":
    main()
