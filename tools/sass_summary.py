"""Counts the Blackwell-specific SASS instructions of every kernel in libnmrgnn_b200.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),
SYNCS = mbarrier, UTCATOM / UTCCP etc. would show tensor-map TMA (none: all bulk copies are 1-D).  Writes markdown."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nmrgnn_b200", "lib", "libnmrgnn_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UTCBAR.MULTICAST", "UBLKCP", "UBLKCP.S.S", "SYNCS", "MUFU", "FFMA",
        "LDG", "STG", "LDS", "STS", "SHFL", "CCTL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("nmr::", "")
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            c = counts[name]
            c["total"] += 1
            base = op.split(".")[0]
            c[base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                c["UTCHMMA.2CTA"] += 1
            if op.startswith("UTCBAR.MULTICAST"):
                c["UTCBAR.MULTICAST"] += 1
            if op.startswith("UBLKCP.S.S"):
                c["UBLKCP.S.S"] += 1
    print("| kernel | SASS instr | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for k, c in counts.items():
        print(f"| `{k}` | {c['total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")


if __name__ == "__main__":
    main()
