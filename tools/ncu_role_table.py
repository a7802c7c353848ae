"""Per-role warp-stall attribution of a warp-specialised kernel from an `ncu --set full --import-source on` report.
Usage: ncu_role_table.py <report.ncu-rep> <kernel: mp | fc>  -> markdown on stdout.
The SASS of a kernel follows the order of its role branches; every SASS instruction is given the role of the nearest
preceding instruction (by address) whose CUDA line lies in kernels_tc.cuh, with the roles' line ranges read from the
"=====" markers of the source.  Samples are the profiler's warp-state samples (one per sampling period and resident warp)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "nmrgnn_b200", "csrc", "kernels_tc.cuh")
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_selected", "stall_not_selected", "stall_barrier", "stall_math",
          "stall_mio", "stall_lg", "stall_dispatch", "stall_branch_resolving", "stall_sleep", "stall_no_inst", "stall_membar"]


def role_ranges(kernel):
    lines = open(SRC).read().split("\n")
    start = next(i for i, l in enumerate(lines) if ("void mp_layer_tc_body" if kernel == "mp" else "void fc_readout_tc_body") in l)
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("template <int ACT>"))
    marks = [(i + 1, re.sub(r"[= /]+", " ", lines[i]).strip()) for i in range(start, end) if "=====================" in lines[i]]
    rng = [(start + 1, "prologue")] + marks
    # the code after the last role branch: the CTA-wide barrier at the end of the kernel (warps that are done wait here)
    tail = next(i for i in range(marks[-1][0], end) if lines[i].startswith("  tc::tc_fence_before();"))
    rng.append((tail + 1, "kernel tail: finished warps waiting for the CTA's last role"))
    return rng, end


HELPERS = [  # tc_common.cuh / common.cuh helpers that are inlined into several roles: (file, first line, last line, label)
    ("tc_common.cuh", 24, 99, "mbarrier waits / arrives (all roles)"),
    ("tc_common.cuh", 305, 328, "mbarrier waits / arrives (all roles)"),
    ("tc_common.cuh", 112, 123, "bulk copies + proxy fences (loaders, producers)"),
    ("tc_common.cuh", 124, 147, "shared-memory loads / stores (producers: edge records, operand tiles)"),
    ("tc_common.cuh", 153, 158, "global 16-byte loads (producers: gathers; epilogue: residual)"),
    ("tc_common.cuh", 159, 219, "tcgen05 fences / ld / alloc (epilogue, MMA)"),
    ("tc_common.cuh", 403, 423, "tcgen05 fences / ld / alloc (epilogue, MMA)"),
    ("tc_common.cuh", 220, 240, "descriptors + tcgen05.mma / commit (MMA issuer)"),
    ("tc_common.cuh", 278, 290, "descriptors + tcgen05.mma / commit (MMA issuer)"),
    ("tc_common.cuh", 387, 402, "fp16 hi / lo split (producers)"),
    ("common.cuh", 1, 200, "activation (epilogue)"),
]


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    rng, end = role_ranges(kernel)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur_file = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = os.path.basename(r[1])
        elif r and r[0] == "Line No":
            hdr = r
            ci = {n: i for i, n in enumerate(hdr)}
        elif hdr and len(r) == len(hdr) and r[0].isdigit():          # per-source-line aggregate over all its SASS
            line = int(r[0])
            if cur_file == "kernels_tc.cuh" and rng[0][0] <= line < end:
                role = [name for l0, name in rng if l0 <= line][-1]
            else:
                role = next((lab for f, l0, l1, lab in HELPERS if f == cur_file and l0 <= line <= l1), f"other ({cur_file})")
            a = agg.setdefault(role, collections.Counter())
            a["samples"] += int(r[ci["# Samples"]])
            a["inst"] += int(r[ci["Instructions Executed"]])
            for st in STALLS:
                if st in ci:
                    a[st] += int(r[ci[st]])
    tot = sum(a["samples"] for a in agg.values())
    print(f"total warp-state samples {tot}")
    print()
    print("| role (lines of the role in kernels_tc.cuh) or inlined helper | samples | share | warp instructions | top stall reasons (share of the row's samples) |")
    print("|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] == 0:
            continue
        top = sorted(((a[st], st) for st in STALLS), reverse=True)[:5]
        print(f"| {name} | {a['samples']} | {100 * a['samples'] / tot:.1f} % | {a['inst']} | "
              + ", ".join(f"{st[6:]} {100 * v / a['samples']:.0f} %" for v, st in top if v) + " |")


if __name__ == "__main__":
    main()
