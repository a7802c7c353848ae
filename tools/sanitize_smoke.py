"""Small forward calls for compute-sanitizer (memcheck / racecheck / synccheck): the reference's unit-test ring graph and
a 300-atom protein fragment through every compute path (exact-FP32 kernels, tensor-core kernels forced on the small
call, edge MLP instead of the table, column-split MP pairs), each checked against its golden peaks.
Usage:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmrgnn_b200  # noqa: E402
from conftest import load_golden, tol_ratio  # noqa: E402


def graph_of(g):
    return g["atoms"], g["nlist"], g["edges"], g["inv_degree"]


def main():
    m = nmrgnn_b200.load_model()
    h = m.handle
    for name in ("ring5_bonded", "prot300"):
        g = load_golden(name)
        graph = graph_of(g)
        for label, opts in (("exact-fp32", {"force_ffma": 1}),
                            ("tcgen05 + edge table", {"tc_min_atoms": 0}),
                            ("tcgen05 + edge MLP", {"tc_min_atoms": 0, "edge_table": 0}),
                            ("tcgen05, column-split MP pairs", {"tc_min_atoms": 0, "mp_nsplit": 1}),
                            ("tcgen05, 2 chain segments", {"tc_min_atoms": 0, "mp_chain_segments": 2}),
                            ("tcgen05, single accumulator", {"tc_min_atoms": 0, "mp_single_acc": 1})):
            for k, v in opts.items():
                h.set_option(k, v)
            y = m(graph)
            print(f"{name:14s} {label:34s} path {h.compute_path:40s} tol_ratio {tol_ratio(y, g['peaks_f64']):.3f}", flush=True)
            for k in opts:
                h.set_option(k, {"tc_min_atoms": 1024, "edge_table": 1, "mp_chain_segments": 1}.get(k, 0))
    m.close()


if __name__ == "__main__":
    main()
