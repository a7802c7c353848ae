"""Small forward calls for compute-sanitizer (memcheck / racecheck / synccheck): the reference's unit-test ring graph and
a 300-atom protein fragment through every compute path (exact-FP32 kernels, tensor-core kernels forced on the small
call -- the pipelined node MLP is their default --, edge MLP instead of the table, column-split MP pairs), each checked against its golden peaks.
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
                            ("tcgen05, single accumulator", {"tc_min_atoms": 0, "mp_single_acc": 1}),
                            ("tcgen05, round-1 node MLP", {"tc_min_atoms": 0, "fc_pipe": 0}),
                            ("tcgen05, node MLP on CTA pairs", {"tc_min_atoms": 0, "fc_pair": 1})):
            for k, v in opts.items():
                h.set_option(k, v)
            y = m(graph)
            print(f"{name:14s} {label:34s} path {h.compute_path:40s} tol_ratio {tol_ratio(y, g['peaks_f64']):.3f}", flush=True)
            for k in opts:
                h.set_option(k, {"tc_min_atoms": 1024, "edge_table": 1, "mp_chain_segments": 1, "fc_pipe": 1}.get(k, 0))
    # GPU graph builder: cell list and brute force on a 900-atom fragment, two graphs in one call
    from nmrgnn_b200 import _capi
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "g108m_structure.npz")) as z:
        pos = np.ascontiguousarray(z["positions_A"].astype(np.float32)[:900] / np.float32(10))
    pos2 = np.ascontiguousarray(np.concatenate([pos, pos[:300] + 1.0], 0))
    offs = np.array([0, 900, 1200], np.int64)
    res = []
    for mode in (1, 0):
        h.set_option("knn_cells", mode)
        nl, ed, inv = np.empty((1200, 16), np.int32), np.empty((1200, 16), np.float32), np.empty(1200, np.float32)
        h.knn_graph(pos2, offs, 1200, 2, 16, 0.0, nl, ed, inv, _capi.MEM_HOST)
        res.append((nl, ed, inv))
    h.set_option("knn_cells", 1)
    print("knn graph builder: cell list == brute force:", all(np.array_equal(a, b) for a, b in zip(*res)), flush=True)
    m.close()


if __name__ == "__main__":
    main()
