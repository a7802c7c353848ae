#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the reference's traced SavedModel
graph (oracle/savedmodel_interp.py) on fixed inputs.  Runs only in the build
container (needs /root/reference); the fixtures travel to the GPU box.

Each fixture holds the input tuple, the float32 outputs of the traced graph
(`peaks`), the same graph evaluated in float64 (`peaks_f64`) and, for the small
cases, per-block intermediates so every kernel has its own parity target.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nmrgnn_b200 import workloads  # noqa: E402
from nmrgnn_b200.graph import batch_graphs, build_graph, inv_degree_from_nlist, read_pdb  # noqa: E402
from oracle.savedmodel_interp import SavedModelInterpreter, reference_dir  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
REF_TESTS = os.path.join(os.environ.get("NMRGNN_REFERENCE", "/root/reference"), "tests")

TRACE = {
    "edge_features": "gnn-model/mul_1",
    "embed": "gnn-model/dense_9/MatMul",
    "mp_nodes_0": "gnn-model/mp-block/add",
    "mp_nodes_1": "gnn-model/mp-block/add_1",
    "mp_nodes_2": "gnn-model/mp-block/add_2",
    "mp_nodes_3": "gnn-model/mp-block/add_3",
    "fc_nodes": "gnn-model/fc-block/dense_7/Softplus",
}


def run(it32, it64, graph, with_trace):
    it32.keep_trace = with_trace
    it32.trace = {}
    peaks = it32(*graph)
    extra = {k: it32.trace[v].astype(np.float32) for k, v in TRACE.items()} if with_trace else {}
    it32.keep_trace = False
    peaks64 = it64(*graph)
    return peaks.astype(np.float32), peaks64.astype(np.float64), extra


def save(name, graph, peaks, peaks64, extra=None, offsets=None, note=""):
    atoms, nlist, edges, inv = graph
    d = dict(atoms=np.asarray(atoms, np.float32), nlist=np.asarray(nlist, np.int32),
             edges=np.asarray(edges, np.float32), inv_degree=np.asarray(inv, np.float32),
             peaks=peaks, peaks_f64=peaks64, note=np.frombuffer(note.encode(), np.uint8))
    if offsets is not None:
        d["graph_offsets"] = np.asarray(offsets, np.int64)
    d.update(extra or {})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: N={atoms.shape[0]} K={nlist.shape[1]} -> {os.path.getsize(path)} bytes")


def per_graph(it32, it64, batch):
    atoms, nlist, edges, inv, offs = batch
    p32 = np.zeros(atoms.shape[0], np.float32)
    p64 = np.zeros(atoms.shape[0], np.float64)
    for g in range(len(offs) - 1):
        a, b = int(offs[g]), int(offs[g + 1])
        gr = (atoms[a:b], nlist[a:b] - a, edges[a:b], inv[a:b])
        p32[a:b], p64[a:b], _ = run(it32, it64, gr, False)
    return p32, p64


def main():
    os.makedirs(OUT, exist_ok=True)
    it32 = SavedModelInterpreter(reference_dir(), np.float32)
    it64 = SavedModelInterpreter(reference_dir(), np.float64)

    # 1. the reference's integration fixture: myoglobin 108M.pdb, kNN-16 (config 1)
    u = read_pdb(os.path.join(REF_TESTS, "108M.pdb"))
    g = build_graph(u.atoms.positions, u.atoms.elements, 16, 10)
    p, p64, _ = run(it32, it64, g, False)
    save("g108m", g, p, p64, note="tests/108M.pdb, 2482 atoms, kNN-16, nm; traced SavedModel graph in NumPy")
    np.savez_compressed(os.path.join(OUT, "g108m_structure.npz"), positions_A=u.atoms.positions,
                        elements=u.atoms.elements.astype("U2"), names=u.atoms.names.astype("U4"),
                        resnames=u.atoms.resnames.astype("U3"), resids=u.atoms.resids)

    # 1b. the reference's trajectory fixture: KRAS NMR ensemble 7lgi.pdb.gz, 10 MODELs x 2 770 atoms
    #     (tests/test_nmrgnn.py:245-257).  Coordinates are stored as milli-Angstrom integers (PDB precision);
    #     golden peaks for the first and the last model on the host-built kNN-16 graph.
    u = read_pdb(os.path.join(REF_TESTS, "7lgi.pdb.gz"))
    frames = np.stack([(u.trajectory[i], u.atoms.positions.copy())[1] for i in range(len(u.trajectory))])
    first = build_graph(frames[0], u.atoms.elements, 16, 10)
    last = build_graph(frames[-1], u.atoms.elements, 16, 10)
    p0, p0_64, _ = run(it32, it64, first, False)
    p9, p9_64, _ = run(it32, it64, last, False)
    np.savez_compressed(os.path.join(OUT, "g7lgi_structure.npz"), positions_mA=np.round(frames.astype(np.float64) * 1000.0).astype(np.int32),
                        elements=u.atoms.elements.astype("U2"), names=u.atoms.names.astype("U4"),
                        resnames=u.atoms.resnames.astype("U3"), resids=u.atoms.resids,
                        peaks_first=p0, peaks_first_f64=p0_64, peaks_last=p9, peaks_last_f64=p9_64)
    print("g7lgi_structure:", frames.shape, "mean((last - first)^2) =", float(np.mean((p9_64 - p0_64) ** 2)))

    # 2. the reference's unit-test ring graph (tests/test_nmrgnn.py:20-31), C=10, K=2
    for tag, d in (("ring5_unit", 1.0), ("ring5_bonded", 0.15)):
        a, nl, e, inv = workloads.ring_graph(5, 10, 2)
        g = (a, nl, e * np.float32(d), inv)
        p, p64, ex = run(it32, it64, g, True)
        save(tag, g, p, p64, ex, note=f"5-node ring, 2 neighbours, edges={d}")

    # 3. small protein-like graph with padded slots + all intermediates
    g = workloads.protein_graph(seed=11, n_lo=300, n_hi=300)
    p, p64, ex = run(it32, it64, g, True)
    save("prot300", g, p, p64, ex, note="synthetic chain, 300 atoms, K=16, ~1% padded slots")

    # 4. edge cases: an isolated atom (all slots padded, inv_degree 0), a genuine
    #    neighbour with index 0 (not counted by inv_degree), non-one-hot atoms row
    a, nl, e, inv = workloads.protein_graph(seed=12, n_lo=64, n_hi=64, pad_fraction=0.2)
    nl[5] = 0
    e[5] = 0
    nl[6, :3] = 0
    e[6, :3] = np.float32([0.11, 0.15, 0.21])
    a = a.copy()
    a[7] = 0
    a[7, 3] = 0.25
    a[7, 4] = 0.75
    inv = inv_degree_from_nlist(nl)
    g = (a, nl, e, inv)
    p, p64, ex = run(it32, it64, g, True)
    save("edge_cases64", g, p, p64, ex, note="isolated atom 5, index-0 neighbours at atom 6, soft one-hot atom 7")

    # 5. batch of small molecules with K=8 (config 3 shape), evaluated graph by graph
    b = workloads.small_molecule_batch(12, first_seed=100)
    p, p64 = per_graph(it32, it64, b)
    save("smallmol12_k8", b[:4], p, p64, offsets=b[4], note="12 small molecules, K=8, ~10% padded; per-graph")

    # 6. batch of 3 protein-like graphs (config 2 shape, reduced), evaluated graph by graph
    b = batch_graphs([workloads.protein_graph(seed=s, n_lo=500, n_hi=700) for s in (21, 22, 23)])
    p, p64 = per_graph(it32, it64, b)
    save("prot3_batch", b[:4], p, p64, offsets=b[4], note="3 synthetic chains 500-700 atoms, K=16; per-graph")


if __name__ == "__main__":
    main()
