"""GPU diagnostic: pipelined node-MLP kernel vs round-1 kernel through nmrgnn_fc_readout on one golden fixture:
where (row in tile, class, column) do they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmrgnn_b200
from nmrgnn_b200 import _capi
from conftest import load_golden

m = nmrgnn_b200.load_model()
h = m.handle
h.set_option("tc_min_atoms", 0)
g = load_golden(sys.argv[1] if len(sys.argv) > 1 else "prot300")
nodes = np.ascontiguousarray(g["mp_nodes_3"], np.float32)
atoms = np.ascontiguousarray(g["atoms"], np.float32)
n = nodes.shape[0]
res = {}
for pipe in (0, 1):
    h.set_option("fc_pipe", pipe)
    peaks = np.zeros(n, np.float32); fcn = np.zeros((n, 128), np.float32)
    h.fc_readout(nodes, atoms, n, peaks, fcn, _capi.MEM_HOST)
    res[pipe] = (peaks.copy(), fcn.copy())
ref_fc = g["fc_nodes"]; ref_pk = g["peaks_f64"]
for pipe in (0, 1):
    pk, fcn = res[pipe]
    sc = np.abs(ref_fc).max()
    e = np.abs(fcn - ref_fc) / sc
    print(f"pipe={pipe}: fc_nodes max err/scale {e.max():.3e} rms {np.sqrt((e**2).mean()):.3e}; peaks max abs err {np.abs(pk - ref_pk).max():.4e}")
    bad_rows = np.where(e.max(1) > 1e-4)[0]
    print("   bad rows:", len(bad_rows), bad_rows[:40])
    bad_cols = np.where(e.max(0) > 1e-4)[0]
    print("   bad cols:", len(bad_cols), bad_cols[:40])
    cls = atoms.argmax(1)
    pe = np.abs(pk - ref_pk) / (1e-4 * np.abs(ref_pk) + 1e-4)
    for c in np.unique(cls):
        print(f"   class {c}: n={int((cls == c).sum())} peak tol_ratio max {pe[cls == c].max():.3f}")
