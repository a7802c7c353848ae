"""GPU diagnostic: error distribution of both compute paths against the fp64 oracle on a large seeded batch
(default: 16 config-2 protein graphs, ~40 k atoms; `--small` = 512 config-3 small molecules).  Prints only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402
from oracle import forward as orc  # noqa: E402


def main():
    small = "--small" in sys.argv
    if small:
        b = workloads.small_molecule_batch(512, first_seed=0)
    else:
        b = workloads.protein_batch(16, first_seed=100)
    atoms, nlist, edges, inv, offs = b
    m = nmrgnn_b200.load_model()
    t0 = time.time()
    ref = orc.forward(m.params, atoms, nlist, edges, inv, dtype=np.float64)
    ref32 = orc.forward(m.params, atoms, nlist, edges, inv, dtype=np.float32)
    print(f"{atoms.shape[0]} atoms, K={nlist.shape[1]}, oracle fp64+fp32 in {time.time() - t0:.1f} s")
    tol = 1e-4 * np.abs(ref) + 1e-4
    rows = [("oracle fp32", ref32)]
    for path in ("tc", "tc-nocomp", "ffma"):
        m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
        m.handle.set_option("tc_compensate", 0 if path == "tc-nocomp" else 1)
        rows.append((path, m((atoms, nlist, edges, inv))))
    for name, y in rows:
        e = np.abs(y - ref) / tol
        nz = ref != 0
        rel = np.abs(y - ref)[nz] / np.abs(ref[nz])
        print(f"{name:12s}: tol_ratio max {e.max():.3f} p99.9 {np.quantile(e, 0.999):.3f} p99 {np.quantile(e, 0.99):.3f} "
              f"median {np.median(e):.4f} | frac > 1: {np.mean(e > 1):.2e} | strict rel err max {rel.max():.2e} "
              f"p99.9 {np.quantile(rel, 0.999):.2e} | mean signed err/tol {np.mean((y - ref) / tol):+.4f}")
    # edge block alone, against the fp64 oracle
    p = m.params
    e64 = edges.astype(np.float64)
    x = orc.rbf_expansion(e64, p.rbf_low, p.rbf_high, p.rbf_count) * orc.edge_mask(e64)
    ef_ref = orc.edge_fc_block(x, [(W.astype(np.float64), b.astype(np.float64)) for W, b in p.edge_fc], p.fc_activation)
    ef_ref = ef_ref * orc.edge_mask(e64)
    for path in ("tc", "tc-nocomp", "ffma"):
        m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
        m.handle.set_option("tc_compensate", 0 if path == "tc-nocomp" else 1)
        ef = m.edge_fc_block(edges).astype(np.float64)
        e = ef - ef_ref
        sc = np.abs(ef_ref).max()
        slope = (e * ef_ref).sum() / (ef_ref * ef_ref).sum()
        print(f"edge block {path:10s}: max/scale {np.abs(e).max() / sc:.2e} rms/scale {np.sqrt((e ** 2).mean()) / sc:.2e} "
              f"slope {slope / 2 ** -24:+.2f} x 2^-24, per channel rms/scale "
              + " ".join(f"{np.sqrt((e[..., c] ** 2).mean()) / np.abs(ef_ref[..., c]).max():.2e}" for c in range(e.shape[-1])))
    m.handle.set_option("force_ffma", 0)
    m.handle.set_option("tc_compensate", 1)


if __name__ == "__main__":
    main()
