"""GPU diagnostic: both compute paths against the committed full-size golden peaks of BASELINE configs[1] / configs[2]
(tests/golden/full_config{2,3}.npz: the reference's traced graph executed per graph in fp32 and fp64).  Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402
from make_golden_full import graph_digest  # noqa: E402


def report(name, y, ref):
    tol = 1e-4 * np.abs(ref) + 1e-4
    e = np.abs(y - ref) / tol
    w = np.argsort(-e)[:3]
    print(f"  {name:14s}: tol_ratio max {e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} p99.9 {np.quantile(e, 0.999):.3f} "
          f"p99 {np.quantile(e, 0.99):.3f} median {np.median(e):.4f} | > 1: {int((e > 1).sum())} | mean signed "
          f"{np.mean((y - ref) / tol):+.4f} | worst " + " ".join(f"{i}:{e[i]:.2f}({ref[i]:+.2f})" for i in w))


def main():
    m = nmrgnn_b200.load_model()
    print("edge table:", m.handle.edge_table_info(), "| default path:", m.handle.compute_path)
    for cfg, gen in (("full_config2", lambda: workloads.protein_batch(64, first_seed=0)),
                     ("full_config3", lambda: workloads.small_molecule_batch(1024, first_seed=0))):
        z = np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz"))
        atoms, nlist, edges, inv, offs = gen()
        assert np.array_equal(offs, z["graph_offsets"]), "graph sizes differ from the fixture"
        bad = 0
        for g in range(len(offs) - 1):
            a, b = int(offs[g]), int(offs[g + 1])
            bad += graph_digest(atoms[a:b], nlist[a:b] - a, edges[a:b], inv[a:b]) != str(z["digests"][g])
        print(f"{cfg}: {atoms.shape[0]} atoms, K={nlist.shape[1]}, input digests differing: {bad} of {len(offs) - 1}")
        ref, ref32 = z["peaks_f64"], z["peaks"].astype(np.float64)
        report("traced fp32", ref32, ref)
        for path in ("tc", "tc-nocomp", "ffma"):
            m.handle.set_option("tc_min_atoms", 0)
            m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
            m.handle.set_option("tc_compensate", 0 if path == "tc-nocomp" else 1)
            y = m((atoms, nlist, edges, inv)).astype(np.float64)
            report(path + " vs fp64", y, ref)
            report(path + " vs fp32", y, ref32)
        for extra in sys.argv[1:]:
            k, _, v = extra.partition("=")
            m.handle.set_option("force_ffma", 0)
            m.handle.set_option("tc_compensate", 1)
            m.handle.set_option(k, int(v))
            y = m((atoms, nlist, edges, inv)).astype(np.float64)
            report(f"{extra} vs fp64", y, ref)
            m.handle.set_option(k, {"edge_table": 1}.get(k, 0))      # back to the default


if __name__ == "__main__":
    main()
