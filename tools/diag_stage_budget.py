"""GPU diagnostic: where does the peak error come from?  One stage at a time is computed by the CUDA block
(through the per-block C ABI, fed with the fp64 oracle's inputs rounded to fp32) and substituted into the fp64
oracle pipeline; the resulting peak error (units of the tolerance 1e-4 |ref| + 1e-4 ppm) is that stage's
contribution.  Usage: diag_stage_budget.py [seed ...]   (config-2 protein graphs).  Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402
from oracle import forward as orc  # noqa: E402


def tail64(p, h, atoms, l0):
    """fp64 continuation from the node features after MP layer l0-1 (h) to the peaks"""
    return h


def main():
    seeds = [int(s) for s in sys.argv[1:] if s.lstrip("-").isdigit()] or [10, 8]
    m = nmrgnn_b200.load_model()
    m.handle.set_option("tc_min_atoms", 0)
    p = m.params.astype(np.float64)
    for seed in seeds:
        atoms, nlist, edges, inv = workloads.protein_graph(seed)
        nl = nlist.astype(np.int64)
        a64, e64, i64 = atoms.astype(np.float64), edges.astype(np.float64), inv.astype(np.float64)
        mask = orc.edge_mask(e64)
        e3 = orc.edge_fc_block(orc.rbf_expansion(e64, p.rbf_low, p.rbf_high, p.rbf_count) * mask, p.edge_fc,
                               p.fc_activation) * mask
        hs = [a64 @ p.embed]
        for l in range(len(p.mp_w)):
            hs.append(orc.mp_layer(hs[-1], nl, e3, i64, p.mp_w[l], p.mp_activation) + hs[-1])

        def finish(h, l0, e3_=e3):
            for l in range(l0, len(p.mp_w)):
                h = orc.mp_layer(h, nl, e3_, i64, p.mp_w[l], p.mp_activation) + h
            z = orc.fc_block(h, p.fc, p.fc_activation)
            return orc.readout(z, a64, p.out, p.peak_std, p.peak_avg)

        ref = finish(hs[0], 0)
        tol = 1e-4 * np.abs(ref) + 1e-4
        print(f"seed {seed}: {atoms.shape[0]} atoms")
        for path in ("tc", "ffma"):
            m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
            rows = []
            ef = m.edge_fc_block(edges).astype(np.float64)
            rows.append(("edge", finish(hs[0], 0, ef)))
            rows.append(("embed", finish(m.embed_layer(atoms).astype(np.float64), 0)))
            for l in range(len(p.mp_w)):
                out = m.mp_block.mp[l]([hs[l].astype(np.float32), nlist, e3.astype(np.float32), inv])
                # the rounding of the inputs to fp32 is part of what an fp32 pipeline does
                rows.append((f"mp{l}", finish(out.astype(np.float64), l + 1)))
            rows.append(("fc+readout", m.readout(hs[-1].astype(np.float32), atoms).astype(np.float64)))
            rows.append(("whole", m((atoms, nlist, edges, inv)).astype(np.float64)))
            for name, y in rows:
                e = np.abs(y - ref) / tol
                w = int(np.argmax(e))
                print(f"  {path:4s} {name:10s}: max {e.max():.3f} (atom {w}, ref {ref[w]:+.3f}) p99.9 {np.quantile(e, 0.999):.3f} "
                      f"p99 {np.quantile(e, 0.99):.3f} rms {np.sqrt(np.mean(e * e)):.4f} mean signed {np.mean((y - ref) / tol):+.4f}")
        # input-rounding floor: fp64 pipeline fed with fp32-rounded stage inputs
        for l in range(len(p.mp_w)):
            y = finish(orc.mp_layer(hs[l].astype(np.float32).astype(np.float64), nl, e3.astype(np.float32).astype(np.float64),
                                    i64, p.mp_w[l], p.mp_activation) + hs[l].astype(np.float32).astype(np.float64), l + 1)
            e = np.abs(y - ref) / tol
            print(f"  floor (fp32-rounded inputs, exact arithmetic) mp{l}: max {e.max():.3f} p99.9 {np.quantile(e, 0.999):.3f}")
    m.handle.set_option("force_ffma", 0)


if __name__ == "__main__":
    main()
