"""GPU diagnostic: peak error of the tensor-core path on the full BASELINE configs 2 and 3 (vs the fp64 golden peaks) and
the MP-layer launch time, as a function of the MP layers' accumulation options:
  mp_chain_segments (1 / 2 / 4 / 8)  and  mp_pos_comp_x100 (slope of the position-dependent compensation, x 2^-24 / 100).
Usage: diag_chain_segments.py [nseg:pos[:fcpos[:edgepos]] ...]   e.g.  1:0 1:50 2:50:0.  Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402


def stats(y, ref):
    e = np.abs(y - ref) / (1e-4 * np.abs(ref) + 1e-4)
    return (f"max {e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} p99.9 {np.quantile(e, 0.999):.3f} "
            f"rms {np.sqrt(np.mean(e * e)):.4f} >1: {int((e > 1).sum())}")


def main():
    combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(1, 0), (1, 50), (2, 50), (4, 50)]
    m = nmrgnn_b200.load_model()
    h = m.handle
    b2 = workloads.protein_batch(64, first_seed=0)
    b3 = workloads.small_molecule_batch(1024, first_seed=0)
    r2 = np.load(os.path.join(ROOT, "tests", "golden", "full_config2.npz"))["peaks_f64"]
    r3 = np.load(os.path.join(ROOT, "tests", "golden", "full_config3.npz"))["peaks_f64"]
    for combo in combos:
        nseg, pos = combo[:2]
        fcpos = combo[2] if len(combo) > 2 else 50
        if len(combo) > 3:                         # 4th field: edge block by the tcgen05 edge MLP with this slope
            h.set_option("edge_table", 0)
            h.set_option("edge_pos_comp_x100", combo[3])
        else:
            h.set_option("edge_table", 1)
        h.set_option("mp_chain_segments", nseg)
        h.set_option("mp_pos_comp_x100", pos)
        h.set_option("fc_pos_comp_x100", fcpos)
        comp = [round(c, 1) for c in h.tc_compensation()["mp_layers"]]
        y2 = m(b2[:4]).astype(np.float64)
        y3 = m(b3[:4]).astype(np.float64)
        h.set_option("profile", 1)
        ts = []
        for _ in range(5):
            m(b2[:4])
            st = h.stage_times()
            ts.append(np.mean(st["mp_layers"]))
        h.set_option("profile", 0)
        print(f"nseg {nseg} pos {pos / 100:.2f} fc {fcpos / 100:.2f} edge {combo[3] / 100 if len(combo) > 3 else 'table'} comp {comp} | cfg2 {stats(y2, r2)} | cfg3 {stats(y3, r3)} | mp ms {np.median(ts):.4f}",
              flush=True)


if __name__ == "__main__":
    main()
