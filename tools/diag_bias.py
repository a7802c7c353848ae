"""GPU diagnostic (run under gpurun): measures the systematic (multiplicative) error of the
tcgen05 accumulation on real layer data.  A copy of the pretrained model with LINEAR
activations makes the raw contraction observable through the public per-block API:
    mp_layer:  h_out = inv_degree * D + h_in        ->  D = (h_out - h_in) / inv_degree
    edge MLP:  e3 = ((x W0 + b0) W1 + b1) ... (chained linear maps)
The regression slope of (D_gpu - D_fp64) on D_fp64 is the relative bias of the path; the
epilogue correction constants in kernels_tc.cuh are derived from this measurement.
Prints only; asserts nothing."""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nmrgnn_b200  # noqa: E402
from conftest import load_golden  # noqa: E402
from nmrgnn_b200.params import GNNParams, baseline_path  # noqa: E402


def stats(name, d_gpu, d_ref):
    e = d_gpu.astype(np.float64) - d_ref
    sc = np.abs(d_ref).max()
    slope = (e * d_ref).sum() / (d_ref * d_ref).sum()
    rms = np.sqrt((e ** 2).mean()) / sc
    rms_c = np.sqrt(((e - slope * d_ref) ** 2).mean()) / sc
    pos, neg = d_ref > 0, d_ref < 0
    bp = e[pos].sum() / np.abs(d_ref[pos]).sum() / 2 ** -24
    bn = e[neg].sum() / np.abs(d_ref[neg]).sum() / 2 ** -24
    print(f"{name}: slope {slope:+.3e} = {slope / 2 ** -24:+.2f} x 2^-24 | rms/scale {rms:.2e} -> {rms_c:.2e} after "
          f"removing the slope | max/scale {np.abs(e).max() / sc:.2e} | sum(e)/sum|D| for D>0: {bp:+.2f}, D<0: {bn:+.2f} "
          f"(x 2^-24), frac D>0 {pos.mean():.2f}")


def main():
    p = GNNParams.load(baseline_path())
    plin = dataclasses.replace(p, mp_activation="linear", fc_activation="linear")
    m = nmrgnn_b200.GNNModel(plin)
    print("path:", m.handle.compute_path)
    print("tc compensation (x 2^-24):", m.handle.tc_compensation())
    rng = np.random.default_rng(0)
    A = rng.uniform(0.5, 1.5, size=(128, 64)).astype(np.float16).astype(np.float32)
    W = rng.uniform(0.5, 1.5, size=(64, 128)).astype(np.float16).astype(np.float32)
    for sa, sw in ((1, 1), (1, -1), (-1, -1)):
        ref = (sa * A).astype(np.float64) @ (sw * W).astype(np.float64)
        d = m.handle.selftest_gemm(sa * A, sw * W, 3).astype(np.float64)
        e = (d - ref) / np.spacing(np.abs(ref).astype(np.float32))
        print(f"selftest fp16x1 K=64 signs A{sa:+d} W{sw:+d}: err in ulp mean {e.mean():+.3f} min {e.min():+.2f} max {e.max():+.2f}")
    for gname in ("prot300", "edge_cases64"):
        g = load_golden(gname)
        inv = g["inv_degree"].astype(np.float64)
        ok = inv > 0
        for path in ("tc", "ffma"):
            m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
            for l in range(4):
                h = g["embed"] if l == 0 else g[f"mp_nodes_{l - 1}"]
                out = m.mp_block.mp[l]([h, g["nlist"], g["edge_features"], g["inv_degree"]])
                d_gpu = (out.astype(np.float64) - h.astype(np.float64))[ok] / inv[ok, None]
                T = np.einsum("ijn,ijl->iln", g["edge_features"].astype(np.float64), h[g["nlist"]].astype(np.float64))
                d_ref = np.einsum("iln,lmn->im", T, p.mp_w[l].astype(np.float64))[ok]
                stats(f"{gname} {path:4s} mp{l}", d_gpu, d_ref)
            # edge MLP with linear activations: RBF -> 4 chained linear layers
            e3 = m.edge_fc_block(g["edges"]).astype(np.float64)
            from oracle import forward as orc
            x = orc.rbf_expansion(g["edges"].astype(np.float64), p.rbf_low, p.rbf_high, p.rbf_count) * (g["edges"] > 0)[..., None]
            for W, b in p.edge_fc:
                x = x @ W.astype(np.float64) + b.astype(np.float64)
            x = x * (g["edges"] > 0)[..., None]
            stats(f"{gname} {path:4s} edge(linear chain)", e3, x)
    m.handle.set_option("force_ffma", 0)
    # natural (uncompensated) slopes per layer on the synthetic workloads, real-model activations as inputs
    from nmrgnn_b200 import workloads
    from oracle import forward as orc
    m.handle.set_option("tc_compensate", 0)
    m.handle.set_option("tc_min_atoms", 0)
    for wname, b in (("protein_batch", workloads.protein_batch(2, first_seed=7)),
                     ("small_molecules", workloads.small_molecule_batch(96, first_seed=3))):
        atoms, nlist, edges, inv, offs = b
        inter = {}
        orc.forward(p, atoms, nlist, edges, inv, dtype=np.float64, intermediates=inter)
        e3 = inter["edge_features"].astype(np.float32)
        hs = [inter["embed"].astype(np.float32)] + [h.astype(np.float32) for h in inter["mp_nodes"][:-1]]
        ok = inv > 0
        for l in range(4):
            out = m.mp_block.mp[l]([hs[l], nlist, e3, inv])
            d_gpu = (out.astype(np.float64) - hs[l].astype(np.float64))[ok] / inv.astype(np.float64)[ok, None]
            T = np.einsum("ijn,ijl->iln", e3.astype(np.float64), hs[l][nlist].astype(np.float64))
            d_ref = np.einsum("iln,lmn->im", T, p.mp_w[l].astype(np.float64))[ok]
            stats(f"{wname} (uncompensated) mp{l}", d_gpu, d_ref)
    m.handle.set_option("tc_compensate", 1)


if __name__ == "__main__":
    main()
