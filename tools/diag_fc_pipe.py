"""GPU diagnostic: the layer-pipelined node-MLP kernel (option fc_pipe) against the round-1 kernel on BASELINE config 2:
error against the committed fp64 golden peaks for several compensation slopes, stage times, role counters."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads, _capi  # noqa: E402


def report(name, y, ref):
    tol = 1e-4 * np.abs(ref) + 1e-4
    e = np.abs(y - ref) / tol
    print(f"  {name:28s}: tol_ratio max {e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} p99.9 {np.quantile(e, 0.999):.3f} "
          f"rms {np.sqrt(np.mean(e * e)):.4f} | > 1: {int((e > 1).sum())} | mean signed {np.mean((y - ref) / tol):+.4f}", flush=True)


def main():
    m = nmrgnn_b200.load_model()
    h = m.handle
    h.set_option("tc_min_atoms", 0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "full_config2.npz"))
    atoms, nlist, edges, inv, offs = workloads.protein_batch(64, first_seed=0)
    ref = z["peaks_f64"]
    g = (atoms, nlist, edges, inv)
    h.set_option("fc_pipe", 0)
    y0 = m(g).astype(np.float64)
    report("round-1 kernel", y0, ref)
    h.set_option("fc_pipe", 1)
    for c in [int(a) for a in sys.argv[1:]] or [50]:
        h.set_option("fc_pos_comp1_x100", c)
        y = m(g).astype(np.float64)
        report(f"pipelined c'={c / 100:.2f}", y, ref)
        report(f"  ... vs round-1 kernel", y, y0)
    h.set_option("fc_pos_comp1_x100", 50)
    y1 = m(g)
    h.set_option("fc_pair", 1)
    yp = m(g)
    h.set_option("fc_pair", 0)
    report("pipelined, CTA pairs", yp.astype(np.float64), ref)
    print("  pair kernel bit-identical to the one-CTA kernel:", bool(np.array_equal(y1, yp)),
          "max abs diff", float(np.abs(y1 - yp).max()), flush=True)
    # timing: device-resident forward, stage times
    dev = torch.device("cuda", 0)
    n = atoms.shape[0]
    d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in g]
    out = torch.empty(n, dtype=torch.float32, device=dev)
    s = int(torch.cuda.current_stream().cuda_stream) or 1
    for pipe, pair in ((0, 0), (1, 0), (1, 1)):
        h.set_option("fc_pipe", pipe)
        h.set_option("fc_pair", pair)
        h.set_option("profile", 1)
        ts = []
        for it in range(8):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
            h.synchronize(s)
            st = h.stage_times()
            ts.append([st["edge"], st["embed"]] + list(st["mp_layers"]) + [st["fc_readout"]])
        t = np.median(np.array(ts[3:]), axis=0)
        print(f"fc_pipe={pipe} fc_pair={pair}: stage ms {np.round(t, 4).tolist()} total {t.sum():.4f}", flush=True)
        h.set_option("profile", 0)
    for pair in (0, 1):
        h.set_option("fc_pair", pair)
        h.set_option("fc_role_counters", 1)
        h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
        h.synchronize(s)
        print(f"fc_pair={pair}:", end=" ", flush=True)
        h.set_option("fc_role_counters", 2)
        h.set_option("fc_role_counters", 0)
    h.set_option("fc_pair", 0)


if __name__ == "__main__":
    main()
