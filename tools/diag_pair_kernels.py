"""GPU diagnostic: CTA-pair (cta_group::2) forms of the MP-layer and node-MLP kernels against the one-CTA forms:
bit-equality of the peaks and stage times (CUDA events, same stream) of the forward."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nmrgnn_b200
from nmrgnn_b200 import workloads, _capi

m = nmrgnn_b200.load_model()
h = m.handle
h.set_option("tc_min_atoms", 0)
dev = torch.device("cuda", 0)
s = int(torch.cuda.current_stream().cuda_stream) or 1
for n_graphs in (1, 3, 64):          # odd and even tile counts, then BASELINE configs[1]
    atoms, nlist, edges, inv, offs = workloads.protein_batch(n_graphs, first_seed=0)
    n = atoms.shape[0]
    d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)]
    res = {}
    for name, mp_pair, fc_pair in (("one-CTA", 0, 0), ("mp_pair", 1, 0), ("fc_pair", 0, 1)):
        h.set_option("mp_pair", mp_pair)
        h.set_option("fc_pair", fc_pair)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        for _ in range(3):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
        h.synchronize(s)
        h.set_option("profile", 1)
        ts = []
        for _ in range(7):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
            h.synchronize(s)
            st = h.stage_times()
            ts.append([st["edge"], st["embed"], float(np.mean(st["mp_layers"])), st["fc_readout"]])
        h.set_option("profile", 0)
        res[name] = (out.cpu().numpy(), np.median(np.array(ts), axis=0))
    base = res["one-CTA"][0]
    line = f"n_atoms {n} tiles {(n + 127) // 128}:"
    for name in res:
        eq = np.array_equal(res[name][0], base)
        line += f" [{name}: equal {eq} maxdiff {np.abs(res[name][0] - base).max():.2e} edge/embed/mp/fc ms {np.round(res[name][1], 4).tolist()}]"
    print(line)
