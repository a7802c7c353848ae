"""GPU diagnostic: stage times of a single 108M-sized forward (2 482 atoms) and of the GPU graph build."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nmrgnn_b200
from nmrgnn_b200 import _capi
from nmrgnn_b200.graph import one_hot_elements

with np.load(os.path.join(ROOT, "tests", "golden", "g108m_structure.npz")) as z:
    pos = z["positions_A"].astype(np.float32) / np.float32(10)
    elements = [str(e) for e in z["elements"]]
n, k = pos.shape[0], 16
m = nmrgnn_b200.load_model(); h = m.handle
dev = torch.device("cuda", 0)
s = int(torch.cuda.current_stream().cuda_stream) or 1
d_atoms = torch.from_numpy(one_hot_elements(elements, 10)).to(dev)
d_pos = torch.from_numpy(pos).to(dev)
d_nl = torch.empty((n, k), dtype=torch.int32, device=dev)
d_ed = torch.empty((n, k), dtype=torch.float32, device=dev)
d_inv = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(n, dtype=torch.float32, device=dev)
offs = np.array([0, n], np.int64)
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(5):
    h.knn_graph(d_pos, offs, n, 1, k, 0.0, d_nl, d_ed, d_inv, _capi.MEM_DEVICE, s)
    h.forward(d_atoms, d_nl, d_ed, d_inv, n, k, d_out, _capi.MEM_DEVICE, s)
torch.cuda.synchronize()
e0, e1, e2 = ev(), ev(), ev()
e0.record()
for _ in range(50):
    h.knn_graph(d_pos, offs, n, 1, k, 0.0, d_nl, d_ed, d_inv, _capi.MEM_DEVICE, s)
e1.record()
for _ in range(50):
    h.forward(d_atoms, d_nl, d_ed, d_inv, n, k, d_out, _capi.MEM_DEVICE, s)
e2.record(); torch.cuda.synchronize()
print("knn_graph ms/frame", e0.elapsed_time(e1) / 50, " forward ms/frame", e1.elapsed_time(e2) / 50)
h.set_option("profile", 1)
acc = []
for _ in range(10):
    h.forward(d_atoms, d_nl, d_ed, d_inv, n, k, d_out, _capi.MEM_DEVICE, s)
    st = h.stage_times(); acc.append([st["edge"], st["embed"]] + list(st["mp_layers"]) + [st["fc_readout"]])
print("stages ms (edge, embed, mp x4, fc):", np.mean(acc[2:], axis=0).round(4))
