import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import nmrgnn_b200
from nmrgnn_b200 import _capi
from conftest import load_golden
g = load_golden("g108m")
m = nmrgnn_b200.load_model(); h = m.handle
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in ("atoms", "nlist", "edges", "inv_degree")]
n = d[0].shape[0]; out = torch.empty(n, dtype=torch.float32, device=dev)
s = int(torch.cuda.current_stream().cuda_stream) or 1
res = {}
for small in (1, 0):
    h.set_option("mp_small_tiles", small)
    for _ in range(5): h.forward(d[0], d[1], d[2], d[3], n, 16, out, _capi.MEM_DEVICE, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): h.forward(d[0], d[1], d[2], d[3], n, 16, out, _capi.MEM_DEVICE, s)
    e1.record(); torch.cuda.synchronize()
    res[small] = out.cpu().numpy().copy()
    print(f"mp_small_tiles={small}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per 108M forward ({n} atoms)")
print("bit-identical:", np.array_equal(res[0], res[1]))
