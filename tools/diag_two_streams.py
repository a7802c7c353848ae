"""GPU experiment: throughput of independent forwards alternating between TWO handles on two streams (the tail of every
persistent kernel -- the last, partly filled wave of tiles -- and the launch gaps of one forward fill with CTAs of the
other) against one handle on one stream.  Device-resident inputs, BASELINE config 2."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads, _capi  # noqa: E402

atoms, nlist, edges, inv, offs = workloads.protein_batch(64, first_seed=0)
n, k = atoms.shape[0], nlist.shape[1]
dev = torch.device("cuda", 0)
models = [nmrgnn_b200.load_model(), nmrgnn_b200.load_model()]
streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
d_in = [[torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)] for _ in range(2)]
outs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]


def run(n_handles, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:n_handles]:
        s.wait_event(e0)
    for i in range(steps):
        j = i % n_handles
        sp = int(streams[j].cuda_stream)
        models[j].handle.forward(d_in[j][0], d_in[j][1], d_in[j][2], d_in[j][3], n, k, outs[j], _capi.MEM_DEVICE, sp)
    evs = []
    for s in streams[:n_handles]:
        ev = torch.cuda.Event()
        ev.record(s)
        evs.append(ev)
    cur = torch.cuda.current_stream()
    for ev in evs:
        cur.wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for nh in (1, 2, 1, 2):
    run(nh, 6)
    ms = run(nh, 40)
    print(f"{nh} handle(s) / stream(s): {ms:.4f} ms per forward, {n / ms / 1e3:.1f} M atoms/s", flush=True)
assert torch.equal(outs[0], outs[1])
