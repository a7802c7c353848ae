"""GPU diagnostic: MD-stream throughput (BASELINE config 5 workload: 108M.pdb x 512 jittered frames) as a function of the
frames per batch; FrameStream's default is the smallest batch that fills its waves of 128-atom tiles to >= 95 %."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200.mdstream import FrameStream, default_frames_per_batch  # noqa: E402

with np.load(os.path.join(ROOT, "tests", "golden", "g108m_structure.npz")) as z:
    pos = z["positions_A"].astype(np.float32) / np.float32(10)
    elements = [str(e) for e in z["elements"]]
rng = np.random.default_rng(0)
frames = pos[None] + rng.normal(scale=0.02, size=(512,) + pos.shape).astype(np.float32)
m = nmrgnn_b200.load_model()
print("default frames per batch:", default_frames_per_batch(pos.shape[0]))
for B in [int(a) for a in sys.argv[1:]] or [15, 30, 45, 60]:
    fs = FrameStream(m, elements, pos.shape[0], 16, frames_per_batch=B)
    for _ in range(3):
        fs.run(frames[:4 * B], 0, 1)
    rs = [fs.run(frames, 0, 1) for _ in range(5)]
    w = float(np.median([r["seconds"] for r in rs]))
    d = float(np.median([r["device_ms"] for r in rs]))
    tiles = (B * pos.shape[0] + 127) // 128
    print(f"B = {B:3d}: {tiles:4d} tiles = {tiles / 148:.2f} waves | {512 / w:8.0f} frames/s end to end, device {d / 512 * 1e3:.1f} us per frame",
          flush=True)
