"""Frame-stream engine: per-frame inference on a trajectory (BASELINE configs[4], the loop of
``nmrgnn eval-struct``, nmrgnn/main.py:236-275) with the graph build on the GPU.

One frame of a 2 500-atom protein is 20 row tiles — a seventh of a B200 — and eight kernel launches, so frames are
processed B at a time: frames are independent graphs (`batch_graphs` semantics: neighbour indices are batch-global),
B is chosen so that a batch fills the 148 SMs, and the whole batch step (kNN graph build + forward: 7 kernels) is
captured once in a CUDA graph and replayed.  Positions go up and peaks come back through double-buffered pinned
staging on a copy stream, so the copies of batch i+1 / i-1 overlap the kernels of batch i.

    fs = FrameStream(model, elements, n_atoms, neighbor_number=16)
    peaks = fs.run(frames_nm)            # [n_frames, n_atoms, 3] float32, nm  ->  [n_frames, n_atoms] float32

Several GPUs: frames are independent, so rank r of W takes batches r, r + W, ... (`rank`, `world` arguments);
there is no exchange on the data path.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Sequence

import numpy as np

from . import _capi
from .graph import one_hot_elements

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def default_frames_per_batch(n_atoms: int, num_sms: int = 148) -> int:
    """Frames per launch.  The MP and node-MLP kernels are persistent over 128-atom tiles, one CTA per SM, so a batch
    costs ceil(tiles / num_sms) tile-times: 8 frames of 108M.pdb are 155 tiles = 2 waves at 52 % utilisation, 15 frames
    are 291 tiles = 2 waves at 98 %.  Returns the smallest batch (<= 64 frames) that fills its waves to >= 95 %, or the
    best-filled one if none does."""
    best, best_u = 1, 0.0
    for b in range(1, 65):
        tiles = (b * int(n_atoms) + 127) // 128
        u = tiles / (-(-tiles // num_sms) * num_sms)
        if u >= 0.95:
            return b
        if u > best_u + 1e-9:
            best, best_u = b, u
    return best


class FrameStream:
    def __init__(self, model, elements: Sequence[str], n_atoms: Optional[int] = None, neighbor_number: int = 16,
                 frames_per_batch: Optional[int] = None, use_cuda_graph: bool = True, cutoff_nm: float = 0.0):
        if torch is None:
            raise ImportError("FrameStream needs torch for pinned staging buffers and streams")
        self.model = model
        self.k = int(neighbor_number)
        atoms1 = one_hot_elements(elements, model.params.num_elem)
        self.n = int(atoms1.shape[0] if n_atoms is None else n_atoms)
        if atoms1.shape[0] != self.n:
            raise ValueError("elements must have one entry per atom")
        self.B = int(frames_per_batch or default_frames_per_batch(self.n))
        self.cutoff = float(cutoff_nm)
        self.dev = torch.device("cuda", model.device)
        B, n, k = self.B, self.n, self.k
        self.atoms1 = atoms1
        with torch.cuda.device(self.dev):
            self.d_atoms = torch.from_numpy(np.tile(atoms1, (B, 1))).to(self.dev)
            self.d_pos = [torch.zeros((B * n, 3), dtype=torch.float32, device=self.dev) for _ in range(2)]
            self.d_nlist = torch.zeros((B * n, k), dtype=torch.int32, device=self.dev)
            self.d_edges = torch.zeros((B * n, k), dtype=torch.float32, device=self.dev)
            self.d_inv = torch.zeros(B * n, dtype=torch.float32, device=self.dev)
            self.d_peaks = [torch.zeros(B * n, dtype=torch.float32, device=self.dev) for _ in range(2)]
            self.h_pos = [torch.zeros((B * n, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
            self.h_peaks = [torch.zeros(B * n, dtype=torch.float32).pin_memory() for _ in range(2)]
            self.compute = torch.cuda.Stream(device=self.dev)
            self.copy = torch.cuda.Stream(device=self.dev)      # uploads
            self.down = torch.cuda.Stream(device=self.dev)      # downloads (separate: a download waits for its batch's kernels
            #                                                   # and must not hold back the next batch's upload behind it)
        self.offsets = np.arange(B + 1, dtype=np.int64) * n
        self.graphs = [None, None]           # one captured graph per (positions, peaks) buffer pair
        self.graph_captured = False
        self._want_graph = bool(use_cuda_graph)
        self._warm = False

    # -- one batch step on the compute stream: graph build + forward -------------------------------------------------
    def _launch(self, slot: int):
        h = self.model.handle
        s = int(self.compute.cuda_stream)
        h.knn_graph(self.d_pos[slot], self.offsets, self.B * self.n, self.B, self.k, self.cutoff, self.d_nlist, self.d_edges,
                    self.d_inv, _capi.MEM_DEVICE, s)
        h.forward(self.d_atoms, self.d_nlist, self.d_edges, self.d_inv, self.B * self.n, self.k, self.d_peaks[slot],
                  _capi.MEM_DEVICE, s)

    def _prepare(self):
        if self._warm:
            return
        with torch.cuda.stream(self.compute):
            for slot in (0, 1):              # sizes every workspace of the handle; uploads the graph offsets
                self._launch(slot)
        self.compute.synchronize()
        if self._want_graph:
            try:
                for slot in (0, 1):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.compute):
                        self._launch(slot)
                    self.graphs[slot] = g
                self.graph_captured = True
            except Exception:                # capture not possible on this runtime: plain launches
                self.graphs = [None, None]
                self.graph_captured = False
                torch.cuda.synchronize(self.dev)
        self._warm = True

    def _step(self, slot: int):
        if self.graphs[slot] is not None:
            self.graphs[slot].replay()       # replays on the capturing stream's context; ordered by the events below
        else:
            self._launch(slot)

    def run(self, frames_nm: np.ndarray, rank: int = 0, world: int = 1) -> Dict[str, object]:
        """frames_nm: [n_frames, n_atoms, 3] (nm).  Returns {"peaks": [n_mine, n_atoms], "frame_index": [...],
        "seconds": wall, "device_ms": compute-stream time, "batches": ...} for the frames of this rank."""
        frames_nm = np.asarray(frames_nm, np.float32)
        if frames_nm.ndim != 3 or frames_nm.shape[1:] != (self.n, 3):
            raise ValueError(f"frames must be [n_frames, {self.n}, 3]")
        B, n = self.B, self.n
        n_frames = frames_nm.shape[0]
        n_batches = -(-n_frames // B)
        mine = [b for b in range(n_batches) if b % world == rank]
        self._prepare()
        out = np.empty((sum(min(B, n_frames - b * B) for b in mine), n), np.float32)
        index = np.concatenate([np.arange(b * B, min(n_frames, (b + 1) * B)) for b in mine]) if mine else np.zeros(0, np.int64)
        up_done = [torch.cuda.Event() for _ in range(2)]
        k_done = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        down_done = [torch.cuda.Event() for _ in range(2)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def stage(i):                        # host: frames of batch i -> pinned buffer i & 1 (last batch padded by repetition)
            b = mine[i]
            f = frames_nm[b * B:(b + 1) * B]
            buf = self.h_pos[i & 1].numpy().reshape(B, n, 3)
            buf[:f.shape[0]] = f
            if f.shape[0] < B:
                buf[f.shape[0]:] = f[-1]

        def collect(i, row):                 # host: peaks of batch i out of pinned buffer i & 1
            down_done[i & 1].synchronize()
            cnt = min(B, n_frames - mine[i] * B)
            out[row:row + cnt] = self.h_peaks[i & 1].numpy().reshape(B, n)[:cnt]
            return row + cnt

        t0 = time.perf_counter()
        ev0.record(self.compute)
        row = 0
        for i in range(len(mine)):
            if i >= 2:
                row = collect(i - 2, row)    # frees pinned buffers i & 1
            stage(i)
            with torch.cuda.stream(self.copy):
                if i >= 2:
                    self.copy.wait_event(k_done[i & 1])     # batch i-2 (same buffers) has finished with d_pos[i & 1]
                self.d_pos[i & 1].copy_(self.h_pos[i & 1], non_blocking=True)
                up_done[i & 1].record(self.copy)
            self.compute.wait_event(up_done[i & 1])
            if i >= 2:
                self.compute.wait_event(down_done[i & 1])   # d_peaks[i & 1] has been copied out
            with torch.cuda.stream(self.compute):
                self._step(i & 1)
                k_done[i & 1].record(self.compute)
            with torch.cuda.stream(self.down):
                self.down.wait_event(k_done[i & 1])
                self.h_peaks[i & 1].copy_(self.d_peaks[i & 1], non_blocking=True)
                down_done[i & 1].record(self.down)
        ev1.record(self.compute)
        for i in range(max(0, len(mine) - 2), len(mine)):
            row = collect(i, row)
        self.compute.synchronize()
        self.model.synchronize()             # deferred device-side errors (out-of-range index cannot happen here)
        wall = time.perf_counter() - t0
        return {"peaks": out, "frame_index": index, "seconds": wall, "device_ms": ev0.elapsed_time(ev1) if mine else 0.0,
                "batches": len(mine), "frames_per_batch": B, "cuda_graph": self.graph_captured}
