"""Pipelined evaluation of a sequence of graph batches that live in HOST memory.

The synchronous call ``model(host_arrays)`` (``nmrgnn_forward(NMRGNN_MEM_HOST)``) uploads 172 B per atom, computes and
downloads: on one B200 the PCIe copy of BASELINE config 2 (28 MB) costs a quarter of the step.  A caller that has more
than one batch -- `nmrgnn eval-tfrecords`-style loops over records (nmrgnn/main.py:125-127), the README's prediction
loop (README.md:97-105), the benchmark's `e2e` leg -- does not have to pay it serially: ``BatchStream`` keeps two input
slots on the device, uploads batch i + 1 on a copy stream while batch i computes, and brings the peaks of batch i - 1
down behind it (double-buffered pinned staging, CUDA events between the two streams; the same scheme `FrameStream` uses
for trajectory frames).  Every batch's inputs cross PCIe and every batch's peaks come back to the host; only the waiting
is overlapped.  Results are bit-identical to the synchronous call (same kernels on the same data).

Multi-GPU: with a ``PeerGather`` (nmrgnn_b200.sharding) each rank streams its shard of every batch and the peaks of all
ranks come back on every rank (``nmrgnn_forward_sharded`` on device buffers, exchange inside the call)."""
from __future__ import annotations

import time
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _capi

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class BatchStream:
    def __init__(self, model, max_atoms: int, neighbor_number: int = 16, peer=None):
        if torch is None:
            raise ImportError("BatchStream needs torch for pinned staging buffers and streams")
        self.model, self.peer = model, peer
        self.cap, self.k = int(max_atoms), int(neighbor_number)
        self.dev = torch.device("cuda", model.device)
        C = model.params.num_elem
        n, k = self.cap, self.k
        with torch.cuda.device(self.dev):
            mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=self.dev) for _ in range(2)]   # noqa: E731
            self.d_atoms, self.d_nlist = mk((n, C), torch.float32), mk((n, k), torch.int32)
            self.d_edges, self.d_inv = mk((n, k), torch.float32), mk((n,), torch.float32)
            self.d_peaks = mk((n,), torch.float32)
            self.out_len = n if peer is None else peer.world * peer.capacity
            self.h_peaks = [torch.empty(self.out_len, dtype=torch.float32).pin_memory() for _ in range(2)]
            self.compute = torch.cuda.Stream(device=self.dev)
            self.copy = torch.cuda.Stream(device=self.dev)        # uploads
            self.down = torch.cuda.Stream(device=self.dev)        # downloads: on their own stream, or the download of
            #                                                     # batch i (which waits for its kernels) would hold back
            #                                                     # the upload of batch i + 1 behind it

    @staticmethod
    def pin(batch: Sequence[np.ndarray]):
        """(atoms, nlist, edges, inv_degree) as pinned host tensors (the H2D copies are asynchronous only from pinned
        memory; pageable arrays are staged by the driver and serialise the pipeline)."""
        a, nl, e, inv = batch[:4]
        return (torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory(),
                torch.from_numpy(np.ascontiguousarray(nl, np.int32)).pin_memory(),
                torch.from_numpy(np.ascontiguousarray(e, np.float32)).pin_memory(),
                torch.from_numpy(np.ascontiguousarray(inv, np.float32).reshape(-1)).pin_memory())

    def run(self, batches: Iterable[Sequence], keep: bool = True) -> Dict[str, object]:
        """batches: iterable of (atoms, nlist, edges, inv_degree) host tensors / arrays (ideally `pin`ned).  Returns
        {"peaks": [np.ndarray per batch] (every rank's peaks concatenated rank-major with a PeerGather), "seconds": wall,
        "device_ms": compute-stream time}.  keep=False drops the peaks after the download (benchmarking)."""
        h = self.model.handle
        sp = int(self.compute.cuda_stream)
        up = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        down = [torch.cuda.Event() for _ in range(2)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out: List[Optional[np.ndarray]] = []
        sizes: List[int] = []
        pending: List[int] = []

        def collect(i):
            s = i & 1
            down[s].synchronize()
            m = sizes[i] if self.peer is None else self.out_len
            out[i] = self.h_peaks[s][:m].numpy().copy() if keep else None

        t0 = time.perf_counter()
        ev0.record(self.compute)
        for i, b in enumerate(batches):
            s = i & 1
            a, nl, e, inv = [x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x)) for x in b[:4]]
            n = int(a.shape[0])
            if n > self.cap or nl.shape[1] != self.k:
                raise ValueError(f"batch of {n} atoms / {nl.shape[1]} neighbours does not fit this stream ({self.cap}, {self.k})")
            sizes.append(n)
            out.append(None)
            if i >= 2:
                collect(i - 2)                           # slot s: its previous peaks have left the pinned buffer
                pending.remove(i - 2)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(done[s])            # the batch that used input slot s two steps ago has computed
                self.d_atoms[s][:n].copy_(a, non_blocking=True)
                self.d_nlist[s][:n].copy_(nl, non_blocking=True)
                self.d_edges[s][:n].copy_(e, non_blocking=True)
                self.d_inv[s][:n].copy_(inv.reshape(-1), non_blocking=True)
                up[s].record(self.copy)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(up[s])
                self.compute.wait_event(down[s])         # ... and its peaks buffer has been downloaded
                if self.peer is None:
                    h.forward(self.d_atoms[s], self.d_nlist[s], self.d_edges[s], self.d_inv[s], n, self.k, self.d_peaks[s],
                              _capi.MEM_DEVICE, sp)
                    src = self.d_peaks[s][:n]
                else:
                    h.forward_sharded(self.d_atoms[s], self.d_nlist[s], self.d_edges[s], self.d_inv[s], n, self.k, None,
                                      _capi.MEM_DEVICE, sp)
                    src = self.peer.view().reshape(-1)
                done[s].record(self.compute)
            with torch.cuda.stream(self.down):
                self.down.wait_event(done[s])
                self.h_peaks[s][:src.numel()].copy_(src, non_blocking=True)
                down[s].record(self.down)
            pending.append(i)
        ev1.record(self.compute)
        for i in list(pending):
            collect(i)
        self.compute.synchronize()
        h.synchronize(sp)                                # device-side errors (bad neighbour index) surface here
        return {"peaks": out, "seconds": time.perf_counter() - t0, "device_ms": ev0.elapsed_time(ev1)}
