"""Graph-sharded evaluation over several GPUs (one process per GPU).

The reference evaluates one graph per call on one device (nmrgnn/main.py:236-245); graphs never
interact (tf.gather indexes within one graph's node tensor, nmrgnn/layers.py:33), so a batch of
independent graphs shards by whole graphs with no data-path collective.  Every rank runs the forward
on the graphs it owns; the only exchange is the reassembly of the float32 peaks.

    dist.init_process_group("nccl", ...)            # torchrun; only used to exchange 128-byte IPC blobs
    sm = ShardedModel(nmrgnn_b200.load_model(device=local_rank))
    peaks = sm(batch)                               # batch = (atoms, nlist, edges, inv_degree, graph_offsets)

Reassembly backends (`collective=`):
  "peer"  (default with a CUDA model) the library's own peer-memory exchange: every rank stores its peaks into its
          slot of every rank's gather buffer over NVLink and publishes an epoch flag; a wait kernel on the same
          stream acquires all sources (csrc/peer_gather.cuh, nmrgnn_forward_sharded).  Device-resident end to end:
          the gathered [world, capacity] tensor is permuted into the original atom order by one device gather.
  "torch" one torch.distributed all-gather of the padded peaks (NCCL for CUDA tensors, gloo for host tensors).
          Used for the CPU tests (`local_forward` injectable, gloo, world_size 2) and as the comparison arm.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .workloads import shard_graphs, take_graphs

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


class ShardPlan:
    """Which graphs each rank owns and where their atoms sit in the full batch."""

    def __init__(self, graph_offsets: Sequence[int], world_size: int):
        self.offsets = np.asarray(graph_offsets, np.int64)
        if self.offsets.ndim != 1 or self.offsets.size < 1 or self.offsets[0] != 0 or np.any(np.diff(self.offsets) < 0):
            raise ValueError("graph_offsets must start at 0 and be non-decreasing")
        self.world_size = int(world_size)
        self.owned: List[np.ndarray] = shard_graphs(self.offsets, self.world_size)
        sizes = np.diff(self.offsets)
        self.counts = np.array([int(sizes[g].sum()) for g in self.owned], np.int64)   # atoms per rank
        self.max_count = int(self.counts.max()) if self.counts.size else 0

    @property
    def n_atoms(self) -> int:
        return int(self.offsets[-1])

    def atom_index(self, rank: int) -> np.ndarray:
        """Positions (in the full batch) of the atoms of `rank`'s shard, in shard order."""
        parts = [np.arange(self.offsets[g], self.offsets[g + 1]) for g in self.owned[rank]]
        return np.concatenate(parts) if parts else np.zeros(0, np.int64)

    def scatter_back(self, gathered: np.ndarray) -> np.ndarray:
        """[world, max_count] padded per-rank peaks -> [n_atoms] in the original order."""
        gathered = np.asarray(gathered).reshape(self.world_size, self.max_count)
        out = np.empty(self.n_atoms, gathered.dtype)
        for r in range(self.world_size):
            out[self.atom_index(r)] = gathered[r, :self.counts[r]]
        return out

    def gather_index(self, capacity: int) -> np.ndarray:
        """For every atom of the full batch, its flat position in a [world, capacity] gathered tensor."""
        idx = np.empty(self.n_atoms, np.int64)
        for r in range(self.world_size):
            idx[self.atom_index(r)] = r * int(capacity) + np.arange(self.counts[r])
        return idx


def _world(group=None) -> Tuple[int, int]:
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class PeerGather:
    """The library's peer-memory exchange for one model handle: nmrgnn_comm_local -> exchange of the IPC blobs through
    torch.distributed (plumbing only) -> nmrgnn_comm_init."""

    def __init__(self, model, capacity: int, group=None):
        self.model = model
        self.group = group
        self.rank, self.world = _world(group)
        self.capacity = (int(capacity) + 3) // 4 * 4
        h = model.handle
        blob = h.comm_local(self.capacity, self.world)
        blobs = [blob]
        if self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, blob, group=group)
        h.comm_init(self.rank, self.world, blobs)
        if self.world > 1:
            dist.barrier(group=group)      # every rank has mapped every buffer before the first scatter

    def forward(self, graph, stream_ptr=None):
        """Local shard (host arrays or CUDA tensors) in, device view [world, capacity] of everyone's peaks out."""
        atoms, nlist, edges, inv = graph
        n, k = nlist.shape
        from . import _capi
        mem = _capi.MEM_DEVICE if (torch is not None and isinstance(atoms, torch.Tensor) and atoms.is_cuda) else _capi.MEM_HOST
        if mem == _capi.MEM_DEVICE and stream_ptr is None:
            stream_ptr = int(torch.cuda.current_stream(atoms.device).cuda_stream) or 1
        self.model.handle.forward_sharded(atoms, nlist, edges, inv, n, k, None, mem, stream_ptr)
        return self.view()

    def view(self):
        ptr, cap = self.model.handle.comm_buffer()
        dev = torch.device("cuda", self.model.device)
        n = self.world * cap

        class _Arr:      # __cuda_array_interface__ of the library-owned gather buffer
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Arr(), device=dev).view(self.world, cap)


class ShardedModel:
    def __init__(self, model=None, local_forward: Optional[Callable] = None, group=None, collective: Optional[str] = None):
        if model is None and local_forward is None:
            raise ValueError("need a model or a local_forward callable")
        self.model = model
        self._forward = local_forward if local_forward is not None else (lambda g: model(g))
        self.group = group
        if collective is None:
            collective = "peer" if (model is not None and local_forward is None and torch is not None) else "torch"
        if collective not in ("peer", "torch"):
            raise ValueError("collective must be 'peer' or 'torch'")
        self.collective = collective
        self._peer: Optional[PeerGather] = None

    def _peer_for(self, capacity: int) -> PeerGather:
        if self._peer is None or self._peer.capacity < capacity:
            if self._peer is not None:
                self.model.synchronize()
                if self._peer.world > 1:
                    dist.barrier(group=self.group)
            self._peer = PeerGather(self.model, max(capacity, 1), self.group)
        return self._peer

    def __call__(self, batch):
        """Full batch in (every rank passes the same batch, or at least the same graph_offsets), peaks of the WHOLE
        batch out on every rank, in the original atom order: a NumPy array for host inputs, a CUDA tensor for CUDA
        inputs."""
        atoms, nlist, edges, inv, offs = batch
        rank, world = _world(self.group)
        plan = ShardPlan(np.asarray(offs.cpu() if (torch is not None and isinstance(offs, torch.Tensor)) else offs), world)
        on_device = torch is not None and isinstance(atoms, torch.Tensor) and atoms.is_cuda
        if on_device:
            local = self._take_device(atoms, nlist, edges, inv, plan, rank)
        else:
            local = take_graphs((atoms, nlist, edges, inv, plan.offsets), plan.owned[rank])[:4]
        if self.collective == "peer":
            pg = self._peer_for(plan.max_count)
            gathered = pg.forward(local)                                  # [world, capacity] on the device
            index = torch.from_numpy(plan.gather_index(pg.capacity)).to(gathered.device)
            out = gathered.reshape(-1).index_select(0, index)             # original atom order, still on the device
            return out if on_device else out.cpu().numpy()
        y = self._forward(local)
        if on_device:
            y = y.detach().cpu().numpy()
        y = np.asarray(y, np.float32).reshape(-1)
        if y.shape[0] != plan.counts[rank]:
            raise RuntimeError("local forward returned the wrong number of peaks")
        if world == 1:
            out = plan.scatter_back(y[None, :])
        else:
            backend = dist.get_backend(self.group)
            use_cuda = backend == "nccl"
            dev = torch.device("cuda", self.model.device) if use_cuda and self.model is not None else (
                torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu"))
            send = torch.zeros(plan.max_count, dtype=torch.float32, device=dev)
            send[:y.shape[0]] = torch.from_numpy(y).to(dev)
            recv = torch.empty(world * plan.max_count, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            out = plan.scatter_back(recv.cpu().numpy())
        return torch.from_numpy(out).to(atoms.device) if on_device else out

    @staticmethod
    def _take_device(atoms, nlist, edges, inv, plan: ShardPlan, rank: int):
        """This rank's graphs of a device-resident batch, re-offset (device gathers, no host round trip of the data)."""
        dev = atoms.device
        idx = torch.from_numpy(plan.atom_index(rank)).to(dev)
        # new position of every old atom index that this rank owns (neighbours stay inside their graph)
        remap = torch.zeros(max(plan.n_atoms, 1), dtype=torch.int32, device=dev)
        remap[idx] = torch.arange(idx.numel(), dtype=torch.int32, device=dev)
        nl = remap[nlist.index_select(0, idx).long()]
        # padded slots (index = graph start + 0 with edge 0) keep pointing at their graph's first atom
        return (atoms.index_select(0, idx).contiguous(), nl.contiguous(), edges.index_select(0, idx).contiguous(),
                inv.index_select(0, idx).contiguous())

    def close(self):
        if self._peer is not None and self.model is not None:
            self.model.handle.comm_destroy()
            self._peer = None
