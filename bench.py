#!/usr/bin/env python
"""Benchmark of the GNN forward hot path (BASELINE.json metric: atoms/s of the
message-passing forward; HBM GB/s / TFLOP/s against the measured B200 roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward over one batch of synthetic graphs.  --config selects the BASELINE.json configuration
(numbered as in BASELINE.md, 1-based):
  2 (default)  64 protein-like graphs (~2 500 atoms, 16-wide neighbour list) PER GPU: weak scaling over N
  3            1 024 small molecules (~40 atoms, 8-wide neighbour list), sharded by graph over N (strong scaling)
  4            1 024 protein graphs sharded by graph over N GPUs (strong scaling; --graphs-total G for other sizes)
  5            MD stream: 108M.pdb x 512 jittered frames, graph build + forward per frame batch, frames round-robin
               over N GPUs (metric: frames/s; per-frame latency reported)
N > 1 (torchrun, one rank per GPU): every rank runs the forward on its graphs and the peaks of all ranks are
reassembled on every rank by the library's peer-memory exchange (nmrgnn_forward_sharded: stores over NVLink +
epoch flags, no collective-library call; --collective nccl times torch.distributed's all-gather instead).
At every N the default run also measures config 4 (field `config4_strong`) so that a 1..8 GPU sweep carries the
strong-scaling numbers; --no-config4 skips it.

Prints ONE JSON line (rank 0).  `value` = atoms/s with inputs resident in HBM;
`e2e` = the same through the public host-array API (pinned host buffers, H2D and
D2H inside the timed region); `roofline` describes the dominant kernel;
`cpu_baseline` times the torch-CPU restatement of the reference (TensorFlow is
not installable here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GRAPHS_PER_GPU = 64
K_NEIGH = 16
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def mp_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one MP-layer launch from the committed ncu capture
    (profiles/r02_mp_traffic.json; same workload).  None if no capture is on record."""
    path = os.path.join(ROOT, "profiles", "r02_mp_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as f:
                d = json.load(f)
            d["_source"] = "measured"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


# FLOPs / bytes per atom (MAC = 2 FLOP), K = neighbours, pretrained dims
def flops_per_atom(K, C=10, R=128, H=128, E=3, F=256, L=4, n_edge_hidden=3, n_fc_res=3):
    F2 = F // 2
    edge = K * 2 * (R * H + (n_edge_hidden - 1) * H * H + H * E)
    embed = 2 * C * F
    mp = L * (2 * K * F * E + 2 * F * F * E)
    fc = n_fc_res * 2 * F * F + 2 * F * F2 + 2 * F2 * C
    return dict(edge=edge, embed=embed, mp_layer=mp // L, fc=fc, total=edge + embed + mp + fc)


def bytes_per_atom(K, C=10, E=3, F=256, L=4):
    mp_layer = 8 * F + 4 * K + 4 * K * E + 4
    return dict(mp_layer=mp_layer, edge=4 * K + 4 * K * E, embed=4 * C + 4 * F, fc=4 * F + 4 * C + 4,
                total=(4 * K + 4 * K * E) + (4 * C + 4 * F) + L * mp_layer + (4 * F + 4 * C + 4))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(rank: int, world: int, config: int, graphs_total: int = 0):
    """This rank's graphs of the selected configuration (a rank only generates the graphs it owns).
    config 2: weak scaling, 64 graphs per GPU: the 64 N graphs with seeds 0 .. 64 N - 1; configs 3 / 4: the G graphs with
    seeds 0..G-1.  In every case the graphs are assigned to the ranks by the greedy atom-count balance of
    nmrgnn_b200.sharding.ShardPlan (at N = 1 config 2 is exactly seeds 0..63)."""
    from nmrgnn_b200 import workloads
    from nmrgnn_b200.graph import batch_graphs
    from nmrgnn_b200.sharding import ShardPlan
    if config == 2 and graphs_total <= 0:
        if world == 1:
            return workloads.protein_batch(GRAPHS_PER_GPU, first_seed=0, neighbor_number=K_NEIGH)
        graphs_total = GRAPHS_PER_GPU * world
    if config == 3:
        total = graphs_total or 1024
        sizes = np.array([workloads.small_molecule_graph(s)[0].shape[0] for s in range(total)], np.int64) if world > 1 else None
        if world == 1:
            return workloads.small_molecule_batch(total, first_seed=0)
        plan = ShardPlan(np.concatenate([[0], np.cumsum(sizes)]), world)
        return batch_graphs([workloads.small_molecule_graph(int(g)) for g in plan.owned[rank]])
    total = graphs_total or 1024
    sizes = np.array([workloads.protein_graph_size(s) for s in range(total)], np.int64)
    plan = ShardPlan(np.concatenate([[0], np.cumsum(sizes)]), world)
    graphs = workloads.protein_graphs([int(g) for g in plan.owned[rank]], neighbor_number=K_NEIGH)
    return batch_graphs(graphs)


WORKLOAD_TEXT = {
    2: "config[1]: batch of {g} synthetic protein graphs (~2500 atoms, 16-wide nlist), fp32, pretrained weights, per GPU "
       "(N GPUs: the 64 N graphs with seeds 0 .. 64 N - 1, sharded by the library's greedy atom-count balance)",
    3: "config[2]: batch of {g} small-molecule graphs (~40 atoms, 8-wide nlist), fp32, sharded by graph over {n} GPU(s)",
    4: "config[3]: batch of {g} synthetic protein graphs sharded by graph over {n} GPU(s), peaks reassembled on every rank",
    5: "config[4]: MD trajectory stream, 108M.pdb x {g} jittered frames, graph build + forward per frame batch, "
       "frames round-robin over {n} GPU(s)",
}


def workload_config(n_gpus, config=2, graphs=GRAPHS_PER_GPU, **extra):
    cfg = {"workload": WORKLOAD_TEXT[config].format(g=graphs, n=n_gpus),
           "parallelism": (f"graph-sharded x{n_gpus}, peaks reassembled over peer memory (NVLink)" if n_gpus > 1 else "single GPU"),
           "l2_policy": "per-step working set (inputs + node / edge-record buffers, ~2.3 KB per atom) exceeds the 126 MB L2 "
                        "for configs 2 and 4; configs 3 and 5 fit and say so in `l2_note`"}
    if config == 2:
        cfg.update(graphs_per_gpu=GRAPHS_PER_GPU, neighbor_number=K_NEIGH)
    cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path — restated in
    torch (reference einsum order, all host threads), one graph per call like the
    reference — on a bounded sample of this arm's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from nmrgnn_b200.params import GNNParams, baseline_path
    from nmrgnn_b200.workloads import take_graphs
    from oracle.forward_torch import TorchReference
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.config == 3:
        from nmrgnn_b200 import workloads
        batch = workloads.small_molecule_batch(64, first_seed=0)
        what = "64 molecules"
    else:
        from nmrgnn_b200 import workloads
        batch = workloads.protein_batch(2, first_seed=0, neighbor_number=K_NEIGH)
        what = "2 graphs"
    n_atoms = int(batch[0].shape[0])
    ref = TorchReference(GNNParams.load(baseline_path()), reference_order=True)
    for _ in range(args.warmup):
        ref.per_graph(batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.per_graph(batch)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_atoms / dt
    unit, metric = "atoms/s", "atoms/sec MP-GNN forward"
    if args.config == 5:        # one frame = one 108M-sized graph
        value, unit, metric = value / 2482.0, "frames/s", "frames/sec MD-stream inference (108M.pdb-sized frames)"
    sample = (f"{what} ({n_atoms} atoms) of the workload per step, one graph per call, "
              f"torch {torch.__version__} CPU fp32, reference einsum order (graph build not included)")
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.config == 2 else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.gpus, args.config, {2: GRAPHS_PER_GPU, 3: 1024, 4: args.graphs_total or 1024, 5: 512}[args.config],
                                  sample=sample),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(batch, budget_s=12.0, n_graphs=2):
    import torch
    from nmrgnn_b200.params import GNNParams, baseline_path
    from nmrgnn_b200.workloads import take_graphs
    from oracle.forward_torch import TorchReference
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = TorchReference(GNNParams.load(baseline_path()), reference_order=True)
    sub = take_graphs(batch, np.arange(n_graphs))
    ref.per_graph(sub)                      # warm-up
    best, n_runs, t_start = None, 0, time.perf_counter()
    while n_runs < 5 and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        ref.per_graph(sub)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        n_runs += 1
    n_atoms = int(sub[0].shape[0])
    return {"value": n_atoms / best, "unit": "atoms/s", "cores": cores, "kind": "port",
            "sample": f"first {n_graphs} graphs ({n_atoms} atoms) of the workload, one graph per call, best of {n_runs}; "
                      f"torch-CPU restatement of the reference in its einsum order (TensorFlow not installable)"}


class Bench:
    """Device / e2e timing of one sharded workload on this rank (shared by the main measurement and `config4_strong`)."""

    def __init__(self, model, batch, world, rank, dev, collective):
        import torch
        import torch.distributed as dist
        from nmrgnn_b200 import _capi
        from nmrgnn_b200.sharding import PeerGather
        self.torch, self.dist, self.capi = torch, dist, _capi
        self.model, self.h, self.world, self.rank, self.dev = model, model.handle, world, rank, dev
        self.collective = collective
        atoms, nlist, edges, inv, offs = batch
        self.n_atoms = int(atoms.shape[0])
        self.k = int(nlist.shape[1])
        self.n_graphs = len(offs) - 1
        self.d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)]
        self.pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (atoms, nlist, edges, inv)]
        self.pin_np = [t.numpy() for t in self.pin]
        self.h2d = sum(t.numel() * t.element_size() for t in self.pin)
        counts = [self.n_atoms]
        if world > 1:
            c = torch.tensor([self.n_atoms], device=dev, dtype=torch.int64)
            allc = [torch.zeros_like(c) for _ in range(world)]
            dist.all_gather(allc, c)
            counts = [int(x.item()) for x in allc]
        self.counts = counts
        self.total_atoms = sum(counts)
        self.max_n = max(counts)
        self.stream = torch.cuda.current_stream(dev)
        self.sptr = int(self.stream.cuda_stream) or 1
        if world > 1 and collective == "peer":
            self.pg = PeerGather(model, self.max_n)
            cap = self.pg.capacity
            self.out_pin = torch.empty(world * cap, dtype=torch.float32).pin_memory()
        else:
            self.pg = None
            self.d_peaks = torch.zeros(self.max_n, dtype=torch.float32, device=dev)
            self.gather_out = torch.empty(world * self.max_n, dtype=torch.float32, device=dev) if world > 1 else None
            self.out_pin = torch.empty(world * self.max_n if world > 1 else self.n_atoms, dtype=torch.float32).pin_memory()
        self.out_np = self.out_pin.numpy()
        self.d2h = self.out_pin.numel() * 4

    def step_device(self):
        h, d, c = self.h, self.d_in, self.capi
        if self.pg is not None:
            h.forward_sharded(d[0], d[1], d[2], d[3], self.n_atoms, self.k, None, c.MEM_DEVICE, self.sptr)
        else:
            h.forward(d[0], d[1], d[2], d[3], self.n_atoms, self.k, self.d_peaks, c.MEM_DEVICE, self.sptr)
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.gather_out, self.d_peaks)

    def step_e2e(self):
        h, p, c = self.h, self.pin_np, self.capi
        if self.pg is not None:       # host buffers in, every rank's peaks back in host memory: copies + exchange inside
            h.forward_sharded(p[0], p[1], p[2], p[3], self.n_atoms, self.k, self.out_np, c.MEM_HOST, None)
        elif self.world > 1:
            tmp = self.out_np[:self.n_atoms]
            h.forward(p[0], p[1], p[2], p[3], self.n_atoms, self.k, tmp, c.MEM_HOST, None)
            self.d_peaks[:self.n_atoms].copy_(self.out_pin[:self.n_atoms], non_blocking=True)
            self.dist.all_gather_into_tensor(self.gather_out, self.d_peaks)
            self.out_pin.copy_(self.gather_out, non_blocking=True)
            self.torch.cuda.synchronize(self.dev)
        else:
            h.forward(p[0], p[1], p[2], p[3], self.n_atoms, self.k, self.out_np, c.MEM_HOST, None)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps, sampler=None):
        torch = self.torch
        self.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        return e0.elapsed_time(e1), wall * 1e3, clocks

    def reduce_max(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def close(self):
        if self.pg is not None:
            self.h.synchronize(self.sptr)
            self.barrier()
            self.h.comm_destroy()


def run_md_stream(args, model, world, rank, local_rank, dev):
    """config 5: 108M.pdb x 512 frames (seeded Gaussian jitter, sigma = 0.02 nm), graph build on the GPU + forward per
    frame batch, frames round-robin over the ranks.  `value` = frames/s from positions resident in pinned host memory
    to peaks in host memory (that IS the end-to-end path of a trajectory: there is no device-resident variant of a
    stream), `device` = the compute stream's time alone."""
    import torch
    import torch.distributed as dist
    from nmrgnn_b200.mdstream import FrameStream
    n_frames = 512
    with np.load(os.path.join(ROOT, "tests", "golden", "g108m_structure.npz")) as z:
        pos = z["positions_A"].astype(np.float32) / np.float32(10)
        elements = [str(e) for e in z["elements"]]
    rng = np.random.default_rng(0)
    frames = pos[None] + rng.normal(scale=0.02, size=(n_frames,) + pos.shape).astype(np.float32)
    fs = FrameStream(model, elements, pos.shape[0], K_NEIGH)
    for _ in range(max(args.warmup, 3)):
        fs.run(frames[:4 * fs.B], 0, 1)
    steps = max(1, min(args.steps, 20))
    l0 = model.handle.kernel_launches
    walls, devs = [], []
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        r = fs.run(frames, rank, world)
        walls.append(r["seconds"])
        devs.append(r["device_ms"] * 1e-3)
    clocks = sampler.stop() if sampler else None
    wall, devt = float(np.mean(walls)), float(np.mean(devs))
    if world > 1:
        t = torch.tensor([wall, devt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, devt = float(t[0]), float(t[1])
    # where a batch's device time goes: graph build alone vs forward alone, same buffers, plain launches
    from nmrgnn_b200 import _capi
    hh, sp = model.handle, int(fs.compute.cuda_stream)
    split = {}
    with torch.cuda.stream(fs.compute):
        for name, fn in (("knn_graph", lambda: hh.knn_graph(fs.d_pos[0], fs.offsets, fs.B * fs.n, fs.B, fs.k, 0.0, fs.d_nlist, fs.d_edges,
                                                             fs.d_inv, _capi.MEM_DEVICE, sp)),
                         ("forward", lambda: hh.forward(fs.d_atoms, fs.d_nlist, fs.d_edges, fs.d_inv, fs.B * fs.n, fs.k, fs.d_peaks[0],
                                                        _capi.MEM_DEVICE, sp))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn()
            e0.record(fs.compute)
            for _ in range(20):
                fn()
            e1.record(fs.compute)
            fs.compute.synchronize()
            split[name + "_ms_per_batch"] = e0.elapsed_time(e1) / 20
    # single-frame latency: one frame per launch, no batching (what an interactive caller sees)
    one = FrameStream(model, elements, pos.shape[0], K_NEIGH, frames_per_batch=1)
    one.run(frames[:8])
    lat = one.run(frames[:64])
    if rank == 0:
        n = pos.shape[0]
        line = {
            "metric": "frames/sec MD-stream inference (108M.pdb-sized frames)", "value": n_frames / wall, "unit": "frames/s",
            "atoms_per_s": n_frames * n / wall, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (108M.pdb coordinates + seeded jitter)",
            "config": workload_config(world, 5, n_frames, atoms_per_frame=n, frames_per_batch=fs.B, cuda_graph=fs.graph_captured,
                                      compute_path=model.handle.compute_path,
                                      l2_note=f"one batch ({fs.B} frames) is {fs.B * n * 2308 / 1e6:.0f} MB of node / record "
                                              "buffers: L2-resident; frames differ from batch to batch"),
            "clocks": clocks,
            "device": {"frames_per_s": n_frames / devt, "ms_per_frame": devt * 1e3 / n_frames,
                       "note": "compute stream only (graph build + forward), max over ranks", **split},
            "latency": {"ms_per_frame_unbatched": lat["seconds"] * 1e3 / 64, "ms_per_batch": wall * 1e3 / max(1, -(-n_frames // fs.B) // world),
                        "frames_per_batch": fs.B},
            "e2e": {"value": n_frames / wall, "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes // world),
                    "d2h_bytes_per_step": int(n_frames * n * 4 // world),
                    "api": "FrameStream.run (nmrgnn_knn_graph + nmrgnn_forward, CUDA graph per batch, pinned staging)"},
            "gpu_launches": int(model.handle.kernel_launches - l0),
            "gpu_launches_note": "kernels enqueued through the C ABI while capturing; graph replays re-launch 7 kernels per batch",
            "cpu_baseline": None,
        }
        print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nmrgnn_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = nmrgnn_b200.load_model(device=local_rank)
    h = model.handle
    if args.config == 5:
        run_md_stream(args, model, world, rank, local_rank, dev)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    config = 4 if (args.graphs_total > 0 and args.config == 2) else args.config
    batch = make_workload(rank, world, config, args.graphs_total)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu = cpu_baseline(batch, n_graphs=64 if config == 3 else 2)     # before any rank spins in a barrier (N = 1 only)
    b = Bench(model, batch, world, rank, dev, args.collective)
    n_atoms = b.n_atoms

    for _ in range(max(args.warmup, 3)):
        b.step_device()
    h.synchronize(b.sptr)
    l0 = h.kernel_launches
    ms_dev, _, clocks = b.timed(b.step_device, args.steps, ClockSampler(local_rank) if rank == 0 else None)
    launches = h.kernel_launches - l0
    h.synchronize(b.sptr)

    # e2e: the host API is synchronous, so the wall clock covers copies + kernels (+ the exchange for N > 1)
    for _ in range(3):
        b.step_e2e()
    _, wall_e2e, _ = b.timed(b.step_e2e, args.steps)

    # e2e, pipelined: the same host buffers through BatchStream (two batches in flight: upload of step i + 1 and download
    # of step i - 1 overlap the kernels of step i; every step still moves its own inputs and its own peaks over PCIe)
    from nmrgnn_b200.batchstream import BatchStream
    if world > 1 and b.pg is None:        # --collective torch: no pipelined form, the synchronous number stands
        wall_pipe, d2h_pipe = wall_e2e, b.d2h
    else:
        bs = BatchStream(model, b.n_atoms, b.k, peer=b.pg)
        bs.run([b.pin] * 4, keep=False)
        b.barrier()
        t0 = time.perf_counter()
        bs.run([b.pin] * args.steps, keep=False)
        b.barrier()
        wall_pipe = (time.perf_counter() - t0) * 1e3
        d2h_pipe = bs.out_len * 4

    # per-kernel timing for the roofline object: CUDA events recorded by the library on the launching
    # stream around each stage of the same forward (option "profile"), averaged over the timed steps
    h.set_option("profile", 1)
    acc = None
    for _ in range(args.steps):
        b.step_device()
        st = h.stage_times()
        flat = [st["edge"], st["embed"]] + list(st["mp_layers"]) + [st["fc_readout"]]
        acc = flat if acc is None else [a + c for a, c in zip(acc, flat)]
    h.set_option("profile", 0)
    acc = [a / args.steps for a in acc]
    n_mp = len(acc) - 3
    kern = {"edge": acc[0], "embed": acc[1], "mp_layer": sum(acc[2:2 + n_mp]) / n_mp, "fc_readout": acc[-1]}
    h.synchronize(b.sptr)

    # the same step with the edge block evaluated per edge by the tcgen05 edge-MLP kernel instead of the table
    h.set_option("edge_table", 0)
    for _ in range(3):
        b.step_device()
    ms_mlp, _, _ = b.timed(b.step_device, min(args.steps, 10))
    ms_mlp /= min(args.steps, 10)
    path_mlp = h.compute_path
    h.set_option("edge_table", 1)
    # the same step with the MP layers on the single-accumulator kernel (faster, less accurate: see the header)
    h.set_option("mp_single_acc", 1)
    for _ in range(3):
        b.step_device()
    ms_one, _, _ = b.timed(b.step_device, min(args.steps, 10))
    ms_one /= min(args.steps, 10)
    h.set_option("mp_single_acc", 0)
    ms_dev, wall_e2e, ms_mlp, wall_pipe, ms_one = b.reduce_max(ms_dev, wall_e2e, ms_mlp, wall_pipe, ms_one)
    total_atoms, total_graphs = b.total_atoms, None
    if world > 1:
        tg = torch.tensor([b.n_graphs], device=dev, dtype=torch.int64)
        dist.all_reduce(tg)
        total_graphs = int(tg.item())
    else:
        total_graphs = b.n_graphs
    h2d, d2h = b.h2d, b.d2h
    compute_path = h.compute_path
    b.close()

    # strong scaling on config 4 (1 024 protein graphs over the same N ranks) as an extra field of this line
    c4 = None
    if config == 2 and not args.no_config4:
        batch4 = make_workload(rank, world, 4, 1024)
        b4 = Bench(model, batch4, world, rank, dev, args.collective)
        for _ in range(3):
            b4.step_device()
        s4 = 5
        ms4, _, _ = b4.timed(b4.step_device, s4)
        (ms4,) = b4.reduce_max(ms4)
        c4 = {"workload": WORKLOAD_TEXT[4].format(g=1024, n=world), "total_atoms": b4.total_atoms, "ms_per_step": ms4 / s4,
              "value": b4.total_atoms / (ms4 / s4 * 1e-3), "unit": "atoms/s", "steps": s4, "scaling": "strong",
              "atoms_on_largest_rank": b4.max_n}
        b4.close()

    if rank == 0:
        peaks = measured_peaks()
        K = b.k
        fl = flops_per_atom(K)
        by = bytes_per_atom(K)
        ms_step = ms_dev / args.steps
        value = total_atoms / (ms_step * 1e-3)
        e2e_ms = wall_e2e / args.steps
        pipe_ms = wall_pipe / args.steps
        t_mp = kern["mp_layer"] * 1e-3
        mp_tflops = n_atoms * fl["mp_layer"] / t_mp / 1e12
        mp_gbs = (n_atoms * by["mp_layer"] + 4 * 256 * 256 * 3) / t_mp / 1e9
        step_kernel_ms = kern["edge"] + kern["embed"] + 4 * kern["mp_layer"] + kern["fc_readout"]
        roofline = {
            "kernel": "mp_layer_tc_kernel (one of 4 launches per step; " + compute_path + ")", "bound": "tensor",
            "achieved": mp_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": mp_tflops / peaks["bf16_tflops"], "traffic": mp_traffic_bytes() if config == 2 else None,
            "traffic_note": "ncu dram bytes of one MP-layer launch on config 2 (profiles/r02_mp_traffic.json); algorithmic "
                            "bytes of the launch = atoms_per_gpu * 2308 + 786432",
            "peak_source": peaks["_source"] + " bf16 dense (MEASURED_PEAKS.json); achieved counts ALGORITHMIC flops: the "
                           "path needs fp32-accurate products and executes every one as three fp16 tensor-core products "
                           "(fp16x3), so the tensor pipe does 3x this work (executed_frac) and 1/3 of the peak is its ceiling",
            "executed_frac": 3.0 * mp_tflops / peaks["bf16_tflops"],
            "launch_ms": kern["mp_layer"], "share_of_step": 4 * kern["mp_layer"] / step_kernel_ms,
            "hbm": {"achieved": mp_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": mp_gbs / peaks["hbm_gbs"],
                    "bytes_per_atom": by["mp_layer"]},
            "flops_per_atom": fl["mp_layer"],
            "kernels_ms": kern,
            "whole_forward": {"tflops": total_atoms / world * (fl["total"] - fl["edge"]) / (ms_step * 1e-3) / 1e12,
                              "flops_per_atom": fl["total"] - fl["edge"], "bytes_per_atom": by["total"],
                              "note": "the edge MLP's flops (K * 99 072 per atom) are not executed on the table path and "
                                      "not counted here"},
        }
        graphs = {2: GRAPHS_PER_GPU, 3: total_graphs, 4: total_graphs}[config]
        extra = {}
        if config in (3, 4):
            extra["graphs_total"] = total_graphs
        if config == 3:
            extra["l2_note"] = "41 k atoms: the 95 MB per-step working set fits the 126 MB L2 (inputs are re-read from it every step)"
        line = {
            "metric": "atoms/sec MP-GNN forward", "value": value, "unit": "atoms/s",
            "graphs_per_s": (total_graphs if config != 2 else world * GRAPHS_PER_GPU) / (ms_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if config == 2 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, config, graphs, atoms_per_gpu=n_atoms, total_atoms=total_atoms,
                                      compute_path=compute_path, neighbor_number=K,
                                      collective=("none" if world == 1 else
                                                  "peer-memory exchange inside nmrgnn_forward_sharded (NVLink stores + epoch flags)"
                                                  if args.collective == "peer" else "torch.distributed all_gather_into_tensor (NCCL)"),
                                      **extra),
            "clocks": clocks,
            "e2e": {"value": total_atoms / (pipe_ms * 1e-3), "unit": "atoms/s", "ms_per_step": pipe_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h_pipe),
                    "api": "nmrgnn_b200.BatchStream.run over host batches in pinned memory: every step uploads its own "
                           "inputs and downloads its own peaks (all ranks' peaks at N > 1: nmrgnn_forward_sharded, exchange "
                           "inside); two steps in flight, so the copies of step i +- 1 overlap the kernels of step i",
                    "synchronous_call": {
                        "value": total_atoms / (e2e_ms * 1e-3), "unit": "atoms/s", "ms_per_step": e2e_ms,
                        "d2h_bytes_per_step": d2h,
                        "api": ("nmrgnn_forward_sharded(NMRGNN_MEM_HOST): pinned host buffers in, every rank's peaks back "
                                "in host memory" if world > 1 and args.collective == "peer" else
                                "nmrgnn_forward(NMRGNN_MEM_HOST) with pinned host buffers"),
                        "note": "one blocking call per step: upload, kernels and download one after the other"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "edge_mlp_variant": {"compute_path": path_mlp, "ms_per_step": ms_mlp, "value": total_atoms / (ms_mlp * 1e-3),
                                 "unit": "atoms/s", "note": "option edge_table = 0: the edge block evaluated per edge by the "
                                                            "tcgen05 edge-MLP kernel instead of the create-time FP64 table"},
            "mp_single_acc_variant": {"ms_per_step": ms_one, "value": total_atoms / (ms_one * 1e-3), "unit": "atoms/s",
                                      "note": "option mp_single_acc = 1 (not the default): MP layers with main and correction "
                                              "products in one accumulator, epilogue under the next tile's MMAs; max error "
                                              "on this workload 0.83 of the tolerance instead of 0.58 "
                                              "(tests/test_gpu_parity.py::test_single_accumulator_kernel_within_tolerance)"},
            "config4_strong": c4,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (1-based)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: the library's peer-memory exchange (default) or torch.distributed's NCCL all-gather")
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only (ncu)")
    ap.add_argument("--no-config4", action="store_true", help="skip the extra config-4 strong-scaling measurement")
    ap.add_argument("--graphs-total", type=int, default=0,
                    help="configs 3 / 4: number of graphs in the fixed batch (default 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
