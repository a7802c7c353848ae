#!/usr/bin/env python
"""Benchmark of the GNN forward hot path (BASELINE.json metric: atoms/s of the
message-passing forward; HBM GB/s / TFLOP/s against the measured B200 roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward over one batch of synthetic graphs (config 2 of
BASELINE.json: 64 protein-like graphs, ~2 500 atoms each, 16-wide neighbour list,
pretrained weights).  N > 1 (torchrun): every rank owns its own 64-graph shard
(weak scaling), runs the same forward and the ranks all-gather their peaks (NCCL).

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
    (profiles/r01_mp_traffic.json; same workload).  None if no capture is on record."""
    path = os.path.join(ROOT, "profiles", "r01_mp_traffic.json")
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


def make_workload(rank: int, n_graphs: int, world: int = 1, graphs_total: int = 0):
    """Weak scaling (default): every rank generates its own `n_graphs` graphs (seeds rank*n .. rank*n+n-1).
    Strong scaling (--graphs-total G, BASELINE config 4): the G graphs with seeds 0..G-1 are assigned to ranks by the
    greedy atom-count balance of nmrgnn_b200.sharding.ShardPlan; a rank only generates the graphs it owns."""
    from nmrgnn_b200 import workloads
    if graphs_total <= 0:
        return workloads.protein_batch(n_graphs, first_seed=rank * n_graphs, neighbor_number=K_NEIGH)
    from nmrgnn_b200.graph import batch_graphs
    from nmrgnn_b200.sharding import ShardPlan
    sizes = np.array([workloads.protein_graph_size(s) for s in range(graphs_total)], np.int64)
    plan = ShardPlan(np.concatenate([[0], np.cumsum(sizes)]), world)
    graphs = workloads.protein_graphs([int(g) for g in plan.owned[rank]], neighbor_number=K_NEIGH)
    return batch_graphs(graphs)


def workload_config(n_gpus, **extra):
    cfg = {"workload": f"config[1]: batch of {GRAPHS_PER_GPU} synthetic protein graphs (~2500 atoms, 16-wide nlist), "
                       f"fp32, pretrained weights, per GPU",
           "graphs_per_gpu": GRAPHS_PER_GPU, "neighbor_number": K_NEIGH,
           "parallelism": f"graph-sharded x{n_gpus}, all-gather of peaks" if n_gpus > 1 else "single GPU",
           "l2_policy": "per-step working set (inputs 28 MB + node/edge buffers ~360 MB) exceeds the 126 MB L2"}
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
    from oracle.forward_torch import TorchReference
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_graphs = 2
    batch = make_workload(0, sample_graphs)
    n_atoms = int(batch[0].shape[0])
    ref = TorchReference(GNNParams.load(baseline_path()), reference_order=True)
    for _ in range(args.warmup):
        ref.per_graph(batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.per_graph(batch)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_atoms / dt
    sample = (f"{sample_graphs} graphs ({n_atoms} atoms) of the config-2 workload per step, one graph per call, "
              f"torch {torch.__version__} CPU fp32, reference einsum order")
    print(json.dumps({
        "impl": "reference", "metric": "atoms/sec MP-GNN forward", "value": value, "unit": "atoms/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, sample=sample),
        "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(batch, budget_s=12.0):
    import torch
    from nmrgnn_b200.params import GNNParams, baseline_path
    from nmrgnn_b200.workloads import take_graphs
    from oracle.forward_torch import TorchReference
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = TorchReference(GNNParams.load(baseline_path()), reference_order=True)
    sub = take_graphs(batch, np.arange(2))
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
            "sample": f"first 2 graphs ({n_atoms} atoms) of the workload, one graph per call, best of {n_runs}; "
                      f"torch-CPU restatement of the reference in its einsum order (TensorFlow not installable)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nmrgnn_b200
    from nmrgnn_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    batch = make_workload(rank, GRAPHS_PER_GPU, world, args.graphs_total)
    atoms, nlist, edges, inv, offs = batch
    n_atoms = int(atoms.shape[0])
    model = nmrgnn_b200.load_model(device=local_rank)
    h = model.handle

    # device-resident inputs for `value`; pinned host copies for `e2e`
    d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)]
    d_peaks = torch.empty(n_atoms, dtype=torch.float32, device=dev)
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (atoms, nlist, edges, inv)]
    pin_np = [t.numpy() for t in pin]
    out_pin = torch.empty(n_atoms, dtype=torch.float32).pin_memory()
    out_np = out_pin.numpy()
    h2d = sum(t.numel() * t.element_size() for t in pin)
    d2h = out_pin.numel() * 4

    # all-gather plumbing (N > 1): padded to the largest shard
    counts = [n_atoms]
    if world > 1:
        c = torch.tensor([n_atoms], device=dev, dtype=torch.int64)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        counts = [int(x.item()) for x in allc]
    max_n = max(counts)
    gather_in = torch.zeros(max_n, dtype=torch.float32, device=dev)
    gather_out = torch.empty(world * max_n, dtype=torch.float32, device=dev) if world > 1 else None

    stream = torch.cuda.current_stream(dev)
    sptr = int(stream.cuda_stream) or 1

    def step_device():
        h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n_atoms, K_NEIGH, gather_in if world > 1 else d_peaks,
                  _capi.MEM_DEVICE, sptr)
        if world > 1:
            dist.all_gather_into_tensor(gather_out, gather_in)

    def step_e2e():
        h.forward(pin_np[0], pin_np[1], pin_np[2], pin_np[3], n_atoms, K_NEIGH, out_np, _capi.MEM_HOST, None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        return ms, wall * 1e3, clocks

    for _ in range(max(args.warmup, 3)):
        step_device()
    h.synchronize(sptr)
    l0 = h.kernel_launches
    ms_dev, _, clocks = timed(step_device, args.steps, ClockSampler(local_rank) if rank == 0 else None)
    launches = h.kernel_launches - l0
    h.synchronize(sptr)

    # e2e: the host API is synchronous, so the wall clock covers copies + kernels
    for _ in range(3):
        step_e2e()
    _, wall_e2e, _ = timed(step_e2e, args.steps)

    # per-kernel timing for the roofline object: CUDA events recorded by the library on the launching
    # stream around each stage of the same forward (option "profile"), averaged over the timed steps
    h.set_option("profile", 1)
    acc = None
    for _ in range(args.steps):
        step_device()
        st = h.stage_times()
        flat = [st["edge"], st["embed"]] + list(st["mp_layers"]) + [st["fc_readout"]]
        acc = flat if acc is None else [a + b for a, b in zip(acc, flat)]
    h.set_option("profile", 0)
    acc = [a / args.steps for a in acc]
    n_mp = len(acc) - 3
    kern = {"edge": acc[0], "embed": acc[1], "mp_layer": sum(acc[2:2 + n_mp]) / n_mp, "fc_readout": acc[-1]}
    h.synchronize(sptr)

    # max over ranks of the device time
    if world > 1:
        t = torch.tensor([ms_dev, wall_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, wall_e2e = float(t[0]), float(t[1])
        tot = torch.tensor([n_atoms], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        total_atoms = int(tot.item())
    else:
        total_atoms = n_atoms

    if rank == 0:
        peaks = measured_peaks()
        fl = flops_per_atom(K_NEIGH)
        by = bytes_per_atom(K_NEIGH)
        ms_step = ms_dev / args.steps
        value = total_atoms / (ms_step * 1e-3)
        e2e_ms = wall_e2e / args.steps
        t_mp = kern["mp_layer"] * 1e-3
        mp_tflops = n_atoms * fl["mp_layer"] / t_mp / 1e12
        mp_gbs = (n_atoms * by["mp_layer"] + 4 * 256 * 256 * 3) / t_mp / 1e9
        step_kernel_ms = kern["edge"] + kern["embed"] + 4 * kern["mp_layer"] + kern["fc_readout"]
        roofline = {
            "kernel": "mp_layer (" + h.compute_path + ")", "bound": "tensor",
            "achieved": mp_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": mp_tflops / peaks["bf16_tflops"], "traffic": mp_traffic_bytes(),
            "traffic_note": "ncu dram bytes of one MP-layer launch (profiles/r01_final_ncu_summary.md); algorithmic "
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
            "whole_forward": {"tflops": total_atoms / world * fl["total"] / (ms_step * 1e-3) / 1e12,
                              "flops_per_atom": fl["total"], "bytes_per_atom": by["total"]},
        }
        cpu = None if args.skip_cpu_baseline else cpu_baseline(batch)
        line = {
            "metric": "atoms/sec MP-GNN forward", "value": value, "unit": "atoms/s",
            "graphs_per_s": (args.graphs_total if args.graphs_total > 0 else world * GRAPHS_PER_GPU) / (ms_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if args.graphs_total > 0 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, atoms_per_gpu=n_atoms, total_atoms=total_atoms,
                                      compute_path=h.compute_path,
                                      **({"workload": f"config[3]: batch of {args.graphs_total} synthetic protein graphs "
                                                      f"sharded by graph over {world} GPU(s), all-gather of peaks",
                                          "graphs_total": args.graphs_total} if args.graphs_total > 0 else {})),
            "clocks": clocks,
            "e2e": {"value": total_atoms / (e2e_ms * 1e-3), "unit": "atoms/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "nmrgnn_forward(NMRGNN_MEM_HOST) with pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": roofline,
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
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only (ncu)")
    ap.add_argument("--graphs-total", type=int, default=0,
                    help="strong scaling: a fixed batch of this many graphs sharded over the ranks (BASELINE config 4 = 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
