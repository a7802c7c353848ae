"""Seeded synthetic graph batches shaped like the BASELINE.json configs
(SURVEY.md §8d).  Used by bench.py and the tests; inputs only — no model code."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np

from .graph import batch_graphs, inv_degree_from_nlist, knn_graph_host, one_hot_elements

# i.i.d. element mix of a hydrogenated protein (108M.pdb: H .505 / C .319 / N .088 / O .087 / S .001)
_PROTEIN_ELEMENTS = (("H", 0.505), ("C", 0.319), ("N", 0.088), ("O", 0.087), ("S", 0.001))
_SMALL_ELEMENTS = (("H", 0.50), ("C", 0.35), ("N", 0.07), ("O", 0.07), ("S", 0.01))


def _elements(rng, n, table):
    names = [t[0] for t in table]
    p = np.array([t[1] for t in table], np.float64)
    return np.array(names)[rng.choice(len(names), size=n, p=p / p.sum())]


def _chain_positions(rng, n: int, density_per_nm3: float = 55.0, d_min: float = 0.115) -> np.ndarray:
    """Self-avoiding random chain: 0.10-0.15 nm steps, confined to a sphere sized
    for the target atom density, every pair at least ``d_min`` apart (cell-grid
    rejection).  Consecutive indices are spatial neighbours (PDB-like index
    locality).  Density and minimum distance are chosen so that the kNN-16 distance
    distribution (median ~0.24 nm, nothing below a bond length) keeps the
    pretrained model in the regime it was trained on: with denser / overlapping
    atoms its activations blow up and fp32 itself is no longer stable to 1e-4."""
    radius = (3.0 * n / (4.0 * np.pi * density_per_nm3)) ** (1.0 / 3.0)
    r2max = radius * radius
    inv_cell = 1.0 / d_min
    d2min = d_min * d_min
    grid = {}
    pos = np.empty((n, 3), np.float64)
    cur = np.zeros(3)
    n_try = 24
    dirs = rng.normal(size=(n, n_try, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    lens = rng.uniform(0.10, 0.15, size=(n, n_try))
    offsets = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]

    def min_d2(p):
        cx, cy, cz = int(np.floor(p[0] * inv_cell)), int(np.floor(p[1] * inv_cell)), int(np.floor(p[2] * inv_cell))
        best = np.inf
        for a, b, c in offsets:
            for q in grid.get((cx + a, cy + b, cz + c), ()):
                d = pos[q] - p
                d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
                if d2 < best:
                    best = d2
        return best

    for i in range(n):
        chosen = None
        if i == 0:
            chosen = cur
        else:
            cand = cur + dirs[i] * lens[i][:, None]
            inside = np.einsum("ij,ij->i", cand, cand) <= r2max
            fallback, fb_d2 = None, -1.0
            for t in range(n_try):
                if not inside[t]:
                    continue
                d2 = min_d2(cand[t])
                if d2 >= d2min:
                    chosen = cand[t]
                    break
                if d2 > fb_d2:
                    fallback, fb_d2 = cand[t], d2
            if chosen is None:
                # trapped: restart the chain from a random earlier atom, keep the best
                # candidate only if nothing else works
                for _ in range(64):
                    j = int(rng.integers(0, i))
                    v = rng.normal(size=3)
                    c2 = pos[j] + v / np.linalg.norm(v) * rng.uniform(0.10, 0.15)
                    if c2 @ c2 <= r2max and min_d2(c2) >= d2min:
                        chosen = c2
                        break
                if chosen is None:
                    chosen = fallback if fallback is not None else cur + dirs[i, 0] * lens[i, 0]
        pos[i] = chosen
        key = (int(np.floor(chosen[0] * inv_cell)), int(np.floor(chosen[1] * inv_cell)),
               int(np.floor(chosen[2] * inv_cell)))
        grid.setdefault(key, []).append(i)
        cur = chosen
    return pos


def protein_graph(seed: int, n_lo: int = 2000, n_hi: int = 3000, neighbor_number: int = 16,
                  pad_fraction: float = 0.01, num_elem: int = 10):
    """Config-2/4 style graph: N ~ U{n_lo..n_hi}, kNN-16, ~1 % of slots zero-padded."""
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    n = int(rng.integers(n_lo, n_hi + 1))
    pos = _chain_positions(rng, n)
    nlist, edges = knn_graph_host(pos, neighbor_number)
    if pad_fraction > 0:
        pad = rng.random(nlist.shape) < pad_fraction
        nlist[pad] = 0
        edges[pad] = 0.0
    atoms = one_hot_elements(_elements(rng, n, _PROTEIN_ELEMENTS), num_elem, warn=False)
    return atoms, nlist, edges, inv_degree_from_nlist(nlist)


def small_molecule_graph(seed: int, n_lo: int = 20, n_hi: int = 60, neighbor_number: int = 8,
                         pad_fraction: float = 0.10, num_elem: int = 10):
    """Config-3 style graph: N ~ U{20..60}, K=8, bonded-like distances, ~10 % padding."""
    rng = np.random.default_rng(7_000_003 * (seed + 1))
    n = int(rng.integers(n_lo, n_hi + 1))
    pos = _chain_positions(rng, n, density_per_nm3=50.0)
    nlist, edges = knn_graph_host(pos, neighbor_number)
    pad = rng.random(nlist.shape) < pad_fraction
    nlist[pad] = 0
    edges[pad] = 0.0
    atoms = one_hot_elements(_elements(rng, n, _SMALL_ELEMENTS), num_elem, warn=False)
    return atoms, nlist, edges, inv_degree_from_nlist(nlist)


def _gen_protein(args):
    seed, kw = args
    return protein_graph(seed, **kw)


def _gen_small(args):
    seed, kw = args
    return small_molecule_graph(seed, **kw)


def _generate(fn, seeds, kw, workers):
    seeds = list(seeds)
    if workers is None:
        workers = min(os.cpu_count() or 1, 64)
    if workers <= 1 or len(seeds) < 8:
        return [fn((s, kw)) for s in seeds]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(fn, [(s, kw) for s in seeds], chunksize=max(1, len(seeds) // (4 * workers)))


def protein_graphs(seeds, workers: Optional[int] = None, **kw):
    """List of protein-like graphs, generated in parallel on the host cores."""
    return _generate(_gen_protein, seeds, kw, workers)


def protein_batch(n_graphs: int = 64, first_seed: int = 0, workers: Optional[int] = None, **kw):
    return batch_graphs(protein_graphs(range(first_seed, first_seed + n_graphs), workers, **kw))


def small_molecule_batch(n_graphs: int = 1024, first_seed: int = 0, workers: Optional[int] = None, **kw):
    return batch_graphs(_generate(_gen_small, range(first_seed, first_seed + n_graphs), kw, workers))


def protein_graph_size(seed: int, n_lo: int = 2000, n_hi: int = 3000) -> int:
    """Atom count protein_graph(seed) will have, without generating it (used to
    balance shards before each rank generates only its own graphs)."""
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    return int(rng.integers(n_lo, n_hi + 1))


def ring_graph(n: int = 5, num_elem: int = 16, neighbor_number: int = 2):
    """The reference's own unit-test graph (tests/test_nmrgnn.py:20-31,198-210):
    each atom bonded to i-1 and i+1 mod n, edges = 1, inv_degree = 1/2."""
    order = np.array([2, 4, 0, 1, 3])
    idx = np.resize(order, n) % num_elem
    atoms = np.zeros((n, num_elem), np.float32)
    atoms[np.arange(n), idx] = 1.0
    nlist = np.zeros((n, neighbor_number), np.int32)
    for i in range(n):
        for k, j in enumerate(range(-1, 3, 2)):
            if k < neighbor_number:
                nlist[i, k] = (i + j) % n
    edges = np.ones((n, neighbor_number), np.float32)
    inv_degree = np.full(n, 0.5, np.float32)
    return atoms, nlist, edges, inv_degree


def shard_graphs(graph_offsets: np.ndarray, world_size: int) -> List[np.ndarray]:
    """Greedy balance of whole graphs over ranks by atom count (largest first);
    returns, per rank, the sorted graph ids it owns."""
    sizes = np.diff(np.asarray(graph_offsets, np.int64))
    loads = np.zeros(world_size, np.int64)
    owner: List[List[int]] = [[] for _ in range(world_size)]
    for g in np.argsort(-sizes, kind="stable"):
        r = int(np.argmin(loads))
        owner[r].append(int(g))
        loads[r] += sizes[g]
    return [np.array(sorted(o), np.int64) for o in owner]


def take_graphs(batch, graph_ids: np.ndarray):
    """Sub-batch made of the given graphs (re-offsetting nlist)."""
    atoms, nlist, edges, inv, offs = batch
    parts = []
    for g in graph_ids:
        a, b = int(offs[g]), int(offs[g + 1])
        parts.append((atoms[a:b], nlist[a:b].astype(np.int64) - a, edges[a:b], inv[a:b]))
    if not parts:
        c, k = atoms.shape[1], nlist.shape[1]
        return (np.zeros((0, c), np.float32), np.zeros((0, k), np.int32), np.zeros((0, k), np.float32),
                np.zeros((0,), np.float32), np.zeros(1, np.int64))
    return batch_graphs(parts)
