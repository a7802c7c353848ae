"""Graph tuples for the GNN forward: ``(atoms, nlist, edges, inv_degree)``.

Host-side counterpart of what the reference obtains from
``nmrdata.parse_universe`` + the inverse-degree line (nmrgnn/library.py:106-117,
nmrgnn/main.py:239-242).  ``nmrdata`` and MDAnalysis are third-party packages
that are neither vendored in the reference nor installed here, so this module
provides (a) a minimal PDB reader exposing the handful of Universe attributes
the reference touches, (b) a k-nearest-neighbour graph builder with the
reference's padding convention (index 0 / distance 0), and (c) batching of
independent graphs into one concatenated tuple for the CUDA path.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

# element -> one-hot column.  N/C/H are pinned by the non-zero entries of the
# reference's baked peak_std/peak_avg constants (15N, 13C, 1H shift statistics at
# columns 2, 3, 4); the remaining columns follow nmrdata's embedding order as far
# as it could be inferred ("parity unpinned" for this mapping —
# those elements have peak_std == 0, so their predicted peak is exactly 0 and
# they only act as neighbours through the embedding row they select).
ELEMENT_INDEX = {"X": 0, "Z": 1, "N": 2, "C": 3, "H": 4, "O": 5, "S": 6, "P": 7, "F": 8, "CL": 9}


@dataclass
class AtomGroup:
    positions: np.ndarray          # [N,3] Angstrom, current frame
    elements: np.ndarray           # [N] str
    names: np.ndarray              # [N] str
    resnames: np.ndarray           # [N] str
    resids: np.ndarray             # [N] int

    def __len__(self) -> int:
        return len(self.elements)


@dataclass
class Timestep:
    frame: int
    time: float


class Universe:
    """Tiny stand-in for ``MDAnalysis.Universe`` covering what nmrgnn uses:
    ``u.atoms.{positions,elements,names,resnames,resids}`` and iteration over
    ``u.trajectory`` (nmrgnn/main.py:220-259)."""

    def __init__(self, frames: np.ndarray, elements, names, resnames, resids, dt: float = 1.0):
        self._frames = np.asarray(frames, np.float32)
        self.atoms = AtomGroup(self._frames[0].copy(), np.asarray(elements), np.asarray(names),
                               np.asarray(resnames), np.asarray(resids, np.int64))
        self._dt = dt
        self.trajectory = _Trajectory(self)

    @property
    def n_frames(self) -> int:
        return self._frames.shape[0]


class _Trajectory:
    def __init__(self, u: Universe):
        self._u = u

    def __len__(self) -> int:
        return self._u.n_frames

    def _seek(self, i: int) -> Timestep:
        self._u.atoms.positions = self._u._frames[i].copy()
        return Timestep(frame=i, time=i * self._u._dt)

    def __getitem__(self, item):
        if isinstance(item, slice):
            return (self._seek(i) for i in range(*item.indices(len(self))))
        return self._seek(range(len(self))[item])

    def __iter__(self):
        return (self._seek(i) for i in range(len(self)))


def _guess_element(name: str) -> str:
    s = "".join(ch for ch in name if ch.isalpha()).upper()
    if not s:
        return "X"
    return "CL" if s.startswith("CL") else s[0]


def read_pdb(path: str, include_hetatm: bool = True) -> Universe:
    """ATOM/HETATM records of a (possibly gzipped, possibly multi-MODEL) PDB file."""
    opener = gzip.open if str(path).endswith(".gz") else open
    frames: List[List[Tuple[float, float, float]]] = []
    cur: List[Tuple[float, float, float]] = []
    elements: List[str] = []
    names: List[str] = []
    resnames: List[str] = []
    resids: List[int] = []
    first = True
    with opener(path, "rt") as f:
        for line in f:
            rec = line[:6]
            if rec == "ATOM  " or (include_hetatm and rec == "HETATM"):
                cur.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                if first:
                    name = line[12:16].strip()
                    el = line[76:78].strip().upper() if len(line) >= 78 else ""
                    elements.append(el or _guess_element(name))
                    names.append(name)
                    resnames.append(line[17:20].strip())
                    try:
                        resids.append(int(line[22:26]))
                    except ValueError:
                        resids.append(0)
            elif rec.startswith("ENDMDL"):
                if cur:
                    frames.append(cur)
                    cur = []
                    first = False
    if cur:
        frames.append(cur)
    if not frames:
        raise ValueError(f"{path}: no ATOM records")
    n = len(frames[0])
    if any(len(fr) != n for fr in frames):
        raise ValueError(f"{path}: MODELs have different atom counts")
    return Universe(np.asarray(frames, np.float32), elements, names, resnames, resids)


class UnpinnedElementWarning(UserWarning):
    """The one-hot column of this element is not pinned against nmrdata's embedding table (third-party, absent)."""


_PINNED = {"N", "C", "H"}          # columns 2, 3, 4: fixed by the baked peak_std / peak_avg constants of the SavedModel
_warned: set = set()


def one_hot_elements(elements: Iterable[str], num_elem: int = 10, warn: bool = True) -> np.ndarray:
    """One-hot rows for element symbols.  Columns of N / C / H are pinned by the reference's artefact; the columns of
    every other element follow ELEMENT_INDEX, which could not be checked against ``nmrdata.load_embeddings()`` (the
    package is neither vendored nor installable here): the first use of such an element in a process emits an
    ``UnpinnedElementWarning``; symbols missing from the table map to column 0 ("X") with the same warning.  In the
    pretrained model columns 0 and 1 still hold their initial (untrained) embedding rows, columns 2-9 are trained."""
    import warnings
    syms = [str(e).upper() for e in elements]
    for e in (set(syms) - _PINNED - _warned) if warn else ():
        _warned.add(e)
        where = f"column {ELEMENT_INDEX[e]}" if e in ELEMENT_INDEX else "column 0 ('X': unknown element)"
        warnings.warn(f"element {e!r} -> one-hot {where}: this assignment is not pinned against nmrdata's embedding "
                      "table; only N, C, H are fixed by the reference's artefact", UnpinnedElementWarning, stacklevel=2)
    idx = np.array([ELEMENT_INDEX.get(e, 0) for e in syms], np.int64)
    idx = np.where(idx < num_elem, idx, 0)
    atoms = np.zeros((len(idx), num_elem), np.float32)
    atoms[np.arange(len(idx)), idx] = 1.0
    return atoms


def inv_degree_from_nlist(nlist: np.ndarray) -> np.ndarray:
    """``divide_no_nan(1, sum(nlist > 0))`` exactly as nmrgnn/library.py:115-116:
    a genuine neighbour whose index is 0 is *not* counted (kept on purpose —
    inv_degree is an input of the path)."""
    deg = np.sum(np.asarray(nlist) > 0, axis=1).astype(np.float32)
    return np.divide(np.float32(1), deg, out=np.zeros_like(deg), where=deg > 0)


def knn_graph_host(positions_nm: np.ndarray, neighbor_number: int = 16,
                   cutoff_nm: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Brute-force/KD-tree k-nearest-neighbour list on the host (float32 distances,
    neighbours sorted by distance, self excluded, missing slots = index 0 / edge 0).
    Reference for the CUDA builder in csrc/knn.cu; not on the timed path."""
    from scipy.spatial import cKDTree

    pos = np.asarray(positions_nm, np.float64)
    n = pos.shape[0]
    k = neighbor_number
    nlist = np.zeros((n, k), np.int32)
    edges = np.zeros((n, k), np.float32)
    if n <= 1:
        return nlist, edges
    kk = min(k + 1, n)
    tree = cKDTree(pos)
    dist, idx = tree.query(pos, k=kk)
    dist = dist.reshape(n, kk)
    idx = idx.reshape(n, kk)
    # drop self wherever it landed (coincident atoms can displace it from column 0)
    keep = idx != np.arange(n)[:, None]
    no_self = ~keep.all(axis=1)
    keep[~no_self, -1] = False                       # self not among the kk hits: drop the farthest
    order = np.argsort(~keep, axis=1, kind="stable")[:, :kk - 1]
    d = np.take_along_axis(dist, order, 1)
    j = np.take_along_axis(idx, order, 1)
    valid = np.isfinite(d)
    if cutoff_nm is not None:
        valid &= d <= cutoff_nm
    m = d.shape[1]
    nlist[:, :m] = np.where(valid, j, 0)
    edges[:, :m] = np.where(valid, d, 0.0).astype(np.float32)
    return nlist, edges


def build_graph(positions_A: np.ndarray, elements: Sequence[str], neighbor_number: int = 16,
                num_elem: int = 10, cutoff_nm: Optional[float] = None):
    """(atoms, nlist, edges, inv_degree) from Angstrom coordinates; distances are
    converted to nm, the unit the model's RBF grid (0.005-0.2) expects."""
    pos_nm = np.asarray(positions_A, np.float64) / 10.0
    nlist, edges = knn_graph_host(pos_nm, neighbor_number, cutoff_nm)
    atoms = one_hot_elements(elements, num_elem)
    return atoms, nlist, edges, inv_degree_from_nlist(nlist)


def batch_graphs(graphs: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]):
    """Concatenate independent graphs: nlist rows get their graph's atom offset
    added to *every* slot (a padded slot then points at the graph's first atom with
    edge 0, so it still contributes exactly 0), inv_degree is taken as given (it
    was computed before offsetting).  Returns (atoms, nlist, edges, inv_degree,
    graph_offsets[int64, G+1])."""
    offs = np.zeros(len(graphs) + 1, np.int64)
    for g, (a, _, _, _) in enumerate(graphs):
        offs[g + 1] = offs[g] + a.shape[0]
    atoms = np.concatenate([np.asarray(g[0], np.float32) for g in graphs], 0)
    nlist = np.concatenate([np.asarray(g[1]).astype(np.int64) + offs[i] for i, g in enumerate(graphs)], 0)
    if nlist.size and nlist.max() >= 2 ** 31:
        raise ValueError("batch too large for int32 neighbour indices")
    edges = np.concatenate([np.asarray(g[2], np.float32) for g in graphs], 0)
    inv = np.concatenate([np.asarray(g[3], np.float32).reshape(-1) for g in graphs], 0)
    return atoms, nlist.astype(np.int32), edges, inv, offs
