"""ctypes binding of libnmrgnn_b200.so (include/nmrgnn_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is
visible, loading / model creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnmrgnn_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
OK, ERR_BAD_DIMS, ERR_BAD_INDEX, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE, ERR_COMM = 0, -1, -2, -3, -4, -5, -6
COMM_HANDLE_BYTES = 128

EXPORTS = [
    "nmrgnn_abi_version", "nmrgnn_num_weights", "nmrgnn_create", "nmrgnn_destroy", "nmrgnn_forward",
    "nmrgnn_edge_features", "nmrgnn_embed", "nmrgnn_mp_layer", "nmrgnn_fc_readout", "nmrgnn_synchronize",
    "nmrgnn_kernel_launches", "nmrgnn_compute_path", "nmrgnn_last_error", "nmrgnn_knn_graph",
    "nmrgnn_set_option", "nmrgnn_selftest_gemm", "nmrgnn_stage_times", "nmrgnn_tc_compensation", "nmrgnn_edge_table_info",
    "nmrgnn_comm_local", "nmrgnn_comm_init", "nmrgnn_forward_sharded", "nmrgnn_comm_buffer", "nmrgnn_comm_destroy",
]


class Dims(C.Structure):
    _fields_ = [("num_elem", C.c_int32), ("atom_features", C.c_int32), ("edge_features", C.c_int32),
                ("edge_hidden", C.c_int32), ("n_edge_fc", C.c_int32), ("n_mp", C.c_int32), ("n_fc", C.c_int32),
                ("mp_activation", C.c_int32), ("fc_activation", C.c_int32), ("rbf_low", C.c_float),
                ("rbf_high", C.c_float)]


class NmrgnnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"nmrgnn_b200 error {code}: {message}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). nmrgnn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, fp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.nmrgnn_abi_version.restype = C.c_int
    lib.nmrgnn_num_weights.argtypes = [C.POINTER(Dims)]
    lib.nmrgnn_create.argtypes = [C.POINTER(Dims), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(vp)]
    lib.nmrgnn_destroy.argtypes = [vp]
    lib.nmrgnn_destroy.restype = None
    lib.nmrgnn_forward.argtypes = [vp, fp, fp, fp, fp, i64, i32, fp, C.c_int, vp]
    lib.nmrgnn_edge_features.argtypes = [vp, fp, i64, fp, C.c_int, vp]
    lib.nmrgnn_embed.argtypes = [vp, fp, i64, fp, C.c_int, vp]
    lib.nmrgnn_mp_layer.argtypes = [vp, i32, fp, fp, fp, fp, i64, i32, fp, C.c_int, vp]
    lib.nmrgnn_fc_readout.argtypes = [vp, fp, fp, i64, fp, fp, C.c_int, vp]
    lib.nmrgnn_synchronize.argtypes = [vp, vp]
    lib.nmrgnn_kernel_launches.argtypes = [vp]
    lib.nmrgnn_kernel_launches.restype = C.c_int64
    lib.nmrgnn_compute_path.argtypes = [vp]
    lib.nmrgnn_compute_path.restype = C.c_char_p
    lib.nmrgnn_last_error.argtypes = [vp]
    lib.nmrgnn_last_error.restype = C.c_char_p
    lib.nmrgnn_knn_graph.argtypes = [vp, fp, fp, i64, i64, i32, C.c_float, fp, fp, fp, C.c_int, vp]
    lib.nmrgnn_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.nmrgnn_selftest_gemm.argtypes = [vp, fp, fp, fp, C.c_int]
    lib.nmrgnn_stage_times.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.nmrgnn_tc_compensation.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.nmrgnn_edge_table_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.nmrgnn_comm_local.argtypes = [vp, i64, i32, vp]
    lib.nmrgnn_comm_init.argtypes = [vp, i32, i32, vp]
    lib.nmrgnn_forward_sharded.argtypes = [vp, fp, fp, fp, fp, i64, i32, fp, C.c_int, vp]
    lib.nmrgnn_comm_buffer.argtypes = [vp, C.POINTER(i64)]
    lib.nmrgnn_comm_buffer.restype = C.c_void_p
    lib.nmrgnn_comm_destroy.argtypes = [vp]
    lib.nmrgnn_comm_destroy.restype = None
    if lib.nmrgnn_abi_version() != 1:
        raise ImportError("libnmrgnn_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr(x) -> Optional[int]:
    """Raw address of a NumPy array / torch tensor / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


class Handle:
    """Owns one nmrgnn_handle (one model on one GPU)."""

    def __init__(self, dims: Dims, weights: Sequence[np.ndarray], device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        arrs = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        rc = self._lib.nmrgnn_create(C.byref(dims), ptrs, len(arrs), device, C.byref(self._h))
        if rc != OK:
            raise NmrgnnError(rc, (self._lib.nmrgnn_last_error(None) or b"").decode())
        self.dims = dims
        self.device = device

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nmrgnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int) -> None:
        if rc == OK:
            return
        msg = (self._lib.nmrgnn_last_error(self._h) or b"").decode()
        if rc == ERR_BAD_INDEX:
            raise IndexError(msg)
        if rc == ERR_BAD_DIMS:
            raise ValueError(msg)
        if rc == ERR_OOM:
            raise MemoryError(msg)
        raise NmrgnnError(rc, msg)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.nmrgnn_kernel_launches(self._h))

    @property
    def compute_path(self) -> str:
        return (self._lib.nmrgnn_compute_path(self._h) or b"").decode()

    def set_option(self, name: str, value: int) -> None:
        self.check(self._lib.nmrgnn_set_option(self._h, name.encode(), int(value)))

    def stage_times(self) -> dict:
        """Device ms per stage of the last forward run with option "profile" = 1."""
        buf = (C.c_float * 80)()
        n = self._lib.nmrgnn_stage_times(self._h, buf, 80)
        if n < 0:
            self.check(n)
        ms = [float(buf[i]) for i in range(n)]
        return {"edge": ms[0], "embed": ms[1], "mp_layers": ms[2:n - 1], "fc_readout": ms[n - 1]}

    def tc_compensation(self) -> dict:
        """Round-toward-zero compensation constants of the tensor-core path, in units of 2^-24."""
        buf = (C.c_float * 80)()
        n = self._lib.nmrgnn_tc_compensation(self._h, buf, 80)
        if n < 0:
            self.check(n)
        return {"edge": float(buf[0]), "mp_layers": [float(buf[i]) for i in range(1, n)]}

    # ---- multi-GPU reassembly over peer memory (nmrgnn_comm_*, see include/nmrgnn_b200.h)
    def comm_local(self, capacity: int, world: int) -> bytes:
        blob = C.create_string_buffer(COMM_HANDLE_BYTES)
        self.check(self._lib.nmrgnn_comm_local(self._h, int(capacity), int(world), blob))
        return blob.raw

    def comm_init(self, rank: int, world: int, blobs: Sequence[bytes]) -> None:
        buf = C.create_string_buffer(b"".join(blobs), COMM_HANDLE_BYTES * int(world))
        self.check(self._lib.nmrgnn_comm_init(self._h, int(rank), int(world), buf))

    def forward_sharded(self, atoms, nlist, edges, inv_degree, n_local, k, gathered, mem, stream=None):
        self.check(self._lib.nmrgnn_forward_sharded(self._h, _ptr(atoms), _ptr(nlist), _ptr(edges), _ptr(inv_degree),
                                                    int(n_local), int(k), _ptr(gathered), mem, stream))

    def comm_buffer(self):
        """(device pointer, capacity) of the gathered peaks [world][capacity] of the latest forward_sharded call."""
        cap = C.c_int64(0)
        ptr = self._lib.nmrgnn_comm_buffer(self._h, C.byref(cap))
        return ptr, int(cap.value)

    def comm_destroy(self) -> None:
        if self._h:
            self._lib.nmrgnn_comm_destroy(self._h)

    def edge_table_info(self) -> dict:
        """Create-time table of the edge block: {'active', 'intervals', 'rel_error'} (nmrgnn_edge_table_info)."""
        n = C.c_int32(0)
        err = C.c_double(0.0)
        rc = self._lib.nmrgnn_edge_table_info(self._h, C.byref(n), C.byref(err))
        if rc < 0:
            self.check(rc)
        return {"active": bool(rc), "intervals": int(n.value), "rel_error": float(err.value)}

    def synchronize(self, stream: Optional[int] = None) -> None:
        self.check(self._lib.nmrgnn_synchronize(self._h, stream))

    # thin wrappers: pointers are raw addresses (NumPy host arrays or torch CUDA tensors)
    def forward(self, atoms, nlist, edges, inv_degree, n_atoms, k, peaks, mem, stream=None):
        self.check(self._lib.nmrgnn_forward(self._h, _ptr(atoms), _ptr(nlist), _ptr(edges), _ptr(inv_degree),
                                            n_atoms, k, _ptr(peaks), mem, stream))

    def edge_features(self, edges, n_edges, out, mem, stream=None):
        self.check(self._lib.nmrgnn_edge_features(self._h, _ptr(edges), n_edges, _ptr(out), mem, stream))

    def embed(self, atoms, n_atoms, out, mem, stream=None):
        self.check(self._lib.nmrgnn_embed(self._h, _ptr(atoms), n_atoms, _ptr(out), mem, stream))

    def mp_layer(self, layer, nodes_in, nlist, efeat, inv_degree, n_atoms, k, nodes_out, mem, stream=None):
        self.check(self._lib.nmrgnn_mp_layer(self._h, layer, _ptr(nodes_in), _ptr(nlist), _ptr(efeat),
                                             _ptr(inv_degree), n_atoms, k, _ptr(nodes_out), mem, stream))

    def fc_readout(self, nodes, atoms, n_atoms, peaks, fc_nodes, mem, stream=None):
        self.check(self._lib.nmrgnn_fc_readout(self._h, _ptr(nodes), _ptr(atoms), n_atoms, _ptr(peaks),
                                               _ptr(fc_nodes), mem, stream))

    def selftest_gemm(self, A: np.ndarray, W: np.ndarray, mode: int = 0) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        assert A.shape == (128, 64) and W.shape == (64, 128)
        D = np.empty((128, 128), np.float32)
        self.check(self._lib.nmrgnn_selftest_gemm(self._h, _ptr(A), _ptr(W), _ptr(D), mode))
        return D

    def knn_graph(self, positions, graph_offsets, n_atoms, n_graphs, k, cutoff, nlist, edges, inv_degree, mem,
                  stream=None):
        self.check(self._lib.nmrgnn_knn_graph(self._h, _ptr(positions), _ptr(graph_offsets), n_atoms, n_graphs, k,
                                              cutoff, _ptr(nlist), _ptr(edges), _ptr(inv_degree), mem, stream))
