"""TensorFlow-free reader for TF "TensorBundle" checkpoints (``variables.index`` +
``variables.data-00000-of-00001``), the format the reference's pretrained model
and its ``ModelCheckpoint``/``model.save`` outputs use
(reference: nmrgnn/library.py:92-103 loads ``nmrgnn/models/baseline`` through
``tf.keras.models.load_model``; nmrgnn/main.py:63-68,82 write the same format).

The index file is a LevelDB-style SSTable with uncompressed blocks:

* footer = last 48 bytes: varint64 (offset,size) handles of the metaindex and
  index blocks, zero padding, 8-byte little-endian magic ``0xdb4775248b80fb57``;
* block = prefix-compressed entries
  ``varint shared | varint non_shared | varint value_len | key_suffix | value``
  followed by ``uint32 restarts[n], uint32 n``; every block is followed on disk
  by a 5-byte trailer (compression type + crc32c);
* index-block values are handles (varint offset, varint size) of data blocks;
* key ``""`` holds a ``BundleHeaderProto``; every other key holds a
  ``BundleEntryProto`` {1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset,
  5: size, 6: crc32c}.  Tensor bytes live in the data shard at [offset,
  offset+size), little-endian, row-major.

Only what inference needs is implemented: float32/float64/int32/int64 tensors,
one or more shards, no slices, no compression.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_FOOTER_LEN = 48

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}
DT_STRING = 7


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Minimal protobuf wire-format walker: yields (field, wire_type, value)."""
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims: List[int] = []
    for field, _, val in _proto_fields(buf):
        if field == 2:  # repeated Dim
            size = 0
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


@dataclass(frozen=True)
class BundleEntry:
    key: str
    dtype: int
    shape: Tuple[int, ...]
    shard_id: int
    offset: int
    size: int
    crc32c: int


def _parse_entry(key: str, buf: bytes) -> BundleEntry:
    dtype = 0
    shape: Tuple[int, ...] = ()
    shard = offset = size = crc = 0
    for field, wt, val in _proto_fields(buf):
        if field == 1:
            dtype = val
        elif field == 2:
            shape = _parse_shape(val)
        elif field == 3:
            shard = val
        elif field == 4:
            offset = val
        elif field == 5:
            size = val
        elif field == 6 and wt == 5:
            crc = struct.unpack("<I", val)[0]
    return BundleEntry(key, dtype, shape, shard, offset, size, crc)


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        val = block[pos:pos + vlen]
        pos += vlen
        yield key, val


_CRC_TABLE = None


def crc32c(data: bytes) -> int:
    """Castagnoli CRC (table-driven, numpy-free; used only in tests/verify)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = tbl
    c = 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


class TensorBundle:
    """Random-access view of one checkpoint prefix (e.g. ``.../variables/variables``)."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        index_path = prefix + ".index"
        with open(index_path, "rb") as f:
            raw = f.read()
        if len(raw) < _FOOTER_LEN:
            raise ValueError(f"{index_path}: too short to be a TensorBundle index")
        footer = raw[-_FOOTER_LEN:]
        if struct.unpack("<Q", footer[-8:])[0] != _MAGIC:
            raise ValueError(f"{index_path}: bad SSTable magic")
        pos = 0
        _, pos = _varint(footer, pos)  # metaindex offset
        _, pos = _varint(footer, pos)  # metaindex size
        idx_off, pos = _varint(footer, pos)
        idx_size, pos = _varint(footer, pos)
        self.entries: Dict[str, BundleEntry] = {}
        self.num_shards = 1
        for _, handle in _block_entries(raw[idx_off:idx_off + idx_size]):
            off, p = _varint(handle, 0)
            size, p = _varint(handle, p)
            if raw[off + size] != 0:
                raise ValueError("compressed TensorBundle index blocks are not supported")
            for key, val in _block_entries(raw[off:off + size]):
                if key == b"":
                    for field, _, v in _proto_fields(val):
                        if field == 1:
                            self.num_shards = v
                        elif field == 2 and v != 0:
                            raise ValueError("big-endian bundles are not supported")
                    continue
                k = key.decode("utf-8")
                self.entries[k] = _parse_entry(k, val)

    def keys(self) -> List[str]:
        return list(self.entries)

    def _shard_path(self, shard: int) -> str:
        return f"{self.prefix}.data-{shard:05d}-of-{self.num_shards:05d}"

    def read_raw(self, key: str) -> bytes:
        e = self.entries[key]
        with open(self._shard_path(e.shard_id), "rb") as f:
            f.seek(e.offset)
            data = f.read(e.size)
        if len(data) != e.size:
            raise ValueError(f"{key}: truncated data shard")
        return data

    def read(self, key: str, verify_crc: bool = False) -> np.ndarray:
        e = self.entries[key]
        if e.dtype not in _DTYPES:
            raise TypeError(f"{key}: unsupported dtype enum {e.dtype}")
        data = self.read_raw(key)
        if verify_crc and masked_crc32c(data) != e.crc32c:
            raise ValueError(f"{key}: crc32c mismatch")
        return np.frombuffer(data, dtype=_DTYPES[e.dtype]).reshape(e.shape).copy()

    def object_graph_names(self) -> Dict[str, str]:
        """checkpoint key -> Keras variable full_name, decoded from the
        ``_CHECKPOINTABLE_OBJECT_GRAPH`` string tensor (a TrackableObjectGraph proto)."""
        key = "_CHECKPOINTABLE_OBJECT_GRAPH"
        if key not in self.entries:
            return {}
        raw = self.read_raw(key)
        ln, pos = _varint(raw, 0)  # string tensor: varint length, 4-byte crc, bytes
        proto = raw[pos + 4:pos + 4 + ln]
        out: Dict[str, str] = {}
        for field, _, node in _proto_fields(proto):
            if field != 1:
                continue
            for f2, _, attr in _proto_fields(node):
                if f2 != 2:  # SerializedTensor attributes
                    continue
                full_name = ckpt_key = None
                for f3, _, v3 in _proto_fields(attr):
                    if f3 == 2:
                        full_name = v3.decode()
                    elif f3 == 3:
                        ckpt_key = v3.decode()
                if ckpt_key:
                    out[ckpt_key] = full_name or ""
        return out


_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def resolve_prefix(path: str) -> str:
    """Accept a SavedModel dir, a ``variables`` dir, or a checkpoint prefix."""
    path = os.fspath(path)
    cands = [path, os.path.join(path, "variables", "variables"), os.path.join(path, "variables")]
    for c in cands:
        if os.path.isfile(c + ".index"):
            return c
    raise FileNotFoundError(f"no TensorBundle index found under {path!r}")


def load_gnn_variables(path: str, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """Read the GNNModel inference tensors of a checkpoint written by the
    reference: ``out_layer/{kernel,bias}``,
    ``embed_layer/kernel`` and ``variables/<i>``, skipping optimizer slots and
    metric scalars.  Keys are returned without the ``/.ATTRIBUTES/...`` suffix."""
    tb = TensorBundle(resolve_prefix(path))
    out: Dict[str, np.ndarray] = {}
    for key in tb.keys():
        if not key.endswith(_SUFFIX) or ".OPTIMIZER_SLOT" in key:
            continue
        short = key[:-len(_SUFFIX)]
        if short.startswith(("out_layer/", "embed_layer/")) or (
                short.startswith("variables/") and short.split("/", 1)[1].isdigit()):
            out[short] = tb.read(key, verify_crc=verify_crc)
    return out
