"""CPU tests for the TensorFlow-free checkpoint reader and the params container."""
import os
import struct

import numpy as np
import pytest

from nmrgnn_b200.params import GNNParams, baseline_path, baseline_standards, rbf_centers
from nmrgnn_b200 import tensorbundle as tbm

REF = "/root/reference/nmrgnn/models/baseline"

# (key, shape, data offset)
TABLE = [
    ("out_layer/kernel", (128, 10), 24), ("out_layer/bias", (10,), 5144), ("embed_layer/kernel", (10, 256), 5184),
    ("variables/0", (128, 128), 15424), ("variables/1", (128,), 80960), ("variables/6", (128, 3), 213568),
    ("variables/7", (3,), 215104), ("variables/8", (256, 256, 3), 215116), ("variables/11", (256, 256, 3), 2574412),
    ("variables/12", (256, 256), 3360844), ("variables/18", (256, 128), 4150348), ("variables/19", (128,), 4281420),
]


def _varint(n):
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _write_bundle(prefix, tensors):
    """Minimal TensorBundle writer (one data block, one shard) for round-trip tests."""
    data = b""
    entries = []
    for key in sorted(tensors):
        arr = np.ascontiguousarray(tensors[key])
        raw = arr.tobytes()
        shape = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(s) for s in arr.shape))
        dtype = {np.dtype("float32"): 1, np.dtype("int64"): 9}[arr.dtype]
        proto = b"\x08" + _varint(dtype) + b"\x12" + _varint(len(shape)) + shape
        if len(data):
            proto += b"\x20" + _varint(len(data))
        proto += b"\x28" + _varint(len(raw)) + b"\x35" + struct.pack("<I", tbm.masked_crc32c(raw))
        entries.append((key.encode(), proto))
        data += raw
    header = b"\x08\x01"  # num_shards = 1
    block = b""
    for k, v in [(b"", header)] + entries:
        block += _varint(0) + _varint(len(k)) + _varint(len(v)) + k + v
    block += struct.pack("<II", 0, 1)
    trailer = b"\x00" + b"\x00\x00\x00\x00"
    meta = struct.pack("<II", 0, 1)
    off_meta = len(block) + 5
    handle = _varint(0) + _varint(len(block))
    index = _varint(0) + _varint(1) + _varint(len(handle)) + b"~" + handle + struct.pack("<II", 0, 1)
    off_index = off_meta + len(meta) + 5
    footer = _varint(off_meta) + _varint(len(meta)) + _varint(off_index) + _varint(len(index))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    with open(prefix + ".index", "wb") as f:
        f.write(block + trailer + meta + trailer + index + trailer + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)


def test_roundtrip_synthetic_bundle(tmp_path):
    p = GNNParams.random(num_elem=7, atom_feature_size=32, edge_feature_size=2, edge_hidden_size=16,
                         mp_layers=2, fc_layers=3, edge_fc_layers=3, seed=3)
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    tensors = {"out_layer/kernel" + suffix: p.out[0], "out_layer/bias" + suffix: p.out[1],
               "embed_layer/kernel" + suffix: p.embed, "optimizer/iter" + suffix: np.array(5, np.int64)}
    seq = [a for Wb in p.edge_fc for a in Wb] + list(p.mp_w) + [a for Wb in p.fc for a in Wb]
    for i, a in enumerate(seq):
        tensors[f"variables/{i}" + suffix] = a
        tensors[f"variables/{i}/.OPTIMIZER_SLOT/optimizer/m" + suffix] = np.zeros_like(a)
    os.makedirs(tmp_path / "variables")
    _write_bundle(str(tmp_path / "variables" / "variables"), tensors)
    # a bare checkpoint does not say which activations / RBF range the model was trained with: no silent defaults
    with pytest.raises(ValueError, match="mp_activation"):
        GNNParams.from_tf_checkpoint(str(tmp_path), peak_std=p.peak_std, peak_avg=p.peak_avg)
    q = GNNParams.from_tf_checkpoint(str(tmp_path), peak_std=p.peak_std, peak_avg=p.peak_avg, mp_activation="softplus",
                                     fc_activation="softplus", rbf_low=0.005, rbf_high=0.2)
    assert len(q.edge_fc) == 3 and len(q.mp_w) == 2 and len(q.fc) == 3
    for (a, b), (c, d) in zip(p.edge_fc + p.fc, q.edge_fc + q.fc):
        assert np.array_equal(a, c) and np.array_equal(b, d)
    for a, c in zip(p.mp_w, q.mp_w):
        assert np.array_equal(a, c)
    v = tbm.load_gnn_variables(str(tmp_path), verify_crc=True)
    assert "optimizer/iter" not in v


def _pb(field, payload, wt=2):
    """one protobuf field: length-delimited bytes (wt 2) or varint (wt 0)"""
    def varint(x):
        out = bytearray()
        while True:
            b = x & 0x7F
            x >>= 7
            out.append(b | (0x80 if x else 0))
            if not x:
                return bytes(out)
    if wt == 0:
        return varint(field << 3) + varint(payload)
    return varint((field << 3) | 2) + varint(len(payload)) + payload


def test_savedmodel_metadata_overrides_defaults(tmp_path):
    """A reference-trained SavedModel with relu / tanh, another RBF range and other standards loads with exactly
    those (the loader used to assume softplus / 0.005-0.20 / the baseline standards)."""
    import json
    p = GNNParams.random(num_elem=7, atom_feature_size=32, edge_feature_size=2, edge_hidden_size=16, mp_layers=2,
                         fc_layers=3, edge_fc_layers=3, seed=4)
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    tensors = {"out_layer/kernel" + suffix: p.out[0], "out_layer/bias" + suffix: p.out[1],
               "embed_layer/kernel" + suffix: p.embed}
    seq = [a for Wb in p.edge_fc for a in Wb] + list(p.mp_w) + [a for Wb in p.fc for a in Wb]
    for i, a in enumerate(seq):
        tensors[f"variables/{i}" + suffix] = a
    os.makedirs(tmp_path / "variables")
    _write_bundle(str(tmp_path / "variables" / "variables"), tensors)
    hyp = {"class_name": "HyperParameters", "config": {"space": [], "values": {
        "mp_activation": "tanh", "fc_activation": "relu", "rbf_low": 0.01, "rbf_high": 0.3}}}
    metas = [{"class_name": "MPBlock", "name": "mp-block", "config": {"name": "mp-block", "hypers": hyp}},
             {"class_name": "RBFExpansion", "name": "rbf-layer", "config": {"low": 0.01, "high": 0.3, "count": 16}}]
    nodes = b"".join(_pb(1, _pb(4, _pb(1, b"_tf_keras_layer") + _pb(3, json.dumps(m).encode()))) for m in metas)
    std = np.arange(1, 8, dtype=np.float32)
    avg = np.arange(7, dtype=np.float32) * 10

    def const(name, arr):
        tensor = _pb(1, 1, 0) + _pb(2, _pb(2, _pb(1, arr.size, 0))) + _pb(4, arr.astype("<f4").tobytes())
        attr = _pb(1, b"value") + _pb(2, _pb(8, tensor))
        return _pb(3, _pb(1, name.encode()) + _pb(2, b"Const") + _pb(5, attr))
    fn = _pb(1, const("gnn-model/mul_3/y", std) + const("gnn-model/mul_4/y", avg))
    meta_graph = _pb(2, _pb(2, fn)) + _pb(7, nodes)
    (tmp_path / "saved_model.pb").write_bytes(_pb(1, 1, 0) + _pb(2, meta_graph))
    q = GNNParams.from_tf_checkpoint(str(tmp_path))
    assert (q.mp_activation, q.fc_activation, q.rbf_low, q.rbf_high) == ("tanh", "relu", 0.01, 0.3)
    assert np.array_equal(q.peak_std, std) and np.array_equal(q.peak_avg, avg)
    assert q.meta["hypers_from"] == "saved_model.pb"
    # explicit arguments still win
    q = GNNParams.from_tf_checkpoint(str(tmp_path), mp_activation="softplus")
    assert q.mp_activation == "softplus" and q.fc_activation == "relu"


def test_bad_magic(tmp_path):
    (tmp_path / "x.index").write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError):
        tbm.TensorBundle(str(tmp_path / "x"))
    with pytest.raises(FileNotFoundError):
        tbm.resolve_prefix(str(tmp_path / "nothing"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkpoint not present on this box")
def test_reference_bundle_matches_survey_table_and_export():
    tb = tbm.TensorBundle(tbm.resolve_prefix(REF))
    assert len(tb.keys()) == 92
    for key, shape, off in TABLE:
        e = tb.entries[key + "/.ATTRIBUTES/VARIABLE_VALUE"]
        assert e.shape == shape and e.offset == off and e.dtype == 1
    v = tbm.load_gnn_variables(REF, verify_crc=True)
    assert sum(a.size for a in v.values()) == 1070477
    p = GNNParams.from_tf_checkpoint(REF)          # hypers and standards from the reference's saved_model.pb
    q = GNNParams.load(baseline_path())
    assert (p.mp_activation, p.fc_activation, p.rbf_low, p.rbf_high) == ("softplus", "softplus", 0.005, 0.2)
    assert np.array_equal(p.peak_std, q.peak_std) and np.array_equal(p.peak_avg, q.peak_avg)
    assert np.array_equal(p.mp_w[2], q.mp_w[2]) and np.array_equal(p.out[1], q.out[1])
    assert np.array_equal(p.edge_fc[0][0], q.edge_fc[0][0]) and np.array_equal(p.fc[3][0], q.fc[3][0])


def test_exported_baseline_shapes():
    p = GNNParams.load(baseline_path())
    assert (p.num_elem, p.atom_feature_size, p.edge_feature_size, p.edge_hidden_size) == (10, 256, 3, 128)
    assert len(p.edge_fc) == 4 and len(p.mp_w) == 4 and len(p.fc) == 4
    assert np.flatnonzero(p.out[1]).tolist() == [2, 3, 4]      # SURVEY §8c cross-check (iv)
    assert np.flatnonzero(p.peak_std).tolist() == [2, 3, 4]


def test_constants_bit_exact():
    std, avg = baseline_standards()
    assert std[3] == np.float32(50.94121551513672) and avg[4] == np.float32(5.630000114440918)
    c, gap = rbf_centers(0.005, 0.2, 128)
    assert c.dtype == np.float32 and gap.view(np.uint32) == 0x3AC94098
    assert c[0] == np.float32(0.005) and c[-1] == np.float32(0.2)
