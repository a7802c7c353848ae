"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every
symbol include/nmrgnn_b200.h declares; the host mirror fails loudly (no CPU fallback) without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from nmrgnn_b200 import build
    return build.build(force=False)


def declared_symbols():
    with open(os.path.join(ROOT, "include", "nmrgnn_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmrgnn_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from nmrgnn_b200 import _capi
    assert sorted(_capi.EXPORTS) == declared_symbols()


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.nmrgnn_abi_version.restype = ctypes.c_int
    assert lib.nmrgnn_abi_version() == 1


def test_every_runtime_option_is_documented_in_the_header():
    """nmrgnn_set_option: every name the library accepts is described in include/nmrgnn_b200.h (and nothing else is)."""
    with open(os.path.join(ROOT, "nmrgnn_b200", "csrc", "api.cu")) as f:
        src = f.read()
    body = src[src.index("int nmrgnn_set_option("):]
    body = body[:body.index("\n}\n")]
    accepted = set(re.findall(r'std::strcmp\(name, "([a-z0-9_]+)"\)', body))
    assert {"profile", "force_ffma", "tc_min_atoms", "edge_table", "tc_compensate"} <= accepted
    with open(os.path.join(ROOT, "include", "nmrgnn_b200.h")) as f:
        header = f.read()
    doc = header[header.index("Runtime options"):header.index("int nmrgnn_set_option(")]
    documented = set(re.findall(r'"([a-z0-9_]+)"', doc))
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))


def test_num_weights_and_null_handling(lib_path):
    from nmrgnn_b200 import _capi
    lib = _capi.load_library()
    d = _capi.Dims(10, 256, 3, 128, 4, 4, 4, 1, 1, 0.005, 0.2)
    assert lib.nmrgnn_num_weights(ctypes.byref(d)) == 2 * 4 + 1 + 4 + 2 * 4 + 2 + 2
    assert lib.nmrgnn_num_weights(None) == _capi.ERR_BAD_DIMS
    # null handle: status code, no crash
    assert lib.nmrgnn_synchronize(None, None) == _capi.ERR_BAD_DIMS
    assert lib.nmrgnn_kernel_launches(None) == 0
    lib.nmrgnn_destroy(None)


def test_create_fails_loudly_without_device(lib_path):
    """No CPU fallback: without an sm_100 device model creation raises with NMRGNN_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import nmrgnn_b200
    from nmrgnn_b200 import _capi
    with pytest.raises(_capi.NmrgnnError) as ei:
        nmrgnn_b200.load_model()
    assert ei.value.code == _capi.ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "nmrgnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
