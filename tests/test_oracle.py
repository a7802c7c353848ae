"""CPU tests: the NumPy oracle against the golden vectors produced by executing
the reference's traced SavedModel graph (tools/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err, scaled_err, tol_ratio
from oracle import forward as orc

CASES = ["g108m", "ring5_unit", "ring5_bonded", "prot300", "edge_cases64"]
BATCHED = ["smallmol12_k8", "prot3_batch"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("reference_order", [False, True])
def test_oracle_matches_traced_graph(baseline_params, name, reference_order):
    g = load_golden(name)
    y = orc.forward(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"],
                    reference_order=reference_order)
    assert y.dtype == np.float32
    # zeros (elements without shift statistics) must be exact
    assert np.array_equal(y == 0, g["peaks"] == 0)
    assert rel_err(y, g["peaks"]) < 5e-5          # fp32 vs fp32, different summation orders
    assert rel_err(y, g["peaks_f64"]) < 5e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_traced_graph_fp64(baseline_params, name):
    g = load_golden(name)
    y = orc.forward(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"], dtype=np.float64)
    assert rel_err(y, g["peaks_f64"]) < 1e-10


@pytest.mark.parametrize("name", ["prot300", "edge_cases64", "ring5_bonded"])
def test_oracle_intermediates(baseline_params, name):
    g = load_golden(name)
    inter = {}
    orc.forward(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"], intermediates=inter)
    np.testing.assert_allclose(inter["edge_features"], g["edge_features"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(inter["embed"], g["embed"], rtol=1e-6, atol=1e-7)
    # node features: error relative to the tensor's scale (single elements suffer cancellation)
    for l in range(4):
        assert scaled_err(inter["mp_nodes"][l], g[f"mp_nodes_{l}"]) < 2e-5
    assert scaled_err(inter["fc_nodes"], g["fc_nodes"]) < 2e-5
    # padded slots carry exactly-zero edge features
    assert np.all(inter["edge_features"][g["edges"] <= 0] == 0)


@pytest.mark.parametrize("name", BATCHED)
def test_oracle_batched_equals_per_graph(baseline_params, name):
    g = load_golden(name)
    # graphs are independent: one forward over the concatenated batch == per-graph forwards
    y = orc.forward(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"])
    assert tol_ratio(y, g["peaks"]) < 1.0 and tol_ratio(y, g["peaks_f64"]) < 1.0
    y2 = orc.forward_per_graph(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"],
                               g["graph_offsets"])
    assert tol_ratio(y2, g["peaks"]) < 1.0


def test_reference_weak_pins(baseline_params):
    """The reference's own value pins (tests/test_nmrgnn.py:236-243 + library.py:39-46):
    >= 75 % of 108M peaks within 2.5 sigma of their element mean."""
    g = load_golden("g108m")
    y = orc.forward(baseline_params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"])
    idx = np.argmax(g["atoms"], axis=1)
    std, avg = baseline_params.peak_std[idx], baseline_params.peak_avg[idx]
    ok = (std > 0) & ((y - avg) ** 2 <= (2.5 * std) ** 2)
    assert ok.mean() >= 0.75


def test_out_of_range_index_raises(baseline_params):
    g = load_golden("ring5_bonded")
    nl = g["nlist"].copy()
    nl[0, 0] = 5
    with pytest.raises(IndexError):
        orc.forward(baseline_params, g["atoms"], nl, g["edges"], g["inv_degree"])


def test_softplus_thresholds():
    x = np.array([-100, -20, -13.9, -1, 0, 1, 13.9, 20, 100], np.float32)
    y = orc.softplus(x)
    ref = np.log1p(np.exp(-np.abs(x.astype(np.float64)))) + np.maximum(x.astype(np.float64), 0)
    np.testing.assert_allclose(y, ref, rtol=2e-6, atol=1.2e-7)  # TF computes log(exp(x)+1): abs error ~eps/2
