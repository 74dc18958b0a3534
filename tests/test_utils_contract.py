"""Behaviour of ``jaxabm_b200.utils`` pinned on what the reference's own unit tests require of
``jaxabm/utils.py`` (tests/unit/test_utils.py there; arrays are NumPy here, the engine has no JAX)."""
import numpy as np
import pytest

from jaxabm_b200 import utils


def test_convert_to_numpy_walks_containers():                      # utils.py:16-43
    out = utils.convert_to_numpy({"a": np.float32(1.5), "b": [np.arange(3), None], "c": (np.ones(2), "s"), "d": {"e": 2}})
    assert isinstance(out["a"], np.ndarray) or isinstance(out["a"], np.floating)
    assert isinstance(out["b"], list) and isinstance(out["b"][0], np.ndarray) and out["b"][1] is None
    assert isinstance(out["c"], tuple) and out["c"][1] == "s" and out["d"] == {"e": 2}
    assert utils.convert_to_numpy(None) is None


@pytest.mark.parametrize("params,required,ok", [
    ({"k1": 1, "k2": 2, "k3": 3}, ["k1", "k2"], True), ({"k1": 1, "k3": 3}, ["k1", "k2"], False),
    ({}, ["k1"], False), ({"k1": 1}, [], True), ({"k1": None, "k2": 2}, ["k1", "k2"], True)])
def test_is_valid_params(params, required, ok):                    # utils.py:46-57
    assert utils.is_valid_params(params, required) is ok


def test_format_time():                                            # utils.py:60-79
    assert utils.format_time(0) == "0.00s" and utils.format_time(0.001) == "0.00s" and utils.format_time(0.01) == "0.01s"
    assert utils.format_time(45.5) == "45.50s" and utils.format_time(125) == "2m 5.00s"
    assert utils.format_time(3723) == "1h 2m 3.00s"


def test_mean_over_runs():                                         # utils.py:82-107
    assert utils.mean_over_runs([]) == {}
    assert utils.mean_over_runs([{"m": [1, 2, 3]}, {"m": [4, 5, 6]}, {"m": [7, 8, 9]}]) == {"m": [4.0, 5.0, 6.0]}
    assert utils.mean_over_runs([{"m": [1, 2, 3], "n": [4, 5, 6]}]) == {"m": [1.0, 2.0, 3.0], "n": [4.0, 5.0, 6.0]}
    # only metrics present in every run, with one common length, survive
    assert utils.mean_over_runs([{"m": [1, 2], "n": [10, 20]}, {"m": [3, 4]}, {"m": [5, 6], "o": [1, 2]}]) == {"m": [3.0, 4.0]}
    assert utils.mean_over_runs([{"m": [1, 2, 3]}, {"m": [4, 5]}, {"m": [7, 8, 9]}]) == {}
    r = utils.mean_over_runs([{"m": [1.1, 2.2]}, {"m": [4.4, 5.5]}, {"m": [7.7, 8.8]}])
    assert r["m"] == pytest.approx([(1.1 + 4.4 + 7.7) / 3, (2.2 + 5.5 + 8.8) / 3], abs=1e-10)


def test_standardize_metrics():                                    # utils.py:110-125
    r = utils.standardize_metrics({"f": 1.5, "i": 42, "a": np.array(3.14), "s32": np.float32(2.5), "txt": "hello",
                                   "lst": [1, 2, 3], "none": None})
    assert set(r) == {"f", "i", "a", "s32"} and all(type(v) is float for v in r.values())
    assert r["i"] == 42.0 and r["s32"] == 2.5 and r["a"] == pytest.approx(3.14)
    assert utils.standardize_metrics({}) == {}
