"""
GPU parity tests for the rows next to the hot path:
Profile.split and `kpal showbalance` (reference tests/test_klib.py:182-217,
tests/test_kmer.py:195-202) and ProfileDistance with do_positive
(reference kpal/kdistlib.py:143-145) -- device results against vectors made by
the unmodified reference (tests/golden/golden_split.json) and against the oracle
at sizes the reference's Python loops cannot reach.
"""
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, dense
from kpal_b200 import _cabi, kdistlib, klib, kmer, metrics
from oracle import kpal_oracle as ko
from test_host_api import FakeH5

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def golden_split():
    with open(os.path.join(GOLDEN_DIR, "golden_split.json")) as f:
        return json.load(f)


def test_split_golden_bit_exact(golden_split):
    for case in golden_split["split_cases"]:
        counts = dense(case["counts"])
        forward, reverse = klib.Profile(counts.copy()).split()
        assert forward.dtype == np.int64 and reverse.dtype == np.int64
        assert forward.tolist() == case["forward"] and reverse.tolist() == case["reverse"]
        got = _cabi.show_balance(counts)
        assert got == pytest.approx(case["showbalance"], rel=RTOL)
        if case["showbalance"] == 0.0:
            assert got == 0.0


def test_split_large_k_against_oracle():
    rng = np.random.default_rng(31)
    for k in (8, 9, 11, 12):                  # even and odd k: with and without palindromes
        counts = rng.poisson(0.8, 4 ** k).astype(np.int64)
        counts[rng.integers(0, 4 ** k, 50)] = rng.integers(1 << 20, 1 << 40, 50)
        forward, reverse = _cabi.split(counts)
        want_f, want_r = ko.split(counts)
        assert forward.size == (4 ** k + (2 ** k if k % 2 == 0 else 0)) // 2
        assert np.array_equal(forward, want_f) and np.array_equal(reverse, want_r)
        small = rng.poisson(0.8, 4 ** k).astype(np.int64)
        assert _cabi.show_balance(small) == pytest.approx(ko.show_balance(small), rel=RTOL)


def test_showbalance_command_golden(golden, golden_split):
    """reference tests/test_kmer.py:195-202: '1 0.669'."""
    counts = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    store = FakeH5()
    klib.Profile(counts).save(store)                        # unnamed -> '1'
    out = io.StringIO()
    kmer.get_balance(store, out, precision=3)
    assert out.getvalue() == '1 0.669\n'
    assert _cabi.show_balance(counts) == pytest.approx(golden_split["showbalance_sequences_k8"], rel=RTOL)


def test_positive_pair_distance_golden(golden_split):
    fns = {"euclidean": metrics.euclidean, "cosine": metrics.cosine_similarity}
    for case in golden_split["positive_cases"]:
        left, right = dense(case["left"]), dense(case["right"])
        for key, want in case["values"].items():
            bal, sc, dn, metric = key.split("_", 3)
            dist = kdistlib.ProfileDistance(
                do_balance=bal == "bal1", do_positive=True, do_scale=sc == "sc1", down=dn == "dn1",
                distance_function=fns.get(metric),
                pairwise=metrics.pairwise["sum" if metric == "multiset-sum" else "prod"])
            a, b = klib.Profile(left.copy()), klib.Profile(right.copy())
            got = dist.distance(a, b)
            assert got == pytest.approx(want, rel=RTOL, nan_ok=True), key
            assert np.array_equal(a.counts, left) and np.array_equal(b.counts, right)


def test_positive_matrix_goes_pair_by_pair():
    rng = np.random.default_rng(6)
    profiles = [klib.Profile(rng.poisson(lam, 4 ** 5).astype(np.int64), 'p%d' % i)
                for i, lam in enumerate((0.4, 0.9, 1.6, 0.2))]
    out = io.StringIO()
    kdistlib.distance_matrix(profiles, out, 9, kdistlib.ProfileDistance(do_positive=True, do_scale=True))
    lines = out.getvalue().split('\n')
    assert lines[:5] == ['4', 'p0', 'p1', 'p2', 'p3']
    for i in range(1, 4):
        got = [float(x) for x in lines[4 + i].split(' ')]
        want = [ko.distance(profiles[i].counts, profiles[j].counts, do_scale=True, do_positive=True)
                for j in range(i)]
        np.testing.assert_allclose(got, want, rtol=0, atol=0.6e-9)
