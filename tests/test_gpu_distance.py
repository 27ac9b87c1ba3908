"""
GPU parity tests for the distance path (pytest -m gpu, B200 box): through
kpal_b200.kdistlib / the C ABI against the oracle and the reference's golden
values.  Tolerance: 1e-9 relative (north star) -- stated as RTOL below; the
goldens the reference pins exactly are asserted exactly.
Mirrors reference tests/test_kdistlib.py:38-134.
"""
import io

import numpy as np
import pytest

from conftest import dense
from kpal_b200 import _cabi, kdistlib, klib, metrics
from oracle import c_oracle, kpal_oracle as ko

pytestmark = pytest.mark.gpu

RTOL = 1e-9

METRIC_OPTS = {
    "multiset-prod": dict(metric="multiset", pairwise="prod"),
    "multiset-sum": dict(metric="multiset", pairwise="sum"),
    "euclidean": dict(metric="euclidean", pairwise="prod"),
    "cosine": dict(metric="cosine", pairwise="prod"),
}


def profile_distance(bal=False, sc=False, dn=False, metric="multiset-prod"):
    fn = {"euclidean": metrics.euclidean, "cosine": metrics.cosine_similarity}.get(metric)
    pw = metrics.pairwise["sum" if metric == "multiset-sum" else "prod"]
    return kdistlib.ProfileDistance(do_balance=bal, do_scale=sc, down=dn,
                                    distance_function=fn, pairwise=pw)


def test_distance_k2_exact(golden):
    g = golden["distance_k2"]
    a = klib.Profile(np.array(g["left"], dtype=np.int64))
    b = klib.Profile(np.array(g["right"], dtype=np.int64))
    assert kdistlib.ProfileDistance().distance(a, b) == 0.0625   # reference test_kdistlib.py:104-112


def test_distance_k8_goldens(golden):
    left = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8))
    right = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8))
    vals = golden["distance_k8"]["values"]
    np.testing.assert_almost_equal(kdistlib.ProfileDistance().distance(left, right), 0.4626209322)
    cases = {"default": profile_distance(), "balance": profile_distance(bal=True),
             "scale": profile_distance(sc=True), "scale_down": profile_distance(sc=True, dn=True),
             "sum": profile_distance(metric="multiset-sum"),
             "euclidean": profile_distance(metric="euclidean"),
             "cosine": profile_distance(metric="cosine")}
    for tag, dist in cases.items():
        assert dist.distance(left, right) == pytest.approx(vals[tag], rel=RTOL), tag


def test_distance_inputs_unmodified(golden):
    """reference test_kdistlib.py:124-134"""
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    a, b = klib.Profile(left.copy()), klib.Profile(right.copy())
    kdistlib.ProfileDistance(do_balance=True).distance(a, b)
    assert np.array_equal(a.counts, left) and np.array_equal(b.counts, right)


def test_distance_matrix_text_goldens(golden):
    """reference test_kdistlib.py:38-74"""
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    sets = {"1": [klib.Profile(left, "a")],
            "2": [klib.Profile(left, "a"), klib.Profile(right, "b")],
            "3": [klib.Profile(left, "a"), klib.Profile(right, "b"), klib.Profile(left, "c")]}
    for n, profiles in sets.items():
        out = io.StringIO()
        kdistlib.distance_matrix(profiles, out, 2, kdistlib.ProfileDistance())
        assert out.getvalue() == golden["matrix_text_k8_precision2"][n]
    assert out.getvalue().strip().split("\n") == ['3', 'a', 'b', 'c', '0.46', '0.00 0.46']


def test_synthetic_goldens_all_options(golden, golden_profiles):
    profiles = golden_profiles["profiles"]
    for key in golden_profiles.files:
        if key == "profiles":
            continue
        bal, sc, dn, metric = key.split("_", 3)
        got = _cabi.distance_matrix(profiles, do_balance=bal == "bal1", do_scale=sc == "sc1",
                                    down=dn == "dn1", **METRIC_OPTS[metric])
        want = golden_profiles[key]
        low = np.tril_indices(len(profiles), -1)
        np.testing.assert_allclose(got[low], want[low], rtol=RTOL, atol=0, err_msg=key)
        assert np.array_equal(got, got.T, equal_nan=True)
        if metric != "cosine":
            # identical profiles (2,5) -> exactly 0; exact multiple (1,6) -> 0 when scaled
            assert got[5, 2] == 0.0
            if sc == "sc1":
                assert got[6, 1] == 0.0
    plist = [klib.Profile(p.copy(), str(i)) for i, p in enumerate(profiles)]
    out = io.StringIO()
    kdistlib.distance_matrix(plist, out, 10, profile_distance(sc=True))
    want_lines = golden["synthetic_matrix_text_scaled_p10"].split("\n")
    got_lines = out.getvalue().split("\n")
    assert got_lines[:8] == want_lines[:8]
    for gl, wl in zip(got_lines[8:], want_lines[8:]):
        np.testing.assert_allclose([float(x) for x in gl.split()], [float(x) for x in wl.split()],
                                   rtol=0, atol=1.01e-10)


def test_tutorial_matrix(golden, tutorial_texts):
    """reference doc/tutorial.rst:128-144"""
    tut = golden["tutorial"]
    profs = {n: klib.Profile.from_fasta(io.StringIO(t), tut["k"], name=n)
             for n, t in tutorial_texts.items()}
    d = kdistlib.ProfileDistance().distance(profs["c_1"], profs["c_2"])
    assert d == pytest.approx(tut["distance_c_1_c_2"], rel=RTOL)
    assert "%.3f" % d == "0.456"
    merged = []
    for s in "abcd":
        m = profs[s + "_1"].copy()
        m.merge(profs[s + "_2"])
        m.name = "%s_1_%s_2" % (s, s)
        merged.append(m)
    out = io.StringIO()
    kdistlib.distance_matrix(merged, out, 3, kdistlib.ProfileDistance())
    assert out.getvalue() == tut["merged_matrix_text_p3"]


def synthetic_profiles(seed, n, k, lam_lo=0.3, lam_hi=8.0):
    rng = np.random.default_rng(seed)
    lam = np.exp(rng.uniform(np.log(lam_lo), np.log(lam_hi), n))
    return np.stack([rng.poisson(l, 4 ** k) for l in lam]).astype(np.int64)


@pytest.mark.parametrize("n,k", [(2, 3), (3, 1), (9, 6), (12, 8), (13, 5), (70, 6), (131, 7), (200, 4)])
def test_matrix_vs_oracle_all_options(n, k):
    """Small-N kernel (n <= 12) and tile kernel, ragged n, every option."""
    profiles = synthetic_profiles(100 * n + k, n, k)
    if n > 4:
        profiles[3] = profiles[1]                  # equal totals + identical profile
        profiles[4] = 2 * profiles[0]
    for bal in (False, True):
        for sc, dn in ((False, False), (True, False), (True, True)):
            for name, opts in METRIC_OPTS.items():
                got = _cabi.distance_matrix(profiles, do_balance=bal, do_scale=sc, down=dn, **opts)
                want = c_oracle.distance_matrix(profiles, do_balance=bal, do_scale=sc, down=dn,
                                                metric=opts["metric"], pairwise=opts["pairwise"],
                                                threads=c_oracle.max_threads())
                low = np.tril_indices(n, -1)
                np.testing.assert_allclose(got[low], want[low], rtol=RTOL, atol=1e-300,
                                           err_msg="%s bal=%s sc=%s dn=%s" % (name, bal, sc, dn))
                assert np.array_equal(got, got.T)
                if n > 4 and name != "cosine":
                    assert got[3, 1] == 0.0


def test_numpy_oracle_spot_check():
    """The NumPy restatement (identical to the reference) on a few pairs."""
    profiles = synthetic_profiles(77, 20, 7)
    got = _cabi.distance_matrix(profiles, do_scale=True)
    for i, j in ((1, 0), (7, 3), (19, 18), (12, 5)):
        assert got[i, j] == pytest.approx(ko.distance(profiles[i], profiles[j], do_scale=True), rel=RTOL)
    dist = profile_distance(sc=True)
    a, b = klib.Profile(profiles[7]), klib.Profile(profiles[3])
    assert dist.distance(a, b) == pytest.approx(ko.distance(profiles[7], profiles[3], do_scale=True), rel=RTOL)


def test_zero_profiles():
    z = np.zeros(4 ** 4, dtype=np.int64)
    x = synthetic_profiles(1, 1, 4)[0]
    assert _cabi.pair_distance(z, z) == 0.0                        # 0 / (0 + 1)
    assert np.isnan(_cabi.pair_distance(z, x, do_scale=True))      # reference: nan (+ warning)
    assert np.isnan(_cabi.pair_distance(z, z, do_scale=True))
    assert _cabi.pair_distance(z, x) == pytest.approx(ko.distance(z, x), rel=RTOL)
    assert np.isnan(_cabi.pair_distance(z, x, metric="cosine"))


def test_large_counts_and_k10_pairs():
    """Counts far from the Poisson toy range, and the BASELINE k=10 length."""
    rng = np.random.default_rng(3)
    k = 10
    profiles = np.stack([rng.poisson(lam, 4 ** k) for lam in (0.5, 3.0, 8.0, 40.0)]).astype(np.int64)
    profiles[3] *= 100_000
    got = _cabi.distance_matrix(profiles, do_scale=True)
    for i in range(1, 4):
        for j in range(i):
            assert got[i, j] == pytest.approx(ko.distance(profiles[i], profiles[j], do_scale=True), rel=RTOL)


def test_matrix_256_k9_sampled_pairs():
    """Several tile rows/columns and D-slices; sampled pairs vs the C oracle."""
    n, k = 256, 9
    profiles = synthetic_profiles(9, n, k)
    got = _cabi.distance_matrix(profiles, do_scale=True)
    rng = np.random.default_rng(0)
    for _ in range(200):
        i, j = sorted(rng.choice(n, 2, replace=False))[::-1]
        want = c_oracle.distance(profiles[i], profiles[j], do_scale=True)
        assert got[i, j] == pytest.approx(want, rel=RTOL), (i, j)
    assert np.array_equal(got, got.T)
    assert not np.diag(got).any()


def test_host_path_options_still_work(golden):
    """do_positive / do_smooth / custom pairwise stay on the host pipeline."""
    left = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60"], 4))
    right = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 4))
    d = kdistlib.ProfileDistance(do_positive=True).distance(left, right)
    mask = (left.counts != 0) & (right.counts != 0)
    assert d == pytest.approx(ko.multiset(left.counts * mask, right.counts * mask))
    custom = kdistlib.ProfileDistance(pairwise=lambda x, y: abs(x - y) / (x + y + 1))
    assert custom.distance(left, right) == pytest.approx(ko.distance(left.counts, right.counts, pairwise="sum"))


def test_matrix_session_equals_one_shot_call():
    """kpal_matrix_open/push/finish (profiles handed over in slabs of any size,
    the staging buffer reused between pushes) gives the matrix of the one-shot
    kpal_distance_matrix, for every fast-path option family."""
    rng = np.random.default_rng(77)
    n, k = 45, 6
    lam = np.exp(rng.uniform(np.log(0.3), np.log(6.0), n))
    profiles = np.stack([rng.poisson(l, 4 ** k) for l in lam]).astype(np.int64)
    for opts in (dict(do_scale=True), dict(do_balance=True, do_scale=True, down=True),
                 dict(metric="euclidean", do_scale=True), dict(metric="cosine"),
                 dict(pairwise="sum", do_balance=True)):
        want = _cabi.distance_matrix(profiles, **opts)
        with _cabi.MatrixSession(n, k, slab_bytes=7 * 8 * 4 ** k, **opts) as session:
            assert session.slab.shape == (7, 4 ** k)
            at = 0
            for m in (7, 1, 5, 7, 7, 3, 7, 7, 1):           # ragged pushes through the pinned slab
                session.slab[:m] = profiles[at:at + m]
                session.push(m)
                session.slab[:] = -1                         # the slab is free again after push
                at += m
            assert at == n
            got = session.finish()
        assert np.array_equal(got, want, equal_nan=True), opts
        with _cabi.MatrixSession(n, k, **opts) as session:   # rows from ordinary memory, one call
            session.push_rows(profiles)
            assert np.array_equal(session.finish(), want, equal_nan=True)
    with _cabi.MatrixSession(3, 2) as session:
        session.push_rows(np.ones((2, 16), dtype=np.int64))
        with pytest.raises(ValueError):
            session.finish()                                  # one profile missing
        with pytest.raises(ValueError):
            session.push_rows(np.ones((2, 16), dtype=np.int64))   # one too many


def test_matrix_command_streams_datasets(golden):
    """kmer.distance_matrix on the fast path reads the datasets slab by slab
    (kdistlib.distance_matrix_from_file) and writes the reference's text."""
    from kpal_b200 import kmer
    from test_host_api import FakeH5
    rng = np.random.default_rng(5)
    store = FakeH5()
    names = ['s%02d' % i for i in range(14)]
    rows = {}
    for name in names:
        rows[name] = rng.poisson(rng.uniform(0.5, 5.0), 4 ** 5).astype(np.int64)
        klib.Profile(rows[name], name).save(store)
    out = io.StringIO()
    kmer.distance_matrix(store, out, do_scale=True, precision=8)
    lines = out.getvalue().split('\n')
    assert lines[0] == '14' and lines[1:15] == names and lines[-1] == ''
    for i in (1, 6, 13):
        got = [float(x) for x in lines[14 + i].split(' ')]
        want = [ko.distance(rows[names[i]], rows[names[j]], do_scale=True) for j in range(i)]
        assert len(got) == i
        np.testing.assert_allclose(got, want, rtol=0, atol=0.6e-8)
    klib.Profile(np.ones(4 ** 4, dtype=np.int64), 'short').save(store)
    with pytest.raises(ValueError):
        kmer.distance_matrix(store, io.StringIO())           # kmer.py:697-698: lengths differ
