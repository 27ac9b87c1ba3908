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


def test_matrix_600_k10_many_tiles_and_slices():
    """The BASELINE length (k = 10: 16 slices of 65536 elements per tile) with 600 profiles:
    5 tile rows x 10 tile columns, ragged last tiles, 18 000 sampled pairs of 190 profiles
    spread over every tile row and column, against the C oracle."""
    n, k = 600, 10
    rng = np.random.default_rng(10)
    highs = rng.integers(2, 18, n)
    profiles = np.empty((n, 4 ** k), dtype=np.int64)
    for i in range(n):
        profiles[i] = rng.integers(0, highs[i], 4 ** k, dtype=np.int64)
    profiles[17, ::3] = 0                        # more zeros: the union counts differ from 4^k
    profiles[411] *= 1000
    got = _cabi.distance_matrix(profiles, do_scale=True)
    assert np.array_equal(got, got.T) and not np.diag(got).any()
    pick = np.sort(rng.choice(n, 190, replace=False))
    pick[:4] = [0, 17, 127, 128]
    pick[-3:] = [411, 598, 599]
    pick = np.unique(pick)
    want = c_oracle.distance_matrix(profiles[pick], do_scale=True, threads=c_oracle.max_threads())
    low = np.tril_indices(len(pick), -1)
    sub = got[np.ix_(pick, pick)]
    rel = np.abs(sub[low] - want[low]) / np.abs(want[low])
    assert rel.max() <= RTOL, (rel.max(), len(low[0]))
    # the other fast-path metrics on the same set, a handful of pairs each
    for options in (dict(pairwise="sum", do_scale=True, down=True), dict(metric="euclidean", do_scale=True),
                    dict(metric="cosine"), dict(do_balance=True)):
        got = _cabi.distance_matrix(profiles[:200], **options)
        for i, j in ((1, 0), (199, 3), (128, 127), (77, 64), (150, 149)):
            want_ij = c_oracle.distance(profiles[i], profiles[j], **options)
            assert got[i, j] == pytest.approx(want_ij, rel=RTOL), (options, i, j)


def test_concurrent_matrices_do_not_share_accumulators():
    """Two host threads computing distance matrices on one device at the same time (ctypes
    drops the GIL): every call owns its accumulators, so both get the sequential results."""
    import threading
    sets = [synthetic_profiles(seed, 150, 7) for seed in (1, 2, 3, 4)]
    want = [_cabi.distance_matrix(p, do_scale=True) for p in sets]
    results = [None] * len(sets)

    def work(i):
        for _ in range(3):
            results[i] = _cabi.distance_matrix(sets[i], do_scale=True)
            assert np.array_equal(results[i], want[i])
            pair = _cabi.pair_distance(sets[i][0], sets[i][1], do_scale=True)
            assert pair == pytest.approx(want[i][1, 0], rel=1e-12)      # another kernel: other summation order
    errors = []

    def guarded(i):
        try:
            work(i)
        except Exception as exc:            # surfaces in the main thread
            errors.append((i, exc))
    threads = [threading.Thread(target=guarded, args=(i,)) for i in range(len(sets))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------- tensor-core Gram form (distance_gram.cu)
def _gram_option(value):
    _cabi.check(_cabi.load().kpal_set_option(b"gram", int(value)))


#: tolerance of the Gram form for euclidean / cosine: the Gram matrix and the numerators are
#: exact integers (tcgen05 kind::i8 into 32-bit accumulators, 128-bit numerators); what is left
#: is one int -> double conversion, one division and one square root
GRAM_RTOL = 1e-9


@pytest.mark.parametrize("n,k,lam_hi", [(13, 4, 6.0), (129, 5, 8.0), (300, 7, 8.0), (257, 8, 3.0), (40, 9, 20.0)])
def test_gram_euclidean_cosine_vs_oracle(n, k, lam_hi):
    """Euclidean distance and cosine similarity matrices through the exact integer Gram matrix
    on the tensor cores: every scale / down / balance combination against the C oracle, ragged
    tile edges (n not a multiple of 128 / 256), one to 2048 K blocks, several K ranges per tile."""
    profiles = synthetic_profiles(n * 31 + k, n, k, lam_hi=lam_hi)
    assert profiles.max() <= 255
    low = np.tril_indices(n, -1)
    try:
        for metric in ("euclidean", "cosine"):
            for opts in (dict(), dict(do_scale=True), dict(do_scale=True, down=True), dict(do_balance=True),
                         dict(do_balance=True, do_scale=True)):
                if 2 * profiles.max() > 255 and opts.get("do_balance"):
                    continue
                _gram_option(1)
                got = _cabi.distance_matrix(profiles, metric=metric, **opts)
                want = c_oracle.distance_matrix(profiles, metric=metric, threads=c_oracle.max_threads(), **opts)
                rel = np.abs(got[low] - want[low]) / np.maximum(np.abs(want[low]), 1e-300)
                assert rel.max() <= GRAM_RTOL, (metric, opts, rel.max())
                assert np.array_equal(got, got.T)
                _gram_option(0)                      # the element-wise fp64 kernel agrees with it
                plain = _cabi.distance_matrix(profiles, metric=metric, **opts)
                assert np.allclose(plain[low], got[low], rtol=1e-9, atol=0), (metric, opts)
                if metric == "euclidean":
                    assert not np.diag(got).any()
                else:
                    assert np.allclose(np.diag(got), 1.0, rtol=1e-15)
    finally:
        _gram_option(1)


def test_gram_identical_and_zero_profiles_and_fallback():
    """Exactness shows where floating point cancels: identical profiles are at distance exactly 0
    (scaled: proportional profiles too), an all-zero profile gives the reference's nan; counts
    above 255 leave the 8-bit form and take the fp64 kernel, with the same results."""
    rng = np.random.default_rng(8)
    n, k = 64, 6
    profiles = rng.poisson(3.0, (n, 4 ** k)).astype(np.int64)
    profiles[5] = profiles[9]
    profiles[11] = 3 * profiles[9]                   # proportional: scaled euclidean distance 0
    profiles[20] = 0
    got = _cabi.distance_matrix(profiles, metric="euclidean")
    assert got[5, 9] == 0.0 and got[9, 5] == 0.0
    scaled = _cabi.distance_matrix(profiles, metric="euclidean", do_scale=True)
    assert scaled[11, 9] == 0.0 and scaled[5, 11] == 0.0
    assert np.isnan(scaled[20, 3]) and np.isnan(scaled[3, 20]) and np.isnan(scaled[20, 20])
    cos = _cabi.distance_matrix(profiles, metric="cosine")
    assert np.isnan(cos[20, 3]) and cos[5, 9] == pytest.approx(1.0, rel=1e-15)
    assert got[20, 3] == pytest.approx(np.sqrt(float((profiles[3] ** 2).sum())), rel=1e-15)
    big = profiles.copy()
    big[33, 17] = 70_000                              # no longer 8-bit: fp64 tile kernel
    want = c_oracle.distance_matrix(big, metric="euclidean", do_scale=True, threads=c_oracle.max_threads())
    fallback = _cabi.distance_matrix(big, metric="euclidean", do_scale=True)
    low = np.tril_indices(n, -1)
    ok = ~np.isnan(want[low])
    assert np.allclose(fallback[low][ok], want[low][ok], rtol=RTOL, atol=0)
    assert np.array_equal(np.isnan(fallback[low]), np.isnan(want[low]))


def test_gram_accumulation_chunks():
    """Norms of 2^31 and more: one 32-bit accumulation over the whole profile could overflow, so
    the profile is accumulated in chunks of 32768 elements added in 64-bit integers."""
    rng = np.random.default_rng(9)
    n, k = 20, 9
    profiles = rng.integers(200, 256, (n, 4 ** k)).astype(np.int64)     # norms ~ 1.4e10 >> 2^31
    profiles[:, ::5] = 0
    got = _cabi.distance_matrix(profiles, metric="euclidean", do_scale=True)
    cos = _cabi.distance_matrix(profiles, metric="cosine")
    for i, j in ((1, 0), (19, 3), (10, 9)):
        assert got[i, j] == pytest.approx(c_oracle.distance(profiles[i], profiles[j], metric="euclidean", do_scale=True),
                                          rel=GRAM_RTOL)
        assert cos[i, j] == pytest.approx(c_oracle.distance(profiles[i], profiles[j], metric="cosine"), rel=GRAM_RTOL)


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
