"""
CPU tests: pin the oracle (oracle/kpal_oracle.py and oracle/kpal_oracle.c)
against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and -- when the reference tree is present, i.e.
in the build container -- against the reference itself on randomised inputs.
"""
import io
import random

import numpy as np
import pytest

from conftest import dense
from oracle import c_oracle, kpal_oracle as ko, ref_loader


METRIC_OPTS = {
    "multiset-prod": dict(metric="multiset", pairwise="prod"),
    "multiset-sum": dict(metric="multiset", pairwise="sum"),
    "euclidean": dict(metric="euclidean"),
    "cosine": dict(metric="cosine"),
}


def parse_key(key):
    bal, sc, dn, metric = key.split("_", 3)
    opts = dict(do_balance=bal == "bal1", do_scale=sc == "sc1", down=dn == "dn1")
    opts.update(METRIC_OPTS[metric])
    return opts


def test_count_golden_fixtures(golden):
    for case in golden["count_cases"] + golden["odd_cases"]:
        want = dense(case["counts"])
        k = case["k"]
        assert np.array_equal(ko.count_sequences(case["sequences"], k), want)
        assert np.array_equal(ko.count_python(case["sequences"], k), want)
        assert np.array_equal(c_oracle.count_sequences(case["sequences"], k), want)
        assert np.array_equal(c_oracle.count_sequences(case["sequences"], k, threads=3), want)


def test_count_golden_fasta(golden):
    text = golden["fasta_text"]
    for case in golden["fasta_cases"]:
        k = case["k"]
        assert np.array_equal(ko.count_fasta(text, k), dense(case["counts"]))
        got = ko.count_fasta_by_record(text, k, prefix="pre")
        assert [n for n, _ in got] == [n for n, _ in case["by_record"]]
        for (_, counts), (_, want) in zip(got, case["by_record"]):
            assert np.array_equal(counts, dense(want))


def test_balance_golden(golden):
    for case in golden["balance_cases"]:
        before, after = dense(case["before"]), dense(case["after"])
        assert np.array_equal(ko.balance(before), after)
        assert np.array_equal(ko.balance_python(before), after)
        assert np.array_equal(c_oracle.balance(before), after)
    for entry in golden["reverse_complement"]:
        k = entry["k"]
        assert ko.reverse_complement_table(k).tolist() == entry["table"]
        assert [ko.reverse_complement(i, k) for i in range(4 ** k)] == entry["table"]


def test_distance_goldens(golden):
    g = golden["distance_k2"]
    assert ko.distance(g["left"], g["right"]) == 0.0625          # reference test_kdistlib.py:104-112
    assert c_oracle.distance(g["left"], g["right"]) == 0.0625
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    vals = golden["distance_k8"]["values"]
    np.testing.assert_almost_equal(vals["default"], 0.4626209322)  # test_kdistlib.py:114-122
    cases = {"default": {}, "balance": dict(do_balance=True), "scale": dict(do_scale=True),
             "scale_down": dict(do_scale=True, down=True), "sum": dict(pairwise="sum"),
             "euclidean": dict(metric="euclidean"), "cosine": dict(metric="cosine")}
    for tag, opts in cases.items():
        assert ko.distance(left, right, **opts) == vals[tag]
        assert c_oracle.distance(left, right, **opts) == pytest.approx(vals[tag], rel=1e-13)


def test_distance_matrix_goldens(golden, golden_profiles):
    profiles = golden_profiles["profiles"]
    for key in golden_profiles.files:
        if key == "profiles":
            continue
        opts = parse_key(key)
        want = golden_profiles[key]
        got = ko.distance_matrix_values(list(profiles), **opts)
        assert np.array_equal(got, want, equal_nan=True), key
        got_c = c_oracle.distance_matrix(profiles, threads=2, **opts)
        np.testing.assert_allclose(np.tril(got_c, -1), want, rtol=1e-12, atol=1e-300, err_msg=key)
    text = ko.format_matrix([str(i) for i in range(len(profiles))],
                            golden_profiles["bal0_sc1_dn0_multiset-prod"], 10)
    assert text == golden["synthetic_matrix_text_scaled_p10"]


def test_matrix_text_golden(golden):
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    profs = {"1": [left], "2": [left, right], "3": [left, right, left]}
    for n, plist in profs.items():
        values = ko.distance_matrix_values(plist)
        assert ko.format_matrix("abc"[:int(n)], values, 2) == golden["matrix_text_k8_precision2"][n]
    assert golden["matrix_text_k8_precision2"]["3"].split("\n")[:6] == ['3', 'a', 'b', 'c', '0.46', '0.00 0.46']


def test_tutorial_fixture(golden, tutorial_texts):
    """Multi-line (wrapped) FASTA records: reference doc/tutorial.rst:44-144."""
    tut = golden["tutorial"]
    profs = {}
    for name, text in tutorial_texts.items():
        counts = ko.count_fasta(text, tut["k"])
        profs[name] = counts
        assert int(counts.sum()) == tut["profiles"][name]["total"]
        assert int(np.count_nonzero(counts)) == tut["profiles"][name]["non_zero"]
        assert np.array_equal(c_oracle.count_bytes(
            "\n".join(s for _, s in ko.parse_fasta(text)).encode(), tut["k"]), counts)
    assert ko.distance(profs["c_1"], profs["c_2"]) == tut["distance_c_1_c_2"]
    merged = [profs[s + "_1"] + profs[s + "_2"] for s in "abcd"]
    names = ["%s_1_%s_2" % (s, s) for s in "abcd"]
    assert ko.format_matrix(names, ko.distance_matrix_values(merged), 3) == tut["merged_matrix_text_p3"]
    assert ko.format_matrix(names, ko.distance_matrix_values(merged), 10) == tut["merged_matrix_text_p10"]


def test_fasta_reader_rules():
    text = "skipped\n>a desc\nAC GT\r\nNN\n\n>\nTT\n>b\tz\n"
    assert ko.parse_fasta(text) == [("a", "ACGTNN"), ("", "TT"), ("b", "")]


needs_reference = pytest.mark.skipif(not ref_loader.available(),
                                     reason="reference tree not present (GPU box)")


@needs_reference
def test_oracle_equals_reference_counts_randomised():
    klib, _, _ = ref_loader.load()
    rng = random.Random(11)
    alphabet = "ACGT" * 8 + "acgtNnRY-* \t"
    for _ in range(150):
        k = rng.randint(1, 7)
        seqs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 90)))
                for _ in range(rng.randint(0, 5))]
        ref = klib.Profile.from_sequences(seqs, k).counts
        assert np.array_equal(ko.count_sequences(seqs, k), ref)
        assert np.array_equal(c_oracle.count_sequences(seqs, k), ref)
        p = klib.Profile(ref.copy())
        p.balance()
        assert np.array_equal(ko.balance(ref), p.counts)


@needs_reference
def test_oracle_equals_reference_distances_randomised():
    klib, kdistlib, metrics = ref_loader.load()
    rng = np.random.default_rng(5)
    for trial in range(25):
        k = int(rng.integers(2, 6))
        left = rng.poisson(rng.uniform(0.2, 6), 4 ** k)
        right = rng.poisson(rng.uniform(0.2, 6), 4 ** k)
        for bal in (False, True):
            for sc, dn in ((False, False), (True, False), (True, True)):
                for name, opts in METRIC_OPTS.items():
                    fn = {"euclidean": metrics.euclidean,
                          "cosine": metrics.cosine_similarity}.get(opts["metric"])
                    dist = kdistlib.ProfileDistance(
                        do_balance=bal, do_scale=sc, down=dn, distance_function=fn,
                        pairwise=metrics.pairwise[opts.get("pairwise", "prod")])
                    ref = dist.distance(klib.Profile(left.copy()), klib.Profile(right.copy()))
                    got = ko.distance(left, right, do_balance=bal, do_scale=sc, down=dn, **opts)
                    assert got == ref
                    got_c = c_oracle.distance(left, right, do_balance=bal, do_scale=sc, down=dn, **opts)
                    assert got_c == pytest.approx(ref, rel=1e-12)


@needs_reference
def test_reference_fasta_through_reader():
    klib, _, _ = ref_loader.load()
    text = ">x\nACGTACGTTTGA\nACGNNACGTA\n>y\nacgtacgtaa\n"
    prof = klib.Profile.from_fasta(io.StringIO(text), 4)
    assert np.array_equal(prof.counts, ko.count_fasta(text, 4))


def test_split_showbalance_positive_golden():
    """oracle.split / show_balance / distance(do_positive) against vectors made
    by the unmodified reference (tests/golden/make_golden_split.py), including
    the upstream golden '1 0.669' (reference tests/test_kmer.py:195-202)."""
    import json
    import os
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "golden_split.json")) as f:
        g = json.load(f)
    for case in g["split_cases"]:
        counts = dense(case["counts"])
        for fn in (ko.split, ko.split_python):
            forward, reverse = fn(counts)
            assert forward.tolist() == case["forward"] and reverse.tolist() == case["reverse"]
        assert ko.show_balance(counts) == case["showbalance"]
    for case in g["positive_cases"]:
        left, right = dense(case["left"]), dense(case["right"])
        for key, want in case["values"].items():
            got = ko.distance(left, right, do_positive=True, **parse_key(key))
            assert got == want or (np.isnan(got) and np.isnan(want)), key


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_split_showbalance_against_reference_source(golden):
    rklib, _, rmetrics = ref_loader.load()
    counts = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    forward, reverse = rklib.Profile(counts.copy()).split()
    a, b = ko.split(counts)
    assert np.array_equal(a, forward) and np.array_equal(b, reverse)
    value = ko.show_balance(counts)
    assert value == rmetrics.multiset(forward, reverse, rmetrics.pairwise['prod'])
    assert '{0:.3f}'.format(value) == '0.669'
