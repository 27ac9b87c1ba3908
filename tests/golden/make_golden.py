#!/usr/bin/env python
"""
Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (LUMC/kPAL, read-only at /root/reference) through
oracle/ref_loader.py.  Runs only in the build container; the outputs
(golden.json, golden_profiles.npz) are committed and are what the tests on
the GPU box read.

    python tests/golden/make_golden.py

Sources of the vectors:
  * the reference's own fixtures and expected values
    (tests/utils.py:25-59, tests/test_klib.py:32-99,164-180,
    tests/test_kdistlib.py:38-134, doc/tutorial.rst:44-144 via
    doc/downloads/tutorial.zip), re-computed here with the reference code;
  * seeded synthetic inputs (stored verbatim, not re-generated in the tests)
    pushed through the reference for the options its fixtures do not
    discriminate (scale / down / sum / euclidean / cosine / balance).
"""
import importlib.util
import io
import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def sparse(counts):
    nz = np.nonzero(counts)[0]
    return {"size": int(counts.size), "idx": nz.tolist(),
            "val": np.asarray(counts)[nz].tolist()}


def load_ref_test_utils(root):
    spec = importlib.util.spec_from_file_location(
        "_ref_test_utils", os.path.join(root, "tests", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    klib, kdistlib, metrics = ref_loader.load()
    root = ref_loader.reference_root()
    utils = load_ref_test_utils(root)
    golden = {"reference": "LUMC/kPAL (unmodified, via oracle/ref_loader.py)"}

    # ---- counting: the reference's fixtures (tests/test_klib.py:47-99) ----
    fixture_sets = {
        "LENGTH_8": utils.LENGTH_8,
        "LENGTH_8_WITH_N": utils.LENGTH_8_WITH_N,
        "LENGTH_60": utils.LENGTH_60,
        "LENGTH_60_WITH_N": utils.LENGTH_60_WITH_N,
        "LENGTH_60_MORE": utils.LENGTH_60_MORE,
    }
    golden["fixtures"] = fixture_sets
    count_cases = []
    for name, seqs in fixture_sets.items():
        for k in (1, 4, 7, 8):
            for subset, tag in ((seqs[:1], "single"), (seqs, "multi")):
                ref = klib.Profile.from_sequences(subset, k).counts
                # the reference's independent naive oracle (tests/utils.py:70-100)
                naive = utils.as_array(utils.counts(subset, k), k)
                assert np.array_equal(ref, naive)
                count_cases.append({"fixture": name, "which": tag, "k": k,
                                    "sequences": subset,
                                    "counts": sparse(ref)})
    golden["count_cases"] = count_cases

    # ---- counting: odd characters, case, short runs, empty records ----
    odd = ["ACGTNNACGTacgtRYACGT-ACGT*ACG TACGT\tACGTACGTAC", "", "ACG", "acgtacgtacgt",
           "NNNNNNNN", "ACGTACGTNACGTACGTXacgtnACGT", "A", "TTTTTTTTTTTTTTTTTTTTTTTTT"]
    odd_cases = []
    for k in (1, 2, 3, 5, 9):
        ref = klib.Profile.from_sequences(odd, k).counts
        odd_cases.append({"k": k, "sequences": odd, "counts": sparse(ref)})
    golden["odd_cases"] = odd_cases

    # ---- FASTA text: wrapped lines, CRLF, blank lines, empty header ----
    rng = np.random.default_rng(20261017)
    body = "".join(rng.choice(list("ACGTacgtN"), size=1500,
                              p=[.22, .22, .22, .22, .025, .025, .025, .025, .02]))
    recs = [("r1 some description", body[:400]), ("r2", body[400:460]),
            ("", body[460:900]), ("r4\tx", body[900:]), ("r5", "")]
    lines = ["; leading text before the first header is skipped", ""]
    for i, (title, seq) in enumerate(recs):
        lines.append(">" + title)
        for p in range(0, len(seq), 60):
            lines.append(seq[p:p + 60] + ("\r" if i == 1 else ""))
        if i == 2:
            lines.append("")
    fasta_text = "\n".join(lines) + "\n"
    fasta_cases = []
    for k in (3, 6, 10):
        ref = klib.Profile.from_fasta(io.StringIO(fasta_text), k).counts
        by_rec = [(p.name, sparse(p.counts)) for p in
                  klib.Profile.from_fasta_by_record(io.StringIO(fasta_text), k, prefix="pre")]
        fasta_cases.append({"k": k, "counts": sparse(ref), "by_record": by_rec})
    golden["fasta_text"] = fasta_text
    golden["fasta_cases"] = fasta_cases

    # ---- balance (tests/test_klib.py:164-180) ----
    bal_cases = []
    for seqs, k in ((utils.SEQUENCES, 8), (["AATT"], 4), (utils.LENGTH_60_MORE, 5),
                    (utils.SEQUENCES, 3)):
        p = klib.Profile.from_sequences(seqs, k)
        before = p.counts.copy()
        p.balance()
        bal_cases.append({"k": k, "sequences": seqs, "before": sparse(before),
                          "after": sparse(p.counts)})
    golden["balance_cases"] = bal_cases
    golden["reverse_complement"] = [
        {"k": k, "table": [int(klib.Profile(np.zeros(4 ** k, dtype="int64")).reverse_complement(i))
                           for i in range(4 ** k)]} for k in (1, 2, 3, 4)]

    # ---- distances: the reference's goldens (tests/test_kdistlib.py) ----
    from collections import Counter
    ca = Counter(['AC', 'AG', 'AT', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TG', 'TT'])
    cb = Counter(['AC', 'AT', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TC', 'TG', 'TT'])
    pa = klib.Profile(utils.as_array(ca, 2))
    pb = klib.Profile(utils.as_array(cb, 2))
    d2 = kdistlib.ProfileDistance().distance(pa, pb)
    assert d2 == 0.0625                                   # test_kdistlib.py:104-112
    golden["distance_k2"] = {"left": pa.counts.tolist(), "right": pb.counts.tolist(),
                             "distance": d2}

    left = klib.Profile(utils.as_array(utils.counts(utils.SEQUENCES_LEFT, 8), 8), 'a')
    right = klib.Profile(utils.as_array(utils.counts(utils.SEQUENCES_RIGHT, 8), 8), 'b')
    third = klib.Profile(utils.as_array(utils.counts(utils.SEQUENCES_LEFT, 8), 8), 'c')
    k8 = {}
    option_sets = {
        "default": {},
        "balance": {"do_balance": True},
        "scale": {"do_scale": True},
        "scale_down": {"do_scale": True, "down": True},
        "sum": {"pairwise": "sum"},
        "euclidean": {"distance_function": "euclidean"},
        "cosine": {"distance_function": "cosine"},
    }

    def make_dist(opts):
        kw = dict(opts)
        if "pairwise" in kw:
            kw["pairwise"] = metrics.pairwise[kw["pairwise"]]
        if "distance_function" in kw:
            kw["distance_function"] = metrics.vector_distance[kw["distance_function"]]
        return kdistlib.ProfileDistance(**kw)

    for tag, opts in option_sets.items():
        k8[tag] = float(make_dist(opts).distance(left, right))
    np.testing.assert_almost_equal(k8["default"], 0.4626209322)   # test_kdistlib.py:114-122
    golden["distance_k8"] = {"left_fixture": "LENGTH_60", "right_fixture": "LENGTH_60_MORE",
                             "k": 8, "values": k8}
    texts = {}
    for n, profs in ((1, [left]), (2, [left, right]), (3, [left, right, third])):
        out = io.StringIO()
        kdistlib.distance_matrix(profs, out, 2, kdistlib.ProfileDistance())
        texts[str(n)] = out.getvalue()
    assert texts["3"].strip().split("\n") == ['3', 'a', 'b', 'c', '0.46', '0.00 0.46']
    golden["matrix_text_k8_precision2"] = texts                     # test_kdistlib.py:38-74

    # ---- distances on seeded synthetic profiles: all fast-path options ----
    rng = np.random.default_rng(4)
    k = 5
    n_prof = 7
    lam = np.exp(rng.uniform(np.log(0.3), np.log(8.0), n_prof))
    profiles = np.stack([rng.poisson(l, 4 ** k) for l in lam]).astype(np.int64)
    profiles[5] = profiles[2]                      # identical pair -> distance 0
    profiles[6] = profiles[1] * 3                  # exact multiple -> scaled distance 0
    plist = [klib.Profile(p.copy(), str(i)) for i, p in enumerate(profiles)]
    matrices = {}
    for bal in (False, True):
        for sc, dn in ((False, False), (True, False), (True, True)):
            for metric in ("multiset-prod", "multiset-sum", "euclidean", "cosine"):
                opts = {"do_balance": bal, "do_scale": sc, "down": dn}
                if metric == "multiset-sum":
                    opts["pairwise"] = "sum"
                elif metric in ("euclidean", "cosine"):
                    opts["distance_function"] = metric
                dist = make_dist(opts)
                m = np.zeros((n_prof, n_prof))
                for i in range(1, n_prof):
                    for j in range(i):
                        m[i, j] = dist.distance(plist[i], plist[j])
                matrices["bal%d_sc%d_dn%d_%s" % (bal, sc, dn, metric)] = m
    out = io.StringIO()
    kdistlib.distance_matrix(plist, out, 10, make_dist({"do_scale": True}))
    golden["synthetic_matrix_text_scaled_p10"] = out.getvalue()
    np.savez_compressed(os.path.join(HERE, "golden_profiles.npz"),
                        profiles=profiles, **matrices)

    # ---- tutorial fixture (doc/tutorial.rst:36-146, doc/downloads/tutorial.zip) ----
    # The only fixture in the reference tree that pins multi-line (60-column
    # wrapped) FASTA records.  The eight small FASTA files are extracted to
    # tests/golden/tutorial/ (data, not source) so the GPU-box tests can read
    # them; expected values are re-computed here with the reference code and
    # cross-checked against the numbers printed in doc/tutorial.rst.
    zpath = os.path.join(root, "doc", "downloads", "tutorial.zip")
    tut = {"k": 8, "profiles": {}}
    tdir = os.path.join(HERE, "tutorial")
    os.makedirs(tdir, exist_ok=True)
    with zipfile.ZipFile(zpath) as z:
        names = sorted(n for n in z.namelist() if n.endswith(".fa"))
        profs = {}
        for n in names:
            text = z.read(n).decode()
            stem = os.path.splitext(os.path.basename(n))[0]
            with open(os.path.join(tdir, stem + ".fa"), "w") as f:
                f.write(text)
            p = klib.Profile.from_fasta(io.StringIO(text), 8, name=stem)
            profs[stem] = p
            tut["profiles"][stem] = {"total": int(p.total), "non_zero": int(p.non_zero)}
    # doc/tutorial.rst:44-82
    assert [tut["profiles"][s]["non_zero"] for s in ("a_1", "b_1", "c_1", "d_1")] == \
        [16141, 16188, 16148, 16191]
    assert all(tut["profiles"][s]["total"] == 18600 for s in ("a_1", "b_1", "c_1", "d_1"))
    d = float(kdistlib.ProfileDistance().distance(profs["c_1"], profs["c_2"]))
    assert "%.3f" % d == "0.456"                       # doc/tutorial.rst:128-129
    tut["distance_c_1_c_2"] = d
    merged = []
    for s in "abcd":
        m = profs[s + "_1"].copy()
        m.merge(profs[s + "_2"])
        m.name = "%s_1_%s_2" % (s, s)
        merged.append(m)
    assert (int(merged[2].total), int(merged[2].non_zero)) == (37200, 28398)   # rst:107-118
    out = io.StringIO()
    kdistlib.distance_matrix(merged, out, 3, kdistlib.ProfileDistance())
    assert out.getvalue().split("\n")[5:8] == ["0.415", "0.416 0.416", "0.414 0.413 0.414"]
    tut["merged_matrix_text_p3"] = out.getvalue()          # doc/tutorial.rst:136-144
    out = io.StringIO()
    kdistlib.distance_matrix(merged, out, 10, kdistlib.ProfileDistance())
    tut["merged_matrix_text_p10"] = out.getvalue()
    golden["tutorial"] = tut

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=0, sort_keys=True)
    print("wrote golden.json (%d count cases) and golden_profiles.npz (%d matrices)"
          % (len(count_cases), len(matrices)))
    print("tutorial:", tut)


if __name__ == "__main__":
    main()
