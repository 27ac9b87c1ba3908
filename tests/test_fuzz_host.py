"""
Property tests (hypothesis) of the host-side pieces, CPU only:

  * the C++ FASTA packer (csrc/pack.cpp: kpal_fasta_scan / kpal_fasta_pack) against the
    oracle's reader on arbitrary texts over an alphabet rich in the troublesome bytes
    (headers in odd places, '\\r', blanks, tabs, VT/FF, FS..US, empty lines, no final
    newline) -- the corners of Biopython's FastaIterator that the reference does not pin
     must at least be handled identically by both of our restatements;
  * the sequence-list packer against the plain text;
  * h5lite round trips of arbitrary names, dataset sizes and attribute values.
"""
import os
import tempfile

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from kpal_b200 import _cabi, h5lite
from oracle import kpal_oracle as ko
from test_cabi import expected_stream, unpack

ALPHABET = "ACGTacgtNn>;- \t\r\n\x0b\x0c\x1c\x1f*"
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow])


@settings(max_examples=400, **COMMON)
@given(st.text(alphabet=ALPHABET, max_size=300))
def test_fasta_pack_equals_oracle_reader_on_arbitrary_text(text):
    records = ko.parse_fasta(text)
    codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(text)
    assert names == [n for n, _ in records]
    assert unpack(codes, valid, n_bases) == expected_stream([s for _, s in records])
    assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for _, s in records]).tolist()


@settings(max_examples=200, **COMMON)
@given(st.lists(st.text(alphabet="ACGTacgtNRYKM-*", max_size=120), max_size=12))
def test_pack_sequences_equals_text(seqs):
    codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
    assert n_bases == sum(len(s) + 1 for s in seqs)
    assert unpack(codes, valid, n_bases) == expected_stream(seqs)
    assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for s in seqs]).tolist()


names = st.text(alphabet=st.characters(blacklist_characters="/.\x00", blacklist_categories=("Cs",)),
                min_size=1, max_size=24)
values = st.one_of(st.integers(-2 ** 62, 2 ** 62), st.floats(allow_nan=False, width=64),
                   st.text(alphabet=st.characters(blacklist_characters="\x00", blacklist_categories=("Cs",)),
                           max_size=40))


@settings(max_examples=60, **COMMON)
@given(st.dictionaries(names, st.tuples(st.integers(0, 6), st.integers(0, 2 ** 31), st.dictionaries(names, values, max_size=4)),
                       max_size=20))
def test_h5lite_round_trip_of_arbitrary_profiles(profiles):
    with tempfile.TemporaryDirectory() as directory:
        path = os.path.join(directory, 'fuzz.h5')
        expected = {}
        with h5lite.File(path, 'w') as f:
            group = f.create_group('profiles')
            for name, (k, seed, attrs) in profiles.items():
                counts = np.random.default_rng(seed).integers(0, 1000, 4 ** k).astype(np.int64)
                dataset = group.create_dataset(name, data=counts, dtype='int64', compression='gzip')
                for key, value in attrs.items():
                    dataset.attrs[key] = value
                expected[name] = (counts, attrs)
        with h5lite.File(path) as f:
            assert sorted(f['profiles'].keys(), key=lambda n: n.encode('utf-8')) == \
                sorted(expected, key=lambda n: n.encode('utf-8'))
            for name, (counts, attrs) in expected.items():
                dataset = f['profiles'][name]
                assert np.array_equal(dataset[:], counts)
                assert set(dataset.attrs.keys()) == set(attrs)
                for key, value in attrs.items():
                    got = dataset.attrs[key]
                    assert got == value and type(got) is (str if isinstance(value, str) else type(got))


# ---- the segment packer of the hybrid upload (csrc/pack.cpp: kpal_fasta_pack_segment) ----
def _segment_stream(records):
    """What a slot holds: the separator comes first (emitted at the header), then the bases."""
    lut = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = []
    for _, s in records:
        out.append(4)
        out.extend(lut.get(ch.upper(), 4) for ch in s)
    return out


def _check_segment(text, begin=0, end=None):
    end = len(text) if end is None else end
    records = ko.parse_fasta(text[begin:end])
    codes, valid, n_bases = _cabi.fasta_pack_segment(text, begin, end)
    want = _segment_stream(records)
    assert n_bases == len(want)
    assert unpack(codes, valid, n_bases) == want
    # invalid positions carry code 0 and the slot is invalid (zero) behind the emitted bases
    spread = np.zeros(len(codes), dtype=np.uint64)
    for word in range(len(valid)):
        v = int(valid[word])
        hi = sum(3 << (2 * (15 - b)) for b in range(16) if v >> (31 - b) & 1)
        lo = sum(3 << (2 * (15 - b)) for b in range(16) if v >> (15 - b) & 1)
        spread[2 * word], spread[2 * word + 1] = hi, lo
    assert not (codes.astype(np.uint64) & ~spread).any()
    assert not valid[(n_bases + 31) // 32:].any()


@settings(max_examples=400, **COMMON)
@given(st.text(alphabet=ALPHABET, max_size=300))
def test_fasta_pack_segment_equals_oracle_reader_on_arbitrary_text(text):
    _check_segment(text)


lines = st.one_of(st.text(alphabet="ACGTacgtN", min_size=0, max_size=200),
                  st.text(alphabet="ACGTacgtNn-* \t\r\x0b\x1c", min_size=0, max_size=120),
                  st.text(alphabet="ACGT>xyz 1", min_size=0, max_size=40).map(lambda s: ">" + s))


@settings(max_examples=300, **COMMON)
@given(st.lists(lines, max_size=40), st.booleans(), st.data())
def test_fasta_pack_segment_on_long_lines_and_cuts_at_headers(rows, final_newline, data):
    text = "\n".join(rows) + ("\n" if final_newline and rows else "")
    _check_segment(text)
    # any range that begins and ends at header lines is a segment of its own
    heads = [0] + [i + 1 for i in range(len(text) - 1) if text[i] == "\n" and text[i + 1] == ">"] + [len(text)]
    a = data.draw(st.integers(0, len(heads) - 1))
    b = data.draw(st.integers(a, len(heads) - 1))
    if a == 0 or text[heads[a]:heads[a] + 1] == ">":
        _check_segment(text, heads[a], heads[b])


def test_fasta_pack_segment_scalar_build_agrees():
    """The same deterministic texts through the scalar form (KPAL_NO_AVX2=1, read at load time)."""
    import subprocess
    import sys
    code = (
        "import numpy as np, random, sys\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_fuzz_host import _check_segment\n"
        "rng = random.Random(11)\n"
        "for _ in range(60):\n"
        "    rows = []\n"
        "    for _ in range(rng.randint(0, 30)):\n"
        "        kind = rng.random()\n"
        "        if kind < 0.3: rows.append('>' + ''.join(rng.choice('ab c') for _ in range(rng.randint(0, 20))))\n"
        "        elif kind < 0.9: rows.append(''.join(rng.choice('ACGTacgtN') for _ in range(rng.randint(0, 180))))\n"
        "        else: rows.append(''.join(rng.choice('ACGT \\t\\r*') for _ in range(rng.randint(0, 90))))\n"
        "    _check_segment('\\n'.join(rows) + rng.choice(['', '\\n']))\n"
        "print('ok')\n" % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    for env_extra in ({}, {"KPAL_NO_AVX2": "1"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env_extra), capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


# ---- the slotted stream: cuts at any line start + junction records (csrc/slotted.h) ----
def _count_stream(codes, valid, n_bases, k):
    counts = np.zeros(4 ** k, dtype=np.int64)
    stream = unpack(codes, valid, n_bases)
    run, index = 0, 0
    for code in stream:
        if code == 4:
            run, index = 0, 0
            continue
        index = ((index << 2) | code) & (4 ** k - 1)
        run += 1
        if run >= k:
            counts[index] += 1
    return counts


def _check_slotted(text, k, seg):
    try:
        codes, valid, n_bases = _cabi.fasta_pack_slotted(text, k, seg)
    except ValueError:
        # no header line at all (or an empty text): nothing to count either
        assert not any(line.startswith(">") for line in text.split("\n"))
        return
    assert np.array_equal(_count_stream(codes, valid, n_bases, k), ko.count_fasta(text, k)), (k, seg)


@settings(max_examples=300, **COMMON)
@given(st.lists(lines, max_size=40), st.booleans(), st.integers(1, 6), st.integers(64, 400))
def test_slotted_stream_counts_equal_the_oracle(rows, final_newline, k, seg):
    text = "\n".join(rows) + ("\n" if final_newline and rows else "")
    _check_slotted(text, k, seg)


@settings(max_examples=200, **COMMON)
@given(st.text(alphabet=ALPHABET, max_size=400), st.integers(1, 5), st.integers(64, 200))
def test_slotted_stream_on_arbitrary_text(text, k, seg):
    _check_slotted(text, k, seg)


def test_slotted_stream_wrapped_genome():
    """70-column records much longer than a segment, N blocks and soft-masking: every cut falls
    inside a record and the junction records restore the windows that cross it."""
    import random
    rng = random.Random(3)
    records = []
    for r in range(3):
        seq = "".join(rng.choice("ACGTacgt") for _ in range(rng.randint(3000, 9000)))
        at = rng.randint(0, len(seq) - 400)
        seq = seq[:at] + "N" * 300 + seq[at + 300:]
        records.append(seq)
    text = "".join(">chr%d\n%s\n" % (i, "\n".join(s[o:o + 70] for o in range(0, len(s), 70))) for i, s in enumerate(records))
    for k in (1, 2, 7, 12, 15):
        want = ko.count_sequences(records, k) if k <= 7 else None
        for seg in (64, 150, 1000, 5000):
            codes, valid, n_bases = _cabi.fasta_pack_slotted(text, k, seg)
            if want is not None:
                assert np.array_equal(_count_stream(codes, valid, n_bases, k), want), (k, seg)
            else:       # large k: compare the multiset of windows instead of a dense table
                stream = unpack(codes, valid, n_bases)
                got = {}
                run, index = 0, 0
                for code in stream:
                    if code == 4:
                        run, index = 0, 0
                        continue
                    index = ((index << 2) | code) & (4 ** k - 1)
                    run += 1
                    if run >= k:
                        got[index] = got.get(index, 0) + 1
                lut = {"A": 0, "C": 1, "G": 2, "T": 3}
                exp = {}
                for s in records:
                    run, index = 0, 0
                    for ch in s.upper():
                        if ch not in lut:
                            run, index = 0, 0
                            continue
                        index = ((index << 2) | lut[ch]) & (4 ** k - 1)
                        run += 1
                        if run >= k:
                            exp[index] = exp.get(index, 0) + 1
                assert got == exp, (k, seg)


# ---- the one-pass deflate encoder for sparse rows (csrc/rowops.cpp: sparse_deflate) ----
def _inflate_all(blob, sizes):
    import zlib
    out, at = [], 0
    for size in sizes:
        out.append(zlib.decompress(bytes(blob[at:at + size])))
        at += int(size)
    assert at == blob.size
    return out


@settings(max_examples=300, **COMMON)
@given(st.integers(1, 3000), st.integers(0, 2 ** 32 - 1), st.floats(0.0, 0.4), st.integers(1, 4))
def test_sparse_deflate_streams_inflate_to_the_data(n, seed, density, chunks):
    rng = np.random.default_rng(seed)
    data = (rng.integers(0, 256, n * chunks) * (rng.random(n * chunks) < density)).astype(np.uint8)
    for slotted in (False, True):
        if slotted:
            buffer, slot, sizes = _cabi.deflate_chunks(data, n, 4, sparse=True)
            streams = [bytes(buffer[c * slot:c * slot + sizes[c]]) for c in range(chunks)]
            import zlib
            got = [zlib.decompress(s) for s in streams]
        else:
            got = _inflate_all(*_cabi.deflate_chunks_packed(data, n, 4, sparse=True))
        assert got == [data[c * n:(c + 1) * n].tobytes() for c in range(chunks)]


def test_sparse_deflate_on_count_rows_and_run_lengths():
    rng = np.random.default_rng(9)
    rows = np.zeros((6, 4 ** 8), dtype=np.int64)
    for r in range(1, 6):
        np.add.at(rows[r], rng.integers(0, 4 ** 8, 200 * r), 1)
    rows[5, 17] = 2 ** 40 + 5                      # several non-zero bytes in one count
    rows[4, :300] = -1                             # 0xff bytes
    blob, sizes = _cabi.deflate_chunks_packed(rows, 65536, 4, sparse=True)
    raw = rows.view(np.uint8).reshape(-1)
    assert _inflate_all(blob, sizes) == [raw[c * 65536:(c + 1) * 65536].tobytes() for c in range(sizes.size)]
    assert sizes[:8].max() < 150                   # an all-zero 64 KiB chunk: 2 bits per 258 bytes
    # every zero-run length around the match lengths of deflate (3 .. 258) and their multiples
    for run in list(range(0, 12)) + [255, 256, 257, 258, 259, 260, 261, 515, 516, 517, 518, 519, 1031, 1032, 1033]:
        for lead in (0, 1, 9):
            data = np.concatenate([np.full(lead, 7, np.uint8), np.zeros(run, np.uint8), np.full(2, 9, np.uint8)])
            assert _inflate_all(*_cabi.deflate_chunks_packed(data, data.size, 4, sparse=True)) == [data.tobytes()]


def test_row_stats_of_sparse_rows_take_the_median_shortcut_exactly():
    rng = np.random.default_rng(10)
    for cols, hits in ((4 ** 8, 993), (4 ** 6, 2047), (4 ** 6, 2048), (4 ** 6, 2049), (4 ** 6, 4000), (4 ** 3 + 1, 31), (4 ** 3 + 1, 33)):
        rows = np.zeros((5, cols), dtype=np.int64)
        for r in range(5):
            rows[r, rng.choice(cols, hits, replace=False)] = rng.integers(1, 50, hits)
        stats = _cabi.row_stats(rows)
        for r, x in enumerate(rows):
            want = (float(x.sum()), float(np.count_nonzero(x)), x.mean(), np.median(x), x.std())
            assert tuple(stats[r]) == want, (cols, hits, r)
