"""
GPU parity tests for the counting path (run on the B200 box: pytest -m gpu).
Everything goes through the public API / C ABI of kpal_b200 and is compared
bit-exactly with the oracle and with the golden vectors of the reference.
Mirrors reference tests/test_klib.py:32-99,164-180.
"""
import ctypes
import io
import random

import numpy as np
import pytest

from conftest import dense
from kpal_b200 import _cabi, klib
from oracle import c_oracle, kpal_oracle as ko

pytestmark = pytest.mark.gpu


def check_profile(profile, want, k, name=None):
    """Same assertions as reference tests/utils.py:139-148."""
    assert profile.length == k
    assert profile.total == want.sum()
    assert profile.non_zero == np.count_nonzero(want)
    assert profile.counts.dtype == np.int64
    assert np.array_equal(profile.counts, want)
    if name:
        assert profile.name == name


def fasta_of(sequences, names=None):
    """One line per record, like reference tests/utils.py:184-196."""
    names = names or ["sequence_%d" % (i + 1) for i in range(len(sequences))]
    return "\n".join(">" + n + "\n" + s for n, s in zip(names, sequences)) + "\n"


def test_from_fasta_reference_fixtures(golden):
    for case in golden["count_cases"]:
        k, seqs = case["k"], case["sequences"]
        profile = klib.Profile.from_fasta(io.StringIO(fasta_of(seqs)), k, name="abc")
        check_profile(profile, dense(case["counts"]), k, name="abc")
        profile = klib.Profile.from_sequences(seqs, k)
        check_profile(profile, dense(case["counts"]), k)


def test_from_sequences_odd_characters(golden):
    for case in golden["odd_cases"]:
        profile = klib.Profile.from_sequences(iter(case["sequences"]), case["k"])
        check_profile(profile, dense(case["counts"]), case["k"])


def test_from_fasta_by_record(golden):
    seqs = golden["fixtures"]["LENGTH_60"]
    names = [str(i) for i in range(len(seqs))]
    for prefix in (None, "pre"):
        profiles = list(klib.Profile.from_fasta_by_record(
            io.StringIO(fasta_of(seqs, names)), 4, prefix=prefix))
        assert len(profiles) == len(seqs)
        for name, seq, profile in zip(names, seqs, profiles):
            check_profile(profile, ko.count_sequences([seq], 4), 4,
                          name=(prefix + "_" + name) if prefix else name)


def test_fasta_text_golden(golden):
    text = golden["fasta_text"]
    for case in golden["fasta_cases"]:
        k = case["k"]
        check_profile(klib.Profile.from_fasta(io.StringIO(text), k), dense(case["counts"]), k)
        got = list(klib.Profile.from_fasta_by_record(io.StringIO(text), k, prefix="pre"))
        assert [p.name for p in got] == [n for n, _ in case["by_record"]]
        for profile, (_, want) in zip(got, case["by_record"]):
            check_profile(profile, dense(want), k)


def test_balance_golden(golden):
    for case in golden["balance_cases"]:
        profile = klib.Profile(dense(case["before"]))
        profile.balance()
        check_profile(profile, dense(case["after"]), case["k"])


def test_balance_fused_equals_count_then_balance():
    rng = random.Random(2)
    seqs = ["".join(rng.choice("ACGTN") for _ in range(500)) for _ in range(20)]
    for k in (1, 2, 5, 6, 8, 9, 11):
        plain = _cabi.count_sequences(seqs, k)
        fused = _cabi.count_sequences(seqs, k, balance=True)
        assert np.array_equal(fused, ko.balance(plain))
        assert fused.sum() == 2 * plain.sum()
        p = klib.Profile(plain.copy())
        p.balance()
        assert np.array_equal(p.counts, fused)


def test_tutorial_fixture(golden, tutorial_texts):
    """60-column wrapped FASTA: reference doc/tutorial.rst:44-82."""
    tut = golden["tutorial"]
    for name, text in tutorial_texts.items():
        profile = klib.Profile.from_fasta(io.StringIO(text), tut["k"], name=name)
        assert int(profile.total) == tut["profiles"][name]["total"]
        assert int(profile.non_zero) == tut["profiles"][name]["non_zero"]
        assert np.array_equal(profile.counts, ko.count_fasta(text, tut["k"]))


def test_randomised_against_oracle():
    rng = random.Random(7)
    alphabet = "ACGT" * 10 + "acgtNnRY-*"
    for trial in range(60):
        k = rng.randint(1, 13)
        n_rec = rng.choice([0, 1, 2, 5, 40])
        seqs = ["".join(rng.choice(alphabet) for _ in range(rng.choice([0, 1, k - 1, k, k + 1, 63, 64, 65, 200, 3000])))
                for _ in range(n_rec)]
        got = _cabi.count_sequences(seqs, k)
        assert np.array_equal(got, ko.count_sequences(seqs, k)), (k, n_rec)


def test_empty_and_short_inputs():
    for k in (1, 4, 12):
        for seqs in ([], [""], ["A" * (k - 1)], ["N" * 100]):
            assert not _cabi.count_sequences(seqs, k).any()
        assert _cabi.count_sequences(["a" * k], k)[0] == 1
        assert _cabi.count_sequences(["T" * k], k)[-1] == 1
    assert not _cabi.count_fasta("", 3).any()
    assert not _cabi.count_fasta(">x\n", 3).any()


def test_k_out_of_range():
    with pytest.raises(ValueError):
        klib.Profile.from_sequences(["ACGT"], 0)
    with pytest.raises(ValueError):
        klib.Profile.from_sequences(["ACGT"], 16)


def random_reads(seed, n_reads, read_len, p_n=0.001, p_lower=0.05):
    rng = np.random.default_rng(seed)
    n = n_reads * read_len
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
    lower = rng.random(n) < p_lower
    bases = np.where(lower, bases + 32, bases).astype(np.uint8)
    bases[rng.random(n) < p_n] = ord("N")
    return bases.reshape(n_reads, read_len)


def reads_to_fasta(reads):
    n_reads, read_len = reads.shape
    header = np.frombuffer(b">r0000000\n", dtype=np.uint8)
    out = np.empty((n_reads, len(header) + read_len + 1), dtype=np.uint8)
    out[:, :len(header)] = header
    digits = np.arange(n_reads)
    for d in range(7):
        out[:, 8 - d] = ord("0") + (digits // 10 ** d) % 10
    out[:, len(header):-1] = reads
    out[:, -1] = ord("\n")
    return out.tobytes()


@pytest.mark.parametrize("k,n_reads,read_len,balance", [
    (6, 1, 1_000_000, False),         # BASELINE config 1
    (12, 66_667, 150, True),          # BASELINE config 2 at 1/10 size
    (13, 4, 2_500_000, True),
    (9, 20_000, 150, False),
])
def test_large_synthetic_fasta(k, n_reads, read_len, balance):
    reads = random_reads(k * 1000 + n_reads, n_reads, read_len)
    fasta = reads_to_fasta(reads)
    got = _cabi.count_fasta(fasta, k, balance=balance)
    want = c_oracle.count_bytes(np.insert(reads, read_len, ord("\n"), axis=1).tobytes(), k,
                                threads=c_oracle.max_threads())
    if balance:
        want = ko.balance(want)
    assert np.array_equal(got, want)
    windows = want.sum() // (2 if balance else 1)
    assert windows <= n_reads * (read_len - k + 1)


def test_by_record_many_records():
    k = 8
    reads = random_reads(33, 3000, 1000)
    fasta = reads_to_fasta(reads)
    profiles = list(klib.Profile.from_fasta_by_record(io.BytesIO(fasta), k))
    assert len(profiles) == 3000
    assert profiles[0].name == "r0000000" and profiles[-1].name == "r0002999"
    for i in (0, 1, 2, 777, 1500, 2998, 2999):
        assert np.array_equal(profiles[i].counts,
                              c_oracle.count_bytes(reads[i].tobytes(), k)), i
    total = sum(int(p.counts.sum()) for p in profiles)
    assert total == int(c_oracle.count_bytes(
        np.insert(reads, 1000, ord("\n"), axis=1).tobytes(), k).sum())


def test_by_record_20k_records_every_row():
    """BASELINE configs[2] at a fifth of its size: 20 000 records of 300 - 1200 bp at k = 8
    through Profile.from_fasta_by_record (80 device batches); EVERY row is compared with the
    oracle, names included."""
    k = 8
    rng = np.random.default_rng(20)
    lengths = rng.integers(300, 1201, 20_000)
    letters = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    parts, seqs = [], []
    for i, n in enumerate(lengths):
        seq = letters[rng.choice(9, n, p=[.24, .24, .24, .24, .01, .01, .005, .005, .01])]
        seqs.append(seq)
        parts.append(b">rec%05d extra words\n" % i)
        parts.append(seq.tobytes())
        parts.append(b"\n")
    lut = np.full(256, -1, dtype=np.int64)
    for c, v in zip(b"ACGTacgt", (0, 1, 2, 3, 0, 1, 2, 3)):
        lut[c] = v
    n_seen = 0
    for i, profile in enumerate(klib.Profile.from_fasta_by_record(io.BytesIO(b"".join(parts)), k)):
        assert profile.name == "rec%05d" % i
        codes = lut[seqs[i]]
        ok = np.convolve((codes >= 0).astype(np.int64), np.ones(k, dtype=np.int64), "valid") == k
        index = np.zeros(len(codes) - k + 1, dtype=np.int64)
        for j in range(k):
            index = index * 4 + np.maximum(codes[j:len(codes) - k + 1 + j], 0)
        want = np.bincount(index[ok], minlength=4 ** k)          # klib.py:160-168, vectorised
        assert np.array_equal(profile.counts, want), i
        n_seen += 1
    assert n_seen == 20_000


def test_by_record_balance_and_long_records():
    rng = random.Random(9)
    seqs = ["".join(rng.choice("ACGTN") for _ in range(n)) for n in (0, 5, 64, 65, 100_000, 3, 12_345)]
    codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
    for k in (2, 5, 9):
        rows = _cabi.count_by_record(codes, valid, n_bases, rec_starts, 0, len(seqs), k, balance=True)
        for seq, row in zip(seqs, rows):
            assert np.array_equal(row, ko.balance(ko.count_sequences([seq], k)))
        part = _cabi.count_by_record(codes, valid, n_bases, rec_starts, 2, 3, k)
        for seq, row in zip(seqs[2:5], part):
            assert np.array_equal(row, ko.count_sequences([seq], k))


def test_by_record_slab_kernel_limits():
    """by_record_kernel builds rows in 16-bit shared-memory slabs; records whose counts could
    pass 65535 (>= 65536 bases, >= 32768 with balance: a palindrome counts twice) take the RED
    rows instead.  Both sides of either limit, worst-case repeats, every slab count (k = 7: one
    slab of 16384 bins, k = 8: two, k = 9: eight), and the RED rows for everything."""
    seqs = ["A" * 65535, "A" * 65536, "AT" * 16383 + "A", "AT" * 16384, "T" * 32767, "ACGT" * 9000,
            "", "ACG", "N" * 500 + "ACGTTGCAAC" * 40]
    codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
    try:
        for path in (0, 1):
            _set_option("by_record_path", path)
            for k in (1, 2, 7, 8, 9):
                for balance in (False, True):
                    rows = _cabi.count_by_record(codes, valid, n_bases, rec_starts, 0, len(seqs), k,
                                                 balance=balance)
                    for seq, row in zip(seqs, rows):
                        want = ko.count_sequences([seq], k)
                        assert np.array_equal(row, ko.balance(want) if balance else want), (path, k, balance, seq[:8])
        # The host entry point copies the rows as uint16 when no record of the call can overflow
        # 16 bits, as int64 otherwise (cabi.cu kpal_count_by_record): sub-ranges on either side,
        # a count of exactly 65534 / 65535, and the narrow copy switched off.
        _set_option("by_record_path", 0)
        for narrow in (1, 0):
            _set_option("narrow_d2h", narrow)
            for k in (2, 8):
                for first, n in ((0, 1), (2, 7), (4, 1), (6, 3), (1, 2)):
                    for balance in (False, True):
                        rows = _cabi.count_by_record(codes, valid, n_bases, rec_starts, first, n, k, balance=balance)
                        for seq, row in zip(seqs[first:first + n], rows):
                            want = ko.count_sequences([seq], k)
                            assert np.array_equal(row, ko.balance(want) if balance else want), (narrow, k, first, balance)
    finally:
        _set_option("by_record_path", 0)
        _set_option("narrow_d2h", 1)


def test_device_api_counter_widths():
    """kpal_dev_count_packed with 32- and 64-bit counters + finalize."""
    L = _cabi.load()
    reads = random_reads(5, 5000, 200)
    codes, valid, _, n_bases = _cabi.pack_sequences([r.tobytes() for r in reads])
    k = 10
    bins = 4 ** k
    want = ko.count_sequences([r.tobytes().decode() for r in reads], k)
    d_codes = L.kpal_dev_alloc(codes.nbytes)
    d_valid = L.kpal_dev_alloc(valid.nbytes)
    d_counts = L.kpal_dev_alloc(bins * 8)
    d_bal = L.kpal_dev_alloc(bins * 8)
    try:
        _cabi.check(L.kpal_memcpy_h2d(d_codes, _cabi.ptr(codes), codes.nbytes, None))
        _cabi.check(L.kpal_memcpy_h2d(d_valid, _cabi.ptr(valid), valid.nbytes, None))
        for bits in (32, 64):
            d_table = L.kpal_dev_alloc(bins * bits // 8)
            zero = np.zeros(bins * bits // 8, dtype=np.uint8)
            _cabi.check(L.kpal_memcpy_h2d(d_table, _cabi.ptr(zero), zero.nbytes, None))
            _cabi.check(L.kpal_dev_count_packed(d_codes, d_valid, n_bases, k, d_table, bits, None))
            _cabi.check(L.kpal_dev_finalize_counts(d_table, bits, k, 0, d_counts, None))
            _cabi.check(L.kpal_dev_balance(d_counts, d_bal, k, None))
            out = np.empty(bins, dtype=np.int64)
            bal = np.empty(bins, dtype=np.int64)
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(out), d_counts, out.nbytes, None))
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(bal), d_bal, bal.nbytes, None))
            _cabi.check(L.kpal_stream_sync(None))
            L.kpal_dev_free(d_table)
            assert np.array_equal(out, want)
            assert np.array_equal(bal, ko.balance(want))
        assert L.kpal_kernel_launches() > 0
    finally:
        for p in (d_codes, d_valid, d_counts, d_bal):
            L.kpal_dev_free(p)


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (100 Mbp of 150-bp reads, k=12, balance):
    bit-exact against the multi-threaded C oracle, plus size-independent
    properties (balanced total = 2 x windows, symmetry under rc)."""
    k = 12
    reads = random_reads(2, 666_667, 150)
    fasta = reads_to_fasta(reads)
    got = _cabi.count_fasta(fasta, k, balance=True)
    plain = c_oracle.count_bytes(np.insert(reads, 150, ord("\n"), axis=1).tobytes(), k,
                                 threads=c_oracle.max_threads())
    assert got.sum() == 2 * plain.sum()
    rc = ko.reverse_complement_table(k)
    assert np.array_equal(got, got[rc])
    assert np.array_equal(got, plain + plain[rc])


# ------------------------------------------------------------- GPU FASTA packer
def _set_option(name, value):
    _cabi.check(_cabi.load().kpal_set_option(name.encode(), int(value)))


def _unpack(codes, valid, n_bases):
    pos = np.arange(n_bases)
    v = (valid[pos // 32] >> (31 - pos % 32).astype(np.uint32)) & 1
    c = (codes[pos // 16] >> (30 - 2 * (pos % 16)).astype(np.uint32)) & 3
    return np.where(v == 1, c, 4).astype(np.uint8)


FASTA_EDGE_TEXTS = [
    "", "\n", ">only\n", "no header at all\nACGT\n", ">a\nAC GT\r\nNN\n\n>\nTT\n>b z\n",
    ">x y z\n\n\nACGT", "\n\n>q\nA\n", "junk\n>r1\nACGT\nAC>GT\n>r2\n>r3\nGGGG",
    ">crlf\r\nACGT\r\nTTGA\r\n", "ACGT\n>late\nGGCC\n",
]


def test_dev_fasta_pack_stream_matches_host_packer(golden, tutorial_texts):
    """kpal_dev_fasta_pack emits the same base stream as the C++ packer (one
    separator in front of every record instead of behind it)."""
    L = _cabi.load()
    reads = random_reads(4, 3000, 150)
    texts = FASTA_EDGE_TEXTS + [golden["fasta_text"], tutorial_texts["c_2"],
                                reads_to_fasta(reads).decode()]
    for text in texts:
        raw = text.encode("latin-1")
        h_codes, h_valid, _, names, h_n = _cabi.fasta_pack(raw)
        host = _unpack(h_codes, h_valid, h_n)
        n = len(raw)
        cw, vw = ctypes.c_uint64(), ctypes.c_uint64()
        L.kpal_packed_words(n, ctypes.byref(cw), ctypes.byref(vw))
        d_text = L.kpal_dev_alloc(n + 32)
        d_codes = L.kpal_dev_alloc(cw.value * 4)
        d_valid = L.kpal_dev_alloc(vw.value * 4)
        d_scr = L.kpal_dev_alloc(L.kpal_fasta_scratch_bytes(n))
        try:
            if n:
                buf = np.frombuffer(raw, dtype=np.uint8)
                _cabi.check(L.kpal_memcpy_h2d(d_text, _cabi.ptr(buf), n, None))
            _cabi.check(L.kpal_dev_fasta_pack(d_text, n, d_codes, d_valid, d_scr, None))
            codes = np.empty(cw.value, dtype=np.uint32)
            valid = np.empty(vw.value, dtype=np.uint32)
            status = np.empty(3, dtype=np.uint64)
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(codes), d_codes, codes.nbytes, None))
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(valid), d_valid, valid.nbytes, None))
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(status), d_scr, 24, None))
            _cabi.check(L.kpal_stream_sync(None))
        finally:
            for p in (d_text, d_codes, d_valid, d_scr):
                L.kpal_dev_free(p)
        n_gpu = int(status[1])
        assert n_gpu == h_n, text[:40]
        assert int(status[2]) & 0xffffffff == 0
        gpu = _unpack(codes, valid, n_gpu)
        if n_gpu:
            assert gpu[0] == 4
            assert np.array_equal(np.append(gpu[1:], 4), host), text[:40]
        # everything past the packed stream stays zero (the count kernel reads it as invalid)
        assert not codes[(n_gpu + 15) // 16:].any() and not valid[(n_gpu + 31) // 32:].any()


def test_count_fasta_gpu_and_host_packers_agree(golden, tutorial_texts):
    reads = random_reads(8, 20000, 150)
    texts = FASTA_EDGE_TEXTS + [golden["fasta_text"], tutorial_texts["a_1"],
                                reads_to_fasta(reads).decode()]
    try:
        for text in texts:
            for k in (1, 5, 9, 12):
                _set_option("host_fasta", 0)
                gpu = _cabi.count_fasta(text, k, balance=True)
                _set_option("host_fasta", 1)
                host = _cabi.count_fasta(text, k, balance=True)
                assert np.array_equal(gpu, host), (text[:30], k)
                if len(text) < 100000:
                    assert np.array_equal(gpu, ko.balance(ko.count_fasta(text, k)))
    finally:
        _set_option("host_fasta", 0)


def test_count_fasta_exotic_whitespace_falls_back_to_host_packer():
    """Tabs on sequence lines need Python's rstrip() semantics: trailing ones are
    dropped (lines join), inner ones split k-mers.  The GPU packer flags them
    and the library re-packs on the host; the result stays exact."""
    text = ">a\nACGT\t\nACGT\n>b\nAC\tGT\nTTTT\x0b\n"
    assert ko.parse_fasta(text) == [("a", "ACGTACGT"), ("b", "AC\tGTTTTT")]
    for k in (2, 4, 6):
        assert np.array_equal(_cabi.count_fasta(text, k), ko.count_fasta(text, k))


# ------------------------------------------------- radix-partitioned count path
def _composition_bytes(seed, n, letters, p_n=0.0005):
    """One long text: `letters` uniformly, with sparse N and a newline every ~100 kb."""
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(letters.encode(), dtype=np.uint8)
    buf = lut[rng.integers(0, len(lut), n)].copy()
    buf[rng.random(n) < p_n] = ord("N")
    buf[rng.integers(0, n, max(1, n // 100_000))] = ord("\n")
    return buf


def _dev_count(text_bytes, k, bits, balance):
    """kpal_dev_count_packed + kpal_dev_finalize_counts on a packed stream."""
    L = _cabi.load()
    seqs = bytes(text_bytes).split(b"\n")
    codes, valid, _, n_bases = _cabi.pack_sequences(seqs)
    bins = 4 ** k
    d_codes = L.kpal_dev_alloc(codes.nbytes)
    d_valid = L.kpal_dev_alloc(valid.nbytes)
    d_table = L.kpal_dev_alloc(bins * bits // 8)
    d_counts = L.kpal_dev_alloc(bins * 8)
    try:
        zero = np.zeros(bins * bits // 8, dtype=np.uint8)
        _cabi.check(L.kpal_memcpy_h2d(d_codes, _cabi.ptr(codes), codes.nbytes, None))
        _cabi.check(L.kpal_memcpy_h2d(d_valid, _cabi.ptr(valid), valid.nbytes, None))
        _cabi.check(L.kpal_memcpy_h2d(d_table, _cabi.ptr(zero), zero.nbytes, None))
        _cabi.check(L.kpal_dev_count_packed(d_codes, d_valid, n_bases, k, d_table, bits, None))
        _cabi.check(L.kpal_dev_finalize_counts(d_table, bits, k, int(balance), d_counts, None))
        out = np.empty(bins, dtype=np.int64)
        _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(out), d_counts, out.nbytes, None))
        _cabi.check(L.kpal_stream_sync(None))
        # the same table through the device-table -> host-profile entry point (narrow copy)
        direct = np.full(bins, -1, dtype=np.int64)
        _cabi.check(L.kpal_dev_table_to_host(d_table, bits, k, int(balance), _cabi.ptr(direct), None))
        assert np.array_equal(direct, out), "kpal_dev_table_to_host differs from finalize + copy"
        return out
    finally:
        for p in (d_codes, d_valid, d_table, d_counts):
            L.kpal_dev_free(p)


@pytest.mark.parametrize("k,n,letters,payload_bits", [
    (12, 12_000_000, "ACGT", 0),       # several tiles per CTA, default geometry (512 buckets)
    (12, 3_000_000, "ACGT", 14),       # 1024 buckets
    (12, 3_000_000, "ACG", 0),         # skew: some slots and regions overflow into the RED path
    (12, 2_000_000, "AC", 0),          # 32 of 512 buckets used: heavy overflow
    (12, 1_000_000, "A", 0),           # one bin: everything overflows
    (13, 6_000_000, "ACGTacgt", 0),    # 2048 buckets: one pass-1 launch (or two of 1024 buckets each)
    (13, 1_000_000, "AT", 0),
    (14, 3_000_000, "ACGT", 0),        # 8192 buckets, 4 (or 8) sweeps
    (12, 2_000_000, "ACGT", 13),       # forced 2048 buckets at k = 12
    (11, 2_000_000, "ACGT", 0),
    (10, 2_000_000, "ACGT", 0),
    (9, 2_000_000, "ACGT", 0),
    (9, 500_000, "CG", 15),            # 8 buckets of 2^15 bins
    (12, 70, "ACGT", 0),               # far less than one tile
])
def test_radix_count_path_bit_exact(k, n, letters, payload_bits):
    """The one-window two-pass radix path (count_radix.cu) forced on: bit-exact against the C
    oracle for uniform, skewed and degenerate compositions, 32- and 64-bit counters."""
    text = _composition_bytes(k * 7919 + n, n, letters)
    want = c_oracle.count_bytes(text, k, threads=c_oracle.max_threads())
    try:
        _set_option("count_path", 3)
        _set_option("radix_payload_bits", payload_bits)
        many_buckets = k >= 13 or payload_bits == 13
        for max_buckets in ((2048, 1024) if many_buckets else (2048,)):   # buckets per pass-1 launch
            _set_option("radix_max_buckets", max_buckets)
            for shape in (1, 2):           # one 1024-thread CTA per SM / two 512-thread CTAs per SM
                _set_option("radix_shape", shape)
                got32 = _dev_count(text, k, 32, False)
                assert np.array_equal(got32, want), (max_buckets, shape)
                got64 = _dev_count(text, k, 64, True)
                assert np.array_equal(got64, ko.balance(want)), (max_buckets, shape)
        _set_option("count_path", 1)
        assert np.array_equal(_dev_count(text, k, 32, False), want)
    finally:
        _set_option("count_path", 0)
        _set_option("radix_payload_bits", 0)
        _set_option("radix_shape", 0)


@pytest.mark.parametrize("k,n,letters,p_n", [
    (12, 40_000_000, "ACGT", 0.0005),  # ~9 tiles per CTA, several flushes per slot
    (12, 5_000_000, "ACGTacgt", 0.05), # short runs: many windows without a valid partner
    (12, 3_000_000, "ACG", 0.0005),    # skew: slots and regions overflow into the RED path
    (12, 2_000_000, "AC", 0.0005),     # 32 of 1024 buckets used: heavy overflow
    (12, 1_000_000, "A", 0.0005),      # one bin
    (12, 300_000, "AT", 0.3),          # runs barely longer than k
    (11, 4_000_000, "ACGT", 0.001),
    (10, 4_000_000, "ACGT", 0.001),
    (9, 4_000_000, "ACGT", 0.001),
    (9, 500_000, "CG", 0.001),
    (12, 70, "ACGT", 0.0),             # far less than one tile
    (12, 13, "ACGT", 0.0),             # one pair
    (12, 12, "ACGT", 0.0),             # one window, no pair
])
def test_pair_count_path_bit_exact(k, n, letters, p_n):
    """The two-windows-per-payload radix path (count_pairs.cu) forced on: bit-exact against
    the C oracle, 32- and 64-bit counters, every tiling / flush / pass-2 variant."""
    text = _composition_bytes(k * 104729 + n, n, letters, p_n=p_n)
    want = c_oracle.count_bytes(text, k, threads=c_oracle.max_threads())
    try:
        _set_option("count_path", 2)
        # tiles between two slot flushes (0 = automatic; 6 overflows the slots into the RED path),
        # pass 2 as one launch flushed by the TMA unit (cp.reduce.async.bulk) or as two launches per role
        for flush_every, fused in ((0, 1), (1, 0), (6, 1), (2, 0)):
            _set_option("pair_flush_every", flush_every)
            _set_option("pair_fused", fused)
            got32 = _dev_count(text, k, 32, False)
            assert np.array_equal(got32, want), (flush_every, fused)
            got64 = _dev_count(text, k, 64, True)
            assert np.array_equal(got64, ko.balance(want)), (flush_every, fused)
    finally:
        _set_option("count_path", 0)
        _set_option("pair_flush_every", 0)
        _set_option("pair_fused", 1)


@pytest.mark.parametrize("k,n", [(12, 30_000_000), (12, 3_000), (10, 20_000_000), (6, 100_000), (13, 9_000_000)])
def test_count_packed_fresh_zeroes_the_table(k, n):
    """kpal_dev_count_packed_fresh on a table full of garbage: the call zeroes it first, whichever
    count path the input takes; repeated calls."""
    L = _cabi.load()
    text = _composition_bytes(k + n, n, "ACGTacgt", p_n=0.01)
    want = c_oracle.count_bytes(text, k, threads=c_oracle.max_threads())
    codes, valid, _, n_bases = _cabi.pack_sequences(bytes(text).split(b"\n"))
    bins = 4 ** k
    d_codes, d_valid = L.kpal_dev_alloc(codes.nbytes), L.kpal_dev_alloc(valid.nbytes)
    d_table, d_counts = L.kpal_dev_alloc(bins * 4), L.kpal_dev_alloc(bins * 8)
    try:
        _cabi.check(L.kpal_memcpy_h2d(d_codes, _cabi.ptr(codes), codes.nbytes, None))
        _cabi.check(L.kpal_memcpy_h2d(d_valid, _cabi.ptr(valid), valid.nbytes, None))
        garbage = np.full(bins, 0xdeadbeef, dtype=np.uint32)
        for round_ in range(3):
            _cabi.check(L.kpal_memcpy_h2d(d_table, _cabi.ptr(garbage), garbage.nbytes, None))
            _cabi.check(L.kpal_dev_count_packed_fresh(d_codes, d_valid, n_bases, k, d_table, 32, None))
            _cabi.check(L.kpal_dev_finalize_counts(d_table, 32, k, 0, d_counts, None))
            out = np.empty(bins, dtype=np.int64)
            _cabi.check(L.kpal_memcpy_d2h(_cabi.ptr(out), d_counts, out.nbytes, None))
            _cabi.check(L.kpal_stream_sync(None))
            assert np.array_equal(out, want), round_
    finally:
        for ptr in (d_codes, d_valid, d_table, d_counts):
            L.kpal_dev_free(ptr)


def test_pair_count_path_reads():
    """150-bp reads (every record ends a run: the windows without a partner go through
    the RED path) at the default path selection, k = 12 and 10."""
    reads = random_reads(78, 140_000, 150)
    text = np.insert(reads, 150, ord("\n"), axis=1).tobytes()
    for k in (12, 10):
        want = c_oracle.count_bytes(text, k, threads=c_oracle.max_threads())
        assert np.array_equal(_dev_count(text, k, 32, False), want)


def test_radix_count_path_through_host_api():
    """count_path=2 under kpal_count_fasta / kpal_count_sequences (records, N, lower case)."""
    reads = random_reads(77, 30_000, 150)
    fasta = reads_to_fasta(reads)
    want = c_oracle.count_bytes(np.insert(reads, 150, ord("\n"), axis=1).tobytes(), 12,
                                threads=c_oracle.max_threads())
    try:
        _set_option("count_path", 2)
        assert np.array_equal(_cabi.count_fasta(fasta, 12, balance=True), ko.balance(want))
        seqs = [r.tobytes().decode() for r in reads[:2000]]
        assert np.array_equal(_cabi.count_sequences(seqs, 10), ko.count_sequences(seqs, 10))
        assert not _cabi.count_sequences([], 12).any()
        assert not _cabi.count_sequences(["ACGTN" * 2], 12).any()
    finally:
        _set_option("count_path", 0)


@pytest.mark.parametrize("k", [6, 7, 8, 9, 12, 13])
def test_tiled_balance_finalize_matches_plain(k):
    """finalize_balance_tiled_kernel (shared-memory transposition) against the plain gather
    kernel and the oracle, on a dense random table incl. palindromic middles."""
    rng = np.random.default_rng(k)
    n = min(4 ** k * 3, 40_000_000)
    text = _composition_bytes(k, n, "ACGT")
    want = ko.balance(c_oracle.count_bytes(text, k, threads=c_oracle.max_threads()))
    try:
        for bits in (32, 64):
            _set_option("tiled_finalize", 1)
            tiled = _dev_count(text, k, bits, True)
            _set_option("tiled_finalize", 0)
            plain = _dev_count(text, k, bits, True)
            assert np.array_equal(tiled, want)
            assert np.array_equal(plain, want)
    finally:
        _set_option("tiled_finalize", 1)


def test_radix_count_path_k15():
    """Largest supported k: 32768 buckets, 32 pass-1 sweeps (32-bit counters only: the
    int64 result alone is 8 GiB)."""
    k = 15
    text = _composition_bytes(15, 1_500_000, "ACGT")
    want = c_oracle.count_bytes(text, k, threads=c_oracle.max_threads())
    try:
        _set_option("count_path", 2)
        got = _dev_count(text, k, 32, False)
    finally:
        _set_option("count_path", 0)
    assert np.array_equal(got, want)


def test_count_fasta_chunked_upload_pipeline(tutorial_texts):
    """kpal_count_fasta uploads the text in chunks and packs every chunk as it lands; the
    scans carry their state from chunk to chunk.  Any chunk count must give the same bits
    (chunk boundaries fall inside headers, lines and records)."""
    reads = random_reads(21, 4000, 150)
    texts = [reads_to_fasta(reads).decode(), tutorial_texts["b_1"],
             "leading junk\n" * 400 + ">late header\n" + "ACGTNacgt" * 3000 + "\n>x\n" + "G" * 9000,
             ">only one long line\n" + "ACGGT" * 20000]
    try:
        for text in texts:
            want = {k: ko.balance(ko.count_fasta(text, k)) for k in (3, 9)}
            for chunks in (1, 2, 7, 16, 32):
                _set_option("fasta_chunks", chunks)
                for k in (3, 9):
                    assert np.array_equal(_cabi.count_fasta(text, k, balance=True), want[k]), (chunks, k)
    finally:
        _set_option("fasta_chunks", 0)


def test_large_fasta_is_counted_in_two_overlapped_parts():
    """With the option fasta_split, kpal_count_fasta cuts a text of 16 MB or more at a header
    line near three quarters (cabi.cu fasta_gpu_count) and counts the first part while the
    second is uploaded.  Same
    bits with the cut, without it, for every chunking; '>' inside header lines is not a cut
    point; a text without a header in the search window is not cut; exotic white space in
    either part still sends the whole file through the host packer."""
    reads = random_reads(21, 130_000, 150)
    fasta = reads_to_fasta(reads)                                    # 20.9 MB, a header every 161 bytes
    assert len(fasta) > (16 << 20)
    want = {k: c_oracle.count_bytes(np.insert(reads, 150, ord("\n"), axis=1).tobytes(), k,
                                    threads=c_oracle.max_threads()) for k in (6, 12)}
    tricky = fasta.replace(b">r0097", b">r>>97")                      # '>' runs inside the header lines around 75 %
    genome = b">chr\n" + reads[:120_000].tobytes() + b"\n>tail\nACGTACGTAC\n"      # no header near 75 %
    want_genome = ko.count_fasta(genome.decode(), 12)
    tab_late = fasta[:-20] + b"\tAC\t\n" + fasta[-20:]               # a tab in the second part
    want_tab = ko.count_fasta(tab_late.decode(), 6)
    try:
        for split in (1, 0):
            _set_option("fasta_split", split)
            for chunks in (0, 2, 5, 32):
                _set_option("fasta_chunks", chunks)
                for k in (6, 12):
                    assert np.array_equal(_cabi.count_fasta(fasta, k), want[k]), (split, chunks, k)
                assert np.array_equal(_cabi.count_fasta(tricky, 12), want[12]), (split, chunks)
            _set_option("fasta_chunks", 0)
            assert np.array_equal(_cabi.count_fasta(fasta, 12, balance=True), ko.balance(want[12])), split
            assert np.array_equal(_cabi.count_fasta(genome, 12), want_genome), split
            assert np.array_equal(_cabi.count_fasta(tab_late, 6), want_tab), split
    finally:
        _set_option("fasta_split", 0)
        _set_option("fasta_chunks", 0)


def test_hybrid_upload_of_a_large_fasta():
    """A text of 32 MB or more is uploaded in hybrid form (cabi.cu fasta_hybrid_count): the
    head raw, packed by the device, while host threads pack the segments of the tail
    (pack.cpp, csrc/slotted.h) into slots of the same stream.  The host's share adapts to the
    machine from call to call, so the test also forces it.  Same bits as the C oracle, and as
    the plain upload."""
    rng = np.random.default_rng(77)
    n_reads, read_len = 240_000, 150
    reads = random_reads(77, n_reads, read_len, p_n=0.002, p_lower=0.1)
    reads[rng.random(reads.shape) < 0.0005] = ord(" ")              # blanks inside lines are dropped
    extra = np.frombuffer(b" \r", dtype=np.uint8)[rng.integers(0, 2, n_reads)]      # 'ACGT \n' and 'ACGT\r\n' lines
    header = np.frombuffer(b">r0000000 x\n", dtype=np.uint8)
    rows = np.empty((n_reads, len(header) + read_len + 2), dtype=np.uint8)
    rows[:, :len(header)] = header
    for d in range(7):
        rows[:, 8 - d] = ord("0") + (np.arange(n_reads) // 10 ** d) % 10
    rows[:, len(header):-2] = reads
    rows[:, -2] = extra
    rows[:, -1] = ord("\n")
    fasta = b"junk before the first header\nACGT\n" + rows.tobytes()
    assert len(fasta) > (32 << 20)
    body = rows[:, len(header):].reshape(-1)
    body = body[(body != ord(" ")) & (body != ord("\r"))].tobytes()    # what the reader keeps, one read per line
    want = {k: c_oracle.count_bytes(body, k, threads=c_oracle.max_threads()) for k in (6, 12)}
    tricky = fasta.replace(b">r0097", b">r>>97").replace(b">r0200", b">r>>00")      # '>' inside header lines
    tab_late = fasta[:-20] + b"\tAC\t\n" + fasta[-20:]             # tabs in the host's part: packed exactly there
    keep = lambda line: line.replace(b" ", b"").replace(b"\r", b"")
    body_tab = (rows[:-1, len(header):].reshape(-1)[(rows[:-1, len(header):].reshape(-1) != ord(" "))
                                                     & (rows[:-1, len(header):].reshape(-1) != ord("\r"))].tobytes()
                + keep(reads[-1, :132].tobytes()) + b"\tAC" + keep(reads[-1, 132:].tobytes()) + b"\n")   # the lines of a record join
    tab_early = fasta[:5000] + fasta[5000:].replace(b"\n>", b"\t\n>", 1)    # a trailing tab in the device's part: whole-file fallback
    genome = b">chr\n" + reads.tobytes() + b"\n>tail\nACGTACGTAC\n"      # no header lines to cut at
    want_genome = c_oracle.count_bytes(reads.tobytes().replace(b" ", b"") + b"\nACGTACGTAC\n", 12, threads=c_oracle.max_threads())
    # 70-column records far longer than a segment: every cut falls inside a record, the junction
    # records (csrc/slotted.h) restore the windows that cross the cuts
    flat = reads.reshape(-1)
    half = (flat.size // 140) * 70
    wrap = lambda seq: np.concatenate([seq.reshape(-1, 70), np.full((seq.size // 70, 1), 10, dtype=np.uint8)], axis=1).tobytes()
    wrapped = b">chr1 first\n" + wrap(flat[:half]) + b">chr2\n" + wrap(flat[half:2 * half]) + b"ACGTTGCA"
    unwrapped = flat[:half].tobytes().replace(b" ", b"") + b"\n" + flat[half:2 * half].tobytes().replace(b" ", b"") + b"ACGTTGCA\n"
    try:
        for hybrid, share in ((1, 0), (0, 0), (2, 0), (64, 10), (5, 50), (1, 90), (1, 1), (1, 0)):
            _set_option("fasta_hybrid", hybrid)
            _set_option("fasta_hybrid_share", share)         # percent of the text for the host (0: adaptive)
            for k in (6, 12):
                assert np.array_equal(_cabi.count_fasta(fasta, k), want[k]), (hybrid, share, k)
            assert np.array_equal(_cabi.count_fasta(tricky, 12), want[12]), (hybrid, share)
        _set_option("fasta_hybrid", 1)
        assert np.array_equal(_cabi.count_fasta(fasta, 12, balance=True), ko.balance(want[12]))
        assert np.array_equal(_cabi.count_fasta(tab_late, 12), c_oracle.count_bytes(body_tab, 12, threads=c_oracle.max_threads()))
        assert np.array_equal(_cabi.count_fasta(tab_early, 6), want[6])
        assert np.array_equal(_cabi.count_fasta(genome, 12), want_genome)
        for k in (6, 12, 13):
            want_wrapped = c_oracle.count_bytes(unwrapped, k, threads=c_oracle.max_threads())
            for share in (0, 15, 80):
                _set_option("fasta_hybrid_share", share)
                assert np.array_equal(_cabi.count_fasta(wrapped, k), want_wrapped), (k, share)
        _set_option("fasta_hybrid_share", 0)
        # text handles go the same way
        profile = klib.Profile.from_fasta(io.BytesIO(fasta), 6)
        assert np.array_equal(profile.counts, want[6])
    finally:
        _set_option("fasta_hybrid", 1)
        _set_option("fasta_hybrid_share", 0)


def test_narrow_profile_copy_and_its_overflow_path():
    """From k = 10 on the host entry points move the profile over PCIe as uint8 or uint16
    (whichever holds every count) and widen it on the host (cabi.cu finalize_to_host).
    Same bits as the plain int64 copy, for counts that fit -- and for counts that do not
    (a repetitive input pushes one bin past 65535: the device flags send the call down the
    int64 copy -- unless the counts that do not fit are few: they then travel in a side list and
    the host patches their bins, option "narrow_lists").  narrow_d2h: 1 = uint8 / uint16,
    2 = uint16 only, 0 = int64."""
    reads = random_reads(5, 20_000, 150)
    fasta = reads_to_fasta(reads)
    seqs = [r.tobytes().decode() for r in reads]
    repetitive = ">poly\n" + "A" * 70_000 + "\n>mix\n" + "ACGTTGCA" * 30_000 + "\n" + fasta.decode()
    # 3000 different 120-mers, 300 copies each: ~330 000 bins above 255 at k = 10, 11 -- more than
    # the side list of the uint8 form holds, so the profile travels as uint16
    rng = np.random.default_rng(55)
    units = ["".join("ACGT"[c] for c in rng.integers(0, 4, 120)) for _ in range(3000)]
    many_big = "".join(">u%d\n%s\n" % (i, "N".join([u] * 300)) for i, u in enumerate(units))
    try:
        for k in (10, 11):
            want_big = ko.count_fasta(many_big, k)
            assert (want_big > 255).sum() > 4 ** k // 32 and want_big.max() <= 65535
            for lists in (1, 0):
                _set_option("narrow_lists", lists)
                assert np.array_equal(_cabi.count_fasta(many_big, k), want_big), (k, lists)
                assert np.array_equal(_cabi.count_fasta(many_big, k, balance=True), ko.balance(want_big)), (k, lists)
        for k in (10, 11, 12):
            want = ko.count_sequences(seqs, k)
            want_rep = ko.count_fasta(repetitive, k)
            assert want_rep.max() > 65535 and ko.balance(want_rep).max() > 65535
            assert ko.balance(want).max() <= 255            # random reads: the uint8 copy under narrow_d2h = 1
            for narrow, lists in ((1, 1), (2, 1), (0, 1), (1, 0), (2, 0)):
                _set_option("narrow_d2h", narrow)
                _set_option("narrow_lists", lists)        # 0: a count that does not fit sends the whole profile wider
                for balance in (False, True):
                    w = ko.balance(want) if balance else want
                    assert np.array_equal(_cabi.count_fasta(fasta, k, balance=balance), w), (k, narrow, balance)
                    assert np.array_equal(_cabi.count_sequences(seqs, k, balance=balance), w), (k, narrow, balance)
                    wr = ko.balance(want_rep) if balance else want_rep
                    assert np.array_equal(_cabi.count_fasta(repetitive, k, balance=balance), wr), (k, narrow, balance)
            # exactly 255 / 65535 fits, 256 / 65536 does not: both sides of either threshold
            for n_a in (255 + k - 1, 256 + k - 1, 65535 + k - 1, 65536 + k - 1):
                _set_option("narrow_d2h", 1)
                got = _cabi.count_sequences(["A" * n_a, "ACGT" * 50], k)
                assert got[0] == n_a - k + 1 and np.array_equal(got, ko.count_sequences(["A" * n_a, "ACGT" * 50], k))
    finally:
        _set_option("narrow_d2h", 1)
        _set_option("narrow_lists", 1)


def test_narrow_copy_into_a_pinned_profile_with_dma_share():
    """Pinned destination: the last dma_share/16 of the profile is copied as int64 by the DMA
    engine behind the narrow chunks, the rest is widened by the host threads (cabi.cu
    finalize_to_host).  Every share, every width, counts above 255 on either side of the
    split, and the int64 fallback -- always the bits of the oracle."""
    reads = random_reads(6, 20_000, 150)
    fasta = reads_to_fasta(reads).decode()
    low_rep = ">a\n" + "A" * 400 + "\n" + fasta                 # bin 0 (narrow side) above 255
    high_rep = ">t\n" + "T" * 400 + "\n" + fasta                # last bin (int64 side) above 255
    big_rep = ">t\n" + "T" * 70_000 + "\n" + fasta              # ... above 65535: stays narrow-copied
    fallback = ">a\n" + "A" * 70_000 + "\n" + fasta             # narrow side above 65535: int64 copy
    try:
        for k in (10, 12):
            pinned = _cabi.PinnedArray(4 ** k, np.int64)
            for text in (fasta, low_rep, high_rep, big_rep, fallback):
                want = ko.count_fasta(text, k)
                for share in (0, 1, 3, 8):
                    _set_option("dma_share", share)
                    for narrow in (1, 2):
                        _set_option("narrow_d2h", narrow)
                        pinned.array[:] = -1
                        got = _cabi.count_fasta(text, k, out=pinned.array)
                        assert np.array_equal(got, want), (k, share, narrow)
                        assert np.array_equal(_cabi.count_fasta(text, k), want), (k, share, narrow, "pageable")
            pinned.free()
    finally:
        _set_option("narrow_d2h", 1)
        _set_option("dma_share", 0)
