"""
CPU tests of the C-ABI boundary: the library loads, exports every symbol the
header declares, the host packers agree with the oracle's reading of the same
text, and -- without a GPU -- every compute entry point fails loudly instead
of falling back to a CPU path.
"""
import ctypes
import os
import random
import re

import numpy as np
import pytest

from conftest import ROOT
from kpal_b200 import _cabi
from oracle import kpal_oracle as ko


def header_symbols():
    with open(os.path.join(ROOT, "include", "kpal_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kpal_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _cabi.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "libkpal_b200.so does not export %s" % name
    assert sorted(_cabi.SYMBOLS) == declared
    assert L.kpal_abi_version() == 1


def unpack(codes, valid, n_bases):
    """Packed stream -> list of codes (0-3, or 4 for invalid), test helper."""
    out = []
    for p in range(n_bases):
        c = (int(codes[p // 16]) >> (30 - 2 * (p % 16))) & 3
        v = (int(valid[p // 32]) >> (31 - (p % 32))) & 1
        out.append(c if v else 4)
    return out


def expected_stream(seqs):
    lut = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = []
    for s in seqs:
        out.extend(lut.get(ch.upper(), 4) for ch in s)
        out.append(4)
    return out


def test_pack_sequences_matches_text():
    rng = random.Random(5)
    for _ in range(40):
        seqs = ["".join(rng.choice("ACGTacgtNn-x") for _ in range(rng.randint(0, 150)))
                for _ in range(rng.randint(0, 7))]
        codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
        assert n_bases == sum(len(s) + 1 for s in seqs)
        assert unpack(codes, valid, n_bases) == expected_stream(seqs)
        starts = np.cumsum([0] + [len(s) + 1 for s in seqs])
        assert rec_starts.tolist() == starts.tolist()
        # padding stays zero and one halo chunk is present
        assert len(codes) == ((n_bases + 63) // 64 + 1) * 4
        assert not codes[(n_bases + 15) // 16:].any()


def test_pack_sequences_large_parallel():
    # > 1 MiB so that several worker threads cut the output stream
    rng = np.random.default_rng(1)
    seqs = []
    for n in (3_000_001, 0, 17, 2_500_000, 1):
        seqs.append(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes().decode())
    codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
    flat = np.array(expected_stream(seqs), dtype=np.uint8)
    pos = np.arange(n_bases)
    got_v = (valid[pos // 32] >> (31 - pos % 32).astype(np.uint32)) & 1
    got_c = (codes[pos // 16] >> (30 - 2 * (pos % 16)).astype(np.uint32)) & 3
    assert np.array_equal(got_v == 1, flat != 4)
    assert np.array_equal(got_c[flat != 4], flat[flat != 4])
    assert not got_c[flat == 4].any()


def test_fasta_pack_matches_oracle_reader(golden, tutorial_texts):
    texts = [golden["fasta_text"], tutorial_texts["a_1"], "", "no header at all\nACGT\n",
             ">only\n", ">a\nAC GT\r\nNN\n\n>\nTT\n>b\tz\n", ">x y z\n\n\nACGT", "\n\n>q\nA\n"]
    for text in texts:
        records = ko.parse_fasta(text)
        codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(text)
        assert names == [n for n, _ in records]
        assert unpack(codes, valid, n_bases) == expected_stream([s for _, s in records])
        assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for _, s in records]).tolist()


def test_fasta_pack_large_parallel():
    rng = np.random.default_rng(2)
    recs = []
    lines = []
    for i in range(3000):
        seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8),
                         size=int(rng.integers(0, 2500))).tobytes().decode()
        recs.append(("r%07d" % i, seq))
        lines.append(">r%07d some text" % i)
        lines.extend(seq[p:p + 70] for p in range(0, len(seq), 70))
    text = "\n".join(lines) + "\n"
    assert len(text) > (3 << 20)
    codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(text)
    assert names == [n for n, _ in recs]
    assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for _, s in recs]).tolist()
    flat = np.array(expected_stream([s for _, s in recs]), dtype=np.uint8)
    pos = np.arange(n_bases)
    got_v = (valid[pos // 32] >> (31 - pos % 32).astype(np.uint32)) & 1
    got_c = (codes[pos // 16] >> (30 - 2 * (pos % 16)).astype(np.uint32)) & 3
    assert np.array_equal(got_v == 1, flat != 4)
    assert np.array_equal(got_c[flat != 4], flat[flat != 4])


def test_argument_errors_map_to_valueerror():
    with pytest.raises(ValueError):
        _cabi._check_k(0)
    with pytest.raises(ValueError):
        _cabi._check_k(16)
    with pytest.raises(ValueError):
        _cabi._k_of(20)
    L = _cabi.load()
    assert L.kpal_dev_count_packed(None, None, 10, 3, None, 32, None) == _cabi.KPAL_EINVAL
    assert b"null" in L.kpal_last_error()


def test_format_matrix_equals_python_format():
    """kpal_format_matrix writes the lower triangle exactly as the reference's
    '{0:.{precision}f}'.format loop (kpal/kdistlib.py:179-186), including nan,
    inf, negative zero, exact ties and huge values."""
    rng = np.random.default_rng(11)
    for n, precision in ((1, 3), (2, 10), (3, 2), (9, 0), (40, 10), (130, 3), (25, 17)):
        v = rng.random((n, n)) * rng.choice([1e-12, 1.0, 1e6, 1e15], (n, n))
        if n > 1:
            v[1, 0] = float('nan')
        if n > 2:
            v[2, 0], v[2, 1] = float('inf'), -0.0
        if n > 5:
            v[5, :5] = [-np.nan, -np.inf, 0.5, 2.5, 1e22]
            v[4, :4] = [0.125, 0.375, 1e-320, 123456789.987654321]
        template = '{{0:.{0}f}}'.format(precision)
        want = ''.join(' '.join(template.format(v[i, j]) for j in range(i)) + '\n'
                       for i in range(1, n))
        assert _cabi.format_matrix(v, precision) == want
        # a view with a larger leading dimension (sub-block of a bigger matrix)
        big = np.zeros((n + 3, n + 5))
        big[:n, :n] = v
        assert _cabi.format_matrix(big[:n, :n], precision) == want
    L = _cabi.load()
    length = ctypes.c_uint64()
    v = np.ones((3, 3))
    assert L.kpal_format_matrix(_cabi.ptr(v), 3, 3, 2, None, 0, ctypes.byref(length)) == _cabi.KPAL_EOVERFLOW
    assert length.value == len('1.00\n1.00 1.00\n')
    assert L.kpal_format_matrix(_cabi.ptr(v), 3, 2, 2, None, 0, ctypes.byref(length)) == _cabi.KPAL_EINVAL


@pytest.mark.parametrize("width", (2, 1))
def test_widen_stage_is_exact(width):
    """The host stage of the narrow profile copy (kpal_widen_u16 / kpal_widen_u8): narrow
    counts -> the int64 counts of Profile.counts (klib.py:170), chunked, any alignment /
    length."""
    L = _cabi.load()
    fn = L.kpal_widen_u16 if width == 2 else L.kpal_widen_u8
    dtype, top = (np.uint16, 65535) if width == 2 else (np.uint8, 255)
    rng = np.random.default_rng(11)
    for n, chunk, offset in ((1, 0, 0), (7, 3, 1), (1000, 0, 0), (2500, 1000, 1), (1 << 20, 1 << 17, 0),
                             ((1 << 22) + 5, 1 << 18, 1), (1 << 16, 1 << 20, 0)):
        narrow = rng.integers(0, top + 1, n, dtype=dtype)
        narrow[:3] = (top, 0, 1)[:min(3, n)]
        backing = np.full(n + 2, -1, dtype=np.int64)
        out = backing[offset:offset + n]            # offset 1: only 8-byte aligned
        for _ in range(3):                          # the pool is reused from call to call
            out[:] = -1
            assert fn(_cabi.ptr(narrow), n, chunk, out.ctypes.data) == _cabi.KPAL_OK
            assert np.array_equal(out, narrow.astype(np.int64))
        assert backing[offset + n] == -1 and (offset == 0 or backing[0] == -1)
    assert fn(None, 0, 0, None) == _cabi.KPAL_OK
    assert fn(None, 4, 0, None) == _cabi.KPAL_EINVAL


def test_widen_u16_from_many_threads():
    """Concurrent callers share one worker pool; every call must still be exact."""
    import threading
    L = _cabi.load()
    rng = np.random.default_rng(12)
    srcs = [rng.integers(0, 65536, (1 << 18) + i, dtype=np.uint16) for i in range(6)]
    outs = [np.empty(s.size, dtype=np.int64) for s in srcs]

    def run(i):
        for _ in range(5):
            assert L.kpal_widen_u16(_cabi.ptr(srcs[i]), srcs[i].size, 1 << 15, outs[i].ctypes.data) == 0
    threads = [threading.Thread(target=run, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for s, o in zip(srcs, outs):
        assert np.array_equal(o, s.astype(np.int64))


def test_row_stats_are_numpy_bit_for_bit():
    """kpal_row_stats (host threads): total, non_zero, mean, median, std of every row, bit-identical
    to the NumPy calls behind the reference's Profile properties (kpal/klib.py:192-225)."""
    rng = np.random.default_rng(21)
    for shape, high in (((40, 65536), 3), ((9, 4096), 50), ((5, 1024), 10 ** 6), ((12, 16), 4), ((3, 4 ** 9), 2)):
        rows = rng.integers(0, high, shape).astype(np.int64)
        rows[0] = 0
        rows[-1, ::5] = 10 ** 10
        stats = _cabi.row_stats(rows)
        for r, x in enumerate(rows):
            want = (float(x.sum()), float(np.count_nonzero(x)), x.mean(), np.median(x), x.std())
            assert tuple(stats[r]) == want, (shape, r)


def test_deflate_chunks_are_zlib_streams():
    import zlib
    rng = np.random.default_rng(22)
    rows = rng.poisson(0.02, (33, 16384)).astype(np.int64)
    for chunk_bytes, level in ((16384, 4), (131072, 1), (8 * 16384, 9)):
        blob, sizes = _cabi.deflate_chunks_packed(rows, chunk_bytes, level)
        raw = rows.view(np.uint8).reshape(-1)
        at = 0
        for c, size in enumerate(sizes):
            stream = bytes(blob[at:at + size])
            assert zlib.decompress(stream) == raw[c * chunk_bytes:(c + 1) * chunk_bytes].tobytes()
            assert stream == zlib.compress(raw[c * chunk_bytes:(c + 1) * chunk_bytes].tobytes(), level)
            at += size
        assert at == blob.size
    with pytest.raises(ValueError):
        _cabi.deflate_chunks_packed(rows, 1000, 4)


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product path must refuse, not emulate."""
    from kpal_b200 import klib, kdistlib
    with pytest.raises(RuntimeError):
        klib.Profile.from_sequences(["ACGTACGT"], 2)
    p = klib.Profile(np.zeros(16, dtype=np.int64))
    with pytest.raises(RuntimeError):
        p.balance()
    with pytest.raises(RuntimeError):
        kdistlib.ProfileDistance().distance(p, p)
    L = _cabi.load()
    out = np.zeros(16, dtype=np.int64)
    blob = b"ACGT"
    off = np.array([0, 4], dtype=np.uint64)
    rc = L.kpal_count_sequences(ctypes.c_char_p(blob), _cabi.ptr(off), 1, 2, 0, _cabi.ptr(out))
    assert rc == _cabi.KPAL_ECUDA
    assert b"no CPU fallback" in L.kpal_last_error()
