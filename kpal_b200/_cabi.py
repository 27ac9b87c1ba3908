"""
ctypes binding of ``libkpal_b200.so`` (C ABI declared in ``include/kpal_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing or no
CUDA device is visible, every compute call raises.  Argument errors map to
``ValueError`` (as the reference raises for bad input), CUDA failures to
``RuntimeError``, allocation failures to ``MemoryError``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: KPAL_B200_LIB selects another build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("KPAL_B200_LIB") or os.path.join(_HERE, "libkpal_b200.so")

KPAL_OK, KPAL_EINVAL, KPAL_ECUDA, KPAL_ENOMEM, KPAL_EOVERFLOW = range(5)
METRICS = {"multiset": 0, "euclidean": 1, "cosine": 2}
PAIRWISE = {"prod": 0, "sum": 1}
MAX_K = 15

#: every symbol include/kpal_b200.h declares (checked by tests/test_cabi.py)
SYMBOLS = (
    "kpal_abi_version", "kpal_last_error", "kpal_device_count", "kpal_set_device", "kpal_get_device",
    "kpal_host_alloc", "kpal_host_free", "kpal_dev_alloc", "kpal_dev_free",
    "kpal_memcpy_h2d", "kpal_memcpy_d2h", "kpal_dev_memset", "kpal_stream_sync",
    "kpal_packed_words", "kpal_pack_sequences", "kpal_fasta_scan", "kpal_fasta_pack", "kpal_fasta_pack_segment", "kpal_fasta_slotted_bases", "kpal_fasta_pack_slotted",
    "kpal_count_sequences", "kpal_count_fasta", "kpal_count_by_record", "kpal_balance",
    "kpal_distance_matrix", "kpal_pair_distance",
    "kpal_matrix_open", "kpal_matrix_push", "kpal_matrix_finish", "kpal_matrix_close",
    "kpal_format_matrix", "kpal_widen_u16", "kpal_widen_u8", "kpal_pair_distance_positive",
    "kpal_row_stats", "kpal_deflate_bound", "kpal_deflate_chunks", "kpal_deflate_chunks_sparse", "kpal_deflate_packed_begin", "kpal_deflate_packed_finish", "kpal_rows_stats_deflate_begin", "kpal_compact_slots",
    "kpal_split_length", "kpal_split", "kpal_show_balance",
    "kpal_ipc_export", "kpal_ipc_open", "kpal_ipc_close", "kpal_peer_inbox_bytes",
    "kpal_dev_reduce_push", "kpal_dev_reduce_collect", "kpal_dev_count_packed_push",
    "kpal_slice_inbox_bytes", "kpal_slice_begin", "kpal_dev_slice_push", "kpal_dev_slice_signal", "kpal_dev_slice_collect",
    "kpal_dev_slice_collect_to_host",
    "kpal_dev_count_packed", "kpal_dev_count_packed_fresh", "kpal_count_fasta_to_dev",
    "kpal_count_fasta_dev_table", "kpal_dev_finalize_counts", "kpal_dev_table_to_host", "kpal_dev_balance",
    "kpal_dev_count_by_record", "kpal_prepared_stride", "kpal_dev_profiles_prepare",
    "kpal_dev_order_by_total", "kpal_distance_num_tiles", "kpal_dev_distance_tiles",
    "kpal_distance_tile_elems", "kpal_dev_distance_tiles_packed", "kpal_dev_distance_unpack_tiles",
    "kpal_gram_row_stride", "kpal_dev_gram_prepare", "kpal_dev_gram_distances",
    "kpal_fasta_scratch_bytes", "kpal_dev_fasta_pack", "kpal_set_option",
    "kpal_kernel_launches", "kpal_reset_kernel_launches", "kpal_last_upload",
)

_lib = None


class KpalB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KpalB200Error(
            "libkpal_b200.so not found at %s; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C kpal_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    c = ctypes
    L = c.CDLL(LIB_PATH)
    u64, i32, vp = c.c_uint64, c.c_int, c.c_void_p
    pu64 = c.POINTER(c.c_uint64)

    def sig(name, restype, *argtypes):
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("kpal_abi_version", i32)
    sig("kpal_last_error", c.c_char_p)
    sig("kpal_device_count", i32)
    sig("kpal_set_device", i32, i32)
    sig("kpal_get_device", i32)
    sig("kpal_host_alloc", vp, c.c_size_t)
    sig("kpal_host_free", None, vp)
    sig("kpal_dev_alloc", vp, c.c_size_t)
    sig("kpal_dev_free", None, vp)
    sig("kpal_memcpy_h2d", i32, vp, vp, c.c_size_t, vp)
    sig("kpal_memcpy_d2h", i32, vp, vp, c.c_size_t, vp)
    sig("kpal_dev_memset", i32, vp, i32, c.c_size_t, vp)
    sig("kpal_stream_sync", i32, vp)
    sig("kpal_packed_words", None, u64, pu64, pu64)
    sig("kpal_pack_sequences", i32, vp, vp, u64, vp, vp, vp, pu64)
    sig("kpal_fasta_scan", i32, vp, u64, pu64, pu64, pu64)
    sig("kpal_fasta_pack", i32, vp, u64, vp, vp, vp, vp)
    sig("kpal_fasta_pack_segment", i32, vp, u64, u64, u64, vp, vp, u64, pu64)
    sig("kpal_fasta_slotted_bases", u64, u64, u64)
    sig("kpal_fasta_pack_slotted", i32, vp, u64, i32, u64, vp, vp, pu64)
    sig("kpal_count_sequences", i32, vp, vp, u64, i32, i32, vp)
    sig("kpal_count_fasta", i32, vp, u64, i32, i32, vp)
    sig("kpal_count_by_record", i32, vp, vp, u64, vp, u64, u64, i32, i32, vp)
    sig("kpal_balance", i32, vp, i32)
    sig("kpal_distance_matrix", i32, vp, u64, i32, i32, i32, i32, i32, i32, vp)
    sig("kpal_pair_distance", i32, vp, vp, i32, i32, i32, i32, i32, i32, vp)
    sig("kpal_matrix_open", i32, u64, i32, i32, i32, i32, i32, i32, c.POINTER(vp))
    sig("kpal_matrix_push", i32, vp, vp, u64)
    sig("kpal_matrix_finish", i32, vp, vp)
    sig("kpal_matrix_close", None, vp)
    sig("kpal_format_matrix", i32, vp, u64, u64, i32, vp, u64, pu64)
    sig("kpal_widen_u16", i32, vp, u64, u64, vp)
    sig("kpal_widen_u8", i32, vp, u64, u64, vp)
    sig("kpal_pair_distance_positive", i32, vp, vp, i32, i32, i32, i32, i32, i32, vp)
    sig("kpal_row_stats", i32, vp, u64, u64, vp)
    sig("kpal_deflate_bound", u64, u64)
    sig("kpal_deflate_chunks", i32, vp, u64, u64, i32, vp, u64, vp)
    sig("kpal_deflate_chunks_sparse", i32, vp, u64, u64, i32, vp, u64, vp)
    sig("kpal_deflate_packed_begin", i32, vp, u64, u64, i32, i32, vp, c.POINTER(vp), pu64)
    sig("kpal_deflate_packed_finish", i32, vp, vp)
    sig("kpal_rows_stats_deflate_begin", i32, vp, u64, u64, u64, i32, i32, vp, vp, c.POINTER(vp), pu64)
    sig("kpal_compact_slots", u64, vp, u64, vp, u64, vp)
    sig("kpal_split_length", u64, i32)
    sig("kpal_split", i32, vp, i32, vp, vp)
    sig("kpal_show_balance", i32, vp, i32, vp)
    sig("kpal_ipc_export", i32, vp, vp)
    sig("kpal_ipc_open", i32, vp, c.POINTER(vp))
    sig("kpal_ipc_close", i32, vp)
    sig("kpal_peer_inbox_bytes", u64, i32, i32, i32)
    sig("kpal_dev_reduce_push", i32, vp, i32, i32, i32, i32, c.POINTER(vp), vp)
    sig("kpal_dev_reduce_collect", i32, vp, i32, i32, i32, i32, vp, vp)
    sig("kpal_dev_count_packed_push", i32, vp, vp, u64, i32, vp, i32, i32, i32, c.POINTER(vp), vp,
        c.POINTER(i32))
    sig("kpal_slice_inbox_bytes", u64, i32, i32)
    sig("kpal_slice_begin", u64, i32, i32, i32)
    sig("kpal_dev_slice_push", i32, vp, i32, i32, i32, i32, c.POINTER(vp), u64, i32, vp)
    sig("kpal_dev_slice_signal", i32, i32, i32, i32, c.POINTER(vp), u64, i32, vp)
    sig("kpal_dev_slice_collect", i32, c.POINTER(vp), i32, i32, i32, u64, i32, vp, vp)
    sig("kpal_dev_slice_collect_to_host", i32, c.POINTER(vp), i32, i32, i32, u64, i32, vp, vp)
    sig("kpal_dev_count_packed", i32, vp, vp, u64, i32, vp, i32, vp)
    sig("kpal_dev_count_packed_fresh", i32, vp, vp, u64, i32, vp, i32, vp)
    sig("kpal_count_fasta_to_dev", i32, vp, u64, i32, vp, i32, vp, pu64)
    sig("kpal_count_fasta_dev_table", i32, vp, u64, i32, c.POINTER(vp), c.POINTER(i32), vp)
    sig("kpal_dev_finalize_counts", i32, vp, i32, i32, i32, vp, vp)
    sig("kpal_dev_table_to_host", i32, vp, i32, i32, i32, vp, vp)
    sig("kpal_dev_balance", i32, vp, vp, i32, vp)
    sig("kpal_dev_count_by_record", i32, vp, vp, vp, u64, u64, i32, i32, vp, vp)
    sig("kpal_prepared_stride", u64, i32)
    sig("kpal_dev_profiles_prepare", i32, vp, u64, i32, i32, i32, vp, vp, vp, vp, vp, vp)
    sig("kpal_dev_order_by_total", i32, vp, u64, i32, vp, vp)
    sig("kpal_distance_num_tiles", u64, u64)
    sig("kpal_dev_distance_tiles", i32, vp, vp, vp, vp, vp, vp, u64, i32, i32, i32, i32, i32,
        u64, u64, vp, vp)
    sig("kpal_distance_tile_elems", u64)
    sig("kpal_dev_distance_tiles_packed", i32, vp, vp, vp, vp, vp, vp, u64, i32, i32, i32, i32, i32,
        u64, u64, vp, vp)
    sig("kpal_dev_distance_unpack_tiles", i32, vp, vp, vp, vp, u64, i32, i32, i32, u64, u64, i32, vp, vp)
    sig("kpal_gram_row_stride", u64, i32)
    sig("kpal_dev_gram_prepare", i32, vp, u64, i32, i32, vp, vp, vp, vp, vp)
    sig("kpal_dev_gram_distances", i32, vp, vp, vp, u64, u64, i32, i32, i32, i32, vp, vp, vp)
    sig("kpal_fasta_scratch_bytes", u64, u64)
    sig("kpal_dev_fasta_pack", i32, vp, u64, vp, vp, vp, vp)
    sig("kpal_set_option", i32, c.c_char_p, i32)
    sig("kpal_kernel_launches", u64)
    sig("kpal_reset_kernel_launches", None)
    sig("kpal_last_upload", None, pu64, pu64)
    _lib = L
    return L


def check(code):
    """Map a C-ABI return code to a Python exception."""
    if code == KPAL_OK:
        return
    msg = (load().kpal_last_error() or b"").decode("utf-8", "replace")
    if code == KPAL_EINVAL:
        raise ValueError(msg)
    if code == KPAL_ENOMEM:
        raise MemoryError(msg)
    if code == KPAL_EOVERFLOW:
        raise OverflowError(msg)
    raise KpalB200Error(msg)


def device_count():
    return int(load().kpal_device_count())


def require_gpu():
    if device_count() < 1:
        raise KpalB200Error("no CUDA device visible: kpal_b200 runs its hot path on a "
                            "B200 GPU only (there is no CPU fallback)")


def ptr(a):
    """Host pointer of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


class PinnedArray(object):
    """NumPy view over page-locked host memory owned by the library
    (fast H2D / D2H).  Keep the object alive while the array is in use."""

    def __init__(self, shape, dtype):
        L = load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape).tolist())
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = L.kpal_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise MemoryError("kpal_host_alloc(%d) failed" % nbytes)
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            load().kpal_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------- host API

def _join_sequences(sequences):
    """bytes blob + uint64 offsets for an iterable of str / bytes."""
    parts = []
    offsets = [0]
    total = 0
    for s in sequences:
        if isinstance(s, str):
            s = s.encode("latin-1", "replace")
        else:
            s = bytes(s)
        parts.append(s)
        total += len(s)
        offsets.append(total)
    return b"".join(parts), np.asarray(offsets, dtype=np.uint64)


def count_sequences(sequences, k, balance=False):
    """int64[4**k] counts of an iterable of sequences (kpal_count_sequences)."""
    _check_k(k)
    L = load()
    require_gpu()
    blob, offsets = _join_sequences(sequences)
    out = np.empty(4 ** k, dtype=np.int64)
    check(L.kpal_count_sequences(ctypes.c_char_p(blob) if blob else None, ptr(offsets),
                                 len(offsets) - 1, int(k), int(bool(balance)), ptr(out)))
    return out


def count_fasta(text, k, balance=False, out=None):
    """int64[4**k] counts of FASTA text (str or bytes) (kpal_count_fasta).  `out`:
    optional C-contiguous int64[4**k] destination (e.g. a PinnedArray's array)."""
    _check_k(k)
    L = load()
    require_gpu()
    if isinstance(text, str):
        text = text.encode("latin-1", "replace")
    if out is None:
        out = np.empty(4 ** k, dtype=np.int64)
    elif out.dtype != np.int64 or out.shape != (4 ** k,) or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous int64 array of length 4**k")
    check(L.kpal_count_fasta(ctypes.c_char_p(text) if text else None, len(text), int(k),
                             int(bool(balance)), ptr(out)))
    return out


def fasta_pack(text):
    """Host-side scan + pack of FASTA text.  Returns (codes, valid, rec_starts,
    names, n_bases).  No GPU needed."""
    L = load()
    if isinstance(text, str):
        text = text.encode("latin-1", "replace")
    n_rec, n_bases, name_bytes = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    buf = ctypes.c_char_p(text) if text else None
    check(L.kpal_fasta_scan(buf, len(text), ctypes.byref(n_rec), ctypes.byref(n_bases),
                            ctypes.byref(name_bytes)))
    cw, vw = ctypes.c_uint64(), ctypes.c_uint64()
    L.kpal_packed_words(n_bases.value, ctypes.byref(cw), ctypes.byref(vw))
    codes = np.zeros(cw.value, dtype=np.uint32)
    valid = np.zeros(vw.value, dtype=np.uint32)
    rec_starts = np.zeros(n_rec.value + 1, dtype=np.uint64)
    names = ctypes.create_string_buffer(max(1, name_bytes.value))
    check(L.kpal_fasta_pack(buf, len(text), ptr(codes), ptr(valid), ptr(rec_starts), names))
    name_list = names.raw[:name_bytes.value].split(b"\0")[:n_rec.value] if n_rec.value else []
    return codes, valid, rec_starts, [n.decode("latin-1") for n in name_list], n_bases.value


def fasta_pack_segment(text, begin=0, end=None):
    """Host-side pack of the records of text[begin:end] into a slot of its own (the unit
    of the hybrid upload of kpal_count_fasta).  Returns (codes, valid, n_bases); the slot
    holds len(segment) rounded up to 64 bases.  No GPU needed."""
    L = load()
    if isinstance(text, str):
        text = text.encode("latin-1", "replace")
    end = len(text) if end is None else end
    cap = (end - begin + 63) // 64 * 64 + 64
    codes = np.full(cap // 16, 0xdeadbeef, dtype=np.uint32)       # the call must define every word
    valid = np.full(cap // 32, 0xdeadbeef, dtype=np.uint32)
    n_bases = ctypes.c_uint64()
    buf = ctypes.c_char_p(text) if text else None
    check(L.kpal_fasta_pack_segment(buf, len(text), begin, end, ptr(codes), ptr(valid), cap,
                                    ctypes.byref(n_bases)))
    return codes, valid, n_bases.value


def fasta_pack_slotted(text, k, seg_bytes):
    """Host-side pack of a FASTA text as a slotted stream (segments of about seg_bytes cut at
    line starts + junction records for window length k).  Returns (codes, valid, n_bases).
    No GPU needed."""
    L = load()
    if isinstance(text, str):
        text = text.encode("latin-1", "replace")
    cap = L.kpal_fasta_slotted_bases(len(text), seg_bytes)
    cw, vw = ctypes.c_uint64(), ctypes.c_uint64()
    L.kpal_packed_words(cap, ctypes.byref(cw), ctypes.byref(vw))
    codes = np.full(cw.value, 0xdeadbeef, dtype=np.uint32)
    valid = np.full(vw.value, 0xdeadbeef, dtype=np.uint32)
    n_bases = ctypes.c_uint64()
    check(L.kpal_fasta_pack_slotted(ctypes.c_char_p(text), len(text), k, seg_bytes, ptr(codes), ptr(valid),
                                    ctypes.byref(n_bases)))
    return codes, valid, n_bases.value


def pack_sequences(sequences):
    """Host-side pack of a sequence list: (codes, valid, rec_starts, n_bases)."""
    L = load()
    blob, offsets = _join_sequences(sequences)
    n_bases = ctypes.c_uint64()
    n_rec = len(offsets) - 1
    check(L.kpal_pack_sequences(None, ptr(offsets), n_rec, None, None, None, ctypes.byref(n_bases)))
    cw, vw = ctypes.c_uint64(), ctypes.c_uint64()
    L.kpal_packed_words(n_bases.value, ctypes.byref(cw), ctypes.byref(vw))
    codes = np.zeros(cw.value, dtype=np.uint32)
    valid = np.zeros(vw.value, dtype=np.uint32)
    rec_starts = np.zeros(n_rec + 1, dtype=np.uint64)
    check(L.kpal_pack_sequences(ctypes.c_char_p(blob) if blob else None, ptr(offsets), n_rec,
                                ptr(codes), ptr(valid), ptr(rec_starts), ctypes.byref(n_bases)))
    return codes, valid, rec_starts, n_bases.value


def count_by_record(codes, valid, n_bases, rec_starts, first, n, k, balance=False, out=None):
    """Dense [n][4**k] int64 rows for records [first, first+n).  `out`: a C-contiguous int64
    array with room for the rows, reused by the caller from batch to batch (a fresh 256 MB
    array per batch costs more in page faults than the GPU call it receives)."""
    _check_k(k)
    L = load()
    require_gpu()
    if out is None:
        out = np.empty((n, 4 ** k), dtype=np.int64)
    else:
        if out.dtype != np.int64 or not out.flags.c_contiguous or out.size < n * 4 ** k:
            raise ValueError("out must be a C-contiguous int64 array of at least n * 4**k elements")
        out = out.reshape(-1)[:n * 4 ** k].reshape(n, 4 ** k)
    check(L.kpal_count_by_record(ptr(codes), ptr(valid), int(n_bases), ptr(rec_starts), int(first),
                                 int(n), int(k), int(bool(balance)), ptr(out)))
    return out


def balance(counts):
    """In-place Profile.balance on a C-contiguous int64 array."""
    L = load()
    require_gpu()
    k = _k_of(counts.size)
    if counts.dtype != np.int64 or not counts.flags.c_contiguous:
        raise ValueError("counts must be a C-contiguous int64 array")
    check(L.kpal_balance(ptr(counts), k))
    return counts


def distance_matrix(profiles, metric="multiset", pairwise="prod", do_balance=False,
                    do_scale=False, down=False):
    """Symmetric [n][n] float64 matrix for a C-contiguous [n][4**k] int64 array."""
    L = load()
    require_gpu()
    profiles = np.ascontiguousarray(profiles, dtype=np.int64)
    n, size = profiles.shape
    k = _k_of(size)
    out = np.empty((n, n), dtype=np.float64)
    check(L.kpal_distance_matrix(ptr(profiles), n, k, METRICS[metric], PAIRWISE[pairwise],
                                 int(bool(do_balance)), int(bool(do_scale)), int(bool(down)),
                                 ptr(out)))
    return out


class MatrixSession(object):
    """Distance matrix over profiles handed over in slabs (kpal_matrix_open /
    push / finish): the caller never holds the whole ``[n][4**k]`` set.

    ``slab`` is a pinned ``[rows][4**k]`` int64 staging array owned by the
    session: fill ``slab[:m]`` (for instance ``dataset.read_direct(slab[i])``)
    and call ``push(m)``; or ``push_rows(array)`` for rows held elsewhere."""

    def __init__(self, n, k, metric="multiset", pairwise="prod", do_balance=False,
                 do_scale=False, down=False, slab_bytes=256 << 20):
        L = load()
        require_gpu()
        _check_k(k)
        self.n, self.k = int(n), int(k)
        self._handle = ctypes.c_void_p()
        check(L.kpal_matrix_open(self.n, self.k, METRICS[metric], PAIRWISE[pairwise],
                                 int(bool(do_balance)), int(bool(do_scale)), int(bool(down)),
                                 ctypes.byref(self._handle)))
        rows = max(1, min(self.n, slab_bytes // (8 * 4 ** self.k)))
        self._pinned = PinnedArray((rows, 4 ** self.k), np.int64)
        self.slab = self._pinned.array

    def push(self, m):
        """Upload the first `m` rows of ``self.slab``."""
        check(load().kpal_matrix_push(self._handle, self._pinned._ptr, int(m)))

    def push_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        if rows.ndim != 2 or rows.shape[1] != 4 ** self.k:
            raise ValueError("rows must be [m][4**k]")
        check(load().kpal_matrix_push(self._handle, ptr(rows), rows.shape[0]))

    def finish(self):
        """Run the distance kernels; symmetric ``[n][n]`` float64 result."""
        out = np.empty((self.n, self.n), dtype=np.float64)
        check(load().kpal_matrix_finish(self._handle, ptr(out)))
        return out

    def close(self):
        if self._handle:
            load().kpal_matrix_close(self._handle)
            self._handle = ctypes.c_void_p()
        if self._pinned is not None:
            self.slab = None
            self._pinned.free()
            self._pinned = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def row_stats(rows):
    """``[n][5]`` float64: total, non_zero, mean, median, std of every row of a C-contiguous
    ``[n][m]`` int64 array, bit-identical to the NumPy calls (host, multi-threaded)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    if rows.ndim != 2 or rows.shape[1] == 0:
        raise ValueError("rows must be [n][m] with m > 0")
    out = np.empty((rows.shape[0], 5), dtype=np.float64)
    check(load().kpal_row_stats(ptr(rows), rows.shape[0], rows.shape[1], ptr(out)))
    return out


def deflate_chunks(data, chunk_bytes, level, sparse=False):
    """zlib streams of the equal-sized chunks of `data` (a C-contiguous array whose size is a
    multiple of `chunk_bytes`): ``(buffer, slot_bytes, sizes)`` -- chunk ``c`` is
    ``buffer[c * slot_bytes : c * slot_bytes + sizes[c]]``.  Host, multi-threaded.
    `sparse`: the one-pass encoder for mostly-zero data (valid zlib streams, not the bytes
    zlib itself would write)."""
    L = load()
    raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    if chunk_bytes < 1 or raw.size % chunk_bytes:
        raise ValueError("data must be whole chunks")
    n_chunks = raw.size // chunk_bytes
    slot = int(L.kpal_deflate_bound(chunk_bytes))
    buffer = np.empty(n_chunks * slot, dtype=np.uint8)
    sizes = np.empty(n_chunks, dtype=np.uint32)
    call = L.kpal_deflate_chunks_sparse if sparse else L.kpal_deflate_chunks
    check(call(ptr(raw), n_chunks, int(chunk_bytes), int(level), ptr(buffer), slot, ptr(sizes)))
    return buffer, slot, sizes


def deflate_chunks_packed(data, chunk_bytes, level, sparse=False):
    """The same streams back to back: ``(blob, sizes)`` (chunk ``c`` starts at
    ``sizes[:c].sum()``)."""
    L = load()
    raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    if chunk_bytes < 1 or raw.size % chunk_bytes:
        raise ValueError("data must be whole chunks")
    n_chunks = raw.size // chunk_bytes
    sizes = np.empty(n_chunks, dtype=np.uint32)
    handle, total = ctypes.c_void_p(), ctypes.c_uint64()
    check(L.kpal_deflate_packed_begin(ptr(raw), n_chunks, int(chunk_bytes), int(level), 1 if sparse else 0,
                                      ptr(sizes), ctypes.byref(handle), ctypes.byref(total)))
    blob = np.empty(total.value, dtype=np.uint8)
    check(L.kpal_deflate_packed_finish(handle, ptr(blob)))
    return blob, sizes


def rows_stats_deflate(rows, chunk_bytes, level, sparse=True):
    """``row_stats(rows)`` and ``deflate_chunks_packed(rows, chunk_bytes, level, sparse)`` in one
    pass over the C-contiguous int64 `rows`: ``(stats, blob, sizes)``."""
    L = load()
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n, m = rows.shape
    if chunk_bytes < 1 or (m * 8) % chunk_bytes:
        raise ValueError("chunk_bytes must divide a row")
    stats = np.empty((n, 5), dtype=np.float64)
    sizes = np.empty(n * (m * 8 // chunk_bytes), dtype=np.uint32)
    handle, total = ctypes.c_void_p(), ctypes.c_uint64()
    check(L.kpal_rows_stats_deflate_begin(ptr(rows), n, m, int(chunk_bytes), int(level), 1 if sparse else 0,
                                          ptr(stats), ptr(sizes), ctypes.byref(handle), ctypes.byref(total)))
    blob = np.empty(total.value, dtype=np.uint8)
    check(L.kpal_deflate_packed_finish(handle, ptr(blob)))
    return stats, blob, sizes


def format_matrix(values, precision):
    """Lower triangle of a square float64 matrix as the text rows of
    kdistlib.distance_matrix (kpal_format_matrix; host C++, no GPU needed)."""
    L = load()
    values = np.asarray(values, dtype=np.float64)
    if values.ndim != 2 or values.shape[0] != values.shape[1]:
        raise ValueError("values must be a square matrix")
    if values.strides[1] != 8 or values.strides[0] % 8 or values.strides[0] < 8 * values.shape[1]:
        values = np.ascontiguousarray(values)
    n, ld = values.shape[0], values.strides[0] // 8 if values.shape[0] > 1 else values.shape[1]
    cells = n * (n - 1) // 2
    capacity = cells * (int(precision) + 12) + n + 16
    length = ctypes.c_uint64()
    for _ in range(2):
        buf = ctypes.create_string_buffer(capacity)
        code = L.kpal_format_matrix(ptr(values), n, ld, int(precision), buf, capacity,
                                    ctypes.byref(length))
        if code != KPAL_EOVERFLOW:
            break
        capacity = length.value + 1
    check(code)
    return buf.raw[:length.value].decode("ascii")


def split(counts):
    """Profile.split on the device: (forward, reverse) int64 arrays."""
    L = load()
    require_gpu()
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    k = _k_of(counts.size)
    half = int(L.kpal_split_length(k))
    forward = np.empty(half, dtype=np.int64)
    reverse = np.empty(half, dtype=np.int64)
    check(L.kpal_split(ptr(counts), k, ptr(forward), ptr(reverse)))
    return forward, reverse


def show_balance(counts):
    """multiset/prod distance between the two halves of Profile.split."""
    L = load()
    require_gpu()
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    out = ctypes.c_double()
    check(L.kpal_show_balance(ptr(counts), _k_of(counts.size), ctypes.byref(out)))
    return out.value


def pair_distance(left, right, metric="multiset", pairwise="prod", do_balance=False,
                  do_scale=False, down=False, do_positive=False):
    L = load()
    require_gpu()
    left = np.ascontiguousarray(left, dtype=np.int64)
    right = np.ascontiguousarray(right, dtype=np.int64)
    if left.shape != right.shape or left.ndim != 1:
        raise ValueError("profiles must be 1-D and of equal length")
    k = _k_of(left.size)
    out = ctypes.c_double()
    entry = L.kpal_pair_distance_positive if do_positive else L.kpal_pair_distance
    check(entry(ptr(left), ptr(right), k, METRICS[metric], PAIRWISE[pairwise],
                int(bool(do_balance)), int(bool(do_scale)), int(bool(down)), ctypes.byref(out)))
    return out.value


def _check_k(k):
    if not (1 <= int(k) <= MAX_K):
        raise ValueError("k-mer length %r out of range [1, %d]" % (k, MAX_K))


def _k_of(size):
    k = int(size).bit_length() // 2
    if size < 4 or 4 ** k != size:
        raise ValueError("profile length %d is not a power of 4" % size)
    _check_k(k)
    return k
