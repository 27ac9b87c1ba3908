"""
*k*-mer profiles -- drop-in mirror of the reference's ``kpal.klib.Profile``
(same constructor, class methods, properties and error behaviour; reference
kpal/klib.py:25-487) with the hot path on the GPU:

===============================  ==============================================
reference (file:line)            here
===============================  ==============================================
from_sequences  klib.py:135-170  ``kpal_count_sequences`` (CUDA, bit exact)
from_fasta      klib.py:97-112   ``kpal_count_fasta`` (C++ FASTA scan + CUDA)
from_fasta_by_record  114-133    ``kpal_fasta_pack`` + ``kpal_count_by_record``
balance         klib.py:285-298  ``kpal_balance`` (CUDA, bit exact)
everything else                  host NumPy, same semantics (not on the hot path)
===============================  ==============================================

There is no CPU fallback for the four rows above: without the CUDA library or a
GPU they raise.
"""
import itertools
import math

import numpy as np

from . import _cabi, metrics


class Profile(object):
    """
    A *k*-mer profile: ``counts`` is an ``int64`` array of length ``4**k``
    ordered alphabetically by *k*-mer (A < C < G < T, first base most
    significant), ``name`` an optional profile name.

    Use the construction methods :meth:`from_fasta`, :meth:`from_sequences`,
    :meth:`from_fasta_by_record`, :meth:`from_file` rather than the
    constructor.
    """
    #: nucleotide -> 2-bit code (kpal/klib.py:43-48)
    _nucleotide_to_binary = {
        'A': 0, 'a': 0,
        'C': 1, 'c': 1,
        'G': 2, 'g': 2,
        'T': 3, 't': 3,
    }
    #: 2-bit code -> nucleotide (kpal/klib.py:51-56)
    _binary_to_nucleotide = {0: 'A', 1: 'C', 2: 'G', 3: 'T'}

    #: rows per device batch in :meth:`from_fasta_by_record` are capped so a
    #: batch of dense rows stays below this many bytes
    _BY_RECORD_BATCH_BYTES = 256 << 20

    def __init__(self, counts, name=None):
        # Same float-log derivation as the reference (kpal/klib.py:59).
        self.length = int(math.log(len(counts), 4))
        self.counts = counts
        self.name = name

    # ------------------------------------------------------------------ I/O
    @classmethod
    def from_file(cls, handle, name=None):
        """Load a profile from an open HDF5 *k*-mer profile file
        (kpal/klib.py:63-76); `name` defaults to the first profile."""
        name = name or sorted(handle['profiles'].keys())[0]
        return cls(handle['profiles/' + name][:], name=name)

    @classmethod
    def from_file_old_format(cls, handle, name=None):
        """Load a profile from the old plain-text format: three header lines
        (k, total, non-zero) then one count per line (kpal/klib.py:78-95)."""
        for _ in range(3):
            next(handle)
        return cls(np.loadtxt(handle, dtype='int64'), name=name)

    def save(self, handle, name=None):
        """
        Write the profile to an open, writable HDF5 profile file as dataset
        ``/profiles/<name>`` (int64, gzip) with the ``length, total, non_zero,
        mean, median, std`` attributes (kpal/klib.py:227-256,
        doc/fileformat.rst:23-39).  Returns the name used: `name`, else the
        profile's own name, else the first free number from 1.
        """
        if name and ('/' in name or '.' in name):
            raise ValueError('Profile name may not contain / or . characters.')
        if not name:
            name = self.name
        if not name:
            taken = handle['profiles']
            name = next(str(n) for n in itertools.count(1) if str(n) not in taken)

        dataset = handle.create_dataset('profiles/' + name, data=self.counts,
                                        dtype='int64', compression='gzip')
        for attribute in ('length', 'total', 'non_zero', 'mean', 'median', 'std'):
            dataset.attrs[attribute] = getattr(self, attribute)
        handle.flush()
        return name

    # ----------------------------------------------------------- construction
    @classmethod
    def from_fasta(cls, handle, length, name=None):
        """
        Count all *k*-mers of every record of an open FASTA file into one
        profile (kpal/klib.py:97-112).  The text is scanned and 2-bit packed
        by the C++ side of the library and counted on the GPU.
        """
        return cls(_cabi.count_fasta(_read_text(handle), length), name=name)

    @classmethod
    def from_fasta_by_record(cls, handle, length, prefix=None):
        """
        Generator of one profile per FASTA record, named ``<prefix>_`` +
        record name, or the 1-based record number for a nameless record
        (kpal/klib.py:114-133).  Records are counted on the GPU in batches of
        dense rows; the batching is invisible to the caller.
        """
        _cabi._check_k(length)
        prefix = prefix + '_' if prefix else ''
        codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(_read_text(handle))
        n_records = len(names)
        batch = max(1, cls._BY_RECORD_BATCH_BYTES // (8 * 4 ** length))
        for first in range(0, n_records, batch):
            n = min(batch, n_records - first)
            rows = _cabi.count_by_record(codes, valid, n_bases, rec_starts, first, n, length)
            for i in range(n):
                record = first + i
                yield cls(rows[i].copy(), name=prefix + (names[record] or str(record + 1)))

    @classmethod
    def record_batches(cls, handle, length, prefix=None):
        """
        The same profiles as :meth:`from_fasta_by_record`, a device batch at a time: a
        generator of ``(names, rows)`` with `rows` the dense ``[n][4**length]`` int64 counts of
        `n` consecutive records.  What ``kpal count --by-record`` feeds to
        :func:`save_profiles` (no Profile object, no row copy per record).  `rows` lives in one
        of two reused buffers: it is valid until the next batch is asked for (copy it to keep it).
        """
        _cabi._check_k(length)
        prefix = prefix + '_' if prefix else ''
        codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(_read_text(handle))
        n_records = len(names)
        batch = max(1, cls._BY_RECORD_BATCH_BYTES // (8 * 4 ** length))
        starts = list(range(0, n_records, batch))
        if not starts:
            return
        # Two row buffers, reused: the device counts batch b + 1 (on a helper thread; the C call
        # releases the GIL) while the caller works on batch b.  `rows` is only valid until the
        # next batch is asked for.
        from concurrent.futures import ThreadPoolExecutor
        buffers = [np.empty((min(batch, n_records), 4 ** length), dtype=np.int64) for _ in range(min(2, len(starts)))]

        device = _cabi.load().kpal_get_device()         # the helper thread works on the caller's device

        def count(slot, first):
            n = min(batch, n_records - first)
            if device >= 0:
                _cabi.check(_cabi.load().kpal_set_device(device))
            return _cabi.count_by_record(codes, valid, n_bases, rec_starts, first, n, length, out=buffers[slot])

        # ONE helper thread for all batches: a thread's first CUDA call costs milliseconds
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(count, 0, starts[0])
            for b, first in enumerate(starts):
                rows = pending.result()
                if b + 1 < len(starts):
                    pending = pool.submit(count, (b + 1) % 2, starts[b + 1])
                n = rows.shape[0]
                yield [prefix + (names[first + i] or str(first + i + 1)) for i in range(n)], rows

    @classmethod
    def from_sequences(cls, sequences, length, name=None):
        """
        Count all *k*-mers in an iterable of sequence strings
        (kpal/klib.py:135-170).  Every character outside ``ACGTacgt`` splits
        its sequence; windows never span two sequences.  Runs on the GPU.
        """
        return cls(_cabi.count_sequences(sequences, length), name=name)

    # ------------------------------------------------------------- properties
    @property
    def name(self):
        """Profile name (may not contain ``/`` or ``.``)."""
        return self._name

    @name.setter
    def name(self, name):
        if name and ('/' in name or '.' in name):
            raise ValueError('Profile name may not contain / or . characters.')
        self._name = name

    @property
    def number(self):
        """Number of possible *k*-mers of this length."""
        return len(self.counts)

    @property
    def non_zero(self):
        """Number of *k*-mers with a non-zero count."""
        return np.count_nonzero(self.counts)

    @property
    def total(self):
        """Sum of the counts."""
        return self.counts.sum()

    @property
    def mean(self):
        """Mean of the counts."""
        return self.counts.mean()

    @property
    def median(self):
        """Median of the counts."""
        return np.median(self.counts)

    @property
    def std(self):
        """Standard deviation of the counts."""
        return self.counts.std()

    # -------------------------------------------------------------- operations
    def copy(self):
        """Deep copy (kpal/klib.py:258-267)."""
        return type(self)(self.counts.copy(), name=self.name)

    def merge(self, profile, merger=metrics.mergers["sum"]):
        """Merge `profile` into this one with a vectorised pairwise `merger`
        (kpal/klib.py:269-283)."""
        self.counts = merger(self.counts, profile.counts)

    def balance(self):
        """
        Add to every *k*-mer the count of its reverse complement and vice
        versa; palindromes are doubled (kpal/klib.py:285-298).  In place, on
        the GPU (``kpal_balance``).
        """
        counts = self.counts
        if not isinstance(counts, np.ndarray) or counts.dtype.kind not in 'iu':
            raise TypeError('Profile.balance on the device needs integer counts')
        if counts.dtype != np.int64 or not counts.flags.c_contiguous:
            work = np.ascontiguousarray(counts, dtype=np.int64)
            _cabi.balance(work)
            counts[...] = work
        else:
            _cabi.balance(counts)

    def _rc_table(self):
        """``rc(i)`` for every index (vectorised :meth:`reverse_complement`)."""
        index = np.arange(self.number, dtype=np.int64)
        rest = ~index
        table = np.zeros_like(index)
        for _ in range(self.length):
            table = (table << 2) | (rest & 3)
            rest >>= 2
        return table

    def split(self):
        """
        Forward / reverse-complement halves of the profile, every position of
        the first array facing its reverse complement in the second; counts
        are doubled except for palindromes, which appear once in both
        (kpal/klib.py:300-327).  On the GPU (``kpal_split``).
        """
        return _cabi.split(self.counts)

    def shrink(self, factor=1):
        """Reduce *k* by `factor`, summing groups of ``4**factor`` neighbours
        (kpal/klib.py:329-352).  Host path."""
        if self.length <= factor:
            raise ValueError(
                "Reduction factor should be smaller than k-mer size.")
        group = 4 ** factor
        self.counts = np.asarray(self.counts, dtype='int64').reshape(-1, group).sum(axis=1)
        self.length -= factor

    def shuffle(self):
        """Randomise the profile in place with NumPy's global RNG
        (kpal/klib.py:354-358)."""
        np.random.shuffle(self.counts)

    def dna_to_binary(self, sequence):
        """Index of a DNA string (kpal/klib.py:360-375)."""
        result = 0
        for base in sequence:
            result = (result << 2) | self._nucleotide_to_binary[base]
        return result

    def binary_to_dna(self, number):
        """DNA string of an index (kpal/klib.py:377-392)."""
        bases = []
        for _ in range(self.length):
            bases.append(self._binary_to_nucleotide[number & 3])
            number >>= 2
        return ''.join(reversed(bases))

    def reverse_complement(self, number):
        """Index of the reverse complement of the *k*-mer with index `number`
        (kpal/klib.py:394-412): complement = bitwise NOT, then the 2-bit
        groups are reversed."""
        number = ~number
        result = 0
        for _ in range(self.length):
            result = (result << 2) | (number & 3)
            number >>= 2
        return result

    def print_counts(self):
        """Print ``<k-mer> <count>`` lines (kpal/klib.py:460-465)."""
        for i in range(self.number):
            print(self.binary_to_dna(i), self.counts[i])


def save_profiles(handle, names, rows):
    """
    ``Profile(rows[i], names[i]).save(handle)`` for every row (reference kpal/klib.py:227-256:
    dataset ``/profiles/<name>``, int64, gzip, attributes ``length, total, non_zero, mean,
    median, std``) -- for the 100 000 profiles of a ``--by-record`` run.  With the in-tree
    HDF5 writer the statistics of all rows come from ``kpal_row_stats`` (host threads,
    bit-identical to the NumPy calls behind the reference's properties), the chunks from
    ``kpal_deflate_chunks``, and the datasets are created with one call; with h5py it is the
    plain loop.
    """
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    for name in names:
        if not name or '/' in name or '.' in name:
            raise ValueError('Profile name may not contain / or . characters.')
    group = handle['profiles']
    if not hasattr(group, 'create_datasets') or len(names) < 2:
        for name, row in zip(names, rows):
            Profile(row, name=name).save(handle)
        return
    layout = group.bulk_layout(rows) if hasattr(group, 'bulk_layout') else None
    streams = None
    if layout:      # statistics and the chunks' zlib streams in one pass over the rows
        stats, blob, sizes = _cabi.rows_stats_deflate(rows, layout[0] * 8, layout[1])
        streams = (blob, sizes)
    else:
        stats = _cabi.row_stats(rows)
    length = int(math.log(rows.shape[1], 4))
    group.create_datasets(list(names), rows, streams=streams, attrs={
        'length': np.full(len(names), length, dtype=np.int64),
        'total': stats[:, 0].astype(np.int64), 'non_zero': stats[:, 1].astype(np.int64),
        'mean': stats[:, 2].copy(), 'median': stats[:, 3].copy(), 'std': stats[:, 4].copy()})
    handle.flush()


def _read_text(handle):
    """
    Whole content of an open FASTA handle (text or binary), as ``str`` or ``bytes``.

    A text-mode file positioned at its start is read through its binary buffer when the
    bytes are plain ASCII without ``\\r``: decoding 100 MB to ``str`` and encoding it again
    for the C ABI costs ~50 times the GPU call it feeds.  Anything else (universal-newline
    translation, another encoding, a partly consumed or unseekable handle, ``StringIO``)
    goes through ``handle.read()`` exactly as before.  The shortcut is only taken when the
    handle's encoding maps bytes below 0x80 to the same ASCII characters (UTF-16 / UTF-32
    text without a BOM is all "ASCII bytes" too, but NUL-interleaved).
    """
    buffer = getattr(handle, 'buffer', None)
    if buffer is not None and hasattr(handle, 'encoding'):
        try:
            at_start = handle.seekable() and handle.tell() == 0
        except (OSError, ValueError):
            at_start = False
        if at_start and _ascii_superset(getattr(handle, 'encoding', None)):
            data = buffer.read()
            if data.isascii() and b'\r' not in data:
                return data
            handle.seek(0)                     # let the text layer do what it does
    return handle.read()


#: codecs under which a byte below 0x80 decodes to the ASCII character of the same value
_ASCII_SUPERSETS = frozenset((
    'ascii', 'utf-8', 'utf-8-sig', 'latin-1', 'iso8859-1', 'iso8859-15', 'cp1252', 'cp1250',
    'cp1251', 'cp437', 'cp850', 'mac-roman', 'iso8859-2', 'koi8-r', 'gbk', 'gb2312', 'gb18030',
    'euc-jp', 'euc-kr', 'big5', 'shift_jis'))


def _ascii_superset(encoding):
    import codecs
    try:
        return codecs.lookup(encoding or '').name in _ASCII_SUPERSETS
    except (LookupError, TypeError):
        return False
