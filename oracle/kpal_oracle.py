"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of kPAL's hot path.

This module is the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product package ``kpal_b200`` never
does (it fails loudly when its CUDA library is missing instead).

Parity status: PINNED.  Every function below is checked (tests/test_oracle.py)
against (a) the golden vectors the reference's own tests hold
(reference tests/utils.py:25-59, tests/test_kdistlib.py:38-134,
tests/test_klib.py:32-99,164-180), committed under tests/golden/, and (b) in
the build container, the unmodified reference sources executed through
``oracle/ref_loader.py`` on randomised inputs.

Unpinned corner: Biopython's FASTA reader is an un-vendored dependency of the
reference (``biopython``, no version pin, reference setup.py:8); the reader
below restates the documented FastaIterator behaviour (see parse_fasta).

All ``file:line`` citations are relative to the reference tree.
"""
import math

import numpy as np

# --------------------------------------------------------------------------
# FASTA text -> (name, sequence) records
# --------------------------------------------------------------------------


def parse_fasta(text):
    """
    Restatement of ``Bio.SeqIO.parse(handle, 'fasta')`` as used at
    kpal/klib.py:111 and kpal/klib.py:131-132 (third-party: biopython,
    unpinned).  Published FastaIterator behaviour: lines before the first
    '>' are skipped; title = header line without '>' right-stripped; name = id
    = first whitespace-delimited token of the title ('' when empty); sequence
    = all following lines up to the next '>' line, each right-stripped,
    joined, with spaces and '\\r' removed.

    Pinned by the reference: one-line records (tests/utils.py:184-196),
    first-token names (tests/test_klib.py:38-45), multi-line joining
    (doc/tutorial.rst:44-82 fixture).  Everything else: parity unpinned.
    """
    if isinstance(text, bytes):
        text = text.decode("latin-1")
    records = []
    name = None
    chunks = []
    for line in text.split("\n"):
        if line.startswith(">"):
            if name is not None:
                records.append((name, "".join(chunks)))
            title = line[1:].rstrip()
            parts = title.split(None, 1)
            name = parts[0] if parts else ""
            chunks = []
        elif name is not None:
            chunks.append(line.rstrip().replace(" ", "").replace("\r", ""))
    if name is not None:
        records.append((name, "".join(chunks)))
    return records


# --------------------------------------------------------------------------
# counting
# --------------------------------------------------------------------------

#: kpal/klib.py:43-48 -- A/a 0, C/c 1, G/g 2, T/t 3; anything else splits
#: (regex '[^AaCcGgTt]', kpal/klib.py:152).
_LUT = np.full(256, 255, dtype=np.uint8)
for _ch, _code in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _LUT[ord(_ch)] = _code
    _LUT[ord(_ch.lower())] = _code


def count_python(sequences, k):
    """
    Literal scalar restatement of ``Profile.from_sequences``
    (kpal/klib.py:149-170): per sequence, per maximal ACGTacgt run of length
    >= k, roll ``binary = ((binary << 2) | code) & mask`` and bump
    ``counts[binary]``.  Pure Python -- small cases only.
    """
    number = 4 ** k
    mask = number - 1
    counts = [0] * number
    for sequence in sequences:
        run = 0
        binary = 0
        for ch in sequence:
            code = int(_LUT[ord(ch) & 0xFF]) if ord(ch) < 256 else 255
            if code == 255:
                run = 0
                binary = 0
                continue
            binary = ((binary << 2) | code) & mask
            run += 1
            if run >= k:
                counts[binary] += 1
    return np.array(counts, dtype=np.int64)


def _count_bytes(buf, k, counts, chunk=1 << 24):
    """Vectorised window count over a uint8 buffer in which every byte that
    is not ACGTacgt (including the record separators inserted by the callers)
    splits, exactly like the regex split of kpal/klib.py:152-156."""
    n = buf.size
    if n < k:
        return
    number = 4 ** k
    start = 0
    while start < n - k + 1:
        stop = min(n, start + chunk + k - 1)       # windows start in [start, stop-k]
        codes = _LUT[buf[start:stop]]
        bad = (codes == 255)
        m = codes.size
        nwin = m - k + 1
        # window i is valid iff no bad byte in [i, i+k)
        cbad = np.concatenate(([0], np.cumsum(bad, dtype=np.int64)))
        ok = (cbad[k:k + nwin] - cbad[:nwin]) == 0
        c64 = np.where(bad, 0, codes).astype(np.int64)
        idx = np.zeros(nwin, dtype=np.int64)
        for j in range(k):                          # first base most significant
            idx = (idx << 2) | c64[j:j + nwin]      # kpal/klib.py:160-167
        counts += np.bincount(idx[ok], minlength=number)
        start += nwin


def count_sequences(sequences, k):
    """
    Vectorised restatement of ``Profile.from_sequences`` (kpal/klib.py:135-170).
    Returns ``int64[4**k]``.  Windows never span two sequences
    (kpal/klib.py:154: each sequence is split separately).
    """
    counts = np.zeros(4 ** k, dtype=np.int64)
    parts = []
    for s in sequences:
        if isinstance(s, str):
            s = s.encode("latin-1", "replace")
        parts.append(np.frombuffer(s, dtype=np.uint8))
        parts.append(np.zeros(1, dtype=np.uint8))    # separator = invalid byte
    if parts:
        _count_bytes(np.concatenate(parts), k, counts)
    return counts


def count_fasta(text, k):
    """``Profile.from_fasta`` (kpal/klib.py:97-112)."""
    return count_sequences((seq for _, seq in parse_fasta(text)), k)


def count_fasta_by_record(text, k, prefix=None):
    """``Profile.from_fasta_by_record`` (kpal/klib.py:114-133): list of
    (name, int64[4**k]); name = prefix_ + (record.name or 1-based index)."""
    prefix = prefix + "_" if prefix else ""
    out = []
    for i, (name, seq) in enumerate(parse_fasta(text)):
        out.append((prefix + (name or str(i + 1)), count_sequences([seq], k)))
    return out


# --------------------------------------------------------------------------
# balance
# --------------------------------------------------------------------------


def reverse_complement(number, k):
    """kpal/klib.py:394-412: complement = bitwise NOT, then reverse the k
    2-bit groups."""
    number = ~number
    result = 0
    for _ in range(k):
        result = (result << 2) | (number & 3)
        number >>= 2
    return result


def reverse_complement_table(k):
    """rc(i) for all i in [0, 4**k) (vectorised reverse_complement)."""
    idx = np.arange(4 ** k, dtype=np.int64)
    comp = ~idx
    out = np.zeros_like(idx)
    for _ in range(k):
        out = (out << 2) | (comp & 3)
        comp >>= 2
    return out


def balance(counts):
    """
    ``Profile.balance`` (kpal/klib.py:285-298): every pair (i, rc(i)) gets
    counts[i] + counts[rc(i)]; palindromes (i == rc(i)) are doubled.  Returns
    a new array.
    """
    counts = np.asarray(counts)
    k = int(round(math.log(counts.size, 4)))
    return counts + counts[reverse_complement_table(k)]


def balance_python(counts):
    """Literal loop form of kpal/klib.py:290-298 (small k only)."""
    counts = np.array(counts, dtype=np.int64)
    k = int(round(math.log(counts.size, 4)))
    for i in range(counts.size):
        i_rc = reverse_complement(i, k)
        if i < i_rc:
            temp = counts[i]
            counts[i] += counts[i_rc]
            counts[i_rc] += temp
        elif i == i_rc:
            counts[i] += counts[i]
    return counts


def split(counts):
    """
    ``Profile.split`` (kpal/klib.py:300-327): forward / reverse lists over the
    indices i <= rc(i) in index order; counts doubled for i < rc(i), taken once
    for palindromes.
    """
    counts = np.asarray(counts)
    k = int(round(math.log(counts.size, 4)))
    index = np.arange(counts.size, dtype=np.int64)
    partner = reverse_complement_table(k)
    keep = index <= partner
    factor = np.where(index[keep] < partner[keep], 2, 1)
    return counts[index[keep]] * factor, counts[partner[keep]] * factor


def split_python(counts):
    """Literal loop form of kpal/klib.py:314-327 (small k only)."""
    counts = np.asarray(counts)
    k = int(round(math.log(counts.size, 4)))
    forward, reverse = [], []
    for i in range(counts.size):
        i_rc = reverse_complement(i, k)
        if i < i_rc:
            forward.append(counts[i] * 2)
            reverse.append(counts[i_rc] * 2)
        elif i == i_rc:
            forward.append(counts[i])
            reverse.append(counts[i])
    return np.array(forward), np.array(reverse)


def show_balance(counts):
    """The figure ``kpal showbalance`` prints (kmer.get_balance,
    kpal/kmer.py:240-245): multiset/prod distance of the two split lists."""
    forward, reverse = split(counts)
    return multiset(forward, reverse, pairwise_prod)


def positive(vector, mask):
    """``metrics.positive`` (kpal/metrics.py:89-98)."""
    return np.multiply(vector, np.asanyarray(mask, dtype=bool))


# --------------------------------------------------------------------------
# metrics + distance
# --------------------------------------------------------------------------


def get_scale(left, right):
    """kpal/metrics.py:49-72."""
    left_scale = 1.0
    right_scale = 1.0
    left_sum = np.sum(left)
    right_sum = np.sum(right)
    with np.errstate(divide="ignore", invalid="ignore"):
        if left_sum < right_sum:
            left_scale = right_sum / left_sum
        else:
            right_scale = left_sum / right_sum
    return left_scale, right_scale


def scale_down(left, right):
    """kpal/metrics.py:75-86."""
    factor = max(left, right)
    return left / factor, right / factor


def pairwise_prod(x, y):
    """kpal/metrics.py:160."""
    return abs(x - y) / ((x + 1) * (y + 1))


def pairwise_sum(x, y):
    """kpal/metrics.py:161."""
    return abs(x - y) / (x + y + 1)


PAIRWISE = {"prod": pairwise_prod, "sum": pairwise_sum}


def multiset(left, right, pairwise=pairwise_prod):
    """kpal/metrics.py:101-123."""
    left = np.asanyarray(left)
    right = np.asanyarray(right)
    nonzero = np.where(np.logical_or(left, right))
    distances = pairwise(left[nonzero], right[nonzero])
    return distances.sum() / (len(distances) + 1)


def euclidean(left, right):
    """kpal/metrics.py:126-135 with vector_length 36-46."""
    d = np.subtract(left, right)
    return np.sqrt(np.dot(d, d))


def cosine_similarity(left, right):
    """kpal/metrics.py:138-147."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.dot(left, right) / (np.sqrt(np.dot(left, left)) *
                                      np.sqrt(np.dot(right, right)))


def distance(left, right, do_balance=False, do_scale=False, down=False,
             metric="multiset", pairwise="prod", do_positive=False):
    """
    ``ProfileDistance.distance`` (kpal/kdistlib.py:126-161) without smoothing:
    copy -> balance -> positive -> scale -> metric.  ``metric`` in
    {'multiset','euclidean','cosine'}.
    """
    left = np.array(left, dtype=np.int64)
    right = np.array(right, dtype=np.int64)
    if do_balance:
        left = balance(left)
        right = balance(right)
    if do_positive:                                   # kpal/kdistlib.py:143-145
        left, right = positive(left, right), positive(right, left)
    if do_scale:
        ls, rs = get_scale(left, right)
        if down:
            ls, rs = scale_down(ls, rs)
        with np.errstate(invalid="ignore"):
            left = left * ls
            right = right * rs
    with np.errstate(invalid="ignore", divide="ignore"):
        if metric == "multiset":
            return multiset(left, right, PAIRWISE[pairwise])
        if metric == "euclidean":
            return euclidean(left, right)
        if metric == "cosine":
            return cosine_similarity(left, right)
    raise ValueError(metric)


def distance_matrix_values(profiles, **opts):
    """N x N float64 with d(p_i, p_j) in the strict lower triangle (row i,
    col j < i) as produced by kpal/kdistlib.py:179-184; upper triangle and
    diagonal are zero."""
    n = len(profiles)
    out = np.zeros((n, n), dtype=np.float64)
    for i in range(1, n):
        for j in range(i):
            out[i, j] = distance(profiles[i], profiles[j], **opts)
    return out


def format_matrix(names, values, precision):
    """Text layout of kpal/kdistlib.py:176-186."""
    n = len(names)
    lines = [str(n)] + [str(x) for x in names]
    fmt = "{{0:.{0}f}}".format(precision)
    for i in range(1, n):
        lines.append(" ".join(fmt.format(values[i][j]) for j in range(i)))
    return "\n".join(lines) + "\n"
