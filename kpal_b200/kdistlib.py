"""
*k*-mer profile distances -- drop-in mirror of the reference's
``kpal.kdistlib`` (``ProfileDistance``, ``distance_matrix``; reference
kpal/kdistlib.py:21-186).

GPU fast path (``kpal_distance_matrix`` / ``kpal_pair_distance``): every
combination of ``do_balance``, ``do_scale``, ``down`` with the multiset
distance (built-in ``prod`` / ``sum`` pairwise) or the euclidean / cosine
vector functions.  ``do_positive`` (pair-dependent mask and totals) runs on
the device pair by pair (``kpal_pair_distance_positive``).  For those options
nothing is computed on the host, and a missing library or GPU raises.

Host path (unchanged NumPy pipeline, as the north star prescribes):
``do_smooth`` (recursive collapse with data-dependent control flow) and
user-supplied ``pairwise`` / ``distance_function`` / ``summary`` callables.
"""
import numpy as np

from . import _cabi, metrics

LENGTH_ERROR = 'k-mer lengths of the files differ'


class ProfileDistance(object):
    """
    Distance function object (kpal/kdistlib.py:21-51).

    :arg bool do_balance: balance the profiles first.
    :arg bool do_positive: only use positions that are non-zero in both.
    :arg bool do_smooth: dynamic smoothing.
    :arg function summary: summary function for dynamic smoothing.
    :arg int threshold: threshold for the summary function.
    :arg bool do_scale: scale the profiles to equal totals.
    :arg bool down: normalise the scaling factors between 0 and 1.
    :arg function distance_function: a vector distance instead of multiset.
    :arg function pairwise: pairwise function for the multiset distance.
    """
    def __init__(self, do_balance=False, do_positive=False, do_smooth=False,
                 summary=metrics.summary['min'], threshold=0, do_scale=False,
                 down=False, distance_function=None,
                 pairwise=metrics.pairwise['prod']):
        self._do_balance = do_balance
        self._do_positive = do_positive
        self._do_smooth = do_smooth
        self._threshold = threshold
        self._do_scale = do_scale
        self._down = down
        self._distance_function = distance_function
        self._pairwise = pairwise
        self._function = summary

    # ------------------------------------------------------------ GPU dispatch
    def _gpu_options(self):
        """Keyword arguments for the C ABI when this object's options are in
        the device fast path, else ``None``."""
        if self._do_smooth:
            return None
        if self._distance_function is None:
            for key in ('prod', 'sum'):
                if self._pairwise is metrics.pairwise[key]:
                    metric, pairwise = 'multiset', key
                    break
            else:
                return None
        elif self._distance_function is metrics.euclidean:
            metric, pairwise = 'euclidean', 'prod'
        elif self._distance_function is metrics.cosine_similarity:
            metric, pairwise = 'cosine', 'prod'
        else:
            return None
        options = dict(metric=metric, pairwise=pairwise,
                       do_balance=bool(self._do_balance),
                       do_scale=bool(self._do_scale), down=bool(self._down))
        if self._do_positive:
            # pair-dependent mask and totals: device path per pair only (no matrix form)
            options['do_positive'] = True
        return options

    # -------------------------------------------------------------- host path
    def _collapse(self, vector, start, length):
        """Sum the four quarters of ``vector[start:start+length]``
        (kpal/kdistlib.py:53-69)."""
        return np.reshape(vector[start:start + length], (4, length // 4)).sum(axis=1)

    def _dynamic_smooth(self, left, right, start, length):
        """Recursive collapse of sub-profiles whose summary is at or below the
        threshold in either profile (kpal/kdistlib.py:71-114)."""
        if length == 1:
            return
        left_c = self._collapse(left.counts, start, length)
        right_c = self._collapse(right.counts, start, length)
        if min(self._function(left_c), self._function(right_c)) <= self._threshold:
            left.counts[start] = left_c.sum()
            right.counts[start] = right_c.sum()
            left.counts[start + 1:start + length] = 0
            right.counts[start + 1:start + length] = 0
            return
        quarter = length // 4
        for i in range(4):
            self._dynamic_smooth(left, right, start + i * quarter, quarter)

    def dynamic_smooth(self, left, right):
        """Smooth two profiles in place (kpal/kdistlib.py:116-124)."""
        self._dynamic_smooth(left, right, 0, left.number)

    def _host_distance(self, left, right):
        """The reference pipeline for the options outside the GPU scope
        (kpal/kdistlib.py:136-161)."""
        left = left.copy()
        right = right.copy()
        if self._do_balance:
            left.balance()
            right.balance()
        if self._do_positive:
            left.counts, right.counts = (metrics.positive(left.counts, right.counts),
                                         metrics.positive(right.counts, left.counts))
        if self._do_smooth:
            self.dynamic_smooth(left, right)
        if self._do_scale:
            left_scale, right_scale = metrics.get_scale(left.counts, right.counts)
            if self._down:
                left_scale, right_scale = metrics.scale_down(left_scale, right_scale)
            left.counts = left.counts * left_scale
            right.counts = right.counts * right_scale
        if not self._distance_function:
            return metrics.multiset(left.counts, right.counts, self._pairwise)
        return self._distance_function(left.counts, right.counts)

    # ------------------------------------------------------------------ public
    def distance(self, left, right):
        """
        Distance between two profiles (kpal/kdistlib.py:126-161).  The inputs
        are never modified.
        """
        options = self._gpu_options()
        if options is None or not (_integer_counts(left) and _integer_counts(right)):
            # the device path works on int64 counts; anything else (float counts from a
            # custom merger, say) keeps the reference's dtype-agnostic NumPy pipeline
            return self._host_distance(left, right)
        return np.float64(_cabi.pair_distance(left.counts, right.counts, **options))


def _integer_counts(profile):
    """True when `profile` holds integer counts that int64 represents exactly (the device
    path never truncates: other dtypes are routed to the host pipeline)."""
    counts = profile.counts
    dtype = getattr(counts, 'dtype', None)
    if dtype is None:
        counts = np.asarray(counts)
        dtype = counts.dtype
    if dtype.kind == 'u' and dtype.itemsize == 8:
        return bool(counts.size == 0 or counts.max() <= np.iinfo(np.int64).max)
    return dtype.kind in 'iub'


def distance_matrix_values(profiles, dist):
    """Symmetric ``[n][n]`` float64 matrix of ``dist`` over `profiles`; one
    GPU call when the options are in the fast path."""
    n = len(profiles)
    options = dist._gpu_options()
    if (options is not None and n > 0 and not options.get('do_positive')
            and all(_integer_counts(profile) for profile in profiles)):
        stacked = np.empty((n, profiles[0].number), dtype=np.int64)
        for i, profile in enumerate(profiles):
            if profile.number != stacked.shape[1]:
                raise ValueError('k-mer lengths of the profiles differ')
            stacked[i] = profile.counts
        return _cabi.distance_matrix(stacked, **options)
    values = np.zeros((n, n), dtype=np.float64)
    for i in range(1, n):
        for j in range(i):
            values[i, j] = values[j, i] = dist.distance(profiles[i], profiles[j])
    return values


def distance_matrix(profiles, output, precision, dist):
    """
    Write the distance matrix of `profiles` to `output`: the number of
    profiles, their names, then the lower triangle row by row with
    `precision` decimals (kpal/kdistlib.py:164-186).
    """
    n = len(profiles)
    values = distance_matrix_values(profiles, dist) if n > 1 else None
    write_matrix([profile.name for profile in profiles], values, output, precision)


def write_matrix(names, values, output, precision):
    """The text of kpal/kdistlib.py:176-186 for a computed ``[n][n]`` matrix:
    ``n``, the names, then the lower triangle (``kpal_format_matrix`` writes
    the digits exactly as ``'{0:.{precision}f}'.format`` does)."""
    n = len(names)
    header = '\n'.join([str(n)] + [str(name) for name in names]) + '\n'
    output.write(header)
    if n > 1:
        output.write(_cabi.format_matrix(values, precision))


def distance_matrix_from_file(input_handle, names, output, precision, dist):
    """
    ``kpal matrix`` without the list of N in-memory profiles of
    kpal/kmer.py:694-698: datasets are read straight into a pinned slab and
    uploaded slab by slab (``kpal_matrix_push``).  Returns ``False`` when
    `dist` needs the host pipeline (the caller then takes the reference route).
    """
    options = dist._gpu_options()
    if options is None or options.get('do_positive'):
        return False
    group = input_handle['profiles']
    first = group[names[0]]
    number = int(first.shape[0])
    length = _length_of(number)
    with _cabi.MatrixSession(len(names), length, **options) as session:
        filled = 0
        for name in names:
            dataset = group[name]
            if int(dataset.shape[0]) != number:
                raise ValueError(LENGTH_ERROR)
            row = session.slab[filled]
            if hasattr(dataset, 'read_direct'):
                dataset.read_direct(row)
            else:
                row[...] = dataset[...]
            filled += 1
            if filled == session.slab.shape[0]:
                session.push(filled)
                filled = 0
        if filled:
            session.push(filled)
        values = session.finish()
    write_matrix(names, values, output, precision)
    return True


def _length_of(number):
    """*k* of a count vector of `number` entries (kpal/klib.py:58-61)."""
    length = int(number).bit_length() // 2
    if number < 4 or 4 ** length != number:
        raise ValueError('profile length %d is not a power of 4' % number)
    return length
