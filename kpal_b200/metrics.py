"""
Metrics and helper functions -- host-side mirror of the reference's
``kpal.metrics`` (same names, arguments and results; reference
kpal/metrics.py:22-179).

These NumPy functions are the *host path* the north star keeps on the CPU:
they serve the analysis helpers (``showbalance``, ``distr``, mergers) and the
pair-dependent options of ``ProfileDistance`` that are out of the GPU scope
(``do_positive``, ``do_smooth``, custom ``pairwise`` callables; SURVEY.md
section 8a row D7).  The N x N matrix / pair distance for the built-in options
never comes through here: ``kdistlib`` sends it to the CUDA kernels and fails
loudly if they are unavailable.
"""
from collections import Counter

import numpy as np


def distribution(vector):
    """Sorted ``(value, count)`` pairs of the values in `vector`
    (kpal/metrics.py:22-33)."""
    return sorted(Counter(vector).items())


def vector_length(vector):
    """Euclidean length of `vector` (kpal/metrics.py:36-46)."""
    return np.sqrt(np.dot(vector, vector))


def get_scale(left, right):
    """Scaling factors from the totals: the vector with the smaller sum gets
    ``big / small``, the other 1.0 (kpal/metrics.py:49-72)."""
    left_sum = np.sum(left)
    right_sum = np.sum(right)
    if left_sum < right_sum:
        return right_sum / left_sum, 1.0
    return 1.0, left_sum / right_sum


def scale_down(left, right):
    """Normalise two scaling factors by the larger one
    (kpal/metrics.py:75-86)."""
    largest = max(left, right)
    return left / largest, right / largest


def positive(vector, mask):
    """`vector` with the positions where `mask` is zero set to zero
    (kpal/metrics.py:89-98)."""
    return np.multiply(vector, np.asanyarray(mask, dtype=bool))


def multiset(left, right, pairwise):
    """Multiset distance: mean of ``pairwise`` over the positions where
    either vector is non-zero, with ``+ 1`` in the denominator
    (kpal/metrics.py:101-123).  `pairwise` must be vectorised."""
    left = np.asanyarray(left)
    right = np.asanyarray(right)
    keep = np.nonzero(np.logical_or(left, right))
    terms = pairwise(left[keep], right[keep])
    return terms.sum() / (len(terms) + 1)


def euclidean(left, right):
    """Euclidean distance (kpal/metrics.py:126-135)."""
    return vector_length(np.subtract(left, right))


def cosine_similarity(left, right):
    """Cosine similarity (kpal/metrics.py:138-147)."""
    return np.dot(left, right) / (vector_length(left) * vector_length(right))


#: Vector distance functions (kpal/metrics.py:151-155).
vector_distance = {
    "default": None,
    "euclidean": euclidean,
    "cosine": cosine_similarity,
}


def _pairwise_prod(x, y):
    return abs(x - y) / ((x + 1) * (y + 1))


def _pairwise_sum(x, y):
    return abs(x - y) / (x + y + 1)


#: Pairwise distance functions (kpal/metrics.py:159-162).
pairwise = {
    "prod": _pairwise_prod,
    "sum": _pairwise_sum,
}

#: Summary functions (kpal/metrics.py:166-170).
summary = {
    "min": np.min,
    "average": np.mean,
    "median": np.median,
}

#: Merge functions (kpal/metrics.py:174-179).
mergers = {
    "sum": lambda x, y: x + y,
    "xor": lambda x, y: (x + y) * np.logical_xor(x, y),
    "int": lambda x, y: x * np.asanyarray(y, dtype=bool),
    "nint": lambda x, y: x * np.logical_not(y),
}
