"""
TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/kpal_oracle.c (the
plain-C restatement of the kPAL hot path).  Used by tests/ and by bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkpal_oracle.so")
_lib = None

METRICS = {"multiset": 0, "euclidean": 1, "cosine": 2}
PAIRWISE = {"prod": 0, "sum": 1}


def build(force=False):
    src = os.path.join(_HERE, "kpal_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B",
                               "libkpal_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        L.oracle_count.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.c_void_p]
        L.oracle_count.restype = None
        L.oracle_count_mt.argtypes = [c.c_void_p, c.c_size_t, c.c_int,
                                      c.c_void_p, c.c_int]
        L.oracle_count_mt.restype = None
        L.oracle_balance.argtypes = [c.c_void_p, c.c_int]
        L.oracle_balance.restype = None
        L.oracle_reverse_complement.argtypes = [c.c_uint64, c.c_int]
        L.oracle_reverse_complement.restype = c.c_uint64
        L.oracle_distance.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t,
                                      c.c_int, c.c_int, c.c_int, c.c_int]
        L.oracle_distance.restype = c.c_double
        L.oracle_distance_matrix.argtypes = [
            c.c_void_p, c.c_size_t, c.c_size_t, c.c_int, c.c_int, c.c_int,
            c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int]
        L.oracle_distance_matrix.restype = None
        L.oracle_max_threads.restype = c.c_int
        _lib = L
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def count_bytes(buf, k, threads=1):
    """Count over raw bytes (every non-ACGTacgt byte splits)."""
    buf = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8)
                               if not isinstance(buf, np.ndarray) else buf)
    counts = np.zeros(4 ** k, dtype=np.int64)
    lib().oracle_count_mt(buf.ctypes.data, buf.size, k, counts.ctypes.data,
                          int(threads))
    return counts


def count_sequences(sequences, k, threads=1):
    parts = []
    for s in sequences:
        parts.append(s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s))
    return count_bytes(b"\n".join(parts), k, threads)


def balance(counts):
    counts = np.array(counts, dtype=np.int64)
    k = int(round(np.log(counts.size) / np.log(4)))
    lib().oracle_balance(counts.ctypes.data, k)
    return counts


def distance(left, right, do_balance=False, do_scale=False, down=False,
             metric="multiset", pairwise="prod"):
    left = np.ascontiguousarray(left, dtype=np.int64)
    right = np.ascontiguousarray(right, dtype=np.int64)
    if do_balance:
        left, right = balance(left), balance(right)
    return float(lib().oracle_distance(
        left.ctypes.data, right.ctypes.data, left.size, int(do_scale),
        int(down), METRICS[metric], PAIRWISE[pairwise]))


def distance_matrix(profiles, do_balance=False, do_scale=False, down=False,
                    metric="multiset", pairwise="prod", threads=1):
    profiles = np.ascontiguousarray(profiles, dtype=np.int64)
    n_prof, n = profiles.shape
    k = int(round(np.log(n) / np.log(4)))
    out = np.zeros((n_prof, n_prof), dtype=np.float64)
    lib().oracle_distance_matrix(
        profiles.ctypes.data, n_prof, n, k, int(do_balance), int(do_scale),
        int(down), METRICS[metric], PAIRWISE[pairwise], out.ctypes.data,
        int(threads))
    return out
