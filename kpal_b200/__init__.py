"""
kpal_b200 -- B200-native (sm_100a) implementation of kPAL's data-parallel hot
path: *k*-mer profile construction from FASTA (with reverse-complement
balancing) and the N x N profile distance matrix, behind kPAL's own Python
API (``klib.Profile``, ``kdistlib.ProfileDistance`` / ``distance_matrix``,
``metrics``) and ``kpal count`` / ``kpal matrix`` command lines.

The kernels live in ``libkpal_b200.so`` (C ABI: ``include/kpal_b200.h``; sources
in ``kpal_b200/csrc``) and are reached through ctypes (``kpal_b200._cabi``).
"""
__version__ = '0.1.0'

from . import _cabi, metrics, klib, kdistlib  # noqa: E402,F401
from .klib import Profile  # noqa: E402,F401
from .kdistlib import ProfileDistance, distance_matrix  # noqa: E402,F401
