"""
TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference sources.

Imports ``kpal.klib`` / ``kpal.kdistlib`` / ``kpal.metrics`` straight from the
read-only reference tree (``/root/reference`` in the build container) so that

  * ``tests/golden/make_golden.py`` can generate golden vectors, and
  * the CPU test-suite can pin ``oracle/kpal_oracle.py`` against the real thing.

The reference cannot be imported as is: ``future``, ``h5py``, ``Bio`` and
``semantic_version`` are not installed in this image.
None of them carries hot-path arithmetic, so tiny stand-in modules are put in
``sys.modules`` *before* the import.  ``Bio.SeqIO.parse`` is served by the
oracle's own FASTA reader (documented Biopython FastaIterator behaviour).

Nothing in the product package (``kpal_b200``) may import this module.
``/root/reference`` does not exist on the GPU box; the offline install of the
reference in ``baseline/_ref`` (git-ignored, it travels with the snapshot) does,
and ``bench.py --impl reference`` uses it there to time the reference's own
Python beside the C port.  The ``-m gpu`` tests and ``smoke()`` never do.
"""
import builtins
import contextlib
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
#: the read-only reference tree of the build container, else the (git-ignored)
#: offline install `pip install --no-deps --target baseline/_ref /root/reference`
REFERENCE_ROOTS = ("/root/reference",
                   os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))


def reference_root():
    for root in REFERENCE_ROOTS:
        if os.path.isfile(os.path.join(root, "kpal", "klib.py")):
            return root
    return None


def available():
    return reference_root() is not None


class _Record(object):
    """Minimal stand-in for Bio.SeqRecord (only .seq and .name are used,
    reference kpal/klib.py:111,131-132)."""

    def __init__(self, name, seq):
        self.name = name
        self.id = name
        self.seq = seq


def _seqio_parse(handle, fmt):
    from . import kpal_oracle
    assert fmt == "fasta"
    text = handle.read()
    for name, seq in kpal_oracle.parse_fasta(text):
        yield _Record(name, seq)


def _install_stubs():
    if "future" not in sys.modules:
        future = types.ModuleType("future")
        fb = types.ModuleType("future.builtins")
        for n in ("next", "range", "str", "zip", "int", "object", "bytes",
                  "map", "filter", "open", "super"):
            setattr(fb, n, getattr(builtins, n))
        fu = types.ModuleType("future.utils")
        fu.native = lambda x: x
        fs = types.ModuleType("future.standard_library")
        fs.hooks = contextlib.nullcontext
        fs.install_aliases = lambda: None
        future.builtins, future.utils, future.standard_library = fb, fu, fs
        sys.modules.update({"future": future, "future.builtins": fb,
                            "future.utils": fu,
                            "future.standard_library": fs})
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.File = object
        sys.modules["h5py"] = h5
    if "semantic_version" not in sys.modules:
        sv = types.ModuleType("semantic_version")
        sv.Version = lambda s: s
        sv.Spec = sv.SimpleSpec = lambda s: s
        sys.modules["semantic_version"] = sv
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        seqio = types.ModuleType("Bio.SeqIO")
        seqio.parse = _seqio_parse
        bioseq = types.ModuleType("Bio.Seq")
        bioseq.Seq = str
        bio.SeqIO, bio.Seq = seqio, bioseq
        sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio,
                            "Bio.Seq": bioseq})


_cache = {}


def load():
    """Return (klib, kdistlib, metrics) of the unmodified reference, or raise
    RuntimeError when the reference tree is not present."""
    if "mods" in _cache:
        return _cache["mods"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected at %s)"
                           % (REFERENCE_ROOTS,))
    _install_stubs()
    if "kpal" in sys.modules and not getattr(
            sys.modules["kpal"], "__file__", "").startswith(root):
        raise RuntimeError("a different 'kpal' package is already imported")
    sys.path.insert(0, root)
    try:
        klib = importlib.import_module("kpal.klib")
        kdistlib = importlib.import_module("kpal.kdistlib")
        metrics = importlib.import_module("kpal.metrics")
    finally:
        sys.path.remove(root)
    _cache["mods"] = (klib, kdistlib, metrics)
    return _cache["mods"]
