"""
Multi-GPU drivers for the two shardable parts of the path,
one process per GPU (``torchrun``), ``torch.distributed`` for the plumbing:

* counting -- records are sharded over the ranks (cut only at record
  boundaries, so no halo is needed), every rank accumulates its windows into
  its own ``4**k`` counter table on its GPU, the tables are summed onto rank 0
  with one ``reduce`` (NCCL over NVLink) and finalised (widen + balance) once.
  Balance is linear, so it commutes with the sum (``klib.py:285-298``).
* distance matrix -- every rank holds the prepared profile set, takes a
  contiguous slice of the upper-triangle tile list, and the (disjointly filled)
  ``N x N`` results are summed onto rank 0.

The sharding helpers are pure host logic and are tested with the ``gloo``
backend on CPU (``tests/test_multigpu.py``); the compute itself always goes
through the CUDA library.
"""
import ctypes

import numpy as np

from . import _cabi


# ------------------------------------------------------------------ sharding
def balanced_ranges(sizes, parts):
    """Cut ``len(sizes)`` consecutive items into `parts` contiguous ranges of
    about equal total size.  Returns ``parts`` ``(begin, end)`` pairs covering
    every item exactly once (ranges may be empty)."""
    sizes = np.asarray(sizes, dtype=np.float64)
    n = len(sizes)
    if parts < 1:
        raise ValueError('parts must be >= 1')
    cum = np.concatenate(([0.0], np.cumsum(sizes)))
    total = cum[-1]
    cuts = [0]
    for p in range(1, parts):
        target = total * p / parts
        cut = int(np.searchsorted(cum, target, side='left'))
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(parts)]


def split_fasta(text, parts):
    """Byte ranges that cut FASTA `text` (bytes) at record boundaries (a ``>``
    at the start of a line) into `parts` shards of about equal size.  Text
    before the first record stays with the first shard."""
    if isinstance(text, str):
        text = text.encode('latin-1', 'replace')
    n = len(text)
    bounds = [0]
    for p in range(1, parts):
        pos = max(n * p // parts, bounds[-1])
        if pos == 0 and text[:1] == b'>':
            cut = 0
        else:
            hit = text.find(b'\n>', max(pos - 1, 0))
            cut = n if hit < 0 else hit + 1
        bounds.append(max(cut, bounds[-1]))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]


def tile_range(n_tiles, rank, world):
    """Contiguous slice of the tile list for `rank`."""
    return n_tiles * rank // world, n_tiles * (rank + 1) // world


# ------------------------------------------------------------------ counting
def _gpu_count_shard(fasta_bytes, k, device):
    """This rank's windows as a device ``int32`` tensor of ``4**k`` u32 counters."""
    import torch
    L = _cabi.load()
    table = torch.zeros(4 ** k, dtype=torch.int32, device=device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    buf = np.frombuffer(fasta_bytes, dtype=np.uint8)
    n_bases = ctypes.c_uint64()
    _cabi.check(L.kpal_count_fasta_to_dev(_cabi.ptr(buf) if buf.size else None, buf.size, int(k),
                                          table.data_ptr(), 32, stream, ctypes.byref(n_bases)))
    return table


def _gpu_finalize(table, k, balance):
    import torch
    L = _cabi.load()
    out = torch.empty(4 ** k, dtype=torch.int64, device=table.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)
    _cabi.check(L.kpal_dev_finalize_counts(table.data_ptr(), 32, int(k), int(bool(balance)),
                                           out.data_ptr(), stream))
    return out.cpu().numpy()


def count_fasta_distributed(fasta_shard, k, balance=False, group=None, device=None,
                            count_shard=None, finalize=None):
    """
    ``Profile.from_fasta`` (+ ``balance``) over FASTA text that is sharded
    across the ranks of `group`: every rank passes ITS shard (cut at record
    boundaries, see :func:`split_fasta`); rank 0 gets the ``int64[4**k]``
    profile, the other ranks ``None``.

    `count_shard` / `finalize` are injection points for the CPU (gloo) tests
    of the sharding + reduce logic; by default both run on the GPU.
    """
    import torch
    import torch.distributed as dist
    _cabi._check_k(k)
    if isinstance(fasta_shard, str):
        fasta_shard = fasta_shard.encode('latin-1', 'replace')
    if count_shard is None:
        _cabi.require_gpu()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        table = _gpu_count_shard(fasta_shard, k, device)
    else:
        table = count_shard(fasta_shard, k)
    distributed = dist.is_available() and dist.is_initialized()
    if distributed and dist.get_world_size(group) > 1:
        # u32 counters are exact while the GLOBAL window total stays below 2**32
        total = torch.tensor([len(fasta_shard)], dtype=torch.int64, device=table.device)
        dist.all_reduce(total, group=group)
        if int(total.item()) >= 2 ** 32:
            table = table.to(torch.int64) & 0xffffffff
        dist.reduce(table, dst=0, group=group)
        if dist.get_rank(group) != 0:
            return None
    if table.dtype == torch.int64:
        counts = table.cpu().numpy()
        if balance:
            work = np.ascontiguousarray(counts)
            counts = _cabi.balance(work) if finalize is None else finalize(work, k, True)
        return counts
    return (finalize or _gpu_finalize)(table, k, balance)


# ------------------------------------------------------------------ distances
def distance_matrix_distributed(profiles, metric='multiset', pairwise='prod', do_balance=False,
                                do_scale=False, down=False, group=None, device=None):
    """
    The symmetric ``[n][n]`` distance matrix of `profiles` (C-contiguous
    ``[n][4**k]`` int64, the same array on every rank) with the upper-triangle
    tiles sharded over the ranks of `group`.  Rank 0 gets the matrix.
    """
    import torch
    import torch.distributed as dist
    _cabi.require_gpu()
    L = _cabi.load()
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    profiles = np.ascontiguousarray(profiles, dtype=np.int64)
    n, size = profiles.shape
    k = _cabi._k_of(size)
    stride = int(L.kpal_prepared_stride(k))
    sp = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    need_p = metric == 'multiset' and pairwise == 'prod'
    F = torch.empty((n, stride), dtype=torch.float64, device=device)
    P = torch.empty((n, stride), dtype=torch.float64, device=device) if need_p else None
    bitmap = torch.empty((n, stride // 32), dtype=torch.int32, device=device)
    totals = torch.empty(n, dtype=torch.float64, device=device)
    norm2 = torch.empty(n, dtype=torch.float64, device=device)
    slab = max(1, min(n, (1 << 28) // (size * 8)))
    for r0 in range(0, n, slab):
        m = min(slab, n - r0)
        counts = torch.from_numpy(profiles[r0:r0 + m]).to(device)
        _cabi.check(L.kpal_dev_profiles_prepare(
            counts.data_ptr(), m, k, int(bool(do_balance)), int(bool(do_scale)), F[r0].data_ptr(),
            P[r0].data_ptr() if need_p else None, bitmap[r0].data_ptr(), totals[r0:].data_ptr(),
            norm2[r0:].data_ptr(), sp))
    order = None
    if do_scale:
        order = torch.empty(n, dtype=torch.int32, device=device)
        _cabi.check(L.kpal_dev_order_by_total(totals.data_ptr(), n, int(bool(down)),
                                              order.data_ptr(), sp))
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    begin, end = tile_range(int(L.kpal_distance_num_tiles(n)), rank, world)
    out = torch.zeros((n, n), dtype=torch.float64, device=device)
    _cabi.check(L.kpal_dev_distance_tiles(
        F.data_ptr(), P.data_ptr() if need_p else None, bitmap.data_ptr(), totals.data_ptr(),
        norm2.data_ptr(), order.data_ptr() if order is not None else None, n, k,
        _cabi.METRICS[metric], _cabi.PAIRWISE[pairwise], int(bool(do_scale)), int(bool(down)),
        begin, end, out.data_ptr(), sp))
    if world > 1:
        dist.reduce(out, dst=0, group=group)
        if rank != 0:
            return None
    return out.cpu().numpy()
