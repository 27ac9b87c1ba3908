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


def _table_to_host(table_ptr, k, balance, device):
    """Device u32 table -> the int64 profile as a NumPy array (widen + balance + the
    narrow device->host copy of ``kpal_dev_table_to_host``)."""
    import torch
    L = _cabi.load()
    out = np.empty(4 ** k, dtype=np.int64)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    _cabi.check(L.kpal_dev_table_to_host(table_ptr, 32, int(k), int(bool(balance)), _cabi.ptr(out), stream))
    return out


def _gpu_finalize(table, k, balance):
    return _table_to_host(table.data_ptr(), k, balance, table.device)


def count_fasta_distributed(fasta_shard, k, balance=False, group=None, device=None,
                            count_shard=None, finalize=None, reduce='nccl'):
    """
    ``Profile.from_fasta`` (+ ``balance``) over FASTA text that is sharded
    across the ranks of `group`: every rank passes ITS shard (cut at record
    boundaries, see :func:`split_fasta`); rank 0 gets the ``int64[4**k]``
    profile, the other ranks ``None``.

    `reduce` = ``'nccl'`` (default) sums the tables with ``dist.reduce``,
    ``'peer'`` over NVLink peer memory (:class:`PeerReducer`; GPU ranks, u32
    counters).  For a single call the NCCL reduce is the faster one (the
    peer path has to map its inboxes first; measured 0.455 vs 0.421 ms per
    100 Mbp step at 2 GPUs even with the mapping kept, profiles/README.md).  `count_shard` / `finalize` are injection points for the
    CPU (gloo) tests of the sharding + reduce logic; by default both run on
    the GPU.
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
        if reduce == 'peer' and table.is_cuda and table.dtype == torch.int32:
            reducer = PeerReducer(k, 32, group=group)
            try:
                stream = ctypes.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)
                summed = reducer.reduce(table.data_ptr(), stream)
                if summed is None:
                    return None
                return _table_to_host(summed, k, balance, table.device)
            finally:
                reducer.close()
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


class PeerReducer(object):
    """
    Sum of the per-rank counter tables onto rank `root` over NVLink peer memory
    (``csrc/peer_reduce.cu``): an all-to-all of table slices into per-rank
    inboxes, a local sum of every inbox, and a peer store of the summed slice
    into the root's table -- two kernels around two stream-ordered barriers
    (1-element all-reduces) instead of ``dist.reduce``.

    One process per GPU; inboxes and the root table are ``cudaMalloc`` buffers
    shared through CUDA IPC handles (exchanged with ``all_gather_object``).
    ``reduce(table_ptr)`` returns the device pointer of the summed table on the
    root (valid until the next call) and ``None`` elsewhere.
    """

    def __init__(self, k, counter_bits=32, group=None, root=0):
        import torch
        import torch.distributed as dist
        self._L = L = _cabi.load()
        self.k, self.bits, self.group, self.root = int(k), int(counter_bits), group, int(root)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._opened = []
        inbox_bytes = int(L.kpal_peer_inbox_bytes(self.k, self.bits, self.world))
        if inbox_bytes == 0:
            raise ValueError('bad k / counter_bits / world for the peer reduce')
        self._inbox = L.kpal_dev_alloc(inbox_bytes)
        self._root_table = L.kpal_dev_alloc(4 ** self.k * self.bits // 8) if self.rank == self.root else None
        if not self._inbox or (self.rank == self.root and not self._root_table):
            raise MemoryError('kpal_dev_alloc failed for the peer-reduce buffers')
        mine = {'inbox': self._export(self._inbox),
                'table': self._export(self._root_table) if self._root_table else None}
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self._inboxes = (ctypes.c_void_p * self.world)()
        error = None
        try:
            for r in range(self.world):
                self._inboxes[r] = self._inbox if r == self.rank else self._open(handles[r]['inbox'])
            self.root_table = (self._root_table if self.rank == self.root
                               else self._open(handles[self.root]['table']))
        except (RuntimeError, ValueError, MemoryError) as exc:        # cudaIpcOpenMemHandle refused (no peer access / IPC namespace)
            error = exc
        # everyone has mapped everyone -- or everyone learns that somebody could not, so that
        # all ranks leave the constructor the same way (no rank is left in a collective)
        self._flag = torch.tensor([0 if error is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(self._flag, group=group)
        failed = int(self._flag.item())
        self._flag.zero_()
        torch.cuda.synchronize()
        if failed:
            self.close()
            raise RuntimeError('peer-memory reduce unavailable: %d rank(s) could not map the peers (%s)'
                               % (failed, error if error is not None else 'failure on another rank'))

    def _export(self, dev_ptr):
        handle = ctypes.create_string_buffer(64)
        _cabi.check(self._L.kpal_ipc_export(dev_ptr, handle))
        return handle.raw

    def _open(self, handle):
        out = ctypes.c_void_p()
        _cabi.check(self._L.kpal_ipc_open(handle, ctypes.byref(out)))
        self._opened.append(out.value)
        return out.value

    def barrier(self):
        """Stream-ordered cross-GPU barrier: nobody's stream passes it before
        every rank's stream has reached it."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def reduce(self, table_ptr, stream):
        L = self._L
        _cabi.check(L.kpal_dev_reduce_push(table_ptr, self.bits, self.k, self.rank, self.world,
                                           self._inboxes, stream))
        self.barrier()
        _cabi.check(L.kpal_dev_reduce_collect(self._inbox, self.bits, self.k, self.rank, self.world,
                                              self.root_table, stream))
        self.barrier()
        return self._root_table if self.rank == self.root else None

    def count_and_reduce(self, codes_ptr, valid_ptr, n_bases, table_ptr, stream):
        """Count this rank's packed stream and sum all ranks' tables onto the
        root in one go (``kpal_dev_count_packed_push``): on the radix path the
        all-to-all is fused into the count's second pass.  `table_ptr` is a
        zeroed scratch table.  Returns like :meth:`reduce`."""
        L = self._L
        fused = ctypes.c_int()
        _cabi.check(L.kpal_dev_count_packed_push(codes_ptr, valid_ptr, int(n_bases), self.k, table_ptr,
                                                 self.bits, self.rank, self.world, self._inboxes,
                                                 stream, ctypes.byref(fused)))
        self.fused = bool(fused.value)
        self.barrier()
        _cabi.check(L.kpal_dev_reduce_collect(self._inbox, self.bits, self.k, self.rank, self.world,
                                              self.root_table, stream))
        self.barrier()
        return self._root_table if self.rank == self.root else None

    def close(self):
        import torch
        torch.cuda.synchronize()
        for p in self._opened:
            self._L.kpal_ipc_close(p)
        self._opened = []
        if self._inbox:
            self._L.kpal_dev_free(self._inbox)
            self._inbox = None
        if self._root_table:
            self._L.kpal_dev_free(self._root_table)
            self._root_table = None


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
