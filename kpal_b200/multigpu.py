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
def bind_to_gpu_cpus(device_index):
    """Pin the calling thread (and the threads it starts later: the library's host workers) to
    the CPUs next to GPU `device_index` (NVML's ideal affinity: the GPU's NUMA node), so that a
    rank's pinned buffers and its widening threads stay on the memory controller its PCIe link
    hangs off.  One process per GPU on a two-socket host otherwise leaves that to chance.  Returns
    the previous affinity (restore it with ``os.sched_setaffinity(0, previous)``) or None when
    NVML or the call is unavailable -- it is an optimisation, never an error."""
    import os
    try:
        import pynvml
        previous = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        ideal &= previous
        if not ideal:
            return None
        os.sched_setaffinity(0, ideal)
        return previous
    except Exception:
        return None


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
    """This rank's windows as a device tensor of ``4**k`` counters: u32 (``int32``), or
    u64 (``int64``) for a shard of 4 Gi bases or more."""
    import torch
    L = _cabi.load()
    # 32-bit counters are exact while the shard has fewer than 2**32 bases (as kpal_count_fasta)
    bits = 64 if len(fasta_bytes) >= 2 ** 32 else 32
    table = torch.zeros(4 ** k, dtype=torch.int64 if bits == 64 else torch.int32, device=device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    buf = np.frombuffer(fasta_bytes, dtype=np.uint8)
    n_bases = ctypes.c_uint64()
    _cabi.check(L.kpal_count_fasta_to_dev(_cabi.ptr(buf) if buf.size else None, buf.size, int(k),
                                          table.data_ptr(), bits, stream, ctypes.byref(n_bases)))
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

    `reduce` = ``'slices'`` (GPU ranks, ``balance=True``, k >= 6): the fused form --
    balance + narrow reduce-scatter over NVLink peer memory + distributed finalize
    (:class:`SliceReducer`), the slices meeting in shared host memory;
    ``'nccl'`` (default) sums the tables with ``dist.reduce``,
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
    distributed_now = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if reduce == 'slices' and count_shard is None and distributed_now and k >= 6 and balance:
        return _count_fasta_slices(fasta_shard, k, group, device)
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
        if int(total.item()) >= 2 ** 32 and table.dtype == torch.int32:
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


def _count_fasta_slices(fasta_shard, k, group, device):
    """count_fasta_distributed(balance=True) through :class:`SliceReducer`: every rank counts its
    shard, the balanced tables are reduce-scattered narrow over NVLink, every rank copies its
    slice of the profile into shared host memory, rank 0 returns the whole."""
    import torch
    L = _cabi.load()
    _cabi.require_gpu()
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    reducer = SliceReducer(k, group=group)
    shared = SharedProfile(k, group=group)
    try:
        buf = np.frombuffer(fasta_shard, dtype=np.uint8)
        table, bits = ctypes.c_void_p(), ctypes.c_int()
        _cabi.check(L.kpal_count_fasta_dev_table(_cabi.ptr(buf) if buf.size else None, buf.size, int(k),
                                                 ctypes.byref(table), ctypes.byref(bits), stream))
        reducer.push(table, bits.value, stream, n_bases=buf.size)
        begin, end = reducer.slice_range()
        reducer.collect_to_host(shared.array[begin:end], stream)
        import torch.distributed as dist
        dist.barrier(group=group)
        return shared.array.copy() if reducer.rank == 0 else None
    finally:
        shared.close()
        reducer.close()


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


class SliceReducer(object):
    """
    The fused form of the multi-GPU table sum (``csrc/peer_reduce.cu``): every rank balances
    its own table and stores the result, one byte per bin, straight into the inbox of the
    rank that owns the bin's slice (``push``: NVLink peer stores + one release flag per
    peer); every rank then sums the rows it received into its int64 slice of the final
    balanced profile (``collect`` / ``collect_to_host``: waits for the peers' flags in its own
    inbox, no host round trip, no NCCL call).  The profile stays sharded by slice
    ``[begin, end)`` = :meth:`slice_range`.

    One process per GPU; inboxes are ``cudaMalloc`` buffers shared through CUDA IPC handles.
    """

    def __init__(self, k, group=None):
        import torch
        import torch.distributed as dist
        self._L = L = _cabi.load()
        self.k, self.group = int(k), group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.epoch = 0
        self._wide = 0
        self._opened = []
        nbytes = int(L.kpal_slice_inbox_bytes(self.k, self.world))
        if nbytes == 0:
            raise ValueError('the sliced reduce needs 6 <= k <= 15 and at most 16 ranks')
        self._inbox = L.kpal_dev_alloc(nbytes)
        if not self._inbox:
            raise MemoryError('kpal_dev_alloc failed for the slice inbox')
        _cabi.check(L.kpal_dev_memset(self._inbox, 0, nbytes, None))
        _cabi.check(L.kpal_stream_sync(None))
        handle = ctypes.create_string_buffer(64)
        _cabi.check(L.kpal_ipc_export(self._inbox, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self._inboxes = (ctypes.c_void_p * self.world)()
        error = None
        try:
            for r in range(self.world):
                if r == self.rank:
                    self._inboxes[r] = self._inbox
                else:
                    out = ctypes.c_void_p()
                    _cabi.check(L.kpal_ipc_open(handles[r], ctypes.byref(out)))
                    self._opened.append(out.value)
                    self._inboxes[r] = out.value
        except (RuntimeError, ValueError, MemoryError) as exc:
            error = exc
        flag = torch.tensor([0 if error is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, group=group)
        torch.cuda.synchronize()
        if int(flag.item()):
            self.close()
            raise RuntimeError('sliced peer reduce unavailable: %d rank(s) could not map the peers (%s)'
                               % (int(flag.item()), error if error is not None else 'failure on another rank'))

    def slice_range(self, rank=None):
        rank = self.rank if rank is None else rank
        return (int(self._L.kpal_slice_begin(self.k, rank, self.world)),
                int(self._L.kpal_slice_begin(self.k, rank + 1, self.world)))

    def push(self, table_ptr, counter_bits, stream, n_bases=None):
        """Balance + narrow push of this rank's table; starts a new epoch.  `n_bases` (the bases
        this rank counted) picks the row form: bytes with escapes for counts >= 255 while the mean
        balanced count 2 * n_bases / 4^k stays below 64, else u32 rows.  The peers are told by
        the collect that follows on the same stream."""
        self.epoch += 1
        self._wide = 1 if (n_bases is not None and 2 * int(n_bases) >= 64 * 4 ** self.k) else 0
        _cabi.check(self._L.kpal_dev_slice_push(table_ptr, int(counter_bits), self.k, self.rank, self.world,
                                                self._inboxes, self.epoch, self._wide, stream))

    def collect(self, slice_ptr, stream):
        """Signal + this rank's int64 slice of the balanced profile -> device memory."""
        _cabi.check(self._L.kpal_dev_slice_collect(self._inboxes, self.k, self.rank, self.world, self.epoch,
                                                   self._wide, slice_ptr, stream))

    def collect_to_host(self, host_slice, stream):
        """... -> `host_slice` (a C-contiguous int64 array of the slice's length)."""
        _cabi.check(self._L.kpal_dev_slice_collect_to_host(self._inboxes, self.k, self.rank, self.world, self.epoch,
                                                           self._wide, _cabi.ptr(host_slice), stream))

    def close(self):
        import torch
        torch.cuda.synchronize()
        for p in self._opened:
            self._L.kpal_ipc_close(p)
        self._opened = []
        if self._inbox:
            self._L.kpal_dev_free(self._inbox)
            self._inbox = None


class SharedProfile(object):
    """An ``int64[4**k]`` profile in POSIX shared memory that every rank's process maps: the
    owners of the slices write their parts side by side (each over its own PCIe link, widened by
    its own host threads) and rank 0 reads the whole -- no gather through one process."""

    def __init__(self, k, group=None):
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.rank = dist.get_rank(group)
        self.group = group
        nbytes = 8 * 4 ** int(k)
        names = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            names[0] = self._shm.name
        dist.broadcast_object_list(names, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=names[0])
            try:                    # the creator unlinks it; keep this process's tracker out of it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, 'shared_memory')
            except Exception:
                pass
        self.array = np.ndarray((4 ** int(k),), dtype=np.int64, buffer=self._shm.buf)

    def close(self):
        import torch.distributed as dist
        self.array = None
        dist.barrier(group=self.group)
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()


# ------------------------------------------------------------- per-record counting
def count_by_record_distributed(fasta_text, k, balance=False, group=None, device=None,
                                gather=False, count_rows=None):
    """
    ``Profile.from_fasta_by_record`` (reference kpal/klib.py:114-133) with the records of
    `fasta_text` (the WHOLE file, the same bytes on every rank) sharded over the ranks of
    `group`: rank r takes the r-th of `world` contiguous byte ranges cut at record
    boundaries (:func:`split_fasta`, about equal bases per rank), counts its records on its
    GPU and gets ``(first, names, rows)`` -- the global index of its first record, the
    names of its records (``''`` for a nameless one: the caller numbers it ``first + i +
    1``, kpal/klib.py:129-132) and the dense ``[n_r][4**k]`` int64 rows.  Rows are
    independent, so there is no data-path collective: only the record
    counts are exchanged, to number the records globally.

    `gather` = True additionally concatenates all rows, in record order, on rank 0 (which
    then gets ``(0, all_names, all_rows)``; the others ``(first, names, None)``) -- for
    small inputs and tests; at scale every rank writes the profiles it counted.
    `count_rows` is the injection point of the CPU (gloo) tests.
    """
    import torch
    import torch.distributed as dist
    _cabi._check_k(k)
    if isinstance(fasta_text, str):
        fasta_text = fasta_text.encode('latin-1', 'replace')
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    begin, end = split_fasta(fasta_text, world)[rank]
    shard = fasta_text[begin:end]
    if rank > 0 and begin > 0 and shard[:1] != b'>':
        shard = b''                 # no record starts in this range (text without headers)
    codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(shard)
    n_local = len(names)
    if count_rows is None:
        _cabi.require_gpu()
        if device is not None:
            _cabi.check(_cabi.load().kpal_set_device(device.index if hasattr(device, 'index') else int(device)))
        rows = np.empty((n_local, 4 ** k), dtype=np.int64)
        batch = max(1, (256 << 20) // (8 * 4 ** k))
        for first in range(0, n_local, batch):
            n = min(batch, n_local - first)
            rows[first:first + n] = _cabi.count_by_record(codes, valid, n_bases, rec_starts, first, n, k,
                                                          balance=balance)
    else:
        rows = count_rows(shard, k, balance)
    first_record = 0
    if world > 1:
        counts = [None] * world
        dist.all_gather_object(counts, n_local, group=group)
        first_record = int(sum(counts[:rank]))
    if not gather or world == 1:
        return first_record, names, rows
    parts = [None] * world if rank == 0 else None
    dist.gather_object((names, rows), parts, dst=0, group=group)
    if rank != 0:
        return first_record, names, None
    all_names = [name for part in parts for name in part[0]]
    return 0, all_names, np.concatenate([part[1] for part in parts], axis=0)


# ------------------------------------------------------------------ distances
def shard_rows(n, rank, world):
    """Rows ``[begin, end)`` of an `n`-row profile set that rank `rank` uploads and
    prepares: equal shares of ``ceil(n / world)`` rows (the last ones may be shorter)."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


class _GpuMatrixOps(object):
    """The device side of :func:`distance_matrix_distributed` (CUDA library + torch
    tensors); the CPU (gloo) tests substitute an oracle-backed double with the same methods."""

    def __init__(self, n, k, options, device):
        import torch
        self.torch = torch
        self.L = _cabi.load()
        self.n, self.k, self.device = n, k, device
        self.metric = _cabi.METRICS[options['metric']]
        self.pairwise = _cabi.PAIRWISE[options['pairwise']]
        self.do_balance = int(bool(options['do_balance']))
        self.do_scale = int(bool(options['do_scale']))
        self.down = int(bool(options['down']))
        self.stride = int(self.L.kpal_prepared_stride(k))
        self.need_p = options['metric'] == 'multiset' and options['pairwise'] == 'prod'
        f64 = dict(dtype=torch.float64, device=device)
        self.F = torch.empty((n, self.stride), **f64)
        self.P = torch.empty((n, self.stride), **f64) if self.need_p else None
        self.bitmap = torch.empty((n, self.stride // 32), dtype=torch.int32, device=device)
        self.totals = torch.empty(n, **f64)
        self.norm2 = torch.empty(n, **f64)
        self.order = None
        self.tile_elems = int(self.L.kpal_distance_tile_elems())

    def _sp(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def arrays(self):
        """The prepared per-profile arrays, row-sharded like the profiles."""
        return [a for a in (self.F, self.P, self.bitmap, self.totals, self.norm2) if a is not None]

    def prepare(self, rows, begin):
        """Upload raw int64 `rows` (host) and prepare them as profiles begin, begin + 1, ..."""
        torch, L = self.torch, self.L
        size = 4 ** self.k
        slab = max(1, min(len(rows), (1 << 28) // (size * 8))) if len(rows) else 1
        for r0 in range(0, len(rows), slab):
            m = min(slab, len(rows) - r0)
            counts = torch.from_numpy(np.ascontiguousarray(rows[r0:r0 + m], dtype=np.int64)).to(self.device)
            at = begin + r0
            _cabi.check(L.kpal_dev_profiles_prepare(
                counts.data_ptr(), m, self.k, self.do_balance, self.do_scale, self.F[at].data_ptr(),
                self.P[at].data_ptr() if self.need_p else None, self.bitmap[at].data_ptr(),
                self.totals[at:].data_ptr(), self.norm2[at:].data_ptr(), self._sp()))

    def make_order(self):
        if self.do_scale:
            self.order = self.torch.empty(self.n, dtype=self.torch.int32, device=self.device)
            _cabi.check(self.L.kpal_dev_order_by_total(self.totals.data_ptr(), self.n, self.down,
                                                       self.order.data_ptr(), self._sp()))

    def num_tiles(self):
        return int(self.L.kpal_distance_num_tiles(self.n))

    def new_packed(self, n_tiles):
        return self.torch.zeros((max(n_tiles, 1), self.tile_elems), dtype=self.torch.float64, device=self.device)

    def tiles_packed(self, begin, end, packed):
        if end > begin:
            _cabi.check(self.L.kpal_dev_distance_tiles_packed(
                self.F.data_ptr(), self.P.data_ptr() if self.need_p else None, self.bitmap.data_ptr(),
                self.totals.data_ptr(), self.norm2.data_ptr(),
                self.order.data_ptr() if self.order is not None else None, self.n, self.k,
                self.metric, self.pairwise, self.do_scale, self.down, begin, end, packed.data_ptr(),
                self._sp()))

    def new_out(self):
        return self.torch.zeros((self.n, self.n), dtype=self.torch.float64, device=self.device)

    def unpack(self, packed, begin, end, diagonal, out):
        _cabi.check(self.L.kpal_dev_distance_unpack_tiles(
            packed.data_ptr(), self.totals.data_ptr(), self.norm2.data_ptr(),
            self.order.data_ptr() if self.order is not None else None, self.n, self.metric,
            self.pairwise, self.do_scale, begin, end, int(bool(diagonal)), out.data_ptr(), self._sp()))

    def to_host(self, out):
        return out.cpu().numpy()


def allgather_rows(arrays, n, group=None):
    """All-gather of row-sharded arrays, in place: every array has `n` rows on every rank,
    rank r holds valid data in rows :func:`shard_rows` ``(n, r, world)`` and receives the
    others' (one ``all_gather_into_tensor`` per array when the shards are equal, which NCCL
    runs in place over NVLink; one broadcast per rank and array otherwise)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    per = (n + world - 1) // world
    for a in arrays:
        if n == per * world:
            try:
                dist.all_gather_into_tensor(a, a[rank * per:(rank + 1) * per], group=group)
                continue
            except (RuntimeError, NotImplementedError):      # backend without the in-place form
                pass
        for r in range(world):
            b, e = shard_rows(n, r, world)
            if e > b:
                dist.broadcast(a[b:e], src=dist.get_global_rank(group, r) if group is not None else r,
                               group=group)


def distance_matrix_distributed(profiles, metric='multiset', pairwise='prod', do_balance=False,
                                do_scale=False, down=False, group=None, device=None, sharded=False,
                                n_total=None, ops=None):
    """
    The symmetric ``[n][n]`` distance matrix of a profile set over the ranks of `group`
    (the loop being cut up is kpal/kdistlib.py:179-184):

    1. rank r uploads and prepares only rows :func:`shard_rows` ``(n, r, world)`` -- 1/world of
       the host-to-device traffic and of the pre-pass per GPU;
    2. the prepared arrays (frequencies, ``x + 1``, non-zero bitmaps, totals, norms) are
       all-gathered over NVLink (:func:`allgather_rows`), so every GPU holds the whole set;
    3. the upper-triangle tiles are dealt out in contiguous, equal ranges (:func:`tile_range`);
    4. every rank's finished tiles travel to rank 0 in ONE gather of compact tile arrays
       (``N^2 / 2`` doubles in total) and are scattered into the matrix there.

    `profiles`: C-contiguous ``[n][4**k]`` int64 -- the whole set on every rank (each rank
    reads only its rows), or with ``sharded=True`` just this rank's rows (then `n_total` is
    required).  Rank 0 gets the matrix, the other ranks ``None``.
    """
    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    profiles = np.asarray(profiles)
    if profiles.ndim != 2:
        raise ValueError('profiles must be [rows][4**k]')
    n = int(n_total) if sharded else profiles.shape[0]
    if sharded and n_total is None:
        raise ValueError('sharded=True needs n_total')
    k = _cabi._k_of(profiles.shape[1])
    begin, end = shard_rows(n, rank, world)
    mine = profiles if sharded else profiles[begin:end]
    if len(mine) != end - begin:
        raise ValueError('rank %d must hold rows [%d, %d) of the profile set' % (rank, begin, end))
    options = dict(metric=metric, pairwise=pairwise, do_balance=do_balance, do_scale=do_scale, down=down)
    if ops is None:
        _cabi.require_gpu()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        ops = _GpuMatrixOps(n, k, options, device)
    ops.prepare(mine, begin)
    allgather_rows(ops.arrays(), n, group=group)
    ops.make_order()
    n_tiles = ops.num_tiles()
    t_begin, t_end = tile_range(n_tiles, rank, world)
    most = max(tile_range(n_tiles, r, world)[1] - tile_range(n_tiles, r, world)[0] for r in range(world))
    packed = ops.new_packed(most)
    ops.tiles_packed(t_begin, t_end, packed)
    if world == 1:
        out = ops.new_out()
        ops.unpack(packed, t_begin, t_end, True, out)
        return ops.to_host(out)
    parts = [ops.new_packed(most) for _ in range(world)] if rank == 0 else None
    dist.gather(packed, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    out = ops.new_out()
    for r in range(world):
        b, e = tile_range(n_tiles, r, world)
        ops.unpack(parts[r], b, e, r == 0, out)
    return ops.to_host(out)
