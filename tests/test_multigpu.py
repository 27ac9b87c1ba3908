"""
N > 1 path.  CPU part: the sharding helpers and the shard -> reduce -> finalise
flow of kpal_b200.multigpu with world_size 2 on the gloo backend (the count of
each shard is injected from the oracle: there is no GPU here, and the product
code has no CPU path of its own).  GPU part (-m gpu, needs >= 2 devices): the
same flow with the CUDA kernels and NCCL.
"""
import os
import socket

import numpy as np
import pytest

from kpal_b200 import _cabi, multigpu
from oracle import kpal_oracle as ko


def make_fasta(seed, n_records, max_len=400):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_records):
        n = int(rng.integers(0, max_len))
        seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=n).tobytes().decode()
        lines.append(">rec%05d desc" % i)
        lines.extend(seq[p:p + 70] for p in range(0, n, 70))
    return ("\n".join(lines) + "\n").encode()


def test_balanced_ranges():
    sizes = [5, 1, 1, 1, 10, 2, 2, 8]
    for parts in (1, 2, 3, 8, 11):
        ranges = multigpu.balanced_ranges(sizes, parts)
        assert len(ranges) == parts
        assert ranges[0][0] == 0 and ranges[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(b <= e for b, e in ranges)
    two = multigpu.balanced_ranges([10, 10, 10, 10], 2)
    assert two == [(0, 2), (2, 4)]
    assert multigpu.balanced_ranges([], 3) == [(0, 0), (0, 0), (0, 0)]


def test_split_fasta_cuts_at_record_boundaries():
    text = make_fasta(3, 200)
    for parts in (1, 2, 3, 7, 64):
        ranges = multigpu.split_fasta(text, parts)
        assert len(ranges) == parts and ranges[0][0] == 0 and ranges[-1][1] == len(text)
        total = np.zeros(4 ** 5, dtype=np.int64)
        for b, e in ranges:
            shard = text[b:e]
            assert shard == b"" or shard.startswith(b">")
            total += ko.count_fasta(shard, 5)
        assert np.array_equal(total, ko.count_fasta(text, 5))
    # leading junk stays with the first shard and is skipped there
    junk = b"junk line\nACGT\n" + text
    ranges = multigpu.split_fasta(junk, 3)
    total = sum(ko.count_fasta(junk[b:e], 4) for b, e in ranges)
    assert np.array_equal(total, ko.count_fasta(junk, 4))


def test_tile_range_partitions():
    for n_tiles in (0, 1, 7, 1056):
        for world in (1, 2, 3, 8):
            got = [multigpu.tile_range(n_tiles, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n_tiles
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, text, k, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = multigpu.split_fasta(text, world)[rank]

        def count_shard(shard, kk):          # stand-in for the CUDA count kernel (test only)
            return torch.from_numpy(ko.count_fasta(shard, kk).astype(np.int32))

        def finalize(table, kk, balance):    # stand-in for finalize_kernel (test only)
            counts = (table.numpy() if hasattr(table, "numpy") else table).astype(np.int64)
            return ko.balance(counts) if balance else counts

        got = multigpu.count_fasta_distributed(text[b:e], k, balance=True, count_shard=count_shard,
                                               finalize=finalize)
        if rank == 0:
            np.save(out_path, got)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_count_distributed_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    text = make_fasta(11, 300)
    k = 6
    out = str(tmp_path / "counts.npy")
    mp.spawn(_gloo_worker, args=(world, _free_port(), text, k, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), ko.balance(ko.count_fasta(text, k)))


class _OracleMatrixOps(object):
    """CPU stand-in for multigpu._GpuMatrixOps (test only): the 'prepared arrays' are the raw
    counts, a 'tile' is a run of 7 pairs of the lower triangle, the distances come from the
    oracle.  What is under test is the host logic around it: who prepares which rows, the
    all-gather, the tile ranges, the gather of packed tiles and their scatter on rank 0."""
    TILE = 7

    def __init__(self, n, size, options):
        import torch
        self.torch, self.n, self.options = torch, n, options
        self.counts = torch.full((n, size), -1, dtype=torch.int64)
        self.pairs = [(i, j) for i in range(1, n) for j in range(i)]

    def arrays(self):
        return [self.counts]

    def prepare(self, rows, begin):
        if len(rows):
            self.counts[begin:begin + len(rows)] = self.torch.from_numpy(np.ascontiguousarray(rows))

    def make_order(self):
        assert int(self.counts.min()) >= 0, "rows missing after the all-gather"

    def num_tiles(self):
        return (len(self.pairs) + self.TILE - 1) // self.TILE

    def new_packed(self, n_tiles):
        return self.torch.zeros((max(n_tiles, 1), self.TILE), dtype=self.torch.float64)

    def tiles_packed(self, begin, end, packed):
        c = self.counts.numpy()
        for t in range(begin, end):
            for e, (i, j) in enumerate(self.pairs[t * self.TILE:(t + 1) * self.TILE]):
                packed[t - begin, e] = ko.distance(c[i], c[j], **self.options)

    def new_out(self):
        return self.torch.full((self.n, self.n), np.nan, dtype=self.torch.float64)

    def unpack(self, packed, begin, end, diagonal, out):
        for t in range(begin, end):
            for e, (i, j) in enumerate(self.pairs[t * self.TILE:(t + 1) * self.TILE]):
                out[i, j] = out[j, i] = packed[t - begin, e]
        if diagonal:
            for i in range(self.n):
                out[i, i] = 0.0

    def to_host(self, out):
        return out.numpy()


def _gloo_matrix_worker(rank, world, port, profiles, sharded, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(profiles)
        options = dict(do_scale=True, metric="multiset", pairwise="sum")
        ops = _OracleMatrixOps(n, profiles.shape[1], options)
        b, e = multigpu.shard_rows(n, rank, world)
        # every rank may only ever look at its own rows
        mine = profiles[b:e] if sharded else np.where(
            ((np.arange(n) >= b) & (np.arange(n) < e))[:, None], profiles, -7)
        got = multigpu.distance_matrix_distributed(mine, sharded=sharded, n_total=n, ops=ops, **options)
        if rank == 0:
            np.save(out_path, got)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,sharded", [(2, 8, False), (3, 9, True), (3, 10, True), (4, 3, True)])
def test_matrix_distributed_gloo(tmp_path, world, n, sharded):
    """Shard -> all-gather -> tile ranges -> gather of packed tiles, on CPU with gloo: equal and
    unequal shards (the in-place all-gather and the broadcast route), fewer rows than ranks."""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(n)
    profiles = rng.poisson(rng.uniform(0.5, 4, (n, 1)), (n, 64)).astype(np.int64)
    out = str(tmp_path / "matrix.npy")
    mp.spawn(_gloo_matrix_worker, args=(world, _free_port(), profiles, sharded, out), nprocs=world, join=True)
    got = np.load(out)
    for i in range(1, n):
        for j in range(i):
            want = ko.distance(profiles[i], profiles[j], do_scale=True, pairwise="sum")
            assert abs(got[i, j] - want) <= 1e-12 * abs(want), (i, j)
    assert np.array_equal(got, got.T) and not np.isnan(got).any()


def _gloo_by_record_worker(rank, world, port, text, k, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def count_rows(shard, kk, balance):      # stand-in for by_record_kernel (test only)
            rows = [ko.count_sequences([seq], kk) for _, seq in ko.parse_fasta(shard)]
            rows = [ko.balance(r) if balance else r for r in rows]
            return np.array(rows, dtype=np.int64).reshape(len(rows), 4 ** kk)

        first, names, rows = multigpu.count_by_record_distributed(text, k, balance=True, count_rows=count_rows)
        np.savez(out_path % rank, first=first, names=np.array(names, dtype=object), rows=rows, allow_pickle=True)
        first, names, rows = multigpu.count_by_record_distributed(text, k, gather=True, count_rows=count_rows)
        if rank == 0:
            np.savez(out_path % 99, first=first, names=np.array(names, dtype=object), rows=rows, allow_pickle=True)
        else:
            assert rows is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_by_record_distributed_gloo(tmp_path, world):
    """Records sharded over the ranks by byte range; the rows come back with the global
    index of the rank's first record, so record order is restored by index."""
    import torch.multiprocessing as mp
    text = make_fasta(12, 41)
    k = 4
    out = str(tmp_path / "rows_%d.npz")
    mp.spawn(_gloo_by_record_worker, args=(world, _free_port(), text, k, out), nprocs=world, join=True)
    records = ko.parse_fasta(text)
    want = np.array([ko.balance(ko.count_sequences([seq], k)) for _, seq in records])
    seen = 0
    for rank in range(world):
        part = np.load(out % rank, allow_pickle=True)
        first, rows = int(part["first"]), part["rows"]
        assert first == seen
        assert list(part["names"]) == [name for name, _ in records[first:first + len(rows)]]
        assert np.array_equal(rows, want[first:first + len(rows)])
        seen += len(rows)
    assert seen == len(records)
    whole = np.load(out % 99, allow_pickle=True)
    assert int(whole["first"]) == 0 and list(whole["names"]) == [name for name, _ in records]
    assert np.array_equal(whole["rows"], np.array([ko.count_sequences([seq], k) for _, seq in records]))


# ---------------------------------------------------------------------- GPU
def _nccl_worker(rank, world, port, text, k, profiles, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    _cabi.check(_cabi.load().kpal_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        b, e = multigpu.split_fasta(text, world)[rank]
        counts = multigpu.count_fasta_distributed(text[b:e], k, balance=True)        # peer-memory reduce
        counts_nccl = multigpu.count_fasta_distributed(text[b:e], k, balance=True, reduce='nccl')
        counts_slices = multigpu.count_fasta_distributed(text[b:e], k, balance=True, reduce='slices')
        matrix = multigpu.distance_matrix_distributed(profiles, do_scale=True)
        # unequal shards (301 rows): the broadcast route; this rank hands over its rows only
        rb, re_ = multigpu.shard_rows(len(profiles), rank, world)
        matrix_sharded = multigpu.distance_matrix_distributed(
            profiles[rb:re_], do_scale=True, pairwise='sum', sharded=True, n_total=len(profiles))
        first, names, rows = multigpu.count_by_record_distributed(text, 6, balance=True)
        np.savez(os.path.join(out_dir, "rows_%d.npz" % rank), first=first, rows=rows,
                 names=np.array(names, dtype=object), allow_pickle=True)
        if rank == 0:
            assert np.array_equal(counts, counts_nccl)
            assert np.array_equal(counts, counts_slices)      # balance + narrow reduce-scatter + shared host memory
            np.save(os.path.join(out_dir, "counts.npy"), counts)
            np.save(os.path.join(out_dir, "matrix.npy"), matrix)
            np.save(os.path.join(out_dir, "matrix_sharded.npy"), matrix_sharded)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_distributed_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from oracle import c_oracle
    text = make_fasta(21, 5000, max_len=1500)
    k = 11
    rng = np.random.default_rng(3)
    lam = np.exp(rng.uniform(np.log(0.5), np.log(8.0), 301))
    profiles = np.stack([rng.poisson(l, 4 ** 6) for l in lam]).astype(np.int64)
    mp.spawn(_nccl_worker, args=(world, _free_port(), text, k, profiles, str(tmp_path)),
             nprocs=world, join=True)
    counts = np.load(str(tmp_path / "counts.npy"))
    assert np.array_equal(counts, ko.balance(ko.count_fasta(text, k)))
    matrix = np.load(str(tmp_path / "matrix.npy"))
    want = c_oracle.distance_matrix(profiles, do_scale=True, threads=c_oracle.max_threads())
    low = np.tril_indices(len(profiles), -1)
    np.testing.assert_allclose(matrix[low], want[low], rtol=1e-9, atol=0)
    assert np.array_equal(matrix, matrix.T)
    matrix = np.load(str(tmp_path / "matrix_sharded.npy"))
    want = c_oracle.distance_matrix(profiles, do_scale=True, pairwise="sum", threads=c_oracle.max_threads())
    np.testing.assert_allclose(matrix[low], want[low], rtol=1e-9, atol=0)
    # per-record rows: every rank's rows, put back in record order by the global index
    records = ko.parse_fasta(text)
    seen = 0
    for rank in range(world):
        part = np.load(str(tmp_path / ("rows_%d.npz" % rank)), allow_pickle=True)
        first, rows = int(part["first"]), part["rows"]
        assert first == seen
        for i in range(0, len(rows), 97):
            assert part["names"][i] == records[first + i][0]
            assert np.array_equal(rows[i], ko.balance(ko.count_sequences([records[first + i][1]], 6)))
        seen += len(rows)
    assert seen == len(records)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3, 8])
def test_matrix_packed_tiles_virtual_world(world):
    """The multi-GPU matrix route on ONE GPU: the tile ranges of `world` virtual ranks are
    computed one after the other as packed tile arrays and scattered into one matrix, exactly
    as rank 0 does with the gathered arrays.  Against the C oracle."""
    import torch
    from oracle import c_oracle
    rng = np.random.default_rng(world)
    n, k = 333, 7
    lam = np.exp(rng.uniform(np.log(0.5), np.log(8.0), n))
    profiles = np.stack([rng.poisson(l, 4 ** k) for l in lam]).astype(np.int64)
    options = dict(metric='multiset', pairwise='prod', do_balance=True, do_scale=True, down=False)
    ops = multigpu._GpuMatrixOps(n, k, options, torch.device('cuda', 0))
    ops.prepare(profiles, 0)
    ops.make_order()
    n_tiles = ops.num_tiles()
    out = ops.new_out()
    for r in range(world):
        b, e = multigpu.tile_range(n_tiles, r, world)
        packed = ops.new_packed(e - b)
        ops.tiles_packed(b, e, packed)
        ops.unpack(packed, b, e, r == 0, out)
    got = ops.to_host(out)
    want = c_oracle.distance_matrix(profiles, do_balance=True, do_scale=True, threads=c_oracle.max_threads())
    low = np.tril_indices(n, -1)
    np.testing.assert_allclose(got[low], want[low], rtol=1e-9, atol=0)
    assert np.array_equal(got, got.T) and not got.diagonal().any()
    # the single-process call of the public function takes the same route
    one = multigpu.distance_matrix_distributed(profiles, **options)
    assert np.array_equal(one, got)


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [32, 64])
def test_peer_reduce_kernels_virtual_world(bits):
    """push / collect kernels of csrc/peer_reduce.cu with every "rank" on ONE device
    (plain pointers instead of IPC mappings): the root table is the exact sum of
    the per-rank tables for any world size, also when 4^k is not divisible by it."""
    import ctypes
    import torch
    L = _cabi.load()
    dev = torch.device("cuda", 0)
    dtype = torch.int32 if bits == 32 else torch.int64
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device=dev)
    gen.manual_seed(9)
    for k, world in ((2, 1), (2, 3), (3, 5), (6, 2), (6, 7), (9, 8), (10, 16), (11, 3)):
        bins = 4 ** k
        tables = [torch.randint(0, 2 ** 31 - 1, (bins,), dtype=dtype, device=dev, generator=gen)
                  for _ in range(world)]
        if bits == 64:
            tables = [t << 20 for t in tables]
        slot_bytes = int(L.kpal_peer_inbox_bytes(k, bits, world))
        assert slot_bytes >= bins * bits // 8
        inboxes = [torch.full((slot_bytes,), 0xAB, dtype=torch.uint8, device=dev) for _ in range(world)]
        ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in inboxes])
        root = torch.full((bins,), -1, dtype=dtype, device=dev)
        for r in range(world):
            _cabi.check(L.kpal_dev_reduce_push(tables[r].data_ptr(), bits, k, r, world, ptrs, sp))
        for r in range(world):
            _cabi.check(L.kpal_dev_reduce_collect(inboxes[r].data_ptr(), bits, k, r, world,
                                                  root.data_ptr(), sp))
        want = torch.stack(tables).sum(dim=0, dtype=dtype)          # wraps like the counters do
        assert torch.equal(root, want), (k, world)
    ptrs = (ctypes.c_void_p * 5)()
    assert L.kpal_dev_reduce_push(root.data_ptr(), bits, 2, 0, 5, ptrs, sp) == _cabi.KPAL_EINVAL
    assert L.kpal_dev_reduce_push(root.data_ptr(), 16, 6, 0, 2, ptrs, sp) == _cabi.KPAL_EINVAL
    assert L.kpal_dev_reduce_collect(root.data_ptr(), bits, 6, 2, 2, root.data_ptr(), sp) == _cabi.KPAL_EINVAL


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [32, 64])
def test_slice_push_collect_virtual_world(bits):
    """The fused form of the table sum (balance + narrow reduce-scatter + distributed finalize,
    csrc/peer_reduce.cu) with every "rank" on ONE device: the owners' int64 slices, put together,
    equal balance(sum of the per-rank tables) -- for narrow (<= 255) and wide counts, world sizes
    that do not divide 4^k, three epochs (both inbox parities), device slices and the narrow
    device->host copy of a slice."""
    import ctypes
    import torch
    L = _cabi.load()
    dev = torch.device("cuda", 0)
    dtype = torch.int32 if bits == 32 else torch.int64
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(bits)
    for k, world in ((6, 1), (6, 3), (7, 2), (9, 8), (10, 5), (11, 16), (12, 8)):
        bins = 4 ** k
        inbox_bytes = int(L.kpal_slice_inbox_bytes(k, world))
        inboxes = [torch.zeros(inbox_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
        ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in inboxes])
        begins = [int(L.kpal_slice_begin(k, r, world)) for r in range(world + 1)]
        assert begins[0] == 0 and begins[-1] == bins and all(b % 64 == 0 for b in begins)
        for epoch, high in ((1, 40), (2, 100_000), (3, 9)):
            host = [rng.integers(0, high, bins).astype(np.int64) for _ in range(world)]
            if epoch == 2:
                for t in host[1:]:
                    t[...] = rng.integers(0, 30, bins)          # only rank 0 sends wide rows
            tables = [torch.from_numpy(t).to(dev).to(dtype) for t in host]
            # rows as bytes with escapes; the last rank of the last epoch sends u32 rows
            wide_rows = [1 if (epoch == 3 and r == world - 1) else 0 for r in range(world)]
            for r in range(world):
                _cabi.check(L.kpal_dev_slice_push(tables[r].data_ptr(), bits, k, r, world, ptrs, epoch, wide_rows[r], sp))
                if r:           # rank 0's signal is sent by its collect kernel below
                    _cabi.check(L.kpal_dev_slice_signal(k, r, world, ptrs, epoch, wide_rows[r], sp))
            want = ko.balance(np.sum(host, axis=0))
            for r in range(world):
                n = begins[r + 1] - begins[r]
                piece = torch.full((max(n, 1),), -1, dtype=torch.int64, device=dev)
                _cabi.check(L.kpal_dev_slice_collect(ptrs, k, r, world, epoch, wide_rows[0] if r == 0 else -1, piece.data_ptr(), sp))
                assert np.array_equal(piece[:n].cpu().numpy(), want[begins[r]:begins[r + 1]]), (k, world, epoch, r)
                out = np.full(max(n, 1), -1, dtype=np.int64)
                _cabi.check(L.kpal_dev_slice_collect_to_host(ptrs, k, r, world, epoch, -1, _cabi.ptr(out), sp))
                assert np.array_equal(out[:n], want[begins[r]:begins[r + 1]]), (k, world, epoch, r, "host")
    ptrs = (ctypes.c_void_p * 5)()
    assert L.kpal_dev_slice_push(tables[0].data_ptr(), bits, 5, 0, 2, ptrs, 1, 0, sp) == _cabi.KPAL_EINVAL     # k < 6
    assert L.kpal_dev_slice_push(tables[0].data_ptr(), bits, 8, 0, 2, ptrs, 0, 0, sp) == _cabi.KPAL_EINVAL     # epoch 0


@pytest.mark.gpu
def test_count_push_fused_virtual_world():
    """kpal_dev_count_packed_push with every "rank" on one device: the radix count's
    second pass stores its histograms straight into the owners' inboxes (fused = 1
    when the world divides the bucket count), otherwise count + push kernel; either
    way collect() yields the exact sum of the per-rank counts."""
    import ctypes
    import torch
    from oracle import c_oracle
    L = _cabi.load()
    dev = torch.device("cuda", 0)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(17)
    letters = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    try:
        for k, world, path, want_fused in ((9, 2, 2, 1), (10, 8, 2, 1), (11, 4, 2, 1), (10, 3, 2, 0),
                                           (9, 4, 1, 0), (6, 2, 0, 0), (12, 16, 2, 1)):
            _cabi.check(L.kpal_set_option(b"count_path", path))
            bins = 4 ** k
            slot_bytes = int(L.kpal_peer_inbox_bytes(k, 32, world))
            inboxes = [torch.full((slot_bytes,), 0xCD, dtype=torch.uint8, device=dev) for _ in range(world)]
            ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in inboxes])
            want = np.zeros(bins, dtype=np.int64)
            keep = []
            for r in range(world):
                n = int(rng.integers(150_000, 400_000))
                seq = letters[rng.choice(9, n, p=[.24, .24, .24, .24, .01, .01, .005, .005, .01])]
                seq[rng.integers(0, n, 40)] = 10                       # record breaks
                records = bytes(seq).split(b"\n")
                want += c_oracle.count_sequences([x.decode() for x in records], k)
                codes, valid, _, n_bases = _cabi.pack_sequences(records)
                d_codes = torch.from_numpy(codes.view(np.int32)).to(dev)
                d_valid = torch.from_numpy(valid.view(np.int32)).to(dev)
                table = torch.zeros(bins, dtype=torch.int32, device=dev)
                fused = ctypes.c_int(-1)
                _cabi.check(L.kpal_dev_count_packed_push(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                                         table.data_ptr(), 32, r, world, ptrs, sp,
                                                         ctypes.byref(fused)))
                assert fused.value == want_fused, (k, world, path)
                keep.append((d_codes, d_valid, table))
            root = torch.full((bins,), -1, dtype=torch.int32, device=dev)
            for r in range(world):
                _cabi.check(L.kpal_dev_reduce_collect(inboxes[r].data_ptr(), 32, k, r, world,
                                                      root.data_ptr(), sp))
            assert np.array_equal(root.cpu().numpy().astype(np.int64), want), (k, world, path)
    finally:
        _cabi.check(L.kpal_set_option(b"count_path", 0))
