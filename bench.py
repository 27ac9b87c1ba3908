#!/usr/bin/env python
"""
bench.py -- headline benchmark of the kPAL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|count|matrix]
                    [--impl ours|reference]

BASELINE.json's metric has two halves and the default run (`--workload all`) measures both
and prints ONE JSON line: the count record (configs[1], the one `metric` is quoted on) with
the matrix record (configs[3]) as its "matrix" sub-record.

Count workload: `kpal count` at k=12 over 100 Mbp of synthetic 150-bp reads, with balance.
A step = one pass of the hot path over the whole batch:

  value : packed sequence already resident in HBM -> int64 balanced profile in
          HBM (memset + count kernels + [table sum over the ranks] + widen/balance
          kernel), timed with CUDA events on the launching stream, L2 flushed
          between steps (untimed), max over ranks;
  e2e   : the same job through the host-buffer C ABI (kpal_count_fasta): pinned
          FASTA bytes -> hybrid upload (the head of the text raw, scanned / packed
          by the GPU; the tail packed by idle host threads, 0.375 B/base over the
          bus) -> count kernels -> [table sum] -> narrow (uint8 / uint16) D2H of
          the profile, widened to int64 by host threads; wall clock around the
          call with device syncs.  h2d_bytes_per_step = the bytes that crossed
          the bus (kpal_last_upload), text_bytes_per_step = the FASTA bytes.

Multi-GPU (weak scaling): every rank counts its own shard of records (same
size per rank); `--reduce auto` = `slices`: every rank balances its table and
stores it, one byte per bin, into the slice owners' inboxes over NVLink peer
memory, every owner sums its slice (multigpu.SliceReducer; no NCCL call on
the data path), the profile stays sharded by slice.  `--reduce fused|peer|nccl`:
the round-1 forms (u32 tables summed onto rank 0).  `--config 5`: one GPU's
shard of BASELINE configs[4] (k=13 genome-like records).

Matrix workload: BASELINE.json configs[3], the 4096-profile k=10 scaled multiset distance
matrix (profile-pairs/s), with its own small step count (<= 2).  At N > 1 (strong scaling)
every rank generates / uploads and prepares 1/N of the profiles, the prepared set is
all-gathered over NVLink, the upper-triangle tiles are dealt out in equal ranges and the
finished tiles are gathered on rank 0 as compact tile arrays (kpal_b200/multigpu.py).

`--config 1`: BASELINE configs[0] (k=6, one 1 Mbp record, no balance); `--config 5`: one
GPU's shard of configs[4] (k=13; the full 3 Gbp job at --gpus 8).

`--impl reference`: the CPU baseline -- the oracle's C port of the reference
algorithm on all host threads (`value`), plus `reference_python`: the unmodified
reference's own pure-Python loop on one core and a bounded sample, when its
offline install (baseline/_ref) is present.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_COUNT = 12
N_READS = 666_667
READ_LEN = 150
K_MATRIX = 10
N_PROFILES = 4096
FLUSH_BYTES = 512 << 20


# ----------------------------------------------------------------- synthetic
def synthetic_reads(seed, n_reads=N_READS, read_len=READ_LEN):
    """SURVEY.md section 8d, cfg 2: uniform ACGT, 0.1 % N, 5 % lower case."""
    rng = np.random.default_rng(seed)
    n = n_reads * read_len
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
    bases = np.where(rng.random(n, dtype=np.float32) < 0.05, bases + 32, bases).astype(np.uint8)
    bases[rng.random(n, dtype=np.float32) < 0.001] = ord("N")
    return bases.reshape(n_reads, read_len)


def reads_to_fasta(reads, wrap=70):
    """'>rNNNNNNN' headers, sequence wrapped at 70 columns, '\\n' line ends."""
    n_reads, read_len = reads.shape
    header = np.frombuffer(b">r0000000\n", dtype=np.uint8)
    n_lines = (read_len + wrap - 1) // wrap
    width = len(header) + read_len + n_lines
    out = np.empty((n_reads, width), dtype=np.uint8)
    out[:, :len(header)] = header
    idx = np.arange(n_reads)
    for d in range(7):
        out[:, 8 - d] = ord("0") + (idx // 10 ** d) % 10
    col = len(header)
    for line in range(n_lines):
        seg = reads[:, line * wrap:(line + 1) * wrap]
        out[:, col:col + seg.shape[1]] = seg
        col += seg.shape[1]
        out[:, col] = ord("\n")
        col += 1
    return out.reshape(-1)


def synthetic_chromosomes(seed, n_records, record_len):
    """SURVEY.md section 8d, cfg 5: long records with 2-6 blocks of N (10 k - 1 M bases) and
    ~40 % soft-masked (lower case) stretches.  Returns a list of uint8 arrays."""
    rng = np.random.default_rng(seed)
    records = []
    for _ in range(n_records):
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, record_len, dtype=np.uint8)]
        # soft-masked stretches: alternate segments of ~2-40 kb, 40 % of them lower case
        n_seg = max(2, record_len // 20_000)
        cuts = np.sort(rng.integers(0, record_len, n_seg - 1))
        masked = rng.random(n_seg) < 0.4
        lower = np.repeat(masked, np.diff(np.concatenate(([0], cuts, [record_len]))))
        seq = np.where(lower, seq + 32, seq).astype(np.uint8)
        for _ in range(int(rng.integers(2, 7))):
            size = int(min(rng.integers(10_000, 1_000_001), max(1, record_len // 8)))
            at = int(rng.integers(0, max(1, record_len - size)))
            seq[at:at + size] = ord("N")
        records.append(seq)
    return records


def records_to_fasta(records, wrap=70, first=0):
    """Vectorised 70-column FASTA of a few long records ('>chrNN' headers)."""
    parts = []
    for i, seq in enumerate(records):
        parts.append(np.frombuffer((">chr%02d\n" % (first + i)).encode(), dtype=np.uint8))
        full = len(seq) // wrap
        body = np.empty((full, wrap + 1), dtype=np.uint8)
        body[:, :wrap] = seq[:full * wrap].reshape(full, wrap)
        body[:, wrap] = ord("\n")
        parts.append(body.reshape(-1))
        if len(seq) > full * wrap:
            parts.append(seq[full * wrap:])
            parts.append(np.frombuffer(b"\n", dtype=np.uint8))
    return np.concatenate(parts)


# -------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region
    (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak(key, fallback):
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)[key]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return fallback, "fallback (B200_PROFILING.md)"


def measured_fp64_peak():
    """FP64 FMA peak in TFLOP/s: MEASURED_PEAKS.json has no fp64 entry, so the denominator is
    this repo's own DFMA microbenchmark on a B200 of this pool (kpal_b200/csrc/microbench.cu,
    result committed in profiles/r01_microbench.jsonl), else the nominal figure."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_microbench.jsonl")) as f:
            for line in f:
                rec = json.loads(line)
                if rec.get("bench") == "dfma":
                    return float(rec["tflops"]), "measured DFMA microbenchmark (profiles/r01_microbench.jsonl)"
    except Exception:
        pass
    return 2 * 64 * 148 * 1.965e9 / 1e12, "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz"


def host_mem_available():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except Exception:
        pass
    return None


def host_threads():
    """Threads for the CPU legs: every core this process may run on.  (torchrun exports
    OMP_NUM_THREADS=1 to its workers; the C port takes its thread count as an argument, so
    the reference arm and the oracle checks still use the whole host.)"""
    try:
        return max(1, len(_FULL_AFFINITY if _FULL_AFFINITY is not None else os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def traffic_record(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture, with the
    file it came from (the bench does not run under a profiler)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            table = json.load(f)
        return table.get(kernel), table.get("_source", {}).get(kernel)
    except Exception:
        return None, None


# ------------------------------------------------------- count workload inputs
def count_shard_input(config, rank, k, mbp_per_gpu, composition="uniform"):
    """(records as the oracle reads them [bytes, one record per line], FASTA bytes, sequence
    bases, windows or None, description) of rank `rank`'s shard -- deterministic per rank, so
    rank 0 can rebuild every shard for the oracle check."""
    if config == 5:
        rec_len = int(mbp_per_gpu * 1e6) // 3
        records = synthetic_chromosomes(5000 + rank, 3, rec_len)
        fasta = records_to_fasta(records, first=3 * rank)
        oracle_text = np.concatenate([np.append(r, np.uint8(10)) for r in records])
        bases = sum(len(r) for r in records)
        what = ("kpal count k=%d, %d x %.1f Mbp records per GPU with N blocks and soft-masking, balance; "
                "BASELINE configs[4]" % (k, 3, rec_len / 1e6))
        return oracle_text, fasta, bases, None, what
    if config == 1:
        rng = np.random.default_rng(1 + rank)
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1_000_000, dtype=np.uint8)]
        fasta = records_to_fasta([seq], first=rank)
        oracle_text = np.append(seq, np.uint8(10))
        return (oracle_text, fasta, seq.size, seq.size - (k - 1),
                "kpal count k=%d, one 1 Mbp record, no balance; BASELINE configs[0]" % k)
    reads = synthetic_reads(1000 + rank)
    if composition == "skewed":
        # 10 % of the reads are low-complexity: homopolymer runs and short tandem repeats
        rng = np.random.default_rng(77 + rank)
        low = np.flatnonzero(rng.random(N_READS) < 0.10)
        units = [b"A", b"T", b"AC", b"AT", b"CAG", b"TTAGGG", b"AAAAT"]
        for u in range(len(units)):
            unit = np.frombuffer(units[u], dtype=np.uint8)
            rows = low[u::len(units)]
            reads[rows] = np.resize(unit, READ_LEN)
    fasta = reads_to_fasta(reads)
    oracle_text = np.insert(reads, READ_LEN, ord("\n"), axis=1).reshape(-1)
    what = ("kpal count k=%d, 100 Mbp of 150-bp reads (666667 records/GPU)%s, balance; BASELINE configs[1]"
            % (k, ", 10 %% low-complexity reads" if composition == "skewed" else ""))
    return oracle_text, fasta, reads.size, reads.size - N_READS * (k - 1) if composition == "uniform" else None, what


# ----------------------------------------------------------- reference arm
def run_reference(args):
    """CPU baseline: the oracle's C port on all host threads, for both halves of the metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    threads = host_threads()

    def timed(step):
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        return (time.perf_counter() - t0) / args.steps

    def matrix_leg():
        n = 48
        rng = np.random.default_rng(4)
        lam = np.exp(rng.uniform(np.log(0.5), np.log(8.0), n))
        profiles = np.stack([rng.poisson(l, 4 ** K_MATRIX) for l in lam]).astype(np.int64)
        dt = timed(lambda: c_oracle.distance_matrix(profiles, do_scale=True, threads=threads))
        sample = "leading %d profiles (%d pairs) of the 4096-profile k=10 set (C port, OpenMP)" % (
            n, n * (n - 1) // 2)
        value = n * (n - 1) / 2 / dt
        return {"impl": "reference", "metric": "profile_pairs_per_sec_k10_multiset", "value": value,
                "unit": "profile-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "kpal matrix multiset/prod scaled, 4096 profiles k=10", "k": K_MATRIX},
                "cpu_baseline": {"value": value, "unit": "profile-pairs/s", "cores": threads, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "profile-pairs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}

    def count_leg():
        k = args.k or (13 if args.config == 5 else 6 if args.config == 1 else K_COUNT)
        balance = args.config != 1
        mbp = min(args.mbp_per_gpu, 90.0) if args.config == 5 else args.mbp_per_gpu   # bounded sample
        text, _, n_bases, _, what = count_shard_input(args.config, 0, k, mbp, args.composition)

        def step():
            counts = c_oracle.count_bytes(text, k, threads=threads)
            return c_oracle.balance(counts) if balance else counts      # literal klib.py:285-298 loop in C
        dt = timed(step)
        value = n_bases / 1e9 / dt
        sample = ("full workload of one GPU: %d bases, k=%d%s (C port, OpenMP)"
                  % (n_bases, k, ", balance" if balance else ""))
        if args.config == 5 and mbp < args.mbp_per_gpu:
            sample = "bounded sample: 3 x %.0f Mbp records of the same generator, k=%d, balance (C port, OpenMP)" % (
                mbp / 3, k)
        return {"impl": "reference", "metric": "gbases_per_sec_counted_k%d" % k, "value": value,
                "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int64", "data": "synthetic", "config": {"workload": what, "k": k},
                "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": threads, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}

    def python_leg():
        """The UNMODIFIED reference's own pure-Python counting loop (kpal/klib.py:135-170 + balance,
        285-298), when its sources travelled with the repo (baseline/_ref: `pip install --no-deps
        --target baseline/_ref` of the reference, see __graft_entry__.build): one core, a bounded
        sample of the same reads, and its counts compared with the C port's on that sample."""
        try:
            from oracle import ref_loader
            if not ref_loader.available():
                return {"unavailable": "baseline/_ref (offline install of the reference) is not present"}
            klib = ref_loader.load()[0]
            k = args.k or K_COUNT
            reads = synthetic_reads(1000)[:20000]                       # 3 Mbp of the config-2 reads
            sequences = [row.tobytes().decode("ascii") for row in reads]
            t0 = time.perf_counter()
            profile = klib.Profile.from_sequences(sequences, k)
            t1 = time.perf_counter()
            profile.balance()                                           # a Python loop over the 4^k bins
            t2 = time.perf_counter()
            port = c_oracle.balance(c_oracle.count_bytes(np.insert(reads, READ_LEN, ord("\n"), axis=1).reshape(-1), k,
                                                         threads=threads))
            return {"value": reads.size / 1e9 / (t1 - t0), "unit": "Gbases/s", "cores": 1,
                    "count_seconds": t1 - t0, "balance_seconds": t2 - t1,
                    "sample": "first 20000 reads (3 Mbp) of the config-2 input, k=%d: kpal.klib.Profile.from_sequences "
                              "of the unmodified reference (value = its counting loop alone), then Profile.balance "
                              "(independent of the input size)" % k,
                    "counts_equal_c_port": bool(np.array_equal(np.asarray(profile.counts), port))}
        except Exception as error:          # never fail the arm over the optional leg
            return {"unavailable": "%s: %s" % (type(error).__name__, error)}

    if args.workload == "matrix":
        out = matrix_leg()
    else:
        out = count_leg()
        if args.workload == "all":
            out["matrix"] = matrix_leg()
        if args.config == 2 and args.composition == "uniform":
            out["reference_python"] = python_leg()
    print(json.dumps(out))


# ------------------------------------------------------------------ our arm
def init_dist(args):
    # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION/INFO level
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


_FULL_AFFINITY = None


def numa_bind(args, world, local):
    """N > 1, one process per GPU: keep the rank's pinned buffers and host threads next to its GPU."""
    global _FULL_AFFINITY
    if world > 1 and args.numa_bind:
        from kpal_b200 import multigpu
        previous = multigpu.bind_to_gpu_cpus(local)
        if _FULL_AFFINITY is None:
            _FULL_AFFINITY = previous


def whole_host():
    """Undo init_dist's binding (for the untimed oracle checks on rank 0)."""
    if _FULL_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, _FULL_AFFINITY)
        except OSError:
            pass


def barrier_sync(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_count(args, world, rank, local):
    import torch
    import torch.distributed as dist
    from kpal_b200 import _cabi

    numa_bind(args, world, local)
    L = _cabi.load()
    _cabi.check(L.kpal_set_device(local))
    _cabi.check(L.kpal_set_option(b"count_path", args.count_path))
    _cabi.check(L.kpal_set_option(b"radix_payload_bits", args.radix_payload_bits))
    _cabi.check(L.kpal_set_option(b"pair_flush_every", args.pair_flush_every))
    _cabi.check(L.kpal_set_option(b"pair_fused", args.pair_fused))
    if args.radix_max_buckets:
        _cabi.check(L.kpal_set_option(b"radix_max_buckets", args.radix_max_buckets))
    _cabi.check(L.kpal_set_option(b"radix_debug", args.radix_debug))
    _cabi.check(L.kpal_set_option(b"radix_shape", args.radix_shape))
    _cabi.check(L.kpal_set_option(b"fasta_chunks", args.fasta_chunks))
    _cabi.check(L.kpal_set_option(b"narrow_d2h", args.narrow_d2h))
    _cabi.check(L.kpal_set_option(b"dma_share", args.dma_share))
    _cabi.check(L.kpal_set_option(b"fasta_split", args.fasta_split))
    _cabi.check(L.kpal_set_option(b"fasta_hybrid", args.fasta_hybrid))
    _cabi.check(L.kpal_set_option(b"fasta_hybrid_share", args.fasta_hybrid_share))
    dev = torch.device("cuda", local)

    # ---- this rank's shard of records (same size on every rank: weak scaling)
    k = args.k or (13 if args.config == 5 else 6 if args.config == 1 else K_COUNT)
    balance = 0 if args.config == 1 else 1
    oracle_text, fasta_np, seq_bases, n_windows, workload = count_shard_input(
        args.config, rank, k, args.mbp_per_gpu, args.composition)
    if world > 1 or rank != 0:
        oracle_text = None              # rank 0 rebuilds every shard for the check at the end
    bins = 4 ** k
    n_fasta = fasta_np.size
    pinned_fasta = _cabi.PinnedArray(n_fasta, np.uint8)
    pinned_fasta.array[:] = fasta_np
    pinned_out = _cabi.PinnedArray(bins, np.int64)
    codes, valid, _, _, n_bases = _cabi.fasta_pack(fasta_np.tobytes())
    del fasta_np
    d_codes = torch.from_numpy(codes.view(np.int32)).to(dev)
    d_valid = torch.from_numpy(valid.view(np.int32)).to(dev)
    d_table = torch.zeros(bins, dtype=torch.int32, device=dev)
    d_counts = torch.zeros(bins, dtype=torch.int64, device=dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    reducer = None
    slicer, shared, d_slice = None, None, None
    reduce_note = None
    if args.reduce == "auto":
        # balance + narrow reduce-scatter + distributed finalize over NVLink peer memory
        # (multigpu.SliceReducer) whenever the step balances; else the NCCL reduce
        args.reduce = "slices" if (balance and k >= 6) else "nccl"
    if world > 1 and args.reduce == "slices":
        from kpal_b200 import multigpu
        peer_ok = torch.ones(1, dtype=torch.int32, device=dev)
        for a in range(world):
            if a != local and not torch.cuda.can_device_access_peer(local, a):
                peer_ok.zero_()
        dist.all_reduce(peer_ok, op=dist.ReduceOp.MIN)
        if int(peer_ok.item()):
            try:
                slicer = multigpu.SliceReducer(k)
            except Exception as exc:        # all ranks fail together (collective decision inside)
                slicer = None
                reduce_note = "sliced peer reduce unavailable (%s): NCCL reduce used" % (exc,)
        else:
            reduce_note = "no peer access between all GPU pairs: NCCL reduce used"
        if slicer is not None:
            shared = multigpu.SharedProfile(k)
            sb, se = slicer.slice_range()
            d_slice = torch.zeros(max(se - sb, 1), dtype=torch.int64, device=dev)
        else:
            args.reduce = "nccl"
    if world > 1 and args.reduce in ("peer", "fused"):
        from kpal_b200 import multigpu
        # CUDA IPC between the ranks can be refused by the box (container without a shared
        # PID/IPC namespace, GPUs without peer access): every rank then takes the NCCL reduce,
        # and the JSON line says so.  The decision is collective so no rank is left waiting.
        peer_ok = torch.ones(1, dtype=torch.int32, device=dev)
        for a in range(world):
            if a != local and not torch.cuda.can_device_access_peer(local, a):
                peer_ok.zero_()
        dist.all_reduce(peer_ok, op=dist.ReduceOp.MIN)
        if int(peer_ok.item()):
            try:
                reducer = multigpu.PeerReducer(k, 32)
            except Exception as exc:        # ranks fail together (IPC open) or not at all
                reducer = None
                reduce_note = "peer-memory reduce unavailable (%s): NCCL reduce used" % (exc,)
        else:
            reduce_note = "no peer access between all GPU pairs: NCCL reduce used"
        ok = torch.tensor([1 if reducer is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()) and reducer is not None:
            reducer.close()
            reducer = None
            reduce_note = "peer-memory reduce unavailable on another rank: NCCL reduce used"

    def reduce_tables():
        """Sum of the per-rank u32 tables onto rank 0; returns its device pointer there."""
        summed = d_table.data_ptr()
        if reducer is not None:
            summed = reducer.reduce(d_table.data_ptr(), sp)       # peer-memory all-to-all + collect
        elif world > 1:
            dist.reduce(d_table, dst=0, op=dist.ReduceOp.SUM)
        return summed

    def reduce_and_finalize():
        """... + widen/balance there."""
        summed = reduce_tables()
        if rank == 0:
            _cabi.check(L.kpal_dev_finalize_counts(summed, 32, k, balance, d_counts.data_ptr(), sp))

    def device_step(ev=None):
        if ev:
            ev[0].record(stream)
        if slicer is not None:
            # count, then balance + narrow push into the owners' inboxes, then this rank's slice
            _cabi.check(L.kpal_dev_count_packed_fresh(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                                      d_table.data_ptr(), 32, sp))
            if ev:
                ev[1].record(stream)
            slicer.push(d_table.data_ptr(), 32, sp, n_bases=n_bases)
            if ev:
                ev[2].record(stream)
            slicer.collect(d_slice.data_ptr(), sp)
            return
        if reducer is not None and args.reduce == "fused":
            d_table.zero_()
            # count + all-to-all in one call: pass 2 of the radix count stores into the inboxes
            summed = reducer.count_and_reduce(d_codes.data_ptr(), d_valid.data_ptr(), n_bases,
                                              d_table.data_ptr(), sp)
            if ev:
                ev[1].record(stream)
            if rank == 0:
                _cabi.check(L.kpal_dev_finalize_counts(summed, 32, k, balance, d_counts.data_ptr(), sp))
            return
        # the table is zeroed by the call (a memset on the stream, inside the timed step)
        if args.fresh:
            _cabi.check(L.kpal_dev_count_packed_fresh(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                                      d_table.data_ptr(), 32, sp))
        else:           # A/B: memset + count into the zeroed table
            d_table.zero_()
            _cabi.check(L.kpal_dev_count_packed(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                                d_table.data_ptr(), 32, sp))
        if ev:
            ev[1].record(stream)
        reduce_and_finalize()

    def e2e_step():
        if world == 1:
            _cabi.check(L.kpal_count_fasta(pinned_fasta._ptr, n_fasta, k, balance, pinned_out._ptr))
        elif slicer is not None:
            table, bits = ctypes.c_void_p(), ctypes.c_int()
            _cabi.check(L.kpal_count_fasta_dev_table(pinned_fasta._ptr, n_fasta, k, ctypes.byref(table),
                                                     ctypes.byref(bits), sp))
            slicer.push(table, bits.value, sp, n_bases=n_bases)
            slicer.collect_to_host(shared.array[sb:se], sp)      # this rank's slice, into shared host memory
        else:
            d_table.zero_()
            nb = ctypes.c_uint64()
            _cabi.check(L.kpal_count_fasta_to_dev(pinned_fasta._ptr, n_fasta, k, d_table.data_ptr(),
                                                  32, sp, ctypes.byref(nb)))
            summed = reduce_tables()
            if rank == 0:       # widen + balance + narrow D2H into the host profile
                _cabi.check(L.kpal_dev_table_to_host(summed, 32, k, balance, pinned_out._ptr, sp))
        torch.cuda.synchronize()

    # ---- warm-up (>= 3)
    for _ in range(max(args.warmup, 3)):
        device_step()
        flush.zero_()
    barrier_sync(world)

    # ---- timed: exactly K steps, CUDA events per step, L2 flushed (untimed) between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.kpal_reset_kernel_launches()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
    barrier_sync(world)
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        step_ev[i][0].record(stream)
        device_step(kern_ev[i])
        step_ev[i][1].record(stream)
    barrier_sync(world)
    wall = time.perf_counter() - wall0
    launches = int(L.kpal_kernel_launches())
    step_ms = sum(a.elapsed_time(b) for a, b in step_ev) / args.steps
    kern_ms = sum(e[0].elapsed_time(e[1]) for e in kern_ev) / args.steps
    tail_ms = None
    if slicer is not None:      # this rank's view: count | balance + push | wait for the peers + collect
        tail_ms = {"push_ms": sum(e[1].elapsed_time(e[2]) for e in kern_ev) / args.steps,
                   "wait_collect_ms": sum(e[2].elapsed_time(s[1]) for e, s in zip(kern_ev, step_ev)) / args.steps}
    step_ms = max_over_ranks(step_ms, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C ABI
    for _ in range(2):
        e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    h2d_sum, host_text_sum = 0, 0
    up, packed_text = ctypes.c_uint64(), ctypes.c_uint64()
    for _ in range(args.steps):
        e2e_step()
        L.kpal_last_upload(ctypes.byref(up), ctypes.byref(packed_text))     # (two loads; this rank's call)
        h2d_sum += up.value
        host_text_sum += packed_text.value
    barrier_sync(world)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps, world)
    h2d_mean, host_text_mean = h2d_sum / args.steps, host_text_sum / args.steps

    # ---- parity of what was just measured, against the ORACLE: the C port counts the
    # concatenation of every rank's shard (rank 0 rebuilds the shards of the other ranks from
    # their seeds; untimed).  At N > 1 the sum of the per-rank tables taken with a plain NCCL
    # reduce is checked as well.
    result_ok, cpu, parity = None, None, None
    gathered = None
    if world > 1 and slicer is not None:
        device_step()
        cap = max(slicer.slice_range(r)[1] - slicer.slice_range(r)[0] for r in range(world))
        mine = torch.zeros(cap, dtype=torch.int64, device=dev)
        mine[:se - sb].copy_(d_slice[:se - sb])
        parts = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            gathered = torch.cat([parts[r][:slicer.slice_range(r)[1] - slicer.slice_range(r)[0]]
                                  for r in range(world)]).cpu().numpy()
    if world > 1:
        device_step()                       # the result under test: rank 0's d_counts / pinned_out
        check = d_table.clone()
        d_table.zero_()
        _cabi.check(L.kpal_dev_count_packed(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                            d_table.data_ptr(), 32, sp))
        check.copy_(d_table)
        dist.reduce(check, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        from oracle import c_oracle
        whole_host()
        threads = host_threads()
        want = np.zeros(bins, dtype=np.int64)
        cpu_s = 0.0
        for r in range(world):
            text = oracle_text if world == 1 else count_shard_input(
                args.config, r, k, args.mbp_per_gpu, args.composition)[0]
            t0 = time.perf_counter()
            part = c_oracle.count_bytes(text, k, threads=threads)
            if world == 1 and balance:
                part = c_oracle.balance(part)             # literal klib.py:285-298 loop in C
            cpu_s += time.perf_counter() - t0
            want += part
            del text
        if world > 1 and balance:
            want = c_oracle.balance(want)
        if n_windows is None:
            n_windows = int(want.sum()) // (2 if balance else 1) // world
        dev_ok = bool(np.array_equal(gathered if gathered is not None else d_counts.cpu().numpy(), want))
        host_ok = bool(np.array_equal(shared.array if shared is not None else pinned_out.array, want))
        parity = {"checked_against": "C port of klib.py:149-170 + 285-298 on the concatenation of all %d shard(s)" % world,
                  "device_result_equals_oracle": dev_ok, "host_result_equals_oracle": host_ok,
                  "total_windows": int(want.sum()) // (2 if balance else 1)}
        result_ok = dev_ok and host_ok
        if world > 1:
            want_dev = torch.empty_like(d_counts)
            _cabi.check(L.kpal_dev_finalize_counts(check.data_ptr(), 32, k, balance, want_dev.data_ptr(), sp))
            torch.cuda.synchronize()
            parity["device_result_equals_nccl_sum"] = bool(
                np.array_equal(want_dev.cpu().numpy(), gathered) if gathered is not None else torch.equal(want_dev, d_counts))
            result_ok = result_ok and parity["device_result_equals_nccl_sum"]
        else:
            cpu = {"value": seq_bases / 1e9 / cpu_s, "unit": "Gbases/s", "cores": threads, "kind": "port",
                   "sample": "full workload once (C port of klib.py:149-170 + balance, OpenMP); the "
                             "reference's pure-Python loop measured 0.001-0.0036 Gbases/s on 1 core "
                             "(BASELINE.md)"}

    out = None
    if rank == 0:
        total_bases = seq_bases * world
        peak, peak_src = measured_peak("hbm_gbs", 6650.0)
        # packed stream read once + the int64 table written once (the balance
        # is fused, + 0).  The whole step is charged: memset + count kernels + widen/balance.
        alg_bytes = 0.375 * n_bases + 8 * bins
        radix = args.count_path >= 2 or (args.count_path == 0 and k >= 9 and
                                         n_bases >= ((16 << 20) if k <= 12 else (4 << 20)))
        pairs = radix and k <= 12 and args.count_path != 3
        narrow = bool(args.narrow_d2h and bins >= (1 << 20))
        # bytes of the narrow copy: uint8 when every count of the result fits, else uint16
        # (what finalize_to_host decides from the device's flag words), + the 8 flag bytes
        # -- for the bins below the split; the last dma_share/16 of the (pinned) profile is
        # copied as int64 by the DMA engine
        split = bins // 16 * (16 - args.dma_share)
        host_result = (shared.array if shared is not None else pinned_out.array)[:split]
        n_over8, n_over16 = int((host_result > 255).sum()), int((host_result > 65535).sum())
        cap8, cap16 = max(4096, bins // 32), max(4096, bins // 256)       # side lists (cabi.cu finalize_to_host)
        listed = 0
        if not narrow:
            narrow_width = 8
        elif args.narrow_d2h == 1 and n_over8 <= cap8:
            narrow_width, listed = 1, n_over8
        elif n_over16 <= cap16:
            narrow_width, listed = 2, n_over16
        else:
            narrow_width = 8
        d2h_bytes = (bins * 8 if narrow_width == 8 else
                     split * narrow_width + (bins - split) * 8 + 16 + listed * 16)
        kernels = (["pair_partition_kernel", "pair_histogram_kernel"] if pairs else
                   ["radix_partition_kernel", "radix_histogram_kernel"] if radix else
                   ["count_smem_kernel" if k <= 7 else "count_global_kernel"])
        step_kernels = ["memset(table)"] + kernels + [
            "finalize_balance_tiled_kernel" if (balance and k >= 6) else "finalize_kernel"]
        traffic, traffic_src = None, None
        if args.config == 2 and k == K_COUNT and args.composition == "uniform":
            parts = [traffic_record(name) for name in step_kernels]
            if all(p[0] is not None for p in parts):
                traffic = int(sum(p[0] for p in parts))
                traffic_src = sorted(set(p[1] for p in parts if p[1]))
        achieved = alg_bytes / (step_ms * 1e-3) / 1e9 if world == 1 else alg_bytes / (kern_ms * 1e-3) / 1e9
        out = {
            "metric": "gbases_per_sec_counted_k%d" % k, "value": total_bases / 1e9 / (step_ms * 1e-3),
            "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 counters -> int64", "data": "synthetic",
            "config": {"workload": workload,
                       "k": k, "bases_per_gpu": int(seq_bases), "packed_bases_per_gpu": int(n_bases),
                       "l2": "512 MB memset between steps (untimed); the table memset is inside the step",
                       "parallelism": ("records sharded per GPU; every rank balances its table and stores it, 1 byte per "
                                       "bin, into the slice owners' inboxes over NVLink peer memory; every rank sums "
                                       "the rows it received into its int64 slice of the profile (result sharded by "
                                       "slice over the GPUs; e2e: the slices meet in shared host memory)"
                                       if slicer is not None else
                                       "records sharded per GPU, u32 tables summed onto rank 0 " +
                                       ("over NVLink peer memory (all-to-all fused into the count's histogram pass + collect kernel)"
                                        if (reducer is not None and args.reduce == "fused") else
                                        "over NVLink peer memory (all-to-all push + collect kernels)"
                                        if reducer is not None else "with an NCCL reduce"))
                       if world > 1 else "1 GPU",
                       **({"reduce_note": reduce_note} if reduce_note else {})},
            "e2e": {"value": total_bases / 1e9 / e2e_s, "unit": "Gbases/s",
                    "h2d_bytes_per_step": int(h2d_mean * world),
                    "text_bytes_per_step": int(n_fasta * world),
                    "host_packed_text_frac": round(host_text_mean / n_fasta, 4),
                    "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": e2e_s * 1e3,
                    "path": ("pinned FASTA bytes -> kpal_count_fasta (hybrid upload: the head of the text goes up raw and is packed by the GPU, "
                             "while host threads pack the tail and only its 0.375 B/base cross the bus; "
                             "count + balance kernels, " if world == 1 else
                             "per rank: pinned FASTA bytes -> kpal_count_fasta_dev_table (H2D in chunks, GPU scan/pack, "
                             "count), kpal_dev_slice_push, kpal_dev_slice_collect_to_host (this rank's slice: "
                             if slicer is not None else
                             "per rank: pinned FASTA bytes -> kpal_count_fasta_to_dev (H2D in chunks, GPU scan/pack, "
                             "count); table sum onto rank 0; there kpal_dev_table_to_host (widen + balance, ") +
                            (("D2H as uint%d in chunks%s, widened to the int64 profile by host threads)" % (
                                8 * narrow_width, " + a side list of the %d larger counts" % listed if listed else "")
                              if args.dma_share == 0 else
                              "D2H of the first %d/16 of the profile as uint%d in chunks, widened to int64 by host "
                              "threads, the rest as int64 by the copy engine meanwhile)" % (16 - args.dma_share, 8 * narrow_width))
                             if narrow and narrow_width < 8 else "D2H int64)")},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm",
                         "kernel": " + ".join(step_kernels if world == 1 else kernels),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes,
                         "accounting": "0.375 B/base (2-bit codes + validity, read once) + 8 B x 4^k (int64 "
                                       "profile written once, balance fused): SURVEY.md section 8d; time = "
                                       + ("the whole step" if world == 1 else "this rank's count kernels"),
                         "step_ms": step_ms, "count_kernels_ms": kern_ms, "peak_source": peak_src,
                         "windows_per_s": n_windows / (kern_ms * 1e-3)},
            "cpu_baseline": cpu, "clocks": clocks, "parity_ok": result_ok, "parity": parity,
            "wall_s_timed_region": wall,
        }
        if tail_ms is not None:
            out["reduce_tail"] = dict(tail_ms, note="rank 0's stream: balance + narrow push kernels | wait for "
                                      "the peers' signals + collect of its slice (CUDA events)")
    if reducer is not None:
        reducer.close()
    if shared is not None:
        shared.close()
    if slicer is not None:
        slicer.close()
    del d_codes, d_valid, d_table, d_counts, flush
    torch.cuda.empty_cache()
    return out


def matrix_slab(torch, gen, lam, slab_index, slab, n, d, dev):
    """int64 counts of profiles [slab_index * slab, ...) of the synthetic set (SURVEY.md 8d cfg 4:
    Poisson(lambda_i) per bin): seeded per slab, so any rank can (re)build any slab."""
    r0 = slab_index * slab
    m = min(slab, n - r0)
    gen.manual_seed(4000 + slab_index)
    rates = lam[r0:r0 + m, None].expand(m, d).to(torch.float32)
    return torch.poisson(rates, generator=gen).to(torch.int64)


def bench_matrix(args, world, rank, local):
    import torch
    import torch.distributed as dist
    from kpal_b200 import _cabi, multigpu

    numa_bind(args, world, local)
    L = _cabi.load()
    _cabi.check(L.kpal_set_device(local))
    dev = torch.device("cuda", local)
    n, k = args.profiles, K_MATRIX
    d = 4 ** k
    steps = max(1, min(args.steps, 2))
    stride = int(L.kpal_prepared_stride(k))
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    # ---- synthetic profile set: rank r generates and prepares rows shard_rows(n, r, world) only
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    lam = torch.exp(torch.empty(n, device=dev, dtype=torch.float64).uniform_(
        float(np.log(0.5)), float(np.log(8.0)), generator=gen))
    F = torch.empty((n, stride), dtype=torch.float64, device=dev)
    R = torch.empty((n, stride), dtype=torch.float64, device=dev)
    bitmap = torch.empty((n, stride // 32), dtype=torch.int32, device=dev)
    totals = torch.empty(n, dtype=torch.float64, device=dev)
    norm2 = torch.empty(n, dtype=torch.float64, device=dev)
    order = torch.empty(n, dtype=torch.int32, device=dev)
    slab = 256
    row_b, row_e = multigpu.shard_rows(n, rank, world)
    # end-to-end leg: this rank's rows as int64 in pinned host memory (what kmer.distance_matrix
    # reads from the profile file); at N = 1 bounded by the host's free RAM
    n_e2e = 0 if args.no_e2e else n
    if n_e2e and world == 1:
        avail = host_mem_available()
        while n_e2e > 64 and avail is not None and n_e2e * d * 8 * 2.5 > avail:
            n_e2e //= 2
    e2e_b, e2e_e = (row_b, row_e) if world > 1 else (0, n_e2e)
    host_rows = _cabi.PinnedArray((max(e2e_e - e2e_b, 1), d), np.int64) if n_e2e else None
    h_t = torch.from_numpy(host_rows.array) if n_e2e else None
    for s_i in range(row_b // slab, (row_e + slab - 1) // slab):
        counts = matrix_slab(torch, gen, lam, s_i, slab, n, d, dev)
        r0 = s_i * slab
        lo, hi = max(r0, row_b), min(r0 + len(counts), row_e)
        part = counts[lo - r0:hi - r0]
        if n_e2e:
            a, b = max(lo, e2e_b), min(hi, e2e_e)
            if b > a:
                h_t[a - e2e_b:b - e2e_b].copy_(counts[a - r0:b - r0])
        _cabi.check(L.kpal_dev_profiles_prepare(
            part.data_ptr(), hi - lo, k, 0, 1, F[lo].data_ptr(), R[lo].data_ptr(),
            bitmap[lo].data_ptr(), totals[lo:].data_ptr(), norm2[lo:].data_ptr(), sp))
        del counts, part
    torch.cuda.synchronize()
    barrier_sync(world)
    t0 = time.perf_counter()
    multigpu.allgather_rows([F, R, bitmap, totals, norm2], n)
    torch.cuda.synchronize()
    allgather_s = max_over_ranks(time.perf_counter() - t0, world)
    _cabi.check(L.kpal_dev_order_by_total(totals.data_ptr(), n, 0, order.data_ptr(), sp))
    tiles = int(L.kpal_distance_num_tiles(n))
    t_begin, t_end = multigpu.tile_range(tiles, rank, world)
    tile_elems = int(L.kpal_distance_tile_elems())
    out = torch.zeros((n, n), dtype=torch.float64, device=dev) if rank == 0 else None
    if world > 1:
        most = max(multigpu.tile_range(tiles, r, world)[1] - multigpu.tile_range(tiles, r, world)[0]
                   for r in range(world))
        packed = torch.zeros((most, tile_elems), dtype=torch.float64, device=dev)
        parts = [torch.zeros((most, tile_elems), dtype=torch.float64, device=dev) for _ in range(world)] \
            if rank == 0 else None

    def device_step():
        if world == 1:
            _cabi.check(L.kpal_dev_distance_tiles(
                F.data_ptr(), R.data_ptr(), bitmap.data_ptr(), totals.data_ptr(), norm2.data_ptr(),
                order.data_ptr(), n, k, 0, 0, 1, 0, t_begin, t_end, out.data_ptr(), sp))
            return
        _cabi.check(L.kpal_dev_distance_tiles_packed(
            F.data_ptr(), R.data_ptr(), bitmap.data_ptr(), totals.data_ptr(), norm2.data_ptr(),
            order.data_ptr(), n, k, 0, 0, 1, 0, t_begin, t_end, packed.data_ptr(), sp))
        dist.gather(packed, parts, dst=0)
        if rank == 0:
            for r in range(world):
                b, e = multigpu.tile_range(tiles, r, world)
                _cabi.check(L.kpal_dev_distance_unpack_tiles(
                    parts[r].data_ptr(), totals.data_ptr(), norm2.data_ptr(), order.data_ptr(), n, 0, 0, 1,
                    b, e, int(r == 0), out.data_ptr(), sp))

    for _ in range(max(min(args.warmup, 1), 1)):
        device_step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.kpal_reset_kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(steps)]
    barrier_sync(world)
    for i in range(steps):
        if rank == 0:
            out.zero_()
        evs[i][0].record(stream)
        device_step()
        evs[i][1].record(stream)
    barrier_sync(world)
    launches = int(L.kpal_kernel_launches())
    step_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity against the oracle (rank 0): the full leading 64 x 64 block and 2000 random
    # pairs among 96 profiles drawn from the whole set
    parity = None
    if rank == 0:
        from oracle import c_oracle
        whole_host()
        threads = host_threads()
        rng = np.random.default_rng(44)
        lead = min(64, n)
        far = np.sort(rng.choice(n, size=min(96, n), replace=False))
        need = sorted(set(range(lead)) | set(int(i) for i in far))
        rows = {}
        for s_i in sorted(set(i // slab for i in need)):
            counts = matrix_slab(torch, gen, lam, s_i, slab, n, d, dev)
            for i in need:
                if i // slab == s_i:
                    rows[i] = counts[i - s_i * slab].cpu().numpy()
            del counts
        got = out.cpu().numpy()
        t0 = time.perf_counter()
        block = c_oracle.distance_matrix(np.stack([rows[i] for i in range(lead)]), do_scale=True, threads=threads)
        cpu_s = time.perf_counter() - t0
        low = np.tril_indices(lead, -1)
        rel_block = float(np.max(np.abs(got[:lead, :lead][low] - block[low]) / np.abs(block[low]))) if lead > 1 else 0.0
        sub = c_oracle.distance_matrix(np.stack([rows[int(i)] for i in far]), do_scale=True, threads=threads)
        ii, jj = np.tril_indices(len(far), -1)
        pick = rng.choice(len(ii), size=min(2000, len(ii)), replace=False)
        a, b = far[ii[pick]], far[jj[pick]]
        rel_far = float(np.max(np.abs(got[a, b] - sub[ii[pick], jj[pick]]) / np.abs(sub[ii[pick], jj[pick]])))
        sym = bool(np.array_equal(got[a, b], got[b, a]))
        parity = {"checked_against": "C port of kdistlib.py:126-161 / metrics.py:49-123",
                  "leading_block_pairs": int(len(low[0])), "max_rel_err_leading_block": rel_block,
                  "random_pairs": int(len(pick)), "random_pairs_drawn_from_profiles": int(len(far)),
                  "max_rel_err_random_pairs": rel_far, "symmetric": sym, "tolerance": 1e-9}
        cpu = {"value": len(low[0]) / cpu_s, "unit": "profile-pairs/s", "cores": threads, "kind": "port",
               "sample": "leading %d profiles (%d pairs), C port, OpenMP" % (lead, len(low[0]))}
        check_block = got[:lead, :lead].copy()
        del got

    # ---- euclidean distances of the same set through the exact integer Gram matrix on the
    # tensor cores (tcgen05 kind::i8), next to the element-wise fp64 tile kernel (N = 1 only)
    gram_rec = None
    if world == 1 and not args.no_gram:
        dp = int(L.kpal_gram_row_stride(k))
        x8 = torch.zeros((n, dp), dtype=torch.uint8, device=dev)
        g_tot = torch.zeros(n, dtype=torch.int64, device=dev)
        g_norm = torch.zeros(n, dtype=torch.int64, device=dev)
        g_flags = torch.zeros(4, dtype=torch.int32, device=dev)
        for s_i in range((n + slab - 1) // slab):
            counts = matrix_slab(torch, gen, lam, s_i, slab, n, d, dev)
            r0 = s_i * slab
            _cabi.check(L.kpal_dev_gram_prepare(counts.data_ptr(), len(counts), k, 0, x8[r0].data_ptr(),
                                                g_tot[r0:].data_ptr(), g_norm[r0:].data_ptr(), g_flags.data_ptr(), sp))
            del counts
        torch.cuda.synchronize()
        if int(g_flags[0].item()) == 0:
            norm_max = int(g_norm.max().item())
            g_mat = torch.empty((n, n), dtype=torch.int64, device=dev)
            g_out = torch.empty((n, n), dtype=torch.float64, device=dev)

            def gram_step():
                _cabi.check(L.kpal_dev_gram_distances(x8.data_ptr(), g_tot.data_ptr(), g_norm.data_ptr(), norm_max,
                                                      n, k, 1, 1, 0, g_mat.data_ptr(), g_out.data_ptr(), sp))
            gram_step()
            torch.cuda.synchronize()
            g_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            for a_ev, b_ev in g_ev:
                a_ev.record(stream)
                gram_step()
                b_ev.record(stream)
            torch.cuda.synchronize()
            gram_ms = sum(x.elapsed_time(y) for x, y in g_ev) / len(g_ev)
            # the same matrix from the element-wise fp64 tile kernel, once
            t_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            t_ev[0].record(stream)
            _cabi.check(L.kpal_dev_distance_tiles(
                F.data_ptr(), R.data_ptr(), bitmap.data_ptr(), totals.data_ptr(), norm2.data_ptr(),
                order.data_ptr(), n, k, 1, 0, 1, 0, 0, tiles, out.data_ptr(), sp))
            t_ev[1].record(stream)
            torch.cuda.synchronize()
            tile_ms = t_ev[0].elapsed_time(t_ev[1])
            g_host, t_host = g_out.cpu().numpy(), out.cpu().numpy()
            from oracle import c_oracle
            want = c_oracle.distance_matrix(np.stack([rows[i] for i in range(lead)]), metric="euclidean",
                                            do_scale=True, threads=host_threads())
            lowb = np.tril_indices(lead, -1)
            rel_gram = float(np.max(np.abs(g_host[:lead, :lead][lowb] - want[lowb]) / np.abs(want[lowb])))
            iu = np.triu_indices(n, 1)
            rel_tile = float(np.max(np.abs(g_host[iu] - t_host[iu]) / np.abs(t_host[iu])))
            pairs_all = n * (n - 1) // 2
            tensor_peak, tensor_src = measured_peak("bf16_tflops", 1590.0)
            ops = 2.0 * d * (n * n / 2.0)            # multiply-adds of the triangle
            gram_rec = {"metric": "profile_pairs_per_sec_k10_euclidean_scaled", "value": pairs_all / (gram_ms * 1e-3),
                        "unit": "profile-pairs/s", "ms_per_step": gram_ms, "steps": len(g_ev),
                        "kernel": "gram_u8_kernel (tcgen05.mma kind::i8, TMEM accumulators, TMA 128B-swizzled loads) "
                                  "+ gram_finalize_kernel",
                        "fp64_tile_kernel_ms": tile_ms, "speedup_vs_fp64_tile_kernel": tile_ms / gram_ms,
                        "roofline": {"bound": "tensor", "achieved": ops / (gram_ms * 1e-3) / 1e12, "unit": "TOP/s",
                                     "peak": 2 * tensor_peak, "frac": ops / (gram_ms * 1e-3) / 1e12 / (2 * tensor_peak),
                                     "peak_source": "2 x the bf16 figure (int8 runs at twice the bf16 rate): " + tensor_src,
                                     "ops": ops},
                        "max_rel_err_vs_oracle_leading_block": rel_gram, "max_rel_diff_vs_fp64_tile_kernel": rel_tile,
                        "tolerance": 1e-9, "parity_ok": bool(rel_gram <= 1e-9 and rel_tile <= 1e-9)}
            del g_mat, g_out
        del x8

    # ---- end to end through the public entry points, host buffers in, host matrix out
    numa_bind(args, world, local)
    e2e = None
    if n_e2e:
        del F, R, bitmap, out
        if world > 1:
            del packed, parts
        torch.cuda.empty_cache()
        host_out = _cabi.PinnedArray((n_e2e, n_e2e), np.float64) if (rank == 0 and world == 1) else None

        def e2e_step():
            if world == 1:
                _cabi.check(L.kpal_distance_matrix(host_rows._ptr, n_e2e, k, 0, 0, 0, 1, 0, host_out._ptr))
                return host_out.array
            return multigpu.distance_matrix_distributed(host_rows.array[:row_e - row_b], do_scale=True,
                                                        sharded=True, n_total=n, device=dev)
        e2e_step()
        e2e_steps = 1 if n_e2e >= 2048 else steps
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = e2e_step()
        barrier_sync(world)
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps, world)
        if rank == 0:
            e2e_pairs = n_e2e * (n_e2e - 1) // 2
            e2e_ok = None
            if n_e2e == n:
                e2e_ok = bool(np.allclose(res[:lead, :lead], check_block, rtol=1e-12, atol=0))
            h2d = int(n_e2e * d * 8)
            e2e = {"value": e2e_pairs / e2e_s, "unit": "profile-pairs/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n_e2e * n_e2e * 8),
                   "ms_per_step": e2e_s * 1e3, "profiles": n_e2e, "steps": e2e_steps,
                   "matches_device_result": e2e_ok,
                   "path": ("pinned int64 profiles -> kpal_distance_matrix (H2D in slabs, prepare, order, "
                            "tile kernel, D2H of the N x N float64 matrix)" if world == 1 else
                            "per rank: its 1/%d of the int64 profiles (pinned host) -> multigpu."
                            "distance_matrix_distributed (H2D + prepare of the shard, NCCL all-gather of the "
                            "prepared set, tile range, gather of packed tiles, scatter + D2H on rank 0); "
                            "h2d bytes summed over the ranks" % world)}

    rec = None
    if rank == 0:
        pairs = n * (n - 1) // 2
        # per-GPU roofline: rank 0's share of the tiles (the tile kernel is >99 % of the step)
        flops = 8.0 * d * pairs * (t_end - t_begin) / max(tiles, 1)
        peak_tf, peak_src = measured_fp64_peak()
        achieved_tf = flops / (step_ms * 1e-3) / 1e12
        traffic, traffic_src = traffic_record("distance_tile_kernel")
        rec = {
            "metric": "profile_pairs_per_sec_k10_multiset", "value": pairs / (step_ms * 1e-3),
            "unit": "profile-pairs/s", "n_gpus": world, "steps": steps,
            "warmup": 1, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kpal matrix, multiset/prod, scaled, %d profiles, k=10; BASELINE configs[3]" % n,
                       "profiles": n, "k": k, "l2": "68 GB working set >> L2",
                       "parallelism": ("1/%d of the profiles prepared per rank, NCCL all-gather of the prepared set "
                                       "(%.1f ms, outside the timed step), upper-triangle tiles in equal ranges, "
                                       "gather of packed tiles onto rank 0 (inside the step)" % (world, allgather_s * 1e3))
                       if world > 1 else "1 GPU"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "kernel": "distance_tile_kernel<prod>",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "flops_per_element_pair": 8, "peak_source": peak_src},
            "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
            "euclidean_gram": gram_rec,
            "parity": parity,
            "parity_ok": bool(parity["max_rel_err_leading_block"] <= 1e-9 and
                              parity["max_rel_err_random_pairs"] <= 1e-9 and parity["symmetric"]),
        }
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "count", "matrix"],
                    help="all (default) = the count line with the matrix record as its 'matrix' sub-record")
    ap.add_argument("--profiles", type=int, default=N_PROFILES)
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 5],
                    help="count workload: 2 = BASELINE configs[1] (default), 1 = configs[0] (k=6, 1 Mbp), "
                         "5 = configs[4] (k=13 genome shards; the full 3 Gbp job at --gpus 8)")
    ap.add_argument("--composition", default="uniform", choices=["uniform", "skewed"],
                    help="config 2: 'skewed' makes 10 %% of the reads low-complexity repeats")
    ap.add_argument("--k", type=int, default=0, help="override the k-mer length of the count workload")
    ap.add_argument("--mbp-per-gpu", type=float, default=375.0, help="config 5: Mbp per GPU")
    ap.add_argument("--count-path", type=int, default=0, choices=[0, 1, 2, 3],
                    help="0 = library default, 1 = scattered-RED kernel, 2 = radix-partitioned path "
                         "(two windows per payload up to k = 12), 3 = one-window radix path")
    ap.add_argument("--fresh", type=int, default=1, choices=[0, 1],
                    help="device step: 1 = the count call zeroes the table itself (inside the first kernel on "
                         "the pair path), 0 = memset, then count")
    ap.add_argument("--pair-flush-every", type=int, default=0, help="pair path: tiles between slot flushes (0 = auto)")
    ap.add_argument("--pair-fused", type=int, default=1, choices=[0, 1],
                    help="pair path, pass 2: 1 = both roles in one launch, flushed with cp.reduce.async.bulk")
    ap.add_argument("--radix-payload-bits", type=int, default=0)
    ap.add_argument("--radix-max-buckets", type=int, default=0, choices=[0, 1024, 2048],
                    help="buckets binned per pass-1 launch of the radix count (0 = library default)")
    ap.add_argument("--radix-debug", type=int, default=0, help="timing experiments (results are wrong)")
    ap.add_argument("--radix-shape", type=int, default=0)
    ap.add_argument("--reduce", default="auto", choices=["auto", "slices", "fused", "peer", "nccl"],
                    help="count workload at N > 1: slices (default) = balance + narrow reduce-scatter + distributed "
                         "finalize over NVLink peer memory; fused / peer = u32 table all-to-all onto rank 0 (fused "
                         "into the one-window radix count / as push + collect kernels); nccl = dist.reduce")
    ap.add_argument("--numa-bind", type=int, default=1, choices=[0, 1],
                    help="N > 1: pin every rank to the CPUs next to its GPU (NVML ideal affinity)")
    ap.add_argument("--no-e2e", action="store_true", help="matrix workload: skip the host-buffer end-to-end leg")
    ap.add_argument("--no-gram", action="store_true",
                    help="matrix workload: skip the euclidean (tensor-core Gram form) measurement")
    ap.add_argument("--fasta-chunks", type=int, default=0, help="chunks of the pipelined FASTA upload (0 = auto)")
    ap.add_argument("--narrow-d2h", type=int, default=1, choices=[0, 1, 2],
                    help="e2e leg: 1 = the profile leaves the device as uint8 / uint16 (the narrowest that "
                         "holds every count) and host threads widen it (library default), 2 = uint16 only, "
                         "0 = plain int64 copy")
    ap.add_argument("--fasta-hybrid", type=int, default=1,
                    help="e2e leg: 1 = host threads pack segments from the end of the text while its head uploads raw "
                         "(hybrid upload, cabi.cu fasta_hybrid_count), 0 = the whole text uploads raw, "
                         "2..64 = at most that many host packers")
    ap.add_argument("--fasta-hybrid-share", type=int, default=0,
                    help="e2e leg, hybrid upload: percent of the text the host packs (0 = adapted from call to call)")
    ap.add_argument("--fasta-split", type=int, default=0, choices=[0, 1],
                    help="e2e leg: 1 = a large FASTA text is cut at a header line into two parts, the first counted "
                         "while the second is uploaded, 0 = one part (library default: the cut measured no gain)")
    ap.add_argument("--dma-share", type=int, default=0,
                    help="e2e leg: sixteenths of the narrow-copied profile that the copy engine moves as int64 "
                         "straight into the pinned result while the host threads widen the rest (library default 0: "
                         "measured slower than widening everything, profiles/r01_e2e_trace.log)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world, rank, local = init_dist(args)
    rec = None
    if args.workload in ("all", "count"):
        rec = bench_count(args, world, rank, local)
    if args.workload in ("all", "matrix"):
        mrec = bench_matrix(args, world, rank, local)
        if rank == 0:
            if rec is None:
                rec = mrec
            else:
                rec["matrix"] = mrec
                rec["gpu_launches"] += mrec["gpu_launches"]
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
