#!/usr/bin/env python
"""
bench.py -- headline benchmark of the kPAL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload count|matrix]
                    [--impl ours|reference]

Default workload (BASELINE.json configs[1], the one `metric` is quoted on):
`kpal count` at k=12 over 100 Mbp of synthetic 150-bp reads, with balance.
A step = one pass of the hot path over the whole batch:

  value : packed sequence already resident in HBM -> int64 balanced profile in
          HBM (memset + count kernels + [table sum over the ranks] + widen/balance
          kernel), timed with CUDA events on the launching stream, L2 flushed
          between steps (untimed), max over ranks;
  e2e   : the same job through the host-buffer C ABI (kpal_count_fasta): pinned
          FASTA bytes -> H2D in chunks -> GPU scan/pack -> count kernels ->
          [table sum] -> narrow (uint8 / uint16) D2H of the profile, widened to
          int64 by host threads; wall clock around the call with device syncs.

Multi-GPU (weak scaling): every rank counts its own shard of records (same
size per rank), the 4^k u32 tables are summed onto rank 0 -- over NVLink peer
memory fused into the count's histogram pass at 2 GPUs, with an NCCL reduce
from 4 GPUs on (`--reduce auto`, chosen by measurement) -- and finalised
(widen + balance) once.  `--config 5`: one GPU's shard of BASELINE configs[4]
(k=13 genome-like records).

`--workload matrix`: BASELINE.json configs[3], the 4096-profile k=10 scaled
multiset distance matrix (profile-pairs/s); tiles sharded over the ranks.

`--impl reference`: the CPU baseline -- the oracle's C port of the reference
algorithm on all host threads (the reference itself is pure Python and cannot
run on the GPU box; BASELINE.md has its measured 1-core figures).
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


def recorded_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------- reference arm
def run_reference(args):
    """CPU baseline: the oracle's C port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle, kpal_oracle as ko
    threads = c_oracle.max_threads()
    if args.workload == "count":
        reads = synthetic_reads(1000)
        text = np.insert(reads, READ_LEN, ord("\n"), axis=1).reshape(-1)
        n_bases = reads.size
        rc = ko.reverse_complement_table(K_COUNT)

        def step():
            counts = c_oracle.count_bytes(text, K_COUNT, threads=threads)
            return c_oracle.balance(counts)            # literal klib.py:285-298 loop in C
        sample = "full workload: %d reads x %d bp, k=%d, balance (C port, OpenMP)" % (
            N_READS, READ_LEN, K_COUNT)
        units, unit, metric = n_bases / 1e9, "Gbases/s", "gbases_per_sec_counted_k12"
        config = {"workload": "kpal count k=12, 100 Mbp of 150-bp reads, balance", "k": K_COUNT}
    else:
        n = 48
        rng = np.random.default_rng(4)
        lam = np.exp(rng.uniform(np.log(0.5), np.log(8.0), n))
        profiles = np.stack([rng.poisson(l, 4 ** K_MATRIX) for l in lam]).astype(np.int64)

        def step():
            return c_oracle.distance_matrix(profiles, do_scale=True, threads=threads)
        sample = "leading %d profiles (%d pairs) of the 4096-profile k=10 set (C port, OpenMP)" % (
            n, n * (n - 1) // 2)
        units, unit, metric = n * (n - 1) / 2, "profile-pairs/s", "profile_pairs_per_sec_k10_multiset"
        config = {"workload": "kpal matrix multiset/prod scaled, 4096 profiles k=10", "k": K_MATRIX}
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = units / dt
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64" if args.workload == "count" else "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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


def bench_count(args):
    import torch
    import torch.distributed as dist
    from kpal_b200 import _cabi

    world, rank, local = init_dist(args)
    L = _cabi.load()
    _cabi.check(L.kpal_set_device(local))
    _cabi.check(L.kpal_set_option(b"count_path", args.count_path))
    _cabi.check(L.kpal_set_option(b"radix_payload_bits", args.radix_payload_bits))
    if args.radix_max_buckets:
        _cabi.check(L.kpal_set_option(b"radix_max_buckets", args.radix_max_buckets))
    _cabi.check(L.kpal_set_option(b"radix_debug", args.radix_debug))
    _cabi.check(L.kpal_set_option(b"radix_shape", args.radix_shape))
    _cabi.check(L.kpal_set_option(b"fasta_chunks", args.fasta_chunks))
    _cabi.check(L.kpal_set_option(b"narrow_d2h", args.narrow_d2h))
    _cabi.check(L.kpal_set_option(b"dma_share", args.dma_share))
    _cabi.check(L.kpal_set_option(b"fasta_split", args.fasta_split))
    dev = torch.device("cuda", local)

    # ---- this rank's shard of records (same size on every rank: weak scaling)
    if args.config == 5:
        # BASELINE configs[4]: k = 13, human-genome-sized FASTA sharded over the GPUs
        # (3 x 125 Mbp records per GPU at the full 8-GPU size)
        k = args.k or 13
        rec_len = int(args.mbp_per_gpu * 1e6) // 3
        records = synthetic_chromosomes(5000 + rank, 3, rec_len)
        fasta_np = records_to_fasta(records, first=3 * rank)
        oracle_text = np.concatenate([np.append(r, np.uint8(10)) for r in records])
        seq_bases_total = sum(len(r) for r in records)
        n_windows = None
        workload = ("kpal count k=%d, %d x %.1f Mbp records per GPU with N blocks and soft-masking, balance; "
                    "BASELINE configs[4]" % (k, 3, rec_len / 1e6))
        del records
    else:
        k = args.k or K_COUNT
        reads = synthetic_reads(1000 + rank)
        fasta_np = reads_to_fasta(reads)
        oracle_text = np.insert(reads, READ_LEN, ord("\n"), axis=1).reshape(-1)
        seq_bases_total = reads.size
        n_windows = reads.size - N_READS * (k - 1)
        workload = ("kpal count k=%d, 100 Mbp of 150-bp reads (666667 records/GPU), balance; "
                    "BASELINE configs[1]" % k)
        del reads
    bins = 4 ** k
    n_fasta = fasta_np.size
    pinned_fasta = _cabi.PinnedArray(n_fasta, np.uint8)
    pinned_fasta.array[:] = fasta_np
    pinned_out = _cabi.PinnedArray(bins, np.int64)
    codes, valid, _, _, n_bases = _cabi.fasta_pack(fasta_np.tobytes())
    seq_bases = seq_bases_total
    d_codes = torch.from_numpy(codes.view(np.int32)).to(dev)
    d_valid = torch.from_numpy(valid.view(np.int32)).to(dev)
    d_table = torch.zeros(bins, dtype=torch.int32, device=dev)
    d_counts = torch.zeros(bins, dtype=torch.int64, device=dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    reducer = None
    reduce_note = None
    if args.reduce == "auto":
        # measured (profiles/README.md): at 2 GPUs the table sum fused into the count's histogram
        # pass is on par with / ahead of the NCCL reduce (0.416 vs 0.421 ms per step), at 4 GPUs the
        # peer stores slow pass 2 down and NCCL wins (0.473 vs 0.427 ms)
        args.reduce = "fused" if world == 2 else "nccl"
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
            _cabi.check(L.kpal_dev_finalize_counts(summed, 32, k, 1, d_counts.data_ptr(), sp))

    def device_step(ev=None):
        d_table.zero_()
        if ev:
            ev[0].record(stream)
        if reducer is not None and args.reduce == "fused":
            # count + all-to-all in one call: pass 2 of the radix count stores into the inboxes
            summed = reducer.count_and_reduce(d_codes.data_ptr(), d_valid.data_ptr(), n_bases,
                                              d_table.data_ptr(), sp)
            if ev:
                ev[1].record(stream)
            if rank == 0:
                _cabi.check(L.kpal_dev_finalize_counts(summed, 32, k, 1, d_counts.data_ptr(), sp))
            return
        _cabi.check(L.kpal_dev_count_packed(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                            d_table.data_ptr(), 32, sp))
        if ev:
            ev[1].record(stream)
        reduce_and_finalize()

    def e2e_step():
        if world == 1:
            _cabi.check(L.kpal_count_fasta(pinned_fasta._ptr, n_fasta, k, 1, pinned_out._ptr))
        else:
            d_table.zero_()
            nb = ctypes.c_uint64()
            _cabi.check(L.kpal_count_fasta_to_dev(pinned_fasta._ptr, n_fasta, k, d_table.data_ptr(),
                                                  32, sp, ctypes.byref(nb)))
            summed = reduce_tables()
            if rank == 0:       # widen + balance + narrow D2H into the host profile
                _cabi.check(L.kpal_dev_table_to_host(summed, 32, k, 1, pinned_out._ptr, sp))
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
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
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
    kern_ms = sum(a.elapsed_time(b) for a, b in kern_ev) / args.steps
    step_ms = max_over_ranks(step_ms, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C ABI
    for _ in range(2):
        e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier_sync(world)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps, world)

    # ---- parity of what was just measured (rank 0, N=1: against the oracle)
    result_ok = None
    if rank == 0 and world == 1:
        from oracle import c_oracle, kpal_oracle as ko
        t0 = time.perf_counter()
        want = c_oracle.count_bytes(oracle_text, k, threads=c_oracle.max_threads())
        want = c_oracle.balance(want)                 # literal klib.py:285-298 loop in C
        cpu_s = time.perf_counter() - t0
        if n_windows is None:
            n_windows = int(want.sum()) // 2
        result_ok = bool(np.array_equal(d_counts.cpu().numpy(), want)
                         and np.array_equal(pinned_out.array, want))
        cpu = {"value": seq_bases / 1e9 / cpu_s, "unit": "Gbases/s", "cores": c_oracle.max_threads(),
               "kind": "port",
               "sample": "full workload once (C port of klib.py:149-170 + balance, OpenMP); the "
                         "reference's pure-Python loop measured 0.001-0.0036 Gbases/s on 1 core "
                         "(BASELINE.md)"}
    else:
        cpu = None
    if world > 1:
        # N > 1: the job's result (rank 0) must equal, bit for bit, the sum of the per-rank
        # tables taken with a plain NCCL reduce and finalised the same way; the host-buffer
        # result of the e2e leg must equal it too.  (Each rank's own table is the N = 1 path,
        # whose oracle parity is the N = 1 run's check.)
        device_step()
        check = d_table.clone()
        d_table.zero_()
        _cabi.check(L.kpal_dev_count_packed(d_codes.data_ptr(), d_valid.data_ptr(), n_bases, k,
                                            d_table.data_ptr(), 32, sp))
        check.copy_(d_table)
        dist.reduce(check, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            want_dev = torch.empty_like(d_counts)
            _cabi.check(L.kpal_dev_finalize_counts(check.data_ptr(), 32, k, 1, want_dev.data_ptr(), sp))
            torch.cuda.synchronize()
            result_ok = bool(torch.equal(want_dev, d_counts)
                             and np.array_equal(pinned_out.array, want_dev.cpu().numpy()))

    if rank == 0:
        total_bases = seq_bases * world
        peak, peak_src = measured_peak("hbm_gbs", 6650.0)
        alg_bytes = 0.375 * n_bases + 4 * bins      # packed stream read once + u32 table written once
        radix = args.count_path == 2 or (args.count_path == 0 and n_bases >= ((16 << 20) if k <= 12 else (4 << 20)))
        if n_windows is None:
            n_windows = seq_bases
        narrow = bool(args.narrow_d2h and bins >= (1 << 20))
        # bytes of the narrow copy: uint8 when every count of the result fits, else uint16
        # (what finalize_to_host decides from the device's flag words), + the 8 flag bytes
        # -- for the bins below the split; the last dma_share/16 of the (pinned) profile is
        # copied as int64 by the DMA engine
        split = bins // 16 * (16 - args.dma_share)
        top = int(pinned_out.array[:split].max())
        narrow_width = 8 if (not narrow or top > 65535) else (1 if (args.narrow_d2h == 1 and top <= 255) else 2)
        d2h_bytes = bins * 8 if narrow_width == 8 else split * narrow_width + (bins - split) * 8 + 8
        count_kernel_name = ("radix_partition_kernel<u32> + radix_histogram_kernel<u32>" if radix
                             else "count_global_kernel<u32>")
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        out = {
            "metric": "gbases_per_sec_counted_k%d" % k, "value": total_bases / 1e9 / (step_ms * 1e-3),
            "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 counters -> int64", "data": "synthetic",
            "config": {"workload": workload,
                       "k": k, "bases_per_gpu": int(seq_bases), "packed_bases_per_gpu": int(n_bases),
                       "l2": "512 MB memset between steps (untimed); table memset is inside the step",
                       "parallelism": ("records sharded per GPU, u32 tables summed onto rank 0 " +
                                       ("over NVLink peer memory (all-to-all fused into the count's histogram pass + collect kernel)"
                                        if (reducer is not None and args.reduce == "fused") else
                                        "over NVLink peer memory (all-to-all push + collect kernels)"
                                        if reducer is not None else "with an NCCL reduce"))
                       if world > 1 else "1 GPU",
                       **({"reduce_note": reduce_note} if reduce_note else {})},
            "e2e": {"value": total_bases / 1e9 / e2e_s, "unit": "Gbases/s",
                    "h2d_bytes_per_step": int(n_fasta * world),
                    "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": e2e_s * 1e3,
                    "path": ("pinned FASTA bytes -> kpal_count_fasta (H2D of the raw text in chunks, GPU scan/pack, "
                             "count + balance kernels, " if world == 1 else
                             "per rank: pinned FASTA bytes -> kpal_count_fasta_to_dev (H2D in chunks, GPU scan/pack, "
                             "count); table sum onto rank 0; there kpal_dev_table_to_host (widen + balance, ") +
                            (("D2H as uint%d in chunks, widened to the int64 profile by host threads)" % (8 * narrow_width)
                              if args.dma_share == 0 else
                              "D2H of the first %d/16 of the profile as uint%d in chunks, widened to int64 by host "
                              "threads, the rest as int64 by the copy engine meanwhile)" % (16 - args.dma_share, 8 * narrow_width))
                             if narrow and narrow_width < 8 else "D2H int64)")},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": count_kernel_name, "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # the committed ncu capture is of the default workload (config 2, k = 12)
                         "traffic": (recorded_traffic(count_kernel_name.split("<")[0].split(" ")[0])
                                     if args.config == 2 and k == K_COUNT else None),
                         "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms, "peak_source": peak_src,
                         "windows_per_s": n_windows / (kern_ms * 1e-3)},
            "cpu_baseline": cpu, "clocks": clocks, "parity_ok": result_ok,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(out))
    if reducer is not None:
        reducer.close()
    if world > 1:
        dist.destroy_process_group()


def bench_matrix(args):
    import torch
    import torch.distributed as dist
    from kpal_b200 import _cabi

    world, rank, local = init_dist(args)
    L = _cabi.load()
    _cabi.check(L.kpal_set_device(local))
    dev = torch.device("cuda", local)
    n, k = args.profiles, K_MATRIX
    d = 4 ** k
    stride = int(L.kpal_prepared_stride(k))
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    # ---- synthetic profile set (SURVEY.md 8d cfg 4): Poisson(lambda_i), generated on device
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
    out = torch.zeros((n, n), dtype=torch.float64, device=dev)
    slab = 256
    first_rows = None
    # end-to-end leg: the same profile set as int64 rows in pinned host memory (the array
    # kmer.distance_matrix hands to kpal_distance_matrix); bounded by the host's free RAM
    n_e2e = n if (world == 1 and not args.no_e2e) else 0
    if n_e2e:
        avail = host_mem_available()
        while n_e2e > 64 and avail is not None and n_e2e * d * 8 * 2.5 > avail:
            n_e2e //= 2
        host_profiles = _cabi.PinnedArray((n_e2e, d), np.int64)
        host_out = _cabi.PinnedArray((n_e2e, n_e2e), np.float64)
        h_t = torch.from_numpy(host_profiles.array)
    for r0 in range(0, n, slab):
        m = min(slab, n - r0)
        rates = lam[r0:r0 + m, None].expand(m, d).to(torch.float32)
        counts = torch.poisson(rates, generator=gen).to(torch.int64)
        if r0 == 0:
            first_rows = counts[:64].cpu().numpy()
        if r0 < n_e2e:
            mm = min(m, n_e2e - r0)
            h_t[r0:r0 + mm].copy_(counts[:mm])
        _cabi.check(L.kpal_dev_profiles_prepare(
            counts.data_ptr(), m, k, 0, 1, F[r0].data_ptr(), R[r0].data_ptr(),
            bitmap[r0].data_ptr(), totals[r0:].data_ptr(), norm2[r0:].data_ptr(), sp))
        del counts, rates
    _cabi.check(L.kpal_dev_order_by_total(totals.data_ptr(), n, 0, order.data_ptr(), sp))
    tiles = int(L.kpal_distance_num_tiles(n))
    t_begin = tiles * rank // world
    t_end = tiles * (rank + 1) // world

    def device_step():
        _cabi.check(L.kpal_dev_distance_tiles(
            F.data_ptr(), R.data_ptr(), bitmap.data_ptr(), totals.data_ptr(), norm2.data_ptr(),
            order.data_ptr(), n, k, 0, 0, 1, 0, t_begin, t_end, out.data_ptr(), sp))
        if world > 1:
            dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)

    for _ in range(max(args.warmup, 1)):
        out.zero_()
        device_step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.kpal_reset_kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier_sync(world)
    for i in range(args.steps):
        out.zero_()
        evs[i][0].record(stream)
        device_step()
        evs[i][1].record(stream)
    barrier_sync(world)
    launches = int(L.kpal_kernel_launches())
    step_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / args.steps, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C ABI (every rank on its own replica of the
    # call at N>1 would only repeat rank 0: measured on rank 0 at N=1)
    e2e = None
    got = out[:64, :64].cpu().numpy() if rank == 0 else None
    if world == 1 and n_e2e:
        check_block = out[:64, :64].clone()
        del F, R, bitmap, out
        torch.cuda.empty_cache()

        def e2e_step():
            _cabi.check(L.kpal_distance_matrix(host_profiles._ptr, n_e2e, k, 0, 0, 0, 1, 0, host_out._ptr))
        e2e_step()
        e2e_steps = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e_pairs = n_e2e * (n_e2e - 1) // 2
        e2e_ok = None
        if n_e2e == n:
            a, b = host_out.array[:64, :64], check_block.cpu().numpy()
            e2e_ok = bool(np.allclose(a, b, rtol=1e-12, atol=0))
        e2e = {"value": e2e_pairs / e2e_s, "unit": "profile-pairs/s",
               "h2d_bytes_per_step": int(n_e2e * d * 8), "d2h_bytes_per_step": int(n_e2e * n_e2e * 8),
               "ms_per_step": e2e_s * 1e3, "profiles": n_e2e, "steps": e2e_steps,
               "matches_device_result": e2e_ok,
               "path": "pinned int64 profiles -> kpal_distance_matrix (H2D in 1 GiB slabs, prepare, order, "
                       "tile kernel, D2H of the N x N float64 matrix)"}

    if rank == 0:
        pairs = n * (n - 1) // 2
        from oracle import c_oracle
        t0 = time.perf_counter()
        want = c_oracle.distance_matrix(first_rows[:24], do_scale=True, threads=c_oracle.max_threads())
        cpu_s = time.perf_counter() - t0
        low = np.tril_indices(24, -1)
        rel = float(np.max(np.abs(got[:24, :24][low] - want[low]) / np.abs(want[low])))
        # per-GPU roofline: rank 0's share of the tiles (the tile kernel is >99 % of the step)
        flops = 8.0 * d * pairs * (t_end - t_begin) / max(tiles, 1)
        peak_tf, peak_src = measured_fp64_peak()
        achieved_tf = flops / (step_ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "profile_pairs_per_sec_k10_multiset", "value": pairs / (step_ms * 1e-3),
            "unit": "profile-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kpal matrix, multiset/prod, scaled, %d profiles, k=10; BASELINE configs[3]" % n,
                       "profiles": n, "k": k, "l2": "68 GB working set >> L2",
                       "parallelism": "upper-triangle tiles sharded over ranks, NCCL reduce of the result" if world > 1 else "1 GPU"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "kernel": "distance_tile_kernel<prod>",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": recorded_traffic("distance_tile_kernel"),
                         "flops_per_element_pair": 8, "peak_source": peak_src},
            "e2e": e2e,
            "cpu_baseline": {"value": 276 / cpu_s, "unit": "profile-pairs/s",
                             "cores": c_oracle.max_threads(), "kind": "port",
                             "sample": "leading 24 profiles (276 pairs), C port, OpenMP"},
            "clocks": clocks, "max_rel_err_vs_oracle_276_pairs": rel, "parity_ok": bool(rel <= 1e-9),
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="count", choices=["count", "matrix"])
    ap.add_argument("--profiles", type=int, default=N_PROFILES)
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="count workload: 2 = BASELINE configs[1] (default), 5 = configs[4] (k=13 genome shards)")
    ap.add_argument("--k", type=int, default=0, help="override the k-mer length of the count workload")
    ap.add_argument("--mbp-per-gpu", type=float, default=375.0, help="config 5: Mbp per GPU")
    ap.add_argument("--count-path", type=int, default=0, choices=[0, 1, 2],
                    help="0 = library default, 1 = scattered-RED kernel, 2 = radix-partitioned path")
    ap.add_argument("--radix-payload-bits", type=int, default=0)
    ap.add_argument("--radix-max-buckets", type=int, default=0, choices=[0, 1024, 2048],
                    help="buckets binned per pass-1 launch of the radix count (0 = library default)")
    ap.add_argument("--radix-debug", type=int, default=0, help="timing experiments (results are wrong)")
    ap.add_argument("--radix-shape", type=int, default=0)
    ap.add_argument("--reduce", default="auto", choices=["auto", "fused", "peer", "nccl"],
                    help="count workload at N > 1: table sum over NVLink peer memory fused into the count "
                         "(default), as separate push/collect kernels, or with dist.reduce")
    ap.add_argument("--no-e2e", action="store_true", help="matrix workload: skip the host-buffer end-to-end leg")
    ap.add_argument("--fasta-chunks", type=int, default=0, help="chunks of the pipelined FASTA upload (0 = auto)")
    ap.add_argument("--narrow-d2h", type=int, default=1, choices=[0, 1, 2],
                    help="e2e leg: 1 = the profile leaves the device as uint8 / uint16 (the narrowest that "
                         "holds every count) and host threads widen it (library default), 2 = uint16 only, "
                         "0 = plain int64 copy")
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
    if args.workload == "count":
        bench_count(args)
    else:
        bench_matrix(args)


if __name__ == "__main__":
    main()
