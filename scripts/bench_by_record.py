#!/usr/bin/env python
"""BASELINE configs[2]: per-record counting of 100k records x 1 kbp at k=8
(dense int64 rows, 52.4 GB in total) -- device-resident timing in batches."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from kpal_b200 import _cabi  # noqa: E402
from oracle import c_oracle  # noqa: E402

K, N_REC, REC_LEN, BATCH = 8, 100_000, 1000, 20_000
L = _cabi.load()
rng = np.random.default_rng(3)
reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, N_REC * REC_LEN, dtype=np.uint8)]
reads[rng.random(reads.size, dtype=np.float32) < 0.001] = ord("N")
reads = reads.reshape(N_REC, REC_LEN)
t0 = time.perf_counter()
codes, valid, rec_starts, n_bases = _cabi.pack_sequences([r.tobytes() for r in reads])
pack_s = time.perf_counter() - t0
dev = torch.device("cuda", 0)
d_codes = torch.from_numpy(codes.view(np.int32)).to(dev)
d_valid = torch.from_numpy(valid.view(np.int32)).to(dev)
d_starts = torch.from_numpy(rec_starts.view(np.int64)).to(dev)
rows = torch.empty((BATCH, 4 ** K), dtype=torch.int64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
PATH = int(os.environ.get("KPAL_BY_RECORD_PATH", "0"))      # 1 = the zero-fill + RED rows, for comparison
_cabi.check(L.kpal_set_option(b"by_record_path", PATH))
for balance in (0, 1):
    times = []
    for rep in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for first in range(0, N_REC, BATCH):
            _cabi.check(L.kpal_dev_count_by_record(d_codes.data_ptr(), d_valid.data_ptr(),
                                                   d_starts.data_ptr(), first, BATCH, K, balance,
                                                   rows.data_ptr(), sp))
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    got = rows[-3:].cpu().numpy()
    ok = True
    for i in range(3):
        want = c_oracle.count_bytes(reads[N_REC - 3 + i].tobytes(), K)
        if balance:
            want = c_oracle.balance(want)
        ok &= bool(np.array_equal(got[i], want))
    ms = min(times)
    out_bytes = N_REC * 4 ** K * 8
    print(json.dumps({"bench": "by_record", "path": "red rows" if PATH else "shared-memory slabs", "k": K, "records": N_REC, "balance": balance,
                      "ms": ms, "records_per_s": N_REC / ms * 1e3,
                      "write_GBps": out_bytes / ms / 1e6, "parity_ok": ok,
                      "host_pack_s": pack_s}))

# ---- end to end through the host entry point (kpal_count_by_record: packed host stream ->
# dense int64 rows in host memory), on the first E2E_REC records (10.5 GB of rows at 20 k)
E2E_REC = int(os.environ.get("KPAL_BY_RECORD_E2E", "20000"))
if E2E_REC:
    out = np.empty((E2E_REC, 4 ** K), dtype=np.int64)
    out[:] = -1                                              # touch the pages before timing
    for narrow in (1, 0):
        _cabi.check(L.kpal_set_option(b"narrow_d2h", narrow))
        best = None
        for rep in range(2):
            t0 = time.perf_counter()
            _cabi.check(L.kpal_count_by_record(_cabi.ptr(codes), _cabi.ptr(valid), int(n_bases), _cabi.ptr(rec_starts),
                                               0, E2E_REC, K, 0, out.ctypes.data))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        ok = all(np.array_equal(out[i], c_oracle.count_bytes(reads[i].tobytes(), K)) for i in (0, 1, E2E_REC // 2, E2E_REC - 1))
        print(json.dumps({"bench": "by_record_e2e", "k": K, "records": E2E_REC,
                          "rows": "uint16 over PCIe, widened by host threads" if narrow else "int64 over PCIe",
                          "s": best, "records_per_s": E2E_REC / best, "host_row_GBps": out.nbytes / best / 1e9,
                          "parity_ok": bool(ok)}))
    _cabi.check(L.kpal_set_option(b"narrow_d2h", 1))
