#!/usr/bin/env python
"""BASELINE configs[2] end to end, as a user runs it: `kpal count --by-record -k 8 reads.fa out.k`
on 100 000 synthetic records of 1 kbp -- FASTA text from disk in, HDF5 profile file with 100 000
gzip datasets and their six statistics on disk out.  Prints one JSON line; checks a sample of the
written profiles (counts and attributes) against the oracle / NumPy."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpal_b200 import kmer, h5lite  # noqa: E402
from oracle import c_oracle  # noqa: E402

K = 8
N_REC = int(os.environ.get("KPAL_BY_RECORD_N", "100000"))
REC_LEN = 1000
rng = np.random.default_rng(3)
reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, N_REC * REC_LEN, dtype=np.uint8)]
reads[rng.random(reads.size, dtype=np.float32) < 0.001] = ord("N")
reads = reads.reshape(N_REC, REC_LEN)
tmp = tempfile.mkdtemp(prefix="kpal_by_record_")
fasta_path, out_path = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "out.k")
with open(fasta_path, "wb") as f:
    for i in range(N_REC):
        f.write(b">r%07d\n" % (i + 1))
        for c in range(0, REC_LEN, 70):
            f.write(reads[i, c:c + 70].tobytes() + b"\n")
os.environ.setdefault("KPAL_B200_H5LITE", "1")
t0 = time.perf_counter()
kmer.main(["count", "--by-record", "-k", str(K), fasta_path, out_path])
wall = time.perf_counter() - t0
size = os.path.getsize(out_path)
# parity on a sample of the written file
handle = h5lite.File(out_path, "r")
names = handle["profiles"].keys()
ok = len(names) == N_REC
for i in sorted(set([0, 1, N_REC // 3, N_REC // 2, N_REC - 2, N_REC - 1])):
    dataset = handle["profiles/r%07d" % (i + 1)]
    counts = dataset[:]
    want = c_oracle.count_bytes(reads[i].tobytes(), K)
    ok &= bool(np.array_equal(counts, want))
    attrs = dict(dataset.attrs.items())
    ok &= (attrs["length"] == K and attrs["total"] == want.sum() and attrs["non_zero"] == np.count_nonzero(want)
           and attrs["mean"] == want.mean() and attrs["median"] == np.median(want) and attrs["std"] == want.std())
handle.close()
print(json.dumps({"bench": "kpal count --by-record (file to file)", "k": K, "records": N_REC, "record_len": REC_LEN,
                  "wall_s": wall, "records_per_s": N_REC / wall, "file_bytes": size, "host_threads": os.cpu_count(),
                  "includes": "FASTA read + host pack, GPU rows, uint16 D2H + widen, row statistics, deflate, HDF5 write + close",
                  "parity_ok": bool(ok)}))
os.remove(fasta_path)
os.remove(out_path)
os.rmdir(tmp)
