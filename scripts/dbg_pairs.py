import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import test_gpu_count as t
from oracle import c_oracle
mode = sys.argv[1] if len(sys.argv) > 1 else "AC"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
text = t._composition_bytes(5, n, mode)
want = c_oracle.count_bytes(text, 12, threads=4)
t._set_option("count_path", 2)
if len(sys.argv) > 3:
    t._set_option("pair_flush_every", int(sys.argv[3]))
L = t._cabi.load()
seqs = bytes(text).split(b"\n")
codes, valid, _, n_bases = t._cabi.pack_sequences(seqs)
bins = 4 ** 12
d_codes = L.kpal_dev_alloc(codes.nbytes); d_valid = L.kpal_dev_alloc(valid.nbytes)
d_table = L.kpal_dev_alloc(bins * 4); d_counts = L.kpal_dev_alloc(bins * 8)
t._cabi.check(L.kpal_memcpy_h2d(d_codes, t._cabi.ptr(codes), codes.nbytes, None))
t._cabi.check(L.kpal_memcpy_h2d(d_valid, t._cabi.ptr(valid), valid.nbytes, None))
t._cabi.check(L.kpal_dev_count_packed_fresh(d_codes, d_valid, n_bases, 12, d_table, 32, None))
t._cabi.check(L.kpal_dev_finalize_counts(d_table, 32, 12, 0, d_counts, None))
out = np.empty(bins, dtype=np.int64)
t._cabi.check(L.kpal_memcpy_d2h(t._cabi.ptr(out), d_counts, out.nbytes, None))
t._cabi.check(L.kpal_stream_sync(None))
print(mode, n, "equal:", np.array_equal(out, want), "sum", out.sum(), want.sum())
direct = np.full(bins, -1, dtype=np.int64)
t._cabi.check(L.kpal_dev_table_to_host(d_table, 32, 12, 0, t._cabi.ptr(direct), None))
print("table_to_host equal:", np.array_equal(direct, want), "diff bins", int((direct != want).sum()))
