"""Where the multi-GPU tail goes: the slice push / collect kernels timed alone on ONE device
(every "rank" of a virtual world on cuda:0, so no NVLink and no waiting for a peer), next to the
single-GPU finalize they replace.  CUDA events, L2 flushed between launches."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kpal_b200 import _cabi     # noqa: E402

L = _cabi.load()
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bins = 4 ** k
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(3)
table = torch.from_numpy(rng.poisson(6.0, bins).astype(np.int32)).to(dev)
out64 = torch.empty(bins, dtype=torch.int64, device=dev)


def timed(fn):
    total = 0.0
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if _ >= 2:
            total += a.elapsed_time(b)
    return total / reps * 1000.0


print("k=%d finalize(balance) 1 GPU: %.1f us" % (k, timed(
    lambda: _cabi.check(L.kpal_dev_finalize_counts(table.data_ptr(), 32, k, 1, out64.data_ptr(), sp)))))
for world in (2, 4, 8):
    inbox_bytes = int(L.kpal_slice_inbox_bytes(k, world))
    inboxes = [torch.zeros(inbox_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in inboxes])
    begins = [int(L.kpal_slice_begin(k, r, world)) for r in range(world + 1)]
    piece = torch.empty(begins[1] - begins[0] + 64, dtype=torch.int64, device=dev)
    state = {"epoch": 0}

    def push_all():
        state["epoch"] += 1
        for r in range(world):
            _cabi.check(L.kpal_dev_slice_push(table.data_ptr(), 32, k, r, world, ptrs, state["epoch"], 0, sp))
            _cabi.check(L.kpal_dev_slice_signal(k, r, world, ptrs, state["epoch"], 0, sp))

    def push_one():
        # the other ranks' flags of this epoch are already there from push_all of the same epoch
        _cabi.check(L.kpal_dev_slice_push(table.data_ptr(), 32, k, 0, world, ptrs, state["epoch"], 0, sp))

    def collect():
        _cabi.check(L.kpal_dev_slice_collect(ptrs, k, 0, world, state["epoch"], -1, piece.data_ptr(), sp))

    push_all()
    torch.cuda.synchronize()
    t_push = timed(push_one)
    t_collect = timed(collect)
    print("k=%d world=%d: push (one rank, local stores) %.1f us, collect %.1f us" % (k, world, t_push, t_collect))
