"""Frame API (compressBuffer / decompressBuffer) through the C ABI with pinned host buffers: wall time of the call, device
time of its kernel section and the segment statistics of the segment-parallel engine.  (The CPU baseline lives in bench.py;
product tools never touch oracle/.)
Usage: python divortio-lz4_b200/tools/frame_bench.py [kind=log|mixed] [MiB] [--only=CASE] [--once]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus  # noqa: E402
from divortio_lz4_b200.api import FrameInfo, FrameOpts  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "log"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--only=")]
reps = 1 if "--once" in sys.argv else 3
n = mib << 20
ctx = dl.Context(0)
L = dl.lib()


def pinned(nbytes):
    p = L.dlz4_pinned_alloc(nbytes + 64)
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes + 64,))


p_in, h_in = pinned(n)
(corpus.log if kind == "log" else corpus.mixed)(1, n, out=h_in)
cap = int(L.dlz4_frame_bound(n))
p_f, h_f = pinned(cap)
p_o, h_o = pinned(n)
print("%s %d MiB, pinned host buffers; GB/s of uncompressed bytes" % (kind, mib))
cases = ((4194304, False, False, False), (4194304, True, False, False), (65536, False, False, False),
         (65536, True, False, False), (4194304, True, True, True), (4194304, False, True, False))
for ci, (bs, indep, cc, bc) in enumerate(cases):
    if only and ci not in only:
        continue
    opts = FrameOpts(bs, int(indep), int(cc), 1, int(bc))
    flen = C.c_uint64(0)
    best_c = best_d = 1e9
    kc = kd = 0.0
    for rep in range(reps):
        t0 = time.perf_counter()
        ctx.check(L.dlz4_frame_compress(ctx.handle, p_in, n, None, 0, C.byref(opts), p_f, cap, C.byref(flen)))
        t1 = time.perf_counter()
        if t1 - t0 < best_c:
            best_c, kc = t1 - t0, ctx.last_kernel_ms
        stats = ctx.segment_stats
        olen = C.c_uint64(0)
        t2 = time.perf_counter()
        ctx.check(L.dlz4_frame_decompress(ctx.handle, p_f, flen.value, None, 0, 1, p_o, n, C.byref(olen)))
        t3 = time.perf_counter()
        if t3 - t2 < best_d:
            best_d, kd = t3 - t2, ctx.last_kernel_ms
    assert olen.value == n and np.array_equal(h_o[:n], h_in[:n])
    line = ("block %7d %-11s cc=%d bc=%d ratio %5.3f | compress %6.2f GB/s (call %7.1f ms, kernels %7.1f ms; %d segments, %d re-run in %d rounds)"
            " | decompress %6.2f GB/s (call %6.1f ms, kernels %6.1f ms)" %
            (bs, "independent" if indep else "linked", cc, bc, n / flen.value, n / best_c / 1e9, best_c * 1e3, kc, stats[0], stats[1], stats[2],
             n / best_d / 1e9, best_d * 1e3, kd))
    print(line, flush=True)
