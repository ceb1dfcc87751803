"""End-to-end round trips of configs[1] through the host-pointer C ABI on page-locked buffers: the two calls of a step back to
back on one context, and consecutive steps overlapped on two contexts (compress(k+1) beside decompress(k)).
Usage: [CUDA_DEVICE_MAX_CONNECTIONS=32] python divortio-lz4_b200/tools/e2e_pipe.py [MiB] [steps]"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus  # noqa: E402


def pinned(L, n):
    p = L.dlz4_pinned_alloc(n + 64)
    return p, np.ctypeslib.as_array((C.c_uint8 * (n + 64)).from_address(p))


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n, B = mib << 20, 65536
    L = dl.lib()
    ctx = dl.Context(0)
    ctx2 = dl.Context(0)
    pin_in, host = pinned(L, n)
    corpus.mixed(2, n, out=host)
    nblk = (n + B - 1) // B
    h_off = np.arange(nblk, dtype=np.uint64) * B
    h_len = np.minimum(B, n - h_off).astype(np.uint32)
    dst_bytes = nblk * dl.compress_bound(B)
    h_olen = np.zeros(nblk, dtype=np.uint32)
    h_st = np.zeros(nblk, dtype=np.uint8)
    slots = []
    for _ in range(2):
        pc, _a = pinned(L, dst_bytes)
        po, res = pinned(L, n)
        slots.append((pc, np.zeros(nblk, dtype=np.uint32), po, res))

    def comp(c, k):
        pc, hc, _, _ = slots[k & 1]
        c.check(L.dlz4_compress_blocks(c.handle, pin_in, n, h_off.ctypes.data, h_len.ctypes.data, nblk, None, 0, 0, None,
                                       pc, dst_bytes, None, hc.ctypes.data))

    def deco(c, k):
        pc, hc, po, _ = slots[k & 1]
        c.check(L.dlz4_decompress_blocks(c.handle, pc, dst_bytes, None, hc.ctypes.data, nblk, po, n, h_off.ctypes.data,
                                         h_len.ctypes.data, None, 0, 0, h_olen.ctypes.data, h_st.ctypes.data))

    def serial(count):
        tc = td = 0.0
        for k in range(count):
            t0 = time.perf_counter(); comp(ctx, k); t1 = time.perf_counter(); deco(ctx, k); t2 = time.perf_counter()
            tc += t1 - t0; td += t2 - t1
        return tc / count, td / count

    def pipelined(count):
        free = [threading.Semaphore(1), threading.Semaphore(1)]
        full = [threading.Semaphore(0), threading.Semaphore(0)]

        def producer():
            for k in range(count):
                free[k & 1].acquire(); comp(ctx, k); full[k & 1].release()
        th = threading.Thread(target=producer)
        th.start()
        for k in range(count):
            full[k & 1].acquire(); deco(ctx2, k); free[k & 1].release()
        th.join()

    serial(2)
    t0 = time.perf_counter(); tc, td = serial(steps); ts = (time.perf_counter() - t0) / steps
    pipelined(2)
    t0 = time.perf_counter(); pipelined(steps); tp = (time.perf_counter() - t0) / steps
    for s in slots:
        assert np.array_equal(s[3][:n], host[:n])
    print("%d MiB x %d steps, max_connections=%s: serial %.2f ms/step (compress %.2f + decompress %.2f) = %.2f GB/s | "
          "overlapped %.2f ms/step = %.2f GB/s" % (mib, steps, os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default"),
                                                  ts * 1e3, tc * 1e3, td * 1e3, n / ts / 1e9, tp * 1e3, n / tp / 1e9), flush=True)


if __name__ == "__main__":
    main()
