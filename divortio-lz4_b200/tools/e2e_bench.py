"""End-to-end timing of the host-pointer batch calls (pinned buffers), with the plain PCIe copy rates beside them."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus  # noqa: E402

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
BLOCK = 65536
ctx = dl.Context(0)
L = dl.lib()
pin_in = L.dlz4_pinned_alloc(n + 64)
host = np.ctypeslib.as_array(C.cast(pin_in, C.POINTER(C.c_uint8)), shape=(n + 64,))
corpus.mixed(2, n, out=host)
nblk = (n + BLOCK - 1) // BLOCK
h_off = np.arange(nblk, dtype=np.uint64) * BLOCK
h_len = np.minimum(BLOCK, n - h_off).astype(np.uint32)
dst_bytes = nblk * dl.compress_bound(BLOCK)
pin_c = L.dlz4_pinned_alloc(dst_bytes + 64)
pin_o = L.dlz4_pinned_alloc(n + 64)
h_clen = np.zeros(nblk, dtype=np.uint32)
h_olen = np.zeros(nblk, dtype=np.uint32)
h_st = np.zeros(nblk, dtype=np.uint8)
# raw PCIe rates
t = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.from_numpy(host[:n])
for name, fn in (("H2D", lambda: t.copy_(hp, non_blocking=True)), ("D2H", lambda: hp.copy_(t, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%s pinned %.1f GB/s" % (name, n / dt / 1e9))
corpus.mixed(2, n, out=host)
for rep in range(4):
    t0 = time.perf_counter()
    ctx.check(L.dlz4_compress_blocks(ctx.handle, pin_in, n, h_off.ctypes.data, h_len.ctypes.data, nblk, None, 0, 0, None, pin_c, dst_bytes,
                                     None, h_clen.ctypes.data))
    t1 = time.perf_counter()
    kc = ctx.last_kernel_ms
    ctx.check(L.dlz4_decompress_blocks(ctx.handle, pin_c, dst_bytes, None, h_clen.ctypes.data, nblk, pin_o, n, h_off.ctypes.data,
                                       h_len.ctypes.data, None, 0, 0, h_olen.ctypes.data, h_st.ctypes.data))
    t2 = time.perf_counter()
    print("compress e2e %.1f ms (%.1f GB/s; kernel section %.1f ms) | decompress e2e %.1f ms (%.1f GB/s; kernel section %.1f ms) | round trip %.2f GB/s"
          % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9, kc, (t2 - t1) * 1e3, n / (t2 - t1) / 1e9, ctx.last_kernel_ms, n / (t2 - t0) / 1e9))
res = np.ctypeslib.as_array(C.cast(pin_o, C.POINTER(C.c_uint8)), shape=(n,))
assert np.array_equal(res, host[:n])
