"""Throughput of the other BASELINE.json configs on one GPU (device-resident kernels where the API allows, CUDA events).
Usage: python divortio-lz4_b200/tools/config_bench.py [MiB for configs 3/5] [messages for config 4]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, device as dev  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nmsg = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
ctx = dl.Context(0)
d = torch.device("cuda", 0)
s = torch.cuda.Stream()


def timed(fn, reps=3):
    with torch.cuda.stream(s):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        s.synchronize()
    return e0.elapsed_time(e1) / reps


def blocks_case(name, host, block, prefix=None, warm=dl.WARM_NONE, table=None):
    n = host.size
    src = torch.from_numpy(host).to(d)
    stride = (dl.compress_bound(block) + 15) & ~15
    off, ln, nblk, coff = dev.uniform_blocks(n, block, d, stride)
    comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=d)
    clen = torch.zeros(nblk, dtype=torch.int32, device=d)
    out = torch.empty(n + 64, dtype=torch.uint8, device=d)
    olen = torch.zeros(nblk, dtype=torch.int32, device=d)
    st = torch.zeros(nblk, dtype=torch.uint8, device=d)
    pf = torch.from_numpy(prefix).to(d) if prefix is not None else None
    tb = torch.from_numpy(table).to(d) if table is not None else None
    tc = timed(lambda: dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen, prefix=pf, warm=warm, init_table=tb))
    td = timed(lambda: dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, ln, olen, st, dictionary=pf))
    torch.cuda.synchronize()
    assert int(st.max()) == 0 and torch.equal(out[:n], src[:n])
    c = int(clen.sum())
    print("%-44s %6d blocks of %7d B ratio %6.3f | compress %7.2f GB/s | decompress %8.2f GB/s" %
          (name, nblk, block, n / c, n / tc / 1e6, n / td / 1e6), flush=True)


n = mib << 20
host = corpus.mixed(3, n)
blocks_case("config3 kernels: MIXED, 4 MiB blocks", host, 4194304)
blocks_case("           MIXED, 1 MiB blocks", host, 1048576)
blocks_case("           MIXED, 256 KiB blocks", host, 262144)
blocks_case("config2 kernels: MIXED, 64 KiB blocks", host, 65536)

# per-block checksums and the serial content checksum (reported separately)
src = torch.from_numpy(host).to(d)
off, ln, nblk = dev.uniform_blocks(n, 4194304, d)
hs = torch.zeros(nblk, dtype=torch.int32, device=d)
t = timed(lambda: dev.xxh32_batch_dev(ctx, src, off, ln, hs))
print("xxh32_batch over %d x 4 MiB                      %8.2f GB/s" % (nblk, n / t / 1e6))
off, ln, nblk = dev.uniform_blocks(n, 65536, d)
hs = torch.zeros(nblk, dtype=torch.int32, device=d)
t = timed(lambda: dev.xxh32_batch_dev(ctx, src, off, ln, hs))
print("xxh32_batch over %d x 64 KiB                  %8.2f GB/s" % (nblk, n / t / 1e6))
one = torch.zeros(1, dtype=torch.int32, device=d)
m = min(n, 256 << 20)
t = timed(lambda: dev.xxh32_stream_dev(ctx, src[:m], one), reps=1)
print("xxh32_stream (serial chain, one warp), %d MiB    %8.2f GB/s" % (m >> 20, m / t / 1e6))

# (the frame API is timed by tools/frame_bench.py: C ABI with pinned buffers, no Python copies)

# config 4: small messages with a shared dictionary prefix
msgs = corpus.jsonmsgs(4, 0, nmsg)
dic = corpus.json_dictionary(44)
import ctypes as C  # noqa: E402
blocks_case("config4: JSON 4 KiB msgs, no prefix", msgs, 4096)
blocks_case("config4: + 64 KiB prefix, warm=none", msgs, 4096, prefix=dic)
blocks_case("config4: + 64 KiB prefix, warm=jenkins", msgs, 4096, prefix=dic, warm=dl.WARM_JENKINS)
# warm=primed: the table the kernel itself leaves behind after compressing the dictionary (raw API, table is in/out)
tab = np.zeros(16384, dtype=np.int32)
scratch = np.zeros(dl.compress_bound(dic.size), dtype=np.uint8)
dl.compressBlock(dic, scratch, 0, dic.size, tab, 0, ctx=ctx)
blocks_case("config4: + 64 KiB prefix, warm=primed", msgs, 4096, prefix=dic, warm=dl.WARM_TABLE, table=tab)
