"""Throughput of the streaming classes (stream.LZ4Encoder / LZ4Decoder) fed in fixed-size chunks.
Usage: python divortio-lz4_b200/tools/stream_bench.py [MiB total] [MiB per add]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, stream  # noqa: E402

total = (int(sys.argv[1]) if len(sys.argv) > 1 else 512) << 20
step = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
dl.default_context()
data = corpus.mixed(5, total)
for bs, indep, cc in ((4194304, False, False), (65536, False, False), (65536, True, False), (4194304, False, True)):
    for rep in range(2):
        enc = stream.LZ4Encoder(bs, indep, cc)
        t0 = time.perf_counter()
        pieces = []
        for p in range(0, total, step):
            pieces += enc.add(data[p:p + step])
        pieces += enc.finish()
        t1 = time.perf_counter()
        frame = b"".join(bytes(x) for x in pieces)
        dec = stream.LZ4Decoder()
        t2 = time.perf_counter()
        n = 0
        for p in range(0, len(frame), step // 2):
            n += sum(len(c) for c in dec.update(frame[p:p + step // 2]))
        t3 = time.perf_counter()
        assert n == total
    print("block %7d %-11s contentChecksum=%d, %d MiB in adds of %d MiB: encode %6.2f GB/s | decode %6.2f GB/s (Python host, pageable buffers)" %
          (bs, "independent" if indep else "linked", cc, total >> 20, step >> 20, total / (t1 - t0) / 1e9, total / (t3 - t2) / 1e9), flush=True)
