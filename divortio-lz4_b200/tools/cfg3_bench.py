"""BASELINE configs[3] kernels alone: 4 KiB JSON-like messages behind a 64 KiB dictionary with a primed table, device-resident.
Usage: [DLZ4_HY_ACTIVE=k] python divortio-lz4_b200/tools/cfg3_bench.py [messages]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, device as dev  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ctx = dl.Context(0)
d = torch.device("cuda", 0)
host = corpus.jsonmsgs(4, 0, count)
dic = corpus.json_dictionary(44)
primed = np.zeros(16384, dtype=np.int32)                 # the table a raw-API user carries out of the dictionary block
dl.compressBlock(dic, np.empty(dl.compress_bound(dic.size), dtype=np.uint8), 0, dic.size, primed, ctx=ctx)
src = torch.from_numpy(host).to(d)
stride = (dl.compress_bound(4096) + 15) & ~15
doff, dln, nblk, coff = dev.uniform_blocks(count * 4096, 4096, d, stride)
comp = torch.empty(count * stride + 64, dtype=torch.uint8, device=d)
clen = torch.zeros(count, dtype=torch.int32, device=d)
out = torch.empty(count * 4096 + 64, dtype=torch.uint8, device=d)
olen = torch.zeros(count, dtype=torch.int32, device=d)
st = torch.zeros(count, dtype=torch.uint8, device=d)
d_dic = torch.from_numpy(dic).to(d)
d_tab = torch.from_numpy(primed).to(d)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tc = td = 1e9
for rep in range(4):
    ev[0].record()
    dev.compress_blocks_dev(ctx, src, doff, dln, 4096, comp, coff, clen, prefix=d_dic, warm=dl.WARM_TABLE, init_table=d_tab)
    ev[1].record()
    dev.decompress_blocks_dev(ctx, comp, coff, clen, out, doff, dln, olen, st, dictionary=d_dic)
    ev[2].record()
    torch.cuda.synchronize()
    tc, td = min(tc, ev[0].elapsed_time(ev[1]) / 1e3), min(td, ev[1].elapsed_time(ev[2]) / 1e3)
assert int(st.max()) == 0 and torch.equal(out[:count * 4096], src)
tot = count * 4096
print("HY_ACTIVE=%s: %d messages, ratio %.3f | compress %.2f GB/s (%.2f M msgs/s) | decompress %.2f GB/s" %
      (os.environ.get("DLZ4_HY_ACTIVE", "7"), count, tot / int(clen.sum()), tot / tc / 1e9, count / tc / 1e6, tot / td / 1e9), flush=True)
