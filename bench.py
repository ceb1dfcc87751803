#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on its config[1]:
1 GiB synthetic MIXED corpus (log text + zero runs + incompressible random), 64 KiB independent blocks,
batched raw block compress + decompress on one B200 per rank (block ranges sharded over ranks, no collective).

A step = one pass of the hot path over the batch: compress every block, then decompress every block.
  value : uncompressed bytes / (compress + decompress device time), inputs resident in HBM (CUDA events)
  e2e   : the same through the host-pointer C-ABI calls (pinned host buffers, H2D + D2H inside the timed region)
  roofline : the compress kernel, algorithmic bytes (N read + C written) / its event-timed duration vs measured HBM peak
  cpu_baseline : the oracle (bit-exact C port of the reference) on host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lz4_block_roundtrip_throughput"
UNIT = "GB/s"            # 1e9 uncompressed bytes per second, compress + decompress of every block
BLOCK = 65536
WORKLOAD = "MIXED(seed=2) 1 GiB, 16384 x 64 KiB independent raw blocks, compress then decompress (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=1 << 30, help="uncompressed bytes per GPU")
    ap.add_argument("--cpu-sample-bytes", type=int, default=256 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_roundtrip(data, threads):
    """Oracle (bit-exact C port of compressBlock/decompressBlock) over 64 KiB blocks of `data` with `threads` host threads.
    Returns (seconds_compress, seconds_decompress)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    n = data.size
    off = np.arange(0, n, BLOCK, dtype=np.uint64)
    ln = np.minimum(BLOCK, n - off).astype(np.uint32)
    nb = len(off)
    parts = [(i * nb // threads, (i + 1) * nb // threads) for i in range(threads)]
    parts = [p for p in parts if p[1] > p[0]]

    def comp(p):
        return oracle.compress_blocks(data, off[p[0]:p[1]], ln[p[0]:p[1]])       # ctypes releases the GIL

    def decomp(args):
        (dst, doff, clen), p = args
        return oracle.decompress_blocks(dst, doff, clen, (off[p[0]:p[1]] - off[p[0]]), ln[p[0]:p[1]])

    with ThreadPoolExecutor(max_workers=len(parts)) as ex:
        t0 = time.perf_counter()
        comps = list(ex.map(comp, parts))
        t1 = time.perf_counter()
        outs = list(ex.map(decomp, list(zip(comps, parts))))
        t2 = time.perf_counter()
    for (o, olen, st), p in zip(outs, parts):
        assert not st.any()
    csize = sum(int(c[2].sum()) for c in comps)
    return t1 - t0, t2 - t1, csize


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here: the oracle port -- the reference is
    JavaScript and no JS engine exists in this image), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from divortio_lz4_b200 import corpus
    threads = os.cpu_count() or 1
    sample = min(args.bytes, args.cpu_sample_bytes)
    data = corpus.mixed(2, sample)
    for _ in range(args.warmup):
        cpu_roundtrip(data[:min(sample, 32 << 20)], threads)
    tc = td = 0.0
    for _ in range(args.steps):
        a, b, csize = cpu_roundtrip(data, threads)
        tc += a
        td += b
    val = args.steps * sample / (tc + td) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round((tc + td) / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "block_bytes": BLOCK, "note": "CPU arm: bounded sample of the same corpus"},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first %d MiB of the workload per step, %d threads over independent blocks" % (sample >> 20, threads)},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detail": {"compress_gbs": round(args.steps * sample / tc / 1e9, 4), "decompress_gbs": round(args.steps * sample / td / 1e9, 4),
                   "ratio": round(sample / csize, 4)},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import divortio_lz4_b200 as dl
    from divortio_lz4_b200 import corpus, device as dev

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    near = None
    if world > 1:                                 # several ranks share the host: stay on the GPU's own NUMA node
        from divortio_lz4_b200 import sharded
        near = sharded.bind_host_near(local)
    ctx = dl.Context(local)                       # raises if the CUDA extension or the device is missing

    # weak scaling: every rank owns `bytes` of the corpus = a contiguous block range of the N-GPU job
    n = args.bytes
    L = dl.lib()
    import ctypes as C
    pin_in = L.dlz4_pinned_alloc(n + 64)
    host = np.ctypeslib.as_array(C.cast(pin_in, C.POINTER(C.c_uint8)), shape=(n + 64,))
    corpus.mixed(2 + rank, n, out=host)
    nblk = (n + BLOCK - 1) // BLOCK
    stride = (dl.compress_bound(BLOCK) + 15) & ~15
    src = torch.empty(n + 64, dtype=torch.uint8, device=device)
    src[:n].copy_(torch.from_numpy(host[:n]), non_blocking=False)
    off, ln, _, coff = dev.uniform_blocks(n, BLOCK, device, stride)
    comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=device)
    clen = torch.zeros(nblk, dtype=torch.int32, device=device)
    out = torch.empty(n + 64, dtype=torch.uint8, device=device)
    olen = torch.zeros(nblk, dtype=torch.int32, device=device)
    status = torch.zeros(nblk, dtype=torch.uint8, device=device)
    cap = ln.clone()

    def step(ev=None):
        if ev:
            ev[0].record()
        dev.compress_blocks_dev(ctx, src, off, ln, BLOCK, comp, coff, clen)
        if ev:
            ev[1].record()
        dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, cap, olen, status)
        if ev:
            ev[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    assert int(status.max()) == 0 and torch.equal(out[:n], src[:n]), "round trip mismatch on device"
    csize = int(clen.sum())

    sampler = ClockSampler(local)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        step(evs[k])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    tc = sum(e[0].elapsed_time(e[1]) for e in evs) / 1e3
    td = sum(e[1].elapsed_time(e[2]) for e in evs) / 1e3
    t_dev = evs[0][0].elapsed_time(evs[-1][2]) / 1e3

    # ---- end to end through the host-pointer C ABI (the call the reference-side binding makes)
    e2e = None
    if not args.no_e2e:
        h_off = np.arange(nblk, dtype=np.uint64) * BLOCK
        h_len = np.minimum(BLOCK, n - h_off).astype(np.uint32)
        dst_bytes = nblk * dl.compress_bound(BLOCK)
        pin_c = L.dlz4_pinned_alloc(dst_bytes + 64)
        pin_o = L.dlz4_pinned_alloc(n + 64)
        h_clen = np.zeros(nblk, dtype=np.uint32)
        h_olen = np.zeros(nblk, dtype=np.uint32)
        h_st = np.zeros(nblk, dtype=np.uint8)

        def e2e_step():
            # packed output (dst_off = NULL): chunked pipeline, only real bytes cross PCIe
            s1 = L.dlz4_compress_blocks(ctx.handle, pin_in, n, h_off.ctypes.data, h_len.ctypes.data, nblk, None, 0, 0, None,
                                        pin_c, dst_bytes, None, h_clen.ctypes.data)
            ctx.check(s1)
            s2 = L.dlz4_decompress_blocks(ctx.handle, pin_c, dst_bytes, None, h_clen.ctypes.data, nblk, pin_o, n,
                                          h_off.ctypes.data, h_len.ctypes.data, None, 0, 0, h_olen.ctypes.data, h_st.ctypes.data)
            ctx.check(s2)

        e2e_step()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        esteps = max(2, min(args.steps, 3))
        for _ in range(esteps):
            e2e_step()
        barrier()
        te = time.perf_counter() - t0
        res = np.ctypeslib.as_array(C.cast(pin_o, C.POINTER(C.c_uint8)), shape=(n,))
        assert np.array_equal(res, host[:n]), "e2e round trip mismatch"
        cbytes = int(h_clen.astype(np.uint64).sum())
        # per step: compress moves n up + cbytes down, decompress moves cbytes up + n down (+ descriptors)
        e2e = {"t": te / esteps, "h2d": n + cbytes + nblk * (20 + 20), "d2h": cbytes + n + nblk * (4 + 5)}
        L.dlz4_pinned_free(pin_c)
        L.dlz4_pinned_free(pin_o)
    # ---- reported separately: the reference's DEFAULT frame (linked 4 MiB blocks, BASELINE configs[0] shape: 64 MiB log text)
    # through the frame C ABI with pinned host buffers -- segment-parallel compressor and jump decoder
    frames = None
    if rank == 0 and not args.no_e2e:
        from divortio_lz4_b200.api import FrameOpts
        fn = min(n, 64 << 20)
        pin_l = L.dlz4_pinned_alloc(fn + 64)
        hl = np.ctypeslib.as_array(C.cast(pin_l, C.POINTER(C.c_uint8)), shape=(fn + 64,))
        corpus.log(1, fn, out=hl)
        fcap = int(L.dlz4_frame_bound(fn))
        pin_f = L.dlz4_pinned_alloc(fcap + 64)
        pin_b = L.dlz4_pinned_alloc(fn + 64)
        opts = FrameOpts(4194304, 0, 0, 1, 0)
        flen, blen = C.c_uint64(0), C.c_uint64(0)
        tcs, tds = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.check(L.dlz4_frame_compress(ctx.handle, pin_l, fn, None, 0, C.byref(opts), pin_f, fcap, C.byref(flen)))
            t1 = time.perf_counter()
            ctx.check(L.dlz4_frame_decompress(ctx.handle, pin_f, flen.value, None, 0, 1, pin_b, fn, C.byref(blen)))
            t2 = time.perf_counter()
            tcs.append(t1 - t0)
            tds.append(t2 - t1)
        back = np.ctypeslib.as_array(C.cast(pin_b, C.POINTER(C.c_uint8)), shape=(fn,))
        assert blen.value == fn and np.array_equal(back, hl[:fn]), "linked frame round trip mismatch"
        segs, reruns, rounds = ctx.segment_stats
        frames = {"workload": "LOG(seed=1) %d MiB, one linked-block frame (blockIndependence=false, 4 MiB blocks), host to host" % (fn >> 20),
                  "compress_gbs": round(fn / min(tcs) / 1e9, 3), "decompress_gbs": round(fn / min(tds) / 1e9, 3),
                  "ratio": round(fn / flen.value, 4), "segments": segs, "segments_rerun": reruns, "rerun_rounds": rounds}
        for ptr in (pin_l, pin_f, pin_b):
            L.dlz4_pinned_free(ptr)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- max over ranks
    t = torch.tensor([t_dev, tc, td, e2e["t"] if e2e else 0.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cs = torch.tensor([csize], dtype=torch.int64, device=device)
        dist.all_reduce(cs)
        csize_all = int(cs.item())
    else:
        csize_all = csize
    t_dev, tc, td, te = [float(x) for x in t.tolist()]
    total = n * world

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = (n + csize) / (tc / args.steps) / 1e9
        line = {
            "metric": METRIC, "value": round(args.steps * total / (tc + td) / 1e9, 3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round((tc + td) / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "block_bytes": BLOCK, "bytes_per_gpu": n, "blocks_per_gpu": nblk,
                       "l2": "inputs (1 GiB) larger than L2 (126 MB), no flush needed", "sharding": "contiguous block range per rank, no collective",
                       "host_cpus_per_rank": (len(near) if near else None)},
            "clocks": sampler.summary(),
            "e2e": ({"value": round(total / te / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"]}
                    if e2e else None),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_compress_fresh16h", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(ach / peak, 5),
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, ncu --set full
                         # (profiles/r01c_bench_compress_full.txt: 23.48 GB + 3.36 GB).  17x the algorithmic bytes: 28 block
                         # chains per SM keep 4144 x (64 KiB block + 32 KiB table) = 400 MB in flight, three times the L2,
                         # so candidate reads and table sectors miss to HBM.  The 7-chain kernel it replaced moved
                         # 1.53 GB (1.0x) at half the throughput -- DESIGN.md 4.1 has the trade.
                         "traffic": 26842329000 if n == (1 << 30) else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": n + csize},
            "detail": {"compress_gbs": round(args.steps * total / tc / 1e9, 3), "decompress_gbs": round(args.steps * total / td / 1e9, 3),
                       "decompress_hbm_frac": round(((n + csize) / (td / args.steps) / 1e9) / peak, 5),
                       "ratio": round(total / csize_all, 4), "wall_s_timed_region": round(t_wall, 4), "verified_roundtrip": True,
                       "linked_frame": frames},
        }
        if not args.no_cpu_baseline and world >= 1:
            sample = min(n, args.cpu_sample_bytes)
            a, b, cs_cpu = cpu_roundtrip(host[:sample], 1)
            assert cs_cpu == int(clen[: (sample + BLOCK - 1) // BLOCK].sum()), "oracle and GPU compressed sizes differ"
            line["cpu_baseline"] = {"value": round(sample / (a + b) / 1e9, 4), "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "first %d MiB of the same corpus, single thread (analogue of the Node sync path)" % (sample >> 20),
                                    "compress_gbs": round(sample / a / 1e9, 4), "decompress_gbs": round(sample / b / 1e9, 4)}
        print(json.dumps(line), flush=True)
    L.dlz4_pinned_free(pin_in)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
