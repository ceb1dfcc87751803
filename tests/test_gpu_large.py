"""Full-size GPU checks at BASELINE.json sizes: the oracle still finishes in seconds at 1 GiB (C, ~0.7 GB/s), so sizes and
per-block checksums of every compressed block are compared against it, plus size-independent properties (round trip on
the device, checksum of checksums).  Run with -m gpu."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import divortio_lz4_b200 as dl
    from divortio_lz4_b200 import corpus, device as dev
    return dl, corpus, dev, torch, dl.default_context()


def test_config2_1gib_mixed_64k_blocks(env):
    """BASELINE configs[1]: MIXED(2, 2^30), 16384 x 64 KiB independent blocks, device-resident batch."""
    dl, corpus, dev, torch, ctx = env
    n, block = 1 << 30, 65536
    host = corpus.mixed(2, n)
    d = torch.device("cuda", ctx.device)
    src = torch.from_numpy(host).to(d)
    stride = (dl.compress_bound(block) + 15) & ~15
    off, ln, nblk, coff = dev.uniform_blocks(n, block, d, stride)
    comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=d)
    clen = torch.zeros(nblk, dtype=torch.int32, device=d)
    dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
    out = torch.empty(n + 64, dtype=torch.uint8, device=d)
    olen = torch.zeros(nblk, dtype=torch.int32, device=d)
    st = torch.zeros(nblk, dtype=torch.uint8, device=d)
    dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, ln, olen, st)
    hashes = torch.zeros(nblk, dtype=torch.int32, device=d)
    dev.xxh32_batch_dev(ctx, comp, coff, clen, hashes)
    torch.cuda.synchronize()
    # properties
    assert int(st.max()) == 0 and torch.equal(olen, ln) and torch.equal(out[:n], src[:n])
    # oracle on the full corpus
    h_off = np.arange(nblk, dtype=np.uint64) * block
    h_len = np.full(nblk, block, dtype=np.uint32)
    odst, odoff, oclen = oracle.compress_blocks(host, h_off, h_len)
    assert np.array_equal(clen.cpu().numpy().astype(np.uint32), oclen)
    ohash = oracle.xxh32_batch(odst, odoff, oclen)
    assert np.array_equal(hashes.cpu().numpy().view(np.uint32), ohash)          # every block's bytes, via its checksum
    # checksum of checksums (one number to quote)
    assert oracle.xxh32(ohash.tobytes()) == oracle.xxh32(hashes.cpu().numpy().tobytes())


def test_config3_shape_4m_blocks_block_and_content_checksums(env):
    """BASELINE configs[2] shape at one call's worth (the reference takes len|0 < 2 GiB per call): 4 MiB blocks,
    blockChecksum + contentChecksum, frame bytes vs oracle; sharded by block range over 4 emulated ranks."""
    dl, corpus, dev, torch, ctx = env
    from divortio_lz4_b200 import sharded
    n = 512 * 1024 * 1024 + 4321
    host = corpus.mixed(3, n)
    want = oracle.compress_buffer(host, None, 4194304, True, True, True, None, True)
    got = dl.compressBuffer(host, None, 4194304, True, True, True, None, True)
    assert len(got) == len(want) and got == want
    assert dl.decompressBuffer(got, None, True, True) == host.tobytes()
    # the same frame from 4 block-range shards placed in one host frame (ranks emulated by threads, one context each), and
    # the sharded decode of it
    back = np.zeros(n + 64, dtype=np.uint8)
    out = np.zeros(len(want) + 4096, dtype=np.uint8)
    total, got_n = _run_thread_ranks(dl, 4, host, out, back, 4194304, True, True)
    assert bytes(out[:total].tobytes()) == want
    assert got_n == n and np.array_equal(back[:n], host)


def _run_thread_ranks(dl, world, host, out, back, bs, cc, bc, frame_max=None, modes=None):
    """compress_sharded + decompress_sharded with `world` ranks emulated by threads (own dlz4_ctx each) on this GPU."""
    import threading
    from divortio_lz4_b200 import sharded
    comms = sharded.ThreadComm.group(world)
    res = [None] * world
    err = []

    def run(r):
        try:
            be = sharded.GpuBackend(dl.Context(0))
            kw = {} if frame_max is None else {"frame_max": frame_max}
            tm1, tm2 = {}, {}
            t = sharded.compress_sharded(host, out, bs, cc, True, bc, comm=comms[r], backend=be, timings=tm1, **kw)
            g = sharded.decompress_sharded(out[:t], back, True, bc, comm=comms[r], backend=be, timings=tm2)
            res[r] = (t, g)
            if modes is not None:
                modes.append((tm1["checksum_mode"], tm2["checksum_mode"]))
        except Exception as e:  # noqa: BLE001
            err.append(e)
            comms[r].s.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not err, err
    assert all(r == res[0] for r in res)
    return res[0]


def test_sharded_multi_frame_and_ragged_world_sizes(env):
    """The >= 2 GiB rule scaled down: frame_max = 24 MiB cuts a 100 MiB input into 5 frames; 3 ranks (ragged ranges), 64 KiB and
    4 MiB blocks.  The concatenated frames equal the oracle's per-span frames and decode back (sharded decode, relay checksum)."""
    dl, corpus, dev, torch, ctx = env
    from divortio_lz4_b200 import sharded
    n = 100 * 1024 * 1024 + 777
    host = corpus.mixed(13, n)
    for bs in (65536, 4194304):
        fm = 24 << 20
        want = b"".join(oracle.compress_buffer(host[lo:hi], None, bs, True, True, True, None, True)
                        for lo, hi in sharded.frame_spans(n, fm))
        out = np.zeros(len(want) + 4096, dtype=np.uint8)
        back = np.zeros(n + 64, dtype=np.uint8)
        total, got_n = _run_thread_ranks(dl, 3, host, out, back, bs, True, True, frame_max=fm)
        assert bytes(out[:total].tobytes()) == want, bs
        assert got_n == n and np.array_equal(back[:n], host)
        # any frame reader decodes the concatenation as one stream
        assert dl.decompressFrames(out[:total])[0] == host.tobytes()


def test_sharded_frames_in_page_locked_buffers_hash_one_chain_per_frame(env):
    """With page-locked host buffers the content checksum of every frame is its own asynchronous chain (dlz4_xxh32_async, read
    in place over PCIe) instead of the rank-to-rank relay; same bytes either way.  Also the slots on their own: device memory,
    page-locked memory, pageable memory refused, a corrupted checksum detected."""
    import ctypes as C
    dl, corpus, dev, torch, ctx = env
    from divortio_lz4_b200 import api, sharded
    n = 70 * 1024 * 1024 + 333
    fm = 16 << 20
    host = corpus.mixed(17, n)
    want = b"".join(oracle.compress_buffer(host[lo:hi], None, 65536, True, True, True, None, True) for lo, hi in sharded.frame_spans(n, fm))
    out = np.zeros(len(want) + 4096, dtype=np.uint8)
    back = np.zeros(n + 64, dtype=np.uint8)
    bufs = (host, out, back)
    assert all(api.host_register(b) for b in bufs)
    try:
        modes = []
        total, got_n = _run_thread_ranks(dl, 2, host, out, back, 65536, True, True, frame_max=fm, modes=modes)
        assert bytes(out[:total].tobytes()) == want
        assert got_n == n and np.array_equal(back[:n], host)
        assert all(m[0].startswith("one chain per frame") and m[1].startswith("one chain per frame") for m in modes), modes
        # the slots directly
        L = dl.lib()
        h = C.c_uint32()
        for slot, lo, ln in ((0, 0, n), (5, 12345, 1 << 20), (31, 7, 15), (2, 100, 0)):
            assert L.dlz4_xxh32_async(ctx.handle, slot, api._ptr(host[lo:]), ln, 0) == 0
        for slot, lo, ln in ((0, 0, n), (5, 12345, 1 << 20), (31, 7, 15), (2, 100, 0)):
            assert L.dlz4_xxh32_wait(ctx.handle, slot, C.byref(h)) == 0
            assert h.value == oracle.xxh32(host[lo:lo + ln]), (slot, lo, ln)
        d = torch.from_numpy(host[:1 << 22]).to(torch.device("cuda", ctx.device))
        assert L.dlz4_xxh32_async(ctx.handle, 1, d.data_ptr() + 3, (1 << 22) - 3, 9) == 0
        assert L.dlz4_xxh32_wait(ctx.handle, 1, C.byref(h)) == 0 and h.value == oracle.xxh32(host[3:1 << 22], 9)
        assert L.dlz4_xxh32_async(ctx.handle, 1, api._ptr(np.zeros(1 << 20, dtype=np.uint8)), 1 << 20, 0) == 20      # pageable: DLZ4_E_INVALID_ARG
        # a wrong content checksum is still found (every rank raises)
        bad = out[:total].copy()
        bad[total - 1] ^= 0x55
        assert api.host_register(bad)
        try:
            with pytest.raises(Exception) as ei:
                sharded.decompress_sharded(bad, back, True, False, comm=sharded.LocalComm(), backend=sharded.GpuBackend(ctx))
            assert "Content Checksum Error" in str(ei.value)
        finally:
            api.host_unregister(bad)
    finally:
        for b in bufs:
            api.host_unregister(b)


def _two_gpu_worker(rank, world, port, tag, q):
    import os
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import divortio_lz4_b200 as dl
        from divortio_lz4_b200 import corpus, sharded
        import oracle as orc
        n = 192 * 1024 * 1024 + 4321
        cap = n + n // 100 + 65536
        names = ["in", "out", "back"]
        sizes = [n, cap, n + 64]
        bufs = {}
        if rank == 0:
            for nm, sz in zip(names, sizes):
                bufs[nm] = sharded.SharedBuffer("dlz4_t2_%s_%s" % (nm, tag), sz, True)
            bufs["in"].array[:] = corpus.mixed(31, n)
        dist.barrier()
        if rank != 0:
            for nm, sz in zip(names, sizes):
                bufs[nm] = sharded.SharedBuffer("dlz4_t2_%s_%s" % (nm, tag), sz, False)
        be = sharded.GpuBackend(dl.Context(rank))
        comm = sharded.DistComm()
        tm = {}
        total = sharded.compress_sharded(bufs["in"].array, bufs["out"].array, 4194304, True, True, True, comm=comm, backend=be, timings=tm)
        got = sharded.decompress_sharded(bufs["out"].array[:total], bufs["back"].array, True, True, comm=comm, backend=be)
        if rank == 0:
            want = orc.compress_buffer(bufs["in"].array, None, 4194304, True, True, True, None, True)
            q.put((bytes(bufs["out"].array[:total].tobytes()) == want, got == n,
                   bool(np.array_equal(bufs["back"].array[:n], bufs["in"].array)), all(b.registered for b in bufs.values())))
        dist.barrier()
        for b in bufs.values():
            b.close()
    finally:
        dist.destroy_process_group()


def test_sharded_on_two_real_gpus(env):
    """compress_sharded / decompress_sharded with one process per GPU over a page-locked shared host mapping (NCCL carries the
    integers and the checksum state only).  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
    dl, corpus, dev, torch, ctx = env
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import torch.multiprocessing as mp
    c = mp.get_context("spawn")
    q = c.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [c.Process(target=_two_gpu_worker, args=(r, 2, port, str(os.getpid()), q)) for r in range(2)]
    for p in procs:
        p.start()
    same, n_ok, back_ok, pinned = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same and n_ok and back_ok and pinned


def test_config4_shape_small_messages_with_dictionary(env):
    """BASELINE configs[3] at 65536 messages: 4 KiB JSON messages, 64 KiB dictionary prefix, primed table."""
    dl, corpus, dev, torch, ctx = env
    nmsg = 65536
    msgs = corpus.jsonmsgs(4, 0, nmsg)
    dic = corpus.json_dictionary(44)
    off = np.arange(nmsg, dtype=np.uint64) * 4096
    ln = np.full(nmsg, 4096, dtype=np.uint32)
    primed = oracle.new_table()
    oracle.compress_block(dic, 0, dic.size, primed)
    dst, doff, clen = dl.compress_blocks(msgs, off, ln, prefix=dic, warm=dl.WARM_TABLE, init_table=primed)
    odst, odoff, oclen = oracle.compress_blocks_prefix(dic, primed, msgs, off, ln)
    assert np.array_equal(clen, oclen)
    assert np.array_equal(dl.xxh32_batch(dst, doff, clen), oracle.xxh32_batch(odst, odoff, oclen))
    out, olen, status = dl.decompress_blocks(dst, doff, clen, off, ln, dictionary=dic)
    assert not status.any() and np.array_equal(out[:msgs.size], msgs)


def test_config5_shape_decode_reference_and_cli_frames_linked_and_independent(env):
    """BASELINE configs[4] shape: decode-only of oracle (reference-format) and liblz4 (CLI-format) frames."""
    import lz4f
    dl, corpus, dev, torch, ctx = env
    n = 256 * 1024 * 1024
    host = corpus.mixed(5, n)
    raw = host.tobytes()
    frames = [oracle.compress_buffer(host, None, 4194304, True, True, True),
              oracle.compress_buffer(host, None, 65536, True, False, True)]
    if lz4f.available():
        frames += [lz4f.compress_frame(raw, 7, False, True, False, False), lz4f.compress_frame(raw, 4, False, True, False, True)]
    for f in frames:
        assert dl.decompressBuffer(f, None, True, True) == raw
    # linked-block frames are the serial chain: smaller input
    small = host[:24 * 1024 * 1024]
    linked = [oracle.compress_buffer(small, None, 4194304, False, True, True)]
    if lz4f.available():
        linked.append(lz4f.compress_frame(small.tobytes(), 7, True, True, False, False))
    for f in linked:
        assert dl.decompressBuffer(f) == small.tobytes()


def test_reference_default_frame_1gib_linked_blocks_byte_exact(env):
    """The reference's default options on 1 GiB (blockIndependence=false, 4 MiB blocks): one serial chain in the reference, 2048
    speculative segments here.  The frame must equal the oracle's byte for byte (checked through length and xxh32 of the whole
    frame, and through the first difference if they differ), and the jump decoder must return the input."""
    dl, corpus, dev, torch, ctx = env
    n = 1 << 30
    data = corpus.mixed(7, n)
    f = dl.compressBuffer(data, None, 4194304, False, False, True, ctx=ctx)
    segs, reruns, rounds = ctx.segment_stats
    want = oracle.compress_buffer(data, None, 4194304, False, False, True)
    assert len(f) == len(want)
    if f != want:
        a, b = np.frombuffer(f, dtype=np.uint8), np.frombuffer(want, dtype=np.uint8)
        raise AssertionError("first difference at frame byte %d" % int(np.nonzero(a != b)[0][0]))
    assert segs >= 2048 and reruns <= 64, (segs, reruns, rounds)
    back = dl.decompressBuffer(f, None, True, False, ctx=ctx)
    assert len(back) == n and np.array_equal(np.frombuffer(back, dtype=np.uint8), data)


def test_frames_compressed_while_the_input_is_still_arriving(env):
    """Frames of 64 MiB and more are parsed under a chunked host-to-device copy (16 MiB chunks + landed flags).
    Odd sizes, a dictionary (chunk and block boundaries no longer coincide), linked and independent large blocks, pageable
    and page-locked callers' buffers: byte-exact against the oracle, and back."""
    import ctypes as C
    dl, corpus, dev, torch, ctx = env
    n = 80 * 1024 * 1024 + 12345
    data = corpus.mixed(9, n)
    dic = corpus.log(10, 100000)
    L = dl.lib()
    pin = L.dlz4_pinned_alloc(n + 64)
    pinned = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(n + 64,))[:n]
    pinned[:] = data
    try:
        for src in (data, pinned):
            for bs, indep, d in ((4194304, False, None), (1048576, False, dic), (4194304, True, None), (262144, False, None)):
                f = dl.compressBuffer(src, d, bs, indep, False, True, ctx=ctx)
                want = oracle.compress_buffer(data, d, bs, indep, False, True)
                assert len(f) == len(want) and f == want, (bs, indep, d is not None)
                segs, reruns, rounds = ctx.segment_stats
                assert segs >= 64
        back = dl.decompressBuffer(f, None, True, False, ctx=ctx)
        assert len(back) == n and np.array_equal(np.frombuffer(back, dtype=np.uint8), data)
    finally:
        L.dlz4_pinned_free(pin)
