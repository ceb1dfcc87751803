"""Host-side sharding logic on CPU: world_size-2 (and 3) gloo groups over a shared host mapping, and ranks emulated by
threads.  The oracle stands in for the GPU kernels (OracleBackend below) -- these tests check the partition, the length scan,
the in-place placement, the multi-frame cut, the checksum relay and the sharded decode, not the kernels: the frames every
world size writes must equal the single-process oracle frames byte for byte, and decode back."""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend(object):
    """CPU stand-in with the GpuBackend interface (test infrastructure: oracle/ is never imported by the product)."""

    def __init__(self):
        import oracle
        self.o = oracle
        self.body = b""
        self.res_in = self.res_out = b""

    def body_compress(self, piece, block_size, block_checksum):
        f = self.o.compress_buffer(piece, None, block_size, True, False, False, None, block_checksum)
        self.body = f[7:-4]              # header without content size / dict id is 7 bytes; EndMark is 4
        self.res_in = bytes(piece.tobytes())
        return len(self.body)

    def body_fetch(self, dst):
        dst[:] = np.frombuffer(self.body, dtype=np.uint8)

    # the relay state: the oracle has no stateful xxh32, so the stand-in carries the bytes seen so far in a side table keyed by
    # the state's total (what matters here is the ORDER of the relay and that the digest lands in the frame)
    _seen = {}

    def state_new(self):
        from divortio_lz4_b200 import api
        s = api.Xxh32State()
        s.seed = 0
        s.total = 0
        s.v[0] = threading.get_ident() & 0x7FFFFFFF        # a relay id that travels with the state
        s.v[1] = os.getpid() & 0x7FFFFFFF
        return s

    def _key(self, state):
        return (int(state.v[0]), int(state.v[1]))

    def _path(self, state):
        return os.path.join("/tmp", "dlz4_relay_%d_%d.bin" % self._key(state))

    def _append(self, state, data):
        with open(self._path(state), "ab") as fh:
            fh.write(data)
        state.total += len(data)

    def state_update_input(self, state):
        self._append(state, self.res_in)

    def state_update_output(self, state):
        self._append(state, self.res_out)

    def state_digest(self, state):
        p = self._path(state)
        data = open(p, "rb").read() if os.path.exists(p) else b""
        if os.path.exists(p):
            os.unlink(p)
        assert len(data) == state.total
        return self.o.xxh32(data)

    def decompress_range(self, frame, first, count, out, verify_block_checksums=False):
        from divortio_lz4_b200 import api
        info = api.frame_info(frame)
        f = bytes(frame.tobytes())
        # walk the block table on the host
        pos = 7 + (8 if info.has_content_size else 0) + (4 if info.has_dict_id else 0)
        blocks = []
        while True:
            bs = int.from_bytes(f[pos:pos + 4], "little")
            pos += 4
            if bs == 0:
                break
            blocks.append((pos, bs & 0x7FFFFFFF, bs >> 31))
            pos += (bs & 0x7FFFFFFF) + (4 if info.has_block_checksum else 0)
        lens = np.zeros(max(1, count), dtype=np.uint32)
        w = 0
        try:
            for k, (bp, bl, stored) in enumerate(blocks[first:first + count]):
                if stored:
                    piece = f[bp:bp + bl]
                else:
                    hist = bytes(out[:w].tobytes()) if not info.block_independence else b""
                    tmp = np.zeros(int(info.block_max_size), dtype=np.uint8)
                    n = self.o.decompress_block(np.frombuffer(f, dtype=np.uint8), bp, bl, tmp, 0,
                                                np.frombuffer(hist[-65536:], dtype=np.uint8) if hist else None)
                    piece = tmp[:n].tobytes()
                out[w:w + len(piece)] = np.frombuffer(piece, dtype=np.uint8)
                lens[k] = len(piece)
                w += len(piece)
        except self.o.OracleError as e:
            return int(e.code), 0, lens[:count]
        self.res_out = bytes(out[:w].tobytes())
        return 0, w, lens[:count]


def _expected_frames(oracle, data, bs, cc, bc, frame_max):
    from divortio_lz4_b200 import sharded
    return b"".join(oracle.compress_buffer(data[lo:hi], None, bs, True, cc, True, None, bc)
                    for lo, hi in sharded.frame_spans(data.size, frame_max))


def _worker(rank, world, port, n, bs, cc, bc, frame_max, tag, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from divortio_lz4_b200 import corpus, sharded
        data = corpus.mixed(21, n)
        cap = n + n // 200 + 4096
        # the host frame and the decode target are ONE mapping shared by the ranks (created by rank 0)
        if rank == 0:
            out = sharded.SharedBuffer("dlz4_test_out_%s" % tag, cap, True, register=False)
            back = sharded.SharedBuffer("dlz4_test_back_%s" % tag, n + 64, True, register=False)
        dist.barrier()
        if rank != 0:
            out = sharded.SharedBuffer("dlz4_test_out_%s" % tag, cap, False, register=False)
            back = sharded.SharedBuffer("dlz4_test_back_%s" % tag, n + 64, False, register=False)
        comm = sharded.DistComm()
        be = OracleBackend()
        tm = {}
        total = sharded.compress_sharded(data, out.array, bs, cc, True, bc, comm=comm, backend=be, frame_max=frame_max, timings=tm)
        got = sharded.decompress_sharded(out.array[:total], back.array, True, False, comm=comm, backend=be)
        if rank == 0:
            want = _expected_frames(oracle, data, bs, cc, bc, frame_max)
            q.put((bytes(out.array[:total].tobytes()) == want, total, got == n and bytes(back.array[:n].tobytes()) == data.tobytes(),
                   "blocks" in tm and "checksum" in tm))
        dist.barrier()
        out.close()
        back.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,bs,cc,bc,frame_max", [
    (2, 1000000, 65536, True, True, 1 << 40),
    (2, 65536 * 3, 65536, False, False, 1 << 40),
    (3, 700001, 262144, True, False, 1 << 40),
    (2, 1, 65536, True, True, 1 << 40),
    (2, 1500000, 65536, True, True, 8 * 65536),          # several frames (the >= 2 GiB rule, scaled down)
])
def test_sharded_frames_equal_single_process_frames(world, n, bs, cc, bc, frame_max):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n + world) % 2000
    tag = "%d_%d_%d" % (os.getpid(), n, world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bs, cc, bc, frame_max, tag, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, length, roundtrip, timed = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and roundtrip and timed and length > 0


def test_thread_ranks_and_short_inner_block_fallback():
    """Ranks as threads of one process (ThreadComm); a frame with a short inner block makes the sharded decode fall back to one
    rank decoding it in order; a linked frame is decoded whole by its owner rank."""
    import struct
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from divortio_lz4_b200 import corpus, sharded
    data = corpus.log(5, 6 * 65536 + 123)
    raw = data.tobytes()
    # frame A: regular independent frame; frame B: hand-made frame with a short inner block; frame C: linked frame
    fa = oracle.compress_buffer(data, None, 65536, True, True, True)
    pieces = [raw[:30000], raw[30000:30000 + 65536], raw[30000 + 65536:30000 + 2 * 65536]]
    desc = bytes([0x60, 0x40])
    fb = bytearray(struct.pack("<I", 0x184D2204) + desc + bytes([(oracle.xxh32(desc) >> 8) & 0xFF]))
    for p in pieces:
        c = oracle.compress_block_bytes(np.frombuffer(p, dtype=np.uint8))
        fb += struct.pack("<I", len(c)) + c
    fb += struct.pack("<I", 0)
    fc = oracle.compress_buffer(data, None, 65536, False, True, True)
    blob = np.frombuffer(fa + bytes(fb) + fc, dtype=np.uint8)
    want = raw + b"".join(pieces) + raw
    world = 3
    comms = sharded.ThreadComm.group(world)
    out = np.zeros(len(want) + 65536 * 4, dtype=np.uint8)
    res = [None] * world
    err = []

    def run(r):
        try:
            res[r] = sharded.decompress_sharded(blob, out, True, False, comm=comms[r], backend=OracleBackend())
        except Exception as e:  # noqa: BLE001
            err.append(e)
            comms[r].s.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not err, err
    assert res == [len(want)] * world
    assert bytes(out[:len(want)].tobytes()) == want


def test_plan_covers_every_block_once():
    from divortio_lz4_b200 import sharded
    for total in (0, 1, 65536, 65537, 10 ** 6, 8 * 2 ** 20 + 5):
        for world in (1, 2, 4, 8):
            for frame_max in (1 << 40, 4 * 65536):
                bs, frames = sharded.plan(total, 65536, world, frame_max)
                assert frames[0][0] == 0 and frames[-1][1] == total
                for fa, fb in zip(frames, frames[1:]):
                    assert fa[1] == fb[0]
                for lo, hi, nblocks, ranges in frames:
                    assert ranges[0][2] == lo and ranges[-1][3] == hi
                    for a, b in zip(ranges, ranges[1:]):
                        assert a[3] == b[2] and a[0] + a[1] == b[0]
                    assert sum(r[1] for r in ranges) == nblocks


def test_frame_spans_for_inputs_of_2_gib_and_more():
    """DESIGN: an input above 2 GiB - 4 MiB is written as ceil(n / 1 GiB) frames (each frame call takes len|0 < 2 GiB like
    bufferCompress.js:127; more frames = more content-checksum chains side by side); the content size field of every frame is
    that frame's own length."""
    from divortio_lz4_b200 import sharded
    spans = sharded.frame_spans(8 << 30)
    assert len(spans) == 8 and spans[0] == (0, sharded.FRAME_SPLIT) and spans[-1][1] == 8 << 30
    assert sharded.frame_spans(sharded.FRAME_MAX) == [(0, sharded.FRAME_MAX)]            # up to the limit: the reference's one frame
    assert all((hi - lo) % (4 << 20) == 0 for lo, hi in spans[:-1]) and all(hi - lo < (2 << 30) for lo, hi in spans)
    assert sharded.frame_spans(0) == [(0, 0)] and sharded.frame_spans(5) == [(0, 5)]
    # a content size above 32 bits would be written in full (the header writer takes u64)
    h = sharded.frame_header((1 << 32) + 5, 4194304, True, False, True, False)
    assert h[6:14] == ((1 << 32) + 5).to_bytes(8, "little")


def test_header_matches_oracle_header():
    import oracle
    from divortio_lz4_b200 import sharded
    for size in (True, False):
        for cc in (False, True):
            for bc in (False, True):
                h = sharded.frame_header(12345, 65536, True, cc, size, bc)
                want = oracle.compress_buffer(bytes(12345), None, 65536, True, cc, size, None, bc)
                assert want.startswith(h)


def test_bind_host_near_never_widens_the_affinity_and_survives_missing_nvml():
    """One rank per GPU pins its host threads to the GPU's NUMA node when NVML can say which CPUs that is; without a GPU or
    NVML it must leave the process alone and say so."""
    import importlib
    sharded = importlib.import_module("divortio_lz4_b200.sharded")
    before = os.sched_getaffinity(0)
    got = sharded.bind_host_near(0)
    after = os.sched_getaffinity(0)
    if got is None:
        assert after == before
    else:
        assert after == got and got <= before
        os.sched_setaffinity(0, before)
