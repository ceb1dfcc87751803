"""Block-range sharding of the frame path over the GPUs of one box -- compress AND decompress.

One process per GPU, one `dlz4_ctx` each.  Independent blocks shard with no data-path collective:

  compress    an input above FRAME_MAX bytes is cut into frames of FRAME_SPLIT bytes (one frame call takes < 2 GiB like the reference,
              bufferCompress.js:127 `len|0`; an 8 GiB input is 8 frames, which every LZ4 frame reader decodes as one stream of
              concatenated frames).  Inside a frame, rank r owns the contiguous block range dlz4_shard_range(nblocks, world, r),
              compresses it on its GPU into the frame-body bytes of that range ([u32 size|stored][data][u32 xxh32]*, exactly
              what the single-GPU path writes for those blocks), the ranks exchange the LENGTHS of their bodies (one integer
              each), scan them, and every rank copies its body device -> host straight to its place in the shared host frame.
              Rank 0 writes header and EndMark.
  decompress  the host walks the frames; inside an independent-block frame rank r decodes its block range into
              out[frame_base + first * blockMaxSize ...) (bufferDecompress.js:133-192 with the block loop cut by range); a frame
              whose inner blocks are not full, and every linked-block frame, is decoded whole by one rank (frames rotate over
              the ranks).  Ranks exchange decoded lengths and statuses (integers) per frame.

The whole-stream content checksum is one serial chain PER FRAME (xxhash32.js:34-57, four non-associative accumulators):
"replicas only" inside a frame, but the chains of different frames are independent.  When the bytes to hash sit in
page-locked host memory (the shared mappings of a multi-rank job, dlz4_pinned_alloc buffers) frame k's chain is started
asynchronously on rank k mod world (dlz4_xxh32_async: a single-warp kernel on its own stream that reads the host bytes in
place over PCIe) and runs beside the block work of that and the following frames; the digests are collected at the end.
Otherwise (pageable buffers, more frames in flight than slots) it runs as a RELAY over the bytes already resident on each
GPU -- rank r receives the 16-byte accumulator state from rank r-1, runs the stripe loop over its resident piece, passes the
state on (dlz4_xxh32_update_resident) -- so nothing is uploaded twice.  Either way it is timed separately from the block work.

`data`, `frames` and `out` are numpy uint8 arrays every rank can address: the same page-locked shared mapping in a multi-process
run (`SharedBuffer`), plain arrays in a single process.  Communication goes through a small `comm` object (torch.distributed
when initialised -- gloo on CPU, NCCL on the GPUs; `ThreadComm` for ranks emulated by threads of one process).
"""
import ctypes as C
import mmap
import os
import queue
import threading
import time

import numpy as np

from . import api

MAGIC = b"\x04\x22\x4D\x18"
BLOCK_SIZES = {4: 65536, 5: 262144, 6: 1048576, 7: 4194304}
FRAME_MAX = (2 << 30) - (4 << 20)      # content bytes per frame: < 2 GiB (bufferCompress.js:127), whole blocks of every size
FRAME_SPLIT = 1 << 30                  # an input above FRAME_MAX is written as frames of this many bytes: the content checksum is
                                       # one serial chain per frame (2.4 GB/s), so more frames = more chains side by side, while
                                       # the large-block compressor loses efficiency on short calls (8 GiB, one GPU, compress /
                                       # decompress GB/s: 2 GiB frames 7.8 / 6.5, 1 GiB 9.4 / 8.8, 512 MiB 5.5 / 7.0)
STATE_BYTES = C.sizeof(api.Xxh32State)


def block_id_for(max_block_size):
    """getBlockId (src/buffer/bufferCompress.js:77-82)."""
    if not max_block_size or max_block_size <= 65536:
        return 4
    if max_block_size <= 262144:
        return 5
    if max_block_size <= 1048576:
        return 6
    return 7


def frame_spans(total_len, frame_max=FRAME_MAX):
    """Byte ranges of the frames an input of total_len bytes is written as (an empty input is one empty frame).  Up to
    frame_max bytes: one frame, the reference's own output.  More (the reference cannot take such an input at all): frames of
    FRAME_SPLIT bytes -- of frame_max bytes when the caller sets a limit of its own."""
    if total_len == 0:
        return [(0, 0)]
    if total_len <= frame_max:
        return [(0, total_len)]
    step = FRAME_SPLIT if frame_max == FRAME_MAX else frame_max
    return [(lo, min(lo + step, total_len)) for lo in range(0, total_len, step)]


shard_range = api.shard_range          # dlz4_shard_range: block i -> rank floor(i * world / nblocks), contiguous ranges


def plan(total_len, max_block_size, world, frame_max=FRAME_MAX):
    """(block size, [(frame_lo, frame_hi, nblocks, [(first, count, byte_lo, byte_hi) per rank]) per frame])."""
    bs = BLOCK_SIZES[block_id_for(max_block_size)]
    frames = []
    for lo, hi in frame_spans(total_len, frame_max):
        nblocks = (hi - lo + bs - 1) // bs
        ranks = []
        for r in range(world):
            first, count = shard_range(nblocks, world, r)
            ranks.append((first, count, min(lo + first * bs, hi), min(lo + (first + count) * bs, hi)))
        frames.append((lo, hi, nblocks, ranks))
    return bs, frames


def frame_header(content_len, max_block_size, block_independence, content_checksum, add_content_size, block_checksum,
                 dict_id=None):
    """Header bytes (bufferCompress.js:147-178) from the library's one writer, dlz4_frame_header."""
    opts = api.FrameOpts(int(max_block_size or 0) & 0xFFFFFFFF, int(bool(block_independence)), int(bool(content_checksum)),
                         int(bool(add_content_size)), int(bool(block_checksum)))
    buf = (C.c_uint8 * 19)()
    n = api.lib().dlz4_frame_header(C.byref(opts), int(content_len), 0 if dict_id is None else 1,
                                    0 if dict_id is None else int(dict_id) & 0xFFFFFFFF, buf)
    return bytes(buf[:n])


# ------------------------------------------------------------------------------------------------ communication
class LocalComm(object):
    rank, world = 0, 1

    def all_gather_ints(self, vals):
        return [list(vals)]

    def send_bytes(self, b, dst):
        raise RuntimeError("no peer")

    def recv_bytes(self, n, src):
        raise RuntimeError("no peer")

    def barrier(self):
        pass


class DistComm(object):
    """torch.distributed: integers and the 40-byte checksum state only -- never block data."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def all_gather_ints(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    def send_bytes(self, b, dst):
        t = self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).to(self.device)
        self.dist.send(t, dst)

    def recv_bytes(self, n, src):
        t = self.torch.empty(n, dtype=self.torch.uint8, device=self.device)
        self.dist.recv(t, src)
        return bytes(t.cpu().numpy().tobytes())

    def barrier(self):
        self.dist.barrier()


class ThreadComm(object):
    """Ranks emulated by the threads of one process (tests; several contexts on one GPU)."""

    class _Shared(object):
        def __init__(self, world):
            self.world = world
            self.bar = threading.Barrier(world)
            self.slots = [None] * world
            self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def group(cls, world):
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def all_gather_ints(self, vals):
        self.s.slots[self.rank] = list(vals)
        self.s.bar.wait()
        out = [list(v) for v in self.s.slots]
        self.s.bar.wait()
        return out

    def send_bytes(self, b, dst):
        self.s.q[(self.rank, dst)].put(bytes(b))

    def recv_bytes(self, n, src):
        return self.s.q[(src, self.rank)].get(timeout=300)

    def barrier(self):
        self.s.bar.wait()


def default_comm():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return DistComm()
    except ImportError:
        pass
    return LocalComm()


# ------------------------------------------------------------------------------------------------ shared host buffers
class SharedBuffer(object):
    """A host buffer every rank of the box maps: a file in /dev/shm (or `directory`), mmap'ed by each process and -- on a
    GPU box -- page-locked in each with cudaHostRegister so that H2D/D2H run at the pinned PCIe rate without any
    inter-process copy ("D2H into the host frame at those offsets")."""

    def __init__(self, name, nbytes, create, directory=None, register=True):
        directory = directory or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
        self.path = os.path.join(directory, name)
        self.nbytes = int(nbytes)
        self.created = create
        if create:
            with open(self.path, "wb") as fh:
                fh.truncate(max(1, self.nbytes))
        self.fh = open(self.path, "r+b")
        self.mm = mmap.mmap(self.fh.fileno(), max(1, self.nbytes))
        self.array = np.frombuffer(self.mm, dtype=np.uint8, count=self.nbytes)
        self.registered = False
        if register and self.nbytes:
            self.registered = api.host_register(self.array)

    def close(self, unlink=None):
        if self.registered:
            api.host_unregister(self.array)
            self.registered = False
        self.array = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.fh.close()
        if self.created if unlink is None else unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


# ------------------------------------------------------------------------------------------------ backends
class GpuBackend(object):
    """This rank's GPU through the C ABI."""

    def __init__(self, ctx=None):
        self.ctx = ctx or api.default_context()

    def body_compress(self, piece, block_size, block_checksum):
        n = C.c_uint64()
        self.ctx.check(api.lib().dlz4_frame_body_compress(self.ctx.handle, api._ptr(piece), piece.size, int(block_size),
                                                          int(bool(block_checksum)), C.byref(n)))
        return int(n.value)

    def body_fetch(self, dst):
        self.ctx.check(api.lib().dlz4_frame_body_fetch(self.ctx.handle, api._ptr(dst), dst.size))

    def state_new(self):
        s = api.Xxh32State()
        api.lib().dlz4_xxh32_reset(C.byref(s), 0)
        return s

    def state_update_input(self, state):
        self.ctx.check(api.lib().dlz4_xxh32_update_resident(self.ctx.handle, C.byref(state), 0))

    def state_update_output(self, state):
        self.ctx.check(api.lib().dlz4_xxh32_update_resident(self.ctx.handle, C.byref(state), 1))

    def state_digest(self, state):
        return int(api.lib().dlz4_xxh32_digest(C.byref(state)))

    SUM_SLOTS = 32

    def sum_async(self, slot, piece):
        """Starts xxh32(piece) in `slot` (dlz4_xxh32_async); False when the bytes are not page-locked (caller falls back)."""
        st = api.lib().dlz4_xxh32_async(self.ctx.handle, int(slot), api._ptr(piece), piece.size, 0)
        if st == api.E_CUDA:
            self.ctx.check(st)
        return st == 0

    def sum_wait(self, slot):
        h = C.c_uint32()
        self.ctx.check(api.lib().dlz4_xxh32_wait(self.ctx.handle, int(slot), C.byref(h)))
        return int(h.value)

    def decompress_range(self, frame, first, count, out, verify_block_checksums=False):
        """-> (status, decoded bytes, per-block decoded lengths)."""
        n = C.c_uint64()
        lens = np.zeros(max(1, count), dtype=np.uint32)
        st = api.lib().dlz4_frame_decompress_range(self.ctx.handle, api._ptr(frame), frame.size, int(first), int(count), None, 0,
                                                   2 if verify_block_checksums else 0, api._ptr(out), out.size, C.byref(n),
                                                   api._ptr(lens))
        if st == api.E_CUDA:
            self.ctx.check(st)
        return int(st), int(n.value), lens[:count]


def _state_bytes(state):
    return bytes(C.string_at(C.addressof(state), STATE_BYTES))


def _state_from(b):
    s = api.Xxh32State()
    C.memmove(C.addressof(s), b, STATE_BYTES)
    return s


# ------------------------------------------------------------------------------------------------ compress
def compress_sharded(data, out, max_block_size=4194304, content_checksum=False, add_content_size=True, block_checksum=False,
                     comm=None, backend=None, frame_max=FRAME_MAX, timings=None):
    """Independent-block frame(s) of `data` written into `out` by comm.world ranks; every rank calls this with the same
    arguments and gets the total number of bytes written.  timings (dict, optional) receives this rank's seconds in
    'blocks' (compress + length exchange + copy out) and 'checksum' (the serial relay)."""
    comm = comm or default_comm()
    backend = backend or GpuBackend()
    data = api.ensureBuffer(data)
    rank, world = comm.rank, comm.world
    bs, frames = plan(data.size, max_block_size, world, frame_max)
    pos = 0
    t_blocks = t_sum = 0.0
    # content checksums: frame k's chain on rank k mod world, started now, collected behind the block loop
    async_sum = {}                                 # frame index -> slot (on its owner)
    use_async = 0
    if content_checksum and hasattr(backend, "sum_async") and len(frames) <= backend.SUM_SLOTS * world:
        ok = 1
        for k, (lo, hi, _, _) in enumerate(frames):
            if k % world == rank:
                if ok and backend.sum_async(k // world, data[lo:hi]):
                    async_sum[k] = k // world
                else:
                    ok = 0
        use_async = int(all(v[0] for v in comm.all_gather_ints([ok])))     # all or nothing: every rank takes the same path
    sum_at = {}                                    # frame index -> position of its checksum in `out`
    for k, (lo, hi, nblocks, ranks) in enumerate(frames):
        t0 = time.perf_counter()
        first, count, blo, bhi = ranks[rank]
        body_len = backend.body_compress(data[blo:bhi], bs, block_checksum) if count else 0
        lens = [v[0] for v in comm.all_gather_ints([body_len])]
        header = frame_header(hi - lo, max_block_size, True, content_checksum, add_content_size, block_checksum)
        body_pos = pos + len(header)
        total_body = sum(lens)
        end = body_pos + total_body + 4 + (4 if content_checksum else 0)
        if end > out.size:
            raise api.LZ4Error(1, "LZ4: Output Buffer Too Small")
        mine = body_pos + sum(lens[:rank])
        if body_len:
            backend.body_fetch(out[mine:mine + body_len])
        if rank == 0:
            out[pos:body_pos] = np.frombuffer(header, dtype=np.uint8)
            out[body_pos + total_body:body_pos + total_body + 4] = 0              # EndMark (bufferCompress.js:244)
        t1 = time.perf_counter()
        if content_checksum and use_async:
            sum_at[k] = end - 4
        elif content_checksum:                                                    # :248-252, serial: relay in rank order
            state = backend.state_new() if rank == 0 else _state_from(comm.recv_bytes(STATE_BYTES, rank - 1))
            if count:
                backend.state_update_input(state)
            if rank + 1 < world:
                comm.send_bytes(_state_bytes(state), rank + 1)
            else:
                h = backend.state_digest(state)
                out[end - 4:end] = np.frombuffer(int(h).to_bytes(4, "little"), dtype=np.uint8)
        t2 = time.perf_counter()
        t_blocks += t1 - t0
        t_sum += t2 - t1
        pos = end
    t2 = time.perf_counter()
    if use_async:
        for k, slot in async_sum.items():
            h = backend.sum_wait(slot)
            out[sum_at[k]:sum_at[k] + 4] = np.frombuffer(int(h).to_bytes(4, "little"), dtype=np.uint8)
    elif async_sum:
        for slot in async_sum.values():            # started but not used (another rank could not): let them finish
            backend.sum_wait(slot)
    t_sum += time.perf_counter() - t2
    comm.barrier()
    if timings is not None:
        timings["blocks"] = t_blocks
        timings["checksum"] = t_sum
        timings["checksum_mode"] = "one chain per frame, asynchronous, from page-locked host memory" if use_async else "relay over resident bytes"
    return pos


# ------------------------------------------------------------------------------------------------ decompress
def list_frames(frames):
    """[(offset, FrameInfo)] of the LZ4 frames in a buffer of concatenated frames; skippable frames are skipped."""
    f = api.ensureBuffer(frames)
    out = []
    pos = 0
    while pos < f.size:
        if pos + 4 > f.size:
            raise api.LZ4Error(api.E_BAD_MAGIC, "LZ4: Invalid Magic Number")
        magic = int.from_bytes(f[pos:pos + 4].tobytes(), "little")
        if (magic & 0xFFFFFFF0) == 0x184D2A50:
            if pos + 8 > f.size:
                raise api.LZ4Error(2, "LZ4: Malformed Input")
            pos += 8 + int.from_bytes(f[pos + 4:pos + 8].tobytes(), "little")
            continue
        info = api.frame_info(f[pos:])
        out.append((pos, info))
        pos += int(info.frame_bytes)
    return out


def decompress_sharded(frames, out, verify_checksum=True, verify_block_checksums=False, comm=None, backend=None, timings=None):
    """Decodes every frame of `frames` into `out` with comm.world ranks; returns the decoded length (same on every rank).
    Raises LZ4Error with the reference's message on the first failing frame (every rank raises)."""
    comm = comm or default_comm()
    backend = backend or GpuBackend()
    f = api.ensureBuffer(frames)
    rank, world = comm.rank, comm.world
    base = 0
    t_blocks = t_sum = 0.0
    flist = list_frames(f)
    # content checksums of different frames are independent chains: asynchronous, on the frame's owner, over the decoded bytes
    # where they land in (page-locked) `out`; verified behind the frame loop
    use_async = int(verify_checksum and hasattr(backend, "sum_async") and len(flist) <= backend.SUM_SLOTS * world
                    and out.size >= 16 and backend.sum_async(0, out[:16]))
    if use_async:
        backend.sum_wait(0)
    use_async = int(all(v[0] for v in comm.all_gather_ints([use_async])))
    pending = []                                   # (slot, expected digest) on this rank
    for k, (fpos, info) in enumerate(flist):
        t0 = time.perf_counter()
        fr = f[fpos:fpos + int(info.frame_bytes)]
        n, B = int(info.nblocks), int(info.block_max_size)
        owner = k % world
        sharded = bool(info.block_independence) and n >= world and world > 1
        st, got, short = 0, 0, 0
        if sharded:
            first, count = shard_range(n, world, rank)
            o = out[base + first * B:]
            if count:
                st, got, lens = backend.decompress_range(fr, first, count, o, verify_block_checksums)
                inner = lens[:-1] if first + count == n else lens
                short = int(st == 0 and bool((inner != B).any()))
        elif rank == owner:
            st, got, _ = backend.decompress_range(fr, 0, n, out[base:], verify_block_checksums)
        res = comm.all_gather_ints([st, got, short])
        if sharded and any(r[2] for r in res) and not any(r[0] for r in res):
            # an inner block was short: the i * blockMaxSize placement does not hold -> one rank decodes the frame in order
            sharded = False
            st, got = 0, 0
            if rank == owner:
                st, got, _ = backend.decompress_range(fr, 0, n, out[base:], verify_block_checksums)
            res = comm.all_gather_ints([st, got, 0])
        bad = [r[0] for r in res if r[0]]
        if bad:
            raise api.LZ4Error(bad[0], api.lib().dlz4_strerror(bad[0]).decode())
        total = sum(r[1] for r in res)
        t1 = time.perf_counter()
        if info.has_content_checksum and verify_checksum and use_async:
            if rank == owner:
                if not backend.sum_async(k // world, out[base:base + total]):
                    raise RuntimeError("dlz4_xxh32_async refused a buffer it accepted before")
                pending.append((k // world, int.from_bytes(fr[fr.size - 4:].tobytes(), "little")))
        elif info.has_content_checksum and verify_checksum:                       # bufferDecompress.js:213-217, serial relay
            order = list(range(world)) if sharded else [owner]
            ok = 1
            if rank in order:
                i = order.index(rank)
                state = backend.state_new() if i == 0 else _state_from(comm.recv_bytes(STATE_BYTES, order[i - 1]))
                backend.state_update_output(state)
                if i + 1 < len(order):
                    comm.send_bytes(_state_bytes(state), order[i + 1])
                else:
                    want = int.from_bytes(fr[fr.size - 4:].tobytes(), "little")
                    ok = int(backend.state_digest(state) == want)
            if not all(r[0] for r in comm.all_gather_ints([ok])):
                raise api.LZ4Error(api.E_CONTENT_CHECKSUM, "LZ4: Content Checksum Error")
        t2 = time.perf_counter()
        t_blocks += t1 - t0
        t_sum += t2 - t1
        base += total
    t2 = time.perf_counter()
    if use_async:
        ok = int(all(backend.sum_wait(slot) == want for slot, want in pending))
        if not all(r[0] for r in comm.all_gather_ints([ok])):
            raise api.LZ4Error(api.E_CONTENT_CHECKSUM, "LZ4: Content Checksum Error")
    t_sum += time.perf_counter() - t2
    comm.barrier()
    if timings is not None:
        timings["blocks"] = t_blocks
        timings["checksum"] = t_sum
        timings["checksum_mode"] = "one chain per frame, asynchronous, from page-locked host memory" if use_async else "relay over resident bytes"
    return base


def bind_host_near(device_index):
    """One process per GPU: keep this process's threads -- and therefore the pinned staging buffers it allocates next -- on
    the CPUs of the GPU's own NUMA node (NVML's ideal CPU affinity), so that the ranks' host copies do not cross the socket
    interconnect.  Returns the CPU set applied, or None when NVML / the affinity call is unavailable (nothing changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
