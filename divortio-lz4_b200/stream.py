"""Host-side mirror of the reference's streaming codec: `LZ4Encoder` (src/shared/lz4Encode.js:95-332) and
`LZ4Decoder` (src/shared/lz4Decode.js:52-306), same constructor arguments, same methods, same chunk-by-chunk results -- with the
GPU block path beneath them.  What changes is the batching: an `add()` that completes several blocks compresses them in ONE call
(independent blocks: one batch; linked blocks: one chain through the segment-parallel engine, hash table carried in and out),
and an `update()` decodes every complete block it holds in one frame call (jump decoder for linked data).  The emitted pieces
are byte-identical to flushing block by block.
"""
import ctypes as C

import numpy as np

from . import api

MAX_WINDOW_SIZE = 65536
BLOCK_MAX_SIZES = {4: 65536, 5: 262144, 6: 1048576, 7: 4194304}


def _block_id(nbytes):                       # lz4Encode.js:337-343
    if not nbytes or nbytes <= 65536:
        return 4
    if nbytes <= 262144:
        return 5
    if nbytes <= 1048576:
        return 6
    return 7


def _u32(v):
    return int(v & 0xFFFFFFFF).to_bytes(4, "little")


def _jenkins_slots(buf):
    """lz4Encode.js:155-167: table[JENKINS(seq_i)] = i + 1 for ascending i (the last writer of a slot wins)."""
    n = buf.size
    table = np.zeros(16384, dtype=np.int32)
    if n < 4:
        return table
    b = buf.astype(np.uint32)
    h = b[:n - 3] | (b[1:n - 2] << 8) | (b[2:n - 1] << 16) | (b[3:] << 24)
    with np.errstate(over="ignore"):
        h = h + np.uint32(2127912214) + (h << np.uint32(12))
        h = h ^ np.uint32(3345072700) ^ (h >> np.uint32(19))
        h = h + np.uint32(374761393) + (h << np.uint32(5))
        h = (h + np.uint32(3550635116)) ^ (h << np.uint32(9))
        h = h + np.uint32(4251993797) + (h << np.uint32(3))
        h = h ^ np.uint32(3042594569) ^ (h >> np.uint32(16))
    slot = (h >> np.uint32(18)) & np.uint32(16383)
    table[slot] = np.arange(1, n - 2, dtype=np.int32)          # ascending assignment: the highest i per slot stays
    return table


def create_frame_header(block_independence, content_checksum, bd_id, dict_id, ctx=None):
    """lz4Encode.js:61-94 (no content size in stream frames; a dictId of 0 counts as absent, `if (dictId)`), written by the
    library's one header writer (dlz4_frame_header)."""
    from .sharded import frame_header
    return frame_header(0, BLOCK_MAX_SIZES.get(bd_id & 7, 4194304), block_independence, content_checksum, False, False,
                        dict_id if dict_id else None)


class LZ4Encoder(object):
    """new LZ4Encoder(maxBlockSize=4194304, blockIndependence=false, contentChecksum=false, dictionary=null)"""

    def __init__(self, maxBlockSize=4194304, blockIndependence=False, contentChecksum=False, dictionary=None, ctx=None):
        self._ctx = ctx or api.default_context()
        self.blockIndependence = bool(blockIndependence)
        self.contentChecksum = bool(contentChecksum)
        self.blockSize = BLOCK_MAX_SIZES.get(_block_id(maxBlockSize), 4194304)
        self.bdId = _block_id(self.blockSize)
        self.buffer = np.zeros(0, dtype=np.uint8)
        self.hasWrittenHeader = False
        self.isClosed = False
        self.hashTable = np.zeros(16384, dtype=np.int32)
        self.dictSize = 0
        self.hasher = api.XXHash32(0, ctx=self._ctx) if self.contentChecksum else None
        self.dictId = None
        if dictionary is not None and len(dictionary) > 0:
            d = api.ensureBuffer(dictionary)
            self.dictId = api.xxHash32(d, 0, ctx=self._ctx)
            window = d[max(0, d.size - MAX_WINDOW_SIZE):]                    # _initDictionary, :140-168
            self.buffer = window.copy()
            self.dictSize = int(window.size)
            self.hashTable = _jenkins_slots(self.buffer)

    # ---- add / finish ---------------------------------------------------------------------------------------------
    # `buffer` holds what the reference's this.buffer holds BEFORE the chunk of the current add(): the window (dictSize bytes,
    # at most 64 KiB) followed by the pending bytes of an incomplete block.  The new chunk is never appended to it: the C side
    # joins the two pieces on the device (dlz4_stream_blocks), only the bytes behind the last full block are copied.
    def add(self, chunk):
        if self.isClosed:
            raise RuntimeError("Stream is closed")
        data = api.ensureBuffer(chunk)
        if data.size == 0:
            return []
        results = []
        if not self.hasWrittenHeader:
            results.append(create_frame_header(self.blockIndependence, self.contentChecksum, self.bdId, self.dictId, self._ctx))
            self.hasWrittenHeader = True
        pending = self.buffer.size - self.dictSize
        full = (pending + data.size) // self.blockSize                          # `while (buffer.length >= dictSize + blockSize)`
        if full:
            take = full * self.blockSize - pending                              # bytes of `data` inside the full blocks
            results.extend(self._flush_blocks(data[:take], full * self.blockSize))
            data = data[take:]
        if data.size:
            self.buffer = np.concatenate([self.buffer, data])                    # < one block
        return results

    def finish(self):
        if self.isClosed:
            return []
        self.isClosed = True
        frames = []
        if not self.hasWrittenHeader:
            frames.append(create_frame_header(self.blockIndependence, self.contentChecksum, self.bdId, self.dictId, self._ctx))
        rest = self.buffer.size - self.dictSize
        if rest > 0:
            frames.extend(self._flush_blocks(np.zeros(0, dtype=np.uint8), rest))  # the final short block
        frames.append(_u32(0))
        if self.hasher:
            frames.append(_u32(self.hasher.digest()))
        return frames

    # ---- _flushBlock over `total` pending bytes at once (:215-298): working buffer = self.buffer ++ tail ---------------
    def _flush_blocks(self, tail, total):
        bs = self.blockSize
        start = self.dictSize
        head = self.buffer
        n = (total + bs - 1) // bs
        cap = total + 8 * n + 64
        body = np.empty(cap, dtype=np.uint8)
        plen = np.zeros(n, dtype=np.uint32)
        blen = C.c_uint64(0)
        linked = not self.blockIndependence
        if linked:
            table = np.ascontiguousarray(self.hashTable, dtype=np.int32)
        else:
            table = None
        st = api.lib().dlz4_stream_blocks(self._ctx.handle, api._ptr(head), head.size, api._ptr(tail), tail.size, int(start), int(total),
                                          int(bs), int(linked), api._ptr(table), api._ptr(body), cap, C.byref(blen), api._ptr(plen))
        self._ctx.check(st)
        if self.hasher:
            # content checksum (:110, :181): bytes are flushed once and in order, so hashing every flush hashes the stream; the
            # flushed bytes are still on the device
            if not self.hasher.update_resident(0):
                self.hasher.update(np.concatenate([head, tail])[start:start + total])
        out = []
        mv = memoryview(body)
        pos = 0
        for k in range(n):
            e = pos + int(plen[k])
            out.append(bytes(mv[pos:e]))                                         # [u32 size | stored bit][payload], :263-273
            pos = e
        consumed_end = start + total
        if linked:
            keep = min(consumed_end, MAX_WINDOW_SIZE)                             # :276-296 slide the window, rebase the table
            shift = consumed_end - keep
            self.buffer = _tail_of(head, tail, consumed_end, keep)
            self.dictSize = keep
            self.hashTable = np.where(table > shift, table - shift, 0).astype(np.int32)
        else:
            self.hashTable[:] = 0                                                # table cleared before every block (:240-242): what the
            self.buffer = np.zeros(0, dtype=np.uint8)                            # last parse leaves is not observable
            self.dictSize = 0
        return out


def _tail_of(head, tail, end, keep):
    """The last `keep` bytes of (head ++ tail)[:end] as a new array."""
    lo = end - keep
    if lo >= head.size:
        return tail[lo - head.size:end - head.size].copy()
    return np.concatenate([head[lo:], tail[:end - head.size]])


class _Pinned(object):
    """Grow-only page-locked staging buffer (the frame calls copy at the full PCIe rate from / to it)."""

    def __init__(self):
        self.ptr, self.arr = None, None

    def get(self, nbytes):
        if self.arr is None or self.arr.size < nbytes:
            self.free()
            cap = max(int(nbytes), 1 << 20)
            cap += cap >> 2
            self.ptr = api.lib().dlz4_pinned_alloc(cap)
            if not self.ptr:
                raise MemoryError("dlz4_pinned_alloc(%d)" % cap)
            self.arr = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(cap,))
        return self.arr

    def free(self):
        if self.ptr:
            self.arr = None
            api.lib().dlz4_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LZ4Decoder(object):
    """new LZ4Decoder(dictionary=null, verifyChecksum=true); update(chunk) -> list of decoded chunks (one per block)."""

    def __init__(self, dictionary=None, verifyChecksum=True, ctx=None):
        self._ctx = ctx or api.default_context()
        self.dictionary = api.ensureBuffer(dictionary) if dictionary is not None and len(dictionary) > 0 else None
        self.verifyChecksum = bool(verifyChecksum)
        self.state = "magic"
        self.buffer = bytearray()
        self.hasher = None
        self.window = np.zeros(0, dtype=np.uint8)
        if self.dictionary is not None:
            self.window = self.dictionary[max(0, self.dictionary.size - MAX_WINDOW_SIZE):].copy()   # _initWindow
        self.blockIndependence = True
        self.hasBlockChecksum = self.hasContentChecksum = self.hasContentSize = self.hasDictId = False
        self._bd = 0x70
        self._pin_frame, self._pin_out = _Pinned(), _Pinned()

    def update(self, chunk):
        self.buffer += memoryview(np.ascontiguousarray(api.ensureBuffer(chunk)))
        output = []
        while True:
            if self.state == "magic":                                            # lz4Decode.js:118-131
                if len(self.buffer) < 4:
                    break
                if int.from_bytes(self.buffer[:4], "little") != 0x184D2204:
                    raise api.LZ4Error(api.E_BAD_MAGIC, "LZ4: Invalid Magic Number")
                del self.buffer[:4]
                self.state = "header"
                self.hasher = None
            if self.state == "header":                                           # :134-178
                if len(self.buffer) < 2:
                    break
                flg = self.buffer[0]
                self.blockIndependence = bool(flg & 0x20)
                self.hasBlockChecksum = bool(flg & 0x10)
                self.hasContentSize = bool(flg & 0x08)
                self.hasContentChecksum = bool(flg & 0x04)
                self.hasDictId = bool(flg & 0x01)
                need = 2 + (8 if self.hasContentSize else 0) + (4 if self.hasDictId else 0) + 1
                if len(self.buffer) < need:
                    break
                self._bd = self.buffer[1]
                if self.hasDictId:
                    cur = 2 + (8 if self.hasContentSize else 0)
                    expected = int.from_bytes(self.buffer[cur:cur + 4], "little")
                    if self.dictionary is None:
                        raise api.LZ4Error(api.E_DICT_OOB, "LZ4: Archive requires a Dictionary, but none was provided.")
                    actual = api.xxHash32(self.dictionary, 0, ctx=self._ctx)
                    if actual != expected:
                        raise api.LZ4Error(api.E_DICT_OOB, "LZ4: Dictionary ID Mismatch. Header: 0x%x, Provided: 0x%x" % (expected, actual))
                del self.buffer[:need]
                # (the hasher only matters for a frame that carries a content checksum, :246-258)
                self.hasher = api.XXHash32(0, ctx=self._ctx) if (self.verifyChecksum and self.hasContentChecksum) else None
                self.state = "blocks"
            if self.state == "blocks":                                           # :181-243, every complete block at once
                nblocks, pos, end_mark, body_end = 0, 0, False, 0
                have = len(self.buffer)
                while True:
                    if have - pos < 4:
                        break
                    val = int.from_bytes(self.buffer[pos:pos + 4], "little")
                    if val == 0:
                        pos += 4
                        end_mark = True
                        break
                    need = (val & 0x7FFFFFFF) + (4 if self.hasBlockChecksum else 0)
                    if have - pos - 4 < need:
                        break
                    nblocks += 1
                    pos += 4 + need
                    body_end = pos
                if nblocks:
                    output.extend(self._decode(nblocks, body_end))
                del self.buffer[:pos]
                if not end_mark:
                    break
                self.state = "checksum"
            if self.state == "checksum":                                         # :246-266
                if self.hasContentChecksum:
                    if len(self.buffer) < 4:
                        break
                    if self.verifyChecksum and self.hasher:
                        if int.from_bytes(self.buffer[:4], "little") != self.hasher.digest():
                            raise api.LZ4Error(api.E_CONTENT_CHECKSUM, "LZ4: Content Checksum Error")
                    del self.buffer[:4]
                self.state = "magic"
                self.hasher = None
                if len(self.buffer) == 0:
                    break
        return output

    def _decode(self, nblocks, body_end):
        """The `nblocks` complete blocks in buffer[:body_end] as a synthetic frame (size words, data and block checksums as
        they came, copied once into page-locked memory): linked blocks see window ++ earlier output (:215-222,:240).  Block
        checksums are skipped like in the reference's stream decoder."""
        from .sharded import frame_header
        bmax = BLOCK_MAX_SIZES.get((self._bd >> 4) & 7, 4194304)
        hdr = frame_header(0, bmax, self.blockIndependence, False, False, self.hasBlockChecksum)
        total = len(hdr) + body_end + 4
        f = self._pin_frame.get(total + 16)
        f[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        src = np.frombuffer(self.buffer, dtype=np.uint8, count=body_end)
        f[len(hdr):len(hdr) + body_end] = src
        del src                                                                  # (the bytearray is resized by the caller)
        f[len(hdr) + body_end:total] = 0                                         # EndMark
        # output room: the library's own bound for this frame (sum of min(len * 255, blockMaxSize) over the blocks), not
        # nblocks * blockMaxSize -- thousands of small blocks under a 4 MiB descriptor would ask for tens of GiB of pinned memory
        info = api.FrameInfo()
        self._ctx.check(api.lib().dlz4_frame_info(api._ptr(f), total, C.byref(info)))
        cap = int(info.max_decoded)
        out = self._pin_out.get(cap + 16)
        n = C.c_uint64(0)
        olen = np.zeros(nblocks, dtype=np.uint32)
        hist = self.window if (not self.blockIndependence and self.window.size) else None
        st = api.lib().dlz4_frame_decompress_ex(self._ctx.handle, api._ptr(f), total, api._ptr(hist), hist.size if hist is not None else 0, 0,
                                                api._ptr(out), cap, C.byref(n), api._ptr(olen))
        self._ctx.check(st)
        chunks, p = [], 0
        for k in range(nblocks):
            chunks.append(out[p:p + int(olen[k])].tobytes())
            p += int(olen[k])
        decoded = out[:p]
        if self.hasher and p:
            if not self.hasher.update_resident(1):                                # the decoded bytes are still on the device
                self.hasher.update(decoded)
        if not self.blockIndependence and p:                                     # _updateWindow, :278-304: the last 64 KiB
            if p >= MAX_WINDOW_SIZE:
                self.window = decoded[p - MAX_WINDOW_SIZE:].copy()
            else:
                self.window = np.concatenate([self.window, decoded])[-MAX_WINDOW_SIZE:].copy()
        return chunks
