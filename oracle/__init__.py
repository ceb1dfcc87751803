"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/lz4_oracle.c, the CPU restatement of the divortio-lz4 hot path
(compressBlock / decompressBlock / compressBuffer / decompressBuffer / xxHash32; file:line
citations are in the C source).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  The product package
(divortio_lz4_b200) never does.

Parity pinning status: the reference is pure JavaScript and cannot be executed in this image
(no JS engine).  The oracle is pinned against the reference's own golden vectors and KATs
(tests/golden.test.mjs, tests/xxhash32/xxhash32.test.mjs), the SURVEY A.3 scratch KATs, an
independent literal transliteration (tests/jsref.py) and liblz4/libxxhash as decoders.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblz4_oracle.so")

E_OUTPUT_TOO_SMALL, E_MALFORMED, E_OFFSET_ZERO, E_DICT_OOB = -1, -2, -3, -4
E_BAD_MAGIC, E_BAD_VERSION, E_CONTENT_CHECKSUM, E_RANGE = -5, -6, -7, -8

MESSAGES = {
    E_OUTPUT_TOO_SMALL: "LZ4: Output Buffer Too Small",
    E_MALFORMED: "LZ4: Malformed Input",
    E_OFFSET_ZERO: "LZ4: Invalid Offset 0",
    E_DICT_OOB: "LZ4: Dictionary Offset Out of Bounds",
    E_BAD_MAGIC: "LZ4: Invalid Magic Number",
    E_BAD_VERSION: "LZ4: Unsupported Version",
    E_CONTENT_CHECKSUM: "LZ4: Content Checksum Error",
    E_RANGE: "RangeError: offset is out of bounds",
}


class OracleError(Exception):
    def __init__(self, code, extra=""):
        self.code = code
        super().__init__(MESSAGES.get(code, "LZ4: error %d" % code) + extra)


def build(force=False):
    src = os.path.join(_HERE, "lz4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_xxh32.restype = C.c_uint32
        L.orc_xxh32.argtypes = [u8p, C.c_uint64, C.c_uint32]
        L.orc_xxh32_batch.restype = None
        L.orc_xxh32_batch.argtypes = [u8p, u64p, u32p, C.c_uint32, C.c_uint32, u32p]
        L.orc_compress_bound.restype = C.c_uint64
        L.orc_compress_bound.argtypes = [C.c_uint64]
        L.orc_frame_bound.restype = C.c_uint64
        L.orc_frame_bound.argtypes = [C.c_uint64]
        L.orc_compress_block.restype = C.c_int32
        L.orc_compress_block.argtypes = [u8p, C.c_int32, C.c_int32, i32p, u8p, C.c_int64, C.c_int32]
        L.orc_compress_blocks.restype = None
        L.orc_compress_blocks.argtypes = [u8p, u64p, u32p, C.c_uint32, u8p, u64p, u32p]
        L.orc_compress_blocks_prefix.restype = None
        L.orc_compress_blocks_prefix.argtypes = [u8p, C.c_uint32, i32p, u8p, u64p, u32p, C.c_uint32, u8p, u64p, u32p]
        L.orc_warm_table_jenkins.restype = None
        L.orc_warm_table_jenkins.argtypes = [u8p, C.c_int32, i32p]
        L.orc_decompress_block.restype = C.c_int64
        L.orc_decompress_block.argtypes = [u8p, C.c_int64, C.c_int64, u8p, C.c_int64, C.c_int64, u8p, C.c_int64]
        L.orc_decompress_block_literal.restype = C.c_int64
        L.orc_decompress_block_literal.argtypes = [u8p, C.c_int64, C.c_int64, C.c_int64, u8p, C.c_int64, C.c_int64, u8p, C.c_int64]
        L.orc_decompress_blocks.restype = None
        L.orc_decompress_blocks.argtypes = [u8p, u64p, u32p, C.c_uint32, u8p, u64p, u32p, u8p, C.c_uint32, u32p, i32p]
        L.orc_compress_buffer.restype = C.c_int64
        L.orc_compress_buffer.argtypes = [u8p, C.c_int64, u8p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int64]
        L.orc_decompress_buffer.restype = C.c_int64
        L.orc_decompress_buffer.argtypes = [u8p, C.c_int64, u8p, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.orc_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _u8(x):
    """bytes / bytearray / ndarray -> contiguous uint8 ndarray (no copy when possible)."""
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    return np.frombuffer(bytes(x) if not isinstance(x, (bytes, bytearray, memoryview)) else x, dtype=np.uint8)


def _p(a):
    return a.ctypes.data if a is not None and a.size else (a.ctypes.data if a is not None else None)


def xxh32(data, seed=0):
    a = _u8(data)
    return int(lib().orc_xxh32(a.ctypes.data, a.size, seed))


def xxh32_batch(base, off, length, seed=0):
    base = _u8(base)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    out = np.zeros(len(off), dtype=np.uint32)
    lib().orc_xxh32_batch(base.ctypes.data, off.ctypes.data, length.ctypes.data, len(off), seed, out.ctypes.data)
    return out


def compress_bound(n):
    return int(lib().orc_compress_bound(n))


def new_table():
    return np.zeros(16384, dtype=np.int32)


def compress_block(src, src_start=0, src_len=None, table=None, out=None, out_offset=0):
    """compressBlock(src, output, srcStart, srcLen, hashTable, outputOffset) -> bytes written.
    Returns (n, out) where out is the output array (allocated to the bound when not given)."""
    src = _u8(src)
    if src_len is None:
        src_len = src.size - src_start
    if table is None:
        table = new_table()
    if out is None:
        out = np.zeros(out_offset + compress_bound(src_len), dtype=np.uint8)
    # pad so the 4-byte loads at the last probe never leave the buffer
    n = lib().orc_compress_block(src.ctypes.data, src_start, src_len, table.ctypes.data,
                                 out.ctypes.data, out.size, out_offset)
    return int(n), out


def compress_block_bytes(src, src_start=0, src_len=None, table=None):
    n, out = compress_block(src, src_start, src_len, table)
    return out[:n].tobytes()


def compress_blocks(src, off, length):
    """Independent raw blocks, fresh table each.  Returns (dst, dst_off, comp_len)."""
    src = _u8(src)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    bounds = length.astype(np.uint64) + length.astype(np.uint64) // 255 + 16
    dst_off = np.zeros(len(off), dtype=np.uint64)
    if len(off):
        dst_off[1:] = np.cumsum(bounds)[:-1]
    dst = np.zeros(int(bounds.sum()), dtype=np.uint8)
    comp = np.zeros(len(off), dtype=np.uint32)
    lib().orc_compress_blocks(src.ctypes.data, off.ctypes.data, length.ctypes.data, len(off),
                              dst.ctypes.data, dst_off.ctypes.data, comp.ctypes.data)
    return dst, dst_off, comp


def compress_blocks_prefix(prefix, init_table, src, off, length):
    prefix = _u8(prefix)
    src = _u8(src)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    bounds = length.astype(np.uint64) + length.astype(np.uint64) // 255 + 16
    dst_off = np.zeros(len(off), dtype=np.uint64)
    if len(off):
        dst_off[1:] = np.cumsum(bounds)[:-1]
    dst = np.zeros(int(bounds.sum()), dtype=np.uint8)
    comp = np.zeros(len(off), dtype=np.uint32)
    tp = None if init_table is None else np.ascontiguousarray(init_table, dtype=np.int32).ctypes.data
    lib().orc_compress_blocks_prefix(prefix.ctypes.data, prefix.size, tp, src.ctypes.data, off.ctypes.data,
                                     length.ctypes.data, len(off), dst.ctypes.data, dst_off.ctypes.data,
                                     comp.ctypes.data)
    return dst, dst_off, comp


def warm_table_jenkins(work, dict_len):
    work = _u8(work)
    t = new_table()
    lib().orc_warm_table_jenkins(work.ctypes.data, dict_len, t.ctypes.data)
    return t


def decompress_block(inp, in_off, in_size, out, out_off=0, dictionary=None, literal=False):
    """decompressBlock(input, inputOffset, inputSize, output, outputOffset, dictionary) -> bytes written.
    `out` is a writable uint8 ndarray (the WHOLE output array).  Raises OracleError."""
    inp = _u8(inp)
    d = _u8(dictionary) if dictionary is not None else None
    dp, dl = (d.ctypes.data, d.size) if d is not None and d.size else (None, 0)
    if literal:
        n = lib().orc_decompress_block_literal(inp.ctypes.data, inp.size, in_off, in_size,
                                               out.ctypes.data, out.size, out_off, dp, dl)
    else:
        n = lib().orc_decompress_block(inp.ctypes.data, in_off, in_size, out.ctypes.data, out.size, out_off, dp, dl)
    if n < 0:
        raise OracleError(int(n))
    return int(n)


def decompress_blocks(src, off, length, dst_off, dst_cap, dictionary=None):
    src = _u8(src)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
    dst_cap = np.ascontiguousarray(dst_cap, dtype=np.uint32)
    total = int((dst_off + dst_cap).max()) if len(off) else 0
    dst = np.zeros(total, dtype=np.uint8)
    out_len = np.zeros(len(off), dtype=np.uint32)
    status = np.zeros(len(off), dtype=np.int32)
    d = _u8(dictionary) if dictionary is not None else None
    dp, dl = (d.ctypes.data, d.size) if d is not None and d.size else (None, 0)
    lib().orc_decompress_blocks(src.ctypes.data, off.ctypes.data, length.ctypes.data, len(off), dst.ctypes.data,
                                dst_off.ctypes.data, dst_cap.ctypes.data, dp, dl, out_len.ctypes.data,
                                status.ctypes.data)
    return dst, out_len, status


def frame_bound(n):
    return int(lib().orc_frame_bound(n))


def compress_buffer(inp, dictionary=None, max_block_size=4194304, block_independence=False,
                    content_checksum=False, add_content_size=True, output_buffer=None, block_checksum=False):
    """compressBuffer(input, dictionary, maxBlockSize, blockIndependence, contentChecksum,
    addContentSize, outputBuffer) -> frame bytes (bufferCompress.js:100)."""
    inp = _u8(inp)
    d = _u8(dictionary) if dictionary is not None else None
    dp, dl = (d.ctypes.data, d.size) if d is not None and d.size else (None, 0)
    if output_buffer is None:
        worst = (19 + inp.size + (inp.size // 255) + 64 + 8)            # bufferCompress.js:140
        worst = max(worst, frame_bound(inp.size))
        out = np.zeros(worst, dtype=np.uint8)
    else:
        out = output_buffer
    n = lib().orc_compress_buffer(inp.ctypes.data, inp.size, dp, dl, int(max_block_size or 0),
                                  int(bool(block_independence)), int(bool(content_checksum)),
                                  int(bool(add_content_size)), int(bool(block_checksum)),
                                  out.ctypes.data, out.size)
    return out[:min(int(n), out.size)].tobytes()


def decompress_buffer(data, dictionary=None, verify_checksum=True):
    """decompressBuffer(input, dictionary, verifyChecksum) -> bytes (bufferDecompress.js:51)."""
    data = _u8(data)
    d = _u8(dictionary) if dictionary is not None else None
    dp, dl = (d.ctypes.data, d.size) if d is not None and d.size else (None, 0)
    outp = C.c_void_p()
    bad = C.c_int(0)
    n = lib().orc_decompress_buffer(data.ctypes.data, data.size, dp, dl, int(bool(verify_checksum)),
                                    C.byref(outp), C.byref(bad))
    if n < 0:
        raise OracleError(int(n), " %d" % bad.value if n == E_BAD_VERSION else "")
    try:
        return C.string_at(outp.value, int(n)) if n else b""
    finally:
        lib().orc_free(outp)
