"""ctypes binding of the system liblz4.so.1 (1.9.4) with hand-declared prototypes -- third-party cross-check only
(stands in for "lz4-CLI-produced frames": the CLI is a thin wrapper over LZ4F_*).  Never used for parity of
compressed bytes (liblz4's match finder differs), only as an independent decoder / frame producer."""
import ctypes as C
import ctypes.util

import numpy as np


def _load():
    for name in ("liblz4.so.1", ctypes.util.find_library("lz4")):
        if not name:
            continue
        try:
            return C.CDLL(name)
        except OSError:
            continue
    return None


L = _load()


class FrameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int), ("frameType", C.c_int),
                ("contentSize", C.c_ulonglong), ("dictID", C.c_uint), ("blockChecksumFlag", C.c_int)]


class Prefs(C.Structure):
    _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint), ("favorDecSpeed", C.c_uint),
                ("reserved", C.c_uint * 3)]


if L is not None:
    L.LZ4F_compressFrameBound.restype = C.c_size_t
    L.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.POINTER(Prefs)]
    L.LZ4F_compressFrame.restype = C.c_size_t
    L.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Prefs)]
    L.LZ4F_isError.restype = C.c_uint
    L.LZ4F_isError.argtypes = [C.c_size_t]
    L.LZ4F_getErrorName.restype = C.c_char_p
    L.LZ4F_getErrorName.argtypes = [C.c_size_t]
    L.LZ4F_createDecompressionContext.restype = C.c_size_t
    L.LZ4F_createDecompressionContext.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
    L.LZ4F_freeDecompressionContext.restype = C.c_size_t
    L.LZ4F_freeDecompressionContext.argtypes = [C.c_void_p]
    L.LZ4F_decompress.restype = C.c_size_t
    L.LZ4F_decompress.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
    L.LZ4F_decompress_usingDict.restype = C.c_size_t
    L.LZ4F_decompress_usingDict.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t),
                                            C.c_void_p, C.c_size_t, C.c_void_p]
    L.LZ4_decompress_safe.restype = C.c_int
    L.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.LZ4_decompress_safe_usingDict.restype = C.c_int
    L.LZ4_decompress_safe_usingDict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.LZ4_compress_default.restype = C.c_int
    L.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.LZ4_compressBound.restype = C.c_int
    L.LZ4_compressBound.argtypes = [C.c_int]


def available():
    return L is not None


def compress_frame(data, block_size_id=7, linked=False, content_checksum=True, content_size=False, block_checksum=False):
    """LZ4F_compressFrame with CLI-like preferences (the lz4 CLI default: 4 MiB independent blocks, content checksum on,
    no content size)."""
    data = np.frombuffer(bytes(data), dtype=np.uint8)
    p = Prefs()
    p.frameInfo.blockSizeID = block_size_id
    p.frameInfo.blockMode = 0 if linked else 1          # LZ4F_blockLinked = 0, LZ4F_blockIndependent = 1
    p.frameInfo.contentChecksumFlag = 1 if content_checksum else 0
    p.frameInfo.contentSize = data.size if content_size else 0
    p.frameInfo.blockChecksumFlag = 1 if block_checksum else 0
    bound = L.LZ4F_compressFrameBound(data.size, C.byref(p))
    out = np.zeros(bound, dtype=np.uint8)
    n = L.LZ4F_compressFrame(out.ctypes.data, bound, data.ctypes.data if data.size else None, data.size, C.byref(p))
    if L.LZ4F_isError(n):
        raise RuntimeError(L.LZ4F_getErrorName(n).decode())
    return out[:n].tobytes()


def decompress_frame(frame, max_out, dictionary=None):
    frame = np.frombuffer(bytes(frame), dtype=np.uint8)
    ctx = C.c_void_p()
    r = L.LZ4F_createDecompressionContext(C.byref(ctx), 100)
    if L.LZ4F_isError(r):
        raise RuntimeError(L.LZ4F_getErrorName(r).decode())
    out = np.zeros(max_out + 64, dtype=np.uint8)
    ip, op = 0, 0
    d = np.frombuffer(bytes(dictionary), dtype=np.uint8) if dictionary else None
    try:
        while ip < frame.size:
            src_sz = C.c_size_t(frame.size - ip)
            dst_sz = C.c_size_t(out.size - op)
            if d is not None:
                r = L.LZ4F_decompress_usingDict(ctx, out.ctypes.data + op, C.byref(dst_sz), frame.ctypes.data + ip, C.byref(src_sz),
                                                d.ctypes.data, d.size, None)
            else:
                r = L.LZ4F_decompress(ctx, out.ctypes.data + op, C.byref(dst_sz), frame.ctypes.data + ip, C.byref(src_sz), None)
            if L.LZ4F_isError(r):
                raise RuntimeError(L.LZ4F_getErrorName(r).decode())
            ip += src_sz.value
            op += dst_sz.value
            if r == 0:
                break
            if src_sz.value == 0 and dst_sz.value == 0:
                raise RuntimeError("LZ4F_decompress made no progress")
    finally:
        L.LZ4F_freeDecompressionContext(ctx)
    return out[:op].tobytes()


def decompress_block(block, max_out, dictionary=None):
    block = np.frombuffer(bytes(block), dtype=np.uint8)
    out = np.zeros(max_out + 8, dtype=np.uint8)
    if dictionary:
        d = np.frombuffer(bytes(dictionary), dtype=np.uint8)
        n = L.LZ4_decompress_safe_usingDict(block.ctypes.data, out.ctypes.data, block.size, max_out, d.ctypes.data, d.size)
    else:
        n = L.LZ4_decompress_safe(block.ctypes.data, out.ctypes.data, block.size, max_out)
    if n < 0:
        raise RuntimeError("LZ4_decompress_safe failed: %d" % n)
    return out[:n].tobytes()
