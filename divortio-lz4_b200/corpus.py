"""Synthetic corpora of SURVEY.md 8d (ctypes over tools/corpus.c).  Host helper for tests and bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "tools", "libdlz4_corpus.so")
_lib = None


def _L():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "tools", "corpus.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
        _lib = C.CDLL(_SO)
        _lib.corpus_benchjson_reclen.restype = C.c_uint64
    return _lib


def _buf(n, out):
    if out is None:
        return np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n
    return out


def log(seed, n, out=None):
    b = _buf(n, out)
    _L().corpus_log(C.c_uint64(seed), C.c_void_p(b.ctypes.data), C.c_uint64(n))
    return b


def zero(n, out=None):
    b = _buf(n, out)
    b[:n] = 0
    return b


def rand(seed, n, out=None):
    b = _buf(n, out)
    _L().corpus_rand(C.c_uint64(seed), C.c_void_p(b.ctypes.data), C.c_uint64(n))
    return b


def mixed(seed, n, out=None):
    b = _buf(n, out)
    _L().corpus_mixed(C.c_uint64(seed), C.c_void_p(b.ctypes.data), C.c_uint64(n))
    return b


def jsonmsgs(seed, first, count, out=None):
    b = _buf(count * 4096, out)
    _L().corpus_jsonmsgs(C.c_uint64(seed), C.c_uint64(first), C.c_uint64(count), C.c_void_p(b.ctypes.data))
    return b


def json_dictionary(seed):
    """First 65536 bytes of JSONMSG(seed, 0..15) concatenated."""
    return jsonmsgs(seed, 0, 16)[:65536].copy()


def benchjson(n, out=None):
    """The reference benchmark's 219-byte record repeated (benchmark/src/base/benchUtils.js:7-22)."""
    b = _buf(n, out)
    _L().corpus_benchjson(C.c_void_p(b.ctypes.data), C.c_uint64(n))
    return b
