"""Host-side mirror of the reference's off-thread offload: `LZ4Worker.compress / decompress`
(src/webWorker/workerClient.js:114-152, worker side src/webWorker/lz4.worker.js:70-83).  The reference posts the buffer to ONE
lazily created Web Worker and resolves a Promise with the result; here the single worker is a thread that owns its own
`dlz4_ctx` (own CUDA streams and scratch, so it never contends with the caller's context) and the Promise is a
`concurrent.futures.Future`.  ctypes releases the GIL during the C call: the caller's thread keeps running while the GPU works,
and several pending tasks queue up in order like messages to the worker."""
import threading
from concurrent.futures import Future, ThreadPoolExecutor

from . import api

_lock = threading.Lock()
_executor = None
_ctx = None


def _worker_ctx():
    global _ctx
    if _ctx is None:
        _ctx = api.Context(api.default_context().device)
    return _ctx


def _get_worker():
    global _executor
    with _lock:
        if _executor is None:                                  # getWorker(): created on first use, then reused (workerClient.js:28-35)
            _executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="LZ4-Worker")
    return _executor


def _run(task, data, options):
    options = options or {}
    ctx = _worker_ctx()
    if task == "compress":                                     # lz4.worker.js:76-78
        return api.compressBuffer(data, options.get("dictionary"), options.get("maxBlockSize", 4194304),
                                  options.get("blockIndependence", False), options.get("contentChecksum", False), ctx=ctx)
    if task == "decompress":                                   # :80-82
        return api.decompressBuffer(data, options.get("dictionary"), options.get("verifyChecksum", True), ctx=ctx)
    raise ValueError('LZ4 Worker: Unknown task "%s"' % task)


class _LZ4Worker(object):
    @staticmethod
    def compress(data, options=None) -> Future:
        return _get_worker().submit(_run, "compress", data, options)

    @staticmethod
    def decompress(data, options=None) -> Future:
        return _get_worker().submit(_run, "decompress", data, options)


LZ4Worker = _LZ4Worker()
