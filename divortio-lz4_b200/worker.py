"""Host-side mirror of the reference's off-thread offload: `LZ4Worker` of src/webWorker/workerClient.js --
buffer tasks `compress / decompress` (:114-130, worker side src/webWorker/lz4.worker.js:70-83) and stream tasks
`compressStream / decompressStream` (:96-110, :143-152, worker side lz4.worker.js:30-69).

The reference posts the buffer -- or transfers the streams -- to ONE lazily created Web Worker and resolves a Promise.  Here
the Promise is a `concurrent.futures.Future` and the worker side is a small pool of threads, each owning its own `dlz4_ctx`
(own CUDA streams, scratch and staging ring, so it never contends with the caller's context or with the other worker).
ctypes releases the GIL during the C calls: the caller keeps running, and with two workers the copy-out of one queued task
overlaps the copy-in and kernels of the next instead of the tasks running strictly one after the other.  Results are tied to
their Future, not to completion order, exactly like the reference's `pendingTasks` map (:14, :37-58).

Stream tasks pipe `readable` (any iterable of byte chunks) through the stream codec of src/stream/* (here
`stream.LZ4Encoder / LZ4Decoder`, the classes createCompressStream / createDecompressStream wrap) into `writable` (an object
with write(bytes) and optionally close(), or a callable); the Future resolves to None when the stream is complete, and the
streams belong to the worker from the call on (the reference transfers them, :107)."""
import threading
from concurrent.futures import Future, ThreadPoolExecutor

from . import api

_lock = threading.Lock()
_executor = None
_local = threading.local()
WORKERS = 2


def _worker_ctx():
    if getattr(_local, "ctx", None) is None:
        _local.ctx = api.Context(api.default_context().device)
    return _local.ctx


def _get_worker():
    global _executor
    with _lock:
        if _executor is None:                                  # getWorker(): created on first use, then reused (workerClient.js:28-35)
            _executor = ThreadPoolExecutor(max_workers=WORKERS, thread_name_prefix="LZ4-Worker")
    return _executor


def _sink(writable):
    if callable(writable):
        return writable, None
    return writable.write, getattr(writable, "close", None)


def _run(task, data, options, writable=None):
    options = options or {}
    ctx = _worker_ctx()
    if task == "compress":                                     # lz4.worker.js:76-78
        return api.compressBuffer(data, options.get("dictionary"), options.get("maxBlockSize", 4194304),
                                  options.get("blockIndependence", False), options.get("contentChecksum", False), ctx=ctx)
    if task == "decompress":                                   # :80-82
        return api.decompressBuffer(data, options.get("dictionary"), options.get("verifyChecksum", True), ctx=ctx)
    if task in ("stream-compress", "stream-decompress"):       # :30-69: readable.pipeThrough(transform).pipeTo(writable)
        from . import stream
        write, close = _sink(writable)
        if task == "stream-compress":
            codec = stream.LZ4Encoder(options.get("maxBlockSize", 4194304), options.get("blockIndependence", False),
                                      options.get("contentChecksum", False), options.get("dictionary"), ctx=ctx)
            step, finish = codec.add, codec.finish
        else:
            codec = stream.LZ4Decoder(options.get("dictionary"), options.get("verifyChecksum", True), ctx=ctx)
            step, finish = codec.update, lambda: []
        for chunk in data:
            for piece in step(chunk):
                write(bytes(piece))
        for piece in finish():
            write(bytes(piece))
        if close:
            close()
        return None
    raise ValueError('LZ4 Worker: Unknown task "%s"' % task)


class _LZ4Worker(object):
    @staticmethod
    def compress(data, options=None) -> Future:
        return _get_worker().submit(_run, "compress", data, options)

    @staticmethod
    def decompress(data, options=None) -> Future:
        return _get_worker().submit(_run, "decompress", data, options)

    @staticmethod
    def compressStream(readable, writable, options=None) -> Future:
        return _get_worker().submit(_run, "stream-compress", readable, options, writable)

    @staticmethod
    def decompressStream(readable, writable, options=None) -> Future:
        return _get_worker().submit(_run, "stream-decompress", readable, options, writable)


LZ4Worker = _LZ4Worker()
