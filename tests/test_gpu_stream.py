"""GPU parity, streaming codec: divortio_lz4_b200.stream.LZ4Encoder / LZ4Decoder against the block-by-block
restatement of src/shared/lz4Encode.js / lz4Decode.js (tests/jsref_stream.py) -- the same pieces, call for call, whatever the
chunking.  The GPU classes batch every block an add()/update() completes (segment-parallel chain, jump decoder).  -m gpu."""
import numpy as np
import pytest

import oracle
from jsref_stream import RefDecoder, RefEncoder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def st():
    import divortio_lz4_b200 as m
    m.default_context()
    from divortio_lz4_b200 import stream
    return stream


def _chunks(data, rng, lo, hi):
    out, p = [], 0
    while p < len(data):
        n = int(rng.randint(lo, hi))
        out.append(data[p:p + n])
        p += n
    return out


def _data(kind, n):
    from divortio_lz4_b200 import corpus
    return {"log": corpus.log(17, n), "mixed": corpus.mixed(18, n), "zero": np.zeros(n, dtype=np.uint8), "rand": corpus.rand(19, n),
            "json": corpus.jsonmsgs(20, 0, n // 4096 + 1)[:n]}[kind].tobytes()


@pytest.mark.parametrize("indep", [False, True])
@pytest.mark.parametrize("kind", ["log", "mixed", "zero", "rand", "json"])
def test_encoder_pieces_equal_the_reference_call_for_call(st, kind, indep):
    rng = np.random.RandomState(11)
    data = _data(kind, 1500000 + 77)
    for bs, cc, dic, lo, hi in ((65536, True, None, 1, 400000), (262144, False, None, 50000, 900000),
                                (65536, True, data[5000:90000], 1, 200000), (1048576, True, data[:300], 100000, 1500000)):
        ref = RefEncoder(bs, indep, cc, dic)
        gpu = st.LZ4Encoder(bs, indep, cc, dic)
        for c in _chunks(data, rng, lo, hi):
            want = [bytes(x) for x in ref.add(c)]
            got = [bytes(x) for x in gpu.add(c)]
            assert got == want, (kind, indep, bs, len(c))
        assert [bytes(x) for x in gpu.finish()] == [bytes(x) for x in ref.finish()]
        assert gpu.finish() == []


def test_encoder_one_large_add_takes_the_segment_engine(st):
    import divortio_lz4_b200 as dl
    data = _data("log", 6 * 1024 * 1024 + 5)
    ref, gpu = RefEncoder(65536, False, True), st.LZ4Encoder(65536, False, True)
    assert [bytes(x) for x in gpu.add(data)] == [bytes(x) for x in ref.add(data)]
    assert dl.default_context().segment_stats[0] >= 8                         # the 96 full blocks went through as segments
    assert [bytes(x) for x in gpu.finish()] == [bytes(x) for x in ref.finish()]
    # a second encoder fed the same bytes in two adds: the table carried between the calls must be the serial loop's
    ref2, gpu2 = RefEncoder(65536, False, False), st.LZ4Encoder(65536, False, False)
    for part in (data[:3 * 1024 * 1024 + 11], data[3 * 1024 * 1024 + 11:]):
        assert [bytes(x) for x in gpu2.add(part)] == [bytes(x) for x in ref2.add(part)]
        assert np.array_equal(gpu2.hashTable, ref2.hash_table)
    assert [bytes(x) for x in gpu2.finish()] == [bytes(x) for x in ref2.finish()]


@pytest.mark.parametrize("indep", [False, True])
def test_decoder_chunks_equal_the_reference(st, indep):
    rng = np.random.RandomState(12)
    for kind in ("log", "mixed", "zero"):
        data = _data(kind, 1300000)
        for bs, cc, dic in ((65536, True, None), (262144, True, data[100:70000])):
            enc = RefEncoder(bs, indep, cc, dic)
            frame = b"".join(enc.add(data) + enc.finish())
            frame = frame + frame                                              # two concatenated frames (lz4Decode.js:262-266)
            ref, gpu = RefDecoder(dic), st.LZ4Decoder(dic)
            for c in _chunks(frame, rng, 1, 200000):
                assert gpu.update(c) == ref.update(c), (kind, indep, bs)
            assert gpu.buffer == b"" and gpu.state == "magic"


def test_decoder_errors(st):
    import divortio_lz4_b200 as dl
    data = _data("log", 300000)
    enc = RefEncoder(65536, False, True)
    frame = bytearray(b"".join(enc.add(data) + enc.finish()))
    frame[-1] ^= 1
    with pytest.raises(dl.LZ4Error, match="Content Checksum Error"):
        st.LZ4Decoder().update(bytes(frame))
    assert b"".join(st.LZ4Decoder(None, False).update(bytes(frame))) == data
    with pytest.raises(dl.LZ4Error, match="Invalid Magic Number"):
        st.LZ4Decoder().update(b"\\x00\\x01\\x02\\x03\\x04")
    enc = RefEncoder(65536, False, False, data[:5000])
    with pytest.raises(dl.LZ4Error, match="requires a Dictionary"):
        st.LZ4Decoder().update(b"".join(enc.add(data) + enc.finish()))


def test_stateful_xxh32_matches_one_shot(st):
    import divortio_lz4_b200 as dl
    rng = np.random.RandomState(13)
    data = _data("mixed", 400000)
    h = dl.XXHash32(0)
    for c in _chunks(data, rng, 1, 5000):
        h.update(c)
    assert h.digest() == oracle.xxh32(data) == dl.xxHash32(data)
    assert dl.XXHash32(0).digest() == 0x02CC5D05                               # tests/xxhash32/xxhash32.test.mjs:13
    assert dl.XXHash32(0).update(b"Hello World").digest() == 0xB1FD16EE         # :20
    ref = bytes((i * 31 + 17) & 0xFF for i in range(1024))                     # xxhash32Stateful.test.mjs:29-46
    assert dl.XXHash32(7).update(ref[:100]).update(ref[100:]).digest() == oracle.xxh32(ref, 7)


def test_worker_offload_returns_futures_with_the_sync_results(st):
    """LZ4.compressWorker / decompressWorker (src/webWorker/workerClient.js:114-152) -- one off-thread worker."""
    import divortio_lz4_b200 as dl
    data = _data("log", 2 * 1024 * 1024 + 9)
    futs = [dl.LZ4.compressWorker(data, {"maxBlockSize": bs, "blockIndependence": ind, "contentChecksum": True})
            for bs, ind in ((65536, True), (4194304, False), (262144, True))]
    want = [oracle.compress_buffer(data, None, bs, ind, True) for bs, ind in ((65536, True), (4194304, False), (262144, True))]
    assert dl.compressBuffer(data, None, 65536, True, True) == want[0]        # the caller's own context stays usable meanwhile
    got = [f.result(timeout=120) for f in futs]
    assert got == want
    backs = [dl.LZ4.decompressWorker(g, {"verifyChecksum": True}) for g in got]
    assert all(b.result(timeout=120) == data for b in backs)
    bad = bytearray(got[0]); bad[-1] ^= 1
    with pytest.raises(dl.LZ4Error, match="Content Checksum Error"):
        dl.LZ4.decompressWorker(bytes(bad)).result(timeout=120)


def test_worker_stream_tasks_pipe_through_the_stream_codec(st):
    """LZ4Worker.compressStream / decompressStream (src/webWorker/workerClient.js:96-110,143-152; worker side
    lz4.worker.js:30-69): readable -> stream codec -> writable on a worker thread, Future resolves when the stream is complete.
    The pieces are the stream encoder's (checked against the restated classes), and several queued tasks all complete."""
    import io
    import divortio_lz4_b200 as dl
    data = _data("mixed", 3 * 1024 * 1024 + 77)
    chunks = [data[i:i + 300000] for i in range(0, len(data), 300000)]
    opts = {"maxBlockSize": 262144, "blockIndependence": False, "contentChecksum": True}
    sinks = [io.BytesIO() for _ in range(3)]
    futs = [dl.LZ4.compressWorkerStream(iter(chunks), s.write, opts) for s in sinks]
    assert all(f.result(timeout=300) is None for f in futs)
    enc = RefEncoder(262144, False, True)
    want = b"".join(bytes(p) for c in chunks for p in enc.add(c)) + b"".join(bytes(p) for p in enc.finish())
    assert all(s.getvalue() == want for s in sinks)
    assert dl.decompressBuffer(want) == data
    back = io.BytesIO()
    frame = sinks[0].getvalue()
    dl.LZ4.decompressWorkerStream((frame[i:i + 123457] for i in range(0, len(frame), 123457)), back.write, {"verifyChecksum": True}).result(timeout=300)
    assert back.getvalue() == data
    bad = bytearray(frame); bad[-1] ^= 1
    with pytest.raises(dl.LZ4Error, match="Content Checksum Error"):
        dl.LZ4.decompressWorkerStream([bytes(bad)], io.BytesIO()).result(timeout=300)
