"""CPU: the restatement of the reference's streaming codec (tests/jsref_stream.py, src/shared/lz4Encode.js / lz4Decode.js) is
pinned the only way the reference allows (its own stream tests are stale): every stream it writes must be a valid
LZ4 frame that the buffer decoder restatement and liblz4 decode to the input, whatever the chunking; its header bytes are
those of lz4Encode.js:61-94; and its decoder must invert it chunk for chunk."""
import numpy as np
import pytest

import oracle
from jsref_stream import RefDecoder, RefEncoder


def _chunks(data, rng, lo=1, hi=300000):
    out, p = [], 0
    while p < len(data):
        n = int(rng.randint(lo, hi))
        out.append(data[p:p + n])
        p += n
    return out


def _data(kind, n):
    from divortio_lz4_b200 import corpus
    return {"log": corpus.log(7, n), "mixed": corpus.mixed(8, n), "zero": np.zeros(n, dtype=np.uint8), "rand": corpus.rand(9, n)}[kind].tobytes()


@pytest.mark.parametrize("indep", [False, True])
@pytest.mark.parametrize("kind", ["log", "mixed", "zero", "rand"])
def test_stream_frames_are_valid_and_round_trip(kind, indep):
    rng = np.random.RandomState(3)
    data = _data(kind, 700001)
    for bs, cc, dic in ((65536, True, None), (262144, False, None), (65536, True, data[1000:50000])):
        enc = RefEncoder(bs, indep, cc, dic)
        pieces = []
        for c in _chunks(data, rng):
            pieces += enc.add(c)
        pieces += enc.finish()
        frame = b"".join(pieces)
        # header as lz4Encode.js:61-94 writes it: no content size, version 1
        assert frame[:4] == bytes.fromhex("04224d18") and (frame[4] >> 6) == 1 and not (frame[4] & 0x08)
        assert bool(frame[4] & 0x20) == indep and bool(frame[4] & 0x04) == cc and bool(frame[4] & 0x01) == bool(dic)
        if not (indep and dic):       # an independent stream never references its dictionary; the buffer decoder needs none either
            assert oracle.decompress_buffer(frame, dic) == data
        else:
            assert oracle.decompress_buffer(frame, dic) == data
        dec = RefDecoder(dic)
        got = []
        for c in _chunks(frame, rng, 1, 90000):
            got += dec.update(c)
        assert b"".join(got) == data
        assert all(len(g) <= bs for g in got)


def test_stream_frames_decode_in_liblz4():
    import lz4f
    if not lz4f.available():
        pytest.skip("liblz4 not present")
    rng = np.random.RandomState(4)
    data = _data("mixed", 900000)
    for indep in (False, True):
        enc = RefEncoder(65536, indep, True)
        pieces = []
        for c in _chunks(data, rng):
            pieces += enc.add(c)
        pieces += enc.finish()
        assert lz4f.decompress_frame(b"".join(pieces), len(data)) == data


def test_empty_and_tiny_streams():
    enc = RefEncoder()
    assert enc.add(b"") == []
    f = b"".join(enc.finish())
    assert f == bytes.fromhex("04224d18407000") [:6] + f[6:7] + bytes(4) and oracle.decompress_buffer(f) == b""
    enc = RefEncoder(65536, True)
    f = b"".join(enc.add(b"abc") + enc.finish())
    assert oracle.decompress_buffer(f) == b"abc"
