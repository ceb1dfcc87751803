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


# Header vectors derived by hand from the reference's writer (src/shared/lz4Encode.js:61-94: magic, FLG = version 1 << 6 |
# blockIndependence 0x20 | contentChecksum 0x04 | dictId 0x01, BD = (bdId & 7) << 4, dictID little-endian, HC = second byte of
# xxh32(FLG..dictID)) with the HC byte computed by the SYSTEM libxxhash (XXH32), not by the oracle; the first two are the
# well-known headers every lz4 CLI file starts with.
STREAM_HEADER_VECTORS = [
    ((False, False, 7, None), "04224d184070df"),
    ((True, True, 4, None), "04224d186440a7"),
    ((True, False, 7, None), "04224d18607073"),
    ((False, True, 5, None), "04224d184450e6"),
    ((False, False, 7, 0x12345678), "04224d18417078563412d9"),
    ((True, True, 6, 0xDEADBEEF), "04224d186560efbeaddecc"),
]


@pytest.mark.parametrize("args,want", STREAM_HEADER_VECTORS)
def test_stream_header_bytes_match_hand_derived_vectors(args, want):
    import jsref_stream
    assert jsref_stream.create_frame_header(*args).hex() == want


def test_stream_frames_with_a_dictionary_decode_in_liblz4():
    """The dictId case end to end: a linked stream primed with a dictionary (lz4Encode.js:140-168) is decoded by liblz4's
    LZ4F_decompress_usingDict (which also verifies the header checksum byte and the content checksum); the dictID field is
    xxh32(dictionary) (lz4Encode.js:120)."""
    import lz4f
    if not lz4f.available():
        pytest.skip("liblz4 not present")
    rng = np.random.RandomState(5)
    data = _data("log", 600000)
    dic = _data("log", 150000)[20000:120000]
    for indep in (False, True):
        enc = RefEncoder(65536, indep, True, dic)
        pieces = []
        for c in _chunks(data, rng):
            pieces += enc.add(c)
        pieces += enc.finish()
        frame = b"".join(pieces)
        assert frame[4] & 0x01 and int.from_bytes(frame[6:10], "little") == oracle.xxh32(dic)
        assert lz4f.decompress_frame(frame, len(data), dictionary=dic) == data
