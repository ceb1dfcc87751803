"""Documents a defect of the reference decoder and shows the oracle's spec decoder differs from the literal JS only there.

src/block/blockDecompress.js:219-250: for an in-buffer match with offset >= 8 the "double-copy tail" writes the 8 bytes
ending at endMatch unconditionally (the literal path has a `literalLen >= 8` guard, the match path has none).  With
matchLen 4..7 that window starts BEFORE the match (tailOut = endMatch-8 < outPos), so up to 4 already-decoded bytes are
overwritten with output[tailOut-offset ..].  The reference's own round trip therefore fails on ordinary text; its tests
only round-trip random bytes (no matches), runs of one byte (offset 1) and a periodic payload (long matches).
The north star requires exact round trip, so the product decoder follows the LZ4 block format, not this defect.
"""
import numpy as np

import jsref
import oracle

TEXT = (b"2026-10-18 INFO svc[1001]: request ok path=/api/v1/users status=200\n"
        b"2026-10-18 WARN db[1003]: slow query path=/api/v1/items status=500\n"
        b"2026-10-18 INFO svc[1002]: request ok path=/api/v1/items status=201\n")


def _block_of(frame):
    size = int.from_bytes(frame[15:19], "little")
    return frame[19:19 + size]


def test_literal_js_decoder_corrupts_short_matches():
    frame = oracle.compress_buffer(TEXT, None, 65536, True)
    assert oracle.decompress_buffer(frame) == TEXT                       # spec semantics: exact
    got = jsref.js_decompress_buffer(frame)                              # literal JS semantics
    assert got != TEXT
    bad = [i for i in range(len(TEXT)) if got[i] != TEXT[i]]
    assert 0 < len(bad) <= 8


def test_c_literal_decoder_equals_python_literal_decoder():
    frame = oracle.compress_buffer(TEXT, None, 65536, True)
    blk = _block_of(frame)
    out_c = np.zeros(len(TEXT), dtype=np.uint8)
    oracle.decompress_block(blk, 0, len(blk), out_c, 0, None, literal=True)
    out_py = jsref.U8(bytes(len(TEXT)))
    jsref.js_decompress_block(jsref.U8(blk), 0, len(blk), out_py, 0, None)
    assert out_c.tobytes() == bytes(out_py.b)


def test_divergence_is_confined_to_short_matches_at_offset_ge_8():
    """Hand-built blocks: literal and spec decoders agree unless (4 <= matchLen <= 7 and offset >= 8)."""
    base = bytes(range(65, 65 + 20))                                    # 20 literals
    for offset in (1, 2, 3, 4, 7, 8, 9, 15, 20):
        for mlen in (4, 5, 7, 8, 9, 16, 17, 40):
            tok = (15 << 4) | min(mlen - 4, 15)
            blk = bytes([tok, 20 - 15]) + base + bytes([offset, 0])
            if mlen - 4 >= 15:
                blk += bytes([mlen - 4 - 15])
            blk += bytes([0x10]) + b"!"                                   # last sequence: 1 literal
            spec = np.zeros(128, dtype=np.uint8)
            lit = np.zeros(128, dtype=np.uint8)
            n1 = oracle.decompress_block(blk, 0, len(blk), spec)
            n2 = oracle.decompress_block(blk, 0, len(blk), lit, literal=True)
            assert n1 == n2 == 20 + mlen + 1
            same = spec.tobytes() == lit.tobytes()
            assert same == (not (4 <= mlen <= 7 and offset >= 8)), (offset, mlen)
