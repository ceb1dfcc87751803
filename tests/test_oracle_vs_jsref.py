"""Differential test: C oracle vs the literal Python transliteration (tests/jsref.py) on seeded inputs.  CPU only."""
import numpy as np
import pytest

import jsref
import oracle
from conftest import edge_corpora


def _jsblock(data, start=0, table=None):
    src = jsref.U8(data)
    n = len(data) - start
    out = jsref.U8(bytes(n + n // 255 + 16))
    t = [0] * 16384 if table is None else table
    k = jsref.js_compress_block(src, out, start, n, t, 0)
    return bytes(out.b[:k]), t


def test_blocks_match_on_edge_corpora():
    for name, data in edge_corpora().items():
        if len(data) > 40000:
            data = data[:40000]                # jsref is pure Python
        got = oracle.compress_block_bytes(data)
        want, _ = _jsblock(data)
        assert got == want, name


def test_table_state_carries_over_like_the_js():
    rng = np.random.RandomState(5)
    data = rng.choice(np.frombuffer(b"abcdefgh ", dtype=np.uint8), 9000).tobytes()
    t_c = oracle.new_table()
    t_js = [0] * 16384
    src = jsref.U8(data)
    for start, n in ((0, 3000), (3000, 3000), (6000, 3000)):
        k_c, out_c = oracle.compress_block(data, start, n, t_c)
        out_js = jsref.U8(bytes(n + n // 255 + 16))
        k_js = jsref.js_compress_block(src, out_js, start, n, t_js, 0)
        assert k_c == k_js and out_c[:k_c].tobytes() == bytes(out_js.b[:k_js])
        assert t_c.tolist() == t_js


def test_negative_filled_table_is_empty():
    # tests/raw/raw.test.mjs:14 passes a -1 filled table
    data = b"hello hello hello hello hello hello"
    t = np.full(16384, -1, dtype=np.int32)
    assert oracle.compress_block_bytes(data, table=t) == oracle.compress_block_bytes(data)


def test_undersized_output_drops_stores_like_typed_arrays():
    data = b"abcabcabcabcabcabcabcabcabcabcabcabcabcabc" * 10
    full = oracle.compress_block_bytes(data)
    small = np.zeros(10, dtype=np.uint8)
    n, _ = oracle.compress_block(data, 0, len(data), None, small)
    assert n == len(full) and small.tobytes() == full[:10]


@pytest.mark.parametrize("seed", range(12))
def test_frames_match_random_options(seed):
    rng = np.random.RandomState(seed)
    n = int(rng.randint(0, 6000))
    alphabet = np.frombuffer(b"ab" if seed % 3 == 0 else b"abcdefghijklmnop \n", dtype=np.uint8)
    data = rng.choice(alphabet, n).tobytes()
    dictionary = rng.choice(alphabet, int(rng.randint(4, 400))).tobytes() if seed % 2 else None
    indep, cc, size = bool(rng.randint(2)), bool(rng.randint(2)), bool(rng.randint(2))
    got = oracle.compress_buffer(data, dictionary, 65536, indep, cc, size)
    want = jsref.js_compress_buffer(data, dictionary, 65536, indep, cc, size)
    assert got == want
    assert oracle.decompress_buffer(got, dictionary) == data


def test_multiblock_frames_match():
    rng = np.random.RandomState(99)
    data = rng.choice(np.frombuffer(b"abcdefgh", dtype=np.uint8), 70000).tobytes() + bytes(3000)
    for indep in (False, True):
        assert oracle.compress_buffer(data, None, 65536, indep) == jsref.js_compress_buffer(data, None, 65536, indep)


def test_xxh32_matches_jsref():
    rng = np.random.RandomState(3)
    for n in list(range(0, 40)) + [63, 64, 65, 1000]:
        d = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        for seed in (0, 12345):
            assert oracle.xxh32(d, seed) == jsref.js_xxh32(d, seed)


def test_spec_decoder_equals_literal_js_decoder_when_the_tail_defect_is_not_hit():
    # matches of length >= 8 only (period-3 data): the literal JS decoder and the spec decoder agree
    data = b"xyz" * 3000 + bytes(500) + b"0123456789abcdef" * 100
    f = oracle.compress_buffer(data, None, 65536, True)
    assert jsref.js_decompress_buffer(f) == data == oracle.decompress_buffer(f)
