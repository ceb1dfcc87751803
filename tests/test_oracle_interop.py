"""Cross-checks with liblz4 1.9.4 (third party): every oracle frame decodes in LZ4F_decompress and every liblz4 frame
(CLI-equivalent preferences) decodes in the oracle's decompressBuffer restatement.  CPU only."""
import numpy as np
import pytest

import lz4f
import oracle
from conftest import edge_corpora

pytestmark = pytest.mark.skipif(not lz4f.available(), reason="liblz4.so.1 not present")


def test_oracle_frames_decode_in_liblz4():
    for name, data in edge_corpora().items():
        for indep in (False, True):
            for cc in (False, True):
                for bc in (False, True):
                    f = oracle.compress_buffer(data, None, 65536, indep, cc, True, None, bc)
                    assert lz4f.decompress_frame(f, len(data)) == data, (name, indep, cc, bc)


def test_oracle_multiblock_frames_decode_in_liblz4():
    from divortio_lz4_b200 import corpus
    data = corpus.mixed(11, 700000).tobytes()
    for bs in (65536, 262144):
        for indep in (False, True):
            f = oracle.compress_buffer(data, None, bs, indep, True, True, None, True)
            assert lz4f.decompress_frame(f, len(data)) == data


def test_oracle_blocks_decode_in_liblz4():
    for name, data in edge_corpora().items():
        assert lz4f.decompress_block(oracle.compress_block_bytes(data), len(data)) == data, name


def test_liblz4_frames_decode_in_oracle():
    from divortio_lz4_b200 import corpus
    data = corpus.mixed(12, 600000).tobytes()
    for bsid in (4, 5, 7):
        for linked in (False, True):
            for size in (False, True):
                for bc in (False, True):
                    f = lz4f.compress_frame(data, bsid, linked, True, size, bc)
                    assert oracle.decompress_buffer(f) == data, (bsid, linked, size, bc)


def test_xxh32_matches_libxxhash_on_corpora():
    xxhash = pytest.importorskip("xxhash")
    for name, data in edge_corpora().items():
        assert oracle.xxh32(data) == xxhash.xxh32(data, seed=0).intdigest(), name
