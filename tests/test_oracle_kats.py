"""Pins the C oracle against the reference's own golden vectors / KATs, the SURVEY A.3 scratch KATs and the
jsref-generated fixtures (tests/golden/kats.json).  CPU only."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_OPTS, golden_inputs


def test_reference_xxh32_kats(kats):
    # tests/xxhash32/xxhash32.test.mjs:13,20
    for k in kats["reference_golden"]["xxh32"]:
        assert oracle.xxh32(k["text"].encode(), k["seed"]) == k["hash"]


def test_reference_golden_decode_frames(kats):
    # tests/golden.test.mjs:23,39,52
    for k in kats["reference_golden"]["decode_frames"]:
        assert oracle.decompress_buffer(bytes.fromhex(k["hex"])) == k["text"].encode()


def test_header_checksum_bytes_from_reference_golden():
    # HC(60 40)=0x82, HC(60 70)=0x73, HC(64 40)=0xA7
    for flg_bd, hc in ((b"\x60\x40", 0x82), (b"\x60\x70", 0x73), (b"\x64\x40", 0xA7)):
        assert (oracle.xxh32(flg_bd) >> 8) & 0xFF == hc


def test_xxh32_stateful_pattern_matches_libxxhash():
    xxhash = pytest.importorskip("xxhash")
    data = bytes((i * 31 + 17) & 0xFF for i in range(1024))      # tests/xxhash32/xxhash32Stateful.test.mjs:8-14
    for n in list(range(0, 70)) + [255, 256, 1000, 1024]:
        for seed in (0, 1, 0x9E3779B1):
            assert oracle.xxh32(data[:n], seed) == xxhash.xxh32(data[:n], seed=seed).intdigest()


SURVEY_KATS = [  # (input key, opts key, frame length, xxh32(frame) or None, hex or None)
    ("K1_A10000", "default", 73, None, "04224d1848701027000000000000c2320000001f410100" + "ff" * 39 + "1e50414141414100000000"),
    ("K2_hello", "indep64k", 34, None, "04224d1868400b00000000000000580b00008048656c6c6f20576f726c6400000000"),
    ("K3_lin1024", "default", 294, 0x5254C232, None),
    ("K4_mod256_70000", "linked64k", 580, 0x7145F0C8, None),
    ("K4_mod256_70000", "indep64k", 836, 0x205863BD, None),
    ("K4_mod256_70000", "indep64k_cc", 840, 0x90993376, None),
    ("K5_benchjson_x4789", "indep64k", 7977, 0x973F5184, None),
    ("K5_benchjson_x4789", "linked64k", 6267, 0x2EA87190, None),
    ("K7_empty", "default", 19, 0x18C13F36, None),
    ("K8_abcd12", "default", 35, 0x61A12BB2, None),
    ("K9_abcd13", "default", 36, 0x18D88818, "04224d1848700d00000000000000d20d0000806162636461626364616263646100000000"),
    ("K10_zero65536", "indep64k", 290, 0xFE935BF2, None),
]


@pytest.mark.parametrize("inp,opt,length,digest,hexs", SURVEY_KATS)
def test_survey_kats(inp, opt, length, digest, hexs):
    f = oracle.compress_buffer(golden_inputs()[inp], None, **GOLDEN_OPTS[opt])
    assert len(f) == length
    if digest is not None:
        assert oracle.xxh32(f) == digest
    if hexs is not None:
        assert f.hex() == hexs


def test_survey_k5_4m_independent():
    f = oracle.compress_buffer(golden_inputs()["K5_benchjson_x4789"], None, 4194304, True, False, True)
    assert len(f) == 4357 and oracle.xxh32(f) == 0xEF687FE6


def test_survey_d1_dictionary_is_not_matched():
    s = b"CommonPrefix_SharedData_Reference_1234567890_UniquePartA"
    with_dict = oracle.compress_buffer(s, s[:44])
    assert len(with_dict) == 83 and len(oracle.compress_buffer(s)) == 79
    assert with_dict[4] == 0x49 and with_dict[14:18] == (0xCA508D42).to_bytes(4, "little") and with_dict[18] == 0xCE


def test_fixture_frames(kats):
    ins = golden_inputs()
    for rec in kats["frames"]:
        f = oracle.compress_buffer(ins[rec["input"]], None, **GOLDEN_OPTS[rec["opts"]])
        assert len(f) == rec["len"], rec
        assert oracle.xxh32(f) == rec["xxh32"], rec
        if "hex" in rec:
            assert f.hex() == rec["hex"]
        assert oracle.decompress_buffer(f) == ins[rec["input"]]


def test_fixture_blocks(kats):
    ins = golden_inputs()
    for rec in kats["blocks"]:
        b = oracle.compress_block_bytes(ins[rec["input"]])
        assert len(b) == rec["len"] and oracle.xxh32(b) == rec["xxh32"], rec


def test_fixture_dictionary_frames(kats):
    for rec in kats["dictionary"]:
        data = bytes.fromhex(rec["input_hex"])
        d = bytes.fromhex(rec["dict_hex"]) if "dict_hex" in rec else data[:rec["dict_len"]]
        if rec["name"] == "D1":
            f = oracle.compress_buffer(data, d)
        elif rec["name"].endswith("indep_cc"):
            f = oracle.compress_buffer(data, d, 65536, True, True, True)
        else:
            f = oracle.compress_buffer(data, d, 65536, False, False, True)
        assert f.hex() == rec["hex"], rec["name"]
        assert oracle.decompress_buffer(f, d) == data


# behaviour pinned by the reference's tests
def test_behaviour_10000_A_compresses_small():          # tests/buffer/bufferCompress.test.mjs:17-24
    assert len(oracle.compress_buffer(b"A" * 10000)) < 100


def test_behaviour_content_checksum_adds_4_bytes():      # tests/buffer/bufferCompress.test.mjs:27-36
    d = b"checksum me " * 50
    assert len(oracle.compress_buffer(d, None, 4194304, False, True)) == len(oracle.compress_buffer(d)) + 4


def test_behaviour_checksum_error_and_magic():          # tests/buffer/bufferDecompress.test.mjs:24-56
    d = b"some payload that repeats, some payload that repeats" * 20
    f = bytearray(oracle.compress_buffer(d, None, 4194304, False, True))
    f[-1] ^= 0xFF
    with pytest.raises(oracle.OracleError, match="Checksum Error"):
        oracle.decompress_buffer(bytes(f))
    assert oracle.decompress_buffer(bytes(f), None, False) == d
    with pytest.raises(oracle.OracleError, match="Invalid Magic"):
        oracle.decompress_buffer(b"\x00\x01\x02\x03\x04\x05\x06\x07")
    with pytest.raises(oracle.OracleError, match="Unsupported Version"):
        oracle.decompress_buffer(bytes.fromhex("04224D18A0400000"))


def test_behaviour_random_roundtrip_64k_plus_500():      # tests/buffer/bufferDecompress.test.mjs:16-22
    d = np.random.RandomState(7).randint(0, 256, 65536 + 500).astype(np.uint8).tobytes()
    for indep in (False, True):
        assert oracle.decompress_buffer(oracle.compress_buffer(d, None, 65536, indep)) == d


def test_block_decoder_errors():
    out = np.zeros(16, dtype=np.uint8)
    with pytest.raises(oracle.OracleError, match="Output Buffer Too Small"):
        oracle.decompress_block(b"\xF0\x20" + b"x" * 47, 0, 49, out)
    with pytest.raises(oracle.OracleError, match="Malformed Input"):
        oracle.decompress_block(b"\x50abc", 0, 4, out)
    with pytest.raises(oracle.OracleError, match="Invalid Offset 0"):
        oracle.decompress_block(b"\x10a\x00\x00\x00", 0, 5, out)
    with pytest.raises(oracle.OracleError, match="Dictionary Offset Out of Bounds"):
        oracle.decompress_block(b"\x10a\x05\x00\x00", 0, 5, out)
    # the same reference resolved by a dictionary (index dictLen + copySrc, blockDecompress.js:147)
    n = oracle.decompress_block(b"\x10a\x05\x00\x00", 0, 5, out, 0, b"WXYZ")
    assert n == 5 and out[:n].tobytes() == b"aWXYZ"
