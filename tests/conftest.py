import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as fh:
        return json.load(fh)


BENCH_REC = ('{"id":1,"type":"benchmark_event","tags":["performance","compression","lz4","javascript","v8"],'
             '"meta":{"valid":true,"scores":[100,205,300,400,500]},'
             '"payload":"Repeated data is the key to high compression ratios in LZ4."}').encode()


def golden_inputs():
    """Same inputs as tests/golden/make_golden.py."""
    return {
        "K1_A10000": b"A" * 10000,
        "K2_hello": b"Hello World",
        "K3_lin1024": bytes((i * 31 + 17) & 0xFF for i in range(1024)),
        "K4_mod256_70000": bytes(i % 256 for i in range(70000)),
        "K5_benchjson_x4789": BENCH_REC * 4789,
        "K7_empty": b"",
        "K8_abcd12": b"abcdabcdabcd",
        "K9_abcd13": b"abcdabcdabcda",
        "K10_zero65536": bytes(65536),
        "T_text": (b"2026-10-18 INFO svc[1001]: request ok path=/api/v1/users status=200\n"
                   b"2026-10-18 WARN db[1003]: slow query path=/api/v1/items status=500\n"
                   b"2026-10-18 INFO svc[1002]: request ok path=/api/v1/items status=201\n") * 40,
    }


GOLDEN_OPTS = {
    "default": dict(),
    "indep64k": dict(max_block_size=65536, block_independence=True),
    "linked64k": dict(max_block_size=65536, block_independence=False),
    "indep64k_cc": dict(max_block_size=65536, block_independence=True, content_checksum=True),
    "indep4m_nosize": dict(max_block_size=4194304, block_independence=True, add_content_size=False),
}


def edge_corpora():
    """Small deterministic inputs that exercise the block codec's corner cases."""
    from divortio_lz4_b200 import corpus
    rng = np.random.RandomState(1234)
    out = {}
    for n in (0, 1, 4, 5, 11, 12, 13, 14, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 270, 271, 272, 300):
        out["abc%d" % n] = rng.choice(np.frombuffer(b"abc", dtype=np.uint8), n).tobytes()
        out["zero%d" % n] = bytes(n)
    out["log64k"] = corpus.log(1, 65536).tobytes()
    out["log_short"] = corpus.log(2, 5000).tobytes()
    out["rand64k"] = corpus.rand(1, 65536).tobytes()
    out["rand1000"] = corpus.rand(2, 1000).tobytes()
    out["zero64k"] = bytes(65536)
    out["zero65535"] = bytes(65535)
    out["bench64k"] = corpus.benchjson(65536).tobytes()
    out["mod256"] = bytes(i % 256 for i in range(65536))
    out["2sym"] = rng.choice(np.frombuffer(b"ab", dtype=np.uint8), 30000).tobytes()
    out["4sym"] = rng.choice(np.frombuffer(b"abcd", dtype=np.uint8), 40000).tobytes()
    r = bytearray(corpus.rand(9, 65536).tobytes())
    r[30000:30100] = r[100:200]
    r[50000:50040] = r[40000:40040]
    r[65000:65300] = r[64000:64300]
    out["rand_with_repeats"] = bytes(r)
    # long literal runs then a long match (exercises 255-run length bytes on both nibbles)
    out["lit_then_match"] = corpus.rand(3, 4000).tobytes() + corpus.rand(3, 4000).tobytes()
    # periods 1..40 (overlapping match copies with every small offset)
    out["periods"] = b"".join(bytes((j * 7 + p) & 0xFF for j in range(p)) * (600 // p + 3) for p in range(1, 41))
    return out
