"""Generates tests/golden/kats.json from tests/jsref.py (the literal Python transliteration of the reference's
JavaScript -- NOT from the C oracle), so the C oracle and the CUDA path are both checked against a second,
independent statement.  The reference itself cannot run here (no JS engine in the image).

Also embeds verbatim the golden vectors held by the reference's own tests:
  tests/golden.test.mjs:23,39,52 (three decode frames) and tests/xxhash32/xxhash32.test.mjs:13,20 (two xxh32 KATs).

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import jsref  # noqa: E402

BENCH_REC = ('{"id":1,"type":"benchmark_event","tags":["performance","compression","lz4","javascript","v8"],'
             '"meta":{"valid":true,"scores":[100,205,300,400,500]},'
             '"payload":"Repeated data is the key to high compression ratios in LZ4."}').encode()


def inputs():
    yield "K1_A10000", b"A" * 10000
    yield "K2_hello", b"Hello World"
    yield "K3_lin1024", bytes((i * 31 + 17) & 0xFF for i in range(1024))
    yield "K4_mod256_70000", bytes(i % 256 for i in range(70000))
    yield "K5_benchjson_x4789", BENCH_REC * 4789
    yield "K7_empty", b""
    yield "K8_abcd12", b"abcdabcdabcd"
    yield "K9_abcd13", b"abcdabcdabcda"
    yield "K10_zero65536", bytes(65536)
    yield "T_text", (b"2026-10-18 INFO svc[1001]: request ok path=/api/v1/users status=200\n"
                     b"2026-10-18 WARN db[1003]: slow query path=/api/v1/items status=500\n"
                     b"2026-10-18 INFO svc[1002]: request ok path=/api/v1/items status=201\n") * 40


OPTS = [
    ("default", dict()),
    ("indep64k", dict(max_block=65536, indep=True)),
    ("linked64k", dict(max_block=65536, indep=False)),
    ("indep64k_cc", dict(max_block=65536, indep=True, content_cksum=True)),
    ("indep4m_nosize", dict(max_block=4194304, indep=True, add_size=False)),
]


def main():
    out = {"generator": "tests/golden/make_golden.py (tests/jsref.py)", "reference_golden": {
        "decode_frames": [
            {"hex": "04224D186040820B00008048656c6c6f20576f726c6400000000", "text": "Hello World", "src": "tests/golden.test.mjs:23"},
            {"hex": "04224D1860707300000000", "text": "", "src": "tests/golden.test.mjs:39"},
            {"hex": "04224D186440A70B00008048656c6c6f20576f726c6400000000EE16FDB1", "text": "Hello World", "src": "tests/golden.test.mjs:52"},
        ],
        "xxh32": [
            {"text": "", "seed": 0, "hash": 0x02CC5D05, "src": "tests/xxhash32/xxhash32.test.mjs:13"},
            {"text": "Hello World", "seed": 0, "hash": 0xB1FD16EE, "src": "tests/xxhash32/xxhash32.test.mjs:20"},
        ]}, "frames": [], "blocks": [], "dictionary": []}
    for name, data in inputs():
        for oname, o in OPTS:
            f = jsref.js_compress_buffer(data, None, o.get("max_block", 4194304), o.get("indep", False),
                                         o.get("content_cksum", False), o.get("add_size", True))
            rec = {"input": name, "opts": oname, "len": len(f), "xxh32": jsref.js_xxh32(f)}
            if len(f) <= 400:
                rec["hex"] = f.hex()
            out["frames"].append(rec)
        # raw block with a fresh table
        src = jsref.U8(data)
        o = jsref.U8(bytes(len(data) + len(data) // 255 + 16))
        n = jsref.js_compress_block(src, o, 0, len(data), [0] * 16384, 0)
        blk = bytes(o.b[:n])
        out["blocks"].append({"input": name, "len": n, "xxh32": jsref.js_xxh32(blk)})
    s = b"CommonPrefix_SharedData_Reference_1234567890_UniquePartA"
    f = jsref.js_compress_buffer(s, s[:44])
    out["dictionary"].append({"name": "D1", "input_hex": s.hex(), "dict_len": 44, "hex": f.hex()})
    f = jsref.js_compress_buffer((s * 50), s[:44] * 3, 65536, True, True, True)
    out["dictionary"].append({"name": "D1x50_indep_cc", "input_hex": (s * 50).hex(), "dict_hex": (s[:44] * 3).hex(), "hex": f.hex()})
    f = jsref.js_compress_buffer((s * 50), s[:44] * 3, 65536, False, False, True)
    out["dictionary"].append({"name": "D1x50_linked", "input_hex": (s * 50).hex(), "dict_hex": (s[:44] * 3).hex(), "hex": f.hex()})
    with open(os.path.join(HERE, "kats.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["frames"]), "frames,", len(out["blocks"]), "blocks")


if __name__ == "__main__":
    main()
