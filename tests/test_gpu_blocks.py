"""GPU parity, raw block layer: the CUDA path (through the C ABI) vs the oracle, bit-exact.  Run with -m gpu."""
import numpy as np
import pytest

import oracle
from conftest import edge_corpora, golden_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dl():
    import divortio_lz4_b200 as m
    m.default_context()          # raises if the CUDA library or device is missing: no silent fallback
    return m


def _pack(items):
    """Concatenate byte strings; returns (buffer, off, len) with deliberately odd alignment."""
    off, ln, parts, pos = [], [], [], 0
    for k, d in enumerate(items):
        pad = (k * 7 + 3) % 13
        parts.append(bytes(pad))
        pos += pad
        off.append(pos)
        ln.append(len(d))
        parts.append(d)
        pos += len(d)
    return np.frombuffer(b"".join(parts) + bytes(8), dtype=np.uint8), np.array(off, dtype=np.uint64), np.array(ln, dtype=np.uint32)


def _check_blocks(dl, items, **kw):
    buf, off, ln = _pack(items)
    dst, doff, clen = dl.compress_blocks(buf, off, ln, **kw)
    for i, d in enumerate(items):
        got = dst[int(doff[i]):int(doff[i]) + int(clen[i])].tobytes()
        yield i, got


def test_compress_edge_corpora_bit_exact(dl):
    corp = edge_corpora()
    names = list(corp)
    for i, got in _check_blocks(dl, [corp[n] for n in names]):
        want = oracle.compress_block_bytes(corp[names[i]])
        assert got == want, names[i]


def test_compress_golden_blocks(dl, kats):
    ins = golden_inputs()
    recs = [r for r in kats["blocks"] if len(ins[r["input"]]) <= 65536]
    for i, got in _check_blocks(dl, [ins[r["input"]] for r in recs]):
        assert len(got) == recs[i]["len"] and oracle.xxh32(got) == recs[i]["xxh32"], recs[i]["input"]


@pytest.mark.parametrize("kind", ["mixed", "log", "rand", "bench"])
def test_compress_64k_blocks_of_corpus(dl, kind):
    from divortio_lz4_b200 import corpus
    n = 8 * 1024 * 1024 + 12345
    data = {"mixed": lambda: corpus.mixed(2, n), "log": lambda: corpus.log(3, n), "rand": lambda: corpus.rand(4, n),
            "bench": lambda: corpus.benchjson(n)}[kind]()
    off = np.arange(0, n, 65536, dtype=np.uint64)
    ln = np.minimum(65536, n - off).astype(np.uint32)
    dst, doff, clen = dl.compress_blocks(data, off, ln)
    odst, odoff, oclen = oracle.compress_blocks(data, off, ln)
    assert np.array_equal(clen, oclen)
    for i in range(len(off)):
        a = dst[int(doff[i]):int(doff[i]) + int(clen[i])]
        b = odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])]
        assert np.array_equal(a, b), (kind, i)


@pytest.mark.parametrize("bs", [262144, 1048576, 4194304])
def test_compress_large_blocks_use_int32_table(dl, bs):
    from divortio_lz4_b200 import corpus
    n = 3 * bs + 777
    data = corpus.mixed(bs, n)
    off = np.arange(0, n, bs, dtype=np.uint64)
    ln = np.minimum(bs, n - off).astype(np.uint32)
    dst, doff, clen = dl.compress_blocks(data, off, ln)
    odst, odoff, oclen = oracle.compress_blocks(data, off, ln)
    assert np.array_equal(clen, oclen)
    for i in range(len(off)):
        assert np.array_equal(dst[int(doff[i]):int(doff[i]) + int(clen[i])], odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])]), i


@pytest.mark.parametrize("bs", [262144, 1048576, 4194304])
def test_decompress_large_blocks_batch_api_jump_decoder_and_fallbacks(dl, bs):
    """Few large blocks through the batch API: uniform batches take the jump decoder (host and device variants), ragged ones
    (a short inner block, scattered destinations) and a malformed block take one warp per block; same bytes, same statuses."""
    import torch
    from divortio_lz4_b200 import corpus, device as dev
    n = 5 * bs + 4321
    data = corpus.mixed(bs + 1, n)
    off = np.arange(0, n, bs, dtype=np.uint64)
    ln = np.minimum(bs, n - off).astype(np.uint32)
    dst, doff, clen = oracle.compress_blocks(data, off, ln)
    out, olen, st = dl.decompress_blocks(dst, doff, clen, off, ln)
    assert not st.any() and np.array_equal(olen, ln) and np.array_equal(out[:n], data)
    # device variant
    ctx = dl.default_context()
    d = torch.device("cuda", ctx.device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    d_out = torch.zeros(n + 64, dtype=torch.uint8, device=d)
    d_olen = torch.zeros(len(off), dtype=torch.int32, device=d)
    d_st = torch.zeros(len(off), dtype=torch.uint8, device=d)
    dev.decompress_blocks_dev(ctx, t(dst), t(doff.view(np.int64)), t(clen.view(np.int32)), d_out, t(off.view(np.int64)), t(ln.view(np.int32)), d_olen, d_st)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0 and np.array_equal(d_out[:n].cpu().numpy(), data)
    # a short inner block: blocks 0..5 where block 2 holds only half a block of data, destinations still i * bs
    pieces = [data[i * bs:(i + 1) * bs] for i in range(2)] + [data[2 * bs:2 * bs + bs // 2]] + [data[3 * bs:4 * bs]]
    poff = np.zeros(len(pieces), dtype=np.uint64)
    pl = np.array([p.size for p in pieces], dtype=np.uint32)
    poff[1:] = np.cumsum(pl.astype(np.uint64))[:-1]
    src = np.concatenate(pieces)
    cdst, cdoff, cclen = oracle.compress_blocks(src, poff, pl)
    where = np.arange(len(pieces), dtype=np.uint64) * bs
    caps = np.full(len(pieces), bs, dtype=np.uint32)
    out2, olen2, st2 = dl.decompress_blocks(cdst, cdoff, cclen, where, caps)
    assert not st2.any() and np.array_equal(olen2, pl)
    for i, p in enumerate(pieces):
        assert np.array_equal(out2[int(where[i]):int(where[i]) + p.size], p), i
    # a malformed block among good ones: its own status, the others decode
    bad = dst.copy()
    bad[int(doff[1]) + int(clen[1]) // 2:int(doff[1]) + int(clen[1])] = 0
    out3, olen3, st3 = dl.decompress_blocks(bad, doff, clen, off, ln, check=False)
    oo, ool, ost = oracle.decompress_blocks(bad, doff, clen, off, ln)
    assert np.array_equal(st3.astype(np.int64), np.abs(ost.astype(np.int64)))      # the oracle reports -status
    for i in range(len(off)):
        if not ost[i]:
            assert np.array_equal(out3[int(off[i]):int(off[i]) + int(ln[i])], data[int(off[i]):int(off[i]) + int(ln[i])]), i


def test_compress_with_shared_prefix_three_table_modes(dl):
    """BASELINE config 4 at test size: 4 KiB JSON messages with a 64 KiB dictionary as shared prefix."""
    from divortio_lz4_b200 import corpus
    nmsg = 512
    msgs = corpus.jsonmsgs(4, 100, nmsg)
    dic = corpus.json_dictionary(44)
    off = np.arange(nmsg, dtype=np.uint64) * 4096
    ln = np.full(nmsg, 4096, dtype=np.uint32)
    work = np.concatenate([dic, np.zeros(8, dtype=np.uint8)])
    tables = {
        "none": (dl.WARM_NONE, None, None),
        "jenkins": (dl.WARM_JENKINS, None, oracle.warm_table_jenkins(work, dic.size)),
    }
    primed = oracle.new_table()
    oracle.compress_block(dic, 0, dic.size, primed)       # table state a raw-API user leaves behind
    tables["primed"] = (dl.WARM_TABLE, primed, primed)
    sizes = {}
    for mode, (warm, init, oracle_table) in tables.items():
        dst, doff, clen = dl.compress_blocks(msgs, off, ln, prefix=dic, warm=warm, init_table=init)
        odst, odoff, oclen = oracle.compress_blocks_prefix(dic, oracle_table, msgs, off, ln)
        assert np.array_equal(clen, oclen), mode
        for i in range(nmsg):
            assert np.array_equal(dst[int(doff[i]):int(doff[i]) + int(clen[i])],
                                  odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])]), (mode, i)
        sizes[mode] = int(clen.sum())
        # packed output through the chunked pipeline: the same blocks back to back
        pdst, pdoff, pclen = dl.compress_blocks(msgs, off, ln, prefix=dic, warm=warm, init_table=init, packed=True)
        assert np.array_equal(pclen, oclen), mode
        want = np.concatenate([odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])] for i in range(nmsg)])
        assert np.array_equal(pdst, want), mode
        pout, polen, pstatus = dl.decompress_blocks(pdst, None, pclen, off, ln, dictionary=dic, hist_mode=dl.HIST_RAW)
        assert not pstatus.any() and np.array_equal(pout[:nmsg * 4096], msgs)
        # decode side: every message has its own output base whose index 0 is the dictionary boundary
        out, olen, status = dl.decompress_blocks(dst, doff, clen, off, ln, dictionary=dic, hist_mode=dl.HIST_RAW)
        assert not status.any() and np.array_equal(olen, ln)
        assert np.array_equal(out[:nmsg * 4096], msgs)
    assert sizes["primed"] < sizes["none"]                # only the kernel's own hash makes the prefix useful


def test_overlay_tables_ragged_messages_odd_prefixes(dl):
    """Small blocks behind a shared prefix with a warmed / primed table (k_compress_overlay): ragged lengths, prefixes of odd
    lengths and alignments, enough messages for every warp to wrap its 4-bit epoch several times."""
    from divortio_lz4_b200 import corpus
    rng = np.random.RandomState(77)
    nmsg = 9000
    raw = corpus.jsonmsgs(9, 0, nmsg)
    ln = rng.randint(0, 4097, size=nmsg).astype(np.uint32)
    ln[:8] = [0, 1, 12, 13, 66, 67, 4095, 4096]
    off = np.arange(nmsg, dtype=np.uint64) * 4096
    for plen in (65536, 65535, 4097, 1000, 37):
        dic = corpus.jsonmsgs(45, 0, 17).tobytes()[3:3 + plen]
        dic = np.frombuffer(dic, dtype=np.uint8).copy()
        work = np.concatenate([dic, np.zeros(8, dtype=np.uint8)])
        primed = oracle.new_table()
        oracle.compress_block(dic, 0, dic.size, primed)
        for mode, (warm, init, otab) in {"jenkins": (dl.WARM_JENKINS, None, oracle.warm_table_jenkins(work, dic.size)),
                                         "primed": (dl.WARM_TABLE, primed, primed)}.items():
            dst, doff, clen = dl.compress_blocks(raw, off, ln, prefix=dic, warm=warm, init_table=init)
            odst, odoff, oclen = oracle.compress_blocks_prefix(dic, otab, raw, off, ln)
            assert np.array_equal(clen, oclen), (plen, mode)
            bad = [i for i in range(nmsg) if not np.array_equal(dst[int(doff[i]):int(doff[i]) + int(clen[i])],
                                                                 odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])])]
            assert not bad, (plen, mode, bad[:5])


def test_overlay_caller_table_with_entries_beyond_16_bits(dl):
    """A caller-supplied initial table may hold anything (Int32): positions beyond the prefix, -1, huge values must give the
    oracle's bytes all the same (k_compress_overlay reads the table as it is)."""
    from divortio_lz4_b200 import corpus
    rng = np.random.RandomState(5)
    nmsg = 3000
    raw = corpus.jsonmsgs(11, 0, nmsg)
    off = np.arange(nmsg, dtype=np.uint64) * 4096
    ln = np.full(nmsg, 4096, dtype=np.uint32)
    dic = corpus.json_dictionary(44)
    primed = oracle.new_table()
    oracle.compress_block(dic, 0, dic.size, primed)
    weird = primed.copy()
    idx = rng.choice(weird.size, size=4000, replace=False)
    weird[idx[:1000]] = -1
    weird[idx[1000:2000]] = rng.randint(65536, 70000, size=1000)          # inside the first message's virtual range
    weird[idx[2000:3000]] = rng.randint(70000, 1 << 30, size=1000)
    weird[idx[3000:]] = 65536                                            # position 65535 + 1: the last that fits is 65535
    for tab in (weird, primed):
        dst, doff, clen = dl.compress_blocks(raw, off, ln, prefix=dic, warm=dl.WARM_TABLE, init_table=tab)
        odst, odoff, oclen = oracle.compress_blocks_prefix(dic, tab, raw, off, ln)
        assert np.array_equal(clen, oclen)
        bad = [i for i in range(nmsg) if not np.array_equal(dst[int(doff[i]):int(doff[i]) + int(clen[i])],
                                                             odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])])]
        assert not bad, bad[:5]


def test_decompress_edge_corpora(dl):
    corp = edge_corpora()
    names = list(corp)
    blocks = [oracle.compress_block_bytes(corp[n]) for n in names]
    buf, off, ln = _pack(blocks)
    cap = np.array([len(corp[n]) for n in names], dtype=np.uint32)
    doff = np.zeros(len(names), dtype=np.uint64)
    doff[1:] = np.cumsum(cap.astype(np.uint64) + 5)[:-1]
    out, olen, status = dl.decompress_blocks(buf, off, ln, doff, cap)
    assert not status.any()
    for i, n in enumerate(names):
        assert int(olen[i]) == len(corp[n]), n
        assert out[int(doff[i]):int(doff[i]) + int(olen[i])].tobytes() == corp[n], n


def test_decompress_liblz4_blocks(dl):
    import lz4f
    if not lz4f.available():
        pytest.skip("liblz4 not present")
    corp = edge_corpora()
    names = [n for n in corp if len(corp[n]) > 0]
    blocks = []
    for n in names:
        d = np.frombuffer(corp[n], dtype=np.uint8)
        o = np.zeros(lz4f.L.LZ4_compressBound(d.size), dtype=np.uint8)
        k = lz4f.L.LZ4_compress_default(d.ctypes.data, o.ctypes.data, d.size, o.size)
        blocks.append(o[:k].tobytes())
    buf, off, ln = _pack(blocks)
    cap = np.array([len(corp[n]) for n in names], dtype=np.uint32)
    doff = np.zeros(len(names), dtype=np.uint64)
    doff[1:] = np.cumsum(cap.astype(np.uint64))[:-1]
    out, olen, status = dl.decompress_blocks(buf, off, ln, doff, cap)
    assert not status.any()
    for i, n in enumerate(names):
        assert out[int(doff[i]):int(doff[i]) + int(olen[i])].tobytes() == corp[n], n


def test_decompress_errors_match_oracle(dl):
    cases = [
        (b"\xF0\x20" + b"x" * 47, 16, None, 1),          # Output Buffer Too Small
        (b"\x50abc", 16, None, 2),                        # Malformed Input
        (b"\x10a\x00\x00\x00", 16, None, 3),              # Invalid Offset 0
        (b"\x10a\x05\x00\x00", 16, None, 4),              # Dictionary Offset Out of Bounds
        (b"\x1Fa\x01\x00\xFF\xFF", 5000, None, 2),        # length run hits the end of input
        (b"\x1Fa\x01\x00\xFF\x10\x00", 64, None, 1),      # match longer than the output
    ]
    for blk, cap, dic, want in cases:
        out = np.zeros(cap, dtype=np.uint8)
        with pytest.raises(oracle.OracleError) as e1:
            oracle.decompress_block(blk, 0, len(blk), out, 0, dic)
        assert -e1.value.code == want
        with pytest.raises(dl.LZ4Error) as e2:
            dl.decompressBlock(blk, 0, len(blk), np.zeros(cap, dtype=np.uint8), 0, dic)
        assert e2.value.status == want and str(e2.value) == oracle.MESSAGES[-want]
    # batch: one bad block must not disturb its neighbours
    good = oracle.compress_block_bytes(b"neighbour " * 30)
    buf, off, ln = _pack([good, b"\x10a\x00\x00\x00", good])
    cap = np.array([300, 16, 300], dtype=np.uint32)
    doff = np.array([0, 300, 316], dtype=np.uint64)
    out, olen, status = dl.decompress_blocks(buf, off, ln, doff, cap, check=False)
    assert status.tolist() == [0, 3, 0]
    assert out[:300].tobytes() == b"neighbour " * 30 == out[316:616].tobytes()


def test_dictionary_reference_resolves(dl):
    out = np.zeros(16, dtype=np.uint8)
    n = dl.decompressBlock(b"\x10a\x05\x00\x00", 0, 5, out, 0, b"WXYZ")
    assert n == 5 and out[:5].tobytes() == b"aWXYZ"


def test_single_block_raw_api_roundtrips_table_state(dl):
    """LZ4.compressRaw semantics incl. the in/out hash table (linked use across calls)."""
    from divortio_lz4_b200 import corpus
    data = corpus.log(7, 150000)
    t_gpu, t_cpu = np.zeros(16384, dtype=np.int32), oracle.new_table()
    for start, n in ((0, 65536), (65536, 65536), (131072, 150000 - 131072)):
        out = np.zeros(dl.compress_bound(n) + 10, dtype=np.uint8)
        k = dl.compressBlock(data, out, start, n, t_gpu, 10)
        k_cpu, o_cpu = oracle.compress_block(data, start, n, t_cpu)
        assert k == k_cpu and out[10:10 + k].tobytes() == o_cpu[:k_cpu].tobytes()
        assert np.array_equal(t_gpu, t_cpu)
    # -1 filled table == empty table (tests/raw/raw.test.mjs:14); undersized output truncates silently
    t = np.full(16384, -1, dtype=np.int32)
    small = np.zeros(20, dtype=np.uint8)
    msg = b"hello hello hello hello hello hello hello, and a tail that does not repeat: 0123456789"
    k = dl.compressBlock(msg, small, 0, len(msg), t, 0)
    want = oracle.compress_block_bytes(msg)
    assert k == len(want) > 20 and small.tobytes() == want[:20]
    # decompressRaw into the middle of a larger array, history = earlier bytes of that array
    blk = oracle.compress_block_bytes(data[:70000].tobytes(), 65536, 70000 - 65536, oracle.new_table())
    arr = np.zeros(70000, dtype=np.uint8)
    arr[:65536] = data[:65536]
    n = dl.decompressBlock(blk, 0, len(blk), arr, 65536)
    assert n == 70000 - 65536 and np.array_equal(arr, data[:70000])


def test_xxh32_single_and_batch(dl):
    corp = edge_corpora()
    for name, d in corp.items():
        assert dl.xxHash32(d) == oracle.xxh32(d), name
        assert dl.xxHash32(d, 0x9E3779B1) == oracle.xxh32(d, 0x9E3779B1), name
    names = list(corp)
    buf, off, ln = _pack([corp[n] for n in names])
    got = dl.xxh32_batch(buf, off, ln)
    assert got.tolist() == [oracle.xxh32(corp[n]) for n in names]
    assert dl.xxHash32(b"") == 0x02CC5D05 and dl.xxHash32("Hello World") == 0xB1FD16EE   # reference KATs


def test_xxh32_batch_every_alignment_and_stripe_boundary(dl):
    """k_xxh32_batch hashes items at any byte alignment (the packed payloads of a frame body) eight stripes per step; lengths
    around the 128-byte steps and the 16-byte stripes, every alignment 0..15, two seeds."""
    rng = np.random.RandomState(123)
    lens = [0, 1, 3, 4, 15, 16, 17, 31, 32, 127, 128, 129, 143, 144, 145, 159, 160, 255, 256, 257, 271, 272, 273, 1000, 4097, 65541]
    items, off, pos = [], [], 0
    parts = []
    for a in range(16):
        for n in lens:
            pad = (a - pos) % 16
            parts.append(bytes(pad))
            pos += pad
            d = rng.randint(0, 256, size=n, dtype=np.uint8).tobytes()
            off.append(pos)
            items.append(d)
            parts.append(d)
            pos += n
    buf = np.frombuffer(b"".join(parts) + bytes(8), dtype=np.uint8)
    off = np.array(off, dtype=np.uint64)
    ln = np.array([len(d) for d in items], dtype=np.uint32)
    assert sorted(set(int(o) % 16 for o in off)) == list(range(16))
    for seed in (0, 0x9E3779B1):
        got = dl.xxh32_batch(buf, off, ln, seed) if seed else dl.xxh32_batch(buf, off, ln)
        assert got.tolist() == [oracle.xxh32(d, seed) for d in items], seed


def test_packed_pipelined_host_path(dl):
    """dst_off == NULL: packed output through the chunked H2D/compute/D2H pipeline; bytes identical to the strided call."""
    from divortio_lz4_b200 import corpus
    n = 300 * 1024 * 1024 + 777                       # > 2 chunks of 128 MiB
    data = corpus.mixed(31, n)
    off = np.arange(0, n, 65536, dtype=np.uint64)
    ln = np.minimum(65536, n - off).astype(np.uint32)
    dst, doff, clen = dl.compress_blocks(data, off, ln, packed=True)
    odst, odoff, oclen = oracle.compress_blocks(data, off, ln)
    assert np.array_equal(clen, oclen) and dst.size == int(clen.sum())
    assert np.array_equal(oracle.xxh32_batch(dst, doff, clen), oracle.xxh32_batch(odst, odoff, oclen))
    out, olen, status = dl.decompress_blocks(dst, None, clen, off, ln)
    assert not status.any() and np.array_equal(olen, ln) and np.array_equal(out[:n], data)


def _fuzz_block(rng, n):
    """A block built from pieces that exercise different paths of the match finder: copies of earlier pieces at random distances
    (short and long matches, same-slot pairs, stale table entries), runs, periodic patterns, text-like words, noise."""
    words = [b"status=", b"GET ", b"/api/v1/", b"user", b"\n2026-10-", b"INFO ", b"error", b"=200 ", b"abcabc", b"    "]
    out = bytearray()
    while len(out) < n:
        k = rng.randint(0, 8)
        if k == 0 and len(out) > 8:                                  # copy of an earlier stretch
            src = rng.randint(0, len(out) - 4)
            ln = int(min(rng.choice([4, 5, 7, 12, 31, 32, 33, 63, 64, 65, 130, 300, 2000]), len(out) - src))
            out += out[src:src + ln]
        elif k == 1:                                                 # run
            out += bytes([rng.randint(0, 256)]) * int(rng.choice([1, 3, 4, 5, 17, 64, 255, 256, 1000, 5000]))
        elif k == 2:                                                 # periodic
            p = rng.randint(1, 40)
            unit = rng.bytes(p)
            out += unit * int(rng.randint(1, 60))
        elif k == 3:                                                 # noise
            out += rng.bytes(int(rng.choice([1, 2, 3, 8, 40, 200, 3000])))
        elif k == 4:                                                 # few-symbol noise (many hash collisions)
            out += bytes(rng.randint(0, 3, size=int(rng.randint(4, 400))).astype(np.uint8))
        else:                                                        # words
            for _ in range(rng.randint(1, 12)):
                out += words[rng.randint(0, len(words))]
                if rng.randint(0, 3) == 0:
                    out += str(rng.randint(0, 100000)).encode()
    return bytes(out[:n])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_compress_fuzz_random_structured_blocks_bit_exact(dl, seed):
    """Differential fuzz of the producer/walker match finder against the oracle: 1500 random structured blocks of every length
    class (empty, below the 13-byte minimum, around the window sizes 32 / 64 / 112, up to 64 KiB) in one batch, compressed bytes
    compared one by one, then the batch decoded back."""
    rng = np.random.RandomState(1000 + seed)
    lens = [0, 1, 4, 5, 11, 12, 13, 14, 31, 32, 33, 63, 64, 65, 95, 96, 97, 111, 112, 113, 127, 128, 129, 143, 144, 145, 4095, 4096, 4097,
            65535, 65536]
    lens += [int(x) for x in rng.randint(0, 2000, size=600)] + [int(x) for x in rng.randint(2000, 65537, size=869)]
    items = [_fuzz_block(rng, n) for n in lens]
    buf, off, ln = _pack(items)
    dst, doff, clen = dl.compress_blocks(buf, off, ln)
    odst, odoff, oclen = oracle.compress_blocks(buf, off, ln)
    bad = [i for i in range(len(items)) if clen[i] != oclen[i] or not np.array_equal(dst[int(doff[i]):int(doff[i]) + int(clen[i])],
                                                                                       odst[int(odoff[i]):int(odoff[i]) + int(oclen[i])])]
    assert not bad, ("blocks differ from the oracle", [(i, len(items[i])) for i in bad[:10]])
    out, olen, status = dl.decompress_blocks(dst, doff, clen, off, ln)
    assert not status.any() and np.array_equal(olen, ln)
    for i in range(0, len(items), 37):
        assert out[int(off[i]):int(off[i]) + int(ln[i])].tobytes() == items[i], i
