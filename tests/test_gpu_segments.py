"""GPU parity, segment-parallel compression (k_compress_segments): linked-block chains (bufferCompress.js:182,219,234) and
independent blocks > 64 KiB are cut into speculative segments whose start states are verified against the serial parse.
Whatever the speculation does -- hit, miss and re-run, many rounds -- the frame must equal the oracle's byte for byte.  -m gpu."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dl():
    import divortio_lz4_b200 as m
    m.default_context()
    return m


class seg_env:
    """DLZ4_SEG_KIB / DLZ4_SEG_WARM_KIB ... are read once, by dlz4_init: the default context is replaced by one created under
    the wanted environment and dropped again on exit."""

    def __init__(self, seg_kib=None, warm_kib=None, group=None, unit_kib=None):
        self.new = {"DLZ4_SEG_KIB": seg_kib, "DLZ4_SEG_WARM_KIB": warm_kib, "DLZ4_SEG_GROUP": group, "DLZ4_SEG_UNIT_KIB": unit_kib}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.new}
        for k, v in self.new.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        self._fresh()

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        self._fresh()

    @staticmethod
    def _fresh():
        from divortio_lz4_b200 import api
        api._default.clear()                                    # (contexts other modules still hold stay alive)


def _corpora(n):
    from divortio_lz4_b200 import corpus
    r = bytearray(corpus.rand(5, n).tobytes())
    for k in range(40):                                        # sparse long-distance repeats in random data
        a, b = 1000 + 37 * k, n // 2 + 50000 * k % (n // 2 - 5000)
        r[b:b + 300] = r[a:a + 300]
    return {
        "log": corpus.log(11, n).tobytes(),
        "mixed": corpus.mixed(12, n).tobytes(),
        "zero": bytes(n),
        "rand": corpus.rand(13, n).tobytes(),
        "rand_repeats": bytes(r),
        "bench": corpus.benchjson(n).tobytes(),
        "json": corpus.jsonmsgs(14, 0, n // 4096 + 1).tobytes()[:n],
    }


def _check(dl, data, bs, indep, **kw):
    f = dl.compressBuffer(data, kw.get("dictionary"), bs, indep, False, True, None, False)
    want = oracle.compress_buffer(data, kw.get("dictionary"), bs, indep, False, True, None, False)
    assert len(f) == len(want) and f == want
    return dl.default_context().segment_stats


@pytest.mark.parametrize("seg_kib,warm_kib", [(None, None), (64, 0), (64, 64), (128, 256), (256, 512), (1024, 128)])
def test_linked_chain_every_corpus(dl, seg_kib, warm_kib):
    n = 3 * 1024 * 1024 + 12345
    with seg_env(seg_kib, warm_kib):
        for name, data in _corpora(n).items():
            for bs in (65536, 262144, 4194304):
                segs, reruns, rounds = _check(dl, data, bs, False)
                assert segs >= 2, (name, bs)


@pytest.mark.parametrize("seg_kib,warm_kib", [(None, None), (64, 0), (128, 128), (512, 512)])
def test_large_independent_blocks_every_corpus(dl, seg_kib, warm_kib):
    n = 9 * 1024 * 1024 + 777
    with seg_env(seg_kib, warm_kib):
        for name, data in _corpora(n).items():
            for bs in (262144, 1048576, 4194304):
                _check(dl, data, bs, True)


def test_failed_speculation_is_rerun_not_emitted(dl):
    """No warm-up at all: nearly every speculative start state is wrong on text, so the re-run path produces the frame."""
    from divortio_lz4_b200 import corpus
    data = corpus.log(21, 4 * 1024 * 1024).tobytes()
    with seg_env(64, 0):
        segs, reruns, rounds = _check(dl, data, 4194304, False)
    assert segs == 64 and reruns >= 32 and rounds >= 1
    with seg_env(128, 512):
        segs2, reruns2, _ = _check(dl, data, 4194304, False)
    assert reruns2 <= 2, "512 KiB of warm-up should converge on log text"
    # stretches of two and four 128 KiB verification units without warm-up: the head of every stretch fails, its re-run
    # must hand a state to the next member that either verifies or re-runs that one as well
    for kib in (256, 512):
        with seg_env(kib, 0):
            segs3, reruns3, rounds3 = _check(dl, data, 4194304, False)
        assert segs3 == 32 and reruns3 >= 4 and rounds3 >= 1, (kib, segs3, reruns3, rounds3)


@pytest.mark.parametrize("seg_kib,warm_kib,group,unit_kib", [(64, 128, 3, None), (128, 0, 2, 64), (256, 64, 5, 64), (128, 512, 4, 128)])
def test_segment_groups_and_verification_units(dl, seg_kib, warm_kib, group, unit_kib):
    """Groups of stretches on the shared-memory warps (two queues, kSegCont / kSegMore) and verification units smaller than
    a stretch, forced at sizes where the defaults would not use them; short warm-ups make heads fail, so re-runs chase into
    members that continued from a wrong state.  Linked chains and independent large blocks, every corpus."""
    n = 6 * 1024 * 1024 + 4321
    with seg_env(seg_kib, warm_kib, group, unit_kib):
        for name, data in _corpora(n).items():
            for bs, indep in ((4194304, False), (65536, False), (4194304, True), (1048576, True)):
                segs, reruns, rounds = _check(dl, data, bs, indep)
                assert segs >= 6, (name, bs, indep, segs)


def test_linked_chain_with_dictionary(dl):
    from divortio_lz4_b200 import corpus
    data = corpus.jsonmsgs(4, 0, 600).tobytes()                 # 2.4 MiB
    dic = corpus.json_dictionary(44).tobytes()
    with seg_env(128, 256):
        for bs in (65536, 1048576):
            for indep in (False, True):
                f = dl.compressBuffer(data, dic, bs, indep, True)
                assert f == oracle.compress_buffer(data, dic, bs, indep, True), (bs, indep)
                assert dl.decompressBuffer(f, dic) == data


def test_default_frame_64mib_log_config0(dl):
    """BASELINE configs[0] shape: 64 MiB log text, 4 MiB blocks -- linked (the reference's default) and independent."""
    from divortio_lz4_b200 import corpus
    data = corpus.log(1, 64 * 1024 * 1024)
    for indep in (False, True):
        f = dl.compressBuffer(data, None, 4194304, indep, False, True)
        assert f == oracle.compress_buffer(data, None, 4194304, indep, False, True)
        segs, reruns, rounds = dl.default_context().segment_stats
        assert segs >= 128 and reruns <= segs // 8, (segs, reruns, rounds)
        assert dl.decompressBuffer(f) == data.tobytes()
