"""Host-side sharding logic on CPU: world_size-2 (and 3) gloo groups; each rank builds its frame segment (the oracle stands
in for the GPU kernels here -- this test checks the partition / concatenation logic, not the kernels) and rank 0's
concatenated frame must equal the single-process frame byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_segment(data_slice, max_block_size, block_checksum):
    import oracle
    f = oracle.compress_buffer(data_slice, None, max_block_size, True, False, False, None, block_checksum)
    return f[7:-4]


def _worker(rank, world, port, n, bs, cc, bc, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from divortio_lz4_b200 import corpus, sharded
        data = corpus.mixed(21, n)
        frame = sharded.compress_sharded(data, bs, cc, True, bc, rank=rank, world=world, segment_fn=_oracle_segment,
                                         xxh32_fn=oracle.xxh32)
        if rank == 0:
            want = oracle.compress_buffer(data, None, bs, True, cc, True, None, bc)
            q.put((frame == want, len(frame), oracle.decompress_buffer(frame) == data.tobytes()))
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,bs,cc,bc", [(2, 1000000, 65536, True, True), (2, 65536 * 3, 65536, False, False),
                                              (3, 700001, 262144, True, False), (2, 1, 65536, True, True)])
def test_sharded_frame_equals_single_process_frame(world, n, bs, cc, bc):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bs, cc, bc, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, length, roundtrip = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and roundtrip and length > 0


def test_plan_covers_every_block_once():
    from divortio_lz4_b200 import sharded
    for total in (0, 1, 65536, 65537, 10 ** 6, 8 * 2 ** 20 + 5):
        for world in (1, 2, 4, 8):
            bs, nblocks, ranges = sharded.plan(total, 65536, world)
            assert ranges[0][2] == 0 and ranges[-1][3] == total
            for a, b in zip(ranges, ranges[1:]):
                assert a[3] == b[2] and a[0] + a[1] == b[0]
            assert sum(r[1] for r in ranges) == nblocks


def test_header_matches_oracle_header():
    import oracle
    from divortio_lz4_b200 import sharded
    for size in (True, False):
        for cc in (False, True):
            for bc in (False, True):
                h = sharded.frame_header(12345, 65536, True, cc, size, bc, oracle.xxh32)
                want = oracle.compress_buffer(bytes(12345), None, 65536, True, cc, size, None, bc)
                assert want.startswith(h)


def test_bind_host_near_never_widens_the_affinity_and_survives_missing_nvml():
    """One rank per GPU pins its host threads to the GPU's NUMA node when NVML can say which CPUs that is; without a GPU or
    NVML it must leave the process alone and say so."""
    import os
    import importlib
    sharded = importlib.import_module("divortio_lz4_b200.sharded")
    before = os.sched_getaffinity(0)
    got = sharded.bind_host_near(0)
    after = os.sched_getaffinity(0)
    if got is None:
        assert after == before
    else:
        assert after == got and got <= before
        os.sched_setaffinity(0, before)
