"""Block-range sharding of the frame path over the GPUs of one box.

Independent blocks shard with no data-path collective: rank r owns the contiguous block range
dlz4_shard_range(nblocks, world, r), produces the frame *segment* of that range on its GPU
([u32 size|stored][data][u32 xxh32]* -- exactly the bytes the single-GPU path writes for those blocks) and the host
concatenates header ++ segments ++ EndMark [++ content checksum].  Only segment LENGTHS cross ranks (a host-side
exclusive scan); torch.distributed is used for that bookkeeping and to gather segments to rank 0 when one process
wants the whole frame.

Linked blocks and the whole-stream content checksum are serial chains: "replicas only" -- rank 0 runs them (the
checksum on its side stream), they are never split.
"""
import numpy as np

from . import api

MAGIC = b"\x04\x22\x4D\x18"
BLOCK_SIZES = {4: 65536, 5: 262144, 6: 1048576, 7: 4194304}


def block_id_for(max_block_size):
    """getBlockId (src/buffer/bufferCompress.js:77-82)."""
    if not max_block_size or max_block_size <= 65536:
        return 4
    if max_block_size <= 262144:
        return 5
    if max_block_size <= 1048576:
        return 6
    return 7


def plan(total_len, max_block_size, world):
    """Per-rank (first_block, block_count, byte_start, byte_end)."""
    bs = BLOCK_SIZES[block_id_for(max_block_size)]
    nblocks = (total_len + bs - 1) // bs
    out = []
    for r in range(world):
        first, count = api.shard_range(nblocks, world, r)
        out.append((first, count, min(first * bs, total_len), min((first + count) * bs, total_len)))
    return bs, nblocks, out


def frame_header(total_len, max_block_size, block_independence, content_checksum, add_content_size, block_checksum, header_hash):
    """Header bytes (bufferCompress.js:147-178).  header_hash(bytes)->u32 supplies xxh32 (the GPU one in production)."""
    flg = (1 << 6) | (0x20 if block_independence else 0) | (0x04 if content_checksum else 0) | (0x08 if add_content_size else 0) \
        | (0x10 if block_checksum else 0)
    desc = bytes([flg, (block_id_for(max_block_size) & 7) << 4])
    if add_content_size:
        desc += int(total_len & 0xFFFFFFFF).to_bytes(4, "little") + (0).to_bytes(4, "little")
    return MAGIC + desc + bytes([(header_hash(desc) >> 8) & 0xFF])


def segment_gpu(data_slice, max_block_size, block_checksum, ctx=None):
    """Frame segment of an independent-block range on this rank's GPU: compress as a frame and strip header/EndMark."""
    frame = api.compressBuffer(data_slice, None, max_block_size, True, False, False, None, block_checksum, ctx=ctx)
    return frame[7:-4]          # header without content size / dict id is 7 bytes; EndMark is 4


def assemble(header, segments, content_hash=None):
    tail = (0).to_bytes(4, "little") + (b"" if content_hash is None else int(content_hash).to_bytes(4, "little"))
    return header + b"".join(segments) + tail


def compress_sharded(data, max_block_size=4194304, content_checksum=False, add_content_size=True, block_checksum=False,
                     rank=0, world=1, segment_fn=None, xxh32_fn=None, gather=None):
    """Independent-block frame of `data` produced by `world` ranks.  Every rank calls this with the same arguments.
    segment_fn(slice, max_block_size, block_checksum) -> bytes   (default: this rank's GPU)
    xxh32_fn(bytes) -> u32                                       (default: the GPU xxh32)
    gather(obj) -> list of every rank's obj on rank 0 (None elsewhere)   (default: torch.distributed.gather_object)
    Returns the frame on rank 0, None on the other ranks."""
    data = api.ensureBuffer(data)
    segment_fn = segment_fn or segment_gpu
    xxh32_fn = xxh32_fn or api.xxHash32
    bs, nblocks, ranges = plan(data.size, max_block_size, world)
    first, count, lo, hi = ranges[rank]
    seg = bytes(segment_fn(data[lo:hi], bs, block_checksum)) if count else b""
    if gather is None:
        gather = _dist_gather
    segs = gather(seg)
    if rank != 0:
        return None
    header = frame_header(data.size, max_block_size, True, content_checksum, add_content_size, block_checksum, xxh32_fn)
    chash = xxh32_fn(data) if content_checksum else None       # serial chain: rank 0 only
    return assemble(header, segs, chash)


def bind_host_near(device_index):
    """One process per GPU: keep this process's threads -- and therefore the pinned staging buffers it allocates next -- on
    the CPUs of the GPU's own NUMA node (NVML's ideal CPU affinity), so that the ranks' host copies do not cross the socket
    interconnect.  Returns the CPU set applied, or None when NVML / the affinity call is unavailable (nothing changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def _dist_gather(obj):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def segment_offsets(lengths):
    """Exclusive scan of segment byte counts: where each rank's segment lands in the host frame (after the header)."""
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.asarray(lengths, dtype=np.int64))
    return off
