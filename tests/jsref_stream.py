"""CPU restatement of the reference's streaming codec, block by block exactly as written:
LZ4Encoder (src/shared/lz4Encode.js:95-332) and LZ4Decoder (src/shared/lz4Decode.js:52-306; the decompressBlock call at
:232 is stale in the reference -- wrong argument order -- and is restated with the intended meaning: decode the block with the
window as dictionary).  Test infrastructure: block arithmetic comes from the oracle (oracle/lz4_oracle.c)."""
import numpy as np

import oracle

MAX_WINDOW_SIZE = 65536
BLOCK_MAX_SIZES = {4: 65536, 5: 262144, 6: 1048576, 7: 4194304}


def get_block_id(n):
    if not n or n <= 65536:
        return 4
    if n <= 262144:
        return 5
    if n <= 1048576:
        return 6
    return 7


def u32(v):
    return int(v & 0xFFFFFFFF).to_bytes(4, "little")


def create_frame_header(block_independence, content_checksum, bd_id, dict_id):
    flg = 1 << 6
    if block_independence:
        flg |= 0x20
    if content_checksum:
        flg |= 0x04
    if dict_id:
        flg |= 0x01
    body = bytes([flg, (bd_id & 7) << 4]) + (u32(dict_id) if dict_id else b"")
    return u32(0x184D2204) + body + bytes([(oracle.xxh32(body) >> 8) & 0xFF])


class RefEncoder(object):
    def __init__(self, max_block_size=4194304, block_independence=False, content_checksum=False, dictionary=None):
        self.block_independence = block_independence
        self.content_checksum = content_checksum
        self.block_size = BLOCK_MAX_SIZES.get(get_block_id(max_block_size), 4194304)
        self.bd_id = get_block_id(self.block_size)
        self.buffer = np.zeros(0, dtype=np.uint8)
        self.has_written_header = False
        self.is_closed = False
        self.hash_table = oracle.new_table()
        self.dict_size = 0
        self.hashed = bytearray()                      # stands in for XXHash32.update(): digest() == xxh32 of everything fed
        self.dict_id = None
        if dictionary:
            d = np.frombuffer(bytes(dictionary), dtype=np.uint8)
            self.dict_id = oracle.xxh32(d)
            window = d[max(0, d.size - MAX_WINDOW_SIZE):]
            self.buffer = window.copy()
            self.dict_size = int(window.size)
            # :147-167 -- identical loop to bufferCompress.js:186-204
            work = np.concatenate([self.buffer, np.zeros(8, dtype=np.uint8)])
            self.hash_table = oracle.warm_table_jenkins(work, self.buffer.size)

    def add(self, chunk):
        assert not self.is_closed
        data = np.frombuffer(bytes(chunk), dtype=np.uint8)
        if data.size == 0:
            return []
        if self.content_checksum:
            self.hashed += bytes(chunk)
        self.buffer = np.concatenate([self.buffer, data])
        results = []
        if not self.has_written_header:
            results.append(create_frame_header(self.block_independence, self.content_checksum, self.bd_id, self.dict_id))
            self.has_written_header = True
        while self.buffer.size >= self.dict_size + self.block_size:
            results.append(self._flush_block(False))
        return results

    def _flush_block(self, final):
        available = self.buffer.size - self.dict_size
        if available == 0 and not final:
            return b""
        block_size = self.block_size
        if available < block_size:
            if final:
                block_size = available
            else:
                return b""
        src_start = self.dict_size
        max_output = block_size + 1024
        output = np.zeros(max_output + 4, dtype=np.uint8)
        if self.block_independence:
            self.hash_table[:] = 0
        comp_size, _ = oracle.compress_block(self.buffer, src_start, block_size, self.hash_table, output, 4)
        if 0 < comp_size < block_size:
            res = u32(comp_size) + output[4:4 + comp_size].tobytes()
        else:
            res = u32(block_size | 0x80000000) + self.buffer[src_start:src_start + block_size].tobytes()
        if not self.block_independence:
            consumed_end = src_start + block_size
            preserve = min(consumed_end, MAX_WINDOW_SIZE)
            start = consumed_end - preserve
            self.buffer = self.buffer[start:].copy()
            self.dict_size = preserve
            t = self.hash_table
            t[:] = np.where(t > start, t - start, 0)
        else:
            self.buffer = self.buffer[self.dict_size + block_size:].copy()
            self.dict_size = 0
        return res

    def finish(self):
        if self.is_closed:
            return []
        self.is_closed = True
        frames = []
        if not self.has_written_header:
            frames.append(create_frame_header(self.block_independence, self.content_checksum, self.bd_id, self.dict_id))
        while self.buffer.size - self.dict_size > 0:
            frames.append(self._flush_block(True))
        frames.append(u32(0))
        if self.content_checksum:
            frames.append(u32(oracle.xxh32(bytes(self.hashed))))
        return frames


class RefDecoder(object):
    def __init__(self, dictionary=None, verify_checksum=True):
        self.dictionary = bytes(dictionary) if dictionary else None
        self.verify = verify_checksum
        self.state = 0
        self.buffer = b""
        self.window = bytearray()
        if self.dictionary:
            self.window = bytearray(self.dictionary[-MAX_WINDOW_SIZE:])
        self.hashed = None
        self.bd = 0x70

    def update(self, chunk):
        self.buffer += bytes(chunk)
        out = []
        while True:
            if self.state == 0:
                if len(self.buffer) < 4:
                    break
                if int.from_bytes(self.buffer[:4], "little") != 0x184D2204:
                    raise ValueError("LZ4: Invalid Magic Number")
                self.buffer = self.buffer[4:]
                self.state = 1
                self.hashed = bytearray() if self.verify else None
            if self.state == 1:
                if len(self.buffer) < 2:
                    break
                flg = self.buffer[0]
                self.indep = bool(flg & 0x20)
                self.has_bc = bool(flg & 0x10)
                self.has_cs = bool(flg & 0x08)
                self.has_cc = bool(flg & 0x04)
                self.has_did = bool(flg & 0x01)
                need = 2 + (8 if self.has_cs else 0) + (4 if self.has_did else 0) + 1
                if len(self.buffer) < need:
                    break
                self.bd = self.buffer[1]
                self.buffer = self.buffer[need:]
                self.state = 2
            if self.state == 2:
                if len(self.buffer) < 4:
                    break
                val = int.from_bytes(self.buffer[:4], "little")
                self.buffer = self.buffer[4:]
                if val == 0:
                    self.state = 4
                    continue
                self.uncompressed = bool(val & 0x80000000)
                self.cur = val & 0x7FFFFFFF
                self.state = 3
            if self.state == 3:
                need = self.cur + (4 if self.has_bc else 0)
                if len(self.buffer) < need:
                    break
                block = self.buffer[:self.cur]
                self.buffer = self.buffer[need:]
                if self.uncompressed:
                    dec = bytes(block)
                else:
                    dic = None if self.indep else (bytes(self.window) if self.window else None)
                    ws = np.zeros(4194304 + 16, dtype=np.uint8)
                    n = oracle.decompress_block(np.frombuffer(block, dtype=np.uint8), 0, len(block), ws, 0, dic)
                    dec = ws[:n].tobytes()
                out.append(dec)
                if self.hashed is not None:
                    self.hashed += dec
                if not self.indep:
                    self.window = (self.window + dec)[-MAX_WINDOW_SIZE:]
                self.state = 2
            if self.state == 4:
                if self.has_cc:
                    if len(self.buffer) < 4:
                        break
                    if self.verify and self.hashed is not None:
                        if int.from_bytes(self.buffer[:4], "little") != oracle.xxh32(bytes(self.hashed)):
                            raise ValueError("LZ4: Content Checksum Error")
                    self.buffer = self.buffer[4:]
                self.state = 0
                self.hashed = None
                if len(self.buffer) == 0:
                    break
        return out
