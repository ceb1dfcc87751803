"""tests/jsref.py -- second, independent statement of the reference semantics, in pure Python.

Written from the JavaScript text with JS typed-array semantics modelled explicitly (int32 wrap,
out-of-range loads read as 0 / `undefined`, out-of-range stores dropped).  Slow: small cases only.
Its job is to cross-check oracle/lz4_oracle.c (tests/test_oracle_vs_jsref.py) and to generate the
golden fixtures under tests/golden/ (tests/golden/make_golden.py).

  js_compress_block     src/block/blockCompress.js:31-233
  js_decompress_block   src/block/blockDecompress.js:30-275   (literal, incl. the :232-250 tail)
  js_compress_buffer    src/buffer/bufferCompress.js:100-259
  js_decompress_buffer  src/buffer/bufferDecompress.js:51-220
  js_xxh32              src/xxhash32/xxhash32.js:21-97
"""

M32 = 0xFFFFFFFF


def i32(x):
    x &= M32
    return x - (1 << 32) if x & 0x80000000 else x


def imul(a, b):
    return i32((a & M32) * (b & M32))


class U8(object):
    """Uint8Array model: OOB load -> 0 (undefined in integer context), OOB store dropped."""

    def __init__(self, data):
        self.b = bytearray(data)

    def __len__(self):
        return len(self.b)

    def g(self, i):
        return self.b[i] if 0 <= i < len(self.b) else 0

    def s(self, i, v):
        if 0 <= i < len(self.b):
            self.b[i] = v & 0xFF

    def le32(self, i):
        return i32(self.g(i) | (self.g(i + 1) << 8) | (self.g(i + 2) << 16) | (self.g(i + 3) << 24))


def rotl(x, r):
    x &= M32
    return ((x << r) | (x >> (32 - r))) & M32


def js_xxh32(data, seed=0):
    P1, P2, P3, P4, P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393
    a = U8(data)
    n = len(a)
    p = 0
    if n >= 16:
        v = [(seed + P1 + P2) & M32, (seed + P2) & M32, seed & M32, (seed - P1) & M32]
        while p <= n - 16:
            for k in range(4):
                v[k] = (rotl(v[k] + (a.le32(p + 4 * k) & M32) * P2, 13) * P1) & M32
            p += 16
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M32
    else:
        h = (seed + P5) & M32
    h = (h + n) & M32
    while p <= n - 4:
        h = (rotl(h + (a.le32(p) & M32) * P3, 17) * P4) & M32
        p += 4
    while p < n:
        h = (rotl(h + a.g(p) * P5, 11) * P1) & M32
        p += 1
    h ^= h >> 15
    h = (h * P2) & M32
    h ^= h >> 13
    h = (h * P3) & M32
    h ^= h >> 16
    return h


def _emit_len(out, token_pos, d, n):
    if n >= 15:
        out.s(token_pos, 0xF0)
        rest = n - 15
        while rest >= 255:
            out.s(d, 255)
            d += 1
            rest -= 255
        out.s(d, rest)
        d += 1
    else:
        out.s(token_pos, n << 4)
    return d


def js_compress_block(src, out, src_start, src_len, table, out_off):
    """src: U8, out: U8, table: list of 16384 ints (mutated).  Returns bytes written."""
    s = src_start
    s_end = src_start + src_len
    mflimit = s_end - 12
    match_limit = s_end - 5
    d = out_off
    anchor = s
    smc = 67
    while s < mflimit:
        seq = src.le32(s)
        h = ((imul(seq, i32(2654435761)) & M32) >> 18) & 16383
        m = i32(table[h] - 1)
        table[h] = s + 1
        if m < 0 or s == m or (((s - m) & M32) >> 16) > 0 or src.le32(m) != seq:
            s += smc >> 6
            smc += 1
            continue
        smc = 67
        lit = s - anchor
        tok = d
        d = _emit_len(out, tok, d + 1, lit)
        for k in range(lit):
            out.s(d + k, src.g(anchor + k))
        d += lit
        sp, mp = s + 4, m + 4
        while sp < match_limit and src.g(sp) == src.g(mp):
            sp += 1
            mp += 1
        off = s - m
        out.s(d, off & 0xFF)
        out.s(d + 1, (off >> 8) & 0xFF)
        d += 2
        code = sp - s - 4
        if code >= 15:
            out.s(tok, out.g(tok) | 0x0F)
            rest = code - 15
            while rest >= 255:
                out.s(d, 255)
                d += 1
                rest -= 255
            out.s(d, rest)
            d += 1
        else:
            out.s(tok, out.g(tok) | code)
        s = anchor = sp
    lit = s_end - anchor
    tok = d
    d = _emit_len(out, tok, d + 1, lit)
    for k in range(lit):
        out.s(d + k, src.g(anchor + k))
    d += lit
    return d - out_off


class JsError(Exception):
    pass


def js_decompress_block(inp, in_off, in_size, out, out_off, dictionary=None):
    """Literal blockDecompress.js, including the unguarded double-copy tail (:232-250)."""
    ip, in_end, op = in_off, in_off + in_size, out_off
    out_len = len(out)
    dlen = len(dictionary) if dictionary is not None else 0
    while ip < in_end:
        token = inp.g(ip)
        ip += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = inp.g(ip)
                ip += 1
                lit += b
                if b != 255:
                    break
        end_lit = op + lit
        if end_lit > out_len:
            raise JsError("LZ4: Output Buffer Too Small")
        if ip + lit > in_end:
            raise JsError("LZ4: Malformed Input")
        for k in range(lit):
            out.s(op + k, inp.g(ip + k))
        op, ip = end_lit, ip + lit
        if ip >= in_end:
            break
        offset = inp.g(ip) | (inp.g(ip + 1) << 8)
        ip += 2
        if offset == 0:
            raise JsError("LZ4: Invalid Offset 0")
        ml = token & 15
        if ml == 15:
            while True:
                b = inp.g(ip)
                ip += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        cs = op - offset
        if cs < 0:
            from_dict = min(-cs, ml)
            cs = dlen + cs
            if cs < 0 or cs + from_dict > dlen:
                raise JsError("LZ4: Dictionary Offset Out of Bounds")
            for k in range(from_dict):
                out.s(op + k, dictionary.g(cs + k))
            op += from_dict
            rem = ml - (op - end_lit)
            rp = op - offset
            for k in range(max(rem, 0)):
                out.s(op + k, out.g(rp + k))
            op += max(rem, 0)
        elif offset == 1:
            v = out.g(cs)
            for k in range(ml):
                out.s(op + k, v)
            op += ml
        elif offset >= ml and ml > 16:
            chunk = [out.g(cs + k) for k in range(ml)]
            for k in range(ml):
                out.s(op + k, chunk[k])
            op += ml
        else:
            end_match, rp = op + ml, cs
            if offset >= 8:
                while op < end_match - 8:
                    for _ in range(8):
                        out.s(op, out.g(rp))
                        op += 1
                        rp += 1
                if op < end_match:
                    tail_out = end_match - 8
                    tail_src = rp + (end_match - op) - 8
                    for k in range(8):
                        out.s(tail_out + k, out.g(tail_src + k))
                    op = end_match
            else:
                while op < end_match:
                    out.s(op, out.g(rp))
                    op += 1
                    rp += 1
    return op - out_off


def _jenkins_slot(seq):
    h = seq
    h = i32(h + 2127912214 + (h << 12))
    h = i32(h ^ -949894596 ^ ((h & M32) >> 19))
    h = i32(h + 374761393 + (h << 5))
    h = i32((h + -744332180) ^ (h << 9))
    h = i32(h + -42973499 + (h << 3))
    h = i32(h ^ -1252372727 ^ ((h & M32) >> 16))
    return ((h & M32) >> 18) & 16383


def js_compress_buffer(data, dictionary=None, max_block=4194304, indep=False, content_cksum=False,
                       add_size=True, out_cap=None):
    raw = bytes(data)
    n = len(raw)
    work, start, dict_id = raw, 0, None
    if dictionary:
        dict_id = js_xxh32(dictionary, 0)
        win = bytes(dictionary)[-65536:]
        work, start = win + raw, len(win)
    if not max_block or max_block <= 65536:
        bd = 4
    elif max_block <= 262144:
        bd = 5
    elif max_block <= 1048576:
        bd = 6
    else:
        bd = 7
    bsz = {4: 65536, 5: 262144, 6: 1048576, 7: 4194304}[bd]
    out = U8(bytes(out_cap if out_cap is not None else 19 + n + n // 255 + 64 + 8))
    for k, v in enumerate((0x04, 0x22, 0x4D, 0x18)):
        out.s(k, v)
    flg = (1 << 6) | (0x20 if indep else 0) | (0x04 if content_cksum else 0) | (0x01 if dict_id is not None else 0) | (0x08 if add_size else 0)
    out.s(4, flg)
    out.s(5, (bd & 7) << 4)
    op = 6

    def w32(v, at):
        for k in range(4):
            out.s(at + k, (v >> (8 * k)) & 0xFF)

    if add_size:
        w32(n & M32, op)
        w32(0, op + 4)
        op += 8
    if dict_id is not None:
        w32(dict_id, op)
        op += 4
    out.s(op, (js_xxh32(bytes(out.b[4:op]), 0) >> 8) & 0xFF)
    op += 1
    table = [0] * 16384
    w = U8(work)
    if start > 0:
        for i in range(0, start - 4 + 1):
            table[_jenkins_slot(w.le32(i))] = i + 1
    pos, total_end = start, start + n
    while pos < total_end:
        end = min(pos + bsz, total_end)
        size_pos = op
        op += 4
        c = js_compress_block(w, out, pos, end - pos, table, op)
        if 0 < c < end - pos:
            w32(c, size_pos)
            op += c
        else:
            w32((end - pos) | 0x80000000, size_pos)
            for k in range(end - pos):
                out.s(op + k, w.g(pos + k))
            op += end - pos
        if indep:
            table = [0] * 16384
        pos = end
    w32(0, op)
    op += 4
    if content_cksum:
        w32(js_xxh32(raw, 0), op)
        op += 4
    return bytes(out.b[:op])


def js_decompress_buffer(frame, dictionary=None, verify=True, block_decoder=js_decompress_block):
    d = U8(frame)
    n = len(d)
    if n < 4 or (d.le32(0) & M32) != 0x184D2204:
        raise JsError("LZ4: Invalid Magic Number")
    flg = d.g(4)
    if (flg & 0xC0) >> 6 != 1:
        raise JsError("LZ4: Unsupported Version %d" % ((flg & 0xC0) >> 6))
    pos = 6
    expected = 0
    if flg & 0x08:
        expected = ((d.le32(pos + 4) & M32) << 32) + (d.le32(pos) & M32)
        pos += 8
    if flg & 0x01:
        pos += 4
    pos += 1
    direct = expected > 0
    dic = U8(dictionary) if dictionary else None
    if direct:
        result = U8(bytes(expected))
        rpos = 0
    else:
        chunks = []
        window = bytearray((bytes(dictionary)[-65536:] if dictionary else b""))
    while pos < n:
        bs = d.le32(pos) & M32
        pos += 4
        if bs == 0:
            break
        stored, actual = bool(bs & 0x80000000), bs & 0x7FFFFFFF
        if direct:
            if stored:
                if rpos + actual > len(result):
                    raise JsError("RangeError: offset is out of bounds")
                result.b[rpos:rpos + actual] = d.b[pos:pos + actual]
                rpos += actual
            else:
                rpos += block_decoder(d, pos, actual, result, rpos, dic)
        else:
            if stored:
                chunk = bytes(d.b[pos:pos + actual])
            else:
                ws = U8(bytes(4194304 if actual > 2048 else 1 << 20))
                got = block_decoder(d, pos, actual, ws, 0, U8(window) if window else None)
                chunk = bytes(ws.b[:got])
            chunks.append(chunk)
            window = (window + chunk)[-65536:]
        pos += actual
        if flg & 0x10:
            pos += 4
    res = bytes(result.b) if direct else b"".join(chunks)
    if (flg & 0x04) and verify:
        if (d.le32(pos) & M32) != js_xxh32(res, 0):
            raise JsError("LZ4: Content Checksum Error")
    return res
