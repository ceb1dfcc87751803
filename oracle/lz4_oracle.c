/*
 * oracle/lz4_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C11, no dependencies) of the divortio-lz4 hot path:
 *   compressBlock    src/block/blockCompress.js:31-233
 *   decompressBlock  src/block/blockDecompress.js:30-275
 *   compressBuffer   src/buffer/bufferCompress.js:100-259
 *   decompressBuffer src/buffer/bufferDecompress.js:51-220
 *   xxHash32         src/xxhash32/xxhash32.js:21-97
 *
 * It is the parity checker for the CUDA path and the timed CPU baseline.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product library never links or calls it.
 *
 * Pinning: the reference is pure JavaScript and no JS engine exists in this image,
 * so the reference cannot be executed here.  This file is pinned against
 *   - every golden vector / KAT the reference's tests hold for this path
 *     (tests/golden.test.mjs:23,39,52; tests/xxhash32/xxhash32.test.mjs:13,20),
 *   - the scratch KATs K1..K10/D1 of SURVEY.md A.3 (derived from the JS text),
 *   - an independent literal Python transliteration (tests/jsref.py),
 *   - liblz4 1.9.4 / libxxhash as third-party decoders of every frame it emits.
 * See tests/test_oracle_*.py.
 *
 * Error codes mirror the JS exception strings one-to-one (ORC_E_*).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define ORC_OK 0
#define ORC_E_OUTPUT_TOO_SMALL  (-1) /* "LZ4: Output Buffer Too Small"          blockDecompress.js:74  */
#define ORC_E_MALFORMED         (-2) /* "LZ4: Malformed Input"                  blockDecompress.js:75  */
#define ORC_E_OFFSET_ZERO       (-3) /* "LZ4: Invalid Offset 0"                 blockDecompress.js:128 */
#define ORC_E_DICT_OOB          (-4) /* "LZ4: Dictionary Offset Out of Bounds"  blockDecompress.js:151 */
#define ORC_E_BAD_MAGIC         (-5) /* "LZ4: Invalid Magic Number"             bufferDecompress.js:60 */
#define ORC_E_BAD_VERSION       (-6) /* "LZ4: Unsupported Version N"            bufferDecompress.js:67 */
#define ORC_E_CONTENT_CHECKSUM  (-7) /* "LZ4: Content Checksum Error"           bufferDecompress.js:216 */
#define ORC_E_RANGE             (-8) /* JS RangeError from TypedArray.set (bufferDecompress.js:148) */

#define HASH_ENTRIES 16384

static inline uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void wr32(uint8_t *b, uint32_t v) {
    b[0] = (uint8_t)v; b[1] = (uint8_t)(v >> 8); b[2] = (uint8_t)(v >> 16); b[3] = (uint8_t)(v >> 24);
}
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* ---------------------------------------------------------------- xxHash32 */
/* xxhash32.js:9-13 */
#define P32_1 2654435761u
#define P32_2 2246822519u
#define P32_3 3266489917u
#define P32_4  668265263u
#define P32_5  374761393u

/* xxhash32.js:21-97.  The JS takes len|0, i.e. inputs must be < 2 GiB. */
uint32_t orc_xxh32(const uint8_t *in, uint64_t len64, uint32_t seed) {
    uint32_t len = (uint32_t)len64;
    const uint8_t *p = in, *end = in + len;
    uint32_t h;
    if (len >= 16) {                                   /* :27-65 */
        const uint8_t *limit = end - 16;
        uint32_t v1 = seed + P32_1 + P32_2, v2 = seed + P32_2, v3 = seed, v4 = seed - P32_1;
        do {
            v1 = rotl32(v1 + rd32(p) * P32_2, 13) * P32_1;
            v2 = rotl32(v2 + rd32(p + 4) * P32_2, 13) * P32_1;
            v3 = rotl32(v3 + rd32(p + 8) * P32_2, 13) * P32_1;
            v4 = rotl32(v4 + rd32(p + 12) * P32_2, 13) * P32_1;
            p += 16;
        } while (p <= limit);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + P32_5;                              /* :67 */
    }
    h += len;                                          /* :70 */
    while (p + 4 <= end) {                             /* :73-80 */
        h = rotl32(h + rd32(p) * P32_3, 17) * P32_4;
        p += 4;
    }
    while (p < end) {                                  /* :83-88 */
        h = rotl32(h + (uint32_t)(*p) * P32_5, 11) * P32_1;
        p++;
    }
    h ^= h >> 15; h *= P32_2; h ^= h >> 13; h *= P32_3; h ^= h >> 16;   /* :91-95 */
    return h;
}

void orc_xxh32_batch(const uint8_t *base, const uint64_t *off, const uint32_t *len,
                     uint32_t n, uint32_t seed, uint32_t *out) {
    for (uint32_t i = 0; i < n; i++) out[i] = orc_xxh32(base + off[i], len[i], seed);
}

/* ------------------------------------------------------------ compressBlock */
/* JS typed-array stores past the end are dropped silently (blockCompress.js has no
 * bounds checks).  PUT/COPY reproduce that when `cap` is short; when the caller
 * supplies >= orc_compress_bound() bytes the checks never fire. */
typedef struct { uint8_t *out; int64_t cap; } sink_t;
static inline void put(sink_t *s, int64_t i, uint8_t v) { if (i >= 0 && i < s->cap) s->out[i] = v; }
static inline void copy_in(sink_t *s, int64_t d, const uint8_t *src, int64_t n) {
    if (n <= 0) return;
    if (d >= 0 && d + n <= s->cap) { memcpy(s->out + d, src, (size_t)n); return; }
    for (int64_t k = 0; k < n; k++) put(s, d + k, src[k]);
}

uint64_t orc_compress_bound(uint64_t n) { return n + n / 255 + 16; }

/* Emit the literal-length part of a sequence: token high nibble + 255-run.
 * blockCompress.js:76-89 and :180-192.  Returns the new write index. */
static inline int64_t put_lit_len(sink_t *s, int64_t tokenPos, int64_t d, int32_t litLen) {
    if (litLen >= 15) {
        put(s, tokenPos, 0xF0);
        int32_t l = litLen - 15;
        while (l >= 255) { put(s, d++, 255); l -= 255; }
        put(s, d++, (uint8_t)l);
    } else {
        put(s, tokenPos, (uint8_t)(litLen << 4));
    }
    return d;
}

/*
 * blockCompress.js:31-233.  `src` is the whole working buffer (history lies in
 * [0,srcStart)), `table` is Int32[16384] holding position+1 (<=0 means empty) and is
 * read AND written (state carried to the next call in linked mode).
 * Returns bytes "written" (the JS count, which may exceed cap when cap is short).
 */
int32_t orc_compress_block(const uint8_t *src, int32_t srcStart, int32_t srcLen,
                           int32_t *table, uint8_t *out, int64_t outCap, int32_t outOffset) {
    sink_t s = { out, outCap };
    int32_t sIndex = srcStart;                       /* :32 */
    const int32_t sEnd = srcStart + srcLen;          /* :33 */
    const int32_t mflimit = sEnd - 12;               /* :34 */
    const int32_t matchLimit = sEnd - 5;             /* :35 */
    int64_t d = outOffset;                           /* :37 */
    int32_t anchor = sIndex;                         /* :38 */
    int32_t searchMatchCount = (1 << 6) + 3;         /* :40 */

    while (sIndex < mflimit) {                       /* :48 */
        uint32_t seq = rd32(src + sIndex);           /* :50 */
        uint32_t h = ((seq * 2654435761u) >> 18) & 16383u;   /* :53 */
        int32_t m = table[h] - 1;                    /* :54 */
        table[h] = sIndex + 1;                       /* :55  insert happens before the test */
        if (m < 0 || m == sIndex || (((uint32_t)(sIndex - m)) >> 16) > 0 || rd32(src + m) != seq) {  /* :62-63 */
            sIndex += (searchMatchCount++ >> 6);     /* :66-67 */
            continue;
        }
        searchMatchCount = (1 << 6) + 3;             /* :71 */

        int32_t litLen = sIndex - anchor;            /* :75 */
        int64_t tokenPos = d++;                      /* :76 */
        d = put_lit_len(&s, tokenPos, d, litLen);    /* :79-89 */
        copy_in(&s, d, src + anchor, litLen);        /* :92-140, all three strategies == memcpy */
        d += litLen;

        int32_t sp = sIndex + 4, mp = m + 4;         /* :143-144 */
        while (sp < matchLimit && src[sp] == src[mp]) { sp++; mp++; }   /* :147-150 forward only */

        int32_t matchLen = sp - sIndex;              /* :152 */
        int32_t offset = sIndex - m;                 /* :153 */
        put(&s, d++, (uint8_t)(offset & 0xff));      /* :156 */
        put(&s, d++, (uint8_t)((offset >> 8) & 0xff)); /* :157 */

        int32_t lenCode = matchLen - 4;              /* :160 */
        if (lenCode >= 15) {                         /* :161-168 */
            if (tokenPos >= 0 && tokenPos < s.cap) s.out[tokenPos] |= 0x0F;
            int32_t l = lenCode - 15;
            while (l >= 255) { put(&s, d++, 255); l -= 255; }
            put(&s, d++, (uint8_t)l);
        } else {
            if (tokenPos >= 0 && tokenPos < s.cap) s.out[tokenPos] |= (uint8_t)lenCode;   /* :170 */
        }
        sIndex = sp;                                 /* :173  positions inside the match are not inserted */
        anchor = sp;                                 /* :174 */
    }

    /* :179-230 last literals, always emitted (srcLen==0 -> one 0x00 byte) */
    int32_t litLen = sEnd - anchor;
    int64_t tokenPos = d++;
    d = put_lit_len(&s, tokenPos, d, litLen);
    copy_in(&s, d, src + anchor, litLen);
    d += litLen;
    return (int32_t)(d - outOffset);                 /* :232 */
}

/* Batched independent raw blocks, each with a fresh zero table and no history
 * (== compressBuffer's per-block behaviour with blockIndependence, bufferCompress.js:219,234-236,
 * for blocks that start at index 0 of their own view; positions are relative so the bytes are
 * identical to the frame path where positions are absolute). */
void orc_compress_blocks(const uint8_t *src, const uint64_t *off, const uint32_t *len, uint32_t n,
                         uint8_t *dst, const uint64_t *dst_off, uint32_t *comp_len) {
    int32_t *table = (int32_t *)malloc(HASH_ENTRIES * sizeof(int32_t));
    for (uint32_t i = 0; i < n; i++) {
        memset(table, 0, HASH_ENTRIES * sizeof(int32_t));
        comp_len[i] = (uint32_t)orc_compress_block(src + off[i], 0, (int32_t)len[i], table,
                                                   dst + dst_off[i], (int64_t)orc_compress_bound(len[i]), 0);
    }
    free(table);
}

/* Raw-batch with a shared prefix (BASELINE config 4): per message the caller-side state is
 * `prefix ++ msg` and a copy of `init_table` (NULL -> zero table); blockCompress.js:27,31. */
void orc_compress_blocks_prefix(const uint8_t *prefix, uint32_t prefix_len, const int32_t *init_table,
                                const uint8_t *src, const uint64_t *off, const uint32_t *len, uint32_t n,
                                uint8_t *dst, const uint64_t *dst_off, uint32_t *comp_len) {
    int32_t *table = (int32_t *)malloc(HASH_ENTRIES * sizeof(int32_t));
    uint32_t maxlen = 0;
    for (uint32_t i = 0; i < n; i++) if (len[i] > maxlen) maxlen = len[i];
    uint8_t *work = (uint8_t *)malloc((size_t)prefix_len + maxlen + 8);
    memcpy(work, prefix, prefix_len);
    for (uint32_t i = 0; i < n; i++) {
        if (init_table) memcpy(table, init_table, HASH_ENTRIES * sizeof(int32_t));
        else memset(table, 0, HASH_ENTRIES * sizeof(int32_t));
        memcpy(work + prefix_len, src + off[i], len[i]);
        comp_len[i] = (uint32_t)orc_compress_block(work, (int32_t)prefix_len, (int32_t)len[i], table,
                                                   dst + dst_off[i], (int64_t)orc_compress_bound(len[i]), 0);
    }
    free(work);
    free(table);
}

/* Dictionary warm-up hash used by compressBuffer (bufferCompress.js:191-203).  NOT the hash the
 * match finder probes with; reproduced because the warmed entries are state. */
static inline uint32_t jenkins_slot(uint32_t seq) {
    uint32_t h = seq;
    h = h + 2127912214u + (h << 12);                 /* :195 */
    h = h ^ 3345072700u ^ (h >> 19);                 /* :196  -949894596 */
    h = h + 374761393u + (h << 5);                   /* :197 */
    h = (h + 3550635116u) ^ (h << 9);                /* :198  JS precedence: (h + -744332180) ^ (h << 9) */
    h = h + 4251993797u + (h << 3);                  /* :199  -42973499 */
    h = h ^ 3042594569u ^ (h >> 16);                 /* :200  -1252372727 */
    return (h >> 18) & 16383u;                       /* :201 */
}

void orc_warm_table_jenkins(const uint8_t *work, int32_t dictLen, int32_t *table) {
    for (int32_t i = 0; i <= dictLen - 4; i++) table[jenkins_slot(rd32(work + i))] = i + 1;   /* :191-203 */
}

/* ---------------------------------------------------------- decompressBlock */
/*
 * blockDecompress.js:30-275 with LZ4-spec match-copy semantics (byte-by-byte forward copy).
 * `out` / `outLen` are the WHOLE output array (history = dictionary ++ out[0..outPos)), as in
 * the JS where output index 0 is the dictionary boundary (:142-147).
 *
 * Deviations from the literal JS, both deliberate and both documented in DESIGN.md:
 *  (1) match writes and length-byte reads are bounds-checked (the JS silently drops OOB stores
 *      and reads `undefined`); the error returned is the nearest reference error.
 *  (2) the JS "double-copy tail" for in-buffer matches (:219-250) is unguarded for
 *      matchLen 4..7 with offset >= 8 and overwrites up to 4 already-decoded bytes BEFORE the
 *      match (tailOut = endMatch-8 < outPos).  That is a reference defect (its own round trip
 *      fails on text); orc_decompress_block_literal() below reproduces it so the tests can
 *      show the divergence is confined to that case.  The north star requires exact round trip.
 * Returns bytes written (>=0) or ORC_E_*.
 */
int64_t orc_decompress_block(const uint8_t *in, int64_t inOff, int64_t inSize,
                             uint8_t *out, int64_t outLen, int64_t outOff,
                             const uint8_t *dict, int64_t dictLen) {
    int64_t ip = inOff, inEnd = inOff + inSize, op = outOff;
    if (!dict) dictLen = 0;
    while (ip < inEnd) {                                             /* :55 */
        uint32_t token = in[ip++];                                   /* :58 */
        int64_t litLen = token >> 4;                                 /* :61 */
        if (litLen == 15) {                                          /* :62-68 */
            uint32_t b;
            do {
                if (ip >= inEnd) return ORC_E_MALFORMED;
                b = in[ip++]; litLen += b;
            } while (b == 255);
        }
        if (op + litLen > outLen) return ORC_E_OUTPUT_TOO_SMALL;     /* :74 */
        if (ip + litLen > inEnd) return ORC_E_MALFORMED;             /* :75 */
        memcpy(out + op, in + ip, (size_t)litLen);                   /* :79-121 == memcpy */
        op += litLen; ip += litLen;
        if (ip >= inEnd) break;                                      /* :123 */

        if (ip + 2 > inEnd) return ORC_E_MALFORMED;
        int64_t offset = (int64_t)in[ip] | ((int64_t)in[ip + 1] << 8);   /* :126 */
        ip += 2;
        if (offset == 0) return ORC_E_OFFSET_ZERO;                   /* :128 */
        int64_t matchLen = token & 15;                               /* :131 */
        if (matchLen == 15) {                                        /* :132-138 */
            uint32_t b;
            do {
                if (ip >= inEnd) return ORC_E_MALFORMED;
                b = in[ip++]; matchLen += b;
            } while (b == 255);
        }
        matchLen += 4;                                               /* :139 */

        int64_t copySrc = op - offset;                               /* :142 */
        if (copySrc < 0) {                                           /* :145 dictionary branch */
            int64_t fromDict = -copySrc;
            int64_t di = dictLen + copySrc;
            if (fromDict > matchLen) fromDict = matchLen;
            if (di < 0 || di + fromDict > dictLen) return ORC_E_DICT_OOB;   /* :150-152 */
            if (op + matchLen > outLen) return ORC_E_OUTPUT_TOO_SMALL;
            memcpy(out + op, dict + di, (size_t)fromDict);           /* :157-191 == memcpy */
            op += fromDict;
            int64_t rem = matchLen - fromDict;                       /* :194 */
            const uint8_t *r = out + (op - offset);                  /* :197 == out index 0 */
            for (int64_t k = 0; k < rem; k++) out[op + k] = r[k];    /* :198 */
            op += rem;
        } else {
            if (op + matchLen > outLen) return ORC_E_OUTPUT_TOO_SMALL;
            if (offset >= matchLen) memcpy(out + op, out + copySrc, (size_t)matchLen);
            else for (int64_t k = 0; k < matchLen; k++) out[op + k] = out[copySrc + k];  /* :202-271 */
            op += matchLen;
        }
    }
    return op - outOff;                                              /* :274 */
}

/*
 * Line-faithful blockDecompress.js:30-275, including the unguarded double-copy tail (:232-250),
 * silent dropping of out-of-range stores and zero-reads past the input.  Used only to document
 * where the reference decoder itself departs from LZ4 (tests/test_oracle_reference_defect.py).
 */
static inline uint8_t ld(const uint8_t *a, int64_t n, int64_t i) { return (i >= 0 && i < n) ? a[i] : 0; }
static inline void st(uint8_t *a, int64_t n, int64_t i, uint8_t v) { if (i >= 0 && i < n) a[i] = v; }

int64_t orc_decompress_block_literal(const uint8_t *in, int64_t inTotal, int64_t inOff, int64_t inSize,
                                     uint8_t *out, int64_t outLen, int64_t outOff,
                                     const uint8_t *dict, int64_t dictLen) {
    int64_t ip = inOff, inEnd = inOff + inSize, op = outOff;
    if (!dict) dictLen = 0;
    while (ip < inEnd) {
        uint32_t token = ld(in, inTotal, ip++);
        int64_t litLen = (token >> 4) & 15;
        if (litLen == 15) { uint32_t b; do { b = ld(in, inTotal, ip++); litLen += b; } while (b == 255); }
        int64_t endLit = op + litLen;
        if (endLit > outLen) return ORC_E_OUTPUT_TOO_SMALL;
        if (ip + litLen > inEnd) return ORC_E_MALFORMED;
        for (int64_t k = 0; k < litLen; k++) st(out, outLen, op + k, ld(in, inTotal, ip + k));
        op = endLit; ip += litLen;
        if (ip >= inEnd) break;
        int64_t offset = ld(in, inTotal, ip) | (ld(in, inTotal, ip + 1) << 8);
        ip += 2;
        if (offset == 0) return ORC_E_OFFSET_ZERO;
        int64_t matchLen = token & 15;
        if (matchLen == 15) { uint32_t b; do { b = ld(in, inTotal, ip++); matchLen += b; } while (b == 255); }
        matchLen += 4;
        int64_t copySrc = op - offset;
        if (copySrc < 0) {
            int64_t fromDict = -copySrc;
            copySrc = dictLen + copySrc;
            if (fromDict > matchLen) fromDict = matchLen;
            if (copySrc < 0 || copySrc + fromDict > dictLen) return ORC_E_DICT_OOB;
            for (int64_t k = 0; k < fromDict; k++) st(out, outLen, op + k, dict[copySrc + k]);
            op += fromDict;
            int64_t rem = matchLen - (op - endLit);
            if (rem > 0) {
                int64_t endMatch = op + rem, rp = op - offset;
                while (op < endMatch) { st(out, outLen, op, ld(out, outLen, rp)); op++; rp++; }
            }
        } else if (offset == 1) {                                    /* :204-207 fill */
            uint8_t v = ld(out, outLen, copySrc);
            for (int64_t k = 0; k < matchLen; k++) st(out, outLen, op + k, v);
            op += matchLen;
        } else if (offset >= matchLen && matchLen > 16) {            /* :209-212 copyWithin */
            int64_t n = matchLen;
            if (op + n > outLen) n = outLen - op;
            if (copySrc + n > outLen) n = outLen - copySrc;
            if (n > 0) memmove(out + op, out + copySrc, (size_t)n);
            op += matchLen;
        } else {
            int64_t endMatch = op + matchLen, rp = copySrc;
            if (offset >= 8) {                                       /* :219-251 */
                int64_t body = endMatch - 8;
                while (op < body) for (int k = 0; k < 8; k++) { st(out, outLen, op, ld(out, outLen, rp)); op++; rp++; }
                if (op < endMatch) {
                    int64_t tailOut = endMatch - 8;
                    int64_t tailSrc = rp + (endMatch - op) - 8;
                    for (int k = 0; k < 8; k++) st(out, outLen, tailOut + k, ld(out, outLen, tailSrc + k));
                    op = endMatch;
                }
            } else {                                                 /* :252-268 == forward byte copy */
                while (op < endMatch) { st(out, outLen, op, ld(out, outLen, rp)); op++; rp++; }
            }
        }
    }
    return op - outOff;
}

/* Batched raw decode, each block with its own output base and only `dict` as history
 * (SURVEY 8d config 4 decode side: decompressBlock(msgComp,0,len,out_i,0,dict)). */
void orc_decompress_blocks(const uint8_t *src, const uint64_t *off, const uint32_t *len, uint32_t n,
                           uint8_t *dst, const uint64_t *dst_off, const uint32_t *dst_cap,
                           const uint8_t *dict, uint32_t dictLen, uint32_t *out_len, int32_t *status) {
    for (uint32_t i = 0; i < n; i++) {
        int64_t r = orc_decompress_block(src + off[i], 0, len[i], dst + dst_off[i], dst_cap[i], 0, dict, dictLen);
        status[i] = r < 0 ? (int32_t)r : 0;
        out_len[i] = r < 0 ? 0 : (uint32_t)r;
    }
}

/* ----------------------------------------------------------- compressBuffer */
/* bufferCompress.js:77-82 */
static int block_id_for(int64_t bytes) {
    if (bytes <= 0 || bytes <= 65536) return 4;
    if (bytes <= 262144) return 5;
    if (bytes <= 1048576) return 6;
    return 7;
}
static const int32_t BLOCK_MAX[8] = { 0, 0, 0, 0, 65536, 262144, 1048576, 4194304 };   /* :43-48 */

uint64_t orc_frame_bound(uint64_t len) {
    /* worst case of THIS writer: header 19 + per-block (4 size + stored data [+4 cksum]) at 64 KiB + 8 */
    return 19 + len + (len / 65536 + 1) * 8 + 8 + 64;
}

/*
 * bufferCompress.js:100-259.  `blockChecksum` is an extension that does not exist in the
 * reference writer: FLG |= 0x10 and xxh32(stored block bytes) after each block,
 * per the LZ4 frame spec; with blockChecksum=0 the bytes are the reference's.
 * Returns the frame length; writes at most outCap bytes (JS drops the rest).
 */
int64_t orc_compress_buffer(const uint8_t *input, int64_t inLen,
                            const uint8_t *dictionary, int64_t dictTotal,
                            int64_t maxBlockSize, int blockIndependence, int contentChecksum,
                            int addContentSize, int blockChecksum,
                            uint8_t *out, int64_t outCap) {
    const int32_t len = (int32_t)inLen;                         /* :127  len|0 */
    const uint8_t *work = input;
    uint8_t *owned = NULL;
    int32_t dictLen = 0, start = 0;
    int haveDict = 0;
    uint32_t dictId = 0;
    if (dictionary && dictTotal > 0) {                          /* :109-125 */
        haveDict = 1;
        dictId = orc_xxh32(dictionary, (uint64_t)dictTotal, 0); /* :112 */
        const uint8_t *win = dictTotal > 65536 ? dictionary + (dictTotal - 65536) : dictionary;   /* :115 */
        dictLen = (int32_t)(dictTotal > 65536 ? 65536 : dictTotal);
        owned = (uint8_t *)malloc((size_t)dictLen + (size_t)len + 8);
        memcpy(owned, win, (size_t)dictLen);
        memcpy(owned + dictLen, input, (size_t)len);
        work = owned; start = dictLen;
    }
    const int bdId = block_id_for(maxBlockSize);                /* :128 */
    const int32_t resolved = BLOCK_MAX[bdId];                   /* :129 */

    /* Work in a private worst-case buffer and copy out min(n,outCap) at the end: the observable
     * result equals "stores beyond the caller's buffer are dropped". */
    uint64_t bound = orc_frame_bound((uint64_t)len) + (uint64_t)(len / 255);
    uint8_t *o = (uint8_t *)malloc(bound);
    int64_t op = 0;
    o[op++] = 0x04; o[op++] = 0x22; o[op++] = 0x4D; o[op++] = 0x18;   /* :147 */
    uint8_t flg = 1 << 6;                                       /* :150 */
    if (blockIndependence) flg |= 0x20;                         /* :151 */
    if (contentChecksum) flg |= 0x04;                           /* :152 */
    if (haveDict) flg |= 0x01;                                  /* :153 */
    if (addContentSize) flg |= 0x08;                            /* :154 */
    if (blockChecksum) flg |= 0x10;                             /* extension */
    o[op++] = flg;
    o[op++] = (uint8_t)((bdId & 7) << 4);                       /* :158 */
    const int64_t headerStart = 4;
    if (addContentSize) {                                       /* :163-168 */
        wr32(o + op, (uint32_t)len); op += 4;
        wr32(o + op, 0); op += 4;                               /* (len/2^32)|0 is 0 for int32 len >= 0 */
    }
    if (haveDict) { wr32(o + op, dictId); op += 4; }            /* :171-174 */
    uint32_t hh = orc_xxh32(o + headerStart, (uint64_t)(op - headerStart), 0);   /* :177 */
    o[op++] = (uint8_t)((hh >> 8) & 0xFF);                      /* :178 */

    int32_t *table = (int32_t *)calloc(HASH_ENTRIES, sizeof(int32_t));   /* :182-183 */
    if (dictLen > 0) orc_warm_table_jenkins(work, dictLen, table);       /* :186-204 */

    int32_t srcPos = start;
    const int32_t totalEnd = start + len;
    while (srcPos < totalEnd) {                                 /* :209 */
        int32_t end = srcPos + resolved; if (end > totalEnd || end < srcPos) end = totalEnd;   /* :210 */
        int32_t blockSize = end - srcPos;
        int64_t sizePos = op; op += 4;                          /* :214-215 */
        int32_t compSize = orc_compress_block(work, srcPos, blockSize, table, o, (int64_t)bound, (int32_t)op);  /* :219 */
        int64_t dataPos = op;
        if (compSize > 0 && compSize < blockSize) {             /* :221-224 */
            wr32(o + sizePos, (uint32_t)compSize);
            op += compSize;
        } else {                                                /* :225-231 stored */
            wr32(o + sizePos, (uint32_t)blockSize | 0x80000000u);
            memcpy(o + op, work + srcPos, (size_t)blockSize);
            op += blockSize;
        }
        if (blockChecksum) { wr32(o + op, orc_xxh32(o + dataPos, (uint64_t)(op - dataPos), 0)); op += 4; }
        if (blockIndependence) memset(table, 0, HASH_ENTRIES * sizeof(int32_t));   /* :234-236 */
        srcPos = end;
    }
    wr32(o + op, 0); op += 4;                                   /* :244-245 EndMark */
    if (contentChecksum) { wr32(o + op, orc_xxh32(input, (uint64_t)len, 0)); op += 4; }   /* :248-252 */

    int64_t n = op < outCap ? op : outCap;
    if (n > 0) memcpy(out, o, (size_t)n);
    free(o); free(table); free(owned);
    return op;
}

/* --------------------------------------------------------- decompressBuffer */
/*
 * bufferDecompress.js:51-220.  Returns decoded length or ORC_E_*; *outp receives a malloc'd
 * buffer (free with orc_free).  *badVersion receives the version for ORC_E_BAD_VERSION.
 * Follows the JS strategies: direct write when contentSize > 0 (:97-107), otherwise chunked with a
 * 64 KiB rolling window (:108-124,:157-185).  BD, dictID, header checksum and block checksums are
 * skipped unverified exactly as in the JS (:75,:89,:92,:191).
 */
int64_t orc_decompress_buffer(const uint8_t *data, int64_t len,
                              const uint8_t *dictionary, int64_t dictLen,
                              int verifyChecksum, uint8_t **outp, int *badVersion) {
    *outp = NULL;
    int64_t pos = 0;
    if (len < 4 || rd32(data) != 0x184D2204u) return ORC_E_BAD_MAGIC;      /* :59-61 */
    pos = 4;
    if (pos >= len) { if (badVersion) *badVersion = 0; return ORC_E_BAD_VERSION; }   /* data[4]==undefined -> version 0 */
    uint8_t flg = data[pos++];                                              /* :65 */
    int version = (flg & 0xC0) >> 6;
    if (version != 1) { if (badVersion) *badVersion = version; return ORC_E_BAD_VERSION; }   /* :67 */
    int hasBlockChecksum = (flg & 0x10) != 0, hasContentSize = (flg & 0x08) != 0;
    int hasContentChecksum = (flg & 0x04) != 0, hasDictId = (flg & 0x01) != 0;
    pos++;                                                                  /* :75 BD skipped */
    double expected = 0;
    if (hasContentSize) {                                                   /* :79-86 */
        uint32_t lo = pos + 4 <= len ? rd32(data + pos) : 0, hi = pos + 8 <= len ? rd32(data + pos + 4) : 0;
        pos += 8;
        expected = (double)hi * 4294967296.0 + (double)lo;
    }
    if (hasDictId) pos += 4;                                                /* :89 */
    pos++;                                                                  /* :92 HC skipped */
    const int direct = expected > 0;                                        /* :97 */
    if (!dictionary) dictLen = 0;

    uint8_t *result = NULL; int64_t resultLen = 0, resultPos = 0;
    /* chunked state */
    uint8_t *acc = NULL; int64_t accLen = 0, accCap = 0;
    uint8_t *window = NULL; int64_t windowPos = 0;
    uint8_t *workspace = NULL;
    const int64_t W = 65536, WS = 4194304;
    int64_t rc = 0;

    if (direct) {
        resultLen = (int64_t)expected;
        result = (uint8_t *)calloc((size_t)resultLen + 1, 1);               /* :107 */
    } else {
        window = (uint8_t *)calloc((size_t)W, 1);                           /* :111 */
        workspace = (uint8_t *)malloc((size_t)WS);
        if (dictLen > 0) {                                                  /* :114-123 */
            if (dictLen > W) { memcpy(window, dictionary + (dictLen - W), (size_t)W); windowPos = W; }
            else { memcpy(window, dictionary, (size_t)dictLen); windowPos = dictLen; }
        }
    }

    while (pos < len) {                                                     /* :133 */
        uint32_t bs = pos + 4 <= len ? rd32(data + pos) : 0;                /* :135 */
        pos += 4;
        if (bs == 0) break;                                                 /* :139 */
        int stored = (bs & 0x80000000u) != 0;
        int64_t actual = bs & 0x7FFFFFFFu;
        if (pos + actual > len) { rc = ORC_E_MALFORMED; goto done; }        /* JS would read undefined; refuse */
        if (direct) {
            if (stored) {                                                   /* :147-149 */
                if (resultPos + actual > resultLen) { rc = ORC_E_RANGE; goto done; }
                memcpy(result + resultPos, data + pos, (size_t)actual);
                resultPos += actual;
            } else {                                                        /* :153 */
                int64_t n = orc_decompress_block(data, pos, actual, result, resultLen, resultPos, dictionary, dictLen);
                if (n < 0) { rc = n; goto done; }
                resultPos += n;
            }
        } else {
            const uint8_t *chunk; int64_t chunkLen;
            if (stored) { chunk = data + pos; chunkLen = actual; }          /* :160 */
            else {                                                          /* :164-167 */
                int64_t n = orc_decompress_block(data, pos, actual, workspace, WS, 0,
                                                 windowPos > 0 ? window : NULL, windowPos);
                if (n < 0) { rc = n; goto done; }
                chunk = workspace; chunkLen = n;
            }
            if (accLen + chunkLen > accCap) {
                accCap = (accLen + chunkLen) * 2 + 65536;
                acc = (uint8_t *)realloc(acc, (size_t)accCap);
            }
            memcpy(acc + accLen, chunk, (size_t)chunkLen); accLen += chunkLen;
            if (chunkLen >= W) { memcpy(window, chunk + (chunkLen - W), (size_t)W); windowPos = W; }      /* :173-175 */
            else if (windowPos + chunkLen <= W) { memcpy(window + windowPos, chunk, (size_t)chunkLen); windowPos += chunkLen; }  /* :176-178 */
            else {                                                          /* :179-185 */
                int64_t keep = W - chunkLen;
                memmove(window, window + (windowPos - keep), (size_t)keep);
                memcpy(window + keep, chunk, (size_t)chunkLen);
                windowPos = W;
            }
        }
        pos += actual;                                                      /* :188 */
        if (hasBlockChecksum) pos += 4;                                     /* :191 skipped, not verified */
    }
    if (!direct) { result = acc; acc = NULL; resultLen = accLen; if (!result) result = (uint8_t *)calloc(1, 1); }
    if (hasContentChecksum && verifyChecksum) {                             /* :213-217 */
        uint32_t storedHash = pos + 4 <= len ? rd32(data + pos) : 0;
        if (storedHash != orc_xxh32(result, (uint64_t)resultLen, 0)) { rc = ORC_E_CONTENT_CHECKSUM; goto done; }
    }
    rc = resultLen;
done:
    free(acc); free(window); free(workspace);
    if (rc < 0) { free(result); result = NULL; }
    *outp = result;
    return rc;
}

void orc_free(void *p) { free(p); }
