// dlz4_parse.cuh -- compressBlock (blockCompress.js:31-233) cut into its serial and its parallel half.
//
// The reference interleaves two things that have nothing to do with each other on a GPU: FINDING the matches (a serial chain:
// which positions are probed depends on every earlier match, and the table holds exactly the probed positions) and WRITING
// the sequences (a pure function of the match list and the source bytes).  Only the first is a dependent chain, so only the
// first belongs in the one-warp-per-block kernel whose length is the whole cost of compression:
//
//   k_parse_fresh16    one warp per block, 16-bit table in shared memory: the exact probe schedule, table updates, candidate
//                      verification and match extension of blockCompress.js:48-71,143-150 -- and nothing else.  Output: the
//                      block's matches in order, one 8-byte record each (position, length, offset).  The 64-position window is
//                      the one of dlz4_wide.cuh (two positions per lane, same-slot pairs found by tagging the table, successor
//                      links computed by all lanes, a serial part that only follows the links and writes the records).
//   k_encode_blocks    one warp per block, 32 sequences per step: literal lengths from neighbouring records, an in-warp prefix
//                      sum of the sequence sizes, then every lane writes its own sequence (token, length bytes, literals,
//                      offset, length bytes: blockCompress.js:75-174) and the warp the final literals (:179-230).  No
//                      dependent chain beyond the running output offset; memory-bound.
//
// The bytes are those of the serial loop: the match list IS the serial loop's list (same probes, same table states, same
// extensions), and the encoding of a given list is unique.
#pragma once

namespace dlz4 {

// record (8 bytes): x = position relative to the block start (bits 0..15) | offset (bits 16..31), y = match length
__device__ __forceinline__ uint2 rec_pack(uint32_t pos, uint32_t ml, uint32_t offset) { return make_uint2(pos | (offset << 16), ml); }

// Matches of the block [start, start + len) of `base` (a fresh block: empty table, no history), in order, into rec[];
// returns their number.  All 32 lanes call this with identical arguments.  len <= 65536 (Tab16) is the caller's business.
template <class Tab>
__device__ uint32_t parse_block64_warp(const uint8_t *__restrict__ base, const int32_t start, const int32_t len, Tab &T,
                                       uint64_t *__restrict__ rec) {
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    const int32_t sEnd = start + len;
    const int32_t mflimit = sEnd - 12;                                          // blockCompress.js:34
    const int32_t matchLimit = sEnd - 5;                                        // :35
    int32_t sIndex = start;
    uint32_t smc = 67;                                                          // :40 searchMatchCount
    uint32_t nrec = 0;
    const SrcFlat S{base};

    // virtual byte address of index v: A0 + v, with A0 = the base pointer's offset inside its 128-byte line
    const uint32_t A0 = (uint32_t)(reinterpret_cast<uintptr_t>(base) & 127u);
    const int32_t wlo = start - (int32_t)((A0 + (uint32_t)start) & 3u);       // first word holding block bytes
    const int32_t whi = sEnd;
    uint32_t R0 = 0, R1 = 0, R2 = 0, R3 = 0;      // word `lane` of the lines Lc .. Lc+3 (the next 512 bytes of the block)
    uint32_t Lc = 0;
    bool cold = true;
    auto load_line = [&](uint32_t l) -> uint32_t {
        const int32_t idx = (int32_t)(l * 128u + 4u * lane) - (int32_t)A0;
        return (idx >= wlo && idx < whi) ? __ldg(reinterpret_cast<const uint32_t *>(base + idx)) : 0u;
    };

    while (sIndex < mflimit) {                                                  // :48
        if (smc <= 96u && sIndex + 100 <= sEnd) {
            // ---- dense window: the skip schedule steps by 1 (:66-67 with searchMatchCount < 128), so the next probes are the
            //      64 consecutive positions w .. w+63 until a match is taken
            const int32_t w = sIndex;
            const uint32_t va = A0 + (uint32_t)w;
            const uint32_t wmis = va & 3u;
            const uint32_t L = va >> 7;
            if (cold || L != Lc) {
                const uint32_t d = cold ? 4u : L - Lc;
                if (d == 1u) { R0 = R1; R1 = R2; R2 = R3; R3 = load_line(L + 3u); }
                else if (d == 2u) { R0 = R2; R1 = R3; R2 = load_line(L + 2u); R3 = load_line(L + 3u); }
                else if (d == 3u) { R0 = R3; R1 = load_line(L + 1u); R2 = load_line(L + 2u); R3 = load_line(L + 3u); }
                else { R0 = load_line(L); R1 = load_line(L + 1u); R2 = load_line(L + 2u); R3 = load_line(L + 3u); }
                Lc = L; cold = false;
            }
            // source bytes: word (wa/4 + lane) of the window, then each lane's 64 bytes from 17 shuffles
            const uint32_t wo = ((va & 127u) >> 2) + lane;                     // word offset from the start of line L (< 64)
            const uint32_t x0 = __shfl_sync(FULL, R0, wo), x1 = __shfl_sync(FULL, R1, wo);
            const uint32_t Tw = wo < 32u ? x0 : x1;
            const uint32_t o = wmis + lane, wi = o >> 2, sh = (o & 3u) * 8u;
            uint32_t Sx[16];                                                   // bytes pa .. pa+63 (pb = pa + 32)
            {
                uint32_t tprev = __shfl_sync(FULL, Tw, wi);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t tnext = __shfl_sync(FULL, Tw, wi + k + 1);
                    Sx[k] = __funnelshift_r(tprev, tnext, sh);
                    tprev = tnext;
                }
            }
            const int32_t pa = w + (int32_t)lane, pb = pa + 32;
            const uint32_t ha = (Sx[0] * 2654435761u) >> 18, hb = (Sx[8] * 2654435761u) >> 18;     // :53
            // table look-up (:54, the state from before the window) and candidate loads
            const uint32_t olda = tab_raw(T, ha), oldb = tab_raw(T, hb);
            const int32_t canda = tab_dec(T, olda), candb = tab_dec(T, oldb);
            const bool oka = canda >= 0 && canda != pa && (((uint32_t)(pa - canda)) >> 16) == 0;   // :62
            const bool okb = candb >= 0 && candb != pb && (((uint32_t)(pb - candb)) >> 16) == 0;
            const int32_t la = oka ? canda : pa, lb_ = okb ? candb : pb;       // no candidate: read at the own address (valid, cached)
            const uint32_t csa = (A0 + (uint32_t)la) & 15u, csb = (A0 + (uint32_t)lb_) & 15u;
            const uint4 *cqa = reinterpret_cast<const uint4 *>(base + (la - (int32_t)csa));
            const uint4 *cqb = reinterpret_cast<const uint4 *>(base + (lb_ - (int32_t)csb));
            const uint4 qa0 = __ldg(cqa), qa1 = __ldg(cqa + 1), qa2 = __ldg(cqa + 2);
            const uint4 qb0 = __ldg(cqb), qb1 = __ldg(cqb + 1), qb2 = __ldg(cqb + 2);
            // same-slot pairs inside the window: tag every slot, read it back (dlz4_wide.cuh)
            const uint32_t taga = tab_enc(T, pa), tagb = tab_enc(T, pb);
            __syncwarp();
            tab_set_raw(T, ha, taga);
            __syncwarp();
            tab_set_raw(T, hb, tagb);
            __syncwarp();
            const uint32_t ra = tab_raw(T, ha), rb = tab_raw(T, hb);
            const bool losta = ra != taga, lostb = rb != tagb;
            uint32_t trunc = 64;                                               // positions [0, trunc) take part in this window
            const uint32_t conf = __ballot_sync(FULL, losta || lostb);
            if (conf) {
                uint32_t inv = 64;
                if (losta) { const uint32_t wr = (uint32_t)(tab_dec(T, ra) - w); inv = min(lane, wr); }
                if (lostb) { const uint32_t wr = (uint32_t)(tab_dec(T, rb) - w); inv = min(inv, min(lane + 32u, wr)); }
                trunc = __reduce_min_sync(FULL, inv) + 1u;
                __syncwarp();
                if (lane >= trunc) tab_set_raw(T, ha, olda);
                if (lane + 32u >= trunc) tab_set_raw(T, hb, oldb);
                __syncwarp();
            }
            // verify (:63) + pre-extend to 32 bytes (:147-150)
            const uint32_t va_ = wide_verify(qa0, qa1, qa2, csa, Sx), vb_ = wide_verify(qb0, qb1, qb2, csb, Sx + 8);
            const uint32_t mla = oka ? va_ : 0u, mlb = okb ? vb_ : 0u;
            uint32_t HMa = __ballot_sync(FULL, mla != 0u), HMb = __ballot_sync(FULL, mlb != 0u);
            if (trunc < 64u) {
                HMa &= trunc >= 32u ? FULL : ((1u << trunc) - 1u);
                HMb &= trunc >= 32u ? ((1u << (trunc - 32u)) - 1u) : 0u;
            }
            auto first_hit_from = [&](uint32_t e) -> uint32_t {                // first hit at or behind relative position e (64: none)
                const uint32_t ma = e < 32u ? (HMa >> e) << e : 0u;
                const uint32_t mb = e < 32u ? HMb : (e < 64u ? (HMb >> (e - 32u)) << (e - 32u) : 0u);
                return ma ? (uint32_t)__ffs(ma) - 1u : (mb ? 31u + (uint32_t)__ffs(mb) : 64u);
            };
            // every hit position's successor in the chain of heads: the first hit at or behind its match's end.  (A b-position's
            // match ends behind position 35: only the upper mask matters.)
            const uint32_t suca = first_hit_from(lane + mla);
            const uint32_t eb_ = lane + mlb;                                   // end of b's match, relative to position 32
            const uint32_t mbb = eb_ < 32u ? (HMb >> eb_) << eb_ : 0u;
            const uint32_t sucb = mbb ? 31u + (uint32_t)__ffs(mbb) : 64u;
            uint32_t pack = (mla | (suca << 6)) | ((mlb | (sucb << 6)) << 16);           // ml: 6 bits, successor: 7 bits, per half
            const uint32_t offa = (uint32_t)(pa - canda), offb = (uint32_t)(pb - candb);   // :153 (only a head's is used)
            // ---- the chain of heads (uniform, serial): follow the links; a head's lane notes that it is one
            const uint32_t dl0 = 128u - smc;                                   // dense probing reaches [0, dl0) before the first match
            uint32_t hl = first_hit_from(0u);
            if (hl >= dl0) hl = 64u;
            const bool any_head = hl < 64u;
            uint32_t myhead = 0;                                               // bit 0: position a is a head, bit 1: position b
            uint32_t mlxa = mla, mlxb = mlb;                                   // match lengths, a long match's continued
            uint32_t cur = 0;
            while (hl < 64u) {
                const uint32_t pk = __shfl_sync(FULL, pack, hl);
                const uint32_t fld = (pk >> ((hl >> 1) & 16u)) & 0xFFFFu;
                uint32_t mlh = fld & 63u;
                uint32_t nxt = fld >> 6;
                if (mlh == 32u && matchLimit - (w + (int32_t)hl) > 32) {
                    // long match: continue cooperatively, 128 bytes per round
                    const int32_t s0 = w + (int32_t)hl;
                    const int32_t m0 = __shfl_sync(FULL, (hl & 32u) ? candb : canda, hl);
                    for (int32_t eb = 32;; eb += 128) {
                        const int32_t q = s0 + eb + 4 * (int32_t)lane;
                        int32_t nv = matchLimit - q;
                        nv = nv > 4 ? 4 : nv;
                        int32_t eq = 0;
                        if (nv > 0) {
                            const uint32_t x = S.ld32(q) ^ S.ld32(m0 + eb + 4 * (int32_t)lane);
                            eq = x ? ((__ffs(x) - 1) >> 3) : 4;
                            eq = eq < nv ? eq : nv;
                        }
                        const uint32_t stop = __ballot_sync(FULL, eq < 4);
                        if (stop) {
                            const int l = __ffs(stop) - 1;
                            mlh = (uint32_t)(eb + 4 * l + __shfl_sync(FULL, eq, l));
                            break;
                        }
                    }
                    nxt = first_hit_from(hl + mlh);
                    if (lane == (hl & 31u)) { if (hl & 32u) mlxb = mlh; else mlxa = mlh; }
                }
                if (lane == (hl & 31u)) myhead |= 1u + (hl >> 5);
                cur = hl + mlh;
                hl = nxt;
            }
            // the window ends where dense probing, the cut or the 64 positions end -- or behind the last match
            const uint32_t stop = any_head ? trunc : min(trunc, dl0);         // <= 64
            // ---- everything else by all lanes at once.  Heads as masks; a position is inside a match when the nearest head
            //      below it reaches over it; a head writes its record at the rank of its bit.
            const uint32_t Ha = __ballot_sync(FULL, myhead & 1u), Hb = __ballot_sync(FULL, myhead & 2u);
            // end (exclusive, window-relative, saturated at 255) of every head's match, for the neighbours to read
            const uint32_t enda = min(lane + mlxa, 255u), endb = min(lane + 32u + mlxb, 255u);
            const uint32_t ends = enda | (endb << 8);
            // nearest head strictly below a-position `lane`: in Ha; below b-position: in Hb, else the last head of Ha
            const uint32_t ma_lo = Ha & lt, mb_lo = Hb & lt;
            const int pha = ma_lo ? 31 - __clz(ma_lo) : -1;
            const int phb = mb_lo ? 63 - __clz(mb_lo) : (Ha ? 31 - __clz(Ha) : -1);
            const uint32_t ea = __shfl_sync(FULL, ends, pha & 31), eb2 = __shfl_sync(FULL, ends, phb & 31);
            const uint32_t end_pha = pha < 0 ? 0u : (ea & 255u);
            const uint32_t end_phb = phb < 0 ? 0u : (phb >= 32 ? eb2 >> 8 : eb2 & 255u);
            const bool ina = lane < end_pha, inb = lane + 32u < end_phb;      // strictly inside (a head is never below its own end's start)
            const bool pra = lane < stop && !ina;                              // probed (:55): below the window's end, not inside a match
            const bool prb = lane + 32u < stop && !inb;
            if (conf) {
                if (lane < trunc) tab_set_raw(T, ha, pra ? taga : olda);
                if (lane + 32u < trunc) tab_set_raw(T, hb, prb ? tagb : oldb);
            } else {
                if (!pra) tab_set_raw(T, ha, olda);
                if (!prb) tab_set_raw(T, hb, oldb);
            }
            // records: position | offset << 16, length
            uint2 *const rec2 = reinterpret_cast<uint2 *>(rec) + nrec;
            const uint32_t prel = (uint32_t)(w - start) + lane;
            if (myhead & 1u) rec2[__popc(ma_lo)] = make_uint2(prel | (offa << 16), mlxa);
            if (myhead & 2u) rec2[__popc(Ha) + __popc(mb_lo)] = make_uint2((prel + 32u) | (offb << 16), mlxb);
            nrec += __popc(Ha) + __popc(Hb);
            if (cur < stop) {
                smc = (any_head ? 67u : smc) + (stop - cur);                   // trailing misses
                sIndex = w + (int32_t)stop;
            } else {
                smc = 67u;                                                     // :71
                sIndex = w + (int32_t)cur;
            }
            __syncwarp();
            continue;
        }

        // ---- batch step: sparse schedule and block tail.  Lane k probes the k-th upcoming position of the skip schedule
        //      (:66-67); lanes behind the first hit do not count (identical to compress_block_warp's loop body)
        const uint32_t base_sum = skip_sum(smc);
        const int32_t p = sIndex + (int32_t)(skip_sum(smc + lane) - base_sum);
        const bool valid = p < mflimit;
        uint32_t seq = 0, h = 0x10000u + lane;
        int32_t cand = -1;
        if (valid) {
            seq = S.ld32(p);
            h = (seq * 2654435761u) >> 18;
            cand = T.get(h);
        }
        const uint32_t same = __match_any_sync(FULL, h);
        const uint32_t prev = same & lt;
        const int j = prev ? 31 - __clz(prev) : (int)lane;
        const int32_t pj = __shfl_sync(FULL, p, j);
        const uint32_t sj = __shfl_sync(FULL, seq, j);
        uint32_t cseq = sj;
        if (prev) cand = pj;
        const bool ok = valid && cand >= 0 && cand != p && (((uint32_t)(p - cand)) >> 16) == 0;
        if (ok && !prev) cseq = S.ld32(cand);
        const bool hit = ok && cseq == seq;
        const uint32_t hits = __ballot_sync(FULL, hit);
        const uint32_t vmask = __ballot_sync(FULL, valid);
        const int hl = __ffs(hits) - 1;
        const uint32_t commit = hits ? ((2u << hl) - 1u) : vmask;
        if (((commit >> lane) & 1u) && ((same & commit) >> lane) == 1u) T.put(h, p);
        __syncwarp();
        if (!hits) {
            if (vmask != FULL) break;                            // ran into mflimit: loop ends
            sIndex += (int32_t)(skip_sum(smc + 32u) - base_sum);
            smc += 32u;
            continue;
        }
        const int32_t s0 = __shfl_sync(FULL, p, hl);
        const int32_t m0 = __shfl_sync(FULL, cand, hl);
        smc = 67;
        int32_t ml;
        for (int32_t eb = 4;; eb += 128) {
            const int32_t q = s0 + eb + 4 * (int32_t)lane;
            int32_t nv = matchLimit - q;
            nv = nv > 4 ? 4 : nv;
            int32_t eq = 0;
            if (nv > 0) {
                const uint32_t x = S.ld32(q) ^ S.ld32(m0 + eb + 4 * (int32_t)lane);
                eq = x ? ((__ffs(x) - 1) >> 3) : 4;
                eq = eq < nv ? eq : nv;
            }
            const uint32_t stop = __ballot_sync(FULL, eq < 4);
            if (stop) {
                const int l = __ffs(stop) - 1;
                ml = eb + 4 * l + __shfl_sync(FULL, eq, l);
                break;
            }
        }
        if (lane == 0) reinterpret_cast<uint2 *>(rec)[nrec] = rec_pack((uint32_t)(s0 - start), (uint32_t)ml, (uint32_t)(s0 - m0));
        ++nrec;
        sIndex = s0 + ml;
    }
    return nrec;
}

// Fresh independent blocks <= 64 KiB: the match finder alone.  rec_base + rec_off[b]: room for len/4 + 1 records of block b.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_parse_fresh16(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, const uint32_t *__restrict__ src_len,
                uint32_t nblocks, uint64_t *__restrict__ rec_base, uint64_t rec_stride /* records per block */,
                uint32_t *__restrict__ nrec_out, uint32_t *counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem) + warp * kHashEntries;
    for (;;) {
        const uint32_t b = next_block(counter, lane);
        if (b >= nblocks) break;
        const uint32_t len = src_len[b];
        if (len > 65536u) { if (lane == 0) nrec_out[b] = 0xFFFFFFFFu; continue; }
        uint4 *t4 = reinterpret_cast<uint4 *>(tab);
        for (uint32_t i = lane; i < kHashEntries * 2 / 16; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        Tab16 T{tab, 0};
        const uint32_t n = parse_block64_warp(src + src_off[b], 0, (int32_t)len, T, rec_base + (uint64_t)b * rec_stride);
        if (lane == 0) nrec_out[b] = n;
        __syncwarp();
    }
}

// Sequences from match records (blockCompress.js:75-174 per match, :179-230 for the final literals).  One warp per block.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_encode_blocks(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, const uint32_t *__restrict__ src_len,
                uint32_t nblocks, const uint64_t *__restrict__ rec_base, uint64_t rec_stride, const uint32_t *__restrict__ nrec_in,
                uint8_t *__restrict__ dst, const uint64_t *__restrict__ dst_off, uint32_t *__restrict__ comp_len, uint32_t *counter) {
    const uint32_t lane = lane_id();
    // the queue is walked twice: blocks with many sequences first (a text block takes a warp about as long as the whole kernel
    // runs), the light ones fill the tail
    const bool twice = nblocks <= 0x7FFFFFFFu;
    const uint32_t qlen = twice ? 2u * nblocks : nblocks;
    for (;;) {
        const uint32_t qi = next_block(counter, lane);
        if (qi >= qlen) break;
        const uint32_t b = qi < nblocks ? qi : qi - nblocks;
        const uint32_t n = nrec_in[b];
        if (twice && (n != 0xFFFFFFFFu && n >= 256u) != (qi < nblocks)) continue;
        if (n == 0xFFFFFFFFu) { if (lane == 0) comp_len[b] = 0xFFFFFFFFu; continue; }
        const uint8_t *in = src + src_off[b];
        const uint32_t len = src_len[b];
        const uint64_t *rec = rec_base + (uint64_t)b * rec_stride;
        uint8_t *out = dst + dst_off[b];
        uint32_t D = 0, prev_end = 0;                            // output offset; end of the previous match (= mAnchor, :174)
        for (uint32_t r0 = 0; r0 < n; r0 += 32) {
            const uint32_t i = r0 + lane;
            const bool have = i < n;
            const uint2 rc = have ? reinterpret_cast<const uint2 *>(rec)[i] : make_uint2(0u, 0u);
            const uint32_t pos = rc.x & 0xFFFFu, ml = rc.y, offset = rc.x >> 16;
            const uint32_t end = pos + ml;
            uint32_t pe = __shfl_up_sync(FULL, end, 1);
            if (lane == 0) pe = prev_end;
            const uint32_t lit = have ? pos - pe : 0u;                                   // :74 litLen
            const uint32_t code = ml - 4u;                                               // :160
            const uint32_t litx = lit >= 15u ? 1u + (lit - 15u) / 255u : 0u;
            const uint32_t mlx = have && code >= 15u ? 1u + (code - 15u) / 255u : 0u;
            const uint32_t size = have ? 3u + litx + lit + mlx : 0u;
            uint32_t incl = size;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
            uint32_t d = D + incl - size;                        // this sequence's token
            if (have) {
                out[d++] = (uint8_t)(((lit < 15u ? lit : 15u) << 4) | (code < 15u ? code : 15u));   // :78-90, :161-170
                if (lit >= 15u) {
                    uint32_t rest = lit - 15u;
                    while (rest >= 255u) { out[d++] = 255; rest -= 255u; }
                    out[d++] = (uint8_t)rest;
                }
            }
            // literals (:92-140): short runs by their own lane, long ones by the whole warp
            const uint32_t lits_at = d;
            const bool longrun = have && lit > 24u;
            if (have && !longrun) {
                const uint8_t *s = in + pe;
                for (uint32_t k = 0; k < lit; ++k) out[d + k] = s[k];
            }
            for (uint32_t m = __ballot_sync(FULL, longrun); m; m &= m - 1u) {
                const int l = __ffs(m) - 1;
                const uint32_t o_ = __shfl_sync(FULL, lits_at, l), p_ = __shfl_sync(FULL, pe, l), n_ = __shfl_sync(FULL, lit, l);
                warp_copy(out + o_, in + p_, n_, lane);
            }
            if (have) {
                d += lit;
                out[d] = (uint8_t)offset;                        // :156-157
                out[d + 1] = (uint8_t)(offset >> 8);
                d += 2;
                if (code >= 15u) {
                    uint32_t rest = code - 15u;
                    while (rest >= 255u) { out[d++] = 255; rest -= 255u; }
                    out[d] = (uint8_t)rest;
                }
            }
            D += __shfl_sync(FULL, incl, 31);
            const uint32_t last = min(n - r0, 32u) - 1u;
            prev_end = __shfl_sync(FULL, end, last);
        }
        // final literals (:179-230)
        __syncwarp();
        uint8_t *e = emit_literals(out + D, SrcFlat{in}, (int32_t)prev_end, len - prev_end, 0u, lane);
        if (lane == 0) comp_len[b] = (uint32_t)(e - out);
        __syncwarp();
    }
}

}  // namespace dlz4
