// dlz4_wide.cuh -- the 64-position dense window of the compressor (included by dlz4_kernels.cuh).
//
// Same exact parse as compress_span_warp (blockCompress.js:31-233, bit for bit), re-cut so that one warp covers TWICE the
// positions per dependent step and the serial part of a step is short:
//   * a lane owns two positions, w + lane and w + 32 + lane; their table look-ups, candidate loads (3 x 16 bytes each),
//     verification and pre-extension to 32 bytes are independent instruction streams the scheduler overlaps, so the two
//     memory round trips of a step (table, candidate bytes) are paid once per 64 positions instead of once per 32;
//   * same-slot pairs inside the window are found with the table itself instead of match.any: every position stores its own
//     tag into its slot and reads the slot back.  A position that does not read its own tag shares the slot with another
//     window position; the lowest position involved in any such pair, Lmin, bounds the window to [w, w + Lmin] -- those
//     positions have pairwise different slots, hence each one's candidate is the table state from before the window whatever
//     the parse does inside it.  No fallback step, no false negatives, and no false positives either (tags are positions).
//     After the walk the slots of positions the serial loop never probed (inside matches, behind the window's end) get their
//     old value back; the probed ones already hold what blockCompress.js:55 stores;
//   * the source lines around the window live in four registers per lane (the next 512 bytes of the block, loaded with plain
//     coalesced 128-byte loads up to three lines ahead of use) instead of a shared-memory ring: no cp.async wait in the
//     dependent chain, and all of shared memory is left to the tables;
//   * the walk over the heads keeps everything uniform (no per-head shuffles except the head's pre-extended length) and hands
//     every literal lane the output base of its sequence, so the emission is one predicated byte store per position.
// Only tables in shared memory take this path (the tag store / read-back is a shared-memory round trip; on an L2-resident
// table a store to a sector loaded a moment earlier costs thousands of cycles, profiles/r01b_ubench_l2_table_access.txt).
#pragma once

namespace dlz4 {

__device__ __forceinline__ uint64_t mask_lt64(uint32_t n) { return n >= 64u ? ~0ull : ((1ull << n) - 1ull); }

// candidate bytes (three aligned 16-byte granules, `cs` = byte offset of the candidate inside the first) against the
// position's own 32 bytes Sw[0..7]: returns the match length pre-extended to at most 32, 0 when the first four bytes differ
__device__ __forceinline__ uint32_t wide_verify(const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint32_t cs,
                                                const uint32_t *Sw) {
    uint32_t v[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    if (cs & 8u) {
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = v[k + 2];
    }
    if (cs & 4u) {
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = v[k + 1];
    }
    const uint32_t csh = (cs & 3u) * 8u;
    if (__funnelshift_r(v[0], v[1], csh) != Sw[0]) return 0u;
    uint32_t n = 28;
#pragma unroll
    for (int k = 7; k >= 1; --k) {
        const uint32_t x = Sw[k] ^ __funnelshift_r(v[k], v[k + 1], csh);
        if (x) n = 4u * (uint32_t)(k - 1) + ((uint32_t)(__ffs(x) - 1) >> 3);
    }
    return 4u + n;
}

// Runs the single-warp parse of the block [start, start + len) from state `st`; arguments and return value as
// compress_span_warp (limit / kSpanStopped / kEmit), contiguous source only.  Tab must live in shared memory.
template <class Tab, bool kEmit>
__device__ uint32_t compress_span64_warp(const uint8_t *__restrict__ base, const int32_t start, const int32_t len, Tab &T,
                                         uint8_t *const out, SpanState &st, const int32_t limit) {
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    const int32_t sEnd = start + len;
    const int32_t mflimit = sEnd - 12;
    const int32_t matchLimit = sEnd - 5;
    int32_t sIndex = st.sIndex, anchor = st.anchor;
    uint32_t smc = st.smc;
    uint32_t D = st.D;
    uint32_t pend = st.pend;
    const SrcFlat S{base};

    // virtual byte address of index v: A0 + v, with A0 = the base pointer's offset inside its 128-byte line, so that
    // line numbers and alignments follow from 32-bit arithmetic
    const uint32_t A0 = (uint32_t)(reinterpret_cast<uintptr_t>(base) & 127u);
    const int32_t wlo = start - (int32_t)((A0 + (uint32_t)start) & 3u);      // first word holding block bytes
    const int32_t whi = sEnd;                                                  // words starting below sEnd hold block bytes
    uint32_t R0 = 0, R1 = 0, R2 = 0, R3 = 0;      // word `lane` of the lines Lc, Lc+1, Lc+2, Lc+3
    uint32_t Lc = 0;
    bool cold = true;
    auto load_line = [&](uint32_t l) -> uint32_t {
        const int32_t idx = (int32_t)(l * 128u + 4u * lane) - (int32_t)A0;
        return (idx >= wlo && idx < whi) ? __ldg(reinterpret_cast<const uint32_t *>(base + idx)) : 0u;
    };

    while (sIndex < mflimit) {
        if (smc <= 96u && sIndex + 100 <= sEnd && sIndex + 64 <= limit) {
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
            // ---- source bytes: word (wa/4 + lane) of the window, then each lane's 64 bytes from 17 shuffles
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
            const uint32_t ha = (Sx[0] * 2654435761u) >> 18, hb = (Sx[8] * 2654435761u) >> 18;
            // ---- table look-up (the state from before the window) and candidate loads
            const uint32_t olda = tab_raw(T, ha), oldb = tab_raw(T, hb);
            const int32_t canda = tab_dec(T, olda), candb = tab_dec(T, oldb);
            const bool oka = canda >= 0 && canda != pa && (((uint32_t)(pa - canda)) >> 16) == 0;
            const bool okb = candb >= 0 && candb != pb && (((uint32_t)(pb - candb)) >> 16) == 0;
            uint4 qa0 = make_uint4(0, 0, 0, 0), qa1 = qa0, qa2 = qa0, qb0 = qa0, qb1 = qa0, qb2 = qa0;
            uint32_t csa = 0, csb = 0;
            if (oka) {            // canda + 35 < pa + 35 <= w + 98 < sEnd; the 16-byte granules holding them are read whole
                csa = (A0 + (uint32_t)canda) & 15u;
                const uint4 *cq = reinterpret_cast<const uint4 *>(base + (canda - (int32_t)csa));
                qa0 = __ldg(cq); qa1 = __ldg(cq + 1);
                if (csa + 36u > 32u) qa2 = __ldg(cq + 2);
            }
            if (okb) {
                csb = (A0 + (uint32_t)candb) & 15u;
                const uint4 *cq = reinterpret_cast<const uint4 *>(base + (candb - (int32_t)csb));
                qb0 = __ldg(cq); qb1 = __ldg(cq + 1);
                if (csb + 36u > 32u) qb2 = __ldg(cq + 2);
            }
            // ---- same-slot pairs inside the window: tag every slot, read it back
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
                // positions behind the cut leave the window: their slots get the old value back now (before the included
                // positions store theirs: an included and an excluded position may share a slot)
                __syncwarp();
                if (lane >= trunc) tab_set_raw(T, ha, olda);
                if (lane + 32u >= trunc) tab_set_raw(T, hb, oldb);
                __syncwarp();
            }
            // ---- verify + pre-extend
            uint32_t mla = oka ? wide_verify(qa0, qa1, qa2, csa, Sx) : 0u;
            uint32_t mlb = okb ? wide_verify(qb0, qb1, qb2, csb, Sx + 8) : 0u;
            const uint64_t HM = ((uint64_t)__ballot_sync(FULL, mlb != 0u) << 32) | (uint64_t)__ballot_sync(FULL, mla != 0u);
            const uint32_t mlpack = mla | (mlb << 8);
            // ---- walk (uniform): first hit at or behind `cur`, jump behind its match
            const int32_t a_rel0 = anchor - w;                                 // <= 0: literals pending from earlier windows
            int32_t a_rel = a_rel0;
            uint32_t cur = 0, dl = 128u - smc;                                 // dense probing reaches [cur, dl)
            uint64_t heads = 0, lits = 0;
            uint32_t lit0 = 0, D0 = 0;                                         // first head of the window (re-copy of pending literals)
            uint32_t myDa = 0, myLita = 0, myMla = 0, myDb = 0, myLitb = 0, myMlb = 0;
            uint32_t lba = 0, lbb = 0;                                         // out[lb + rel] is this position's literal byte
            for (;;) {
                const uint32_t bound = min(trunc, dl);
                if (cur >= bound) break;
                const uint64_t m = (HM >> cur << cur) & mask_lt64(bound);
                if (!m) break;
                const int hl = __ffsll((long long)m) - 1;
                const uint32_t pk = __shfl_sync(FULL, mlpack, hl);
                int32_t mlh = (int32_t)((hl & 32) ? (pk >> 8) : (pk & 255u));
                const int32_t s0 = w + hl;
                if (mlh == 32 && matchLimit - s0 > 32) {
                    // long match: continue cooperatively, 128 bytes per round
                    const int32_t m0 = __shfl_sync(FULL, (hl & 32) ? candb : canda, hl);
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
                            mlh = eb + 4 * l + __shfl_sync(FULL, eq, l);
                            break;
                        }
                    }
                }
                const uint32_t lit = (uint32_t)(hl - a_rel);
                const uint32_t code = (uint32_t)(mlh - 4);
                uint32_t litx = 0, mlx = code >= 15u;
                if (lit >= 15u) litx = 1u + (lit - 15u) / 255u;
                if (code >= 15u + 255u) mlx = 1u + (code - 15u) / 255u;
                if (!heads) { lit0 = lit; D0 = D; }
                const int32_t lo = a_rel > 0 ? a_rel : 0;                      // literal lanes of this sequence: [lo, hl)
                const uint32_t lb = D + 1u + litx - (uint32_t)a_rel;
                if ((int32_t)lane >= lo && (int32_t)lane < hl) lba = lb;
                if ((int32_t)lane + 32 >= lo && (int32_t)lane + 32 < hl) lbb = lb;
                lits |= mask_lt64((uint32_t)hl) & ~mask_lt64((uint32_t)lo);
                if (lane == (uint32_t)(hl & 31)) {
                    if (hl & 32) { myDb = D; myLitb = lit; myMlb = (uint32_t)mlh; }
                    else { myDa = D; myLita = lit; myMla = (uint32_t)mlh; }
                }
                D += 3u + litx + lit + mlx;
                heads |= 1ull << hl;
                a_rel = hl + mlh;
                cur = (uint32_t)a_rel;
                dl = cur + 61u;
            }
            // the window ends where dense probing, the cut or the 64 positions end -- or behind the last match
            const uint32_t stop = min(trunc, dl);                              // <= 64
            uint32_t next_rel, tail = 0;
            if (cur < stop) {
                // trailing probed positions [cur, stop): literals of the still-open sequence, stored provisionally
                tail = stop - cur;
                const uint32_t pend0 = heads ? 0u : pend;
                const uint32_t lb = D + 1u + pend0 - cur;
                if (lane >= cur && lane < stop) lba = lb;
                if (lane + 32u >= cur && lane + 32u < stop) lbb = lb;
                lits |= mask_lt64(stop) & ~mask_lt64(cur);
                next_rel = stop;
            } else {
                next_rel = cur;
            }
            const uint64_t probed = lits | heads;
            const bool pra = (probed >> lane) & 1ull, prb = (probed >> (lane + 32u)) & 1ull;
            // ---- table: exactly the probed positions keep their entry (blockCompress.js:55)
            if (conf) {
                if (lane < trunc) tab_set_raw(T, ha, pra ? taga : olda);
                if (lane + 32u < trunc) tab_set_raw(T, hb, prb ? tagb : oldb);
            } else {
                if (!pra) tab_set_raw(T, ha, olda);
                if (!prb) tab_set_raw(T, hb, oldb);
            }
            // ---- emission
            if (kEmit) {
                if (heads && a_rel0 < 0 && lit0 >= 15u) {
                    // the open run reached 15 literals: its provisional bytes sit one length field too low -> re-copy them
                    warp_copy(out + D0 + 2u + (lit0 - 15u) / 255u, base + anchor, (uint32_t)(-a_rel0), lane);
                }
                if ((lits >> lane) & 1ull) out[lba + lane] = (uint8_t)Sx[0];
                if ((lits >> (lane + 32u)) & 1ull) out[lbb + lane + 32u] = (uint8_t)Sx[8];
                if ((heads >> lane) & 1ull) {
                    uint8_t *q = out + myDa;
                    const uint32_t code = myMla - 4u;
                    q[0] = (uint8_t)(((myLita < 15u ? myLita : 15u) << 4) | (code < 15u ? code : 15u));
                    q += 1;
                    if (myLita >= 15u) {
                        uint32_t rest = myLita - 15u;
                        while (rest >= 255u) { *q++ = 255; rest -= 255u; }
                        *q++ = (uint8_t)rest;
                    }
                    q += myLita;
                    const uint32_t offset = (uint32_t)(pa - canda);
                    q[0] = (uint8_t)offset;
                    q[1] = (uint8_t)(offset >> 8);
                    if (code >= 15u) {
                        uint32_t rest = code - 15u;
                        q += 2;
                        while (rest >= 255u) { *q++ = 255; rest -= 255u; }
                        *q = (uint8_t)rest;
                    }
                }
                if ((heads >> (lane + 32u)) & 1ull) {
                    uint8_t *q = out + myDb;
                    const uint32_t code = myMlb - 4u;
                    q[0] = (uint8_t)(((myLitb < 15u ? myLitb : 15u) << 4) | (code < 15u ? code : 15u));
                    q += 1;
                    if (myLitb >= 15u) {
                        uint32_t rest = myLitb - 15u;
                        while (rest >= 255u) { *q++ = 255; rest -= 255u; }
                        *q++ = (uint8_t)rest;
                    }
                    q += myLitb;
                    const uint32_t offset = (uint32_t)(pb - candb);
                    q[0] = (uint8_t)offset;
                    q[1] = (uint8_t)(offset >> 8);
                    if (code >= 15u) {
                        uint32_t rest = code - 15u;
                        q += 2;
                        while (rest >= 255u) { *q++ = 255; rest -= 255u; }
                        *q = (uint8_t)rest;
                    }
                }
            }
            if (heads) {
                anchor = w + a_rel;
                pend = tail;
                smc = 67u + tail;
            } else {
                pend += tail;
                smc += tail;
            }
            sIndex = w + (int32_t)next_rel;
            __syncwarp();
            continue;
        }

        // ---- batch step (identical to compress_block_warp's loop body): sparse schedule, block tail, segment end
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
        if (kEmit) {
            const uint32_t code = (uint32_t)(ml - 4);
            uint8_t *d = emit_literals(out + D, S, anchor, (uint32_t)(s0 - anchor), code < 15u ? code : 15u, lane);
            const uint32_t offset = (uint32_t)(s0 - m0);
            if (lane == 0) { d[0] = (uint8_t)offset; d[1] = (uint8_t)(offset >> 8); }
            d += 2;
            if (code >= 15u) {
                const uint32_t rest = code - 15u, n255 = rest / 255u;
                for (uint32_t i = lane; i < n255; i += 32) d[i] = 255;
                if (lane == 0) d[n255] = (uint8_t)(rest - n255 * 255u);
                d += n255 + 1;
            }
            D = (uint32_t)(d - out);
        }
        pend = 0;
        sIndex = anchor = s0 + ml;
        if (s0 >= limit) {                   // first sequence that starts in the next segment: hand the state over
            st.sIndex = sIndex; st.anchor = anchor; st.smc = smc; st.D = D; st.pend = 0; st.head = s0;
            return kSpanStopped;
        }
    }
    if (!kEmit) return 0u;
    uint8_t *d = emit_literals(out + D, S, anchor, (uint32_t)(sEnd - anchor), 0u, lane);
    return (uint32_t)(d - out);
}

}  // namespace dlz4
