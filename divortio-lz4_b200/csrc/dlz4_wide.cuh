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

// candidate bytes (three aligned 16-byte granules, `cs` = byte offset of the candidate inside the first) against the
// position's own 32 bytes Sw[0..7]: returns the match length pre-extended to at most 32, 0 when the first four bytes differ.
// Branch-free on purpose (selects only): the two positions of a lane are independent instruction streams, and only
// straight-line code lets the scheduler interleave them.
__device__ __forceinline__ uint32_t wide_verify(const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint32_t cs,
                                                const uint32_t *Sw) {
    uint32_t v[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    const bool s8 = (cs & 8u) != 0, s4 = (cs & 4u) != 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = s8 ? v[k + 2] : v[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = s4 ? v[k + 1] : v[k];
    const uint32_t csh = (cs & 3u) * 8u;
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = Sw[k] ^ __funnelshift_r(v[k], v[k + 1], csh);
    // first differing byte among bytes 4..31 (words x[1..7]; a sentinel behind them stands for "all 28 equal"): halving
    // with selects, one find-first-set at the end
    const bool h4 = (x[1] | x[2] | x[3] | x[4]) == 0u;
    const uint32_t y1 = h4 ? x[5] : x[1], y2 = h4 ? x[6] : x[2], y3 = h4 ? x[7] : x[3], y4 = h4 ? 1u : x[4];
    const bool h2 = (y1 | y2) == 0u;
    const uint32_t z1 = h2 ? y3 : y1, z2 = h2 ? y4 : y2;
    const bool h1 = z1 == 0u;
    const uint32_t zz = h1 ? z2 : z1;
    const uint32_t n = (h4 ? 16u : 0u) + (h2 ? 8u : 0u) + (h1 ? 4u : 0u) + ((uint32_t)(__ffs(zz) - 1) >> 3);
    return x[0] == 0u ? 4u + n : 0u;
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
            // candidate bytes: three aligned 16-byte granules each (cand + 35 < p + 35 <= w + 98 < sEnd; the granules holding
            // them are read whole).  A position without a usable candidate reads at its own address instead -- always valid,
            // already in L1 -- so that the loads and the verification below are straight-line code for both positions.
            const int32_t la = oka ? canda : pa, lb_ = okb ? candb : pb;
            const uint32_t csa = (A0 + (uint32_t)la) & 15u, csb = (A0 + (uint32_t)lb_) & 15u;
            const uint4 *cqa = reinterpret_cast<const uint4 *>(base + (la - (int32_t)csa));
            const uint4 *cqb = reinterpret_cast<const uint4 *>(base + (lb_ - (int32_t)csb));
            const uint4 qa0 = __ldg(cqa), qa1 = __ldg(cqa + 1), qa2 = __ldg(cqa + 2);
            const uint4 qb0 = __ldg(cqb), qb1 = __ldg(cqb + 1), qb2 = __ldg(cqb + 2);
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
            const uint32_t va_ = wide_verify(qa0, qa1, qa2, csa, Sx), vb_ = wide_verify(qb0, qb1, qb2, csb, Sx + 8);
            const uint32_t mla = oka ? va_ : 0u, mlb = okb ? vb_ : 0u;
            // hits that take part: positions below the cut
            uint32_t HMa = __ballot_sync(FULL, mla != 0u), HMb = __ballot_sync(FULL, mlb != 0u);
            if (trunc < 64u) {
                HMa &= trunc >= 32u ? FULL : ((1u << trunc) - 1u);
                HMb &= trunc >= 32u ? ((1u << (trunc - 32u)) - 1u) : 0u;          // trunc - 32 < 32 here
            }
            // first hit at or behind relative position e (64: none), from the two 32-bit masks
            auto first_hit_from = [&](uint32_t e) -> uint32_t {
                const uint32_t ma = e < 32u ? (HMa >> e) << e : 0u;
                const uint32_t mb = e < 32u ? HMb : (e < 64u ? (HMb >> (e - 32u)) << (e - 32u) : 0u);
                return ma ? (uint32_t)__ffs(ma) - 1u : (mb ? 31u + (uint32_t)__ffs(mb) : 64u);
            };
            // every hit position's successor in the chain of heads (the first hit at or behind its match's end), computed by
            // all lanes at once; the serial part below only follows these links.  A pre-extended length of 32 may be a longer
            // match: its successor is computed in the chain once the length is known.
            const uint32_t suca = first_hit_from(lane + mla), sucb = first_hit_from(lane + 32u + mlb);
            const uint32_t pack = (mla | (suca << 6)) | ((mlb | (sucb << 6)) << 16);     // ml: 6 bits, successor: 7 bits, per half
            // ---- walk (uniform): follow the links from the first hit inside the dense stretch
            const int32_t a_rel0 = anchor - w;                                 // <= 0: literals pending from earlier windows
            int32_t a_rel = a_rel0;
            const uint32_t dl0 = 128u - smc;                                   // dense probing reaches [0, dl0) before the first match
            uint32_t hl = first_hit_from(0u);
            if (hl >= dl0) hl = 64u;                                           // (dl0 >= 32, trunc already applied)
            const bool any_head = hl < 64u;
            const uint32_t D0 = D, lit0 = (uint32_t)((int32_t)hl - a_rel0);     // first head of the window (re-copy of pending literals)
            constexpr uint32_t kNone = 0xFFFFFFFFu;
            // out[lb + rel] is where a probed position of a closed sequence stores: a literal lane its own byte, the head lane
            // (the one position of the range that has a match) the two offset bytes that follow the literals
            uint32_t lba = kNone, lbb = kNone;
            uint32_t cur = 0;                                                  // next probe position behind the last match
            while (hl < 64u) {
                const uint32_t pk = __shfl_sync(FULL, pack, hl);
                const uint32_t fld = (pk >> ((hl >> 1) & 16u)) & 0xFFFFu;
                int32_t mlh = (int32_t)(fld & 63u);
                uint32_t nxt = fld >> 6;
                if (mlh == 32 && matchLimit - (w + (int32_t)hl) > 32) {
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
                            mlh = eb + 4 * l + __shfl_sync(FULL, eq, l);
                            break;
                        }
                    }
                    nxt = first_hit_from(hl + (uint32_t)mlh);
                }
                // the sequence: token, [literal length], literals, offset, [match length] (blockCompress.js:75-90, :153-170).
                // Everything uniform is stored by lane 0 right here; literals and offset by their own lanes after the walk.
                const uint32_t lit = (uint32_t)((int32_t)hl - a_rel);
                const uint32_t code = (uint32_t)(mlh - 4);
                uint32_t litx = 0, mlx = 0;
                if (kEmit && lane == 0) out[D] = (uint8_t)(((lit < 15u ? lit : 15u) << 4) | (code < 15u ? code : 15u));
                if (lit >= 15u) {                                              // uniform and rare (3 % of sequences on text)
                    const uint32_t rest = lit - 15u, n255 = rest / 255u;
                    if (kEmit) {
                        for (uint32_t i = lane; i < n255; i += 32) out[D + 1u + i] = 255;
                        if (lane == 0) out[D + 1u + n255] = (uint8_t)(rest - n255 * 255u);
                    }
                    litx = n255 + 1u;
                }
                if (code >= 15u) {
                    const uint32_t rest = code - 15u, n255 = rest / 255u;      // n255 > 0 only for a long match (continued above)
                    if (kEmit) {
                        uint8_t *const d = out + D + 3u + litx + lit;
                        for (uint32_t i = lane; i < n255; i += 32) d[i] = 255;
                        if (lane == 0) d[n255] = (uint8_t)(rest - n255 * 255u);
                    }
                    mlx = n255 + 1u;
                }
                // probed positions of this sequence: the literal lanes [max(a_rel, 0), hl) and the head hl itself
                const uint32_t lo = a_rel > 0 ? (uint32_t)a_rel : 0u;
                const uint32_t lb = D + 1u + litx - (uint32_t)a_rel;
                if (lane - lo <= hl - lo) lba = lb;
                if (lane + 32u - lo <= hl - lo) lbb = lb;
                D += 3u + litx + lit + mlx;
                a_rel = (int32_t)hl + mlh;
                hl = nxt;                                                      // (behind a match dense probing reaches 61 positions: past the window)
            }
            if (any_head) cur = (uint32_t)a_rel;
            // the window ends where dense probing, the cut or the 64 positions end -- or behind the last match
            const uint32_t stop = any_head ? trunc : min(trunc, dl0);         // <= 64
            uint32_t next_rel, tail = 0;
            if (cur < stop) {
                // trailing probed positions [cur, stop): literals of the still-open sequence, stored provisionally
                tail = stop - cur;
                const uint32_t pend0 = any_head ? 0u : pend;
                const uint32_t lb = D + 1u + pend0 - cur;
                if (lane - cur < tail) lba = lb;
                if (lane + 32u - cur < tail) lbb = lb;
                next_rel = stop;
            } else {
                next_rel = cur;
            }
            const bool pra = lba != kNone, prb = lbb != kNone;
            // ---- table: exactly the probed positions keep their entry (blockCompress.js:55)
            if (conf) {
                if (lane < trunc) tab_set_raw(T, ha, pra ? taga : olda);
                if (lane + 32u < trunc) tab_set_raw(T, hb, prb ? tagb : oldb);
            } else {
                if (!pra) tab_set_raw(T, ha, olda);
                if (!prb) tab_set_raw(T, hb, oldb);
            }
            // ---- emission by the probed positions: a position with a match is the head of its sequence (the first hit at or
            //      behind the previous match's end), any other one a literal
            if (kEmit) {
                if (any_head && a_rel0 < 0 && lit0 >= 15u) {
                    // the open run reached 15 literals: its provisional bytes sit one length field too low -> re-copy them
                    warp_copy(out + D0 + 2u + (lit0 - 15u) / 255u, base + anchor, (uint32_t)(-a_rel0), lane);
                }
                if (pra) {
                    uint8_t *const q = out + lba + lane;
                    if (mla) { const uint32_t offset = (uint32_t)(pa - canda); q[0] = (uint8_t)offset; q[1] = (uint8_t)(offset >> 8); }
                    else q[0] = (uint8_t)Sx[0];
                }
                if (prb) {
                    uint8_t *const q = out + lbb + lane + 32u;
                    if (mlb) { const uint32_t offset = (uint32_t)(pb - candb); q[0] = (uint8_t)offset; q[1] = (uint8_t)(offset >> 8); }
                    else q[0] = (uint8_t)Sx[8];
                }
            }
            const bool heads = any_head;
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
