// dlz4_team.cuh -- the match finder with FOUR warps per block chain (included by dlz4_kernels.cuh after dlz4_parse.cuh).
//
// Why: a block's parse is one dependent chain and its hash table is 32 KiB whatever is done, so an SM holds 7 chains in shared
// memory.  With one warp per chain (k_parse_fresh16) those 7 warps leave 70 % of the issue slots idle -- every step waits for
// two memory round trips (table, candidate bytes) and for its own dependent instructions -- and more chains only fit with
// their tables in L2 (round 1: 17x DRAM traffic).  The way to use the SM with 7 tables is more lanes PER chain:
//
//   * a team of 4 warps owns one block; a step covers 128 consecutive positions, one per lane.  Per position the work is that
//     of dlz4_parse.cuh (source bytes, hash, look-up, candidate load, verification, pre-extension to 32 bytes), done by 128
//     lanes at once; the memory round trips and the dependent arithmetic of a step are paid once per 128 positions;
//   * the serial part -- following the chain of heads -- is ~25 instructions per match on ONE warp, fed by successor links all
//     lanes computed; the other three warps wait at a barrier (their issue slots go to the other six teams of the SM);
//   * same-slot positions inside a window: OPTIMISTIC.  Every position takes the table state from before the window as its
//     candidate; after the parse only the positions the serial loop really probed store their tag (blockCompress.js:55) and
//     read the slot back.  Two PROBED positions in one slot are the only case where a candidate was wrong (the later one must
//     see the earlier one's entry); then the window is cut behind the lowest probed position involved and the step is redone
//     from the links (rare: profiles/scratch/r02/window_stats.c -- 134.8 of 137.2 possible bytes per 128-position window on
//     log text, against 80.9 for the cut-before-parsing rule of dlz4_parse.cuh).  Entries are written by probed positions only,
//     so nothing has to be taken back in the common case.
//
// One CTA per SM holds the 7 teams (28 warps, 7 x 32 KiB of tables); teams synchronise with named barriers (bar.sync id, 128).
// Output: match records as in dlz4_parse.cuh; k_encode_blocks writes the sequences.
#pragma once

namespace dlz4 {

constexpr int kTeamWarps = 4;
constexpr int kTeamsPerCta = 7;
constexpr int kTeamThreads = kTeamWarps * 32;
constexpr int kTeamWin = kTeamWarps * 32;                 // positions per step
struct TeamShared {                                      // per team, behind the tables
    uint32_t hm[kTeamWarps];                             // hit masks of the step
    uint32_t heads[kTeamWarps];                          // heads of the step (after the chain)
    uint32_t carry_end[kTeamWarps];                      // end (window-relative) of the last match that began below warp t's range
    uint32_t conf[kTeamWarps];                           // per warp: some probed position lost its slot
    uint32_t cur, lmin, block, pad;                      // end of the last match; lowest probed position in a shared slot; work item
    uint16_t pack[kTeamWin];                             // per position: pre-extended length | successor << 6, later the final length
};
constexpr int kTeamSmemBytes = kTeamsPerCta * kHashEntries * 2 + kTeamsPerCta * (int)sizeof(TeamShared);

__device__ __forceinline__ void team_sync(uint32_t team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1u), "n"(kTeamThreads) : "memory");
}

// first set bit at or above position e (< 128) of the 128-bit mask m[0..3]; 128 if none
__device__ __forceinline__ uint32_t first_from128(const uint32_t m0, const uint32_t m1, const uint32_t m2, const uint32_t m3, const uint32_t e) {
    if (e >= 128u) return 128u;
    const uint32_t wsel = e >> 5, sh = e & 31u;
    // words at and above the one holding e, the first one masked below e
    const uint32_t a0 = wsel == 0u ? (m0 >> sh) << sh : 0u;
    const uint32_t a1 = wsel == 1u ? (m1 >> sh) << sh : (wsel < 1u ? m1 : 0u);
    const uint32_t a2 = wsel == 2u ? (m2 >> sh) << sh : (wsel < 2u ? m2 : 0u);
    const uint32_t a3 = wsel == 3u ? (m3 >> sh) << sh : m3;
    if (a0) return (uint32_t)__ffs(a0) - 1u;
    if (a1) return 31u + (uint32_t)__ffs(a1);
    if (a2) return 63u + (uint32_t)__ffs(a2);
    if (a3) return 95u + (uint32_t)__ffs(a3);
    return 128u;
}

__global__ void __launch_bounds__(kTeamsPerCta * kTeamThreads, 1)
k_parse_team(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, const uint32_t *__restrict__ src_len,
             uint32_t nblocks, uint64_t *__restrict__ rec_base, uint64_t rec_stride, uint32_t *__restrict__ nrec_out, uint32_t *counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id();
    const uint32_t team = threadIdx.x / kTeamThreads, tw = (threadIdx.x >> 5) & (kTeamWarps - 1), tl = threadIdx.x & (kTeamThreads - 1);
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t *const tab = reinterpret_cast<uint16_t *>(smem) + team * kHashEntries;
    TeamShared &sh = *reinterpret_cast<TeamShared *>(smem + kTeamsPerCta * kHashEntries * 2 + team * sizeof(TeamShared));
    Tab16 T{tab, 0};

    for (;;) {
        if (tl == 0) sh.block = atomicAdd(counter, 1u);
        team_sync(team);
        const uint32_t b = sh.block;
        if (b >= nblocks) break;
        const int32_t len = (int32_t)src_len[b];
        uint2 *const rec = reinterpret_cast<uint2 *>(rec_base + (uint64_t)b * rec_stride);
        if (len > 65536) { if (tl == 0) nrec_out[b] = 0xFFFFFFFFu; team_sync(team); continue; }
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(tab);
            for (uint32_t i = tl; i < kHashEntries * 2 / 16; i += kTeamThreads) t4[i] = make_uint4(0, 0, 0, 0);
        }
        team_sync(team);
        const uint8_t *__restrict__ base = src + src_off[b];
        const SrcFlat S{base};
        const int32_t sEnd = len, mflimit = sEnd - 12, matchLimit = sEnd - 5;   // blockCompress.js:34-35
        int32_t sIndex = 0;
        uint32_t smc = 67, nrec = 0;                                            // :40
        const uint32_t A0 = (uint32_t)(reinterpret_cast<uintptr_t>(base) & 127u);
        const int32_t wlo = -(int32_t)(A0 & 3u), whi = sEnd;
        uint32_t R0 = 0, R1 = 0, R2 = 0, R3 = 0, Lc = 0;     // this warp's source lines (word `lane` of lines Lc .. Lc+3)
        bool cold = true;
        auto load_line = [&](uint32_t l) -> uint32_t {
            const int32_t idx = (int32_t)(l * 128u + 4u * lane) - (int32_t)A0;
            return (idx >= wlo && idx < whi) ? __ldg(reinterpret_cast<const uint32_t *>(base + idx)) : 0u;
        };

        while (sIndex < mflimit) {                                              // :48
            if (smc <= 96u && sIndex + kTeamWin + 36 <= sEnd) {
                // ================= dense window: positions w .. w+127, this lane's is p = w + rel
                const int32_t w = sIndex;
                const uint32_t rel = tw * 32u + lane;
                const int32_t p = w + (int32_t)rel;
                const uint32_t va = A0 + (uint32_t)w + tw * 32u;               // virtual address of this warp's first position
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
                const uint32_t wo = ((va & 127u) >> 2) + lane;
                const uint32_t x0 = __shfl_sync(FULL, R0, wo), x1 = __shfl_sync(FULL, R1, wo);
                const uint32_t Tw = wo < 32u ? x0 : x1;
                const uint32_t o = wmis + lane, wi = o >> 2, shb = (o & 3u) * 8u;
                uint32_t Sw[8];                                                // bytes p .. p+31
                {
                    uint32_t tprev = __shfl_sync(FULL, Tw, wi);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t tnext = __shfl_sync(FULL, Tw, wi + k + 1);
                        Sw[k] = __funnelshift_r(tprev, tnext, shb);
                        tprev = tnext;
                    }
                }
                const uint32_t h = (Sw[0] * 2654435761u) >> 18;                // :53
                const uint32_t old = tab_raw(T, h);                            // :54 (state from before the window)
                const int32_t cand = tab_dec(T, old);
                const bool ok = cand >= 0 && cand != p && (((uint32_t)(p - cand)) >> 16) == 0;   // :62
                const int32_t la = ok ? cand : p;
                const uint32_t cs = (A0 + (uint32_t)la) & 15u;
                const uint4 *cq = reinterpret_cast<const uint4 *>(base + (la - (int32_t)cs));
                const uint4 q0 = __ldg(cq), q1 = __ldg(cq + 1), q2 = __ldg(cq + 2);
                const uint32_t vlen = wide_verify(q0, q1, q2, cs, Sw);          // :63, :147-150 up to 32 bytes
                const uint32_t ml0 = ok ? vlen : 0u;
                const uint32_t tag = tab_enc(T, p);
                const uint32_t hmine = __ballot_sync(FULL, ml0 != 0u);
                if (lane == 0) sh.hm[tw] = hmine;
                team_sync(team);                                               // ---- B1: hit masks
                uint32_t trunc = kTeamWin;                                     // positions [0, trunc) take part
                for (;;) {
                    // (redone after a cut) hit masks below the cut -> successor links -> the chain (warp 0) -> roles -> tags
                    uint32_t m0 = sh.hm[0], m1 = sh.hm[1], m2 = sh.hm[2], m3 = sh.hm[3];
                    if (trunc < (uint32_t)kTeamWin) {
                        const uint32_t tq = trunc >> 5, tr = (1u << (trunc & 31u)) - 1u;
                        m0 &= tq == 0u ? tr : FULL;
                        m1 &= tq == 1u ? tr : (tq > 1u ? FULL : 0u);
                        m2 &= tq == 2u ? tr : (tq > 2u ? FULL : 0u);
                        m3 &= tq == 3u ? tr : 0u;
                    }
                    const uint32_t suc = first_from128(m0, m1, m2, m3, rel + ml0);
                    sh.pack[rel] = (uint16_t)(ml0 | (suc << 6));               // ml: 6 bits, successor: 8 bits
                    team_sync(team);                                           // ---- B2: links
                    if (tw == 0) {
                        // the chain of heads, serial and uniform: first hit inside the dense stretch, then link by link
                        const uint32_t dl0 = 128u - smc;                       // dense probing reaches [0, dl0) before the first match
                        uint32_t hl = first_from128(m0, m1, m2, m3, 0u);
                        if (hl >= dl0) hl = kTeamWin;
                        uint32_t H0 = 0, H1 = 0, H2 = 0, H3 = 0, c1 = 0, c2 = 0, c3 = 0, cur = 0;
                        while (hl < (uint32_t)kTeamWin) {
                            const uint32_t fld = sh.pack[hl];
                            uint32_t mlh = fld & 63u;
                            uint32_t nxt = fld >> 6;
                            if (mlh == 32u && matchLimit - (w + (int32_t)hl) > 32) {
                                // long match: continue cooperatively, 128 bytes per round (:147-150).  The head's candidate is
                                // still in the table (entries of this window are stored after the chain).
                                const int32_t s0 = w + (int32_t)hl;
                                const int32_t mc = tab_dec(T, tab_raw(T, (S.ld32(s0) * 2654435761u) >> 18));
                                for (int32_t eb = 32;; eb += 128) {
                                    const int32_t q = s0 + eb + 4 * (int32_t)lane;
                                    int32_t nv = matchLimit - q;
                                    nv = nv > 4 ? 4 : nv;
                                    int32_t eq = 0;
                                    if (nv > 0) {
                                        const uint32_t x = S.ld32(q) ^ S.ld32(mc + eb + 4 * (int32_t)lane);
                                        eq = x ? ((__ffs(x) - 1) >> 3) : 4;
                                        eq = eq < nv ? eq : nv;
                                    }
                                    const uint32_t stopm = __ballot_sync(FULL, eq < 4);
                                    if (stopm) {
                                        const int l = __ffs(stopm) - 1;
                                        mlh = (uint32_t)(eb + 4 * l + __shfl_sync(FULL, eq, l));
                                        break;
                                    }
                                }
                                nxt = first_from128(m0, m1, m2, m3, hl + mlh);
                            }
                            if (lane == 0) sh.pack[hl] = (uint16_t)mlh;        // the head's final length, for its own lane
                            const uint32_t bit = 1u << (hl & 31u), hq = hl >> 5;
                            H0 |= hq == 0u ? bit : 0u; H1 |= hq == 1u ? bit : 0u; H2 |= hq == 2u ? bit : 0u; H3 |= hq == 3u ? bit : 0u;
                            cur = hl + mlh;
                            // the last match that began below warp t's range ends at c_t
                            if (hl < 32u) c1 = cur;
                            if (hl < 64u) c2 = cur;
                            if (hl < 96u) c3 = cur;
                            hl = nxt;
                        }
                        if (lane == 0) {
                            sh.heads[0] = H0; sh.heads[1] = H1; sh.heads[2] = H2; sh.heads[3] = H3;
                            sh.carry_end[0] = 0; sh.carry_end[1] = c1; sh.carry_end[2] = c2; sh.carry_end[3] = c3;
                            sh.cur = cur;
                            sh.lmin = kTeamWin;
                        }
                    }
                    team_sync(team);                                           // ---- B3: heads
                    const uint32_t H0 = sh.heads[0], H1 = sh.heads[1], H2 = sh.heads[2], H3 = sh.heads[3];
                    const uint32_t Hm = tw == 0u ? H0 : tw == 1u ? H1 : tw == 2u ? H2 : H3;
                    const uint32_t cur = sh.cur;
                    const bool any = (H0 | H1 | H2 | H3) != 0u;
                    // the window ends where dense probing, the cut or the 128 positions end -- or behind the last match
                    const uint32_t stop = any ? trunc : min(trunc, 128u - smc);
                    const bool head = (Hm >> lane) & 1u;
                    const uint32_t mlfin = head ? (uint32_t)sh.pack[rel] : ml0;
                    // inside a match: the nearest head below this position reaches over it
                    const uint32_t below = Hm & lt;
                    const uint32_t myend = rel + mlfin;                        // (meaningful for heads)
                    const uint32_t pe = __shfl_sync(FULL, myend, below ? 31 - __clz(below) : 0);
                    const uint32_t end_below = below ? pe : sh.carry_end[tw];
                    const bool probed = rel < stop && rel >= end_below;        // :55 exactly the probed positions keep an entry
                    // records at the rank of the head's bit
                    if (head) {
                        const uint32_t rank = (tw > 0u ? __popc(H0) : 0u) + (tw > 1u ? __popc(H1) : 0u) + (tw > 2u ? __popc(H2) : 0u) + __popc(below);
                        rec[nrec + rank] = make_uint2((uint32_t)p | ((uint32_t)(p - cand) << 16), mlfin);
                    }
                    const uint32_t nheads = __popc(H0) + __popc(H1) + __popc(H2) + __popc(H3);
                    // probed positions store their entry, then read the slot back: two probed positions in one slot?
                    if (probed) tab_set_raw(T, h, tag);
                    team_sync(team);                                           // ---- B4: entries
                    const bool lost = probed && tab_raw(T, h) != tag;
                    const uint32_t lostm = __ballot_sync(FULL, lost);
                    if (lane == 0) sh.conf[tw] = lostm;
                    team_sync(team);                                           // ---- B5: verdict
                    if ((sh.conf[0] | sh.conf[1] | sh.conf[2] | sh.conf[3]) == 0u) {
                        // common case: the step stands
                        nrec += nheads;
                        if (cur < stop) { smc = (any ? 67u : smc) + (stop - cur); sIndex = w + (int32_t)stop; }   // trailing misses
                        else { smc = 67u; sIndex = w + (int32_t)cur; }                                      // :71
                        break;
                    }
                    // ---- rare: two probed positions share a slot.  Cut the window behind the lowest probed position involved
                    //      (all probed positions up to it have pairwise different slots and no probed slot-mate below them, so
                    //      their candidates were right), take this step's entries back, redo the step from the links.
                    if (lost) {
                        const uint32_t winner = (uint32_t)(tab_dec(T, tab_raw(T, h)) - w);
                        atomicMin(&sh.lmin, min(rel, winner));
                    }
                    team_sync(team);
                    if (probed) tab_set_raw(T, h, old);
                    trunc = sh.lmin + 1u;
                    team_sync(team);
                }
                team_sync(team);                                               // ---- end of step (entries visible, state agreed)
                continue;
            }

            // ================= batch step: sparse schedule and block tail.  Every warp of the team runs it identically (same
            // loads, same table writes, same record): no exchange needed, one barrier keeps the warps in step.
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
            team_sync(team);                                                   // every warp has read the table before any writes it
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
            if (tw == 0 && ((commit >> lane) & 1u) && ((same & commit) >> lane) == 1u) T.put(h, p);
            team_sync(team);
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
                const uint32_t stopm = __ballot_sync(FULL, eq < 4);
                if (stopm) {
                    const int l = __ffs(stopm) - 1;
                    ml = eb + 4 * l + __shfl_sync(FULL, eq, l);
                    break;
                }
            }
            if (tl == 0) rec[nrec] = rec_pack((uint32_t)s0, (uint32_t)ml, (uint32_t)(s0 - m0));
            ++nrec;
            sIndex = s0 + ml;
        }
        if (tl == 0) nrec_out[b] = nrec;
        team_sync(team);
    }
}

}  // namespace dlz4
