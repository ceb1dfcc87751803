// dlz4_wide.cuh -- 32-byte candidate verification shared by the match finders (dlz4_parse.cuh: one warp per chain, two
// positions per lane; dlz4_pw.cuh: the producers).  Included by dlz4_kernels.cuh.
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

}  // namespace dlz4
