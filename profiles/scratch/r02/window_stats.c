/* window_stats.c -- CPU model of the dense-window schedule (dlz4_parse.cuh) on top of the exact serial parse: how many bytes a
 * window of W positions advances under three same-slot policies.  Design aid for the team (multi-warp) window; not product code.
 *   policy 0: cut at (lowest position involved in any same-slot pair) + 1            (what dlz4_parse.cuh does)
 *   policy 1: cut at the LATER member of the first pair whose members are both probed  (optimistic, needs the parse first)
 *   policy 2: no cut (upper bound)
 * build: gcc -O2 -o /tmp/window_stats window_stats.c ../../../divortio-lz4_b200/tools/corpus.c */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void corpus_log(uint64_t seed, uint8_t *out, uint64_t n);
void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n);
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint8_t probed[65536 + 256]; static int32_t mlen[65536 + 256];
static void parse(const uint8_t *src, int32_t len) {
    static int32_t table[16384];
    memset(table, 0, sizeof table); memset(probed, 0, sizeof probed); memset(mlen, 0, sizeof mlen);
    int32_t s = 0, smc = 67; const int32_t mfl = len - 12, mlim = len - 5;
    while (s < mfl) {
        uint32_t seq = rd32(src + s), h = (seq * 2654435761u) >> 18;
        int32_t m = table[h] - 1; table[h] = s + 1; probed[s] = 1;
        int ok = !(m < 0 || m == s || ((uint32_t)(s - m) >> 16));
        if (!ok || rd32(src + m) != seq) { s += (smc++ >> 6); continue; }
        smc = 67;
        int32_t sp = s + 4, mp = m + 4;
        while (sp < mlim && src[sp] == src[mp]) { ++sp; ++mp; }
        mlen[s] = sp - s; s = sp;
    }
}
int main(int argc, char **argv) {
    const char *kind = argc > 1 ? argv[1] : "log"; uint64_t n = (argc > 2 ? atoll(argv[2]) : 64) << 20;
    uint8_t *b = malloc(n + 64);
    if (!strcmp(kind, "log")) corpus_log(3, b, n); else corpus_mixed(2, b, n);
    for (int W = 32; W <= 256; W *= 2) for (int pol = 0; pol < 3; ++pol) {
        uint64_t windows = 0, bytes = 0, heads = 0;
        for (uint64_t o = 0; o + 65536 <= n; o += 65536) {
            const uint8_t *src = b + o; parse(src, 65536);
            int32_t w = 0;
            while (w + W + 36 <= 65536) {
                /* skip sparse stretches: windows only start at probed positions (approximation: advance to the next probed one) */
                while (w < 65536 - W - 36 && !probed[w]) ++w;
                int cut = W;
                if (pol < 2) {
                    static int32_t last[16384]; static int32_t stamp[16384]; static int32_t gen = 0; ++gen;
                    for (int k = 0; k < W; ++k) {
                        uint32_t h = (rd32(src + w + k) * 2654435761u) >> 18;
                        if (stamp[h] == gen) {
                            int x = last[h];
                            if (pol == 0) { if (x + 1 < cut) cut = x + 1; }
                            else if (probed[w + x] && probed[w + k]) { if (k < cut) cut = k; }
                            if (pol == 1 && probed[w + k]) last[h] = k;   /* keep the latest probed member */
                        } else { stamp[h] = gen; last[h] = k; }
                    }
                }
                /* advance: walk the real parse inside [w, w+cut) */
                int32_t p = w, end = w + cut;
                while (p < end) { if (mlen[p]) { ++heads; p += mlen[p]; } else ++p; }
                ++windows; bytes += p - w; w = p;
            }
        }
        printf("%s W=%3d policy %d: %.1f bytes/window, %.2f heads/window\n", kind, W, pol, (double)bytes / windows, (double)heads / windows);
    }
    return 0;
}
