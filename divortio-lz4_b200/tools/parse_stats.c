/* parse_stats.c -- statistics of the reference's greedy parse (blockCompress.js:31-233) on the synthetic corpora:
 * candidate distances, match lengths, literal runs, probes per sequence.  Design aid for the CUDA compressor
 * (how much history a shared-memory window must hold, how far to pre-extend a match).  Not part of the product.
 * build: gcc -O2 -o /tmp/parse_stats parse_stats.c corpus.c */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void corpus_log(uint64_t seed, uint8_t *out, uint64_t n);
void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n);
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t dist_ok[20], dist_hit[20], mlh[12], lith[8], probes, okc, hits, seqs, bytes, in_window_pair[4];
static int lg(uint32_t d) { int k = 0; while ((1u << k) < d) ++k; return k; }
static void block(const uint8_t *src, int32_t len) {
    static int32_t table[16384];
    memset(table, 0, sizeof table);
    int32_t s = 0, anchor = 0, smc = 67; const int32_t mfl = len - 12, mlim = len - 5;
    while (s < mfl) {
        uint32_t seq = rd32(src + s), h = (seq * 2654435761u) >> 18;
        int32_t m = table[h] - 1; table[h] = s + 1; ++probes;
        int ok = !(m < 0 || m == s || ((uint32_t)(s - m) >> 16));
        if (ok) { ++okc; ++dist_ok[lg(s - m)]; }
        if (!ok || rd32(src + m) != seq) { s += (smc++ >> 6); continue; }
        smc = 67; ++hits; ++dist_hit[lg(s - m)];
        int32_t lit = s - anchor; ++lith[lit == 0 ? 0 : lit < 4 ? 1 : lit < 8 ? 2 : lit < 15 ? 3 : lit < 32 ? 4 : 5];
        int32_t sp = s + 4, mp = m + 4;
        while (sp < mlim && src[sp] == src[mp]) { ++sp; ++mp; }
        int32_t ml = sp - s; ++mlh[ml <= 7 ? 0 : ml <= 11 ? 1 : ml <= 15 ? 2 : ml <= 19 ? 3 : ml <= 27 ? 4 : ml <= 35 ? 5 : ml <= 67 ? 6 : ml < 256 ? 7 : 8];
        ++seqs; s = anchor = sp;
    }
    bytes += len;
}
int main(int argc, char **argv) {
    const char *kind = argc > 1 ? argv[1] : "log"; uint64_t n = (argc > 2 ? atoll(argv[2]) : 64) << 20; int32_t bs = argc > 3 ? atoi(argv[3]) : 65536;
    uint8_t *b = malloc(n + 64);
    if (!strcmp(kind, "log")) corpus_log(1, b, n); else corpus_mixed(2, b, n);
    for (uint64_t o = 0; o < n; o += bs) block(b + o, (int32_t)(n - o < (uint64_t)bs ? n - o : bs));
    printf("%s %llu MiB blocks of %d: %.1f seq/KiB, %.2f probes/seq, %.1f bytes/seq, ok-cand %.1f%% of probes, hits %.1f%% of probes\n", kind,
           (unsigned long long)(n >> 20), bs, seqs * 1024.0 / bytes, (double)probes / seqs, (double)bytes / seqs, 100.0 * okc / probes, 100.0 * hits / probes);
    printf("distance <=2^k : ok-candidates cumulative %% | hits cumulative %%\n");
    uint64_t a = 0, c = 0;
    for (int k = 0; k <= 16; ++k) { a += dist_ok[k]; c += dist_hit[k]; if (k >= 6) printf("  2^%-2d  %6.2f  %6.2f\n", k, 100.0 * a / okc, 100.0 * c / hits); }
    const char *mln[] = {"4-7", "8-11", "12-15", "16-19", "20-27", "28-35", "36-67", "68-255", "256+"};
    printf("match length: "); for (int i = 0; i < 9; ++i) printf("%s %.1f%%  ", mln[i], 100.0 * mlh[i] / seqs); printf("\n");
    const char *ln[] = {"0", "1-3", "4-7", "8-14", "15-31", "32+"};
    printf("literal run : "); for (int i = 0; i < 6; ++i) printf("%s %.1f%%  ", ln[i], 100.0 * lith[i] / seqs); printf("\n");
    return 0;
}
