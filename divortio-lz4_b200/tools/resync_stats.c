/* resync_stats.c -- does the reference's greedy parse (blockCompress.js:31-233) forget its past?
 * Parse a large block exactly (table carried from position 0), and again from a late start Q with an EMPTY table.
 * At Q + W compare the two parser states: next probe position, anchor, searchMatchCount and every table entry that can
 * still be used (entries older than 65535 bytes are rejected by :62 forever, so they count as empty).  If the states are
 * equal the two parses are identical from there on -- which is what a speculative segment-parallel compressor needs.
 * build: gcc -O2 -o /tmp/resync_stats resync_stats.c corpus.c */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void corpus_log(uint64_t seed, uint8_t *out, uint64_t n);
void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n);
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
typedef struct { int32_t s, anchor, smc; int32_t table[16384]; } state_t;
/* run the parse from state st until the probe position reaches `stop` (a probe AT or beyond stop is not executed) */
static void run(const uint8_t *src, int32_t len, state_t *st, int32_t stop) {
    const int32_t mfl = len - 12, mlim = len - 5;
    int32_t s = st->s, anchor = st->anchor, smc = st->smc;
    while (s < mfl && s < stop) {
        uint32_t seq = rd32(src + s), h = (seq * 2654435761u) >> 18;
        int32_t m = st->table[h] - 1; st->table[h] = s + 1;
        if (m < 0 || m == s || ((uint32_t)(s - m) >> 16) || rd32(src + m) != seq) { s += (smc++ >> 6); continue; }
        smc = 67;
        int32_t sp = s + 4, mp = m + 4;
        while (sp < mlim && src[sp] == src[mp]) { ++sp; ++mp; }
        s = anchor = sp;
    }
    st->s = s; st->anchor = anchor; st->smc = smc;
}
static int same(const state_t *a, const state_t *b, int32_t *ndiff) {
    int d = 0;
    const int32_t horizon = a->s - 65535;
    for (int i = 0; i < 16384; ++i) {
        int32_t x = a->table[i] - 1, y = b->table[i] - 1;
        if (x < horizon) x = -1;
        if (y < horizon) y = -1;
        d += x != y;
    }
    *ndiff = d;
    return d == 0 && a->s == b->s && a->anchor == b->anchor && a->smc == b->smc;
}
int main(int argc, char **argv) {
    const char *kind = argc > 1 ? argv[1] : "log";
    const int32_t len = (argc > 2 ? atoi(argv[2]) : 4) << 20;
    uint8_t *b = malloc((size_t)len + 64);
    static state_t truth, spec;
    const int32_t W[] = {65536, 98304, 131072, 196608, 262144, 393216, 524288};
    int ok[7] = {0}, trials = 0;
    for (int seed = 1; seed <= 4; ++seed) {
        if (!strcmp(kind, "log")) corpus_log(seed, b, len); else corpus_mixed(seed, b, len);
        for (int32_t Q = 256 << 10; Q + (576 << 10) < len; Q += 448 << 10) {
            ++trials;
            memset(&spec, 0, sizeof spec); spec.s = spec.anchor = Q; spec.smc = 67;
            memset(&truth, 0, sizeof truth); truth.smc = 67;
            for (int k = 0; k < 7; ++k) {
                /* advance both to the first probe position >= Q + W[k]; they can only be equal if they stop at the same probe */
                run(b, len, &truth, Q + W[k]);
                run(b, len, &spec, Q + W[k]);
                int32_t nd; const int eq = same(&truth, &spec, &nd);
                ok[k] += eq;
                if (seed == 1 && Q == (256 << 10)) printf("  Q=%d W=%7d: truth s=%d smc=%d | spec s=%d smc=%d | %d usable table entries differ\n", Q, W[k], truth.s, truth.smc, spec.s, spec.smc, nd);
            }
        }
    }
    printf("%s, %d MiB blocks, %d trials: state equal after warm-up of", kind, len >> 20, trials);
    for (int k = 0; k < 7; ++k) printf("  %dK: %d", W[k] >> 10, ok[k]);
    printf("\n");
    return 0;
}
