/* pw_stats.c -- CPU model of the producer/walker pipeline (dlz4_pw.cuh) on top of the exact serial parse.
 * Producer warps run `lag` positions ahead of the walker and record, per position, the table entry they saw and the match
 * length against it; the walker steps over up to 32 upcoming probe positions, cut at the first same-slot pair, the first
 * stale entry or the first hit.  Counts walker steps per sequence and how often an entry is stale.  Design aid, not product.
 * build: gcc -O2 -o /tmp/pw_stats pw_stats.c ../../../divortio-lz4_b200/tools/corpus.c */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void corpus_log(uint64_t seed, uint8_t *out, uint64_t n);
void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n);
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint32_t skip_sum(uint32_t c) { /* sum_{k<c} (k >> 6) */ uint32_t q = c >> 6, r = c & 63; return 64u * (q * (q - 1) / 2) + r * q; }
int main(int argc, char **argv) {
    const char *kind = argc > 1 ? argv[1] : "log"; uint64_t n = (argc > 2 ? atoll(argv[2]) : 64) << 20;
    uint8_t *b = malloc(n + 64);
    if (!strcmp(kind, "log")) corpus_log(3, b, n); else corpus_mixed(2, b, n);
    for (int lag = 32; lag <= 512; lag *= 2) {
        uint64_t steps = 0, seqs = 0, probes = 0, stale = 0, cuts = 0, nohit = 0, longm = 0, stale_hit = 0;
        for (uint64_t o = 0; o + 65536 <= n; o += 65536) {
            const uint8_t *src = b + o; const int32_t len = 65536, mfl = len - 12, mlim = len - 5;
            static int32_t table[16384]; memset(table, 0, sizeof table);
            int32_t s = 0; uint32_t smc = 67;
            while (s < mfl) {
                /* one walker step: lanes k = 0..31 at s + skip_sum(smc + k) - skip_sum(smc) */
                ++steps;
                int32_t pos[32]; uint32_t hh[32]; int nl = 0;
                for (int k = 0; k < 32; ++k) { int32_t p = s + (int32_t)(skip_sum(smc + k) - skip_sum(smc)); if (p >= mfl) break; pos[nl] = p; hh[nl] = (rd32(src + p) * 2654435761u) >> 18; ++nl; }
                /* cut at the lowest lane involved in a same-slot pair, inclusive */
                int cut = nl;
                for (int k = 0; k < nl; ++k) for (int j = 0; j < k; ++j) if (hh[j] == hh[k] && j + 1 < cut) cut = j + 1;
                if (cut < nl) ++cuts;
                int done = 0;
                for (int k = 0; k < cut; ++k) {
                    int32_t p = pos[k]; int32_t m = table[hh[k]] - 1; table[hh[k]] = p + 1; ++probes;
                    int is_stale = m >= 0 && m > p - lag && m < p;   /* inserted after the producer looked */
                    int ok = !(m < 0 || m == p || ((uint32_t)(p - m) >> 16));
                    int hit = ok && rd32(src + m) == rd32(src + p);
                    if (is_stale) { ++stale; if (hit) ++stale_hit; }
                    if (hit) {
                        int32_t sp = p + 4, mp = m + 4;
                        while (sp < mlim && src[sp] == src[mp]) { ++sp; ++mp; }
                        if (sp - p > 36) ++longm;
                        ++seqs; s = sp; smc = 67; done = 1; break;
                    }
                    if (is_stale) { /* the walker resolves a stale miss on its own and ends the step behind it */
                        s = p + (int32_t)((smc + k) >> 6); smc += k + 1; done = 1; break; }
                }
                if (!done) { ++nohit; s += (int32_t)(skip_sum(smc + cut) - skip_sum(smc)); smc += cut; }
            }
        }
        printf("%s lag %3d: %.2f steps/seq, %.2f probes/seq, stale %.2f%% of probes (%.2f%% of sequences are stale hits), cut steps %.1f%%, no-hit steps %.1f%%, long (>36) %.1f%% of seqs, %.1f seq/KiB\n",
               kind, lag, (double)steps / seqs, (double)probes / seqs, 100.0 * stale / probes, 100.0 * stale_hit / seqs, 100.0 * cuts / steps, 100.0 * nohit / steps, 100.0 * longm / seqs, seqs / (n / 1024.0));
    }
    return 0;
}
