/* infer_stats.c -- would neighbour inference replace the producers' second stage (bytes 32..63)?  For every match the exact
 * serial parse takes with length >= 32: look at the positions g+1, g+2, ... with the table as it is when g is probed; a link
 * g+k -> g+k+1 holds if table[hash(g+k+1)] == table[hash(g+k)] + 1 (same offset).  The length of g is known without loading
 * more bytes if the chain of links reaches a position whose own 32-byte verification ends (< 32); otherwise the entry stays OPEN.
 * build: gcc -O2 -o /tmp/infer_stats infer_stats.c ../../../divortio-lz4_b200/tools/corpus.c */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void corpus_log(uint64_t seed, uint8_t *out, uint64_t n);
void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n);
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t H(const uint8_t *p) { return (rd32(p) * 2654435761u) >> 18; }
int main(int argc, char **argv) {
    const char *kind = argc > 1 ? argv[1] : "log"; uint64_t n = (argc > 2 ? atoll(argv[2]) : 64) << 20;
    uint8_t *b = malloc(n + 64);
    if (!strcmp(kind, "log")) corpus_log(3, b, n); else corpus_mixed(2, b, n);
    uint64_t seqs = 0, longm = 0, known = 0, open_ = 0, ge64 = 0, open_lt64 = 0;
    for (uint64_t o = 0; o + 65536 <= n; o += 65536) {
        const uint8_t *src = b + o; const int32_t len = 65536, mfl = len - 12, mlim = len - 5;
        static int32_t table[16384]; memset(table, 0, sizeof table);
        int32_t s = 0; uint32_t smc = 67;
        while (s < mfl) {
            uint32_t h = H(src + s); int32_t m = table[h] - 1; table[h] = s + 1;
            int ok = !(m < 0 || m == s || ((uint32_t)(s - m) >> 16));
            if (ok && rd32(src + m) == rd32(src + s)) {
                int32_t sp = s + 4, mp = m + 4;
                while (sp < mlim && src[sp] == src[mp]) { ++sp; ++mp; }
                int32_t ml = sp - s; ++seqs;
                if (ml >= 32 && s + 100 < len) {
                    ++longm;
                    if (ml >= 64) ++ge64;
                    /* chain inference */
                    int32_t c = m; int k = 0; int res = -1; /* -1 open */
                    for (k = 1; k < 32; ++k) {
                        int32_t g = s + k; int32_t cg = table[H(src + g)] - 1;
                        if (cg != c + k) break;                    /* link broken */
                        /* position g's own verification: equal bytes up to 32 */
                        int e = 0; while (e < 32 && src[g + e] == src[cg + e]) ++e;
                        if (e < 32) { res = k + e; break; }
                    }
                    if (res >= 0 || (k >= 32)) ++known; else { ++open_; if (ml < 64) ++open_lt64; }
                }
                s = sp; smc = 67;
            } else { s += (int32_t)(smc >> 6); ++smc; }
        }
    }
    printf("%s: %llu sequences, %.1f%% with ml >= 32 (%.1f%% >= 64); of those: known by inference %.1f%%, open %.1f%% (open and < 64: %.1f%%)\n", kind,
           (unsigned long long)seqs, 100.0 * longm / seqs, 100.0 * ge64 / seqs, 100.0 * known / longm, 100.0 * open_ / longm, 100.0 * open_lt64 / longm);
    return 0;
}
