/*
 * corpus.c -- deterministic synthetic corpora for tests and bench.py.
 * Host-only helper: not on the product path, not part of the oracle.
 *
 *   LOG(seed,n)     newline-terminated service log lines, truncated to n bytes
 *   ZERO(n)         n zero bytes
 *   RAND(seed,n)    SplitMix64 bytes (incompressible -> stored blocks)
 *   MIXED(seed,n)   segments of 40000 + r%120000 bytes, type r%4 in {0,1:LOG, 2:ZERO, 3:RAND}
 *   JSONMSG(seed,i) one 4096-byte JSON-like message (fixed keys, random short values, space padded)
 *   BENCHJSON(n)    the reference benchmark's 219-byte record repeated (benchmark/src/base/benchUtils.js:7-22)
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>

typedef struct { uint64_t x; } sm64_t;
static inline uint64_t sm64_next(sm64_t *s) {
    uint64_t z = (s->x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static const char *LEVELS[20] = { "INFO", "INFO", "INFO", "INFO", "INFO", "INFO", "INFO", "INFO", "INFO", "INFO",
                                  "INFO", "INFO", "INFO", "INFO", "DEBUG", "DEBUG", "DEBUG", "WARN", "WARN", "ERROR" };
static const char *SVCS[16] = { "auth", "gateway", "billing", "search", "inventory", "orders", "payments", "notify",
                                "profile", "catalog", "shipping", "analytics", "scheduler", "storage", "session", "router" };
static const char *MSGS[8] = { "request completed", "request failed upstream timeout", "cache miss refreshing entry",
                               "connection pool exhausted retrying", "user session validated", "rate limit applied to client",
                               "background job finished", "database query slow path" };
static const char *METHODS[4] = { "GET", "POST", "PUT", "DELETE" };
static const char *PATHS[16] = { "/api/v1/users", "/api/v1/orders", "/api/v1/items", "/api/v1/cart", "/api/v1/login",
                                 "/api/v1/logout", "/api/v2/search", "/api/v2/recommend", "/api/v2/payments", "/healthz",
                                 "/metrics", "/api/v1/profile", "/api/v1/shipments", "/api/v1/invoices", "/static/app.js",
                                 "/api/v2/events" };
static const int STATUS[16] = { 200, 200, 200, 200, 200, 200, 200, 200, 200, 200, 200, 200, 201, 204, 404, 500 };

typedef struct { sm64_t rng; uint64_t ms; } logstate_t;

/* one line into buf (>= 320 bytes); returns its length */
static int log_line(logstate_t *st, char *buf) {
    sm64_t *r = &st->rng;
    st->ms += sm64_next(r) % 50;
    uint64_t ms = st->ms;
    uint64_t sec = ms / 1000, msec = ms % 1000;
    uint64_t day = sec / 86400, sod = sec % 86400;
    int h = (int)(sod / 3600), mi = (int)((sod % 3600) / 60), s = (int)(sod % 60);
    int dom = 1 + (int)(day % 28), mon = 1 + (int)((day / 28) % 12), yr = 1965 + (int)(day / 336);
    uint64_t a = sm64_next(r), b = sm64_next(r), c = sm64_next(r);
    return sprintf(buf,
                   "%04d-%02d-%02dT%02d:%02d:%02d.%03dZ %s %s[%d]: %s method=%s path=%s status=%d req_id=%016llx user=%d dur_ms=%d bytes=%d\n",
                   yr, mon, dom, h, mi, s, (int)msec, LEVELS[a % 20], SVCS[(a >> 8) % 16], 1000 + (int)((a >> 16) % 64),
                   MSGS[(a >> 24) % 8], METHODS[(a >> 32) % 4], PATHS[(a >> 40) % 16], STATUS[(a >> 48) % 16],
                   (unsigned long long)b, (int)(c % 100000), (int)((c >> 20) % 2000), (int)((c >> 40) % 65536));
}

static void fill_log(logstate_t *st, uint8_t *out, uint64_t n) {
    char line[512];
    uint64_t pos = 0;
    while (pos < n) {
        int len = log_line(st, line);
        uint64_t take = (uint64_t)len < n - pos ? (uint64_t)len : n - pos;
        memcpy(out + pos, line, take);
        pos += take;
    }
}

static void fill_rand(sm64_t *r, uint8_t *out, uint64_t n) {
    uint64_t pos = 0;
    while (pos + 8 <= n) { uint64_t v = sm64_next(r); memcpy(out + pos, &v, 8); pos += 8; }
    if (pos < n) { uint64_t v = sm64_next(r); memcpy(out + pos, &v, n - pos); }
}

void corpus_log(uint64_t seed, uint8_t *out, uint64_t n) {
    logstate_t st = { { seed }, 1792300000000ull };
    fill_log(&st, out, n);
}
void corpus_zero(uint8_t *out, uint64_t n) { memset(out, 0, n); }
void corpus_rand(uint64_t seed, uint8_t *out, uint64_t n) { sm64_t r = { seed ^ 0xA5A5A5A5ull }; fill_rand(&r, out, n); }

void corpus_mixed(uint64_t seed, uint8_t *out, uint64_t n) {
    sm64_t ctl = { seed * 0x100000001B3ull + 7 };
    logstate_t st = { { seed + 1 }, 1792300000000ull };
    sm64_t rr = { seed + 2 };
    uint64_t pos = 0;
    while (pos < n) {
        uint64_t r = sm64_next(&ctl);
        uint64_t seg = 40000 + (r >> 8) % 120000;
        if (seg > n - pos) seg = n - pos;
        switch (r % 4) {
            case 0: case 1: fill_log(&st, out + pos, seg); break;
            case 2: memset(out + pos, 0, seg); break;
            default: fill_rand(&rr, out + pos, seg); break;
        }
        pos += seg;
    }
}

static const char *JKEYS[40] = { "id", "ts", "type", "source", "region", "tenant", "user_id", "session", "device", "os",
                                 "app_version", "locale", "country", "city", "lat", "lon", "event", "category", "action",
                                 "label", "value", "currency", "price", "quantity", "sku", "cart_id", "order_id", "status",
                                 "latency_ms", "retries", "referrer", "campaign", "experiment", "variant", "flags", "score",
                                 "trace_id", "span_id", "parent_id", "checksum" };
static const char *JWORDS[16] = { "alpha", "bravo", "charlie", "delta", "echo", "foxtrot", "golf", "hotel", "india",
                                  "juliet", "kilo", "lima", "mike", "november", "oscar", "papa" };

void corpus_jsonmsg(uint64_t seed, uint64_t index, uint8_t *out /* 4096 */) {
    sm64_t r = { seed * 0x9E3779B1ull + index * 0x85EBCA77ull + 1 };
    char buf[8192];
    int p = 0;
    buf[p++] = '{';
    for (int round = 0; p < 3900; round++) {
        for (int k = 0; k < 40 && p < 3900; k++) {
            uint64_t v = sm64_next(&r);
            if (round == 0) p += sprintf(buf + p, "\"%s\":", JKEYS[k]);
            else p += sprintf(buf + p, "\"%s_%d\":", JKEYS[k], round);
            switch (v % 4) {
                case 0: p += sprintf(buf + p, "%d,", (int)((v >> 8) % 1000000)); break;
                case 1: p += sprintf(buf + p, "\"%s\",", JWORDS[(v >> 8) % 16]); break;
                case 2: p += sprintf(buf + p, "\"%08x\",", (unsigned)(v >> 16)); break;
                default: p += sprintf(buf + p, "%s,", ((v >> 8) & 1) ? "true" : "false"); break;
            }
        }
    }
    buf[p - 1] = '}';
    if (p > 4096) p = 4096;
    memcpy(out, buf, (size_t)p);
    memset(out + p, ' ', (size_t)(4096 - p));
}

void corpus_jsonmsgs(uint64_t seed, uint64_t first, uint64_t count, uint8_t *out) {
    for (uint64_t i = 0; i < count; i++) corpus_jsonmsg(seed, first + i, out + i * 4096);
}

static const char BENCH_REC[] =
    "{\"id\":1,\"type\":\"benchmark_event\",\"tags\":[\"performance\",\"compression\",\"lz4\",\"javascript\",\"v8\"],"
    "\"meta\":{\"valid\":true,\"scores\":[100,205,300,400,500]},"
    "\"payload\":\"Repeated data is the key to high compression ratios in LZ4.\"}";

uint64_t corpus_benchjson_reclen(void) { return sizeof(BENCH_REC) - 1; }
void corpus_benchjson(uint8_t *out, uint64_t n) {
    const uint64_t L = sizeof(BENCH_REC) - 1;
    for (uint64_t pos = 0; pos < n; pos += L) memcpy(out + pos, BENCH_REC, L < n - pos ? L : n - pos);
}
