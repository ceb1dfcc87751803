// dlz4_api.cu -- C ABI (include/dlz4_b200.h) over the sm_100a kernels in dlz4_kernels.cuh.
// Host code here only moves bytes, launches kernels and writes the frame header/footer; every block is
// compressed, decoded and hashed on the GPU.  There is no CPU path.
#include "../../include/dlz4_b200.h"
#include "dlz4_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace dlz4;

namespace {

constexpr int kWarpsFresh16 = 7;     // 7 x 32 KiB tables = 224 KiB of the 227 KiB a CTA may own
constexpr int kWarpsGeneric32 = 3;   // 3 x 64 KiB
constexpr int kWarpsDecode = 8;
constexpr int kJdGroupEvent = 112;    // evp[112..127]: input groups of the jump decoder (evp[0..111]: its output units)
constexpr int kGtabRegions = 5;      // L2-table regions: work-queue counter 0 (device API) and 1..4 (pipeline lanes)
constexpr uint32_t kMaxSplitBlocks = 1u << 20;   // blocks per launch of the split (parse + encode) path

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

// Host threads for copies between ordinary (pageable) caller memory and the page-locked staging ring: the reference's
// callers own plain Uint8Arrays (bufferCompress.js:100), and one thread's memcpy -- or the driver's own staging inside
// cudaMemcpyAsync -- moves 5-6 GB/s where the PCIe link takes 55.
class CopyPool {
  public:
    explicit CopyPool(int threads) {
        for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    // memcpy(dst, src, n) cut into pieces over the workers and the caller
    void copy(void *dst, const void *src, size_t n) {
        const size_t piece = 1u << 20;
        const size_t parts = (n + piece - 1) / piece;
        if (parts <= 1 || workers_.empty()) { memcpy(dst, src, n); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = (uint8_t *)dst; src_ = (const uint8_t *)src; n_ = n; piece_ = piece; parts_ = parts;
            next_.store(0); done_ = 0; ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [&] { return done_ == parts_; });
    }

  private:
    void work() {
        size_t mine = 0;
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= parts_) break;
            const size_t o = i * piece_;
            memcpy(dst_ + o, src_ + o, std::min(piece_, n_ - o));
            ++mine;
        }
        if (mine) {
            std::lock_guard<std::mutex> g(m_);
            done_ += mine;
            if (done_ == parts_) cv_done_.notify_all();
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    bool stop_ = false;
    uint64_t gen_ = 0;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t n_ = 0, piece_ = 0, parts_ = 0, done_ = 0;
    std::atomic<size_t> next_{0};
};
constexpr int kSumSlots = 32;                  // concurrent whole-stream checksums (dlz4_xxh32_async)
constexpr int kStageSlots = 4;                 // page-locked staging ring for pageable caller buffers
constexpr size_t kStageBytes = 16u << 20;

}  // namespace

struct dlz4_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, side = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaStream_t lanes[4] = {};         // compute streams of the chunked host pipeline (chunks rotate over them)
    uint64_t chunk_bytes = 128ull << 20; // target uncompressed bytes per pipeline chunk (DLZ4_CHUNK_MIB)
    int n_lanes = 4;                    // compute streams in use (DLZ4_LANES)
    cudaEvent_t evp[128] = {};          // event pool of the chunked host pipeline: [0,64) copies landed, [64,128) kernels done
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_side = nullptr, ev_fork = nullptr;
    uint32_t *d_counter = nullptr;      // work-queue heads (one per launch slot)
    uint32_t *d_land = nullptr;         // kLandFlags "input chunk is in memory" flags (segment engine under a chunked H2D copy)
    uint32_t *h_one = nullptr;          // pinned word holding 1: source of the flag copies
    uint64_t seg_overlap_min_bytes = 64ull << 20;            // segment-engine frames at least this long overlap H2D and parse
    uint32_t *d_hash = nullptr;         // small result slots
    uint64_t *d_total = nullptr;
    int32_t *d_table = nullptr;         // int32[16384] scratch table
    int wide = 1;                       // shared-memory-table chains use the 64-position window (dlz4_wide.cuh); 0: A/B runs
    int split = 1;                      // fresh blocks <= 64 KiB: match finder (k_parse_pw / k_parse_fresh16) + encoder (k_encode_blocks); 0: A/B runs
    int pw = 2;                         // producers per chain of the match finder (k_parse_pw<2|3>); 0: one warp per chain (k_parse_fresh16)
    int pw_fused = 0;                   // 1: the encoder is a fourth kind of warp inside k_parse_pw (measured slower: it delays the next block)
    int pw_sleep = 200;                 // ns a producer sleeps when its ring is full
    int pw_lead = 7;                    // windows a producer may run ahead of the walker (3..kPwWin - 1)
    Buf rec;                            // match records of the split path: kGtabRegions regions (one per work-queue counter)
    uint32_t *d_nrec = nullptr;         // matches per block (split path), kGtabRegions regions of kMaxSplitBlocks
    int hybrid = 1;                     // 64 KiB fresh blocks: hybrid kernel (L2-resident tables) instead of the 7-warp one
    int hy_grid = 0;                    // CTAs that fill the device (sm_count x kHyCtasPerSm)
    int hy_active = 0;                  // cap on the warps used per hybrid CTA (DLZ4_HY_ACTIVE, 0 = all 7; A/B runs)
    uint16_t *d_gtabs = nullptr;        // kGtabRegions x hy_grid x kHyGlWarps tables of 16384 x u16 (one region per stream lane)
    Buf work, comp, seg, out, meta, aux, pin;
    std::string last_error;
    uint64_t frame_pipe_min_bytes = 32ull << 20;             // independent 64 KiB-block frames at least this long: chunked pipeline
    uint64_t jump_min_bytes = 256ull << 10;                  // frames at least this long may use the jump decoder (DLZ4_JUMP_MIN_KIB)
    uint64_t seg_min_bytes = 256ull << 10;                   // frames at least this long use the segment-parallel engine (DLZ4_SEG_MIN_KIB)
    uint32_t seg_jobs = 0, seg_reruns = 0, seg_rounds = 0;   // last segment-parallel call: segments, re-run segments, rounds
    uint64_t launches = 0;
    float last_ms = 0.f;
    // measurement knobs, read from the environment ONCE in dlz4_init (A/B runs; none changes an output byte).  -1 / 0 = automatic
    int64_t k_seg_kib = -1, k_seg_warm_kib = -1, k_seg_unit_kib = -1;
    int k_seg_group = -1, k_seg_no_chase = 0, k_jd_unit_mib = 0, k_jd_serial_scan = 0, k_jd_no_overlap = 0, k_debug = 0;
    CopyPool *pool = nullptr;           // host copy threads (pageable caller buffers), created on first use
    int copy_threads = 0;               // 0: min(8, hardware threads / 2)
    uint8_t *stage[kStageSlots] = {};   // page-locked staging ring
    cudaEvent_t stage_ev[kStageSlots] = {};
    bool stage_busy[kStageSlots] = {};
    cudaStream_t sum_stream[kSumSlots] = {};   // dlz4_xxh32_async: one serial chain per slot, each on its own stream
    uint32_t *h_sum = nullptr;          // their results (page-locked, written by the kernels)
    int probe = 0;                      // dlz4_kernel_probe: time the match finder and the encoder of the next batch separately
    cudaEvent_t evq[3] = {};            // before the match finder, between the two kernels, behind the encoder
    // what the last frame-body / frame-range call left resident on the device (sharded frames): the rank's input
    // slice, its packed frame body, the decoded bytes of its block range -- read by dlz4_frame_body_fetch and by the
    // content-checksum relay dlz4_xxh32_update_resident
    const uint8_t *res_in = nullptr, *res_out = nullptr;
    uint64_t res_in_len = 0, res_out_len = 0, res_body_len = 0;
};

namespace {

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof m_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->last_error = m_;                                                                     \
            return DLZ4_E_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define CKS(expr)                       \
    do {                                \
        int s_ = (expr);                \
        if (s_ != DLZ4_OK) return s_;   \
    } while (0)

int reserve(dlz4_ctx *ctx, Buf &b, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes + 256 <= b.cap) return DLZ4_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + 256;            // slack so aligned word reads at the very end stay inside
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return DLZ4_OK;
}

int reserve_pinned(dlz4_ctx *ctx, Buf &b, size_t bytes) {
    if (bytes <= b.cap) return DLZ4_OK;
    if (b.p) CK(cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    CK(cudaMallocHost(&b.p, bytes));
    b.cap = bytes;
    return DLZ4_OK;
}

inline cudaStream_t pick(dlz4_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

// Ordinary host memory (neither cudaMallocHost nor cudaHostRegister)?
bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int stage_init(dlz4_ctx *ctx) {
    if (ctx->stage[0]) return DLZ4_OK;
    for (int i = 0; i < kStageSlots; ++i) {
        CK(cudaMallocHost((void **)&ctx->stage[i], kStageBytes));
        CK(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
        ctx->stage_busy[i] = false;
    }
    int t = ctx->copy_threads;
    if (t <= 0) t = (int)std::min<unsigned>(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    ctx->pool = new CopyPool(t - 1);             // the calling thread copies too
    return DLZ4_OK;
}

// Host -> device on stream `s`.  Page-locked source: one asynchronous copy.  Pageable source of some size: pieces go through
// the staging ring (host threads copy a piece into a page-locked slot, the slot is sent asynchronously, the next piece is
// copied meanwhile); returns when the last piece is queued -- the caller's memory is no longer read after that.
int h2d(dlz4_ctx *ctx, void *d, const void *h, size_t n, cudaStream_t s, int pageable /* -1: ask the driver */ = -1) {
    if (!n) return DLZ4_OK;
    if (pageable < 0) pageable = n >= (4u << 20) && is_pageable(h);
    if (!pageable || n < (4u << 20)) { CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); return DLZ4_OK; }
    CKS(stage_init(ctx));
    int slot = 0;
    for (size_t o = 0; o < n; o += kStageBytes, slot = (slot + 1) % kStageSlots) {
        const size_t len = std::min(kStageBytes, n - o);
        if (ctx->stage_busy[slot]) CK(cudaEventSynchronize(ctx->stage_ev[slot]));
        ctx->pool->copy(ctx->stage[slot], (const uint8_t *)h + o, len);
        CK(cudaMemcpyAsync((uint8_t *)d + o, ctx->stage[slot], len, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(ctx->stage_ev[slot], s));
        ctx->stage_busy[slot] = true;
    }
    return DLZ4_OK;
}

// Device -> host on stream `s`.  Page-locked destination: one asynchronous copy (the caller synchronises).  Pageable
// destination: the pieces arrive in the staging ring and host threads copy them out while the next ones are in flight;
// returns when the caller's memory holds the bytes.
int d2h(dlz4_ctx *ctx, void *h, const void *d, size_t n, cudaStream_t s, int pageable = -1) {
    if (!n) return DLZ4_OK;
    if (pageable < 0) pageable = n >= (4u << 20) && is_pageable(h);
    if (!pageable || n < (4u << 20)) { CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); return DLZ4_OK; }
    CKS(stage_init(ctx));
    for (int i = 0; i < kStageSlots; ++i)
        if (ctx->stage_busy[i]) { CK(cudaEventSynchronize(ctx->stage_ev[i])); ctx->stage_busy[i] = false; }
    const size_t pieces = (n + kStageBytes - 1) / kStageBytes;
    auto issue = [&](size_t i) -> int {
        const size_t o = i * kStageBytes, len = std::min(kStageBytes, n - o);
        CK(cudaMemcpyAsync(ctx->stage[i % kStageSlots], (const uint8_t *)d + o, len, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->stage_ev[i % kStageSlots], s));
        return DLZ4_OK;
    };
    for (size_t i = 0; i < std::min<size_t>(pieces, kStageSlots); ++i) CKS(issue(i));
    for (size_t i = 0; i < pieces; ++i) {
        const size_t o = i * kStageBytes, len = std::min(kStageBytes, n - o);
        CK(cudaEventSynchronize(ctx->stage_ev[i % kStageSlots]));
        ctx->pool->copy((uint8_t *)h + o, ctx->stage[i % kStageSlots], len);
        if (i + kStageSlots < pieces) CKS(issue(i + kStageSlots));
    }
    return DLZ4_OK;
}

// host xxh32 for the <= 14 header bytes only (FLG..dictID -> HC byte, bufferCompress.js:177-178)
uint32_t header_xxh32(const uint8_t *p, size_t len) {
    auto rd = [](const uint8_t *q) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); };
    auto rl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    const uint8_t *end = p + len;
    uint32_t h = 374761393u + (uint32_t)len;        // len < 16 always
    while (p + 4 <= end) { h = rl(h + rd(p) * 3266489917u, 17) * 668265263u; p += 4; }
    while (p < end) { h = rl(h + (*p) * 374761393u, 11) * 2654435761u; ++p; }
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return h;
}

inline void wr32(uint8_t *b, uint32_t v) { b[0] = (uint8_t)v; b[1] = (uint8_t)(v >> 8); b[2] = (uint8_t)(v >> 16); b[3] = (uint8_t)(v >> 24); }
inline uint32_t rd32(const uint8_t *b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }

int block_id_for(uint64_t bytes) {          // bufferCompress.js:77-82
    if (bytes == 0 || bytes <= 65536) return 4;
    if (bytes <= 262144) return 5;
    if (bytes <= 1048576) return 6;
    return 7;
}
const uint32_t kBlockMax[8] = {0, 0, 0, 0, 65536, 262144, 1048576, 4194304};
const uint32_t kLandFlags = 256, kLandShift = 24;      // 16 MiB H2D chunks, one flag each (inputs are < 2 GiB: <= 128 chunks)

// ---- launch helpers ----------------------------------------------------------------------------------
int compress_segmented(dlz4_ctx *ctx, const uint8_t *d_work, int64_t start, int64_t total, int64_t B, uint32_t n, bool linked,
                       const int32_t *init_table, uint8_t *d_comp, const uint64_t *d_coff, uint32_t *d_clen, cudaStream_t st,
                       int32_t *final_table = nullptr, const uint32_t *landed = nullptr, int32_t land_origin = 0,
                       uint32_t land_shift = 0);

int launch_compress(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len, uint32_t n,
                    uint32_t max_len, const uint8_t *prefix, uint32_t prefix_len, const int32_t *init_table, uint8_t *dst,
                    const uint64_t *dst_off, uint32_t *comp_len, cudaStream_t st, uint32_t *counter = nullptr, bool dense = false) {
    if (n == 0) return DLZ4_OK;
    if (!counter) counter = ctx->d_counter;
    CK(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    // a prefix without an initial table is never referenced: every table entry was inserted by the block itself
    // (blockCompress.js:54-55), so the bytes are those of the block compressed alone
    if (init_table == nullptr) { prefix = nullptr; prefix_len = 0; }
    if (max_len <= 65536 && prefix_len == 0 && init_table == nullptr && ctx->split && n <= kMaxSplitBlocks) {
        // match finder, then encoder (dlz4_parse.cuh).  Records: len/4 + 1 per block, one scratch region per work-queue counter
        // so that the chunks of the host pipeline (one counter each) do not share it.
        const size_t region = (size_t)(counter - ctx->d_counter) % kGtabRegions;
        const uint64_t rstride = (uint64_t)max_len / 4 + 2;
        const size_t region_bytes = (((size_t)n * rstride * 8) * 9 / 8 + 255) & ~(size_t)255;      // (headroom: chunks differ a little)
        if (ctx->rec.cap < region_bytes * kGtabRegions + 256) {
            // grow-only scratch: every stream that may still read the old one is idle after this
            CK(cudaDeviceSynchronize());
            CKS(reserve(ctx, ctx->rec, std::max(region_bytes, (size_t)ctx->rec.cap / kGtabRegions) * kGtabRegions));
        }
        const size_t per_region = ((ctx->rec.cap - 256) / kGtabRegions) & ~(size_t)255;
        uint64_t *recs = (uint64_t *)((uint8_t *)ctx->rec.p + region * per_region);
        uint32_t *nrec = ctx->d_nrec + region * (size_t)kMaxSplitBlocks;
        if (ctx->probe) CK(cudaEventRecord(ctx->evq[0], st));
        if (ctx->pw) {
            // teams spread over the SMs first (a small batch uses one chain per SM), six teams per CTA at most
            const int grid = (int)std::min<uint64_t>(n, (uint64_t)ctx->sm_count);
            const uint32_t lead = (uint32_t)ctx->pw_lead, ns = (uint32_t)ctx->pw_sleep;
            const size_t sm = (size_t)kPwChains * kPwChainBytes;
            if (ctx->pw_fused) {
                // walker, producers and encoder of a chain in one kernel
                if (ctx->pw == 2)
                    k_parse_pw<2, true><<<grid, kPwChains * 4 * 32, sm, st>>>(src, src_off, src_len, n, recs, rstride, nrec, counter, lead, ns, dst, dst_off, comp_len);
                else
                    k_parse_pw<3, true><<<grid, kPwChains * 5 * 32, sm, st>>>(src, src_off, src_len, n, recs, rstride, nrec, counter, lead, ns, dst, dst_off, comp_len);
                ctx->launches++;
                CK(cudaGetLastError());
                return DLZ4_OK;
            }
            if (ctx->pw == 2)
                k_parse_pw<2, false><<<grid, kPwChains * 3 * 32, sm, st>>>(src, src_off, src_len, n, recs, rstride, nrec, counter, lead, ns, nullptr, nullptr, nullptr);
            else
                k_parse_pw<3, false><<<grid, kPwChains * 4 * 32, sm, st>>>(src, src_off, src_len, n, recs, rstride, nrec, counter, lead, ns, nullptr, nullptr, nullptr);
        } else
        {
            const int grid = (int)std::min<uint64_t>((n + kWarpsFresh16 - 1) / kWarpsFresh16, (uint64_t)ctx->sm_count);
            k_parse_fresh16<kWarpsFresh16><<<grid, kWarpsFresh16 * 32, kWarpsFresh16 * kHashEntries * 2, st>>>(
                src, src_off, src_len, n, recs, rstride, nrec, counter);
        }
        if (ctx->probe) CK(cudaEventRecord(ctx->evq[1], st));
        CK(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
        const int egrid = (int)std::min<uint64_t>((n + kWarpsDecode - 1) / kWarpsDecode, (uint64_t)ctx->sm_count * 8);
        k_encode_blocks<kWarpsDecode><<<egrid, kWarpsDecode * 32, 0, st>>>(src, src_off, src_len, n, recs, rstride, nrec, dst, dst_off,
                                                                            comp_len, counter);
        if (ctx->probe) CK(cudaEventRecord(ctx->evq[2], st));
        ctx->launches++;
    } else if (max_len <= 65536 && prefix_len == 0 && init_table == nullptr && ctx->hybrid) {
        // one table region per work-queue counter: kernels of different pipeline lanes run concurrently
        const size_t region = (size_t)(counter - ctx->d_counter) % kGtabRegions;
        // a batch on its own spreads over as many SMs as it has blocks (fewer active warps per CTA); a pipeline chunk
        // (`dense`) packs 7 chains per CTA so that the chunks in flight on the other lanes find free CTA slots
        const int grid = (int)std::min<uint64_t>(dense ? (n + kHyWarps - 1) / kHyWarps : n, (uint64_t)ctx->hy_grid);
        uint32_t active = dense ? (uint32_t)kHyWarps : (uint32_t)std::min<uint64_t>((n + grid - 1) / grid, (uint64_t)kHyWarps);
        if (ctx->hy_active) active = std::min<uint32_t>(active, (uint32_t)ctx->hy_active);
        k_compress_fresh16h<<<grid, kHyWarps * 32, kHySmemBytes, st>>>(
            src, src_off, src_len, n, dst, dst_off, comp_len, counter,
            ctx->d_gtabs + region * (size_t)ctx->hy_grid * kHyGlWarps * kHashEntries, active);
    } else if (max_len <= 4096 && prefix_len > 0 && init_table != nullptr && ctx->hybrid && prefix_len < 0x7FFF0000u) {
        // small blocks behind a shared prefix, all from the same initial table: read-only base table + per-warp overlay
        const size_t region = (size_t)(counter - ctx->d_counter) % kGtabRegions;
        const int grid = (int)std::min<uint64_t>(dense ? (n + kHyWarps - 1) / kHyWarps : n, (uint64_t)ctx->hy_grid);
        const uint32_t active = dense ? (uint32_t)kHyWarps : (uint32_t)std::min<uint64_t>((n + grid - 1) / grid, (uint64_t)kHyWarps);
        k_compress_overlay<<<grid, kHyWarps * 32, kHySmemBytes, st>>>(
            src, src_off, src_len, n, prefix, prefix_len, init_table, dst, dst_off, comp_len, counter,
            ctx->d_gtabs + region * (size_t)ctx->hy_grid * kHyGlWarps * kHashEntries, active);
    } else if (max_len <= 65536 && prefix_len == 0 && init_table == nullptr) {
        const int grid = (int)std::min<uint64_t>((n + kWarpsFresh16 - 1) / kWarpsFresh16, (uint64_t)ctx->sm_count);
        if (ctx->wide)
            k_compress_fresh16<kWarpsFresh16, true><<<grid, kWarpsFresh16 * 32, kWarpsFresh16 * (kHashEntries * 2 + kRingBytes), st>>>(
                src, src_off, src_len, n, dst, dst_off, comp_len, counter);
        else
            k_compress_fresh16<kWarpsFresh16, false><<<grid, kWarpsFresh16 * 32, kWarpsFresh16 * (kHashEntries * 2 + kRingBytes), st>>>(
                src, src_off, src_len, n, dst, dst_off, comp_len, counter);
    } else {
        if (max_len > 65536 && prefix_len == 0 && init_table == nullptr && n <= 65536) {
            // blocks > 64 KiB: if they tile one contiguous range uniformly (the usual batch), cut them into segments
            // (k_compress_segments) instead of one 3-per-SM chain per block.  The descriptors live on the device: read them back.
            std::vector<uint64_t> off(n);
            std::vector<uint32_t> len(n);
            CK(cudaMemcpyAsync(off.data(), src_off, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(len.data(), src_len, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const uint64_t B = len[0];
            bool uniform = B > 65536 && (B & (B - 1)) == 0 && off[0] + (uint64_t)n * B < 0x7FFFFFF0ull;
            for (uint32_t i = 0; uniform && i < n; ++i)
                uniform = off[i] == off[0] + (uint64_t)i * B && (len[i] == B || (i + 1 == n && len[i] <= B && len[i] > 0));
            if (uniform) {
                const int64_t total = (int64_t)(n - 1) * (int64_t)B + len[n - 1];
                return compress_segmented(ctx, src, (int64_t)off[0], total, (int64_t)B, n, false, nullptr, dst, dst_off, comp_len, st);
            }
        }
        const int grid = (int)std::min<uint64_t>((n + kWarpsGeneric32 - 1) / kWarpsGeneric32, (uint64_t)ctx->sm_count);
        k_compress_generic32<kWarpsGeneric32><<<grid, kWarpsGeneric32 * 32, kWarpsGeneric32 * (kHashEntries * 4 + kRingBytes), st>>>(
            src, src_off, src_len, n, prefix, prefix_len, init_table, dst, dst_off, comp_len, counter);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

int launch_decompress(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len, uint32_t n,
                      uint8_t *dst, const uint64_t *dst_off, const uint32_t *dst_cap, const uint8_t *dict, uint32_t dict_len,
                      int hist_frame, const uint8_t *stored, uint32_t *out_len, uint8_t *status, cudaStream_t st,
                      uint32_t *counter = nullptr) {
    if (n == 0) return DLZ4_OK;
    if (!counter) counter = ctx->d_counter;
    CK(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    // frame history: a block may read what the blocks before it wrote, so ONE warp takes them in queue order (the parallel
    // route for linked data is the frame call's jump decoder)
    const int grid = hist_frame ? 1 : (int)std::min<uint64_t>((n + kWarpsDecode - 1) / kWarpsDecode, (uint64_t)ctx->sm_count * 8);
    if (dict_len)
        k_decompress_blocks<kWarpsDecode, true><<<grid, hist_frame ? 32 : kWarpsDecode * 32, 0, st>>>(src, src_off, src_len, n, dst, dst_off, dst_cap,
                                                                                                      dict, dict_len, hist_frame, stored, out_len, status, counter);
    else
        k_decompress_blocks<kWarpsDecode, false><<<grid, hist_frame ? 32 : kWarpsDecode * 32, 0, st>>>(src, src_off, src_len, n, dst, dst_off, dst_cap,
                                                                                                       dict, dict_len, hist_frame, stored, out_len, status, counter);
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

int launch_xxh32_batch(dlz4_ctx *ctx, const uint8_t *base, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t seed,
                       uint32_t *out, uint8_t *append_base, cudaStream_t st) {
    if (n == 0) return DLZ4_OK;
    const uint64_t threads = (uint64_t)n * 4;
    k_xxh32_batch<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(base, off, len, n, seed, out, append_base);
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

int launch_xxh32_stream(dlz4_ctx *ctx, const uint8_t *data, uint64_t len, uint32_t seed, uint32_t *out, cudaStream_t st) {
    k_xxh32_stream<1><<<1, 32, 0, st>>>(data, len, seed, out, nullptr, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

// Segment-parallel compression of large blocks / linked chains (k_compress_segments).  Blocks are the uniform blocks of
// size B that tile [start, start + total) of the working buffer; linked: one chain carrying the table (init_table = its
// initial state), otherwise every block is its own chain with a fresh table.  Output: d_comp + d_coff[b], d_clen[b].
int compress_segmented(dlz4_ctx *ctx, const uint8_t *d_work, int64_t start, int64_t total, int64_t B, uint32_t n, bool linked,
                       const int32_t *init_table, uint8_t *d_comp, const uint64_t *d_coff, uint32_t *d_clen, cudaStream_t st,
                       int32_t *final_table /* nullable, device int32[16384]: the chain's table after its last block (linked only) */,
                       const uint32_t *landed /* nullable: input still arriving, see k_compress_segments */, int32_t land_origin,
                       uint32_t land_shift) {
    if (n == 0) return DLZ4_OK;
    // segment size: about 2048 segments over the call, at least 128 KiB; warm-up 512 KiB (tools/resync_stats.c)
    int64_t S = 128 << 10;
    {
        // up to ~130 MiB: as few segments as the shared-memory-table slots (3 per SM: those chains run twice as fast as the
        // L2-table ones and a short call is one wave of (warm-up + segment) / chain speed); beyond that ~2048 segments
        const int64_t slots = (int64_t)ctx->sm_count * kSegCtasPerSm;
        const int64_t fit = (((total + slots - 1) / slots) + 65535) & ~(int64_t)65535;
        if (fit <= (320 << 10)) S = std::max(S, fit);
        else while (S < total / 2048) S <<= 1;
    }
    if (ctx->k_seg_kib >= 0) S = std::max<int64_t>(64, ctx->k_seg_kib) << 10;
    int64_t W = 512 << 10;
    if (ctx->k_seg_warm_kib >= 0) W = ctx->k_seg_warm_kib << 10;
    if (!linked && S > B) S = B;
    // Groups: the chains on shared-memory tables (one per CTA) run about 1.5 times as fast as those on L2 tables when the
    // device is full, so once there are more segments than such slots, G consecutive segments form a group that one
    // shared-memory warp runs back to back -- one warm-up for the group, every later member starts from its predecessor's end
    // state (kSegCont) -- while the L2-table warps take single segments.  G makes the group's work, W + G S, about 1.5 times
    // a single segment's, W + S: both kinds finish together and fewer L2-table chains compete for the L2
    // (profiles/r01c_seg_groups.txt).  Members stay separate segments for verification, so a failed speculation still
    // re-runs S bytes only.
    const int64_t slots = (int64_t)ctx->sm_count * kSegCtasPerSm;
    int64_t G = (total + S - 1) / S > slots ? std::max<int64_t>(2, (W + 3 * S) / (2 * S)) : 0;
    if (ctx->k_seg_group >= 0) G = ctx->k_seg_group;
    if (G < 2) G = 0;
    // Verification unit U <= S: a warp's stretch of S bytes is itself a run of S/U segments (kSegCont / kSegMore), so that a
    // failed speculation re-runs U bytes, not S: the re-run converges to the speculative run's state within the unit and
    // the next member's snapshot then verifies (DLZ4_SEG_UNIT_KIB, 0 = S)
    int64_t U = S % (128 << 10) == 0 ? (128 << 10) : S % (64 << 10) == 0 ? (64 << 10) : S;
    if (ctx->k_seg_unit_kib >= 0) { const int64_t u = ctx->k_seg_unit_kib << 10; U = u > 0 && S % u == 0 ? u : S; }
    const int64_t gs = S / U;
    std::vector<SegJob> jobs;
    std::vector<uint32_t> heads, singles;                  // first-launch queues (job indices)
    std::vector<uint32_t> slot_first(n, 0), slot_count(n, 0);
    uint32_t nslots = 0;
    const uint32_t nchains = linked ? 1u : n;
    for (uint32_t c = 0; c < nchains; ++c) {
        const int64_t cs = linked ? start : start + (int64_t)c * B;
        const int64_t ce = linked ? start + total : std::min<int64_t>(cs + B, start + total);
        const size_t j0 = jobs.size();
        for (int64_t sb = cs; sb < ce; sb += U) {
            SegJob J;
            J.chain_start = (int32_t)cs; J.chain_end = (int32_t)ce;
            J.seg_begin = (int32_t)sb; J.seg_end = (int32_t)std::min<int64_t>(sb + U, ce);
            J.warm_begin = (int32_t)std::max<int64_t>(cs, sb - W);
            J.flags = (sb == cs ? kSegFirst : 0u) | (J.seg_end == J.chain_end ? kSegLast : 0u);
            const int64_t b0 = (sb - cs) / B, b1 = (J.seg_end - 1 - cs) / B;          // blocks of the chain this segment overlaps
            J.first_block = (linked ? 0u : c) + (uint32_t)b0;
            J.first_slot = nslots;
            for (int64_t b = b0; b <= b1; ++b) {
                const uint32_t g = (linked ? 0u : c) + (uint32_t)b;
                if (!slot_count[g]) slot_first[g] = nslots;
                slot_count[g]++;
                nslots++;
            }
            jobs.push_back(J);
        }
        // groups (G stretches) and single stretches of this chain, interleaved evenly
        const int64_t njc = (int64_t)(jobs.size() - j0);
        const int64_t nst = (njc + gs - 1) / gs;                               // stretches of S bytes
        // (never more groups than shared-memory slots over all chains: a group on an L2-table warp would be the tail)
        const int64_t share = ((int64_t)(c + 1) * slots) / nchains - ((int64_t)c * slots) / nchains;
        const int64_t nb = G ? std::min<int64_t>(nst / G, share) : 0, ns = nst - nb * G;
        int64_t j = 0;
        for (int64_t k = 0; k < nb + ns; ++k) {
            const bool group = nb && (k + 1) * nb / (nb + ns) > k * nb / (nb + ns);
            const int64_t members = std::min<int64_t>((group ? G : 1) * gs, njc - j);
            (group ? heads : singles).push_back((uint32_t)(j0 + j));
            for (int64_t m = 0; m < members; ++m, ++j) {
                if (m) jobs[j0 + j].flags |= kSegCont;
                if (m + 1 < members) jobs[j0 + j].flags |= kSegMore;
            }
        }
    }
    const uint32_t nj = (uint32_t)jobs.size();
    const uint64_t bstride = (uint64_t)((B + (B >> 3) + 64 + 15) & ~15ll);
    // device scratch, carved out of ctx->aux
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_tab = carve((size_t)nj * kHashEntries * 4), o_snap = carve((size_t)nj * kHashEntries * 4);
    const size_t o_jobs = carve((size_t)nj * sizeof(SegJob)), o_list = carve((size_t)nj * 4), o_perm = carve((size_t)nj * 4);
    const size_t o_ss = carve((size_t)nj * sizeof(SegState)), o_es = carve((size_t)nj * sizeof(SegState)), o_bad = carve(nj);
    const size_t o_poff = carve((size_t)nslots * 4), o_plen = carve((size_t)nslots * 4);
    const size_t o_sf = carve((size_t)n * 4), o_sc = carve((size_t)n * 4), o_buf = carve((size_t)n * bstride + 64);
    CKS(reserve(ctx, ctx->aux, off));
    uint8_t *A = (uint8_t *)ctx->aux.p;
    int32_t *d_tab = (int32_t *)(A + o_tab), *d_snap = (int32_t *)(A + o_snap);
    SegJob *d_jobs = (SegJob *)(A + o_jobs);
    uint32_t *d_list = (uint32_t *)(A + o_list), *d_perm = (uint32_t *)(A + o_perm);
    SegState *d_ss = (SegState *)(A + o_ss), *d_es = (SegState *)(A + o_es);
    uint8_t *d_bad = A + o_bad;
    uint32_t *d_poff = (uint32_t *)(A + o_poff), *d_plen = (uint32_t *)(A + o_plen), *d_sf = (uint32_t *)(A + o_sf), *d_sc = (uint32_t *)(A + o_sc);
    uint8_t *d_buf = A + o_buf;
    CK(cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)nj * sizeof(SegJob), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sf, slot_first.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sc, slot_count.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_plen, 0, (size_t)nslots * 4, st));
    uint32_t *counter = ctx->d_counter + 8;
    auto launch = [&](const uint32_t *list, uint32_t count, const uint32_t *land, uint32_t nbig, uint32_t follow, const uint8_t *chase) -> int {
        CK(cudaMemsetAsync(counter, 0, 12, st));
        const uint64_t cap = (uint64_t)ctx->sm_count * kSegCtasPerSm;
        int grid;
        uint32_t active;
        if (nbig) {     // warp 0 of every CTA takes the long segments, the others the short ones
            const uint64_t nsmall = count - nbig;
            grid = (int)std::min<uint64_t>(std::max<uint64_t>(nbig, (nsmall + kSegWarps - 2) / (kSegWarps - 1)), cap);
            active = (uint32_t)std::min<uint64_t>(1 + (nsmall + grid - 1) / grid, (uint64_t)kSegWarps);
        } else {
            grid = (int)std::min<uint64_t>(count, cap);
            active = (uint32_t)std::min<uint64_t>((count + grid - 1) / grid, (uint64_t)kSegWarps);
        }
        k_compress_segments<<<grid, kSegWarps * 32, kSegSmemBytes, st>>>(d_work, d_jobs, list, count, (int32_t)B, init_table, d_tab, d_snap,
                                                                                  d_ss, d_es, d_buf, bstride, d_poff, d_plen, counter, active, land,
                                                                                  land_origin, land_shift, nbig, follow, chase, nj, counter + 2);
        ctx->launches++;
        CK(cudaGetLastError());
        return DLZ4_OK;
    };
    const uint32_t nbig = (uint32_t)heads.size();
    std::vector<uint32_t> perm(heads);
    perm.insert(perm.end(), singles.begin(), singles.end());
    CK(cudaMemcpyAsync(d_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice, st));
    CKS(launch(d_perm, (uint32_t)perm.size(), landed, nbig, 1, nullptr));
    ctx->seg_jobs = nj; ctx->seg_reruns = 0; ctx->seg_rounds = 0;
    bool speculative = false;
    for (const SegJob &J : jobs) speculative |= !(J.flags & kSegFirst);
    if (speculative) {
        std::vector<uint8_t> bad(nj);
        std::vector<uint32_t> list;
        for (uint32_t round = 0; round <= nj; ++round) {
            k_seg_verify<<<nj, 256, 0, st>>>(d_jobs, nj, d_tab, d_snap, d_ss, d_es, d_bad);
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(bad.data(), d_bad, nj, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            list.clear();
            for (uint32_t j = 1; j < nj; ++j)
                if (bad[j] && !bad[j - 1]) list.push_back(j);          // predecessor's end state stands: re-run from it is exact
            if (list.empty()) break;
            for (uint32_t j : list) jobs[j].flags |= kSegRerun;
            CK(cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)nj * sizeof(SegJob), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, st));
            CKS(launch(d_list, (uint32_t)list.size(), nullptr, 0, 0, ctx->k_seg_no_chase ? nullptr : d_bad));
            uint32_t chased = 0;
            CK(cudaMemcpyAsync(&chased, counter + 2, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));                                // `list` / `jobs` are reused next round
            ctx->seg_reruns += (uint32_t)list.size() + chased;
            ctx->seg_rounds++;
        }
    }
    if (final_table && linked)      // the last segment's table buffer is the chain's end state (verified above)
        CK(cudaMemcpyAsync(final_table, d_tab + (size_t)(nj - 1) * kHashEntries, kHashEntries * 4, cudaMemcpyDeviceToDevice, st));
    {
        const uint32_t gx = (uint32_t)std::min<uint64_t>(n, 65535), gy = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(64, (uint64_t)ctx->sm_count * 8 / gx));
        k_seg_assemble<<<dim3(gx, gy), 256, 0, st>>>(d_buf, bstride, d_sf, d_sc, d_poff, d_plen, n, d_comp, d_coff, d_clen);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

// Jump decoder (k_jd_*): token scan of every block, then pointer doubling inside units of <= 16 MiB of consecutive blocks.
// Used for linked-block frames and for frames of few large blocks, where one warp per dependent stream would crawl.
// On return *total = decoded bytes and status_h[i] = per-block status (first non-zero one is the frame's error).
int decompress_jump(dlz4_ctx *ctx, const uint8_t *d_frame, uint64_t frame_span, const uint64_t *d_soff, const uint32_t *d_slen, const uint8_t *d_stored,
                    const std::vector<uint32_t> &slen, uint32_t n, uint32_t B, uint8_t *d_out, uint64_t cap_total, const uint8_t *d_dict,
                    uint32_t dwin, bool linked, uint32_t *d_olen, uint8_t *d_status, std::vector<uint8_t> &status_h, uint64_t *total,
                    cudaStream_t st, uint8_t *host_out, uint64_t host_cap, bool *copied_out,
                    const std::vector<uint32_t> *groups = nullptr /* the frame is still arriving: group g = blocks [(*groups)[g],
                                                                     (*groups)[g+1]) is in memory once event evp[kJdGroupEvent + g] fires */) {
    std::vector<uint64_t> seq_base(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) seq_base[i + 1] = seq_base[i] + slen[i] / 3 + B / 2048 + 8;
    // 16 MiB units; 32 MiB for frames of half a GiB and more (fewer launches; the first unit ships later, which only a long
    // frame can afford: 1 GiB 40.2 -> 38.2 ms, 128 MiB 5.8 -> 6.0 ms)
    uint32_t unit_bytes = (uint64_t)n * B >= (512ull << 20) ? 32u << 20 : 16u << 20;
    if (ctx->k_jd_unit_mib > 0) unit_bytes = (uint32_t)ctx->k_jd_unit_mib << 20;
    const uint32_t per_unit = std::max<uint32_t>(1u, unit_bytes / B);
    const uint32_t nunits = (n + per_unit - 1) / per_unit;
    int rounds = 1;
    while ((1ull << (kJdHopBits * rounds)) < (uint64_t)per_unit * B) ++rounds;    // chain depth <= unit bytes, / kJdHops per round
    ++rounds;
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_seq = carve((size_t)seq_base[n] * sizeof(JdSeq)), o_sb = carve((size_t)(n + 1) * 8), o_ns = carve((size_t)n * 4);
    const size_t o_reach = carve((size_t)n * 4), o_base = carve((size_t)(n + 1) * 8), o_P = carve((size_t)per_unit * B * 4);
    const size_t o_todo = carve((size_t)nunits * (rounds + 2) * 4);
    const size_t ntiles = ((size_t)per_unit * B + kJdTile - 1) / kJdTile, o_tile = carve(ntiles);
    // chunked (parallel) token scan for blocks > 64 KiB: per-byte next/advance, per-byte chunk exits, run and slow-token lists
    const bool chunked = B > 65536 && !ctx->k_jd_serial_scan;
    uint64_t fspan = 0;
    uint32_t max_slen = 0;
    std::vector<uint64_t> list_base(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) {
        max_slen = std::max(max_slen, slen[i]);
        list_base[i + 1] = list_base[i] + slen[i] / kJdpChunk + slen[i] / 64 + 8;
    }
    size_t o_nx = 0, o_adv = 0, o_ex = 0, o_lb = 0, o_runs = 0, o_slows = 0, o_nr = 0, o_nsl = 0, o_fb = 0;
    if (chunked) {
        fspan = frame_span;
        o_nx = carve((size_t)fspan * 2); o_adv = carve((size_t)fspan * 2); o_ex = carve((size_t)fspan * sizeof(JdpExit));
        o_lb = carve((size_t)(n + 1) * 8); o_runs = carve((size_t)list_base[n] * sizeof(JdpRun));
        o_slows = carve((size_t)list_base[n] * sizeof(JdpSlow)); o_nr = carve((size_t)n * 4); o_nsl = carve((size_t)n * 4); o_fb = carve(n);
    }
    CKS(reserve(ctx, ctx->aux, off));
    uint8_t *A = (uint8_t *)ctx->aux.p;
    JdSeq *d_seq = (JdSeq *)(A + o_seq);
    uint64_t *d_sb = (uint64_t *)(A + o_sb), *d_base = (uint64_t *)(A + o_base);
    uint32_t *d_ns = (uint32_t *)(A + o_ns), *d_reach = (uint32_t *)(A + o_reach), *d_todo = (uint32_t *)(A + o_todo);
    int32_t *d_P = (int32_t *)(A + o_P);
    uint8_t *d_tile = A + o_tile;
    CK(cudaMemcpyAsync(d_sb, seq_base.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_todo, 0, (size_t)nunits * (rounds + 2) * 4, st));
    CK(cudaMemsetAsync(d_reach, 0, (size_t)n * 4, st));
    CK(cudaMemsetAsync(ctx->d_counter, 0, 4, st));
    const int scan_grid = (int)std::min<uint64_t>((n + 3) / 4, (uint64_t)ctx->sm_count * 8);
    if (chunked) {
        uint16_t *d_nx = (uint16_t *)(A + o_nx), *d_adv = (uint16_t *)(A + o_adv);
        JdpExit *d_ex = (JdpExit *)(A + o_ex);
        uint64_t *d_lb = (uint64_t *)(A + o_lb);
        JdpRun *d_runs = (JdpRun *)(A + o_runs);
        JdpSlow *d_slows = (JdpSlow *)(A + o_slows);
        uint32_t *d_nr = (uint32_t *)(A + o_nr), *d_nsl = (uint32_t *)(A + o_nsl);
        uint8_t *d_fb = A + o_fb;
        CK(cudaMemcpyAsync(d_lb, list_base.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        const uint32_t ngroups = groups ? (uint32_t)groups->size() - 1 : 1;
        for (uint32_t gi = 0; gi < ngroups; ++gi) {
            // per-byte sizing and chunk exits need only a block's own bytes: start on every group of blocks as it lands
            const uint32_t g0 = groups ? (*groups)[gi] : 0, g1 = groups ? (*groups)[gi + 1] : n;
            if (groups) CK(cudaStreamWaitEvent(st, ctx->evp[kJdGroupEvent + gi], 0));
            k_jdp_next<<<dim3((max_slen + 255) / 256, g1 - g0), 256, 0, st>>>(d_frame, d_soff + g0, d_slen + g0, d_stored + g0, d_nx, d_adv);
            k_jdp_exit<<<dim3((max_slen + kJdpChunk - 1) / kJdpChunk, g1 - g0), 256, 0, st>>>(d_soff + g0, d_slen + g0, d_stored + g0, d_nx, d_adv, d_ex);
        }
        k_jdp_hop<<<(n + 7) / 8, 256, 0, st>>>(d_frame, d_soff, d_slen, d_stored, n, B, d_ex, d_runs, d_slows, d_lb, d_nr, d_nsl, d_ns, d_olen,
                                                d_reach, d_status, d_fb);
        k_jdp_emit<<<dim3(64, n), 256, 0, st>>>(d_frame, d_soff, d_slen, d_nx, d_adv, d_runs, d_slows, d_lb, d_nr, d_nsl, d_seq, d_sb, d_reach);
        // blocks the chunked scan refused (malformed input, overflow): the serial scan decides their status
        k_jd_scan<<<scan_grid, 128, 0, st>>>(d_frame, d_soff, d_slen, d_stored, n, B, d_seq, d_sb, d_ns, d_olen, d_reach, d_status,
                                             ctx->d_counter, d_fb);
        ctx->launches += 5;
        if (ctx->k_debug) {
            std::vector<uint32_t> nr(n), nsl(n); std::vector<uint8_t> fb(n);
            CK(cudaMemcpyAsync(nr.data(), d_nr, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(nsl.data(), d_nsl, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(fb.data(), d_fb, n, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint64_t a = 0, b2 = 0, c = 0; uint32_t ma = 0, mb = 0;
            for (uint32_t i = 0; i < n; ++i) { a += nr[i]; b2 += nsl[i]; c += fb[i]; ma = std::max(ma, nr[i]); mb = std::max(mb, nsl[i]); }
            fprintf(stderr, "dlz4: chunked scan: %u blocks, runs %llu (max %u per block), slow tokens %llu (max %u), fallback blocks %llu\n", n,
                    (unsigned long long)a, ma, (unsigned long long)b2, mb, (unsigned long long)c);
        }
    } else {
        k_jd_scan<<<scan_grid, 128, 0, st>>>(d_frame, d_soff, d_slen, d_stored, n, B, d_seq, d_sb, d_ns, d_olen, d_reach, d_status,
                                             ctx->d_counter, nullptr);
        ctx->launches++;
    }
    k_jd_bases<<<1, 32, 0, st>>>(d_olen, d_reach, n, dwin, linked ? 1 : 0, cap_total, d_base, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    status_h.assign(n, 0);
    std::vector<uint64_t> base_h(n + 1);
    CK(cudaMemcpyAsync(status_h.data(), d_status, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(base_h.data(), d_base, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *total = base_h[n];
    for (uint32_t i = 0; i < n; ++i)
        if (status_h[i]) return DLZ4_OK;                                  // the caller maps the first status to the error
    // a unit's bytes are final once its k_jd_emit ran: ship them while the later units resolve
    const bool ship = host_out && *total <= host_cap && nunits <= (uint32_t)kJdGroupEvent;
    if (copied_out) *copied_out = ship;
    const int wide = ctx->sm_count * 8;
    for (uint32_t u = 0; u < nunits; ++u) {
        const uint32_t b0 = u * per_unit, b1 = std::min<uint32_t>(n, b0 + per_unit);
        uint32_t *todo = d_todo + (size_t)u * (rounds + 2);
        const uint32_t cpb = std::max<uint32_t>(1u, 2048u / (b1 - b0));
        k_jd_fill<<<(b1 - b0) * cpb, 256, 0, st>>>(d_frame, d_soff, d_seq, d_sb, d_ns, d_base, b0, cpb, d_out, d_dict, dwin, linked ? 1 : 0, d_P, todo);
        CK(cudaMemsetAsync(d_tile, 1, ntiles, st));
        for (int r = 0; r < rounds; ++r) k_jd_round<<<wide, 256, 0, st>>>(d_P, d_base, b0, b1, todo + r, todo + r + 1, d_tile);
        k_jd_emit<<<wide, 256, 0, st>>>(d_P, d_base, b0, b1, d_out);
        ctx->launches += 2 + rounds;
        if (ship && base_h[b1] > base_h[b0]) {
            CK(cudaEventRecord(ctx->evp[u], st));
            CK(cudaStreamWaitEvent(ctx->copy_out, ctx->evp[u], 0));
            CKS(d2h(ctx, host_out + base_h[b0], d_out + base_h[b0], base_h[b1] - base_h[b0], ctx->copy_out));
        }
    }

    CK(cudaGetLastError());
    return DLZ4_OK;
}

int launch_chain(dlz4_ctx *ctx, const uint8_t *work, int32_t start, int32_t total, int32_t block, uint32_t nblocks,
                 int32_t *table_io, uint8_t *dst, uint64_t stride, uint32_t *comp_len, cudaStream_t st) {
    k_compress_chain<<<1, 32, kHashEntries * 4 + kRingBytes, st>>>(work, start, total, block, nblocks, table_io, dst, stride, comp_len);
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int dlz4_init(int device, dlz4_ctx **out) {
    if (!out) return DLZ4_E_INVALID_ARG;
    *out = nullptr;
    dlz4_ctx *ctx = new dlz4_ctx();
    ctx->device = device;
    *out = ctx;                       // returned even on failure so the caller can read last_error
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { ctx->last_error = "no such CUDA device"; return DLZ4_E_CUDA; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    for (cudaEvent_t &e : ctx->evp) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->ev0));
    CK(cudaEventCreate(&ctx->ev1));
    for (cudaEvent_t &e : ctx->evq) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    for (cudaStream_t &l : ctx->lanes) CK(cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking));
    // every environment knob is read here, once (the reference has no configuration surface: these exist for A/B measurements)
    if (const char *e = getenv("DLZ4_SEG_KIB")) ctx->k_seg_kib = std::max(0, atoi(e));
    if (const char *e = getenv("DLZ4_SEG_WARM_KIB")) ctx->k_seg_warm_kib = std::max(0, atoi(e));
    if (const char *e = getenv("DLZ4_SEG_UNIT_KIB")) ctx->k_seg_unit_kib = std::max(0, atoi(e));
    if (const char *e = getenv("DLZ4_SEG_GROUP")) ctx->k_seg_group = std::max(0, atoi(e));
    ctx->k_seg_no_chase = getenv("DLZ4_SEG_NO_CHASE") != nullptr;
    if (const char *e = getenv("DLZ4_JD_UNIT_MIB")) ctx->k_jd_unit_mib = std::max(1, atoi(e));
    ctx->k_jd_serial_scan = getenv("DLZ4_JD_SERIAL_SCAN") != nullptr;
    ctx->k_jd_no_overlap = getenv("DLZ4_JD_NO_OVERLAP") != nullptr;
    ctx->k_debug = getenv("DLZ4_DEBUG") != nullptr;
    if (const char *e = getenv("DLZ4_CHUNK_MIB")) ctx->chunk_bytes = (uint64_t)std::max(1, atoi(e)) << 20;
    if (const char *e = getenv("DLZ4_LANES")) ctx->n_lanes = std::min(4, std::max(1, atoi(e)));
    if (const char *e = getenv("DLZ4_COPY_THREADS")) ctx->copy_threads = std::min(64, std::max(1, atoi(e)));
    CK(cudaMalloc(&ctx->d_counter, 64));
    CK(cudaMalloc(&ctx->d_land, kLandFlags * 4));
    CK(cudaMallocHost(&ctx->h_one, 64));
    *ctx->h_one = 1u;
    if (const char *e = getenv("DLZ4_SEG_OVERLAP_MIN_MIB")) ctx->seg_overlap_min_bytes = (uint64_t)atoll(e) << 20;   // huge: never
    CK(cudaMalloc(&ctx->d_hash, 64));
    CK(cudaMalloc(&ctx->d_total, 64));
    CK(cudaMalloc(&ctx->d_table, kHashEntries * sizeof(int32_t)));
    CK(cudaFuncSetAttribute(k_compress_fresh16<kWarpsFresh16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            kWarpsFresh16 * (kHashEntries * 2 + kRingBytes)));
    CK(cudaFuncSetAttribute(k_compress_fresh16<kWarpsFresh16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            kWarpsFresh16 * (kHashEntries * 2 + kRingBytes)));
    if (const char *e = getenv("DLZ4_SEG_MIN_KIB")) ctx->seg_min_bytes = (uint64_t)atoll(e) << 10;     // huge value: serial chain only
    if (const char *e = getenv("DLZ4_JUMP_MIN_KIB")) ctx->jump_min_bytes = (uint64_t)atoll(e) << 10;
    if (const char *e = getenv("DLZ4_WIDE")) ctx->wide = atoi(e) != 0;
    if (const char *e = getenv("DLZ4_SPLIT")) ctx->split = atoi(e) != 0;
    if (const char *e = getenv("DLZ4_PW")) ctx->pw = std::max(0, std::min(3, atoi(e)));
    if (const char *e = getenv("DLZ4_PW_FUSED")) ctx->pw_fused = atoi(e) != 0;
    if (const char *e = getenv("DLZ4_PW_SLEEP")) ctx->pw_sleep = std::max(0, atoi(e));
    if (const char *e = getenv("DLZ4_PW_LEAD")) ctx->pw_lead = std::max(3, std::min(kPwWin - 1, atoi(e)));
    if (const char *e = getenv("DLZ4_HYBRID")) ctx->hybrid = atoi(e) != 0;      // 0: the 7-warp shared-memory-only kernel (A/B runs)
    ctx->hy_grid = ctx->sm_count * kHyCtasPerSm;
    if (const char *e = getenv("DLZ4_HY_ACTIVE")) ctx->hy_active = std::max(0, std::min(kHyWarps, atoi(e)));
    CK(cudaMalloc(&ctx->d_gtabs, (size_t)kGtabRegions * ctx->hy_grid * kHyGlWarps * kHashEntries * 2));
    CK(cudaMalloc(&ctx->d_nrec, (size_t)kGtabRegions * kMaxSplitBlocks * 4));
    CK(cudaFuncSetAttribute(k_parse_fresh16<kWarpsFresh16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsFresh16 * kHashEntries * 2));
    CK(cudaFuncSetAttribute(k_parse_pw<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwChains * kPwChainBytes));
    CK(cudaFuncSetAttribute(k_parse_pw<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwChains * kPwChainBytes));
    CK(cudaFuncSetAttribute(k_parse_pw<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwChains * kPwChainBytes));
    CK(cudaFuncSetAttribute(k_parse_pw<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwChains * kPwChainBytes));
    CK(cudaFuncSetAttribute(k_compress_fresh16h, cudaFuncAttributeMaxDynamicSharedMemorySize, kHySmemBytes));
    CK(cudaFuncSetAttribute(k_compress_overlay, cudaFuncAttributeMaxDynamicSharedMemorySize, kHySmemBytes));
    CK(cudaFuncSetAttribute(k_compress_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegSmemBytes));
    CK(cudaFuncSetAttribute(k_compress_generic32<kWarpsGeneric32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            kWarpsGeneric32 * (kHashEntries * 4 + kRingBytes)));
    CK(cudaFuncSetAttribute(k_compress_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, kHashEntries * 4 + kRingBytes));
    return DLZ4_OK;
}

void dlz4_shutdown(dlz4_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->side) cudaStreamSynchronize(ctx->side);
    for (Buf *b : {&ctx->work, &ctx->comp, &ctx->seg, &ctx->out, &ctx->meta, &ctx->aux})
        if (b->p) cudaFree(b->p);
    if (ctx->pin.p) cudaFreeHost(ctx->pin.p);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_land) cudaFree(ctx->d_land);
    if (ctx->h_one) cudaFreeHost(ctx->h_one);
    if (ctx->d_hash) cudaFree(ctx->d_hash);
    if (ctx->d_total) cudaFree(ctx->d_total);
    if (ctx->d_table) cudaFree(ctx->d_table);
    if (ctx->d_gtabs) cudaFree(ctx->d_gtabs);
    if (ctx->d_nrec) cudaFree(ctx->d_nrec);
    if (ctx->rec.p) cudaFree(ctx->rec.p);
    delete ctx->pool;
    for (cudaStream_t s : ctx->sum_stream) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (ctx->h_sum) cudaFreeHost(ctx->h_sum);
    for (int i = 0; i < kStageSlots; ++i) {
        if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->evq) if (e) cudaEventDestroy(e);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (cudaEvent_t e : ctx->evp) if (e) cudaEventDestroy(e);
    for (cudaStream_t l : ctx->lanes) if (l) { cudaStreamSynchronize(l); cudaStreamDestroy(l); }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    delete ctx;
}

const char *dlz4_strerror(int status) {
    switch (status) {
        case DLZ4_OK: return "ok";
        case DLZ4_E_OUTPUT_TOO_SMALL: return "LZ4: Output Buffer Too Small";
        case DLZ4_E_MALFORMED: return "LZ4: Malformed Input";
        case DLZ4_E_OFFSET_ZERO: return "LZ4: Invalid Offset 0";
        case DLZ4_E_DICT_OOB: return "LZ4: Dictionary Offset Out of Bounds";
        case DLZ4_E_BAD_MAGIC: return "LZ4: Invalid Magic Number";
        case DLZ4_E_BAD_VERSION: return "LZ4: Unsupported Version";
        case DLZ4_E_CONTENT_CHECKSUM: return "LZ4: Content Checksum Error";
        case DLZ4_E_BLOCK_CHECKSUM: return "LZ4: Block Checksum Error";
        case DLZ4_E_HEADER_CHECKSUM: return "LZ4: Header Checksum Error";
        case DLZ4_E_INVALID_ARG: return "dlz4: invalid argument";
        case DLZ4_E_TOO_LARGE: return "dlz4: input of 2 GiB or more in one call";
        case DLZ4_E_CUDA: return "dlz4: CUDA error";
        default: return "dlz4: unknown status";
    }
}

const char *dlz4_last_error(const dlz4_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t dlz4_launch_count(const dlz4_ctx *ctx) { return ctx ? ctx->launches : 0; }
float dlz4_last_kernel_ms(const dlz4_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }
int dlz4_kernel_probe(dlz4_ctx *ctx, int enable, float *finder_ms, float *encoder_ms) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    if (finder_ms || encoder_ms) {
        // the most recent probed batch of fresh blocks <= 64 KiB (match finder k_parse_pw, then encoder k_encode_blocks)
        CK(cudaSetDevice(ctx->device));
        CK(cudaEventSynchronize(ctx->evq[2]));
        if (finder_ms) CK(cudaEventElapsedTime(finder_ms, ctx->evq[0], ctx->evq[1]));
        if (encoder_ms) CK(cudaEventElapsedTime(encoder_ms, ctx->evq[1], ctx->evq[2]));
    }
    ctx->probe = enable != 0;
    return DLZ4_OK;
}
void dlz4_segment_stats(const dlz4_ctx *ctx, uint32_t *segments, uint32_t *reruns, uint32_t *rounds) {
    if (segments) *segments = ctx ? ctx->seg_jobs : 0;
    if (reruns) *reruns = ctx ? ctx->seg_reruns : 0;
    if (rounds) *rounds = ctx ? ctx->seg_rounds : 0;
}

uint64_t dlz4_compress_bound(uint64_t n) { return n + n / 255 + 16; }
uint64_t dlz4_frame_bound(uint64_t n) { return 19 + n + (n / 65536 + 1) * 8 + 8 + 64; }

void dlz4_shard_range(uint64_t nblocks, uint32_t world, uint32_t rank, uint64_t *first, uint64_t *count) {
    if (world == 0) world = 1;
    // block i belongs to rank floor(i * world / nblocks): rank r owns [ceil(r*n/w), ceil((r+1)*n/w))
    const uint64_t lo = ((uint64_t)rank * nblocks + world - 1) / world;
    const uint64_t hi = ((uint64_t)(rank + 1) * nblocks + world - 1) / world;
    if (first) *first = lo;
    if (count) *count = hi - lo;
}

#ifdef DLZ4_PHASE_TIMING
// profiling build only: read and clear the per-phase cycle counters
extern "C" int dlz4_phase_counters(unsigned long long *out16) {
    if (cudaMemcpyFromSymbol(out16, dlz4::g_phase, 16 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(dlz4::g_phase, z, sizeof z) != cudaSuccess;
}
#endif

// ---- pinned host memory for callers that want the full PCIe rate (N-API external ArrayBuffers) ---------
void *dlz4_pinned_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void dlz4_pinned_free(void *p) { if (p) cudaFreeHost(p); }
// Page-locks memory the caller already owns (a Node Buffer, a shared mapping the ranks of a box all map): portable, so
// every context of the process sees it as pinned.  Returns DLZ4_OK, or DLZ4_E_CUDA when the driver refuses (the memory
// stays usable, just pageable).
int dlz4_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return DLZ4_E_INVALID_ARG;
    return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess ? DLZ4_OK : (cudaGetLastError(), DLZ4_E_CUDA);
}
int dlz4_host_unregister(void *p) {
    if (!p) return DLZ4_E_INVALID_ARG;
    return cudaHostUnregister(p) == cudaSuccess ? DLZ4_OK : (cudaGetLastError(), DLZ4_E_CUDA);
}

// ---- batched raw blocks ---------------------------------------------------------------------------------
int dlz4_compress_blocks_dev(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len,
                                uint32_t nblocks, uint32_t max_block_len, const uint8_t *prefix, uint32_t prefix_len, int warm,
                                const int32_t *init_table, uint8_t *dst, const uint64_t *dst_off, uint32_t *comp_len,
                                void *stream) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    const int32_t *table = nullptr;
    if (warm == DLZ4_WARM_TABLE) {
        if (!init_table) return DLZ4_E_INVALID_ARG;
        table = init_table;
    } else if (warm == DLZ4_WARM_JENKINS && prefix_len >= 4) {
        CK(cudaMemsetAsync(ctx->d_table, 0, kHashEntries * sizeof(int32_t), st));
        const int n = (int)prefix_len - 3;
        k_warm_jenkins<<<(n + 255) / 256, 256, 0, st>>>(prefix, (int32_t)prefix_len, ctx->d_table);
        ctx->launches++;
        CK(cudaGetLastError());
        table = ctx->d_table;
    } else if (warm != DLZ4_WARM_NONE && warm != DLZ4_WARM_JENKINS) {
        return DLZ4_E_INVALID_ARG;
    }
    return launch_compress(ctx, src, src_off, src_len, nblocks, max_block_len, prefix, prefix_len, table, dst, dst_off, comp_len, st);
}

// Chunked host pipeline: H2D of later chunks, kernels of up to n_lanes chunks and D2H of finished chunks overlap (one copy-in
// stream, n_lanes compute streams with a work-queue counter each, one copy-out stream).  128 MiB chunks measured best:
// the drain is bounded below by the latency of one block chain (~3 ms for a 64 KiB text block) whatever the chunk size,
// and smaller chunks only add launches (profiles/r01_e2e_chunk_sweep.txt).
// Blocks must be ascending and non-overlapping in `src`; output is packed (block i directly after block i-1).
// Error exits of the chunked pipelines: copies into the caller's buffers and kernels on the lanes may still be in flight
// (a later chunk's D2H, queued launches).  Nothing of this call may outlive it -- the caller frees or reuses its buffers and
// the next call reuses the event pool and the scratch -- so every exit that is not the normal one drains all streams first.
struct PipeGuard {
    dlz4_ctx *ctx;
    bool armed = true;
    ~PipeGuard() {
        if (!armed) return;
        cudaStreamSynchronize(ctx->copy_in);
        cudaStreamSynchronize(ctx->copy_out);
        for (int l = 0; l < 4; ++l) if (ctx->lanes[l]) cudaStreamSynchronize(ctx->lanes[l]);
    }
};
static const uint32_t kMaxChunks = 56;

// frame_mode: the packed stream is the body of an LZ4 frame -- [u32 size | stored bit][payload][u32 xxh32]* with the
// stored-block rule of bufferCompress.js:221-231 -- instead of bare compressed blocks; *total_out = its length.
static int compress_blocks_packed(dlz4_ctx *ctx, const uint8_t *src, uint64_t src_bytes, const uint64_t *src_off,
                                  const uint32_t *src_len, uint32_t n, uint32_t max_len, uint8_t *dst, uint64_t dst_bytes,
                                  uint32_t *comp_len, int frame_mode = 0, int block_checksum = 0, uint64_t *total_out = nullptr) {
    const uint64_t stride = (dlz4_compress_bound(max_len) + 15) & ~15ull;
    const uint64_t src_pad = (src_bytes + 15) & ~(uint64_t)15;
    CKS(reserve(ctx, ctx->work, src_pad + 16));
    CKS(reserve(ctx, ctx->comp, (uint64_t)n * stride + 64));
    CKS(reserve(ctx, ctx->seg, (uint64_t)n * (stride + 16) + 64));
    CKS(reserve(ctx, ctx->meta, (size_t)n * (8 + 8 + 4 + 4) + ((size_t)n + 64) * 8 + 256));
    CKS(reserve(ctx, ctx->aux, (size_t)n * 12 + 256));                   // frame mode: payload offsets / lengths for the block checksums
    uint64_t *d_doff = (uint64_t *)ctx->aux.p;
    uint32_t *d_dlen = (uint32_t *)(d_doff + n);
    CKS(reserve_pinned(ctx, ctx->pin, 64 * 8 + (size_t)n * 4));      // chunk totals + comp_len staging (pageable D2H would block)
    uint8_t *d_src = (uint8_t *)ctx->work.p, *d_comp = (uint8_t *)ctx->comp.p, *d_pack = (uint8_t *)ctx->seg.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_coff = d_soff + n, *d_pos = d_coff + n;     // d_pos: per chunk n_c + 1 entries
    uint32_t *d_slen = (uint32_t *)(d_pos + n + 64), *d_clen = d_slen + n;
    volatile uint64_t *h_tot = (volatile uint64_t *)ctx->pin.p;
    uint32_t *h_clen = (uint32_t *)((uint8_t *)ctx->pin.p + 64 * 8);
    cudaStream_t *sks = ctx->lanes, si = ctx->copy_in, so = ctx->copy_out;
    const uint32_t nl = (uint32_t)ctx->n_lanes;
    cudaStream_t sk = sks[0];
    PipeGuard guard{ctx};

    // chunk boundaries by source bytes
    std::vector<uint32_t> cb{0};
    uint64_t total_src = 0;
    for (uint32_t i = 0; i < n; ++i) total_src += src_len[i];
    const uint64_t target = std::max<uint64_t>(ctx->chunk_bytes, total_src / kMaxChunks + 1);
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        acc += src_len[i];
        if (acc >= target || i + 1 == n) { cb.push_back(i + 1); acc = 0; }      // (small first chunks measured worse here: 26.7 vs 25.4 ms)
    }
    const uint32_t nc = (uint32_t)cb.size() - 1;            // <= kMaxChunks + 1
    // descriptors once (compressed scratch is worst-case strided)
    std::vector<uint64_t> coff(n);
    for (uint32_t i = 0; i < n; ++i) coff[i] = (uint64_t)i * stride;
    CK(cudaMemcpyAsync(d_soff, src_off, (size_t)n * 8, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_coff, coff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_slen, src_len, (size_t)n * 4, cudaMemcpyHostToDevice, sk));
    CK(cudaEventRecord(ctx->ev0, sk));
    CK(cudaEventRecord(ctx->ev_fork, sk));
    for (uint32_t l = 1; l < nl; ++l) CK(cudaStreamWaitEvent(sks[l], ctx->ev_fork, 0));   // descriptors visible to every lane

    uint64_t host_pos = 0;
    const int src_pg = src_bytes >= (4u << 20) && is_pageable(src), dst_pg = dst_bytes >= (4u << 20) && is_pageable(dst);
    auto drain = [&](uint32_t c) -> int {          // chunk c's kernels are done: ship its packed bytes
        CK(cudaEventSynchronize(ctx->evp[64 + c]));
        const uint64_t tot = h_tot[c];
        if (host_pos + tot > dst_bytes) return DLZ4_E_OUTPUT_TOO_SMALL;
        CKS(d2h(ctx, dst + host_pos, d_pack + (uint64_t)cb[c] * (stride + 16), tot, so, dst_pg));
        host_pos += tot;
        return DLZ4_OK;
    };
    // everything is enqueued up front (the host only blocks in drain(), which needs each chunk's packed size to place it):
    // a chunk's H2D copy on the copy-in stream (a pageable source through the staging ring: the host copies chunk c + 1 while
    // chunk c's kernels run), then its kernels on its lane, gated by the chunk's copy event
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t b0 = cb[c], b1 = cb[c + 1], m = b1 - b0;
        {
            const uint64_t lo = src_off[b0], hi = src_off[b1 - 1] + src_len[b1 - 1];
            if (hi > lo) CKS(h2d(ctx, d_src + lo, src + lo, hi - lo, si, src_pg));
            CK(cudaEventRecord(ctx->evp[c], si));
        }
        // chunks rotate over the compute streams (own work-queue counter and L2-table region each): the next chunks' CTAs
        // take over SM slots as the previous chunk's last blocks drain
        sk = sks[c % nl];
        CK(cudaStreamWaitEvent(sk, ctx->evp[c], 0));
        CKS(launch_compress(ctx, d_src, d_soff + b0, d_slen + b0, m, max_len, nullptr, 0, nullptr, d_comp, d_coff + b0, d_clen + b0, sk,
                            ctx->d_counter + 1 + (c % nl), nc > 1));
        uint64_t *pos = d_pos + b0 + c;                                   // m + 1 entries
        uint8_t *pack_c = d_pack + (uint64_t)b0 * (stride + 16);
        k_frame_layout<<<1, 1024, 0, sk>>>(d_slen + b0, d_clen + b0, m, block_checksum, pos, frame_mode ? d_doff + b0 : nullptr,
                                           frame_mode ? d_dlen + b0 : nullptr, frame_mode ? 0 : 1);
        k_frame_gather<<<(int)std::min<uint64_t>(m, (uint64_t)ctx->sm_count * 8), 256, 0, sk>>>(
            d_src, d_soff + b0, d_slen + b0, d_comp, d_coff + b0, d_clen + b0, m, pos, pack_c, frame_mode ? 0 : 1);
        ctx->launches += 2;
        CK(cudaGetLastError());
        if (frame_mode && block_checksum) CKS(launch_xxh32_batch(ctx, pack_c, d_doff + b0, d_dlen + b0, m, 0, nullptr, pack_c, sk));
        CK(cudaMemcpyAsync((void *)(h_tot + c), pos + m, 8, cudaMemcpyDeviceToHost, sk));
        if (comp_len) CK(cudaMemcpyAsync(h_clen + b0, d_clen + b0, (size_t)m * 4, cudaMemcpyDeviceToHost, sk));
        CK(cudaEventRecord(ctx->evp[64 + c], sk));
    }
    for (uint32_t c = 0; c < nc; ++c) CKS(drain(c));
    CK(cudaStreamSynchronize(so));
    for (uint32_t l = 1; l < nl; ++l) { CK(cudaEventRecord(ctx->ev_side, sks[l])); CK(cudaStreamWaitEvent(sks[0], ctx->ev_side, 0)); }
    CK(cudaEventRecord(ctx->ev1, sks[0]));
    CK(cudaStreamSynchronize(sks[0]));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    guard.armed = false;
    if (total_out) *total_out = host_pos;
    if (comp_len) {
        memcpy(comp_len, h_clen, (size_t)n * 4);
        for (uint32_t i = 0; i < n; ++i)
            if (comp_len[i] == 0xFFFFFFFFu) return DLZ4_E_INVALID_ARG;
    }
    return DLZ4_OK;
}

int dlz4_compress_blocks(dlz4_ctx *ctx, const uint8_t *src, uint64_t src_bytes, const uint64_t *src_off, const uint32_t *src_len,
                         uint32_t nblocks, const uint8_t *prefix, uint32_t prefix_len, int warm, const int32_t *init_table,
                         uint8_t *dst, uint64_t dst_bytes, const uint64_t *dst_off, uint32_t *comp_len) {
    if (!ctx || (nblocks && (!src_off || !src_len || !comp_len || !dst))) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint32_t max_len = 0;
    if (!dst_off) {
        // packed output: block i directly follows block i-1 in dst (offsets = running sum of comp_len)
        if (prefix_len || warm != DLZ4_WARM_NONE) return DLZ4_E_INVALID_ARG;
        for (uint32_t i = 0; i < nblocks; ++i) {
            if (src_off[i] + src_len[i] > src_bytes) return DLZ4_E_INVALID_ARG;
            if (i && src_off[i] < src_off[i - 1] + src_len[i - 1]) return DLZ4_E_INVALID_ARG;     // ascending, disjoint
            if (src_len[i] > max_len) max_len = src_len[i];
            if (src_len[i] > 0x7FFFFFF0u) return DLZ4_E_TOO_LARGE;
        }
        if (!nblocks) return DLZ4_OK;
        return compress_blocks_packed(ctx, src, src_bytes, src_off, src_len, nblocks, max_len, dst, dst_bytes, comp_len);
    }
    for (uint32_t i = 0; i < nblocks; ++i) {
        if (src_off[i] + src_len[i] > src_bytes) return DLZ4_E_INVALID_ARG;
        if (dst_off[i] + dlz4_compress_bound(src_len[i]) > dst_bytes) return DLZ4_E_OUTPUT_TOO_SMALL;
        if (src_len[i] > max_len) max_len = src_len[i];
        if (src_len[i] > 0x7FFFFFF0u - prefix_len) return DLZ4_E_TOO_LARGE;
    }
    const size_t meta_bytes = (size_t)nblocks * (8 + 4 + 8 + 4) + 64;
    const uint64_t src_pad = (src_bytes + 15) & ~(uint64_t)15;
    CKS(reserve(ctx, ctx->work, src_pad + prefix_len + 16));
    CKS(reserve(ctx, ctx->comp, dst_bytes));
    CKS(reserve(ctx, ctx->meta, meta_bytes));
    uint8_t *d_src = (uint8_t *)ctx->work.p;
    uint8_t *d_prefix = d_src + src_pad;
    uint8_t *d_dst = (uint8_t *)ctx->comp.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p;
    uint64_t *d_doff = d_soff + nblocks;
    uint32_t *d_slen = (uint32_t *)(d_doff + nblocks);
    uint32_t *d_clen = d_slen + nblocks;
    CK(cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, st));
    if (prefix_len) CK(cudaMemcpyAsync(d_prefix, prefix, prefix_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_soff, src_off, (size_t)nblocks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_doff, dst_off, (size_t)nblocks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_slen, src_len, (size_t)nblocks * 4, cudaMemcpyHostToDevice, st));
    const int32_t *d_init = nullptr;
    if (warm == DLZ4_WARM_TABLE) {
        if (!init_table) return DLZ4_E_INVALID_ARG;
        CKS(reserve(ctx, ctx->aux, kHashEntries * 4));
        CK(cudaMemcpyAsync(ctx->aux.p, init_table, kHashEntries * 4, cudaMemcpyHostToDevice, st));
        d_init = (const int32_t *)ctx->aux.p;
    }
    CK(cudaEventRecord(ctx->ev0, st));
    CKS(dlz4_compress_blocks_dev(ctx, d_src, d_soff, d_slen, nblocks, max_len, prefix_len ? d_prefix : nullptr, prefix_len, warm,
                                    d_init, d_dst, d_doff, d_clen, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(comp_len, d_clen, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // bring back only the bytes each block produced
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < nblocks; ++i) {
        if (comp_len[i] == 0xFFFFFFFFu) return DLZ4_E_INVALID_ARG;
        lo = std::min<uint64_t>(lo, dst_off[i]);
        hi = std::max<uint64_t>(hi, dst_off[i] + comp_len[i]);
    }
    if (nblocks && hi > lo) CK(cudaMemcpyAsync(dst + lo, d_dst + lo, hi - lo, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return DLZ4_OK;
}

// Decode side of the chunked pipeline: packed compressed input (block i directly after block i-1), ascending disjoint outputs.
// src_off_in (nullable): the blocks' positions in `src` when they are not back to back (blocks of a frame: size words and
// checksums lie between them; ascending); stored_in (nullable): 1 = the block is stored raw (bufferDecompress.js:147-149).
static int decompress_blocks_packed(dlz4_ctx *ctx, const uint8_t *src, uint64_t src_bytes, const uint32_t *src_len, uint32_t n,
                                    uint8_t *dst, uint64_t dst_bytes, const uint64_t *dst_off, const uint32_t *dst_cap,
                                    const uint8_t *dict, uint32_t dict_len, int hist_mode, uint32_t *out_len, uint8_t *status,
                                    const uint64_t *src_off_in = nullptr, const uint8_t *stored_in = nullptr) {
    std::vector<uint64_t> soff(n + 1, 0);
    uint64_t total_out = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (src_off_in) soff[i] = src_off_in[i];
        soff[i + 1] = soff[i] + src_len[i];
        total_out += dst_cap[i];
    }
    if (soff[n] > src_bytes) return DLZ4_E_INVALID_ARG;
    const uint64_t src_pad = ((src_off_in ? src_bytes : soff[n]) + 15) & ~(uint64_t)15;
    CKS(reserve(ctx, ctx->work, src_pad + dict_len + 32));
    CKS(reserve(ctx, ctx->out, dst_bytes + 64));
    CKS(reserve(ctx, ctx->meta, (size_t)n * (8 + 8 + 4 + 4 + 4 + 1 + 1) + 64));
    uint8_t *d_src = (uint8_t *)ctx->work.p, *d_dict = d_src + src_pad, *d_dst = (uint8_t *)ctx->out.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_doff = d_soff + n;
    uint32_t *d_slen = (uint32_t *)(d_doff + n), *d_cap = d_slen + n, *d_olen = d_cap + n;
    uint8_t *d_status = (uint8_t *)(d_olen + n), *d_stored = d_status + n;
    cudaStream_t *sks = ctx->lanes, si = ctx->copy_in, so = ctx->copy_out;
    const uint32_t nl = (uint32_t)ctx->n_lanes;
    cudaStream_t sk = sks[0];
    PipeGuard guard{ctx};
    const uint64_t target = std::max<uint64_t>(ctx->chunk_bytes, total_out / kMaxChunks + 1);
    std::vector<uint32_t> cb{0};
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        acc += dst_cap[i];
        // the first chunks are small (16, 32, 64 MiB ...): the copy back (the bottleneck of the decode side) starts early
        const uint64_t want = std::min<uint64_t>(target, (16ull << 20) << std::min<size_t>(cb.size() - 1, 8));
        if (acc >= want || i + 1 == n) { cb.push_back(i + 1); acc = 0; }
    }
    const uint32_t nc = (uint32_t)cb.size() - 1;
    const int src_pg = src_bytes >= (4u << 20) && is_pageable(src), dst_pg = dst_bytes >= (4u << 20) && is_pageable(dst);
    if (dict_len) CK(cudaMemcpyAsync(d_dict, dict, dict_len, cudaMemcpyHostToDevice, sk));
    if (stored_in) CK(cudaMemcpyAsync(d_stored, stored_in, n, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_soff, soff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_doff, dst_off, (size_t)n * 8, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_slen, src_len, (size_t)n * 4, cudaMemcpyHostToDevice, sk));
    CK(cudaMemcpyAsync(d_cap, dst_cap, (size_t)n * 4, cudaMemcpyHostToDevice, sk));
    CK(cudaEventRecord(ctx->ev0, sk));
    CK(cudaEventRecord(ctx->ev_fork, sk));
    for (uint32_t l = 1; l < nl; ++l) CK(cudaStreamWaitEvent(sks[l], ctx->ev_fork, 0));
    // frame-history mode: a block may read the previous blocks' output, so chunks must run in order on one stream
    const uint32_t lanes_used = hist_mode == DLZ4_HIST_FRAME ? 1u : nl;
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t b0 = cb[c], b1 = cb[c + 1], m = b1 - b0;
        const uint64_t s_lo = soff[b0], s_hi = soff[b1 - 1] + src_len[b1 - 1];
        if (s_hi > s_lo) CKS(h2d(ctx, d_src + s_lo, src + s_lo, s_hi - s_lo, si, src_pg));
        CK(cudaEventRecord(ctx->evp[c], si));
        cudaStream_t sc = sks[c % lanes_used];
        CK(cudaStreamWaitEvent(sc, ctx->evp[c], 0));
        CKS(launch_decompress(ctx, d_src, d_soff + b0, d_slen + b0, m, d_dst, d_doff + b0, d_cap + b0, dict_len ? d_dict : nullptr, dict_len,
                              hist_mode == DLZ4_HIST_FRAME, stored_in ? d_stored + b0 : nullptr, d_olen + b0, d_status + b0, sc,
                              ctx->d_counter + 1 + (c % lanes_used)));
        CK(cudaEventRecord(ctx->evp[64 + c], sc));
        if (dst_pg) continue;                                             // copied out below, chunk by chunk, once everything is queued
        CK(cudaStreamWaitEvent(so, ctx->evp[64 + c], 0));
        const uint64_t lo = dst_off[b0], hi = dst_off[b1 - 1] + dst_cap[b1 - 1];
        if (hi > lo) CK(cudaMemcpyAsync(dst + lo, d_dst + lo, hi - lo, cudaMemcpyDeviceToHost, so));
    }
    for (uint32_t l = 1; l < lanes_used; ++l) { CK(cudaEventRecord(ctx->ev_side, sks[l])); CK(cudaStreamWaitEvent(sk, ctx->ev_side, 0)); }
    CK(cudaEventRecord(ctx->ev1, sk));
    if (dst_pg) {
        // pageable destination: through the staging ring (a direct copy would block the host inside the loop above)
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t b0 = cb[c], b1 = cb[c + 1];
            CK(cudaStreamWaitEvent(so, ctx->evp[64 + c], 0));
            const uint64_t lo = dst_off[b0], hi = dst_off[b1 - 1] + dst_cap[b1 - 1];
            if (hi > lo) CKS(d2h(ctx, dst + lo, d_dst + lo, hi - lo, so, 1));
        }
    }
    CK(cudaMemcpyAsync(out_len, d_olen, (size_t)n * 4, cudaMemcpyDeviceToHost, sk));
    CK(cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, sk));
    CK(cudaStreamSynchronize(sk));
    CK(cudaStreamSynchronize(so));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    guard.armed = false;
    for (uint32_t i = 0; i < n; ++i)
        if (status[i]) return status[i];
    return DLZ4_OK;
}

// Device-resident batch decode: one warp per block, or the jump decoder for few large blocks.
static int decompress_dev_routed(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len, uint32_t nblocks,
                                 uint8_t *dst, const uint64_t *dst_off, const uint32_t *dst_cap, const uint8_t *dict, uint32_t dict_len,
                                 int hist_mode, uint32_t *out_len, uint8_t *status, cudaStream_t st) {
    if (nblocks && nblocks <= 65536 && hist_mode != DLZ4_HIST_FRAME && !dict_len) {
        // few large blocks: one warp per block decodes a 4 MiB block at ~40 MB/s.  If the batch is the usual one -- uniform blocks
        // > 64 KiB tiling one output range -- the jump decoder takes it (token scan and pointer doubling in parallel inside
        // every block).  The descriptors live on the device: read them back.
        std::vector<uint64_t> so(nblocks), doff(nblocks);
        std::vector<uint32_t> sl(nblocks), cap(nblocks);
        CK(cudaMemcpyAsync(cap.data(), dst_cap, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(doff.data(), dst_off, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint64_t B = cap[0];
        bool uniform = B > 65536 && B <= 4194304 && (B & (B - 1)) == 0;
        for (uint32_t i = 0; uniform && i < nblocks; ++i)
            uniform = doff[i] == doff[0] + (uint64_t)i * B && (cap[i] == B || (i + 1 == nblocks && cap[i] <= B && cap[i] > 0));
        if (uniform) {
            CK(cudaMemcpyAsync(so.data(), src_off, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(sl.data(), src_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint64_t span = 0;
            for (uint32_t i = 0; i < nblocks; ++i) { span = std::max<uint64_t>(span, so[i] + sl[i]); uniform = uniform && sl[i] > 0; }
            if (uniform && span < 0x7FFF0000ull) {
                CKS(reserve(ctx, ctx->seg, (size_t)nblocks + 256));        // (meta holds the host variant's descriptors, aux the decoder's scratch)
                uint8_t *d_stored = (uint8_t *)ctx->seg.p;
                CK(cudaMemsetAsync(d_stored, 0, nblocks, st));
                const uint64_t cap_total = (uint64_t)(nblocks - 1) * B + cap[nblocks - 1];
                std::vector<uint8_t> status_h;
                uint64_t total = 0;
                CKS(decompress_jump(ctx, src, span, src_off, src_len, d_stored, sl, nblocks, (uint32_t)B, dst + doff[0], cap_total, nullptr, 0,
                                    false, out_len, status, status_h, &total, st, nullptr, 0, nullptr));
                bool clean = true;
                for (uint32_t i = 0; i < nblocks; ++i) clean = clean && status_h[i] == 0;
                if (clean) {
                    // block i was written at the running sum of the decoded lengths: that is dst_off[i] only if every inner block is full
                    std::vector<uint32_t> ol(nblocks);
                    CK(cudaMemcpyAsync(ol.data(), out_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    for (uint32_t i = 0; clean && i + 1 < nblocks; ++i) clean = ol[i] == B;
                    clean = clean && ol[nblocks - 1] <= cap[nblocks - 1];
                }
                if (clean) return DLZ4_OK;
                // short inner blocks, or a block in error: one warp per block places and reports each on its own
            }
        }
    }
    return launch_decompress(ctx, src, src_off, src_len, nblocks, dst, dst_off, dst_cap, dict_len ? dict : nullptr, dict_len,
                             hist_mode == DLZ4_HIST_FRAME, nullptr, out_len, status, st);
}

int dlz4_decompress_blocks_dev(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len, uint32_t nblocks,
                               uint8_t *dst, const uint64_t *dst_off, const uint32_t *dst_cap, const uint8_t *dict, uint32_t dict_len,
                               int hist_mode, uint32_t *out_len, uint8_t *status, void *stream) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    return decompress_dev_routed(ctx, src, src_off, src_len, nblocks, dst, dst_off, dst_cap, dict, dict_len, hist_mode, out_len, status,
                                 pick(ctx, stream));
}

int dlz4_decompress_blocks(dlz4_ctx *ctx, const uint8_t *src, uint64_t src_bytes, const uint64_t *src_off, const uint32_t *src_len,
                           uint32_t nblocks, uint8_t *dst, uint64_t dst_bytes, const uint64_t *dst_off, const uint32_t *dst_cap,
                           const uint8_t *dict, uint32_t dict_len, int hist_mode, uint32_t *out_len, uint8_t *status) {
    if (!ctx || (nblocks && (!src_len || !dst_off || !dst_cap || !out_len || !status))) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!src_off) {
        // packed input (what the packed compress call returns): block i starts at the running sum of src_len
        if (!nblocks) return DLZ4_OK;
        bool ascending = true;
        for (uint32_t i = 0; i < nblocks; ++i) {
            if (dst_off[i] + dst_cap[i] > dst_bytes) return DLZ4_E_INVALID_ARG;
            if (i && dst_off[i] < dst_off[i - 1] + dst_cap[i - 1]) ascending = false;
        }
        if (!ascending) return DLZ4_E_INVALID_ARG;
        if (dict_len > 65536) { dict += dict_len - 65536; dict_len = 65536; }
        return decompress_blocks_packed(ctx, src, src_bytes, src_len, nblocks, dst, dst_bytes, dst_off, dst_cap, dict, dict_len,
                                        hist_mode, out_len, status);
    }
    for (uint32_t i = 0; i < nblocks; ++i) {
        if (src_off[i] + src_len[i] > src_bytes) return DLZ4_E_INVALID_ARG;
        if (dst_off[i] + dst_cap[i] > dst_bytes) return DLZ4_E_INVALID_ARG;
    }
    const uint64_t src_pad = (src_bytes + 15) & ~(uint64_t)15;
    CKS(reserve(ctx, ctx->work, src_pad + dict_len + 16));
    CKS(reserve(ctx, ctx->out, dst_bytes));
    CKS(reserve(ctx, ctx->meta, (size_t)nblocks * (8 + 8 + 4 + 4 + 4 + 1) + 64));
    uint8_t *d_src = (uint8_t *)ctx->work.p, *d_dict = d_src + src_pad, *d_dst = (uint8_t *)ctx->out.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_doff = d_soff + nblocks;
    uint32_t *d_slen = (uint32_t *)(d_doff + nblocks), *d_cap = d_slen + nblocks, *d_olen = d_cap + nblocks;
    uint8_t *d_status = (uint8_t *)(d_olen + nblocks);
    CKS(h2d(ctx, d_src, src, src_bytes, st));
    if (dict_len) CK(cudaMemcpyAsync(d_dict, dict, dict_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_soff, src_off, (size_t)nblocks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_doff, dst_off, (size_t)nblocks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_slen, src_len, (size_t)nblocks * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cap, dst_cap, (size_t)nblocks * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CKS(decompress_dev_routed(ctx, d_src, d_soff, d_slen, nblocks, d_dst, d_doff, d_cap, dict_len ? d_dict : nullptr, dict_len, hist_mode,
                              d_olen, d_status, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(out_len, d_olen, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, d_status, nblocks, cudaMemcpyDeviceToHost, st));
    CKS(d2h(ctx, dst, d_dst, dst_bytes, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    for (uint32_t i = 0; i < nblocks; ++i)
        if (status[i]) return status[i];
    return DLZ4_OK;
}

// ---- single raw block (LZ4.compressRaw / LZ4.decompressRaw) ----------------------------------------------------
int dlz4_compress_block(dlz4_ctx *ctx, const uint8_t *src, uint64_t src_total, int32_t src_start, int32_t src_len, int32_t *table,
                        uint8_t *output, uint64_t output_total, int32_t output_offset, int32_t *written) {
    if (!ctx || !table || !written || src_start < 0 || src_len < 0 || output_offset < 0) return DLZ4_E_INVALID_ARG;
    if ((uint64_t)src_start + (uint64_t)src_len > src_total) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t need = (uint64_t)src_start + src_len;
    const uint64_t bound = dlz4_compress_bound((uint64_t)src_len);
    CKS(reserve(ctx, ctx->work, need + 16));
    CKS(reserve(ctx, ctx->comp, bound));
    CKS(reserve(ctx, ctx->aux, kHashEntries * 4 + 64));
    int32_t *d_table = (int32_t *)ctx->aux.p;
    uint32_t *d_clen = (uint32_t *)((uint8_t *)ctx->aux.p + kHashEntries * 4);
    if (need) CK(cudaMemcpyAsync(ctx->work.p, src, need, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_table, table, kHashEntries * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CKS(launch_chain(ctx, (const uint8_t *)ctx->work.p, src_start, src_len, src_len > 0 ? src_len : 1, 1, d_table,
                     (uint8_t *)ctx->comp.p, 0, d_clen, st));
    CK(cudaEventRecord(ctx->ev1, st));
    uint32_t clen = 0;
    CK(cudaMemcpyAsync(&clen, d_clen, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(table, d_table, kHashEntries * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // the JS drops stores beyond output.length silently and still returns the full count (blockCompress.js has no checks)
    uint64_t room = (uint64_t)output_offset < output_total ? output_total - output_offset : 0;
    uint64_t ncopy = std::min<uint64_t>(clen, room);
    if (ncopy) CK(cudaMemcpy(output + output_offset, ctx->comp.p, ncopy, cudaMemcpyDeviceToHost));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    *written = (int32_t)clen;
    return DLZ4_OK;
}

int dlz4_decompress_block(dlz4_ctx *ctx, const uint8_t *input, uint64_t input_total, int64_t input_offset, int64_t input_size,
                          uint8_t *output, uint64_t output_total, int64_t output_offset, const uint8_t *dictionary, uint64_t dict_len,
                          int64_t *written) {
    if (!ctx || !written || input_offset < 0 || input_size < 0 || output_offset < 0) return DLZ4_E_INVALID_ARG;
    if ((uint64_t)input_offset + (uint64_t)input_size > input_total || input_size > 0x7FFFFFFF) return DLZ4_E_INVALID_ARG;
    if ((uint64_t)output_offset > output_total) return DLZ4_E_OUTPUT_TOO_SMALL;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!dictionary) dict_len = 0;
    // Matches reach at most 65535 bytes back: upload that much of output[0..outputOffset) as history and
    // the tail of the dictionary (it is indexed from its end, blockDecompress.js:147).
    const uint64_t hist = std::min<uint64_t>((uint64_t)output_offset, 65536);
    const uint64_t dwin = std::min<uint64_t>(dict_len, 65536);
    const uint64_t room = output_total - (uint64_t)output_offset;
    const uint64_t cap = std::min<uint64_t>(room, (uint64_t)input_size * 255 + 64);   // a block cannot expand more than 255x
    const uint64_t in_pad = ((uint64_t)input_size + 15) & ~15ull;
    CKS(reserve(ctx, ctx->work, in_pad + dwin + 32));
    CKS(reserve(ctx, ctx->out, hist + cap + 16));
    CKS(reserve(ctx, ctx->meta, 256));
    uint8_t *d_in = (uint8_t *)ctx->work.p, *d_dict = d_in + in_pad, *d_out = (uint8_t *)ctx->out.p;
    if (input_size) CK(cudaMemcpyAsync(d_in, input + input_offset, (size_t)input_size, cudaMemcpyHostToDevice, st));
    if (dwin) CK(cudaMemcpyAsync(d_dict, dictionary + (dict_len - dwin), dwin, cudaMemcpyHostToDevice, st));
    if (hist) CK(cudaMemcpyAsync(d_out, output + (output_offset - hist), hist, cudaMemcpyHostToDevice, st));
    // one block in frame-history mode: the device array's index 0 is `hist` bytes before the block
    struct { uint64_t soff, doff; uint32_t slen, cap, olen; uint8_t status; } h = {0, hist, (uint32_t)input_size, (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFFFu), 0, 0};
    uint8_t *m = (uint8_t *)ctx->meta.p;
    CK(cudaMemcpyAsync(m + 0, &h.soff, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m + 8, &h.doff, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m + 16, &h.slen, 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m + 20, &h.cap, 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    // history shorter than the real one is only possible when output_offset > 65536, where no offset can reach index 0,
    // so the dictionary branch (copySrc < 0) is taken exactly when the reference takes it.
    const bool truncated = (uint64_t)output_offset > hist;
    CKS(launch_decompress(ctx, d_in, (uint64_t *)(m + 0), (uint32_t *)(m + 16), 1, d_out, (uint64_t *)(m + 8), (uint32_t *)(m + 20),
                          (dwin && !truncated) ? d_dict : nullptr, truncated ? 0 : (uint32_t)dwin, 1, nullptr, (uint32_t *)(m + 24), m + 28, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(&h.olen, m + 24, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h.status, m + 28, 1, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    if (h.status) {
        // a dictionary window shorter than the caller's dictionary can only turn an in-range reference into DICT_OOB when
        // the reference reaches more than 64 KiB back, which the 16-bit offset cannot express
        return h.status;
    }
    if (h.olen) CK(cudaMemcpy(output + output_offset, d_out + hist, h.olen, cudaMemcpyDeviceToHost));
    *written = h.olen;
    return DLZ4_OK;
}

// ---- xxHash32 ---------------------------------------------------------------------------------------------------
int dlz4_xxh32_batch_dev(dlz4_ctx *ctx, const uint8_t *base, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t seed,
                         uint32_t *out, void *stream) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    return launch_xxh32_batch(ctx, base, off, len, n, seed, out, nullptr, pick(ctx, stream));
}

int dlz4_xxh32_stream_dev(dlz4_ctx *ctx, const uint8_t *data, uint64_t len, uint32_t seed, uint32_t *out, void *stream) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    return launch_xxh32_stream(ctx, data, len, seed, out, pick(ctx, stream));
}

// Whole-stream checksums of independent streams (the content checksums of the frames of a multi-frame job) are independent
// serial chains: each runs as its own single-warp kernel on its own stream, beside the block kernels, and reads its bytes where
// they already are -- device memory, or the caller's PAGE-LOCKED host memory directly over PCIe (no second upload).
int dlz4_xxh32_async(dlz4_ctx *ctx, int slot, const uint8_t *data, uint64_t len, uint32_t seed) {
    if (!ctx || slot < 0 || slot >= kSumSlots || (len && !data)) return DLZ4_E_INVALID_ARG;
    if (len >= 0x80000000ull) return DLZ4_E_TOO_LARGE;       // xxhash32.js:23 len|0
    CK(cudaSetDevice(ctx->device));
    const uint8_t *dptr = data;
    bool on_host = false;
    if (len) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, data) != cudaSuccess) { cudaGetLastError(); return DLZ4_E_INVALID_ARG; }
        if (a.type == cudaMemoryTypeUnregistered) return DLZ4_E_INVALID_ARG;      // pageable: the caller uses dlz4_xxh32 / _update
        if (a.type == cudaMemoryTypeHost) {
            void *dp = nullptr;
            if (cudaHostGetDevicePointer(&dp, (void *)data, 0) != cudaSuccess) { cudaGetLastError(); return DLZ4_E_INVALID_ARG; }
            dptr = (const uint8_t *)dp;
            on_host = true;
        }
    }
    if (!ctx->h_sum) CK(cudaHostAlloc((void **)&ctx->h_sum, kSumSlots * sizeof(uint32_t), cudaHostAllocMapped));
    if (!ctx->sum_stream[slot]) CK(cudaStreamCreateWithFlags(&ctx->sum_stream[slot], cudaStreamNonBlocking));
    void *res = nullptr;
    CK(cudaHostGetDevicePointer(&res, ctx->h_sum + slot, 0));
    if (on_host) k_xxh32_stream<3><<<1, 32, 0, ctx->sum_stream[slot]>>>(dptr, len, seed, (uint32_t *)res, nullptr, nullptr);
    else k_xxh32_stream<1><<<1, 32, 0, ctx->sum_stream[slot]>>>(dptr, len, seed, (uint32_t *)res, nullptr, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    return DLZ4_OK;
}

int dlz4_xxh32_wait(dlz4_ctx *ctx, int slot, uint32_t *out) {
    if (!ctx || slot < 0 || slot >= kSumSlots || !out || !ctx->sum_stream[slot] || !ctx->h_sum) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->sum_stream[slot]));
    *out = ctx->h_sum[slot];
    return DLZ4_OK;
}

int dlz4_xxh32(dlz4_ctx *ctx, const uint8_t *data, uint64_t len, uint32_t seed, uint32_t *out) {
    if (!ctx || !out) return DLZ4_E_INVALID_ARG;
    if (len >= 0x80000000ull) return DLZ4_E_TOO_LARGE;       // xxhash32.js:23 len|0
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CKS(reserve(ctx, ctx->work, len + 16));
    if (len) CK(cudaMemcpyAsync(ctx->work.p, data, len, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CKS(launch_xxh32_stream(ctx, (const uint8_t *)ctx->work.p, len, seed, ctx->d_hash, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(out, ctx->d_hash, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return DLZ4_OK;
}

int dlz4_xxh32_batch(dlz4_ctx *ctx, const uint8_t *base, uint64_t base_bytes, const uint64_t *off, const uint32_t *len, uint32_t n,
                     uint32_t seed, uint32_t *out) {
    if (!ctx || (n && (!off || !len || !out))) return DLZ4_E_INVALID_ARG;
    for (uint32_t i = 0; i < n; ++i)
        if (off[i] + len[i] > base_bytes) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CKS(reserve(ctx, ctx->work, base_bytes + 16));
    CKS(reserve(ctx, ctx->meta, (size_t)n * 16 + 64));
    uint64_t *d_off = (uint64_t *)ctx->meta.p;
    uint32_t *d_len = (uint32_t *)(d_off + n), *d_out = d_len + n;
    if (base_bytes) CK(cudaMemcpyAsync(ctx->work.p, base, base_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off, off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_len, len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CKS(launch_xxh32_batch(ctx, (const uint8_t *)ctx->work.p, d_off, d_len, n, seed, d_out, nullptr, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return DLZ4_OK;
}

// ---- frame packing on the device -----------------------------------------------------------------------------------
int dlz4_frame_pack_dev(dlz4_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len, const uint8_t *comp,
                        const uint64_t *comp_off, const uint32_t *comp_len, uint32_t nblocks, int block_checksum, uint8_t *segment,
                        uint64_t *block_pos, void *stream) {
    if (!ctx) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    CKS(reserve(ctx, ctx->aux, (size_t)nblocks * 12 + kHashEntries * 4 + 256));
    uint64_t *d_doff = (uint64_t *)((uint8_t *)ctx->aux.p + kHashEntries * 4 + 64);
    uint32_t *d_dlen = (uint32_t *)(d_doff + nblocks);
    k_frame_layout<<<1, 1024, 0, st>>>(src_len, comp_len, nblocks, block_checksum, block_pos, d_doff, d_dlen, 0);
    ctx->launches++;
    CK(cudaGetLastError());
    if (nblocks) {
        const int grid = (int)std::min<uint64_t>(nblocks, (uint64_t)ctx->sm_count * 8);
        k_frame_gather<<<grid, 256, 0, st>>>(src, src_off, src_len, comp, comp_off, comp_len, nblocks, block_pos, segment, 0);
        ctx->launches++;
        CK(cudaGetLastError());
        if (block_checksum) CKS(launch_xxh32_batch(ctx, segment, d_doff, d_dlen, nblocks, 0, nullptr, segment, st));
    }
    return DLZ4_OK;
}

// ---- frame header (bufferCompress.js:147-178): the one writer every path uses (frame calls, sharded frames, stream encoder)
size_t dlz4_frame_header(const dlz4_frame_opts *opts, uint64_t content_len, int have_dict, uint32_t dict_id, uint8_t out[19]) {
    size_t hp = 0;
    wr32(out, 0x184D2204u); hp = 4;                                     // :147 magic
    uint8_t flg = 1 << 6;                                               // :150 version 01
    if (opts->block_independence) flg |= 0x20;
    if (opts->content_checksum) flg |= 0x04;
    if (have_dict) flg |= 0x01;
    if (opts->add_content_size) flg |= 0x08;
    if (opts->block_checksum) flg |= 0x10;                              // addition (LZ4 frame spec)
    out[hp++] = flg;
    out[hp++] = (uint8_t)((block_id_for(opts->max_block_size) & 7) << 4);   // :160 BD
    if (opts->add_content_size) {                                       // :163-168 u64 from len|0 (frames are < 2 GiB)
        wr32(out + hp, (uint32_t)content_len); wr32(out + hp + 4, (uint32_t)(content_len >> 32)); hp += 8;
    }
    if (have_dict) { wr32(out + hp, dict_id); hp += 4; }                // :171-175
    out[hp] = (uint8_t)((header_xxh32(out + 4, hp - 4) >> 8) & 0xFF); hp++;   // :177-178 HC
    return hp;
}

// ---- frame compress (compressBuffer) ---------------------------------------------------------------------------------
// The block loop of compressBuffer (bufferCompress.js:209-239) for one input of < 2 GiB, everything on the device: stages
// dictionary window ++ input, launches the content checksum on the side stream, compresses every block and packs
// [u32 size | stored bit][payload][u32 xxh32]* into ctx->seg.  On return the stream is idle, *seg_len = bytes of that body,
// hdr[0, *hp) the frame header; the caller adds EndMark / checksum.  The input stays resident (ctx->res_in) for
// dlz4_xxh32_update_resident.
static int frame_body(dlz4_ctx *ctx, const uint8_t *input, uint64_t input_len, const uint8_t *dictionary, uint64_t dict_len,
                      const dlz4_frame_opts *opts, uint8_t *hdr, size_t *hp_out, uint64_t *seg_len_out) {
    cudaStream_t st = ctx->stream;
    const bool have_dict = dictionary && dict_len > 0;                  // :109
    const uint64_t dwin = have_dict ? std::min<uint64_t>(dict_len, 65536) : 0;   // :115
    const int bd = block_id_for(opts->max_block_size);                  // :128
    const uint32_t B = kBlockMax[bd];
    const uint32_t n = (uint32_t)((input_len + B - 1) / B);
    const uint64_t stride = (dlz4_compress_bound(B) + 15) & ~15ull;
    const uint64_t dpad = (dwin + 15) & ~15ull;                          // keep the input 16-byte aligned
    CKS(reserve(ctx, ctx->work, dict_len + dpad + input_len + 64));
    uint8_t *d_work = (uint8_t *)ctx->work.p + (dpad - dwin);            // dictionary window directly before the input
    uint8_t *d_in = d_work + dwin;
    uint8_t *d_fulldict = (uint8_t *)ctx->work.p + dpad + ((input_len + 15) & ~15ull);
    CKS(reserve(ctx, ctx->comp, (uint64_t)n * stride + 64));
    CKS(reserve(ctx, ctx->seg, input_len + (uint64_t)n * 8 + 64));
    CKS(reserve(ctx, ctx->meta, (size_t)n * (8 + 8 + 4 + 4) + 8 * ((size_t)n + 1) + 64));
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_coff = d_soff + n, *d_pos = d_coff + n;
    uint32_t *d_slen = (uint32_t *)(d_pos + n + 1), *d_clen = d_slen + n;

    // Segment-engine frames of 64 MiB and more parse while the input is still arriving: the copy goes out in 16 MiB chunks on
    // the copy stream, each followed by a 4-byte flag copy, and a segment waits for the flags of what it reads.
    const bool segments = input_len >= ctx->seg_min_bytes && (!opts->block_independence || (B > 65536 && !dwin));
    const bool overlap = segments && input_len >= ctx->seg_overlap_min_bytes;
    if (overlap) {
        CK(cudaMemsetAsync(ctx->d_land, 0, kLandFlags * 4, st));
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_fork, 0));
        const uint64_t chunk = 1ull << kLandShift;
        for (uint64_t c = 0, o = 0; o < input_len; ++c, o += chunk) {
            CKS(h2d(ctx, d_in + o, input + o, (size_t)std::min<uint64_t>(chunk, input_len - o), ctx->copy_in));
            CK(cudaMemcpyAsync(ctx->d_land + c, ctx->h_one, 4, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        CK(cudaEventRecord(ctx->evp[0], ctx->copy_in));                 // whole input in memory
    } else if (input_len) {
        CKS(h2d(ctx, d_in, input, input_len, st));
    }
    if (dwin) CK(cudaMemcpyAsync(d_work, dictionary + (dict_len - dwin), dwin, cudaMemcpyHostToDevice, st));

    // header (host, bufferCompress.js:147-178)
    uint32_t dict_id = 0;
    if (have_dict) {                                                    // :112 dictId = xxHash32(dict) on the GPU
        if (dict_len > dwin) {
            CK(cudaMemcpyAsync(d_fulldict, dictionary, dict_len, cudaMemcpyHostToDevice, ctx->side));
            CKS(launch_xxh32_stream(ctx, d_fulldict, dict_len, 0, ctx->d_hash + 1, ctx->side));
        } else {
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            CKS(launch_xxh32_stream(ctx, d_work, dwin, 0, ctx->d_hash + 1, ctx->side));
        }
        CK(cudaMemcpyAsync(&dict_id, ctx->d_hash + 1, 4, cudaMemcpyDeviceToHost, ctx->side));
        CK(cudaStreamSynchronize(ctx->side));
    }
    const size_t hp = dlz4_frame_header(opts, input_len, have_dict ? 1 : 0, dict_id, hdr);

    // content checksum: serial xxh32 over the whole input on the side stream, overlapped with the block kernels (:248-252)
    if (opts->content_checksum) {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        if (overlap) CK(cudaStreamWaitEvent(ctx->side, ctx->evp[0], 0));
        CKS(launch_xxh32_stream(ctx, d_in, input_len, 0, ctx->d_hash, ctx->side));
        CK(cudaEventRecord(ctx->ev_side, ctx->side));
    }

    CK(cudaEventRecord(ctx->ev0, st));
    uint64_t seg_len = 0;
    ctx->seg_jobs = ctx->seg_reruns = ctx->seg_rounds = 0;
    if (n) {
        k_uniform_blocks<<<(n + 255) / 256, 256, 0, st>>>((uint64_t)(d_in - d_work), input_len, B, n, d_soff, d_slen, d_coff, stride);
        ctx->launches++;
        CK(cudaGetLastError());
        uint8_t *d_comp = (uint8_t *)ctx->comp.p;
        if (!opts->block_independence) {
            // linked blocks: one serial chain, table and history carried across blocks (:182,:219,:234)
            CK(cudaMemsetAsync(ctx->d_table, 0, kHashEntries * 4, st));
            if (dwin >= 4) { k_warm_jenkins<<<((int)dwin - 3 + 255) / 256, 256, 0, st>>>(d_work, (int32_t)dwin, ctx->d_table); ctx->launches++; }
            if (input_len >= ctx->seg_min_bytes)
                // long chain: speculative segments, verified against the serial parse's state (k_compress_segments)
                CKS(compress_segmented(ctx, d_work, (int64_t)dwin, (int64_t)input_len, (int64_t)B, n, true, ctx->d_table, d_comp, d_coff, d_clen, st,
                                       nullptr, overlap ? ctx->d_land : nullptr, (int32_t)dwin, kLandShift));
            else
                CKS(launch_chain(ctx, d_work, (int32_t)dwin, (int32_t)input_len, (int32_t)B, n, ctx->d_table, d_comp, stride, d_clen, st));
        } else {
            uint32_t first = 0;
            if (dwin) {
                // block 0 alone sees the dictionary prefix and the Jenkins-warmed table (:186-204); the table is
                // cleared after it (:234-236), so every later block is a fresh independent block
                CK(cudaMemsetAsync(ctx->d_table, 0, kHashEntries * 4, st));
                if (dwin >= 4) { k_warm_jenkins<<<((int)dwin - 3 + 255) / 256, 256, 0, st>>>(d_work, (int32_t)dwin, ctx->d_table); ctx->launches++; }
                const int32_t l0 = (int32_t)std::min<uint64_t>(B, input_len);
                CKS(launch_chain(ctx, d_work, (int32_t)dwin, l0, l0, 1, ctx->d_table, d_comp, stride, d_clen, st));
                first = 1;
            }
            if (B > 65536 && n > first && input_len >= ctx->seg_min_bytes)
                // large independent blocks: segments inside every block (the first segment of a block starts exactly)
                CKS(compress_segmented(ctx, d_work, (int64_t)dwin + (int64_t)first * B, (int64_t)input_len - (int64_t)first * B, (int64_t)B,
                                       n - first, false, nullptr, d_comp, d_coff + first, d_clen + first, st, nullptr,
                                       overlap ? ctx->d_land : nullptr, (int32_t)dwin, kLandShift));
            else
                CKS(launch_compress(ctx, d_work, d_soff + first, d_slen + first, n - first, B, nullptr, 0, nullptr, d_comp, d_coff + first,
                                    d_clen + first, st));
        }
        if (overlap) CK(cudaStreamWaitEvent(st, ctx->evp[0], 0));       // (stored blocks are copied from the input)
        CKS(dlz4_frame_pack_dev(ctx, d_work, d_soff, d_slen, d_comp, d_coff, d_clen, n, opts->block_checksum, (uint8_t *)ctx->seg.p,
                                d_pos, st));
        CK(cudaMemcpyAsync(&seg_len, d_pos + n, 8, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));

    ctx->res_in = d_in; ctx->res_in_len = input_len;
    *hp_out = hp;
    *seg_len_out = seg_len;
    return DLZ4_OK;
}

int dlz4_frame_compress(dlz4_ctx *ctx, const uint8_t *input, uint64_t input_len, const uint8_t *dictionary, uint64_t dict_len,
                        const dlz4_frame_opts *opts, uint8_t *output, uint64_t output_cap, uint64_t *output_len) {
    if (!ctx || !opts || !output_len || (input_len && !input)) return DLZ4_E_INVALID_ARG;
    if (input_len >= 0x7FFF0000ull) return DLZ4_E_TOO_LARGE;            // bufferCompress.js:127 len|0
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool have_dict = dictionary && dict_len > 0;                  // :109
    const uint64_t dwin = have_dict ? std::min<uint64_t>(dict_len, 65536) : 0;   // :115
    const int bd = block_id_for(opts->max_block_size);                  // :128
    const uint32_t B = kBlockMax[bd];
    const uint32_t n = (uint32_t)((input_len + B - 1) / B);
    const uint64_t stride = (dlz4_compress_bound(B) + 15) & ~15ull;

    // device staging: work = dictionary window ++ input (the reference's workingBuffer, :121-124)
    if (opts->block_independence && B <= 65536 && !have_dict && !opts->content_checksum && input_len >= ctx->frame_pipe_min_bytes &&
        output_cap >= dlz4_frame_bound(input_len)) {     // (with a content checksum the serial xxh32 is the whole cost: old path)
        // large frame of small independent blocks: the chunked host pipeline (H2D / kernels / D2H overlapped) writes the
        // frame body straight into `output`; header, EndMark and the content checksum are added around it
        uint8_t hdr[32];
        const size_t hp = dlz4_frame_header(opts, input_len, 0, 0, hdr);
        memcpy(output, hdr, hp);
        std::vector<uint64_t> off(n);
        std::vector<uint32_t> len(n);
        for (uint32_t i = 0; i < n; ++i) { off[i] = (uint64_t)i * B; len[i] = (uint32_t)std::min<uint64_t>(B, input_len - off[i]); }
        uint64_t body = 0;
        ctx->seg_jobs = ctx->seg_reruns = ctx->seg_rounds = 0;
        CKS(compress_blocks_packed(ctx, input, input_len, off.data(), len.data(), n, B, output + hp, output_cap - hp - 8, nullptr, 1,
                                   opts->block_checksum, &body));
        uint8_t foot[8] = {0, 0, 0, 0, 0, 0, 0, 0};                         // EndMark (:244)
        memcpy(output + hp + body, foot, opts->content_checksum ? 8 : 4);
        *output_len = hp + body + 4 + (opts->content_checksum ? 4 : 0);
        return DLZ4_OK;
    }
    uint8_t hdr[32];
    size_t hp = 0;
    uint64_t seg_len = 0;
    CKS(frame_body(ctx, input, input_len, dictionary, dict_len, opts, hdr, &hp, &seg_len));

    const uint64_t total = hp + seg_len + 4 + (opts->content_checksum ? 4 : 0);
    *output_len = total;
    // an undersized outputBuffer truncates silently in the reference (typed-array stores are dropped); same here
    std::vector<uint8_t> tail;
    uint64_t pos = 0;
    auto put = [&](const uint8_t *p, uint64_t len) {
        if (pos < output_cap) memcpy(output + pos, p, (size_t)std::min<uint64_t>(len, output_cap - pos));
        pos += len;
    };
    put(hdr, hp);
    if (seg_len) {
        if (pos < output_cap)
            CKS(d2h(ctx, output + pos, ctx->seg.p, (size_t)std::min<uint64_t>(seg_len, output_cap - pos), st));
        pos += seg_len;
    }
    uint8_t foot[8] = {0, 0, 0, 0, 0, 0, 0, 0};                           // EndMark (:244)
    if (opts->content_checksum) {
        uint32_t hsh = 0;
        CK(cudaMemcpyAsync(&hsh, ctx->d_hash, 4, cudaMemcpyDeviceToHost, ctx->side));
        CK(cudaStreamSynchronize(ctx->side));
        wr32(foot + 4, hsh);
    }
    put(foot, opts->content_checksum ? 8 : 4);
    CK(cudaStreamSynchronize(st));
    return DLZ4_OK;
}

// ---- sharded frames: one rank's contiguous range of independent blocks ------------------------------------
int dlz4_frame_body_compress(dlz4_ctx *ctx, const uint8_t *input, uint64_t input_len, uint32_t max_block_size, int block_checksum,
                             uint64_t *body_len) {
    if (!ctx || !body_len || (input_len && !input)) return DLZ4_E_INVALID_ARG;
    if (input_len >= 0x7FFF0000ull) return DLZ4_E_TOO_LARGE;
    CK(cudaSetDevice(ctx->device));
    dlz4_frame_opts o{max_block_size, 1, 0, 0, block_checksum};
    uint8_t hdr[32];
    size_t hp = 0;
    ctx->res_body_len = 0;
    CKS(frame_body(ctx, input, input_len, nullptr, 0, &o, hdr, &hp, body_len));
    ctx->res_body_len = *body_len;
    return DLZ4_OK;
}

int dlz4_frame_body_fetch(dlz4_ctx *ctx, uint8_t *dst, uint64_t dst_cap) {
    if (!ctx || (ctx->res_body_len && !dst)) return DLZ4_E_INVALID_ARG;
    if (ctx->res_body_len > dst_cap) return DLZ4_E_OUTPUT_TOO_SMALL;
    CK(cudaSetDevice(ctx->device));
    if (ctx->res_body_len) CKS(d2h(ctx, dst, ctx->seg.p, ctx->res_body_len, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DLZ4_OK;
}

int dlz4_xxh32_update_resident(dlz4_ctx *ctx, dlz4_xxh32_state *s, int which) {
    if (!ctx || !s || (which != DLZ4_RESIDENT_INPUT && which != DLZ4_RESIDENT_OUTPUT)) return DLZ4_E_INVALID_ARG;
    const uint8_t *d = which == DLZ4_RESIDENT_INPUT ? ctx->res_in : ctx->res_out;
    const uint64_t len = which == DLZ4_RESIDENT_INPUT ? ctx->res_in_len : ctx->res_out_len;
    if (len == 0) return DLZ4_OK;
    if (!d || s->memsize != 0) return DLZ4_E_INVALID_ARG;               // pieces before the last are whole stripes (whole blocks)
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t body = len & ~15ull;
    if (body) {
        CK(cudaMemcpyAsync(ctx->d_hash + 4, s->v, 16, cudaMemcpyHostToDevice, st));
        k_xxh32_stream<1><<<1, 32, 0, st>>>(d, body, s->seed, nullptr, ctx->d_hash + 4, ctx->d_hash + 8);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(s->v, ctx->d_hash + 8, 16, cudaMemcpyDeviceToHost, st));
    }
    if (len > body) CK(cudaMemcpyAsync(s->mem, d + body, len - body, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    s->memsize = (uint32_t)(len - body);
    s->total += len;
    return DLZ4_OK;
}

// ---- frame decompress (decompressBuffer) -------------------------------------------------------------------------------
static int parse_header(const uint8_t *f, uint64_t len, dlz4_frame_info_t *info, uint64_t *body_pos) {
    memset(info, 0, sizeof *info);
    if (len < 4 || rd32(f) != 0x184D2204u) return DLZ4_E_BAD_MAGIC;      // bufferDecompress.js:59-61
    if (len < 5) { info->version = 0; return DLZ4_E_BAD_VERSION; }
    uint64_t pos = 4;
    const uint8_t flg = f[pos++];
    info->flg = flg;
    info->version = (flg & 0xC0) >> 6;
    if (info->version != 1) return DLZ4_E_BAD_VERSION;                  // :67
    info->has_block_checksum = (flg & 0x10) != 0;
    info->has_content_size = (flg & 0x08) != 0;
    info->has_content_checksum = (flg & 0x04) != 0;
    info->has_dict_id = (flg & 0x01) != 0;
    info->block_independence = (flg & 0x20) != 0;
    if (pos >= len) return DLZ4_E_MALFORMED;
    info->bd = f[pos++];                                                // :75 (the JS skips it; we use it to size buffers)
    const int bid = (info->bd >> 4) & 7;
    info->block_max_size = bid >= 4 ? kBlockMax[bid] : 4194304u;
    if (info->has_content_size) {
        if (pos + 8 > len) return DLZ4_E_MALFORMED;
        info->content_size = (uint64_t)rd32(f + pos) | ((uint64_t)rd32(f + pos + 4) << 32);   // :81-85
        pos += 8;
    }
    if (info->has_dict_id) {
        if (pos + 4 > len) return DLZ4_E_MALFORMED;
        info->dict_id = rd32(f + pos);
        pos += 4;
    }
    pos += 1;                                                           // :92 header checksum
    if (pos > len) return DLZ4_E_MALFORMED;
    *body_pos = pos;
    return DLZ4_OK;
}

struct BlockRef { uint64_t off; uint32_t len; uint8_t stored; };

static int walk_blocks(const uint8_t *f, uint64_t len, const dlz4_frame_info_t *info, uint64_t pos, std::vector<BlockRef> *blocks,
                       uint64_t *end_pos) {
    while (pos < len) {                                                  // :133
        if (pos + 4 > len) return DLZ4_E_MALFORMED;
        const uint32_t bs = rd32(f + pos);
        pos += 4;
        if (bs == 0) break;                                              // :139 EndMark
        const uint32_t actual = bs & 0x7FFFFFFFu;
        if (pos + actual > len) return DLZ4_E_MALFORMED;
        if (blocks) blocks->push_back({pos, actual, (uint8_t)((bs >> 31) & 1)});
        pos += actual;
        if (info->has_block_checksum) {                                  // :191
            if (pos + 4 > len) return DLZ4_E_MALFORMED;                   // truncated frame: the checksum bytes must exist
            pos += 4;
        }
    }
    *end_pos = pos;
    return DLZ4_OK;
}

int dlz4_frame_info(const uint8_t *frame, uint64_t frame_len, dlz4_frame_info_t *info) {
    if (!frame || !info) return DLZ4_E_INVALID_ARG;
    uint64_t pos = 0, end = 0;
    int s = parse_header(frame, frame_len, info, &pos);
    if (s) return s;
    std::vector<BlockRef> blocks;
    s = walk_blocks(frame, frame_len, info, pos, &blocks, &end);
    if (s) return s;
    info->nblocks = (uint32_t)blocks.size();
    uint64_t bound = 0;
    for (const BlockRef &b : blocks) bound += b.stored ? b.len : std::min<uint64_t>((uint64_t)b.len * 255, info->block_max_size);
    info->max_decoded = info->content_size ? info->content_size : bound;
    info->frame_bytes = end + (info->has_content_checksum ? 4 : 0);
    if (info->frame_bytes > frame_len) return DLZ4_E_MALFORMED;
    return DLZ4_OK;
}

}  // extern "C"
// Walks concatenated / skippable frames (LZ4 frame spec); fn(frame pointer, info) is called per LZ4 frame.
template <class Fn>
static int for_each_frame(const uint8_t *data, uint64_t len, Fn fn) {
    uint64_t pos = 0;
    while (pos < len) {
        if (pos + 4 > len) return DLZ4_E_BAD_MAGIC;
        const uint32_t magic = rd32(data + pos);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {                       // skippable frame
            if (pos + 8 > len) return DLZ4_E_MALFORMED;
            const uint64_t sz = rd32(data + pos + 4);
            if (pos + 8 + sz > len) return DLZ4_E_MALFORMED;
            pos += 8 + sz;
            continue;
        }
        dlz4_frame_info_t info;
        int s = dlz4_frame_info(data + pos, len - pos, &info);
        if (s) return s;
        s = fn(data + pos, info);
        if (s) return s;
        pos += info.frame_bytes;
    }
    return DLZ4_OK;
}
extern "C" {

int dlz4_frames_info(const uint8_t *data, uint64_t data_len, uint64_t *max_decoded, uint32_t *frames) {
    if (!data || !max_decoded) return DLZ4_E_INVALID_ARG;
    uint64_t total = 0;
    uint32_t count = 0;
    const int s = for_each_frame(data, data_len, [&](const uint8_t *, const dlz4_frame_info_t &info) { total += info.max_decoded; ++count; return DLZ4_OK; });
    *max_decoded = total;
    if (frames) *frames = count;
    return s;
}

int dlz4_frame_decompress(dlz4_ctx *ctx, const uint8_t *frame, uint64_t frame_len, const uint8_t *dictionary, uint64_t dict_len,
                          uint32_t flags, uint8_t *output, uint64_t output_cap, uint64_t *output_len) {
    return dlz4_frame_decompress_ex(ctx, frame, frame_len, dictionary, dict_len, flags, output, output_cap, output_len, nullptr);
}

static int frame_decompress_impl(dlz4_ctx *ctx, const uint8_t *frame_in, uint64_t frame_len_in, const uint8_t *dictionary, uint64_t dict_len,
                                 uint32_t flags, uint8_t *output, uint64_t output_cap, uint64_t *output_len, uint32_t *block_out_len,
                                 uint32_t first_block, uint32_t block_count /* 0xFFFFFFFF: to the end */) {
    if (!ctx || !frame_in || !output_len) return DLZ4_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    dlz4_frame_info_t info;
    uint64_t pos = 0, end = 0;
    CKS(parse_header(frame_in, frame_len_in, &info, &pos));
    if ((flags & 4u)) {
        const uint64_t hc_pos = pos - 1;
        if ((uint8_t)((header_xxh32(frame_in + 4, (size_t)(hc_pos - 4)) >> 8) & 0xFF) != frame_in[hc_pos]) return DLZ4_E_HEADER_CHECKSUM;
    }
    std::vector<BlockRef> blocks;
    CKS(walk_blocks(frame_in, frame_len_in, &info, pos, &blocks, &end));
    // A block range (sharded decode, SURVEY 8e: rank r decodes blocks dlz4_shard_range(...) into its slice of the output):
    // the range is treated as a frame of its own whose bytes start at its first block's size word.  Independent blocks only
    // (a linked block needs the output before it), inner blocks must be full (checked by the caller through block_out_len),
    // and the whole-stream content checksum is the caller's business (dlz4_xxh32_update_resident relay).
    const uint32_t n_all = (uint32_t)blocks.size();
    const bool ranged = !(first_block == 0 && block_count >= n_all);
    const uint8_t *frame = frame_in;
    uint64_t frame_len = frame_len_in;
    ctx->res_out = nullptr; ctx->res_out_len = 0;
    if (ranged) {
        if (first_block > n_all) return DLZ4_E_INVALID_ARG;
        block_count = std::min<uint32_t>(block_count, n_all - first_block);
        if (!info.block_independence && first_block != 0) return DLZ4_E_INVALID_ARG;
        const uint64_t B0 = info.block_max_size;
        if (info.content_size) {
            const uint64_t lo_c = std::min<uint64_t>(info.content_size, (uint64_t)first_block * B0);
            const uint64_t hi_c = first_block + block_count >= n_all ? info.content_size
                                                                     : std::min<uint64_t>(info.content_size, (uint64_t)(first_block + block_count) * B0);
            info.content_size = hi_c - lo_c;
            info.has_content_size = info.content_size != 0;
        }
        flags &= ~1u;                                                    // content checksum: not over a part
        info.has_content_checksum = 0;
        if (block_count == 0) { *output_len = 0; return DLZ4_OK; }
        const uint64_t lo = blocks[first_block].off - 4;
        const BlockRef &lb = blocks[first_block + block_count - 1];
        const uint64_t hi = lb.off + lb.len + (info.has_block_checksum ? 4 : 0);
        blocks.assign(blocks.begin() + first_block, blocks.begin() + first_block + block_count);
        for (BlockRef &b : blocks) b.off -= lo;
        frame = frame_in + lo;
        frame_len = hi - lo;
        end = frame_len;
    }
    const uint32_t n = (uint32_t)blocks.size();
    if (!dictionary) dict_len = 0;
    const uint64_t dwin = std::min<uint64_t>(dict_len, 65536);
    const uint32_t B = info.block_max_size;

    if (info.block_independence && B <= 65536 && n >= 512 && !dict_len && !(flags & 2u) && info.content_size >= ctx->frame_pipe_min_bytes &&
        info.content_size <= output_cap && (uint64_t)(n - 1) * B < info.content_size && info.content_size <= (uint64_t)n * B) {
        // large frame of small independent blocks with a declared size: the chunked host pipeline (H2D / kernels / D2H
        // overlapped), every block at i * blockMaxSize.  A frame whose inner blocks are not full falls through to the path below.
        std::vector<uint64_t> soff(n), doff(n);
        std::vector<uint32_t> slen(n), cap(n), olen(n);
        std::vector<uint8_t> status(n), stored(n);
        for (uint32_t i = 0; i < n; ++i) {
            soff[i] = blocks[i].off; slen[i] = blocks[i].len; stored[i] = blocks[i].stored;
            doff[i] = (uint64_t)i * B;
            cap[i] = (uint32_t)std::min<uint64_t>(B, info.content_size - doff[i]);
        }
        const int s = decompress_blocks_packed(ctx, frame, frame_len, slen.data(), n, output, info.content_size, doff.data(), cap.data(),
                                               nullptr, 0, DLZ4_HIST_RAW, olen.data(), status.data(), soff.data(), stored.data());
        if (s == DLZ4_E_CUDA || s == DLZ4_E_INVALID_ARG) return s;
        bool full = s == DLZ4_OK;
        for (uint32_t i = 0; full && i < n; ++i) full = olen[i] == cap[i];
        if (full) {
            if (info.has_content_checksum && (flags & 1u)) {             // :213-217 over the decoded bytes still resident in ctx->out
                if (end + 4 > frame_len) return DLZ4_E_CONTENT_CHECKSUM;
                uint32_t h = 0;
                CKS(launch_xxh32_stream(ctx, (const uint8_t *)ctx->out.p, info.content_size, 0, ctx->d_hash, st));
                CK(cudaMemcpyAsync(&h, ctx->d_hash, 4, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (h != rd32(frame + end)) return DLZ4_E_CONTENT_CHECKSUM;
            }
            *output_len = info.content_size;
            if (block_out_len) memcpy(block_out_len, olen.data(), (size_t)n * 4);
            ctx->res_out = (const uint8_t *)ctx->out.p; ctx->res_out_len = info.content_size;
            return DLZ4_OK;
        }
        // an error or a short inner block: the general path below decides (same status order as ever)
    }
    // capacity the decode may use: the reference allocates contentSize when present (:107), otherwise grows as needed
    uint64_t cap_total = info.content_size ? info.content_size : 0;
    if (!info.content_size) {
        for (const BlockRef &b : blocks) cap_total += b.stored ? b.len : std::min<uint64_t>((uint64_t)b.len * 255, B);
    }
    if (cap_total > output_cap) cap_total = output_cap;

    const uint64_t fpad = (frame_len + 15) & ~15ull;
    CKS(reserve(ctx, ctx->work, fpad + dwin + 32));
    CKS(reserve(ctx, ctx->out, cap_total + 64));
    CKS(reserve(ctx, ctx->meta, (size_t)n * (8 + 8 + 4 + 4 + 4 + 1 + 1 + 4) + 256));
    uint8_t *d_frame = (uint8_t *)ctx->work.p, *d_dict = d_frame + fpad, *d_out = (uint8_t *)ctx->out.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_doff = d_soff + n;
    uint32_t *d_slen = (uint32_t *)(d_doff + n), *d_cap = d_slen + n, *d_olen = d_cap + n, *d_bhash = d_olen + n;
    uint8_t *d_status = (uint8_t *)(d_bhash + n), *d_stored = d_status + n;

    // A frame of 8 MiB and more for the jump decoder's chunked scan is copied in up to 16 groups of whole blocks on the copy stream; the scan
    // of a group starts when it has landed (decompress_jump), the rest of the decoder waits for all of them through `st`.
    std::vector<uint32_t> groups;
    const bool jump_scan = n && frame_len >= ctx->jump_min_bytes && (!info.block_independence || (B > 65536 && n < 1024)) && B > 65536 &&
                           !ctx->k_jd_serial_scan;
    if (jump_scan && frame_len >= (8ull << 20) && !((flags & 2u) && info.has_block_checksum) && !ctx->k_jd_no_overlap) {
        const uint64_t target = std::max<uint64_t>(4ull << 20, (frame_len + 14) / 15);      // <= 16 groups
        uint64_t begin = 0;
        groups.push_back(0);
        for (uint32_t i = 0; i < n; ++i) {
            const bool last = i + 1 == n;
            const uint64_t end = last ? frame_len : blocks[i + 1].off;           // (runs into the next block's size word: harmless)
            if (last || end - begin >= target) {
                CKS(h2d(ctx, d_frame + begin, frame + begin, end - begin, ctx->copy_in));
                CK(cudaEventRecord(ctx->evp[kJdGroupEvent + groups.size() - 1], ctx->copy_in));
                groups.push_back(i + 1);
                begin = end;
            }
        }
    } else {
        CKS(h2d(ctx, d_frame, frame, frame_len, st));
    }
    if (dwin) CK(cudaMemcpyAsync(d_dict, dictionary + (dict_len - dwin), dwin, cudaMemcpyHostToDevice, st));

    std::vector<uint64_t> soff(n), doff(n);
    std::vector<uint32_t> slen(n), cap(n), olen(n), bhash(n);
    std::vector<uint8_t> status(n), stored(n);
    for (uint32_t i = 0; i < n; ++i) { soff[i] = blocks[i].off; slen[i] = blocks[i].len; stored[i] = blocks[i].stored; }
    CK(cudaMemcpyAsync(d_soff, soff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_slen, slen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_stored, stored.data(), n, cudaMemcpyHostToDevice, st));

    if ((flags & 2u) && info.has_block_checksum && n) {                  // addition: verify block checksums on the GPU
        CKS(launch_xxh32_batch(ctx, d_frame, d_soff, d_slen, n, 0, d_bhash, nullptr, st));
        CK(cudaMemcpyAsync(bhash.data(), d_bhash, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }

    CK(cudaEventRecord(ctx->ev0, st));
    uint64_t total = 0;
    int first_status = 0;
    bool shipped = false;                                                // the jump decoder copies finished units out as it goes
    if (n) {
        const bool jump = frame_len >= ctx->jump_min_bytes && (!info.block_independence || (B > 65536 && n < 1024));
        if (jump) {
            // linked blocks (block k reads block k-1's output) or few large blocks: token scan + pointer doubling
            CKS(decompress_jump(ctx, d_frame, fpad, d_soff, d_slen, d_stored, slen, n, B, d_out, cap_total, dwin ? d_dict : nullptr, (uint32_t)dwin,
                                !info.block_independence, d_olen, d_status, status, &total, st, output, output_cap, &shipped,
                                groups.empty() ? nullptr : &groups));
            for (uint32_t i = 0; i < n && !first_status; ++i) first_status = status[i];
        } else if (!info.block_independence) {
            // short linked frame: serial chain (one warp)
            k_decompress_chain<<<1, 32, 0, st>>>(d_frame, d_soff, d_slen, d_stored, n, d_out, cap_total, dwin ? d_dict : nullptr,
                                                (uint32_t)dwin, d_olen, d_status, ctx->d_total);
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&total, ctx->d_total, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(status.data(), d_status, n, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < n && !first_status; ++i) first_status = status[i];
        } else {
            // independent blocks: every block is decoded by its own warp at i * blockMaxSize (every writer this library
            // meets -- the reference, liblz4, the lz4 CLI -- emits full blocks except the last); a short inner block is
            // detected afterwards and the output is closed up on the device.  That layout needs room for n - 1 full blocks:
            // a frame that cannot have them (flushed short inner blocks, LZ4F_flush-style writers, a stream's update() that
            // carries several small blocks), or one where a block ran out of its strided room, is decoded at running
            // offsets instead, like the reference's sequential loop (bufferDecompress.js:133-192): token scan -> exact
            // sizes -> exclusive scan -> decode (the jump decoder with independent history).
            bool running = !((uint64_t)(n - 1) * B < cap_total);
            if (!running) {
                for (uint32_t i = 0; i < n; ++i) {
                    doff[i] = (uint64_t)i * B;
                    const uint64_t room = doff[i] < cap_total ? cap_total - doff[i] : 0;
                    cap[i] = (uint32_t)std::min<uint64_t>(room, B);
                }
                CK(cudaMemcpyAsync(d_doff, doff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(d_cap, cap.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
                // history of an independent block is the dictionary only (LZ4 frame spec; for frames the reference writes,
                // blocks > 0 never reach before their own start, so this equals bufferDecompress.js:153 on them)
                CKS(launch_decompress(ctx, d_frame, d_soff, d_slen, n, d_out, d_doff, d_cap, dwin ? d_dict : nullptr, (uint32_t)dwin, 0,
                                      d_stored, d_olen, d_status, st));
                CK(cudaMemcpyAsync(olen.data(), d_olen, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(status.data(), d_status, n, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                bool contiguous = true;
                for (uint32_t i = 0; i < n; ++i) {
                    if (status[i] && !first_status) first_status = status[i];
                    if (i + 1 < n && olen[i] != B) contiguous = false;
                }
                // "too small" may be the strided room, not the caller's: only the running-offset decode can tell
                for (uint32_t i = 0; i < n; ++i)
                    if (status[i] == DLZ4_E_OUTPUT_TOO_SMALL && (uint64_t)i * B + B > cap_total) running = true;
                if (running) first_status = 0;
                if (!running && !first_status && !contiguous) {
                    // close the gaps front to back (forward moves never overlap later data)
                    uint64_t w = 0;
                    for (uint32_t i = 0; i < n; ++i) {
                        const uint64_t gap = doff[i] - w;                 // pieces of <= gap bytes never overlap their source
                        for (uint64_t done = 0; gap && done < olen[i]; done += gap)
                            CK(cudaMemcpyAsync(d_out + w + done, d_out + doff[i] + done, (size_t)std::min<uint64_t>(gap, olen[i] - done),
                                               cudaMemcpyDeviceToDevice, st));
                        w += olen[i];
                    }
                }
                if (!running) for (uint32_t i = 0; i < n; ++i) total += olen[i];
            }
            if (running) {
                CKS(decompress_jump(ctx, d_frame, fpad, d_soff, d_slen, d_stored, slen, n, B, d_out, cap_total, dwin ? d_dict : nullptr,
                                    (uint32_t)dwin, false, d_olen, d_status, status, &total, st, output, output_cap, &shipped, nullptr));
                for (uint32_t i = 0; i < n && !first_status; ++i) first_status = status[i];
            }
        }
    }
    CK(cudaEventRecord(ctx->ev1, st));
    if (shipped) CK(cudaStreamSynchronize(ctx->copy_out));               // no copy into the caller's buffer survives this call
    if (first_status) { CK(cudaStreamSynchronize(st)); return first_status; }

    if ((flags & 2u) && info.has_block_checksum) {
        CK(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < n; ++i)
            if (rd32(frame + blocks[i].off + blocks[i].len) != bhash[i]) return DLZ4_E_BLOCK_CHECKSUM;
    }
    if (info.has_content_checksum && (flags & 1u)) {                     // :213-217
        if (end + 4 > frame_len) return DLZ4_E_CONTENT_CHECKSUM;
        uint32_t h = 0;
        CKS(launch_xxh32_stream(ctx, d_out, total, 0, ctx->d_hash, st));
        CK(cudaMemcpyAsync(&h, ctx->d_hash, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h != rd32(frame + end)) return DLZ4_E_CONTENT_CHECKSUM;
    }
    *output_len = total;
    if (total > output_cap) return DLZ4_E_OUTPUT_TOO_SMALL;
    if (block_out_len && n) CK(cudaMemcpyAsync(block_out_len, d_olen, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (total && !shipped) CKS(d2h(ctx, output, d_out, total, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->res_out = d_out; ctx->res_out_len = total;
    return DLZ4_OK;
}

int dlz4_frame_decompress_ex(dlz4_ctx *ctx, const uint8_t *frame, uint64_t frame_len, const uint8_t *dictionary, uint64_t dict_len,
                             uint32_t flags, uint8_t *output, uint64_t output_cap, uint64_t *output_len, uint32_t *block_out_len) {
    return frame_decompress_impl(ctx, frame, frame_len, dictionary, dict_len, flags, output, output_cap, output_len, block_out_len, 0,
                                 0xFFFFFFFFu);
}

int dlz4_frame_decompress_range(dlz4_ctx *ctx, const uint8_t *frame, uint64_t frame_len, uint32_t first_block, uint32_t block_count,
                                const uint8_t *dictionary, uint64_t dict_len, uint32_t flags, uint8_t *output, uint64_t output_cap,
                                uint64_t *output_len, uint32_t *block_out_len) {
    return frame_decompress_impl(ctx, frame, frame_len, dictionary, dict_len, flags, output, output_cap, output_len, block_out_len,
                                 first_block, block_count);
}

// ---- stateful xxh32 (the reference's XXHash32 class: update / digest, src/xxhash32/xxhash32Stateful.js) ---------------
void dlz4_xxh32_reset(dlz4_xxh32_state *s, uint32_t seed) {
    if (!s) return;
    memset(s, 0, sizeof *s);
    s->seed = seed;
    s->v[0] = seed + 2654435761u + 2246822519u; s->v[1] = seed + 2246822519u; s->v[2] = seed; s->v[3] = seed - 2654435761u;
}

int dlz4_xxh32_update(dlz4_ctx *ctx, dlz4_xxh32_state *s, const uint8_t *data, uint64_t len) {
    if (!ctx || !s || (len && !data)) return DLZ4_E_INVALID_ARG;
    s->total += len;
    if (s->memsize + len < 16) {                                        // not a stripe yet: host-side buffering only
        memcpy(s->mem + s->memsize, data, (size_t)len);
        s->memsize += (uint32_t)len;
        return DLZ4_OK;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // stripes = pending tail ++ data, cut at a multiple of 16; the rest becomes the new tail
    const uint64_t avail = s->memsize + len, body = avail & ~15ull, from_data = body - s->memsize;
    CKS(reserve(ctx, ctx->work, body + 64));
    uint8_t *d = (uint8_t *)ctx->work.p;
    if (s->memsize) CK(cudaMemcpyAsync(d, s->mem, s->memsize, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + s->memsize, data, from_data, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_hash + 4, s->v, 16, cudaMemcpyHostToDevice, st));
    k_xxh32_stream<1><<<1, 32, 0, st>>>(d, body, s->seed, nullptr, ctx->d_hash + 4, ctx->d_hash + 8);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->v, ctx->d_hash + 8, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    s->memsize = (uint32_t)(avail - body);
    memcpy(s->mem, data + from_data, s->memsize);
    return DLZ4_OK;
}

uint32_t dlz4_xxh32_digest(const dlz4_xxh32_state *s) {
    // the accumulators come from the GPU; what is left is the merge and <= 15 tail bytes (xxhash32.js:59-97), like the header byte
    auto rl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    uint32_t h = s->total >= 16 ? rl(s->v[0], 1) + rl(s->v[1], 7) + rl(s->v[2], 12) + rl(s->v[3], 18) : s->seed + 374761393u;
    h += (uint32_t)s->total;
    const uint8_t *p = s->mem, *end = s->mem + s->memsize;
    while (p + 4 <= end) { h = rl(h + rd32(p) * 3266489917u, 17) * 668265263u; p += 4; }
    while (p < end) { h = rl(h + (*p) * 374761393u, 11) * 2654435761u; ++p; }
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return h;
}

// ---- one linked chain with the table in/out (LZ4Encoder._flushBlock over every full block of an add(), lz4Encode.js:215-298) ----
int dlz4_chain_compress(dlz4_ctx *ctx, const uint8_t *work, uint64_t work_len, int32_t start, int32_t total, int32_t block_size,
                        int32_t *table, uint8_t *dst, uint64_t dst_stride, uint32_t *comp_len) {
    if (!ctx || !work || !table || !dst || !comp_len || start < 0 || total < 0 || block_size <= 0) return DLZ4_E_INVALID_ARG;
    if ((uint64_t)start + (uint64_t)total > work_len || work_len >= 0x7FFFFFF0ull) return DLZ4_E_INVALID_ARG;
    if (total == 0) return DLZ4_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)(((int64_t)total + block_size - 1) / block_size);
    const uint64_t stride = (dlz4_compress_bound((uint64_t)block_size) + 15) & ~15ull;
    if (dst_stride < dlz4_compress_bound((uint64_t)std::min(block_size, total))) return DLZ4_E_OUTPUT_TOO_SMALL;
    CKS(reserve(ctx, ctx->work, work_len + 64));
    CKS(reserve(ctx, ctx->comp, (uint64_t)n * stride + 64));
    CKS(reserve(ctx, ctx->meta, (size_t)n * (8 + 8 + 4 + 4) + 64));
    uint8_t *d_work = (uint8_t *)ctx->work.p, *d_comp = (uint8_t *)ctx->comp.p;
    uint64_t *d_soff = (uint64_t *)ctx->meta.p, *d_coff = d_soff + n;
    uint32_t *d_slen = (uint32_t *)(d_coff + n), *d_clen = d_slen + n;
    CK(cudaMemcpyAsync(d_work, work, work_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_table, table, kHashEntries * 4, cudaMemcpyHostToDevice, st));
    k_uniform_blocks<<<(n + 255) / 256, 256, 0, st>>>((uint64_t)start, (uint64_t)total, (uint32_t)block_size, n, d_soff, d_slen, d_coff, stride);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev0, st));
    if ((uint64_t)total >= ctx->seg_min_bytes)
        CKS(compress_segmented(ctx, d_work, start, total, block_size, n, true, ctx->d_table, d_comp, d_coff, d_clen, st, ctx->d_table));
    else
        CKS(launch_chain(ctx, d_work, start, total, block_size, n, ctx->d_table, d_comp, stride, d_clen, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(comp_len, d_clen, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(table, ctx->d_table, kHashEntries * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t k = 0; k < n; ++k) {
        const uint64_t take = std::min<uint64_t>(comp_len[k], dst_stride);       // undersized room truncates like a typed array
        if (take) CK(cudaMemcpyAsync(dst + (uint64_t)k * dst_stride, d_comp + (uint64_t)k * stride, take, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return DLZ4_OK;
}

int dlz4_frames_decompress(dlz4_ctx *ctx, const uint8_t *data, uint64_t data_len, const uint8_t *dictionary, uint64_t dict_len,
                           uint32_t flags, uint8_t *output, uint64_t output_cap, uint64_t *output_len, uint32_t *frames) {
    if (!ctx || !data || !output_len) return DLZ4_E_INVALID_ARG;
    uint64_t written = 0;
    uint32_t count = 0;
    const int s = for_each_frame(data, data_len, [&](const uint8_t *f, const dlz4_frame_info_t &info) {
        uint64_t got = 0;
        const int r = dlz4_frame_decompress(ctx, f, info.frame_bytes, dictionary, dict_len, flags, output ? output + written : nullptr,
                                            output_cap - written, &got);
        if (r) return r;
        written += got;
        ++count;
        return (int)DLZ4_OK;
    });
    *output_len = written;
    if (frames) *frames = count;
    return s;
}

}  // extern "C"
