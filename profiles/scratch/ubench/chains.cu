// chains.cu -- microbenchmarks behind two round-2 decisions:
//  (1) the xxh32 stripe chain v' = rotl(v + x*P2, 13)*P1: cycles per stripe of four algebraically equal forms
//      (one warp, lanes 0-3 active, data from shared memory like k_xxh32_stream);
//  (2) same-slot detection inside a window: __match_any_sync on 32 distinct values vs. store-tag / read-back on a
//      shared-memory table (what dlz4_wide.cuh uses).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chains chains.cu ; run: ./chains
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u;

template <int MODE>
__global__ void k_xxh(const uint32_t *in, int chunks, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint32_t buf[1024];
    const uint32_t lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) buf[i] = in[i];
    __syncwarp();
    uint32_t v = lane * 77u + 1u;
    const long long t0 = clock64();
    if (lane < 4) {
        const uint32_t *w = buf + lane;
        for (int c = 0; c < chunks; ++c) {
            if (MODE == 0) {                 // textbook: three dependent operations
#pragma unroll 16
                for (int t = 0; t < 256; ++t) { v += w[t * 4] * P2; v = __funnelshift_l(v, v, 13); v *= P1; }
            } else if (MODE == 1) {          // round 1: a' = (a >> 19)*P1 + (a*(P1 << 13) + y')
                constexpr uint32_t K1 = P1 << 13;
                uint32_t a = v + w[0] * P2;
#pragma unroll 16
                for (int t = 1; t < 256; ++t) {
                    const uint32_t y = w[t * 4] * P2;
                    uint32_t cc;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(cc) : "r"(a), "r"(K1), "r"(y));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(a >> 19), "r"(P1), "r"(cc));
                }
                v = __funnelshift_l(a, a, 13) * P1;
            } else if (MODE == 2) {          // a' = rotl(a,13)*P1 + y' : shift + multiply-add
                uint32_t a = v + w[0] * P2;
#pragma unroll 16
                for (int t = 1; t < 256; ++t) {
                    const uint32_t y = w[t * 4] * P2;
                    const uint32_t r = __funnelshift_l(a, a, 13);
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(r), "r"(P1), "r"(y));
                }
                v = __funnelshift_l(a, a, 13) * P1;
            } else if (MODE == 3) {          // both links on the multiplier pipe: (a >> 19) as mul.hi(a, 2^13)
                constexpr uint32_t K1 = P1 << 13;
                uint32_t a = v + w[0] * P2;
#pragma unroll 16
                for (int t = 1; t < 256; ++t) {
                    const uint32_t y = w[t * 4] * P2;
                    uint32_t cc, hi;
                    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(8192u));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(cc) : "r"(a), "r"(K1), "r"(y));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(hi), "r"(P1), "r"(cc));
                }
                v = __funnelshift_l(a, a, 13) * P1;
            } else {                         // 64-bit product a * 2^13: (lo, hi) = (a << 13, a >> 19), then two multiply-adds
                uint32_t a = v + w[0] * P2;
#pragma unroll 16
                for (int t = 1; t < 256; ++t) {
                    const uint32_t y = w[t * 4] * P2;
                    const uint64_t wide = (uint64_t)a * 8192ull;
                    const uint32_t r = (uint32_t)wide + (uint32_t)(wide >> 32);
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(r), "r"(P1), "r"(y));
                }
                v = __funnelshift_l(a, a, 13) * P1;
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) *cycles = (unsigned long long)(t1 - t0);
    if (v == 0xdeadbeef) *sink = v;
    if (lane < 4) sink[1 + lane] = v;
}

template <int MODE>
__global__ void k_conf(int iters, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint16_t tab[16384];
    const uint32_t lane = threadIdx.x;
    for (int i = lane; i < 16384; i += 32) tab[i] = 0;
    __syncwarp();
    uint32_t x = lane * 2654435761u + 12345u, acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = (x >> 10) & 16383u;
        if (MODE == 0) {
            const uint32_t same = __match_any_sync(0xffffffffu, h);
            acc += __ballot_sync(0xffffffffu, same != (1u << lane));
        } else if (MODE == 1) {
            const uint32_t old = tab[h];
            __syncwarp();
            tab[h] = (uint16_t)(lane + 1);
            __syncwarp();
            const uint32_t r = tab[h];
            acc += __ballot_sync(0xffffffffu, r != lane + 1) + old;
            __syncwarp();
            tab[h] = (uint16_t)old;
        } else {
            // all lanes the same value (best case for match.any)
            const uint32_t same = __match_any_sync(0xffffffffu, h & 0u);
            acc += __ballot_sync(0xffffffffu, same != (1u << lane));
        }
        x ^= acc;
    }
    const long long t1 = clock64();
    if (lane == 0) *cycles = (unsigned long long)(t1 - t0);
    if (acc == 0xdeadbeef) *sink = acc;
}

int main() {
    uint32_t *in, *sink;
    unsigned long long *cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&sink, 64); cudaMalloc(&cyc, 8);
    uint32_t h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = i * 2654435761u + 99u;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    const int chunks = 2000;
    uint32_t res[5][4];
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: k_xxh<0><<<1, 32>>>(in, chunks, cyc, sink); break;
                case 1: k_xxh<1><<<1, 32>>>(in, chunks, cyc, sink); break;
                case 2: k_xxh<2><<<1, 32>>>(in, chunks, cyc, sink); break;
                case 3: k_xxh<3><<<1, 32>>>(in, chunks, cyc, sink); break;
                default: k_xxh<4><<<1, 32>>>(in, chunks, cyc, sink); break;
            }
            cudaDeviceSynchronize();
        }
        unsigned long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[mode], sink + 1, 16, cudaMemcpyDeviceToHost);
        printf("xxh chain mode %d: %.2f cycles/stripe  (v0=%08x %s)\n", mode, (double)c / (chunks * 256.0), res[mode][0],
               res[mode][0] == res[0][0] && res[mode][3] == res[0][3] ? "same result" : "DIFFERENT");
    }
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k_conf<0><<<1, 32>>>(iters, cyc, sink);
            else if (mode == 1) k_conf<1><<<1, 32>>>(iters, cyc, sink);
            else k_conf<2><<<1, 32>>>(iters, cyc, sink);
            cudaDeviceSynchronize();
        }
        unsigned long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("conflict test mode %d (%s): %.1f cycles/iteration\n", mode,
               mode == 0 ? "match.any, 32 distinct" : mode == 1 ? "smem tag store + read back + restore" : "match.any, all equal",
               (double)c / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
