// gather5.cu -- how long after a ld.cg must a st.cg to the SAME slot wait to avoid the slow path found by gather3?
// Dependent loop: load slot h, spin DELAY cycles (clock64), store slot h; 7 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int BITS>
__global__ void k(uint8_t *tabs, int iters, int delay, unsigned long long *cycles, uint32_t *sink) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t wid = (size_t)blockIdx.x * nw + warp;
    uint16_t *t16 = reinterpret_cast<uint16_t *>(tabs + wid * 65536);
    uint32_t *t32 = reinterpret_cast<uint32_t *>(tabs + wid * 65536);
    uint32_t x = lane * 2654435761u + warp * 40503u + blockIdx.x, acc = 0;
    long long spent = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = (x >> 10) & 16383u;
        uint32_t v = BITS == 16 ? (uint32_t)__ldcg(t16 + h) : __ldcg(t32 + h);
        acc += v;
        const long long a = clock64() + (v & 1);       // wait for the load, then spin
        while (clock64() - a < delay) { }
        spent += delay;
        if (BITS == 16) __stcg(t16 + h, (uint16_t)(i * 32 + lane + 1)); else __stcg(t32 + h, (uint32_t)(i * 32 + lane + 1));
        x ^= v;
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0 - spent));
    if (acc == 0xdeadbeef) *sink = acc;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    const size_t nw = (size_t)sms * 8;
    uint8_t *tabs; cudaMalloc(&tabs, nw * 65536); cudaMemset(tabs, 0, nw * 65536);
    unsigned long long *cyc; cudaMalloc(&cyc, 8); uint32_t *sink; cudaMalloc(&sink, 4);
    const int iters = 300, w = 7;
    for (int bits : {16, 32})
        for (int delay : {0, 200, 500, 1000, 2000, 4000, 8000}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(cyc, 0, 8);
                if (bits == 16) k<16><<<sms, w * 32>>>(tabs, iters, delay, cyc, sink); else k<32><<<sms, w * 32>>>(tabs, iters, delay, cyc, sink);
                cudaDeviceSynchronize();
            }
            unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%d-bit entries, store %5d cycles after the load returned: %8.1f cycles/iter excluding the delay (%s)\n", bits, delay,
                   (double)c / ((double)sms * w * iters), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
