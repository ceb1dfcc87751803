// gather.cu -- microbenchmark: cost of a warp-wide random 16-bit gather (+ scatter) on a 32 KiB table that lives in
// L2-resident global memory, as a function of warps per SM and of the access flavour.  Answers "why are L2-table warps slow".
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather gather.cu ; run: ./gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint16_t *tabs, int iters, unsigned long long *cycles, uint32_t *sink) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint16_t *t = tabs + ((size_t)blockIdx.x * nw + warp) * 16384;
    uint32_t x = lane * 2654435761u + warp * 40503u + blockIdx.x, acc = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = (x >> 10) & 16383u;
        uint32_t v;
        if (MODE == 0) v = t[h];                                   // weak, L1-cached
        else if (MODE == 1) v = __ldcg(t + h);                     // .cg
        else if (MODE == 2) { v = __ldcg(t + h); __stcg(t + h, (uint16_t)(v + 1)); }          // .cg gather + scatter, same slot
        else if (MODE == 3) { v = t[h]; t[h] = (uint16_t)(v + 1); }                              // weak gather + scatter
        else if (MODE == 4) { asm volatile("ld.global.cv.u16 %0, [%1];" : "=r"(v) : "l"(t + h)); }
        else { asm volatile("ld.relaxed.cta.global.u16 %0, [%1];" : "=r"(v) : "l"(t + h)); asm volatile("st.relaxed.cta.global.u16 [%0], %1;" :: "l"(t + h), "r"(v + 1)); }
        acc += v;
        x ^= v;                                                    // dependent chain, like the parse
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
    if (acc == 0xdeadbeef) *sink = acc;
}
int main() {
    int sms = 148; cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); sms = p.multiProcessorCount;
    uint16_t *tabs; cudaMalloc(&tabs, (size_t)sms * 32 * 32768); cudaMemset(tabs, 0, (size_t)sms * 32 * 32768);
    unsigned long long *cyc; cudaMalloc(&cyc, 8); uint32_t *sink; cudaMalloc(&sink, 4);
    const int iters = 2000;
    const char *names[] = {"weak ld (L1)", "ld.cg", "ld.cg + st.cg", "weak ld + st", "ld.cv", "relaxed.cta ld+st"};
    for (int mode = 0; mode < 6; ++mode)
        for (int w : {1, 4, 7, 14, 28}) {
            cudaMemset(cyc, 0, 8);
            auto run = [&](int m) {
                switch (m) { case 0: k<0><<<sms, w * 32>>>(tabs, iters, cyc, sink); break; case 1: k<1><<<sms, w * 32>>>(tabs, iters, cyc, sink); break;
                             case 2: k<2><<<sms, w * 32>>>(tabs, iters, cyc, sink); break; case 3: k<3><<<sms, w * 32>>>(tabs, iters, cyc, sink); break;
                             case 4: k<4><<<sms, w * 32>>>(tabs, iters, cyc, sink); break; default: k<5><<<sms, w * 32>>>(tabs, iters, cyc, sink); }
            };
            run(mode); cudaDeviceSynchronize(); cudaMemset(cyc, 0, 8); run(mode); cudaDeviceSynchronize();
            unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-20s warps/SM %2d : %8.1f cycles per dependent iteration (%s)\n", names[mode], w, (double)c / ((double)sms * w * iters), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
