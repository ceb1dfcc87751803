// gather2.cu -- which piece of the L2-table window sequence is slow?  Each variant adds one piece of the real kernel's
// per-window memory traffic to a dependent loop (7 warps per SM, 1 CTA per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// bit0: scatter st.cg same slots   bit1: read back same slots   bit2: un-insert (half the lanes)   bit3: byte stores to out
// bit4: 3 x LDG.128 __ldg from a 64 KiB source window   bit5: use weak (L1) accesses for the table instead of .cg
template <int F>
__global__ void k(uint16_t *tabs, const uint4 *src, uint8_t *out, int iters, unsigned long long *cycles, uint32_t *sink) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t wid = (size_t)blockIdx.x * nw + warp;
    uint16_t *t = tabs + wid * 16384;
    const uint4 *s = src + wid * 4096;          // 64 KiB per warp
    uint8_t *o = out + wid * 65536;
    uint32_t x = lane * 2654435761u + warp * 40503u + blockIdx.x, acc = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = (x >> 10) & 16383u;
        uint32_t v = (F & 32) ? (uint32_t)t[h] : (uint32_t)__ldcg(t + h);
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0;
        if (F & 16) { const uint4 *c = s + ((v ^ x) & 4093u); q0 = __ldg(c); q1 = __ldg(c + 1); q2 = __ldg(c + 2); }
        const uint16_t mine = (uint16_t)(i * 32 + lane + 1);
        __syncwarp();
        if (F & 1) { if (F & 32) t[h] = mine; else __stcg(t + h, mine); }
        __syncwarp();
        uint32_t rb = mine;
        if (F & 2) rb = (F & 32) ? (uint32_t)t[h] : (uint32_t)__ldcg(t + h);
        const uint32_t conflict = __ballot_sync(0xffffffffu, rb != mine);
        acc += q0.x ^ q1.y ^ q2.z ^ conflict;
        if ((F & 4) && (lane & 1)) { if (F & 32) t[h] = (uint16_t)v; else __stcg(t + h, (uint16_t)v); }
        if (F & 8) o[(i * 32 + lane) & 65535] = (uint8_t)v;
        x ^= v + (acc & 1);
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
    if (acc == 0xdeadbeef) *sink = acc;
}
template <int F> void run(int sms, int w, uint16_t *tabs, uint4 *src, uint8_t *out, unsigned long long *cyc, uint32_t *sink, const char *name) {
    const int iters = 1000;
    k<F><<<sms, w * 32>>>(tabs, src, out, iters, cyc, sink); cudaDeviceSynchronize(); cudaMemset(cyc, 0, 8);
    k<F><<<sms, w * 32>>>(tabs, src, out, iters, cyc, sink); cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-58s warps/SM %2d : %8.1f cycles/iter (%s)\n", name, w, (double)c / ((double)sms * w * iters), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    const size_t nw = (size_t)sms * 32;
    uint16_t *tabs; cudaMalloc(&tabs, nw * 32768); cudaMemset(tabs, 0, nw * 32768);
    uint4 *src; cudaMalloc(&src, nw * 65536 + 64); cudaMemset(src, 1, nw * 65536 + 64);
    uint8_t *out; cudaMalloc(&out, nw * 65536);
    unsigned long long *cyc; cudaMalloc(&cyc, 8); uint32_t *sink; cudaMalloc(&sink, 4);
    for (int w : {1, 7, 14}) {
        run<0>(sms, w, tabs, src, out, cyc, sink, "ld.cg gather");
        run<1>(sms, w, tabs, src, out, cyc, sink, "+ st.cg scatter");
        run<3>(sms, w, tabs, src, out, cyc, sink, "+ st.cg scatter + readback");
        run<7>(sms, w, tabs, src, out, cyc, sink, "+ st.cg scatter + readback + un-insert");
        run<15>(sms, w, tabs, src, out, cyc, sink, "+ scatter + readback + un-insert + out bytes");
        run<31>(sms, w, tabs, src, out, cyc, sink, "+ scatter + readback + un-insert + out bytes + 3xLDG.128");
        run<16>(sms, w, tabs, src, out, cyc, sink, "ld.cg gather + 3xLDG.128");
        run<63>(sms, w, tabs, src, out, cyc, sink, "everything, weak (L1) table accesses");
    }
    return 0;
}
