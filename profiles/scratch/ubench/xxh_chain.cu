// xxh_chain.cu -- cycles per stripe of the xxh32 accumulator chain in three formulations (one warp, 4 active lanes, data in smem).
//   0: v = rotl(v + x*P2, 13) * P1 (three links)   1: a' = (a >> 19)*P1 + (a*K1 + y), shift on the alu pipe (two links, cross-pipe)
//   2: the same with a >> 19 as mul.hi(a, 2^13) (fma pipe only)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o xxh_chain xxh_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, K1 = P1 << 13;
template <int kMode>
__global__ void k(uint32_t *out, long long *cyc, int reps) {
    __shared__ uint32_t buf[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) buf[i] = i * 2654435761u + 12345u;
    __syncwarp();
    uint32_t a = threadIdx.x * 7u + 1u;
    const long long t0 = clock64();
    if (threadIdx.x < 4) {
        const uint32_t *w = buf + threadIdx.x;
        for (int r = 0; r < reps; ++r) {
#pragma unroll 16
            for (int t = 0; t < 256; ++t) {
                const uint32_t y = w[t * 4];
                if (kMode == 0) {
                    a += y * P2; a = (a << 13) | (a >> 19); a *= P1;
                } else if (kMode == 1) {
                    uint32_t c;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(c) : "r"(a), "r"(K1), "r"(y));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(a >> 19), "r"(P1), "r"(c));
                } else {
                    uint32_t c, h;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(c) : "r"(a), "r"(K1), "r"(y));
                    asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(a), "r"(1u << 13));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(h), "r"(P1), "r"(c));
                }
            }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    uint32_t *o; long long *c, h;
    cudaMalloc(&o, 128); cudaMalloc(&c, 8);
    const int reps = 2000;
    for (int m = 0; m < 3; ++m) {
        for (int it = 0; it < 2; ++it) {
            if (m == 0) k<0><<<1, 32>>>(o, c, reps); else if (m == 1) k<1><<<1, 32>>>(o, c, reps); else k<2><<<1, 32>>>(o, c, reps);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %.2f cycles per 16-byte stripe\n", m, (double)h / (256.0 * reps));
    }
    return 0;
}
