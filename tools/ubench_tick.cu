// ubench_tick.cu -- the order-2 PDM tick (glide + pdm2, out_shift 24) in isolation: clocks per
// tick per scheduler as a function of resident warps per scheduler, for the tick formulations
// of k_pdm_v2_ws2 (dither read from shared memory, no PRNG, no stores except one per 16 ticks).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tick ubench_tick.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t s, uint32_t d) { uint32_t a; asm("lop3.b32 %0, %1, 0xFF000000, %2, 0xEA;" : "=r"(a) : "r"(s), "r"(d)); return a; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t t; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c)); return t; }
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c) { uint32_t t; asm("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(t) : "r"(a), "r"(b), "r"(c)); return t; }
__device__ __forceinline__ uint32_t pack_top_bytes(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
    return __byte_perm(__byte_perm(a0, a1, 0x0073), __byte_perm(a2, a3, 0x0073), 0x5410);
}

template <int FORM>
__device__ __forceinline__ uint32_t tick(uint32_t &p, uint32_t v, uint32_t &s1, uint32_t &s2, uint32_t d, uint32_t m1, uint32_t m2) {
    p += v;
    uint32_t a;
    if (FORM == 0) { a = lop3_and_or(s2, d); uint32_t t = imad(a, m1, p); s1 += t; s2 += s1 - a; }
    if (FORM == 1) { uint32_t x = s1 + p; a = lop3_and_or(s2, d); s1 = imad(a, m1, x); s2 = s2 + s1 - a; }
    if (FORM == 2) { uint32_t x = s1 + p; uint32_t u = add3(s2, s1, p); a = lop3_and_or(s2, d); s1 = imad(a, m1, x); s2 = imad(a, m2, u); }
    if (FORM == 3) { uint32_t x = s1 + p; uint32_t u = x + s2; a = lop3_and_or(s2, d); s1 = x - a; s2 = u - a - a; }      // ptxas' choice
    if (FORM == 4) { a = lop3_and_or(s2, d); s1 = add3(s1, p, 0u - a); s2 = add3(s2, s1, 0u - a); }                     // 4-op, all ALU
    return a;
}

template <int FORM>
__global__ void k(uint4 *out, const uint32_t *st, uint32_t m1, uint32_t m2, int groups, long long *clk) {
    __shared__ __align__(16) uint32_t dbuf[16][32][4];
    for (int i = threadIdx.x; i < 16 * 32 * 4; i += blockDim.x) (&dbuf[0][0][0])[i] = (i * 2654435761u) & 0x3FF;
    __syncthreads();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = st[gid & 1023], v = st[(gid + 1) & 1023], s1 = st[(gid + 2) & 1023], s2 = st[(gid + 3) & 1023];
    const int bl = (threadIdx.x & 31) / 3;
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const uint4 dv = *reinterpret_cast<const uint4 *>(&dbuf[(g * 4 + i4) & 15][bl][0]);
            const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
            uint32_t a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tick<FORM>(p, v, s1, s2, d[i], m1, m2);
            w[i4] = pack_top_bytes(a[0], a[1], a[2], a[3]);
        }
        if (s2 == 0x12345678u) out[gid] = make_uint4(w[0], w[1], w[2], w[3]);    // practically never
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
    if (p + s1 + s2 == 0x9abcdef0u) out[gid] = make_uint4(p, s1, s2, 0);
}

template <int FORM>
void run(uint4 *out, uint32_t *st, long long *dclk, int sms) {
    const int groups = 4096;
    printf("FORM %d:", FORM);
    for (int wps : {1, 2, 3, 4, 5, 6, 8, 12, 16}) {
        k<FORM><<<sms, 128 * wps>>>(out, st, 0xFFFFFFFFu, 0xFFFFFFFEu, groups, dclk);
        cudaDeviceSynchronize();
        k<FORM><<<sms, 128 * wps>>>(out, st, 0xFFFFFFFFu, 0xFFFFFFFEu, groups, dclk);
        long long c; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
        printf("  w%d %.1f (%.2f)", wps, (double)c / (groups * 16), (double)c / (groups * 16) / wps);
    }
    printf("   [clk per tick per scheduler (per warp-tick)]\n");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint4 *out; uint32_t *st; long long *dclk;
    cudaMalloc(&out, sizeof(uint4) * p.multiProcessorCount * 2048); cudaMalloc(&st, 4096); cudaMalloc(&dclk, 8);
    cudaMemset(st, 0x5a, 4096);
    run<0>(out, st, dclk, p.multiProcessorCount);
    run<1>(out, st, dclk, p.multiProcessorCount);
    run<2>(out, st, dclk, p.multiProcessorCount);
    run<3>(out, st, dclk, p.multiProcessorCount);
    run<4>(out, st, dclk, p.multiProcessorCount);
    return 0;
}
