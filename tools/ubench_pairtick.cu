// ubench_pairtick.cu -- the instruction mix of k_xvoice_mix2's pair tick with one or two independent voice pairs in flight per thread
// (register accumulators over a 32- or 16-frame chunk), at the kernel's occupancy: how much of the 43 clk per pair-tick is latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pairtick ubench_pairtick.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u2;
__device__ __forceinline__ u2 pk(float a, float b) { u2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi(u2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ u2 fma2(u2 a, u2 b, u2 c) { u2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u2 add2(u2 a, u2 b) { u2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u2 mul2(u2 a, u2 b) { u2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void addalu(uint32_t &x, uint32_t i) { asm("add.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(i)); }
struct Pair { uint32_t phA, phB, incA, incB; u2 f2, nf2, nq2, nlp, bp, nd; float neA, neB, glA, glB, grA, grB; };
__device__ __forceinline__ void init(Pair &v, uint32_t s) {
    v.phA = s * 2654435761u; v.phB = s * 40503u + 7; v.incA = 39370533u + s; v.incB = 23409859u + 3 * s;
    const float f = 0.05f + 1e-6f * (s & 1023), q = 0.9f;
    v.f2 = pk(f, f * 1.1f); v.nf2 = v.f2 ^ 0x8000000080000000ull; v.nq2 = pk(-q, -q * 1.01f); v.nlp = pk(0.f, 0.f); v.bp = pk(0.f, 0.f); v.nd = pk(1e-3f, 2e-3f);
    v.neA = -0.5f; v.neB = -0.25f; v.glA = 0.3f; v.glB = 0.6f; v.grA = 0.7f; v.grB = 0.4f;
}
#define TICK(v, k) { const u2 xi = pk(__int2float_rn((int)v.phA), __int2float_rn((int)v.phB)); addalu(v.phA, v.incA); addalu(v.phB, v.incB); \
    v.nlp = fma2(v.nf2, v.bp, v.nlp); u2 hp = fma2(xi, c31, v.nlp); hp = fma2(v.nq2, v.bp, hp); v.bp = fma2(v.f2, hp, v.bp); \
    const u2 s_ = add2(pk(v.neA, v.neB), v.nd); v.neA = fmaxf(fminf(lo(s_), -0.0f), -1.0f); v.neB = fmaxf(fminf(hi(s_), -0.0f), -1.0f); \
    const u2 y = mul2(v.nlp, pk(v.neA, v.neB)); const float yA = lo(y), yB = hi(y); \
    aL[k] = __fmaf_rn(v.glA, yA, aL[k]); aR[k] = __fmaf_rn(v.grA, yA, aR[k]); aL[k] = __fmaf_rn(v.glB, yB, aL[k]); aR[k] = __fmaf_rn(v.grB, yB, aR[k]); }
template <int ILP, int CH>
__global__ void __launch_bounds__(128) k(float *out, int n_pairs, int n_chunks) {
    const u2 c31 = pk(0x1p-31f, 0x1p-31f);
    float tot = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        float aL[CH], aR[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) { aL[q] = 0.f; aR[q] = 0.f; }
#pragma unroll 1
        for (int j = 0; j < n_pairs; j += ILP) {
            Pair v0, v1;
            init(v0, threadIdx.x + 131 * j + c); if (ILP == 2) init(v1, threadIdx.x + 131 * j + 77 + c);
#pragma unroll
            for (int q = 0; q < CH; ++q) { TICK(v0, q) if (ILP == 2) TICK(v1, q) }
            tot += lo(v0.bp) + v0.neA; if (ILP == 2) tot += lo(v1.bp) + v1.neB;
        }
#pragma unroll
        for (int q = 0; q < CH; ++q) tot += aL[q] + aR[q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}
template <int ILP, int CH> void run(int bps, float *d, int sms, double clk) {
    const int n_pairs = 12, n_chunks = 512 / CH * 2;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<ILP, CH>, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<ILP, CH>);
    if (bps > occ) { printf("ILP %d chunk %d: %d blocks/SM do not fit (%d regs, occupancy %d)\n", ILP, CH, bps, fa.numRegs, occ); return; }
    k<ILP, CH><<<sms * bps, 128>>>(d, n_pairs, n_chunks); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<ILP, CH><<<sms * bps, 128>>>(d, n_pairs, n_chunks); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairticks_per_sched = (double)bps * 4 /*warps*/ / 4 /*schedulers*/ * n_pairs * n_chunks * CH;
    printf("ILP %d chunk %2d, %d blocks/SM (%3d regs): %6.2f clk per pair-tick per scheduler  (%.2fe12 voice-samples/s chip-wide)\n", ILP, CH, bps, fa.numRegs,
           ms * 1e-3 * clk / pairticks_per_sched, 2.0 * sms * bps * 128 * n_pairs * n_chunks * CH / (ms * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *d; cudaMalloc(&d, 4 * sms * 8 * 128);
    for (int bps : {2, 3, 4}) { run<1, 32>(bps, d, sms, khz * 1e3); run<2, 32>(bps, d, sms, khz * 1e3); run<1, 16>(bps, d, sms, khz * 1e3); run<2, 16>(bps, d, sms, khz * 1e3); }
    return 0;
}
