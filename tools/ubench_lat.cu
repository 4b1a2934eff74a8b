// ubench_lat.cu -- dependent-issue latency of FFMA, FFMA2, FADD2, FMUL2 and of the SVF recurrence of k_xvoice_mix2
// (one warp per SM sub-partition, one chain).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_lat ubench_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(float *out, long long *cyc, float a0, float b0) {
    float x = a0 + threadIdx.x, a = a0 * 0.999f, b = b0;
    unsigned long long p = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x * 3.f);
    const unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    unsigned long long nlp = p, bp = pa, f2 = pb, nf2 = pb ^ 0x8000000080000000ull, nq2 = pa ^ 0x8000000080000000ull;
    float lp = x, sbp = a;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
        if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(pa), "l"(pb));
        if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pb));
        if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pa));
        if (OP == 4) {   // the packed SVF recurrence: 4 dependent FFMA2 per tick
            unsigned long long hp;
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(nlp) : "l"(nf2), "l"(bp));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hp) : "l"(pa), "l"(pb), "l"(nlp));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp) : "l"(nq2), "l"(bp));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(bp) : "l"(f2), "l"(hp));
        }
        if (OP == 5) {   // the scalar SVF recurrence
            float hp;
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(lp) : "f"(a), "f"(sbp));
            asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(hp) : "f"(a), "f"(b), "f"(lp));
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(hp) : "f"(b), "f"(sbp));
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(sbp) : "f"(a), "f"(hp));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + lp + sbp + __uint_as_float((uint32_t)p) + __uint_as_float((uint32_t)nlp) + __uint_as_float((uint32_t)bp);
}
template <int OP> void run(const char *name, int per_iter, int warps, float *d, long long *c) {
    k<OP><<<1, 32 * warps>>>(d, c, 1.0f, 1e-3f); cudaDeviceSynchronize();
    k<OP><<<1, 32 * warps>>>(d, c, 1.0f, 1e-3f); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %d warps/SM: %6.2f clk per op, %6.2f clk per iteration\n", name, warps, (double)h / N / per_iter, (double)h / N);
}
int main() {
    float *d; long long *c; cudaMalloc(&d, 4 * 1024); cudaMalloc(&c, 8);
    for (int w : {4, 8, 12, 16}) {
        run<0>("FFMA chain", 1, w, d, c);
        run<1>("FFMA2 chain", 1, w, d, c);
        run<2>("FADD2 chain", 1, w, d, c);
        run<3>("FMUL2 chain", 1, w, d, c);
        run<4>("packed SVF tick (4 dependent FFMA2)", 4, w, d, c);
        run<5>("scalar SVF tick (4 dependent FFMA)", 4, w, d, c);
    }
    return 0;
}
