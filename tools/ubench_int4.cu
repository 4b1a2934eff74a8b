// ubench_int4.cu -- which pipe a two-register IADD3 runs on (SASS checked: cuobjdump -sass must show the named ops).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int4 ubench_int4.cu
// Every op works on its own chain of ILP independent registers; the mixes interleave ops on different chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
#define PRMT(x, b) asm volatile("prmt.b32 %0, %0, %1, 0x7910;" : "+r"(x) : "r"(b))
#define IMAD(x, m, b) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(b))
#define ADD2(x, b) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define SUB2(x, b) asm volatile("sub.u32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define ADD3(x, b, c) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c))
#define ADDSUB3(x, b, c) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c))
template <int OP>
__global__ void k4(uint32_t *out, uint32_t a0, uint32_t b0, uint32_t m1) {
    uint32_t x[ILP], y[ILP], z[ILP], w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a0 + threadIdx.x * (i + 1); y[i] = x[i] * 3; z[i] = x[i] * 5; w[i] = x[i] * 7; }
    uint32_t b = b0 + threadIdx.x, c = b0 * 3 + 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) { ADD2(x[i], b); }
            if (OP == 1) { PRMT(x[i], b); ADD2(y[i], b); }
            if (OP == 2) { IMAD(x[i], m1, b); ADD2(y[i], b); }
            if (OP == 3) { PRMT(x[i], b); IMAD(y[i], m1, b); ADD2(z[i], b); }
            if (OP == 4) { PRMT(x[i], b); IMAD(y[i], m1, b); ADD2(z[i], b); ADD2(w[i], c); }
            if (OP == 5) { PRMT(x[i], b); ADD2(y[i], b); ADD2(z[i], c); ADD2(w[i], b); }
            if (OP == 6) { SUB2(x[i], b); }
            if (OP == 7) { PRMT(x[i], b); ADD3(y[i], b, c); }
            if (OP == 8) { PRMT(x[i], b); SUB2(y[i], b); }
            if (OP == 9) { PRMT(x[i], b); PRMT(y[i], c); IMAD(z[i], m1, b); ADD2(w[i], b); ADD2(x[i], c); ADD2(y[i], b); }   // 2 ALU + 1 FMA + 3 add
            if (OP == 10) { IMAD(x[i], m1, b); IMAD(y[i], m1, c); ADD2(z[i], b); }
            if (OP == 11) { ADDSUB3(x[i], b, c); }
            if (OP == 12) { PRMT(x[i], b); IMAD(y[i], m1, b); ADD2(z[i], b); ADD2(w[i], c); ADDSUB3(x[i], y[i], c); }   // the tick: 2 ALU(+3reg) + FMA + 2 add
        }
    }
    uint32_t s = b + c;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, int per_iter, uint32_t *d, int sms, double clk_hz) {
    int blocks = sms * 8, threads = 256;
    k4<OP><<<blocks, threads>>>(d, 1, 2, 0xFFFFFFFFu);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k4<OP><<<blocks, threads>>>(d, 1, 2, 0xFFFFFFFFu);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 5.0 * blocks * threads * (double)ITERS * ILP * per_iter;
    double per_s = inst / (ms * 1e-3);
    printf("OP%-2d %-44s %8.2f T thread-instr/s = %6.1f /clk/SM = %.2f clk per group per scheduler\n", OP, name, per_s / 1e12, per_s / sms / clk_hz,
           per_iter * 128.0 / (per_s / sms / clk_hz));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s, %d SMs, nominal %d MHz\n", p.name, sms, khz / 1000);
    uint32_t *d; cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256);
    run<0>("IADD3 2r", 1, d, sms, clk);
    run<1>("PRMT + IADD3 2r", 2, d, sms, clk);
    run<2>("IMAD + IADD3 2r", 2, d, sms, clk);
    run<3>("PRMT + IMAD + IADD3 2r", 3, d, sms, clk);
    run<4>("PRMT + IMAD + 2 IADD3 2r", 4, d, sms, clk);
    run<5>("PRMT + 3 IADD3 2r", 4, d, sms, clk);
    run<6>("IADD3 2r (sub)", 1, d, sms, clk);
    run<7>("PRMT + IADD3 3r", 2, d, sms, clk);
    run<8>("PRMT + IADD3 2r (sub)", 2, d, sms, clk);
    run<9>("2 PRMT + IMAD + 3 IADD3 2r", 6, d, sms, clk);
    run<10>("2 IMAD + IADD3 2r", 3, d, sms, clk);
    run<11>("IADD3 3r (a + b - c)", 1, d, sms, clk);
    run<12>("PRMT + IMAD + 2 IADD3 2r + IADD3 3r", 5, d, sms, clk);
    return 0;
}
