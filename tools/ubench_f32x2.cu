// ubench_f32x2.cu -- Blackwell packed fp32 (fma.rn.f32x2 -> FFMA2, add/mul.f32x2): issue rate against scalar FFMA, alone and
// beside ALU-pipe work (SASS checked: cuobjdump -sass must show FFMA2 / FADD2 / FMUL2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
#define FFMA(x, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b))
#define FFMA2(x, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(a), "l"(b))
#define FADD2(x, a) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a))
#define FMUL2(x, a) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a))
#define FFMA3(x, y, z) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x) : "f"(y), "f"(z))
#define FFMA23(x, y, z) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(y), "l"(z))
#define LOP(x, b) asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(x) : "r"(b))
#define FMNMX(x, a) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(a))
#define I2F(f, i) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(i))
template <int OP>
__global__ void k(float *out, float a0, float b0, uint32_t u0) {
    float x[ILP], m[ILP], y3[ILP], z3[ILP]; unsigned long long p[ILP], p3[ILP], q3[ILP]; uint32_t u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { y3[i] = a0 * (0.9f + 0.01f * i) + threadIdx.x * 1e-6f; z3[i] = b0 * (i + 1); p3[i] = ((unsigned long long)__float_as_uint(y3[i]) << 32) | __float_as_uint(y3[i] * 0.99f); q3[i] = ((unsigned long long)__float_as_uint(z3[i]) << 32) | __float_as_uint(z3[i] * 3.f); }
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a0 + threadIdx.x * (i + 1); m[i] = x[i] * 0.5f; p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] * 3.f); u[i] = u0 + i + threadIdx.x; }
    const float a = a0 * 0.999f, b = b0;
    const unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) { FFMA(x[i], a, b); }
            if (OP == 1) { FFMA2(p[i], pa, pb); }
            if (OP == 2) { FADD2(p[i], pb); }
            if (OP == 3) { FMUL2(p[i], pa); }
            if (OP == 4) { FFMA2(p[i], pa, pb); LOP(u[i], u0); }
            if (OP == 5) { FFMA(x[i], a, b); LOP(u[i], u0); }
            if (OP == 6) { FFMA2(p[i], pa, pb); FFMA(x[i], a, b); }
            if (OP == 7) { FFMA2(p[i], pa, pb); FMNMX(m[i], a); }
            if (OP == 8) { FFMA2(p[i], pa, pb); LOP(u[i], u0); FMNMX(m[i], a); }
            if (OP == 9) { FMNMX(m[i], a); }
            if (OP == 10) { I2F(m[i], u[i]); LOP(u[i], u0); }
            if (OP == 11) { FFMA2(p[i], pa, pb); FFMA2(p[i], pb, pa); LOP(u[i], u0); }
            if (OP == 12) { FFMA3(x[i], y3[i], z3[i]); }
            if (OP == 13) { FFMA23(p[i], p3[i], q3[i]); }
            if (OP == 14) { FFMA3(x[i], y3[i], z3[i]); LOP(u[i], u0); }
            if (OP == 15) { FFMA3(x[i], y3[i], z3[(i + 1) % ILP]); FFMA3(m[i], y3[(i + 3) % ILP], z3[i]); }
            if (OP == 16) { FFMA23(p[i], p3[i], q3[i]); LOP(u[i], u0); }
            if (OP == 17) { FFMA23(p[i], p3[i], q3[i]); FFMA3(x[i], y3[i], z3[i]); FFMA3(m[i], y3[(i + 3) % ILP], z3[i]); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += y3[i] + z3[i] + __uint_as_float((uint32_t)p3[i]) + __uint_as_float((uint32_t)q3[i]) + x[i] + m[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, int per_iter, float *d, int sms, double clk_hz) {
    int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1.0f, 1e-3f, 77u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(d, 1.0f, 1e-3f, 77u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 5.0 * blocks * threads * (double)ITERS * ILP * per_iter;
    double per_s = inst / (ms * 1e-3);
    printf("OP%-2d %-36s %8.2f T thread-instr/s = %6.1f /clk/SM = %.2f clk per group per scheduler\n", OP, name, per_s / 1e12, per_s / sms / clk_hz, per_iter * 128.0 / (per_s / sms / clk_hz));
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s, %d SMs, nominal %d MHz\n", pr.name, sms, khz / 1000);
    float *d; cudaMalloc(&d, sizeof(float) * sms * 8 * 256);
    run<0>("FFMA", 1, d, sms, clk);
    run<1>("FFMA2", 1, d, sms, clk);
    run<2>("FADD2", 1, d, sms, clk);
    run<3>("FMUL2", 1, d, sms, clk);
    run<4>("FFMA2 + LOP3", 2, d, sms, clk);
    run<5>("FFMA + LOP3", 2, d, sms, clk);
    run<6>("FFMA2 + FFMA", 2, d, sms, clk);
    run<7>("FFMA2 + FMNMX", 2, d, sms, clk);
    run<8>("FFMA2 + LOP3 + FMNMX", 3, d, sms, clk);
    run<9>("FMNMX", 1, d, sms, clk);
    run<10>("I2F + LOP3", 2, d, sms, clk);
    run<11>("2 FFMA2 (dependent) + LOP3", 3, d, sms, clk);
    run<12>("FFMA, 3 distinct registers", 1, d, sms, clk);
    run<13>("FFMA2, 3 distinct register pairs", 1, d, sms, clk);
    run<14>("FFMA 3-reg + LOP3", 2, d, sms, clk);
    run<15>("2 FFMA 3-reg (crossed operands)", 2, d, sms, clk);
    run<16>("FFMA2 3-pair + LOP3", 2, d, sms, clk);
    run<17>("FFMA2 3-pair + 2 FFMA 3-reg", 3, d, sms, clk);
    return 0;
}
