// ubench_int.cu -- per-SM issue rates of the integer instructions the PDM /
// phasor kernels are made of (the "INT issue roofline" denominators of
// SURVEY.md section 8d must be measured on the box, not assumed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int ubench_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a0 + threadIdx.x * (i + 1);
    uint32_t b = b0 + threadIdx.x, c = b0 * 3 + 1;
    asm volatile(".reg .pred p;");
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                    // IADD3
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));      // IMAD
            if (OP == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));  // SHF
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));        // PRMT
            if (OP == 5) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                   // IADD + IMAD mix
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 6) { asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %2, %2, %2;" : "+r"(x[i]), "+r"(b), "+r"(c)); } // carry chain
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));      // FFMA
            if (OP == 8) asm volatile("shr.u32 %0, %0, 3;" : "+r"(x[i]));                               // SHF imm
            if (OP == 9) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                   // IADD + FFMA mix
                           asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 10) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(x[i]));                          // I2F
            if (OP == 11) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + IMAD mix
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 12) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + SHF mix
                            asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 13) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + PRMT mix
                            asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + IADD3 mix
                            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(i + 1) % ILP]) : "r"(b)); }
            if (OP == 15) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + FFMA mix
                            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[(i + 1) % ILP]) : "r"(b), "r"(c)); }
            if (OP == 16) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + 2x IADD3
                            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(i + 1) % ILP]) : "r"(b));
                            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(i + 2) % ILP]) : "r"(c)); }
            if (OP == 17) asm volatile("add.f32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                     // FADD
            if (OP == 18) asm volatile("setp.lt.f32 p, %0, %1;\n\tselp.b32 %0, %1, %2, p;" : "+r"(x[i]) : "r"(b), "r"(c)); // FSETP+SEL
        }
    }
    uint32_t s = b + c;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int per_iter, uint32_t *d, int sms, double clk_hz) {
    int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1, 2);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(d, 1, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 5.0 * blocks * threads * (double)ITERS * ILP * per_iter;
    double per_s = inst / (ms * 1e-3);
    printf("%-14s %8.2f T thread-instr/s  = %6.1f /clk/SM at %.0f MHz\n", name, per_s / 1e12, per_s / sms / clk_hz, clk_hz / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s, %d SMs, nominal %d MHz\n", p.name, sms, khz / 1000);
    uint32_t *d; cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256);
    run<0>("IADD3", 1, d, sms, clk);
    run<1>("LOP3", 1, d, sms, clk);
    run<2>("IMAD", 1, d, sms, clk);
    run<3>("SHF", 1, d, sms, clk);
    run<8>("SHF.imm", 1, d, sms, clk);
    run<4>("PRMT", 1, d, sms, clk);
    run<5>("IADD3+IMAD", 2, d, sms, clk);
    run<6>("ADD.CC+ADDC", 2, d, sms, clk);
    run<7>("FFMA", 1, d, sms, clk);
    run<9>("IADD3+FFMA", 2, d, sms, clk);
    run<10>("I2F", 1, d, sms, clk);
    run<11>("LOP3+IMAD", 2, d, sms, clk);
    run<12>("LOP3+SHF", 2, d, sms, clk);
    run<13>("LOP3+PRMT", 2, d, sms, clk);
    run<14>("LOP3+IADD3", 2, d, sms, clk);
    run<15>("LOP3+FFMA", 2, d, sms, clk);
    run<16>("LOP3+2IADD3", 3, d, sms, clk);
    run<17>("FADD", 1, d, sms, clk);
    run<18>("FSETP+SEL", 2, d, sms, clk);
    return 0;
}
