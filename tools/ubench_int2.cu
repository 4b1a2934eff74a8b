// ubench_int2.cu -- second integer issue-rate microbenchmark: which pipe do the
// "cheap IMAD" forms (IMAD.IADD / IMAD.SHL / IMAD.MOV) use and at what rate, alone and
// next to true IMADs and LOP3s?  ptxas chooses the opcode, so every case must be read
// together with its SASS histogram (tools/ubench_int2_sass.sh prints it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int2 ubench_int2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k2(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a0 + threadIdx.x * (i + 1);
    uint32_t b = b0 + threadIdx.x, c = b0 * 3 + 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int j = (i + 1) % ILP, l = (i + 2) % ILP, m = (i + 3) % ILP;
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[i]) : "r"(b));
            if (OP == 1) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[i]));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, 8, %1;" : "+r"(x[i]) : "r"(b));
            if (OP == 3) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));      // true IMAD + add
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(b)); }
            if (OP == 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));      // true IMAD + 2 adds
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(b));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[l]) : "r"(c)); }
            if (OP == 5) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + IMAD + add
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[l]) : "r"(b)); }
            if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + IMAD + 2 adds
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[l]) : "r"(b));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[m]) : "r"(c)); }
            if (OP == 7) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + shl
                           asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[j])); }
            if (OP == 8) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + 2 shl
                           asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[j]));
                           asm volatile("shl.b32 %0, %0, 5;" : "+r"(x[l])); }
            if (OP == 9) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));  // LOP3 + 3 adds
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(b));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[l]) : "r"(c));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(x[m]) : "r"(b)); }
            if (OP == 10) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c)); // LOP3 + 2 true IMAD
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[l]) : "r"(c), "r"(b)); }
            if (OP == 11) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));     // true IMAD + mad*1 (IMAD.IADD?)
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[j]) : "r"(b)); }
            if (OP == 12) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));     // true IMAD + shl
                            asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[j])); }
            if (OP == 13) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));     // true IMAD + FFMA
                            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c)); }
            if (OP == 14) { asm volatile("prmt.b32 %0, %0, %1, 0x0073;" : "+r"(x[i]) : "r"(b));           // PRMT imm + add
                            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(b)); }
            if (OP == 15) { asm volatile("mad.hi.u32 %0, %0, 256, %1;" : "+r"(x[i]) : "r"(b)); }          // IMAD.HI imm
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
    k2<OP><<<blocks, threads>>>(d, 1, 2);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k2<OP><<<blocks, threads>>>(d, 1, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 5.0 * blocks * threads * (double)ITERS * ILP * per_iter;
    double per_s = inst / (ms * 1e-3);
    printf("OP%-2d %-22s %8.2f T thread-instr/s  = %6.1f /clk/SM at %.0f MHz\n", OP, name, per_s / 1e12, per_s / sms / clk_hz, clk_hz / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s, %d SMs, nominal %d MHz\n", p.name, sms, khz / 1000);
    uint32_t *d; cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256);
    run<0>("mad x*1+b", 1, d, sms, clk);
    run<1>("shl imm", 1, d, sms, clk);
    run<2>("mad x*8+b", 1, d, sms, clk);
    run<3>("IMAD + add", 2, d, sms, clk);
    run<4>("IMAD + 2 add", 3, d, sms, clk);
    run<5>("LOP3 + IMAD + add", 3, d, sms, clk);
    run<6>("LOP3 + IMAD + 2 add", 4, d, sms, clk);
    run<7>("LOP3 + shl", 2, d, sms, clk);
    run<8>("LOP3 + 2 shl", 3, d, sms, clk);
    run<9>("LOP3 + 3 add", 4, d, sms, clk);
    run<10>("LOP3 + 2 IMAD", 3, d, sms, clk);
    run<11>("IMAD + mad*1", 2, d, sms, clk);
    run<12>("IMAD + shl", 2, d, sms, clk);
    run<13>("IMAD + FFMA", 2, d, sms, clk);
    run<14>("PRMT imm + add", 2, d, sms, clk);
    run<15>("mad.hi imm", 1, d, sms, clk);
    return 0;
}
