// ubench_int3.cu -- issue rate of the exact operand FORMS k_pdm_v2_ws2 is made of
// (2 registers + immediate / uniform operand), alone and in the kernel's mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int3 ubench_int3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k3(uint32_t *out, uint32_t a0, uint32_t b0, uint32_t m1) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a0 + threadIdx.x * (i + 1);
    uint32_t b = b0 + threadIdx.x, c = b0 * 3 + 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int j = (i + 1) % ILP, l = (i + 2) % ILP, m = (i + 3) % ILP, n = (i + 4) % ILP, o = (i + 5) % ILP;
            if (OP == 0) asm volatile("lop3.b32 %0, %0, 0xFF000000, %1, 0xEA;" : "+r"(x[i]) : "r"(b));   // LOP3 2 reg + imm
            if (OP == 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                      // LOP3 2 reg
            if (OP == 2) asm volatile("prmt.b32 %0, %0, %1, 0x0073;" : "+r"(x[i]) : "r"(b));             // PRMT 2 reg + imm
            if (OP == 3) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(b));            // SHF 2 reg + imm
            if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(b));      // IMAD reg, uniform, reg
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                      // IADD3 2 reg
            if (OP == 6) { asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(b), "r"(c)); }   // IADD3 3 reg
            if (OP == 7) { asm volatile("lop3.b32 %0, %0, 0xFF000000, %1, 0xEA;" : "+r"(x[i]) : "r"(b)); // consumer mix form 2
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(m1), "r"(b));
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[l]) : "r"(m1), "r"(c));
                           asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[m]) : "r"(b));
                           asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[n]) : "r"(c));
                           asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[o]) : "r"(b), "r"(c)); }
            if (OP == 8) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                    // producer mix
                           asm volatile("shl.b32 %0, %1, 13;" : "=r"(x[j]) : "r"(x[i]));
                           asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[l]) : "r"(x[j]));
                           asm volatile("shr.u32 %0, %1, 17;" : "=r"(x[m]) : "r"(x[l]));
                           asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[n]) : "r"(x[m]));
                           asm volatile("and.b32 %0, %1, 0x3FF;" : "=r"(x[o]) : "r"(x[n])); }
            if (OP == 9) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                    // XOR + mad*1
                           asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[j]) : "r"(c)); }
            if (OP == 10) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                   // 2 XOR + mad*1
                            asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[l]) : "r"(c));
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[j]) : "r"(c)); }
            if (OP == 11) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                   // XOR + 2 mad*1
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[l]) : "r"(b));
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[j]) : "r"(c)); }
            if (OP == 12) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(b));   // IMAD(UR) + 2 mad*1
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[l]) : "r"(b));
                            asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[j]) : "r"(c)); }
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
    k3<OP><<<blocks, threads>>>(d, 1, 2, 0xFFFFFFFFu);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k3<OP><<<blocks, threads>>>(d, 1, 2, 0xFFFFFFFFu);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 5.0 * blocks * threads * (double)ITERS * ILP * per_iter;
    double per_s = inst / (ms * 1e-3);
    printf("OP%-2d %-26s %8.2f T thread-instr/s  = %6.1f /clk/SM at %.0f MHz\n", OP, name, per_s / 1e12, per_s / sms / clk_hz, clk_hz / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s, %d SMs, nominal %d MHz\n", p.name, sms, khz / 1000);
    uint32_t *d; cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256);
    run<0>("LOP3 R,R,imm,R", 1, d, sms, clk);
    run<1>("XOR R,R,R", 1, d, sms, clk);
    run<2>("PRMT R,R,imm,R", 1, d, sms, clk);
    run<3>("SHF R,R,imm,R", 1, d, sms, clk);
    run<4>("IMAD R,R,UR,R", 1, d, sms, clk);
    run<5>("IADD3 2 reg", 1, d, sms, clk);
    run<6>("IADD3 3 reg", 1, d, sms, clk);
    run<7>("consumer mix (6)", 6, d, sms, clk);
    run<8>("producer mix (6)", 6, d, sms, clk);
    run<9>("XOR + mad*1", 2, d, sms, clk);
    run<10>("2 XOR + mad*1", 3, d, sms, clk);
    run<11>("XOR + 2 mad*1", 3, d, sms, clk);
    run<12>("IMAD(UR) + 2 mad*1", 3, d, sms, clk);
    return 0;
}
