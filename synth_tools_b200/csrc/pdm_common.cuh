// pdm_common.cuh -- arithmetic shared by the PDM kernels (k_pdm.cu: v1, pdmK on a stream, pwm; k_pdm_v2.cu: v2).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// shared helpers

template <int K>
__device__ __forceinline__ uint32_t pdm_step(uint32_t (&s)[K], uint32_t in, uint32_t sh, uint32_t d) {
    // pdm.h:13-24 / 32-40 / 48-57 / 67-77
    uint32_t q = s[K - 1] >> sh;
    uint32_t a = (q << sh) + (K == 1 ? 0u : d);
    s[0] += in - a;
#pragma unroll
    for (int k = 1; k < K; ++k) s[k] += s[k - 1] - a;
    return q;
}

// out_shift == 24 and dither below bit 24: out_a = (s & 0xFF000000) | d is the
// same number as (out_q << 24) + d (no carries), and out_q is its top byte.
// `m1` is the constant 0xFFFFFFFF passed through a kernel parameter so that
// ptxas keeps `in - a` as an IMAD (in + a * m1) on the FMA pipe: the loop is
// bound by the ALU pipe (LOP3 / PRMT / IADD3 share 64 lanes/clk/SM, measured in
// tools/ubench_int.cu), the FMA pipe has slack.
template <int K>
__device__ __forceinline__ uint32_t pdm_step_q24(uint32_t (&s)[K], uint32_t in, uint32_t d, uint32_t m1) {
    uint32_t a;
    if (K == 1) a = s[0] & 0xFF000000u;
    else asm("lop3.b32 %0, %1, 0xFF000000, %2, 0xEA;" : "=r"(a) : "r"(s[K - 1]), "r"(d));   // (s & M) | d, one LOP3
    uint32_t t;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(m1), "r"(in));                 // in - a
    s[0] += t;
#pragma unroll
    for (int k = 1; k < K; ++k) s[k] += s[k - 1] - a;
    return a;       // byte 3 = out_q
}

// byte 3 of four words -> one little-endian word (3 PRMT per 4 samples)
__device__ __forceinline__ uint32_t pack_top_bytes(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
    uint32_t lo = __byte_perm(a0, a1, 0x0073);
    uint32_t hi = __byte_perm(a2, a3, 0x0073);
    return __byte_perm(lo, hi, 0x5410);
}
// low bytes of four words -> one word
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
    uint32_t lo = __byte_perm(q0, q1, 0x0040);
    uint32_t hi = __byte_perm(q2, q3, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t s, uint32_t d) {   // (s & 0xFF000000) | d
    uint32_t a;
    asm("lop3.b32 %0, %1, 0xFF000000, %2, 0xEA;" : "=r"(a) : "r"(s), "r"(d));
    return a;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t t;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c));
    return t;
}

// Three-input add as one IADD3 (ALU pipe).  Left to itself ptxas turns every add of
// the tick loop into IMAD.IADD and the FMA pipe (one warp instruction per 2 clk per
// scheduler, like the ALU pipe: tools/ubench_int3.cu) becomes the limiter.
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t t;
    asm("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(t) : "r"(a), "r"(b), "r"(c));
    return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// xorshift32 is linear over GF(2): M^steps as 4 byte-indexed LUTs
__device__ __forceinline__ uint32_t jump_apply(const uint32_t (*jt)[256], uint32_t x) {
    return jt[0][x & 255u] ^ jt[1][(x >> 8) & 255u] ^ jt[2][(x >> 16) & 255u] ^ jt[3][x >> 24];
}
static inline void jump_table_fill(uint32_t *t, uint32_t steps) {
    for (int k = 0; k < 4; ++k)
        for (uint32_t b = 0; b < 256; ++b) {
            uint32_t x = b << (8 * k);
            for (uint32_t i = 0; i < steps; ++i) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; }
            t[k * 256 + b] = x;
        }
}
