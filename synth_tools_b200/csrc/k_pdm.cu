// k_pdm.cu -- sigma-delta / PDM kernels (integer, serial in time, bit-exact).
//
//   k_pdm_v2_*    stm32f103/mod_pdm_pwm.c:101-141 (glide + pdmK_update) with the
//                 control-rate line generator of mod_controlrate.c:28-40
//   k_pdm_v1_*    stm32f103/mod_pdm.c:230-264 (carry-bit PDM, shared dither)
//   k_pdm_raw     stm32f103/pdm.h:13-77 (pdmK_update on an input stream)
//   k_pwm         stm32f103/mod_pdm.c:167-175
//
// Mapping.  The recurrence is nonlinear (quantiser in the loop), so time stays
// serial and the batch axis carries the parallelism.  One lane owns one dither
// bank -- the B channels that share one random word per tick, i.e. one MCU of
// the reference -- so the PRNG is computed once per bank, and keeps every state
// word in registers for the whole segment; 16 ticks of 8-bit duty leave as one
// 128-bit store per channel.
//
// Scheduling.  A warp of 32 banks is a "chain" of G time groups that must run
// in order.  With 65,536 channels in banks of 3 there are 683 chains for 592
// warp schedulers (148 SMs x 4): a plain grid leaves 91 schedulers with two
// chains and finishes in 2T.  The persistent kernels instead run W <= C worker
// warps and cut the C*G work units into W equal contiguous pieces (McNaughton's
// wrap-around rule for preemptive scheduling): a worker runs the HEAD of its
// last chain first, then its whole chains, then the TAIL of its first chain,
// whose head was run first thing by the previous worker.  The hand-over goes
// through the state arrays in L2 plus a release/acquire progress word; by
// construction the producer is always ahead, so the wait is a safety net.
#include "common.cuh"
#include "planar_bulk.cuh"

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

// one tick of glide + pdmK (out_shift 24, dither below bit 24); returns out_a (byte 3 = out_q)
template <int K, int FORM>
__device__ __forceinline__ uint32_t v2_tick_q24(uint32_t &p0, uint32_t v0, uint32_t (&s)[K], uint32_t d, uint32_t m1, uint32_t m2) {
    p0 += v0;                                                      // mod_pdm_pwm.c:101-104
    if constexpr (K == 2 && FORM == 1) {                           // chain LOP3 -> IMAD -> IADD3
        const uint32_t x = s[0] + p0;
        const uint32_t a = lop3_and_or(s[1], d);
        s[0] = imad(a, m1, x);
        s[1] = s[1] + s[0] - a;
        return a;
    } else if constexpr (K == 2 && FORM == 2) {                    // chain LOP3 -> IMAD; u on the ALU pipe
        const uint32_t x = s[0] + p0;
        const uint32_t u = add3(s[1], s[0], p0);
        const uint32_t a = lop3_and_or(s[1], d);
        s[0] = imad(a, m1, x);
        s[1] = imad(a, m2, u);
        return a;
    } else {
        return pdm_step_q24<K>(s, p0, d, m1);
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// McNaughton wrap-around schedule over C chains x G groups for W workers.
struct Sched {
    uint64_t C, G, W, Lg;            // Lg = ceil(C*G / W) >= G
    unsigned long long *flags;       // [C] progress words: (epoch << 32) | groups done
    unsigned long long epoch;
};

struct Segment { uint64_t chain, g0, g1; bool wait, signal; };

// Segment k of worker w in execution order (k < sched_count(s, w)).
__device__ __forceinline__ uint32_t sched_count(const Sched &s, uint64_t w, uint64_t *ca, uint64_t *ga, uint64_t *cb, uint64_t *gb) {
    const uint64_t total = s.C * s.G;
    const uint64_t u0 = w * s.Lg;
    if (u0 >= total) return 0;
    const uint64_t u1 = u0 + s.Lg < total ? u0 + s.Lg : total;
    *ca = u0 / s.G; *ga = u0 % s.G;
    *cb = (u1 - 1) / s.G; *gb = (u1 - 1) % s.G + 1;
    return (uint32_t)(*cb - *ca + 1);
}
__device__ __forceinline__ Segment sched_segment(const Sched &s, uint32_t k, uint32_t nseg, uint64_t ca, uint64_t ga, uint64_t cb, uint64_t gb) {
    if (nseg == 1) return Segment{ca, ga, gb, ga > 0, gb < s.G};
    if (k == 0) return Segment{cb, 0, gb, false, gb < s.G};          // head of the last chain: no dependency
    if (k == nseg - 1) return Segment{ca, ga, s.G, ga > 0, false};   // tail of the first chain
    return Segment{ca + k, 0, s.G, false, false};                    // whole chains in between
}

__device__ __forceinline__ void sched_wait(const Sched &s, const Segment &sg, uint32_t lane) {
    if (!sg.wait) return;
    if (lane == 0) {
        const unsigned long long want = (s.epoch << 32) | sg.g0;
        while (ld_acquire_u64(s.flags + sg.chain) != want) __nanosleep(100);
    }
    __syncwarp();
}
__device__ __forceinline__ void sched_signal(const Sched &s, const Segment &sg, uint32_t lane) {
    if (!sg.signal) return;
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_u64(s.flags + sg.chain, (s.epoch << 32) | sg.g1);
}

// ---------------------------------------------------------------------------
// v2
struct PdmV2Params {
    uint32_t *st;              // SoA [5+K][npad]
    uint64_t npad, n, n_banks;
    uint32_t bank_size;
    uint32_t *prng;            // [n_banks]
    const uint32_t *dither_ext;// [n_banks][F] or null
    const uint32_t *setpoints; // [n_ctl][n] or null
    uint8_t *out;
    uint64_t F;
    uint32_t count0, ctl_div_log, sh, dmask, layout;
    uint32_t m1;               // 0xFFFFFFFF, opaque to the compiler (see pdm_step_q24)
    Sched sched;
};

template <int K, int B>
struct V2Regs {
    uint32_t sp[B], p0[B], v0[B], p1[B], v1[B], s[B][K];
    __device__ __forceinline__ void load(const uint32_t *st, uint64_t npad, uint64_t c0) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const uint32_t *x = st + c0 + j;
            sp[j] = __ldcg(x); p0[j] = __ldcg(x + npad); v0[j] = __ldcg(x + 2 * npad);
            p1[j] = __ldcg(x + 3 * npad); v1[j] = __ldcg(x + 4 * npad);
#pragma unroll
            for (int k = 0; k < K; ++k) s[j][k] = __ldcg(x + (5 + k) * npad);
        }
    }
    __device__ __forceinline__ void store(uint32_t *st, uint64_t npad, uint64_t c0) const {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            uint32_t *x = st + c0 + j;
            __stcg(x, sp[j]); __stcg(x + npad, p0[j]); __stcg(x + 2 * npad, v0[j]);
            __stcg(x + 3 * npad, p1[j]); __stcg(x + 4 * npad, v1[j]);
#pragma unroll
            for (int k = 0; k < K; ++k) __stcg(x + (5 + k) * npad, s[j][k]);
        }
    }
    // mod_pdm_pwm.c:129-137 (line[0] = line[1]) + mod_controlrate.c:28-40
    __device__ __forceinline__ void boundary(const uint32_t *row, uint64_t c0, uint64_t n, uint32_t L) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (row && c0 + j < n) sp[j] = __ldg(row + c0 + j);
            p0[j] = p1[j]; v0[j] = v1[j];
            p1[j] += v1[j] << L;
            v1[j] = (uint32_t)((int32_t)(sp[j] - p1[j]) >> L);
        }
    }
    // 16 ticks -> 4 packed words per channel
    template <bool FASTQ, bool DEXT>
    __device__ __forceinline__ void group(uint32_t &rng, const uint32_t *dext16, uint32_t sh, uint32_t dmask, uint32_t m1, uint32_t (&w)[B][4]) {
        uint32_t dbuf[16];
        if (DEXT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 v = *reinterpret_cast<const uint4 *>(dext16 + i * 4);
                dbuf[4 * i] = v.x; dbuf[4 * i + 1] = v.y; dbuf[4 * i + 2] = v.z; dbuf[4 * i + 3] = v.w;
            }
        }
        uint32_t a[B][4];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t d;
            if (DEXT) d = dbuf[i] & dmask;
            else { rng = xorshift32_step(rng); asm("and.b32 %0, %1, %2;" : "=r"(d) : "r"(rng), "r"(dmask)); }   // mod_pdm_pwm.c:127 (asm: keep the mask out of the per-channel LOP3)
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if (FASTQ) a[j][i & 3] = v2_tick_q24<K, 1>(p0[j], v0[j], s[j], d, m1, 0u);   // :101-104, :108-116
                else { p0[j] += v0[j]; a[j][i & 3] = pdm_step<K>(s[j], p0[j], sh, d); }
                if ((i & 3) == 3)
                    w[j][i >> 2] = FASTQ ? pack_top_bytes(a[j][0], a[j][1], a[j][2], a[j][3])
                                         : pack_low_bytes(a[j][0], a[j][1], a[j][2], a[j][3]);
            }
        }
    }
};

__device__ __forceinline__ uint64_t v2_rows_before(uint32_t count0, uint32_t div, uint64_t t0) {
    // control boundaries at ticks t in [0, t0) with (count0 + t) % div == 0
    const uint64_t first = count0 == 0 ? 0 : div - count0;
    return t0 > first ? 1 + (t0 - first - 1) / div : 0;
}

// One segment [g0, g1) of one thread's B channels.
template <int K, int B, bool FASTQ, bool DEXT>
__device__ __forceinline__ void v2_run_segment(const PdmV2Params &p, V2Regs<K, B> &r, uint32_t &rng, uint64_t c0,
                                               uint64_t bank, uint64_t g0, uint64_t g1) {
    const uint32_t L = p.ctl_div_log, div_mask = (1u << L) - 1u;
    uint32_t cnt = (p.count0 + (uint32_t)(g0 << 4)) & div_mask;
    uint64_t row = v2_rows_before(p.count0, 1u << L, g0 << 4);
    const uint32_t *dext = DEXT ? p.dither_ext + bank * p.F : nullptr;
    const bool tiled = p.layout == CPROC_CUDA_TILED;
    for (uint64_t g = g0; g < g1; ++g) {
        if (cnt == 0) {
            r.boundary(p.setpoints ? p.setpoints + row * p.n : nullptr, c0, p.n, L);
            ++row;
        }
        uint32_t w[B][4];
        r.template group<FASTQ, DEXT>(rng, DEXT ? dext + (g << 4) : nullptr, p.sh, p.dmask, p.m1, w);
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (c0 + j < p.n) {
                uint8_t *dst = tiled ? p.out + ((g * p.n + c0 + j) << 4) : p.out + (c0 + j) * p.F + (g << 4);
                st_v4_stream(dst, make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]));
            }
        }
        cnt = (cnt + 16) & div_mask;
    }
}

// Persistent, McNaughton-scheduled, thread == bank.
template <int K, int B, bool FASTQ>
__global__ void __launch_bounds__(128, 4) k_pdm_v2_persist(const PdmV2Params p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t worker = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (worker >= p.sched.W) return;
    uint64_t ca, ga, cb, gb;
    const uint32_t nseg = sched_count(p.sched, worker, &ca, &ga, &cb, &gb);
    for (uint32_t k = 0; k < nseg; ++k) {
        const Segment sg = sched_segment(p.sched, k, nseg, ca, ga, cb, gb);
        const uint64_t bank = sg.chain * 32 + lane;
        const bool live = bank < p.n_banks;
        sched_wait(p.sched, sg, lane);
        if (live) {
            const uint64_t c0 = bank * B;
            V2Regs<K, B> r;
            r.load(p.st, p.npad, c0);
            uint32_t rng = __ldcg(p.prng + bank);
            v2_run_segment<K, B, FASTQ, false>(p, r, rng, c0, bank, sg.g0, sg.g1);
            r.store(p.st, p.npad, c0);
            __stcg(p.prng + bank, rng);
        }
        sched_signal(p.sched, sg, lane);
    }
}

// Plain grid: thread == bank (TPB) or thread == channel (B == 1, any bank size:
// every thread of a bank replays the bank PRNG), optional external dither.
template <int K, int B, bool TPB, bool FASTQ, bool DEXT>
__global__ void __launch_bounds__(128) k_pdm_v2_simple(const PdmV2Params p) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (TPB ? p.n_banks : p.n_banks * p.bank_size)) return;
    const uint64_t c0 = tid * B;
    const uint64_t bank = TPB ? tid : tid / p.bank_size;
    const bool rng_owner = TPB ? true : (tid % p.bank_size == 0);
    V2Regs<K, B> r;
    r.load(p.st, p.npad, c0);
    uint32_t rng = p.prng[bank];
    v2_run_segment<K, B, FASTQ, DEXT>(p, r, rng, c0, bank, 0, p.F >> 4);
    r.store(p.st, p.npad, c0);
    if (rng_owner && !DEXT) p.prng[bank] = rng;
}

// Any F, any count alignment, byte stores: the conformance path.
template <int K>
__global__ void k_pdm_v2_any(const PdmV2Params p) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_banks * p.bank_size) return;
    const uint64_t bank = c / p.bank_size;
    uint32_t *x = p.st + c;
    uint32_t sp = x[0], p0 = x[p.npad], v0 = x[2 * p.npad], p1 = x[3 * p.npad], v1 = x[4 * p.npad], s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = x[(5 + k) * p.npad];
    uint32_t rng = p.prng[bank];
    const uint32_t L = p.ctl_div_log, div_mask = (1u << L) - 1u;
    uint32_t cnt = p.count0;
    uint64_t row = 0;
    const uint32_t *dext = p.dither_ext ? p.dither_ext + bank * p.F : nullptr;
    for (uint64_t t = 0; t < p.F; ++t) {
        uint32_t d = dext ? dext[t] : (rng = xorshift32_step(rng));
        d &= p.dmask;
        if (cnt == 0) {
            if (p.setpoints && c < p.n) sp = p.setpoints[row * p.n + c];
            p0 = p1; v0 = v1;
            p1 += v1 << L;
            v1 = (uint32_t)((int32_t)(sp - p1) >> L);
            ++row;
        }
        p0 += v0;
        uint32_t q = pdm_step<K>(s, p0, p.sh, d);
        if (c < p.n) {
            uint64_t idx = p.layout == CPROC_CUDA_TILED ? (((t >> 4) * p.n + c) << 4) + (t & 15)
                         : p.layout == CPROC_CUDA_INTERLEAVED ? t * p.n + c
                         : c * p.F + t;
            p.out[idx] = (uint8_t)q;
        }
        cnt = (cnt + 1) & div_mask;
    }
    x[0] = sp; x[p.npad] = p0; x[2 * p.npad] = v0; x[3 * p.npad] = p1; x[4 * p.npad] = v1;
#pragma unroll
    for (int k = 0; k < K; ++k) x[(5 + k) * p.npad] = s[k];
    if (!dext && c % p.bank_size == 0) p.prng[bank] = rng;
}

// Warp-specialised: one block = 32 banks = 32*B channels.  One PRODUCER warp
// (lane == bank) runs the bank PRNGs and stages masked dither words in shared
// memory, 64 ticks ahead, double buffered; B CONSUMER warps (lane == channel)
// run the modulators and read their bank's dither as a 128-bit broadcast load
// per 4 ticks.  The PRNG is still computed once per bank, but the chip now holds
// (B+1)/B warps per 32 channels instead of 1/B, which is what the issue slots
// need: a single warp per scheduler only reaches IPC ~0.4 on this loop
// (profiles/r1_k_pdm_v2_persist_details.txt).  Hand-off: named barriers
// FULL[s] / EMPTY[s] per buffer slot.
#define WS_T 64                  // ticks per dither batch
#define WS_BAR_FULL 1            // barrier ids 1,2
#define WS_BAR_EMPTY 3           // barrier ids 3,4
template <int ID, int N> __device__ __forceinline__ void bar_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void bar_arrive_i() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
// slot s in {0,1}: immediate barrier ids so the CTA only reserves barriers 0..4
template <int BASE, int N> __device__ __forceinline__ void bar_sync_n(uint32_t s) { if (s) bar_sync_i<BASE + 1, N>(); else bar_sync_i<BASE, N>(); }
template <int BASE, int N> __device__ __forceinline__ void bar_arrive_n(uint32_t s) { if (s) bar_arrive_i<BASE + 1, N>(); else bar_arrive_i<BASE, N>(); }

template <int K, int B, bool FASTQ>
__global__ void __launch_bounds__(32 * (B + 1)) k_pdm_v2_ws(const PdmV2Params p) {
    __shared__ __align__(16) uint32_t dbuf[2][WS_T / 4][32][4];
    constexpr int NT = 32 * (B + 1);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t prod_warp = blockIdx.x % (B + 1);      // spread producers over the 4 schedulers
    const uint64_t bank0 = (uint64_t)blockIdx.x * 32;
    const uint64_t batches = p.F / WS_T;
    if (warp == prod_warp) {
        const uint64_t bank = bank0 + lane;
        const bool live = bank < p.n_banks;
        uint32_t rng = live ? p.prng[bank] : 1u;
        const uint32_t dmask = p.dmask;
        for (uint64_t bt = 0; bt < batches; ++bt) {
            const uint32_t s = (uint32_t)bt & 1u;
            if (bt >= 2) bar_sync_n<WS_BAR_EMPTY, NT>(s);            // slot drained by the consumers
#pragma unroll
            for (int q = 0; q < WS_T / 4; ++q) {
                uint4 v;
                rng = xorshift32_step(rng); v.x = rng & dmask;          // mod_pdm_pwm.c:127
                rng = xorshift32_step(rng); v.y = rng & dmask;
                rng = xorshift32_step(rng); v.z = rng & dmask;
                rng = xorshift32_step(rng); v.w = rng & dmask;
                *reinterpret_cast<uint4 *>(&dbuf[s][q][lane][0]) = v;
            }
            __threadfence_block();
            bar_arrive_n<WS_BAR_FULL, NT>(s);
        }
        if (live) p.prng[bank] = rng;
        return;
    }
    const uint32_t cw = warp - (warp > prod_warp ? 1u : 0u);          // consumer index 0..B-1
    const uint32_t cl = cw * 32 + lane;                               // channel within the block
    const uint32_t bl = cl / B;                                       // its bank within the block
    const uint64_t c = bank0 * B + cl;
    const bool live = c < p.n_banks * B;                              // inside the padded SoA rows
    V2Regs<K, 1> r;
    if (live) r.load(p.st, p.npad, c);
    else { r.sp[0] = r.p0[0] = r.v0[0] = r.p1[0] = r.v1[0] = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) r.s[0][k] = 0; }
    const uint32_t L = p.ctl_div_log, div_mask = (1u << L) - 1u;
    uint32_t cnt = p.count0;
    uint64_t row = 0;
    const bool tiled = p.layout == CPROC_CUDA_TILED;
    const bool store = c < p.n;
    const uint32_t m1 = p.m1;
    for (uint64_t bt = 0; bt < batches; ++bt) {
        const uint32_t s = (uint32_t)bt & 1u;
        bar_sync_n<WS_BAR_FULL, NT>(s);
#pragma unroll
        for (int gq = 0; gq < WS_T / 16; ++gq) {
            if (cnt == 0) {
                r.boundary(p.setpoints ? p.setpoints + row * p.n : nullptr, c, p.n, L);
                ++row;
            }
            uint32_t w[4];
            {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const uint4 dv = *reinterpret_cast<const uint4 *>(&dbuf[s][gq * 4 + i4][bl][0]);
                    const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
                    uint32_t a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        r.p0[0] += r.v0[0];                           // :101-104
                        a[i] = FASTQ ? pdm_step_q24<K>(r.s[0], r.p0[0], d[i], m1) : pdm_step<K>(r.s[0], r.p0[0], p.sh, d[i]);
                    }
                    w[i4] = FASTQ ? pack_top_bytes(a[0], a[1], a[2], a[3]) : pack_low_bytes(a[0], a[1], a[2], a[3]);
                }
            }
            if (store) {
                const uint64_t g = bt * (WS_T / 16) + gq;
                uint8_t *dst = tiled ? p.out + ((g * p.n + c) << 4) : p.out + c * p.F + (g << 4);
                st_v4_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
            }
            cnt = (cnt + 16) & div_mask;
        }
        if (bt + 2 < batches) bar_arrive_n<WS_BAR_EMPTY, NT>(s);
    }
    if (live) r.store(p.st, p.npad, c);
}


// ---------------------------------------------------------------------------
// Warp-specialised, second generation.  Same block shape as k_pdm_v2_ws; two
// changes, both about dependency depth (the first-generation kernel sat at 65 %
// of the issue slots with 1.1 eligible warps per scheduler: every warp is a
// serial chain of 4-cycle ALU ops and there are only ~4.6 warps per scheduler
// at 65,536 channels):
//
//  * consumer (order 2): with x = s1 + p and u = x + s2 formed off the critical
//    path, the tick is  a = (s2 & 0xFF000000) | d;  s1' = x - a;  s2' = u - 2a
//    (pdm.h:32-40 with the two subtractions of out_a folded): the loop-carried
//    chain is LOP3 -> IMAD (FORM 2) instead of LOP3 -> IMAD -> IADD -> IADD3.
//  * producer: xorshift32 is a 6-deep chain per tick.  Each producer lane runs
//    P independent chains of its bank's generator, T/P ticks apart; the start
//    of chain j+1 is the state T/P steps after chain j, obtained with a GF(2)
//    jump table (xorshift is linear: M^(T/P) as 4 byte-indexed LUTs in smem).
#ifndef WS2_T_LOG
#define WS2_T_LOG 6
#endif
#define WS2_T (1 << WS2_T_LOG)   // ticks per dither batch
#define WS2_BAR_FULL 1           // barrier ids 1..NS
#define WS2_BAR_EMPTY 5          // barrier ids 5..4+NS  (NS <= 4)
template <int BASE, int N, int NS> __device__ __forceinline__ void bar_sync_slot(uint32_t s) {
    if (NS > 3 && s == 3) bar_sync_i<BASE + 3, N>();
    else if (NS > 2 && s == 2) bar_sync_i<BASE + 2, N>();
    else if (s == 1) bar_sync_i<BASE + 1, N>();
    else bar_sync_i<BASE, N>();
}
template <int BASE, int N, int NS> __device__ __forceinline__ void bar_arrive_slot(uint32_t s) {
    if (NS > 3 && s == 3) bar_arrive_i<BASE + 3, N>();
    else if (NS > 2 && s == 2) bar_arrive_i<BASE + 2, N>();
    else if (s == 1) bar_arrive_i<BASE + 1, N>();
    else bar_arrive_i<BASE, N>();
}
struct PdmV2Ws2Extra { const uint32_t *jump; uint32_t *sm_rank; uint32_t m2; uint32_t k13, k15, k5; };   // jump: [P-1][4][256], M^(T/P * j); k*: 2^13, 2^15, 2^5 or 0

// xorshift32 with its three shifts on the FMA pipe: x << 13 = x * 2^13 (IMAD), x >> 17 = hi32(x * 2^15)
// (IMAD.HI), x << 5 = x * 2^5.  The multipliers arrive through kernel parameters so that ptxas cannot
// turn them back into shifts: the producer is bound by the ALU pipe (SHF, LOP3), the FMA pipe has slack.
__device__ __forceinline__ uint32_t xorshift32_step_fma(uint32_t x, uint32_t k13, uint32_t k15, uint32_t k5) {
    uint32_t t;
    asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(t) : "r"(x), "r"(k13)); x ^= t;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(x), "r"(k15)); x ^= t;
    asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(t) : "r"(x), "r"(k5)); x ^= t;
    return x;
}

__device__ __forceinline__ uint32_t jump_apply(const uint32_t (*jt)[256], uint32_t x) {
    return jt[0][x & 255u] ^ jt[1][(x >> 8) & 255u] ^ jt[2][(x >> 16) & 255u] ^ jt[3][x >> 24];
}

// The producer is a separate (noinline) function on purpose: ptxas balances the
// ALU and FMA pipes by static instruction counts per function; inlined, the
// LOP3-heavy PRNG pushes every add of the consumer loop onto the FMA pipe
// (IMAD.IADD), which then limits the consumer warps.
template <int NT, int P, int NS>
__device__ __noinline__ void ws2_producer_fma(uint32_t (*dbuf)[WS2_T / 4][32][4], const uint32_t (*jt)[4][256], uint32_t *prng_slot,
                                              uint32_t dmask, uint64_t batches, uint32_t lane, uint32_t k13, uint32_t k15, uint32_t k5) {
    constexpr int QC = WS2_T / 4 / P;
    uint32_t x[P];
    x[0] = prng_slot ? *prng_slot : 1u;
    uint32_t s = 0;
    for (uint64_t bt = 0; bt < batches; ++bt) {
#pragma unroll
        for (int j = 1; j < P; ++j) x[j] = jump_apply(jt[j - 1], x[0]);
        if (bt >= NS) bar_sync_slot<WS2_BAR_EMPTY, NT, NS>(s);
#pragma unroll
        for (int q = 0; q < QC; ++q) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                uint4 v;
                x[j] = xorshift32_step_fma(x[j], k13, k15, k5); v.x = x[j] & dmask;      // mod_pdm_pwm.c:127
                x[j] = xorshift32_step_fma(x[j], k13, k15, k5); v.y = x[j] & dmask;
                x[j] = xorshift32_step_fma(x[j], k13, k15, k5); v.z = x[j] & dmask;
                x[j] = xorshift32_step_fma(x[j], k13, k15, k5); v.w = x[j] & dmask;
                *reinterpret_cast<uint4 *>(&dbuf[s][j * QC + q][lane][0]) = v;
            }
        }
        x[0] = x[P - 1];
        __threadfence_block();
        bar_arrive_slot<WS2_BAR_FULL, NT, NS>(s);
        s = (s + 1 == NS) ? 0 : s + 1;
    }
    if (prng_slot) *prng_slot = x[0];
}

template <int NT, int P, int NS>
__device__ __noinline__ void ws2_producer(uint32_t (*dbuf)[WS2_T / 4][32][4], const uint32_t (*jt)[4][256], uint32_t *prng_slot,
                                          uint32_t dmask, uint64_t batches, uint32_t lane) {
    constexpr int QC = WS2_T / 4 / P;                     // uint4 groups per chain per batch
    uint32_t x[P];
    x[0] = prng_slot ? *prng_slot : 1u;
    uint32_t s = 0;
    for (uint64_t bt = 0; bt < batches; ++bt) {
#pragma unroll
        for (int j = 1; j < P; ++j) x[j] = jump_apply(jt[j - 1], x[0]);   // state (T/P)*j ticks ahead
        if (bt >= NS) bar_sync_slot<WS2_BAR_EMPTY, NT, NS>(s);       // slot drained by the consumers
#pragma unroll
        for (int q = 0; q < QC; ++q) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                uint4 v;
                x[j] = xorshift32_step(x[j]); v.x = x[j] & dmask;      // mod_pdm_pwm.c:127
                x[j] = xorshift32_step(x[j]); v.y = x[j] & dmask;
                x[j] = xorshift32_step(x[j]); v.z = x[j] & dmask;
                x[j] = xorshift32_step(x[j]); v.w = x[j] & dmask;
                *reinterpret_cast<uint4 *>(&dbuf[s][j * QC + q][lane][0]) = v;
            }
        }
        x[0] = x[P - 1];                                         // the last chain ends at tick T
        __threadfence_block();
        bar_arrive_slot<WS2_BAR_FULL, NT, NS>(s);
        s = (s + 1 == NS) ? 0 : s + 1;
    }
    if (prng_slot) *prng_slot = x[0];
}

// NS dither slots of WS2_T ticks; barrier ids WS2_BAR_FULL + slot, WS2_BAR_EMPTY + slot.
// Requires count0 % WS2_T == 0 (a control boundary can only fall on a batch start).
template <int K, int B, int FORM, int P, int NS>
__global__ void __launch_bounds__(32 * (B + 1)) k_pdm_v2_ws2(const PdmV2Params p, const PdmV2Ws2Extra ex) {
    __shared__ __align__(16) uint32_t dbuf[NS][WS2_T / 4][32][4];
    __shared__ uint32_t jt[P > 1 ? P - 1 : 1][4][256];
    constexpr int NT = 32 * (B + 1);
    if constexpr (P > 1) {
        for (uint32_t i = threadIdx.x; i < (P - 1) * 1024; i += NT) (&jt[0][0][0])[i] = __ldg(ex.jump + i);
    }
    // Producer placement.  A warp runs on scheduler (%warpid % 4), and the hardware
    // staggers the warp slots of successive blocks on one SM (measured with
    // tools/probe_place.cu: warp 0 of the 1st..5th block sits in slot 0, 5, 10, 15, 16),
    // so a fixed producer warp index piles the producers of an SM onto one or two
    // schedulers.  Instead the producer is the warp that sits on scheduler
    // (rank % 4), rank = arrival order of this block on its SM (never-reset per-SM
    // counter).  Any choice is correct; this one spreads the load.
    __shared__ uint32_t rank_s, smsp_s[B + 1];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rank_s = atomicAdd(ex.sm_rank + smid, 1u);
    }
    if (lane == 0) {
        uint32_t wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        smsp_s[warp] = wid & 3u;
    }
    __syncthreads();
    uint32_t prod_warp = 0;
#pragma unroll
    for (int w = B; w >= 0; --w) if (smsp_s[w] == (rank_s & 3u)) prod_warp = w;
    const uint64_t bank0 = (uint64_t)blockIdx.x * 32;
    const uint64_t batches = p.F / WS2_T;
    if (warp == prod_warp) {
        const uint64_t bank = bank0 + lane;
        ws2_producer<NT, P, NS>(dbuf, jt, bank < p.n_banks ? p.prng + bank : nullptr, p.dmask, batches, lane);
        return;
    }
    const uint32_t cw = warp - (warp > prod_warp ? 1u : 0u);          // consumer index 0..B-1
    const uint32_t cl = cw * 32 + lane;                               // channel within the block
    const uint32_t bl = cl / B;                                       // its bank within the block
    const uint64_t c = bank0 * B + cl;
    const bool live = c < p.n_banks * B;                              // inside the padded SoA rows
    V2Regs<K, 1> r;
    if (live) r.load(p.st, p.npad, c);
    else { r.sp[0] = r.p0[0] = r.v0[0] = r.p1[0] = r.v1[0] = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) r.s[0][k] = 0; }
    const uint32_t L = p.ctl_div_log;
    const uint32_t period = 1u << (L - WS2_T_LOG);                    // batches per control period (L >= WS2_T_LOG)
    uint32_t until = (p.count0 >> WS2_T_LOG) == 0 ? 0 : period - (p.count0 >> WS2_T_LOG);   // batches until the next boundary
    const uint32_t *sp_row = p.setpoints;
    const bool store = c < p.n;
    const uint32_t m1 = p.m1, m2 = ex.m2;
    uint8_t *dst = p.layout == CPROC_CUDA_TILED ? p.out + (c << 4) : p.out + c * p.F;
    const uint64_t dstep = p.layout == CPROC_CUDA_TILED ? p.n << 4 : 16;
    const uint32_t *dbase = &dbuf[0][0][bl][0];
    uint32_t s = 0;
    for (uint64_t bt = 0; bt < batches; ++bt) {
        if (until == 0) {                                             // uniform over the grid
            r.boundary(sp_row, c, p.n, L);
            if (sp_row) sp_row += p.n;
            until = period;
        }
        --until;
        bar_sync_slot<WS2_BAR_FULL, NT, NS>(s);
        const uint32_t *dslot = dbase + s * (WS2_T / 4 * 32 * 4);
#pragma unroll
        for (int gq = 0; gq < WS2_T / 16; ++gq) {
            uint32_t w[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const uint4 dv = *reinterpret_cast<const uint4 *>(dslot + (gq * 4 + i4) * (32 * 4));
                const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
                uint32_t a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = v2_tick_q24<K, FORM>(r.p0[0], r.v0[0], r.s[0], d[i], m1, m2);   // :108-116
                w[i4] = pack_top_bytes(a[0], a[1], a[2], a[3]);
            }
            if (store) st_v4_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
            dst += dstep;
        }
        if (bt + NS < batches) bar_arrive_slot<WS2_BAR_EMPTY, NT, NS>(s);
        s = (s + 1 == NS) ? 0 : s + 1;
    }
    if (live) r.store(p.st, p.npad, c);
}

// ---------------------------------------------------------------------------
// Third generation: the ws2 block (producer warp + B consumer warps per 32 banks) under
// a dynamic schedule.  65,536 channels in banks of 3 are 683 blocks for 148 SMs: a plain
// grid puts 5 blocks on 91 SMs and 4 on the other 57, and the launch lasts as long as
// the 5-block SMs (ncu, ws2: sm__cycles_active max/avg = 1.17).  Here the launch is cut
// into work items (32-bank group g, time slice sl of `bps` dither batches), numbered
// slice-major; a persistent grid of exactly `ctas_per_sm` blocks per SM takes items
// from an atomic counter.  Item (sl, g) continues item (sl-1, g): the channel and PRNG
// state go through the SoA rows in L2 (7 words per channel per slice) and a
// release/acquire progress word per group.  A predecessor always has a smaller item
// number, i.e. is held by a running block or finished, so waiting cannot deadlock.
struct PdmV2Work {
    uint32_t *counter;             // [2]: next item, blocks finished (both reset by the last block to leave)
    unsigned long long *flags;     // [groups]: (epoch << 32) | slices done
    unsigned long long epoch;
    uint32_t groups, slices, bps;  // bps: batches of WS2_T ticks per slice
};

// PL: PLANAR duty rows ([ch][F] bytes) leave through shared memory -- a consumer lane stages 256
// ticks of its channel (four dither batches) in its own 256-byte row and sends them with one bulk
// store (cp.async.bulk), instead of sixteen scattered 16-byte stores.
#define WS3_PL_BATCHES (256 / WS2_T)
// PL: how PLANAR duty rows [ch][F] leave the block.  0: not PLANAR (TILED: 16-byte stores).  1: every consumer lane
// stages 256 ticks of its channel and sends them with one cp.async.bulk.  2: tensor TMA -- a consumer warp fills a box of
// 128 ticks x 32 channels (SWIZZLE_128B: lane r writes 16-tick chunk c at r*128 + ((c ^ (r & 7)) << 4), conflict free),
// one elected lane stores it through the 2-D map over the duty rows; two boxes per warp alternate.
template <int K, int B, int FORM, int P, int NS, int PL>
__global__ void __launch_bounds__(32 * (B + 1)) k_pdm_v2_ws3(const PdmV2Params p, const PdmV2Ws2Extra ex, const PdmV2Work wk, const __grid_constant__ CUtensorMap tm_out) {
    __shared__ __align__(1024) uint8_t pbox[PL == 2 ? B : 1][PL == 2 ? 2 : 1][PL == 2 ? 4096 : 16];
    __shared__ __align__(16) uint32_t dbuf[NS][WS2_T / 4][32][4];
    __shared__ __align__(16) uint8_t prow[PL == 1 ? B * 32 : 1][PL == 1 ? 272 : 16];      // 256 duty bytes per consumer lane (+16: bank-group skew)
    __shared__ uint32_t jt[P > 1 ? P - 1 : 1][4][256];
    constexpr int NT = 32 * (B + 1);
    if constexpr (P > 1) {
        for (uint32_t i = threadIdx.x; i < (P - 1) * 1024; i += NT) (&jt[0][0][0])[i] = __ldg(ex.jump + i);
    }
    __shared__ uint32_t rank_s, smsp_s[B + 1], item_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rank_s = atomicAdd(ex.sm_rank + smid, 1u);
    }
    if (lane == 0) {
        uint32_t wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        smsp_s[warp] = wid & 3u;
    }
    __syncthreads();
    uint32_t prod_warp = 0;                                           // see k_pdm_v2_ws2: the warp on scheduler (rank % 4)
#pragma unroll
    for (int w = B; w >= 0; --w) if (smsp_s[w] == (rank_s & 3u)) prod_warp = w;
    const uint32_t cw = warp - (warp > prod_warp ? 1u : 0u);          // consumer index 0..B-1
    const uint32_t cl = cw * 32 + lane;                               // channel within the block
    const uint32_t bl = cl / B;                                       // its bank within the block
    const uint32_t L = p.ctl_div_log, period = 1u << (L - WS2_T_LOG); // batches per control period (L >= WS2_T_LOG)
    const uint32_t m1 = p.m1, m2 = ex.m2;
    const uint64_t batches_total = p.F / WS2_T;
    const uint32_t total = wk.groups * wk.slices;
    const uint32_t *dbase = &dbuf[0][0][bl][0];
    for (;;) {
        if (threadIdx.x == 0) {
            const uint32_t idx = atomicAdd(wk.counter, 1u);
            item_s = idx;
            if (idx < total) {
                const uint32_t g = idx % wk.groups, sl = idx / wk.groups;
                const unsigned long long want = (wk.epoch << 32) | sl;
                if (sl) while (ld_acquire_u64(wk.flags + g) != want) __nanosleep(64);
            }
        }
        __syncthreads();
        const uint32_t idx = item_s;
        if (idx >= total) break;
        const uint32_t g = idx % wk.groups, sl = idx / wk.groups;
        const uint64_t bt0 = (uint64_t)sl * wk.bps;
        const uint64_t nb = batches_total - bt0 < wk.bps ? batches_total - bt0 : wk.bps;
        const uint64_t bank0 = (uint64_t)g * 32;
        if (warp == prod_warp) {
            const uint64_t bank = bank0 + lane;
            uint32_t *slot = bank < p.n_banks ? p.prng + bank : nullptr;
            if (ex.k13) ws2_producer_fma<NT, P, NS>(dbuf, jt, slot, p.dmask, nb, lane, ex.k13, ex.k15, ex.k5);
            else ws2_producer<NT, P, NS>(dbuf, jt, slot, p.dmask, nb, lane);
        } else {
            const uint64_t c = bank0 * B + cl;
            const bool live = c < p.n_banks * B;                      // inside the padded SoA rows
            V2Regs<K, 1> r;
            if (live) r.load(p.st, p.npad, c);
            else { r.sp[0] = r.p0[0] = r.v0[0] = r.p1[0] = r.v1[0] = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) r.s[0][k] = 0; }
            // control divider at the first batch of the slice (count0 % WS2_T == 0)
            const uint32_t cb = (uint32_t)(((p.count0 >> WS2_T_LOG) + bt0) & (period - 1));
            uint32_t until = cb == 0 ? 0 : period - cb;               // batches until the next boundary
            const uint32_t *sp_row = p.setpoints ? p.setpoints + v2_rows_before(p.count0, 1u << L, bt0 * WS2_T) * p.n : nullptr;
            const bool store = c < p.n;
            uint8_t *dst = p.layout == CPROC_CUDA_TILED ? p.out + ((bt0 * (WS2_T / 16) * p.n + c) << 4) : p.out + c * p.F + bt0 * WS2_T;
            const uint64_t dstep = p.layout == CPROC_CUDA_TILED ? p.n << 4 : 16;
            const uint32_t prow_s = PL == 1 ? (uint32_t)__cvta_generic_to_shared(&prow[PL == 1 ? cl : 0][0]) : 0u;
            const uint32_t pbox_s = PL == 2 ? (uint32_t)__cvta_generic_to_shared(&pbox[PL == 2 ? cw : 0][0][0]) : 0u;
            uint32_t s = 0;
            for (uint64_t bt = 0; bt < nb; ++bt) {
                if (PL == 1 && (bt % WS3_PL_BATCHES) == 0 && bt) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the row has left
                if (PL == 2 && (bt & 1) == 0 && bt >= 4) {            // this box last left two boxes ago
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                if (until == 0) {                                     // uniform over the block
                    r.boundary(sp_row, c, p.n, L);
                    if (sp_row) sp_row += p.n;
                    until = period;
                }
                --until;
                bar_sync_slot<WS2_BAR_FULL, NT, NS>(s);
                const uint32_t *dslot = dbase + s * (WS2_T / 4 * 32 * 4);
#pragma unroll
                for (int gq = 0; gq < WS2_T / 16; ++gq) {
                    uint32_t w[4];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const uint4 dv = *reinterpret_cast<const uint4 *>(dslot + (gq * 4 + i4) * (32 * 4));
                        const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
                        uint32_t a[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = v2_tick_q24<K, FORM>(r.p0[0], r.v0[0], r.s[0], d[i], m1, m2);   // :108-116
                        w[i4] = pack_top_bytes(a[0], a[1], a[2], a[3]);
                    }
                    if (PL == 1) {
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(prow_s + (uint32_t)(bt % WS3_PL_BATCHES) * WS2_T + gq * 16), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    } else if (PL == 2) {
                        const uint32_t ch = ((uint32_t)bt & 1u) * (WS2_T / 16) + gq;                  // 16-tick chunk within the 128-tick box
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pbox_s + (((uint32_t)bt >> 1) & 1u) * 4096u + lane * 128u + ((ch ^ (lane & 7u)) << 4)),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    } else {
                        if (store) st_v4_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
                        dst += dstep;
                    }
                }
                if (PL == 2 && (bt & 1) == 1) {                       // 128 ticks x 32 channels staged (slices hold an even number of batches)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                                     ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"((int32_t)((bt0 + bt - 1) * WS2_T)), "r"((int32_t)(bank0 * B + cw * 32)),
                                       "r"(pbox_s + (((uint32_t)bt >> 1) & 1u) * 4096u) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (PL == 1 && ((bt + 1) % WS3_PL_BATCHES == 0 || bt + 1 == nb)) {        // 256 ticks staged (or the slice ends)
                    const uint32_t nbytes = (uint32_t)((bt % WS3_PL_BATCHES) + 1) * WS2_T;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (store) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(prow_s), "r"(nbytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    dst += nbytes;
                }
                if (bt + NS < nb) bar_arrive_slot<WS2_BAR_EMPTY, NT, NS>(s);
                s = (s + 1 == NS) ? 0 : s + 1;
            }
            if (PL) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (PL == 2) __syncwarp();
            if (live) r.store(p.st, p.npad, c);
        }
        __syncthreads();                                              // every state word of the item is written
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_u64(wk.flags + g, (wk.epoch << 32) | (sl + 1));
        }
    }
    if (threadIdx.x == 0) {
        const uint32_t left = atomicAdd(wk.counter + 1, 1u);
        if (left == gridDim.x - 1) { wk.counter[0] = 0; wk.counter[1] = 0; __threadfence(); }   // ready for the next launch
    }
}

// host: M^steps of xorshift32 as 4 byte-indexed LUTs (linear over GF(2))
static void jump_table_fill(uint32_t *t, uint32_t steps) {
    for (int k = 0; k < 4; ++k)
        for (uint32_t b = 0; b < 256; ++b) {
            uint32_t x = b << (8 * k);
            for (uint32_t i = 0; i < steps; ++i) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; }
            t[k * 256 + b] = x;
        }
}

static int jump_tables(cproc_cuda_ctx *ctx, int P, const uint32_t **out) {
    if (P < 2) { *out = nullptr; return 0; }
    uint32_t *&d = ctx->d_jump[P];
    if (!d) {
        std::vector<uint32_t> h((size_t)(P - 1) * 1024);
        for (int j = 1; j < P; ++j) jump_table_fill(h.data() + (size_t)(j - 1) * 1024, (uint32_t)(WS2_T / P) * j);
        CK(ctx, cudaMalloc(&d, h.size() * 4));
        CK(ctx, cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = d;
    return 0;
}

// Workers for C chains: as many warps as there are chains, up to `per_sm` warps
// on each of the SMs; always W <= C so that Lg >= G.
static int sched_setup(cproc_cuda_batch *b, Sched *s, uint64_t C, uint64_t G, int per_sm) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (b->n_flags < C) {
        if (b->d_flags) cudaFree(b->d_flags);
        b->d_flags = nullptr; b->n_flags = 0;
        CK(ctx, cudaMalloc(&b->d_flags, sizeof(unsigned long long) * C));
        CK(ctx, cudaMemsetAsync(b->d_flags, 0, sizeof(unsigned long long) * C, ctx->stream));
        b->n_flags = C;
    }
    uint64_t W = (uint64_t)ctx->n_sm * per_sm;
    if (W > C) W = C;
    s->C = C; s->G = G; s->W = W; s->Lg = ceil_div_u64(C * G, W);
    s->flags = b->d_flags;
    s->epoch = ++b->epoch;
    return 0;
}

// pdm_persist: 0 never, 2 always, 1 auto -- only where a plain grid would leave
// the warp schedulers badly unbalanced (between 1 and 1.6 chains per scheduler;
// measured on B200: 683 chains -> +20 %, 1024 chains -> -9 %).
static bool persist_wanted(const cproc_cuda_ctx *ctx, uint64_t C) {
    if (ctx->pdm_persist == 0) return false;
    if (ctx->pdm_persist >= 2) return true;
    const uint64_t smsp = 4ull * ctx->n_sm;
    return C > smsp && C * 10 < smsp * 16;
}

template <int K, bool FASTQ>
static int launch_v2_order(cproc_cuda_batch *b, PdmV2Params &p, bool fast, bool tpb, bool dext) {
    cproc_cuda_ctx *ctx = b->ctx;
    const int blk = ctx->pdm_block;
    if (!fast) {
        k_pdm_v2_any<K><<<(unsigned)ceil_div_u64(p.npad, 128), 128, 0, ctx->stream>>>(p);
        return 0;
    }
    const uint64_t C = ceil_div_u64(p.n_banks, 32);
    if (FASTQ && tpb && !dext && ctx->pdm_ws >= 2 && (p.F % WS2_T) == 0 && (p.count0 % WS2_T) == 0 && p.ctl_div_log >= WS2_T_LOG) {
        const unsigned grid = (unsigned)C;
        PdmV2Ws2Extra ex;
        ex.m2 = 0xFFFFFFFEu;
        ex.k13 = ctx->pdm_prng_fma ? 1u << 13 : 0u; ex.k15 = 1u << 15; ex.k5 = 1u << 5;
        const int P = ctx->pdm_chains, form = (K == 2) ? ctx->pdm_form : 0;
        int rc = jump_tables(ctx, P, &ex.jump);
        if (rc) return rc;
        if (!ctx->d_sm_rank) {
            CK(ctx, cudaMalloc(&ctx->d_sm_rank, 1024 * sizeof(uint32_t)));
            CK(ctx, cudaMemsetAsync(ctx->d_sm_rank, 0, 1024 * sizeof(uint32_t), ctx->stream));
        }
        ex.sm_rank = ctx->d_sm_rank;
        // dynamic schedule: only where the plain grid is unbalanced (more blocks than one
        // even layer over the SMs) and the run is long enough to be cut into slices
        const uint64_t batches = p.F / WS2_T;
        const uint64_t ctas = (uint64_t)ctx->n_sm * ctx->pdm_ctas_per_sm;
        const bool dyn = ctx->pdm_ws >= 3 && K == 2 && p.bank_size == 3 && C > ctas && C < 65536 && batches >= 2 * (uint64_t)ctx->pdm_slice_batches;
        if constexpr (K == 2) if (dyn) {
            if (b->n_flags < C) {
                if (b->d_flags) cudaFree(b->d_flags);
                b->d_flags = nullptr; b->n_flags = 0;
                CK(ctx, cudaMalloc(&b->d_flags, sizeof(unsigned long long) * C));
                CK(ctx, cudaMemsetAsync(b->d_flags, 0, sizeof(unsigned long long) * C, ctx->stream));
                b->n_flags = C;
            }
            if (!ctx->d_work) {
                CK(ctx, cudaMalloc(&ctx->d_work, 2 * sizeof(uint32_t)));
                CK(ctx, cudaMemsetAsync(ctx->d_work, 0, 2 * sizeof(uint32_t), ctx->stream));
            }
            PdmV2Work wk;
            wk.counter = ctx->d_work; wk.flags = b->d_flags; wk.epoch = ++b->epoch;
            wk.groups = (uint32_t)C; wk.bps = (uint32_t)ctx->pdm_slice_batches;
            wk.slices = (uint32_t)ceil_div_u64(batches, wk.bps);
            const int f = form;
            CUtensorMap tm0;
            memset(&tm0, 0, sizeof(tm0));
#define WS3_GO(FF, PP) k_pdm_v2_ws3<2, 3, FF, PP, 2, 0><<<(unsigned)ctas, 128, 0, ctx->stream>>>(p, ex, wk, tm0)
            // PLANAR rows leave through shared memory (default variant only): 64-byte aligned rows, slices of whole 256-tick stages
            if (ctx->pdm_planar_bulk && p.layout == CPROC_CUDA_PLANAR && f == 1 && P == 2 && (p.F % 16) == 0 && ((uintptr_t)p.out & 15) == 0 &&
                ((uint64_t)wk.bps % WS3_PL_BATCHES) == 0) {
                // tensor TMA needs whole 128-tick boxes in every slice (bps is a multiple of 4 batches; the last slice ends at F)
                if (ctx->pdm_planar_bulk >= 2 && (p.F % 128) == 0 && pbulk::encode_rows_u8(&tm0, p.out, p.F, p.n))
                    k_pdm_v2_ws3<2, 3, 1, 2, 2, 2><<<(unsigned)ctas, 128, 0, ctx->stream>>>(p, ex, wk, tm0);
                else
                    k_pdm_v2_ws3<2, 3, 1, 2, 2, 1><<<(unsigned)ctas, 128, 0, ctx->stream>>>(p, ex, wk, tm0);
            } else
            if (P == 4) { if (f == 1) WS3_GO(1, 4); else if (f == 2) WS3_GO(2, 4); else WS3_GO(0, 4); }
            else if (P == 2) { if (f == 1) WS3_GO(1, 2); else if (f == 2) WS3_GO(2, 2); else WS3_GO(0, 2); }
            else { if (f == 1) WS3_GO(1, 1); else if (f == 2) WS3_GO(2, 1); else WS3_GO(0, 1); }
#undef WS3_GO
            return 0;
        }
#define WS2_NS4 (WS2_T_LOG > 6 ? 2 : 4)      /* four slots of 128-tick batches exceed the static shared-memory limit */
#define WS2_GO(BB, FF, PP) do { if (ctx->pdm_slots >= 4) k_pdm_v2_ws2<K, BB, FF, PP, WS2_NS4><<<grid, 32 * (BB + 1), 0, ctx->stream>>>(p, ex); \
                                else k_pdm_v2_ws2<K, BB, FF, PP, 2><<<grid, 32 * (BB + 1), 0, ctx->stream>>>(p, ex); } while (0)
#define WS2_P(BB, FF) do { if (P == 4) WS2_GO(BB, FF, 4); else if (P == 2) WS2_GO(BB, FF, 2); else WS2_GO(BB, FF, 1); } while (0)
#define WS2_F(BB) do { if constexpr (K == 2) { if (form == 1) WS2_P(BB, 1); else if (form == 2) WS2_P(BB, 2); else WS2_P(BB, 0); } \
                       else WS2_P(BB, 0); } while (0)
#ifdef PDM_DEV_FAST
        WS2_F(3);
#else
        switch (p.bank_size) {
        case 1: WS2_F(1); break;
        case 2: WS2_F(2); break;
        case 3: WS2_F(3); break;
        default: WS2_F(4); break;
        }
#endif
#undef WS2_F
#undef WS2_P
#undef WS2_GO
        return 0;
    }
    if (tpb && !dext && ctx->pdm_ws && (p.F % WS_T) == 0) {
        const unsigned grid = (unsigned)C;
        switch (p.bank_size) {
        case 1: k_pdm_v2_ws<K, 1, FASTQ><<<grid, 64, 0, ctx->stream>>>(p); break;
        case 2: k_pdm_v2_ws<K, 2, FASTQ><<<grid, 96, 0, ctx->stream>>>(p); break;
        case 3: k_pdm_v2_ws<K, 3, FASTQ><<<grid, 128, 0, ctx->stream>>>(p); break;
        default: k_pdm_v2_ws<K, 4, FASTQ><<<grid, 160, 0, ctx->stream>>>(p); break;
        }
        return 0;
    }
    if (tpb && !dext && persist_wanted(ctx, C)) {
        int rc = sched_setup(b, &p.sched, C, p.F >> 4, 4 * ctx->pdm_warps_per_smsp);
        if (rc) return rc;
        const unsigned grid = (unsigned)ceil_div_u64(p.sched.W, 4);
        switch (p.bank_size) {
        case 1: k_pdm_v2_persist<K, 1, FASTQ><<<grid, 128, 0, ctx->stream>>>(p); break;
        case 2: k_pdm_v2_persist<K, 2, FASTQ><<<grid, 128, 0, ctx->stream>>>(p); break;
        case 3: k_pdm_v2_persist<K, 3, FASTQ><<<grid, 128, 0, ctx->stream>>>(p); break;
        default: k_pdm_v2_persist<K, 4, FASTQ><<<grid, 128, 0, ctx->stream>>>(p); break;
        }
        return 0;
    }
#define V2_SIMPLE(BB, TT) do { \
        const unsigned grid = (unsigned)ceil_div_u64((TT) ? p.n_banks : p.n_banks * p.bank_size, blk); \
        if (dext) k_pdm_v2_simple<K, BB, TT, FASTQ, true><<<grid, blk, 0, ctx->stream>>>(p); \
        else k_pdm_v2_simple<K, BB, TT, FASTQ, false><<<grid, blk, 0, ctx->stream>>>(p); } while (0)
    if (tpb) {
        switch (p.bank_size) {
        case 1: V2_SIMPLE(1, true); break;
        case 2: V2_SIMPLE(2, true); break;
        case 3: V2_SIMPLE(3, true); break;
        default: V2_SIMPLE(4, true); break;
        }
    } else V2_SIMPLE(1, false);
#undef V2_SIMPLE
    return 0;
}

int launch_pdm_v2(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    const cproc_cuda_config &c = b->cfg;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: out is NULL");
    if (io->layout == CPROC_CUDA_TILED && (F & 15)) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: TILED needs F %% 16 == 0");
    if (F == 0) return 0;
    uint32_t div = 1u << c.ctl_div_log;
    if (io->ctl) {
        uint64_t rows = 0;
        {   // rows consumed = control boundaries met in [count, count+F)
            uint64_t first = b->count == 0 ? 0 : div - b->count;
            rows = F > first ? 1 + (F - first - 1) / div : 0;
        }
        if (rows > io->n_ctl) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: run crosses %llu control boundaries but ctl has %u rows", (unsigned long long)rows, io->n_ctl);
    }
    PdmV2Params p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n; p.n_banks = b->n_banks; p.bank_size = c.bank_size;
    p.prng = b->d_prng; p.dither_ext = (const uint32_t *)io->in2; p.setpoints = (const uint32_t *)io->ctl;
    p.out = (uint8_t *)io->out; p.F = F; p.count0 = b->count; p.ctl_div_log = c.ctl_div_log; p.sh = c.out_shift;
    p.dmask = c.dither_mask; p.layout = io->layout; p.m1 = 0xFFFFFFFFu;
    p.sched = Sched{};
    const bool aligned_ptr = ((uintptr_t)io->out & 15) == 0 && (!io->in2 || ((uintptr_t)io->in2 & 15) == 0);
    const bool fast = (F & 15) == 0 && (b->count & 15) == 0 && c.ctl_div_log >= 4 && aligned_ptr &&
                      (io->layout == CPROC_CUDA_TILED || io->layout == CPROC_CUDA_PLANAR);
    const bool tpb = ctx->pdm_tpb && c.bank_size <= 4;
    const bool fastq = c.out_shift == 24 && (c.dither_mask & 0xFF000000u) == 0;
    const bool dext = io->in2 != nullptr;
    int rc;
#define V2_ORDER(KK) (fastq ? launch_v2_order<KK, true>(b, p, fast, tpb, dext) : launch_v2_order<KK, false>(b, p, fast, tpb, dext))
#ifdef PDM_DEV_FAST     // development builds (SASS inspection): the firmware's order only
    rc = V2_ORDER(2);
#else
    switch (c.order) {
    case 1: rc = V2_ORDER(1); break;
    case 2: rc = V2_ORDER(2); break;
    case 3: rc = V2_ORDER(3); break;
    default: rc = V2_ORDER(4); break;
    }
#endif
#undef V2_ORDER
    if (rc) return rc;
    CK_LAUNCH(ctx, "k_pdm_v2");
    b->count = (uint32_t)((b->count + F) & (div - 1));
    return 0;
}

// ---------------------------------------------------------------------------
// v1: carry-bit PDM.  32 ticks -> one packed word (sample t at bit t & 31).
struct PdmV1Params {
    uint32_t *st;              // SoA [2][npad]: setpoint, accu
    uint64_t npad, n, n_banks;
    uint32_t bank_size;
    uint32_t *prng;
    const uint32_t *dither_ext;
    uint32_t *out;
    uint64_t F;
    uint32_t dmask, layout;
    const uint32_t *jump16;    // M^16 of xorshift32 as 4 byte-indexed LUTs (two-chain PRNG) or null
    Sched sched;
};

// accu += x with the carry out shifted into the LSB of bits (ARM: adds + adc;
// mod_pdm.c:236-240 uses rrx, i.e. the MSB: the final __brev gives the same
// time order LSB-first).
__device__ __forceinline__ void add_carry_shift(uint32_t &accu, uint32_t &bits, uint32_t x) {
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(accu), "+r"(bits) : "r"(x));
}

// 32 ticks of B channels -> one packed word per channel
template <int B, bool DEXT>
__device__ __forceinline__ void v1_word(const uint32_t (&sp)[B], uint32_t (&acc)[B], uint32_t &rng, const uint32_t *dext32,
                                        uint32_t dmask, uint32_t (&wv)[B]) {
    uint32_t bits[B];
#pragma unroll
    for (int j = 0; j < B; ++j) bits[j] = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t d;
        if (DEXT) d = dext32[i] & dmask;
        else { rng = xorshift32_step(rng); d = rng & dmask; }       // mod_pdm.c:261
#pragma unroll
        for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + d);   // :235-240
    }
#pragma unroll
    for (int j = 0; j < B; ++j) wv[j] = __brev(bits[j]);
}

// The same word with the bank's generator run as TWO interleaved chains: xorshift32 is a 6-deep
// dependent chain per tick and there are fewer than two warps per scheduler at the C2 shape, so
// one chain leaves the issue slots idle.  Chain A produces ticks 0..15, chain B -- started 16
// steps ahead with the GF(2) jump table M^16 -- ticks 16..31 into registers; the accumulators
// then consume them in time order.
template <int B>
__device__ __forceinline__ void v1_word2(const uint32_t (&sp)[B], uint32_t (&acc)[B], uint32_t &rng, const uint32_t (*jt)[256],
                                         uint32_t dmask, uint32_t (&wv)[B]) {
    uint32_t bits[B], late[16];
#pragma unroll
    for (int j = 0; j < B; ++j) bits[j] = 0;
    uint32_t xa = rng, xb = jump_apply(jt, rng);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        xa = xorshift32_step(xa);
        xb = xorshift32_step(xb);
        late[i] = xb & dmask;
        const uint32_t d = xa & dmask;                                             // mod_pdm.c:261
#pragma unroll
        for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + d);   // :235-240
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + late[i]);
    }
    rng = xb;
#pragma unroll
    for (int j = 0; j < B; ++j) wv[j] = __brev(bits[j]);
}

// Software-pipelined form of the two-chain word: while the accumulators consume the 32 dither
// values of word w (registers `cur`), the generator produces the 32 values of word w+1 (`nxt`) --
// chain A ticks 0..15, chain B (started with the jump table M^16) ticks 16..31.  The accumulate
// chain (one dependent add per tick) is then never paced by the 6-deep xorshift chain.
// GEN = false for the last word of a group: nothing is generated beyond it.
template <int B, bool GEN>
__device__ __forceinline__ void v1_word_pipe(const uint32_t (&sp)[B], uint32_t (&acc)[B], const uint32_t (&cur)[32], uint32_t (&nxt)[32],
                                             uint32_t &rng, const uint32_t (*jt)[256], uint32_t dmask, uint32_t (&wv)[B]) {
    uint32_t bits[B];
#pragma unroll
    for (int j = 0; j < B; ++j) bits[j] = 0;
    uint32_t xa = rng, xb = GEN ? jump_apply(jt, rng) : 0u;          // rng = generator state at the start of word w+1
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (GEN) {
            xa = xorshift32_step(xa); nxt[i] = xa & dmask;            // mod_pdm.c:261
            xb = xorshift32_step(xb); nxt[16 + i] = xb & dmask;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + cur[2 * i]);       // :235-240
#pragma unroll
        for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + cur[2 * i + 1]);
    }
    if (GEN) rng = xb;                                               // state at the start of word w+2
#pragma unroll
    for (int j = 0; j < B; ++j) wv[j] = __brev(bits[j]);
}

// words [w0, w1) of one thread's B channels
template <int B, bool DEXT>
__device__ __forceinline__ void v1_run_segment(const PdmV1Params &p, const uint32_t (&sp)[B], uint32_t (&acc)[B], uint32_t &rng,
                                               uint64_t c0, uint64_t bank, uint64_t w0, uint64_t w1, const uint32_t (*jt)[256] = nullptr) {
    const uint64_t words = p.F >> 5;
    const uint32_t *dext = DEXT ? p.dither_ext + bank * p.F : nullptr;
    if (p.layout == CPROC_CUDA_TILED) {
        for (uint64_t g = w0; g < w1; g += 4) {
            uint32_t q[4][B];
            if (!DEXT && jt) {
                // four words per TILED group; the first word's dither is generated up front, the last
                // word of the group generates nothing (the next group starts over: one un-overlapped
                // word in four keeps the generator state exact at every group boundary)
                uint32_t da[32], db[32];
                {
                    uint32_t xa = rng, xb = jump_apply(jt, rng);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { xa = xorshift32_step(xa); da[i] = xa & p.dmask; xb = xorshift32_step(xb); da[16 + i] = xb & p.dmask; }
                    rng = xb;
                }
                v1_word_pipe<B, true>(sp, acc, da, db, rng, jt, p.dmask, q[0]);
                v1_word_pipe<B, true>(sp, acc, db, da, rng, jt, p.dmask, q[1]);
                v1_word_pipe<B, true>(sp, acc, da, db, rng, jt, p.dmask, q[2]);
                v1_word_pipe<B, false>(sp, acc, db, da, rng, jt, p.dmask, q[3]);
            } else
#pragma unroll
            for (int k = 0; k < 4; ++k) v1_word<B, DEXT>(sp, acc, rng, DEXT ? dext + ((g + k) << 5) : nullptr, p.dmask, q[k]);
#pragma unroll
            for (int j = 0; j < B; ++j)
                if (c0 + j < p.n) st_v4_stream(p.out + (((g >> 2) * p.n + c0 + j) << 2), make_uint4(q[0][j], q[1][j], q[2][j], q[3][j]));
        }
    } else {
        const bool il = p.layout == CPROC_CUDA_INTERLEAVED;
        for (uint64_t g = w0; g < w1; ++g) {
            uint32_t wv[B];
            v1_word<B, DEXT>(sp, acc, rng, DEXT ? dext + (g << 5) : nullptr, p.dmask, wv);
#pragma unroll
            for (int j = 0; j < B; ++j)
                if (c0 + j < p.n) __stcs(p.out + (il ? g * p.n + c0 + j : (c0 + j) * words + g), wv[j]);
        }
    }
}

template <int B>
__global__ void __launch_bounds__(128, 4) k_pdm_v1_persist(const PdmV1Params p, uint32_t unit_words) {
    __shared__ uint32_t jt[4][256];
    if (p.jump16) {
        for (uint32_t k = threadIdx.x; k < 1024; k += blockDim.x) (&jt[0][0])[k] = __ldg(p.jump16 + k);
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t worker = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (worker >= p.sched.W) return;
    uint64_t ca, ga, cb, gb;
    const uint32_t nseg = sched_count(p.sched, worker, &ca, &ga, &cb, &gb);
    for (uint32_t k = 0; k < nseg; ++k) {
        const Segment sg = sched_segment(p.sched, k, nseg, ca, ga, cb, gb);
        const uint64_t bank = sg.chain * 32 + lane;
        sched_wait(p.sched, sg, lane);
        if (bank < p.n_banks) {
            const uint64_t c0 = bank * B;
            uint32_t sp[B], acc[B];
#pragma unroll
            for (int j = 0; j < B; ++j) { sp[j] = __ldcg(p.st + c0 + j); acc[j] = __ldcg(p.st + p.npad + c0 + j); }
            uint32_t rng = __ldcg(p.prng + bank);
            v1_run_segment<B, false>(p, sp, acc, rng, c0, bank, sg.g0 * unit_words, sg.g1 * unit_words, p.jump16 ? jt : nullptr);
#pragma unroll
            for (int j = 0; j < B; ++j) __stcg(p.st + p.npad + c0 + j, acc[j]);
            __stcg(p.prng + bank, rng);
        }
        sched_signal(p.sched, sg, lane);
    }
}

template <int B, bool TPB, bool DEXT>
__global__ void __launch_bounds__(128) k_pdm_v1_simple(const PdmV1Params p) {
    __shared__ uint32_t jt[(TPB && !DEXT) ? 4 : 1][256];
    const bool two = TPB && !DEXT && p.jump16 != nullptr;
    if (two) {
        for (uint32_t k = threadIdx.x; k < 1024; k += blockDim.x) (&jt[0][0])[k] = __ldg(p.jump16 + k);
        __syncthreads();
    }
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (TPB ? p.n_banks : p.n_banks * p.bank_size)) return;
    const uint64_t c0 = tid * B;
    const uint64_t bank = TPB ? tid : tid / p.bank_size;
    const bool rng_owner = TPB ? true : (tid % p.bank_size == 0);
    uint32_t sp[B], acc[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { sp[j] = p.st[c0 + j]; acc[j] = p.st[p.npad + c0 + j]; }
    uint32_t rng = p.prng[bank];
    v1_run_segment<B, DEXT>(p, sp, acc, rng, c0, bank, 0, p.F >> 5, two ? jt : nullptr);
#pragma unroll
    for (int j = 0; j < B; ++j) p.st[p.npad + c0 + j] = acc[j];
    if (rng_owner && !DEXT) p.prng[bank] = rng;
}

int launch_pdm_v1(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    const cproc_cuda_config &c = b->cfg;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: out is NULL");
    if (F & 31) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: F must be a multiple of 32 (packed bit output)");
    if (io->layout == CPROC_CUDA_TILED && (F & 127)) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: TILED needs F %% 128 == 0");
    if (io->layout == CPROC_CUDA_TILED && ((uintptr_t)io->out & 15)) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: TILED needs a 16-byte aligned out");
    if (F == 0) return 0;
    PdmV1Params p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n; p.n_banks = b->n_banks; p.bank_size = c.bank_size;
    p.prng = b->d_prng; p.dither_ext = (const uint32_t *)io->in2; p.out = (uint32_t *)io->out; p.F = F;
    p.dmask = c.dither_mask; p.layout = io->layout;
    p.sched = Sched{};
    p.jump16 = nullptr;
    const int blk = ctx->pdm_block;
    const bool dext = io->in2 != nullptr;
    const bool tpb = ctx->pdm_tpb && c.bank_size <= 4;
    const uint64_t C = ceil_div_u64(p.n_banks, 32);
    if (ctx->pdm_v1_chains == 2 && tpb && !dext && io->layout == CPROC_CUDA_TILED) {
        if (!ctx->d_jump16) {                                 // M^16 of xorshift32 as 4 byte-indexed LUTs
            std::vector<uint32_t> h(1024);
            jump_table_fill(h.data(), 16);
            CK(ctx, cudaMalloc(&ctx->d_jump16, h.size() * 4));
            CK(ctx, cudaMemcpyAsync(ctx->d_jump16, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(ctx, cudaStreamSynchronize(ctx->stream));
        }
        p.jump16 = ctx->d_jump16;
    }
    if (tpb && !dext && (F & 127) == 0 && persist_wanted(ctx, C)) {
        int rc = sched_setup(b, &p.sched, C, F >> 7, 4 * ctx->pdm_warps_per_smsp);   // unit = 128 ticks = 4 words
        if (rc) return rc;
        const unsigned grid = (unsigned)ceil_div_u64(p.sched.W, 4);
        switch (c.bank_size) {
        case 1: k_pdm_v1_persist<1><<<grid, 128, 0, ctx->stream>>>(p, 4); break;
        case 2: k_pdm_v1_persist<2><<<grid, 128, 0, ctx->stream>>>(p, 4); break;
        case 3: k_pdm_v1_persist<3><<<grid, 128, 0, ctx->stream>>>(p, 4); break;
        default: k_pdm_v1_persist<4><<<grid, 128, 0, ctx->stream>>>(p, 4); break;
        }
    } else {
#define V1_SIMPLE(BB, TT) do { \
        const unsigned grid = (unsigned)ceil_div_u64((TT) ? p.n_banks : p.n_banks * p.bank_size, blk); \
        if (dext) k_pdm_v1_simple<BB, TT, true><<<grid, blk, 0, ctx->stream>>>(p); \
        else k_pdm_v1_simple<BB, TT, false><<<grid, blk, 0, ctx->stream>>>(p); } while (0)
        if (tpb) {
            switch (c.bank_size) {
            case 1: V1_SIMPLE(1, true); break;
            case 2: V1_SIMPLE(2, true); break;
            case 3: V1_SIMPLE(3, true); break;
            default: V1_SIMPLE(4, true); break;
            }
        } else V1_SIMPLE(1, false);
#undef V1_SIMPLE
    }
    CK_LAUNCH(ctx, "k_pdm_v1");
    return 0;
}

// ---------------------------------------------------------------------------
// raw pdmK_update over an input stream (pdm.h), uint32 quantiser output.
struct PdmRawParams {
    uint32_t *st;               // SoA [K][npad]
    const uint32_t *param;      // SoA [1][npad] constant input
    uint64_t npad, n;
    const uint32_t *in;         // [inst][F] / [F][inst] or null
    const uint32_t *dither;     // [F] or null
    uint32_t *out;
    uint64_t F;
    uint32_t sh, layout;
};

template <int K>
__global__ void k_pdm_raw(const PdmRawParams p) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n) return;
    uint32_t s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = p.st[k * p.npad + c];
    const uint32_t cst = p.param[c];
    const bool il = p.layout == CPROC_CUDA_INTERLEAVED;
    for (uint64_t t = 0; t < p.F; ++t) {
        uint64_t idx = il ? t * p.n + c : c * p.F + t;
        uint32_t x = p.in ? p.in[idx] : cst;
        uint32_t d = p.dither ? p.dither[t] : 0u;
        p.out[idx] = pdm_step<K>(s, x, p.sh, d);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) p.st[k * p.npad + c] = s[k];
}

// PLANAR streams through the bulk-staged template (planar_bulk.cuh): the per-thread row walk of
// k_pdm_raw touches 32 different lines per warp instruction.
template <int K, int HAS_IN>
struct PdmRawOp {
    static constexpr int NIN = HAS_IN;
    uint32_t *st; const uint32_t *param; const uint32_t *dither;
    uint64_t npad; uint32_t sh;
    uint32_t s[K], cst;
    __device__ __forceinline__ void load(uint64_t i) {
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = st[k * npad + i];
        cst = param[i];
    }
    __device__ __forceinline__ void store(uint64_t i) {
#pragma unroll
        for (int k = 0; k < K; ++k) st[k * npad + i] = s[k];
    }
    __device__ __forceinline__ uint32_t tick(uint32_t x, uint64_t t) {
        return pdm_step<K>(s, HAS_IN ? x : cst, sh, dither ? __ldg(dither + t) : 0u);    // pdm.h:13-77
    }
};

template <int K>
static int launch_pdm_raw_bulk(cproc_cuda_ctx *ctx, const PdmRawParams &p) {
    if (p.in) {
        PdmRawOp<K, 1> op; op.st = p.st; op.param = p.param; op.dither = p.dither; op.npad = p.npad; op.sh = p.sh;
        return pbulk::launch<64, 3>(ctx, op, p.in, p.out, p.n, p.F);
    }
    PdmRawOp<K, 0> op; op.st = p.st; op.param = p.param; op.dither = p.dither; op.npad = p.npad; op.sh = p.sh;
    return pbulk::launch<64, 3>(ctx, op, p.out, p.out, p.n, p.F);
}

int launch_pdm(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm: out is NULL");
    if (io->layout == CPROC_CUDA_TILED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm: TILED layout not supported");
    if (F == 0) return 0;
    PdmRawParams p;
    p.st = b->d_state; p.param = b->d_param; p.npad = b->npad; p.n = b->n;
    p.in = (const uint32_t *)io->in; p.dither = (const uint32_t *)io->in2; p.out = (uint32_t *)io->out;
    p.F = F; p.sh = b->cfg.out_shift; p.layout = io->layout;
    unsigned grid = (unsigned)ceil_div_u64(p.n, 128);
    if (ctx->planar_bulk && io->layout == CPROC_CUDA_PLANAR && pbulk::usable(F, p.in, p.out)) {
        int rc;
        switch (b->cfg.order) {
        case 1: rc = launch_pdm_raw_bulk<1>(ctx, p); break;
        case 2: rc = launch_pdm_raw_bulk<2>(ctx, p); break;
        case 3: rc = launch_pdm_raw_bulk<3>(ctx, p); break;
        default: rc = launch_pdm_raw_bulk<4>(ctx, p); break;
        }
        if (rc) return rc;
        CK_LAUNCH(ctx, "k_pdm_raw (bulk)");
        return 0;
    }
    switch (b->cfg.order) {
    case 1: k_pdm_raw<1><<<grid, 128, 0, ctx->stream>>>(p); break;
    case 2: k_pdm_raw<2><<<grid, 128, 0, ctx->stream>>>(p); break;
    case 3: k_pdm_raw<3><<<grid, 128, 0, ctx->stream>>>(p); break;
    default: k_pdm_raw<4><<<grid, 128, 0, ctx->stream>>>(p); break;
    }
    CK_LAUNCH(ctx, "k_pdm_raw");
    return 0;
}

// ---------------------------------------------------------------------------
// pwm_update (mod_pdm.c:167-175): nonlinear phase feedback, serial.
__global__ void k_pwm(uint32_t *phase, const uint32_t *speed, uint64_t n, uint64_t F, uint8_t *out, uint32_t layout) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint32_t ph = phase[c];
    const uint32_t sp = speed[c];
    for (uint64_t t = 0; t < F; ++t) {
        uint64_t idx = layout == CPROC_CUDA_INTERLEAVED ? t * n + c : c * F + t;
        out[idx] = (uint8_t)(ph >> 16);
        ph = (ph + sp + (ph >> 9)) & 0xFFFFFFu;
    }
    phase[c] = ph;
}

// PLANAR duty bytes through the bulk-staged template: one "frame" of the template is a 32-bit word
// = four ticks (duty is 8 bit), so a lane's 64-word tile is 256 ticks of its channel.
struct PwmOp {
    static constexpr int NIN = 0;
    uint32_t *phase; const uint32_t *speed;
    uint32_t ph, sp;
    __device__ __forceinline__ void load(uint64_t i) { ph = phase[i]; sp = speed[i]; }
    __device__ __forceinline__ void store(uint64_t i) { phase[i] = ph; }
    __device__ __forceinline__ uint32_t tick(uint32_t, uint64_t) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w |= ((ph >> 16) & 0xFFu) << (8 * k);              // mod_pdm.c:170
            ph = (ph + sp + (ph >> 9)) & 0xFFFFFFu;            // :171-174
        }
        return w;
    }
};

// TILED [F/16][ch][16]: one 128-bit store per 16 ticks, like the PDM v2 duty stream
__global__ void __launch_bounds__(128) k_pwm_tiled(uint32_t *phase, const uint32_t *speed, uint64_t n, uint64_t F, uint8_t *out) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint32_t ph = phase[c];
    const uint32_t sp = speed[c];
    for (uint64_t g = 0; g < F / 16; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            w[q] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[q] |= ((ph >> 16) & 0xFFu) << (8 * k);
                ph = (ph + sp + (ph >> 9)) & 0xFFFFFFu;
            }
        }
        st_v4_stream(out + ((g * n + c) << 4), make_uint4(w[0], w[1], w[2], w[3]));
    }
    phase[c] = ph;
}

int launch_pwm(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pwm: out is NULL");
    if (io->layout == CPROC_CUDA_TILED && ((F & 15) || ((uintptr_t)io->out & 15))) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pwm: TILED needs F %% 16 == 0 and a 16-byte aligned buffer");
    if (F == 0) return 0;
    if (io->layout == CPROC_CUDA_TILED) {
        k_pwm_tiled<<<(unsigned)ceil_div_u64(b->n, 128), 128, 0, ctx->stream>>>(b->d_state, b->d_param, b->n, F, (uint8_t *)io->out);
        CK_LAUNCH(ctx, "k_pwm_tiled");
        return 0;
    }
    if (ctx->planar_bulk && io->layout == CPROC_CUDA_PLANAR && F % 16 == 0 && pbulk::usable(F / 4, io->out, io->out)) {
        PwmOp op; op.phase = b->d_state; op.speed = b->d_param; op.ph = 0; op.sp = 0;
        int rc = pbulk::launch<64, 3>(ctx, op, (const uint32_t *)io->out, (uint32_t *)io->out, b->n, F / 4);
        if (rc) return rc;
        CK_LAUNCH(ctx, "k_pwm (bulk)");
        return 0;
    }
    k_pwm<<<(unsigned)ceil_div_u64(b->n, 128), 128, 0, ctx->stream>>>(b->d_state, b->d_param, b->n, F, (uint8_t *)io->out, io->layout);
    CK_LAUNCH(ctx, "k_pwm");
    return 0;
}
