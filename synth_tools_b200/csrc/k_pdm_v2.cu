// k_pdm_v2.cu -- the v2 PDM modulator: stm32f103/mod_pdm_pwm.c:101-141 (glide + pdmK_update, one
// dither word per tick shared by the channels of a bank) with the control-rate line generator of
// mod_controlrate.c:28-40.  Integer, serial in time, bit-exact.
//
//   k_pdm_v2_ws4     the render kernel: any order 1..4, any bank size, TILED or PLANAR duty
//   k_pdm_v2_simple  thread per bank / per channel (external dither, runs that are not whole batches)
//   k_pdm_v2_any     any F, any counter alignment, any out_shift, byte stores: the conformance path
//
// Mapping.  The recurrence is nonlinear (quantiser in the loop): time stays serial, the channel axis
// carries the parallelism -- 65,536 channels are only 3.5 warps per scheduler, so the kernel is built
// around keeping those few warps issuing.
//
// k_pdm_v2_ws4.  A work GROUP is 32*CW consecutive channels (CW consumer warps, lane == channel; CW =
// bank size for banks of 1..4 -- a group is then exactly 32 banks -- and 4 for larger banks, where a
// group touches at most 27 banks, not aligned to its channel range).  One PRODUCER warp (lane ==
// bank) generates the masked dither of the group's banks a batch of T ticks ahead into shared
// memory, double buffered; the consumers keep the 5+K state words of their channel in registers and
// emit one 128-bit store per 16 ticks.  Hand-off: named barriers FULL / EMPTY per slot.
//
// Dither generator (uc_tools random_u32 stand-in: xorshift32, parity unpinned -- DESIGN.md 2): the recurrence itself,
// two interleaved chains per producer lane T/2 ticks apart (the second start state by a GF(2) jump table: xorshift
// is linear), 7 instructions per bank-tick.  Measured and rejected: a table-driven generator
// (state 8 ticks on and the eight masked 16-bit dither values as XORs of byte-indexed rows, 16-bit packed dither
// unpacked for free by the consumer's PRMT): 2.9 instead of 7 instructions per bank-tick, but the random 128-bit
// shared-memory look-ups cost ~37 clk per warp instruction and the kernel came out 2 % slower.
//
// Schedule.  683 groups for 148 SMs: a plain grid leaves 91 SMs with 5 blocks and 57 with 4.  The
// launch is cut into items (group g, time slice sl of `bps` batches), numbered slice-major; a
// persistent grid of `ctas_per_sm` blocks per SM takes items from an atomic counter.  Item (sl, g)
// continues (sl-1, g): channel state goes through the SoA rows in L2, the generator state through a
// per-group scratch row, ordered by a release/acquire progress word per group.  A predecessor always
// has a smaller item number, i.e. is held by a running block or finished: waiting cannot deadlock.
//
// Generator state is double buffered (read p.prng, the bank's owner writes p.prng_out, the host swaps
// the pointers after the launch): a bank larger than a block, or one that straddles two groups, is
// replayed by every block that needs it, and none of them may see another's write-back.
#include "common.cuh"
#include "pdm_common.cuh"
#include "planar_bulk.cuh"

struct PdmV2Params {
    uint32_t *st;              // SoA [5+K][npad]
    uint64_t npad, n, n_banks;
    uint32_t bank_size;
    const uint32_t *prng;      // [n_banks] generator state at the start of the run
    uint32_t *prng_out;        // [n_banks] ... at its end (a different buffer)
    const uint32_t *dither_ext;// [n_banks][F] or null
    const uint32_t *setpoints; // [n_ctl][n] or null
    uint8_t *out;
    uint64_t F;
    uint32_t count0, ctl_div_log, sh, dmask, layout;
    uint32_t m1;               // 0xFFFFFFFF, opaque to the compiler (see pdm_step_q24)
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
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            sp[j] = p0[j] = v0[j] = p1[j] = v1[j] = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) s[j][k] = 0;
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
                p0[j] += v0[j];                                                                  // :101-104
                if (FASTQ) a[j][i & 3] = pdm_step_q24<K>(s[j], p0[j], d, m1);                    // :108-116
                else a[j][i & 3] = pdm_step<K>(s[j], p0[j], sh, d);
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

// Plain grid: thread == bank (TPB) or thread == channel (B == 1, any bank size: every thread of a
// bank replays the bank's generator), optional external dither.  F % 16 == 0.
template <int K, int B, bool TPB, bool FASTQ, bool DEXT>
__global__ void __launch_bounds__(128) k_pdm_v2_simple(const PdmV2Params p) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (TPB ? p.n_banks : p.n_banks * p.bank_size)) return;
    const uint64_t c0 = tid * B;
    const uint64_t bank = TPB ? tid : tid / p.bank_size;
    const bool rng_owner = TPB ? true : (tid % p.bank_size == 0);
    V2Regs<K, B> r;
    r.load(p.st, p.npad, c0);
    uint32_t rng = __ldg(p.prng + bank);
    const uint32_t L = p.ctl_div_log, div_mask = (1u << L) - 1u;
    uint32_t cnt = p.count0;
    uint64_t row = 0;
    const uint32_t *dext = DEXT ? p.dither_ext + bank * p.F : nullptr;
    const bool tiled = p.layout == CPROC_CUDA_TILED;
    for (uint64_t g = 0; g < (p.F >> 4); ++g) {
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
    r.store(p.st, p.npad, c0);
    if (rng_owner && !DEXT) p.prng_out[bank] = rng;
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
    uint32_t rng = __ldg(p.prng + bank);
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
    if (!dext && c % p.bank_size == 0) p.prng_out[bank] = rng;
}

// ---------------------------------------------------------------------------
// k_pdm_v2_ws4
#define WS4_BAR_FULL 1            // barrier ids 1, 2
#define WS4_BAR_EMPTY 3           // barrier ids 3, 4
template <int ID, int N> __device__ __forceinline__ void bar_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void bar_arrive_i() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int BASE, int N> __device__ __forceinline__ void bar_sync_slot(uint32_t s) { if (s) bar_sync_i<BASE + 1, N>(); else bar_sync_i<BASE, N>(); }
template <int BASE, int N> __device__ __forceinline__ void bar_arrive_slot(uint32_t s) { if (s) bar_arrive_i<BASE + 1, N>(); else bar_arrive_i<BASE, N>(); }

struct PdmV2Work {
    uint32_t *counter;             // [2]: next item, blocks finished (both reset by the last block to leave)
    unsigned long long *flags;     // [groups]: (epoch << 32) | slices done
    unsigned long long epoch;
    uint32_t groups, slices, bps;  // bps: batches of T ticks per slice
    uint32_t *prng_g;              // [groups][32]: generator state of the group's banks between its slices
    uint32_t *sm_rank;             // per-SM block arrival counters (producer placement)
    const uint32_t *jump;          // M^(T/2) as [4][256]
};

// one tick of glide + pdmK for out_shift 24 and dither below bit 24; returns out_a (byte 3 = out_q)
template <int K>
__device__ __forceinline__ uint32_t ws4_tick(uint32_t &p0, uint32_t v0, uint32_t (&s)[K], uint32_t d, uint32_t m1) {
    p0 += v0;                                                      // mod_pdm_pwm.c:101-104
    if constexpr (K == 2) {                                        // chain LOP3 -> IMAD -> IADD3
        const uint32_t x = s[0] + p0;
        const uint32_t a = lop3_and_or(s[1], d);
        s[0] = imad(a, m1, x);                                     // pdm.h:32-40
        s[1] = s[1] + s[0] - a;
        return a;
    } else {
        return pdm_step_q24<K>(s, p0, d, m1);                      // pdm.h:13-24 / 48-77
    }
}

// The producer is a separate (noinline) function on purpose: ptxas balances the ALU and FMA pipes by
// static instruction counts per function; inlined, the LOP3-heavy generator pushes every add of the
// consumer loop onto the FMA pipe (IMAD.IADD), which then limits the consumer warps.
template <int NT, int TLOG>
__device__ __noinline__ uint32_t ws4_producer_arith(uint4 *dbuf, const uint32_t (*jt)[256], uint32_t x0, uint32_t dmask, uint64_t batches, uint32_t lane) {
    constexpr int QW = (1 << TLOG) / 4, QC = QW / 2;             // uint4 rows per slot / per chain
    uint32_t xa = x0, s = 0;
    for (uint64_t bt = 0; bt < batches; ++bt) {
        uint32_t xb = jump_apply(jt, xa);                         // state T/2 ticks ahead
        if (bt >= 2) bar_sync_slot<WS4_BAR_EMPTY, NT>(s);         // slot drained by the consumers
        uint4 *slot = dbuf + s * (QW * 32) + lane;
#pragma unroll 4
        for (int q = 0; q < QC; ++q) {
            uint4 va, vb;
            xa = xorshift32_step(xa); va.x = xa & dmask;           // mod_pdm_pwm.c:127
            xb = xorshift32_step(xb); vb.x = xb & dmask;
            xa = xorshift32_step(xa); va.y = xa & dmask;
            xb = xorshift32_step(xb); vb.y = xb & dmask;
            xa = xorshift32_step(xa); va.z = xa & dmask;
            xb = xorshift32_step(xb); vb.z = xb & dmask;
            xa = xorshift32_step(xa); va.w = xa & dmask;
            xb = xorshift32_step(xb); vb.w = xb & dmask;
            slot[q * 32] = va;
            slot[(QC + q) * 32] = vb;
        }
        xa = xb;                                                   // the second chain ends at tick T
        __threadfence_block();
        bar_arrive_slot<WS4_BAR_FULL, NT>(s);
        s ^= 1u;
    }
    return xa;
}

constexpr size_t ws4_smem_bytes(int CW, int TLOG, int PL) {      // boxes, two dither slots, the jump table, alignment slack
    return (size_t)(PL == 2 ? CW * 2 * 4096 : PL == 1 ? CW * 4096 : 0) + 2 * ((1u << TLOG) / 4) * 32 * 16 + 4096 + 1024;
}

// PL 1: INTERLEAVED duty [tick][ch] -- the order the firmware's ISR produces (one byte per channel per tick).  The consumer warps write
// their byte of every tick into ONE [128 ticks][32 CW channels] tile per block in shared memory (lane == channel: 32 consecutive bytes
// per warp, conflict free); when the batch is done they meet at a consumers-only barrier and move the tile out together as 16-byte
// pieces, consecutive lanes on consecutive pieces of a tick row -- 32 CW contiguous bytes per tick and block (96 for banks of 3: one
// or two L2 requests per row; per-warp rows of 32 bytes, a request each, ran 17 % slower: the pattern is bound by L2 requests, not
// bytes).  n % 16 == 0, 16-byte aligned base.  One shift and one byte store per tick instead of 0.75 PRMT: ~7.4 instead of 6 instructions.
// PL 0: 16-byte stores (TILED [F/16][ch][16], or PLANAR rows as scattered stores).  PL 2: PLANAR duty rows
// [ch][F] leave through shared memory -- a consumer warp fills a box of 128 ticks x 32 channels (SWIZZLE_128B: lane r
// writes 16-tick chunk c at r*128 + ((c ^ (r & 7)) << 4), conflict free), one elected lane stores it through the 2-D
// tensor map over the duty rows; two boxes per warp alternate.
template <int K, int CW, int TLOG, int PL>
__global__ void __launch_bounds__(32 * (CW + 1)) k_pdm_v2_ws4(const PdmV2Params p, const PdmV2Work wk, const __grid_constant__ CUtensorMap tm_out) {
    constexpr int T = 1 << TLOG, NT = 32 * (CW + 1);
    constexpr int QW = T / 4;                                         // uint4 rows per dither slot
    extern __shared__ __align__(1024) uint8_t ws4_smem[];
    // 1024-byte alignment for the swizzled boxes, computed on the shared-window offset so that every pointer below stays an LDS / STS address
    uint8_t *base = ws4_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(ws4_smem) & 1023u)) & 1023u);
    uint8_t *pbox = base;                                             // [CW][2][4096]
    uint4 *dbuf = reinterpret_cast<uint4 *>(base + (PL == 2 ? CW * 2 * 4096 : PL == 1 ? CW * 4096 : 0));   // [2][QW][32]
    uint32_t *tab = reinterpret_cast<uint32_t *>(dbuf + 2 * QW * 32);
    for (uint32_t i = threadIdx.x; i < 1024; i += NT) tab[i] = __ldg(wk.jump + i);
    // Producer placement.  A warp runs on scheduler (%warpid % 4), and the hardware staggers the warp slots of
    // successive blocks on one SM (tools/probe_place.cu: warp 0 of the 1st..5th block sits in slot 0, 5, 10, 15, 16),
    // so a fixed producer warp index piles the producers of an SM onto one or two schedulers.  Instead the producer is
    // the warp that sits on scheduler (rank % 4), rank = arrival order of this block on its SM (never-reset per-SM
    // counter).  Any choice is correct; this one spreads the load.
    __shared__ uint32_t rank_s, smsp_s[CW + 1], item_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rank_s = atomicAdd(wk.sm_rank + smid, 1u);
    }
    if (lane == 0) {
        uint32_t wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        smsp_s[warp] = wid & 3u;
    }
    __syncthreads();
    uint32_t prod_warp = 0;
#pragma unroll
    for (int w = CW; w >= 0; --w) if (smsp_s[w] == (rank_s & 3u)) prod_warp = w;
    const uint32_t cw = warp - (warp > prod_warp ? 1u : 0u);          // consumer index 0..CW-1
    const uint32_t cl = cw * 32 + lane;                               // channel within the group
    const uint32_t L = p.ctl_div_log, period = 1u << (L - TLOG);      // batches per control period (L >= TLOG)
    const uint32_t m1 = p.m1, bs = p.bank_size;
    const uint64_t batches_total = p.F >> TLOG;
    const uint64_t n_ch = p.n_banks * bs;                             // channels inside the padded SoA rows
    const uint32_t total = wk.groups * wk.slices;
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
        const uint64_t c_lo = (uint64_t)g * (32 * CW);
        const uint64_t bank_lo = c_lo / bs;
        if (warp == prod_warp) {
            const uint64_t bank = bank_lo + lane;
            const bool have = bank < p.n_banks;
            uint32_t x = 1u;
            if (have) x = sl ? __ldcg(wk.prng_g + (uint64_t)g * 32 + lane) : __ldg(p.prng + bank);
            x = ws4_producer_arith<NT, TLOG>(dbuf, reinterpret_cast<const uint32_t (*)[256]>(tab), x, p.dmask, nb, lane);
            if (have) {
                if (sl + 1 < wk.slices) __stcg(wk.prng_g + (uint64_t)g * 32 + lane, x);
                else if (bank * bs >= c_lo && bank * bs < c_lo + 32 * CW) p.prng_out[bank] = x;    // the group that holds the bank's first channel owns it
            }
        } else {
            const uint64_t c = c_lo + cl;
            const bool live = c < n_ch;
            const uint32_t bl = live ? (uint32_t)(c / bs - bank_lo) : 0u;   // the channel's bank within the group
            V2Regs<K, 1> r;
            if (live) r.load(p.st, p.npad, c); else r.zero();
            // control divider at the first batch of the slice (count0 % T == 0)
            const uint32_t cb = (uint32_t)(((p.count0 >> TLOG) + bt0) & (period - 1));
            uint32_t until = cb == 0 ? 0 : period - cb;               // batches until the next boundary
            const uint32_t *sp_row = p.setpoints ? p.setpoints + v2_rows_before(p.count0, 1u << L, bt0 << TLOG) * p.n : nullptr;
            const bool store = c < p.n;
            uint8_t *dst = p.layout == CPROC_CUDA_TILED ? p.out + (((bt0 << (TLOG - 4)) * p.n + c) << 4) : p.out + c * p.F + (bt0 << TLOG);
            const uint64_t dstep = p.layout == CPROC_CUDA_TILED ? p.n << 4 : 16;
            const uint32_t pbox_s = PL == 2 ? (uint32_t)__cvta_generic_to_shared(pbox + cw * 8192) : PL == 1 ? (uint32_t)__cvta_generic_to_shared(pbox) : 0u;   // PL 1: the block's tile
            const uint4 *dlane = dbuf + bl;
            uint32_t s = 0;
            for (uint64_t bt = 0; bt < nb; ++bt) {
                if (until == 0) {                                     // uniform over the block
                    r.boundary(sp_row, c, p.n, L);
                    if (sp_row) sp_row += p.n;
                    until = period;
                }
                --until;
                bar_sync_slot<WS4_BAR_FULL, NT>(s);
                const uint4 *dslot = dlane + s * (QW * 32);
#pragma unroll
                for (int g16 = 0; g16 < T / 16; ++g16) {              // 16 ticks -> one 128-bit store
                    uint32_t w[4];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const uint4 dv = dslot[(g16 * 4 + i4) * 32];
                        const uint32_t d[4] = {dv.x, dv.y, dv.z, dv.w};
                        uint32_t a[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = ws4_tick<K>(r.p0[0], r.v0[0], r.s[0], d[i], m1);   // :108-116
                        if constexpr (PL == 1) {
                            const uint32_t row = (((uint32_t)bt << TLOG) + g16 * 16 + i4 * 4) & 127u;         // tick within the tile
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(pbox_s + (row + i) * (32u * CW) + cl), "r"(a[i] >> 24) : "memory");
                        } else w[i4] = pack_top_bytes(a[0], a[1], a[2], a[3]);
                    }
                    if constexpr (PL == 1) {
                        const uint32_t tk = ((uint32_t)bt << TLOG) + g16 * 16;          // ticks since the start of the slice (slices hold whole tiles)
                        if ((tk & 127u) == 112u) {                                      // 128 ticks x 32 CW channels staged
                            bar_sync_i<5, 32 * CW>();                                   // the consumers of the block (the next FULL barrier orders the tile's reuse)
                            const uint64_t tick0 = (bt0 << TLOG) + tk - 112;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t piece = i * (32u * CW) + cl, row = piece / (2u * CW), col = piece % (2u * CW);
                                if (c_lo + col * 16 < p.n) {                            // n % 16 == 0: whole pieces
                                    uint4 v;
                                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(pbox_s + piece * 16u));
                                    st_v4_stream(p.out + (tick0 + row) * p.n + c_lo + col * 16, v);
                                }
                            }
                        }
                    } else if constexpr (PL == 2) {
                        const uint32_t tk = ((uint32_t)bt << TLOG) + g16 * 16;          // ticks since the start of the slice
                        const uint32_t box = (tk >> 7) & 1u, ch = (tk >> 4) & 7u;
                        if (ch == 0 && tk >= 256) {                   // this box last left two boxes ago
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            __syncwarp();
                        }
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pbox_s + box * 4096u + lane * 128u + ((ch ^ (lane & 7u)) << 4)),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                        if (ch == 7) {                                // 128 ticks x 32 channels staged (slices hold whole boxes)
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                                             ::"l"(reinterpret_cast<uint64_t>(&tm_out)), "r"((int32_t)((bt0 << TLOG) + tk - 112)), "r"((int32_t)(c_lo + cw * 32)),
                                               "r"(pbox_s + box * 4096u) : "memory");
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    } else {
                        if (store) st_v4_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
                        dst += dstep;
                    }
                }
                if (bt + 2 < nb) bar_arrive_slot<WS4_BAR_EMPTY, NT>(s);
                s ^= 1u;
            }
            if constexpr (PL == 2) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
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

// ---------------------------------------------------------------------------
// host

static int ws4_jump(cproc_cuda_ctx *ctx, int tlog, const uint32_t **out) {       // M^(T/2) as byte-indexed tables
    uint32_t *&d = ctx->d_jump[tlog - 6];
    if (!d) {
        std::vector<uint32_t> h(1024);
        jump_table_fill(h.data(), 1u << (tlog - 1));
        CK(ctx, cudaMalloc(&d, h.size() * 4));
        CK(ctx, cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = d;
    return 0;
}

template <int K, int CW, int TLOG, int PL>
static int ws4_go(cproc_cuda_ctx *ctx, const PdmV2Params &p, PdmV2Work &wk, const CUtensorMap &tm, uint64_t slice_ticks) {
    auto kern = k_pdm_v2_ws4<K, CW, TLOG, PL>;
    constexpr size_t smem = ws4_smem_bytes(CW, TLOG, PL);
    static int occ = 0;                                               // per instantiation (one device kind: sm_100a)
    if (!occ) {
        CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * (CW + 1), smem));
        if (occ < 1) { occ = 0; return cproc_set_err(ctx, CPROC_CUDA_ECUDA, "pdm_v2: kernel does not fit an SM (%zu bytes of shared memory)", smem); }
    }
    const uint64_t batches = p.F >> TLOG;
    uint64_t bps = slice_ticks >> TLOG;
    if (PL != 0 && TLOG < 7) bps &= ~(uint64_t)((128 >> TLOG) - 1);   // slices hold whole 128-tick boxes / tiles
    if (bps < 1) bps = PL != 0 && TLOG < 7 ? (128 >> TLOG) : 1;
    wk.bps = (uint32_t)bps;
    wk.slices = (uint32_t)ceil_div_u64(batches, bps);
    if ((uint64_t)wk.groups * wk.slices >= 0xFFFFFFFFull) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: too many work items; render fewer ticks per call");
    int want_per_sm = ctx->pdm_ctas_per_sm;
    if (CW == 1 && want_per_sm < 6) want_per_sm = 6;                 // banks of 1: a block is two warps -- six blocks per SM (1.48e12 -> 1.60e12 at the C2 shape)
    const int per_sm = want_per_sm < occ ? want_per_sm : occ;
    uint64_t grid = (uint64_t)ctx->n_sm * per_sm;
    if (grid > wk.groups) grid = wk.groups;
    kern<<<(unsigned)grid, 32 * (CW + 1), smem, ctx->stream>>>(p, wk, tm);
    return 0;
}

template <int K, int CW>
static int ws4_dispatch(cproc_cuda_ctx *ctx, const PdmV2Params &p, PdmV2Work &wk, const CUtensorMap &tm, int tlog, int pl, uint64_t slice_ticks) {
    if (pl == 1) return ws4_go<K, CW, 7, 1>(ctx, p, wk, tm, slice_ticks);   // INTERLEAVED tiles (the caller has checked that 128-tick batches apply)
    if (pl == 2) return ws4_go<K, CW, 6, 2>(ctx, p, wk, tm, slice_ticks);   // boxes + 64-tick slots: four blocks per SM still fit
    if (tlog >= 7) return ws4_go<K, CW, 7, 0>(ctx, p, wk, tm, slice_ticks);
    return ws4_go<K, CW, 6, 0>(ctx, p, wk, tm, slice_ticks);
}

template <int K>
static int launch_v2_ws4(cproc_cuda_batch *b, PdmV2Params &p) {
    cproc_cuda_ctx *ctx = b->ctx;
    const int CW = p.bank_size <= 4 ? (int)p.bank_size : 4;
    // PLANAR duty rows through tensor-TMA boxes: 16-byte aligned rows of whole 128-tick boxes
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    int pl = 0;
    if (p.layout == CPROC_CUDA_PLANAR && ctx->pdm_planar_bulk && (p.F % 128) == 0 && pbulk::encode_rows_u8(&tm, p.out, p.F, p.n)) pl = 2;
    if (p.layout == CPROC_CUDA_INTERLEAVED) pl = 1;                   // launch_pdm_v2 has checked v2_interleaved_fast()
    // batch length (ticks per FULL / EMPTY hand-off): 128 when F, the counter and the control period allow, else 64
    int tlog = pl == 2 ? 6 : pl == 1 ? 7 : ctx->pdm_tlog;
    if (tlog > 6 && (((p.F | p.count0) & 127u) || p.ctl_div_log < 7)) tlog = 6;
    PdmV2Work wk;
    memset(&wk, 0, sizeof(wk));
    wk.groups = (uint32_t)ceil_div_u64(p.n_banks * p.bank_size, 32 * CW);
    if (b->n_flags < wk.groups) {
        if (b->d_flags) cudaFree(b->d_flags);
        if (b->d_prng_g) cudaFree(b->d_prng_g);
        b->d_flags = nullptr; b->d_prng_g = nullptr; b->n_flags = 0;
        CK(ctx, cudaMalloc(&b->d_flags, sizeof(unsigned long long) * wk.groups));
        CK(ctx, cudaMemsetAsync(b->d_flags, 0, sizeof(unsigned long long) * wk.groups, ctx->stream));
        CK(ctx, cudaMalloc(&b->d_prng_g, sizeof(uint32_t) * 32 * wk.groups));
        b->n_flags = wk.groups;
    }
    if (!ctx->d_work) {
        CK(ctx, cudaMalloc(&ctx->d_work, 2 * sizeof(uint32_t)));
        CK(ctx, cudaMemsetAsync(ctx->d_work, 0, 2 * sizeof(uint32_t), ctx->stream));
    }
    if (!ctx->d_sm_rank) {
        CK(ctx, cudaMalloc(&ctx->d_sm_rank, 1024 * sizeof(uint32_t)));
        CK(ctx, cudaMemsetAsync(ctx->d_sm_rank, 0, 1024 * sizeof(uint32_t), ctx->stream));
    }
    wk.counter = ctx->d_work; wk.flags = b->d_flags; wk.epoch = ++b->epoch; wk.prng_g = b->d_prng_g; wk.sm_rank = ctx->d_sm_rank;
    int rc;
    if ((rc = ws4_jump(ctx, tlog, &wk.jump))) return rc;
    const uint64_t slice_ticks = (uint64_t)ctx->pdm_slice_batches * 64;
    switch (CW) {
    case 1: return ws4_dispatch<K, 1>(ctx, p, wk, tm, tlog, pl, slice_ticks);
    case 2: return ws4_dispatch<K, 2>(ctx, p, wk, tm, tlog, pl, slice_ticks);
    case 3: return ws4_dispatch<K, 3>(ctx, p, wk, tm, tlog, pl, slice_ticks);
    default: return ws4_dispatch<K, 4>(ctx, p, wk, tm, tlog, pl, slice_ticks);
    }
}

template <int K, bool FASTQ>
static int launch_v2_order(cproc_cuda_batch *b, PdmV2Params &p, bool fast, bool dext) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!fast) {
        k_pdm_v2_any<K><<<(unsigned)ceil_div_u64(p.npad, 128), 128, 0, ctx->stream>>>(p);
        return 0;
    }
    if constexpr (FASTQ) {
        if (!dext && ctx->pdm_ws && (p.F % 64) == 0 && (p.count0 % 64) == 0 && p.ctl_div_log >= 6) return launch_v2_ws4<K>(b, p);
    }
    const int blk = ctx->pdm_block;
    const bool tpb = ctx->pdm_tpb && p.bank_size <= 4;
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
        // rows consumed = control boundaries met in [count, count+F)
        const uint64_t first = b->count == 0 ? 0 : div - b->count;
        const uint64_t rows = F > first ? 1 + (F - first - 1) / div : 0;
        if (rows > io->n_ctl) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: run crosses %llu control boundaries but ctl has %u rows", (unsigned long long)rows, io->n_ctl);
    }
    if (!b->d_prng2) CK(ctx, cudaMalloc(&b->d_prng2, sizeof(uint32_t) * b->n_banks));
    PdmV2Params p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n; p.n_banks = b->n_banks; p.bank_size = c.bank_size;
    p.prng = b->d_prng; p.prng_out = b->d_prng2; p.dither_ext = (const uint32_t *)io->in2; p.setpoints = (const uint32_t *)io->ctl;
    p.out = (uint8_t *)io->out; p.F = F; p.count0 = b->count; p.ctl_div_log = c.ctl_div_log; p.sh = c.out_shift;
    p.dmask = c.dither_mask; p.layout = io->layout; p.m1 = 0xFFFFFFFFu;
    const bool aligned_ptr = ((uintptr_t)io->out & 15) == 0 && (!io->in2 || ((uintptr_t)io->in2 & 15) == 0);
    const bool fastq = c.out_shift == 24 && (c.dither_mask & 0xFF000000u) == 0;
    const bool dext = io->in2 != nullptr;
    // INTERLEAVED [tick][ch]: only through k_pdm_v2_ws4's tile path (whole 128-tick batches, channels in sixteens); the conformance kernel otherwise
    const bool il_fast = io->layout == CPROC_CUDA_INTERLEAVED && fastq && !dext && ctx->pdm_ws && ((F | b->count) & 127) == 0 && c.ctl_div_log >= 7 && (b->n & 15) == 0;
    const bool fast = (F & 15) == 0 && (b->count & 15) == 0 && c.ctl_div_log >= 4 && aligned_ptr &&
                      (io->layout == CPROC_CUDA_TILED || io->layout == CPROC_CUDA_PLANAR || il_fast);
    int rc;
#define V2_ORDER(KK) (fastq ? launch_v2_order<KK, true>(b, p, fast, dext) : launch_v2_order<KK, false>(b, p, fast, dext))
    switch (c.order) {
    case 1: rc = V2_ORDER(1); break;
    case 2: rc = V2_ORDER(2); break;
    case 3: rc = V2_ORDER(3); break;
    default: rc = V2_ORDER(4); break;
    }
#undef V2_ORDER
    if (rc) return rc;
    CK_LAUNCH(ctx, "k_pdm_v2");
    if (!dext) { uint32_t *t = b->d_prng; b->d_prng = b->d_prng2; b->d_prng2 = t; }   // the generator state now lives in the other buffer
    b->count = (uint32_t)((b->count + F) & (div - 1));
    return 0;
}
