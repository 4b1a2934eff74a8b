// k_pdm.cu -- sigma-delta / PDM kernels (integer, serial in time, bit-exact).
//
//   k_pdm_v2      stm32f103/mod_pdm_pwm.c:101-141 (glide + pdmK_update) with the
//                 control-rate line generator of mod_controlrate.c:28-40
//   k_pdm_v1      stm32f103/mod_pdm.c:230-264 (carry-bit PDM, shared dither)
//   k_pdm_raw     stm32f103/pdm.h:13-77 (pdmK_update on an input stream)
//   k_pwm         stm32f103/mod_pdm.c:167-175
//
// Mapping: the recurrence is nonlinear (quantiser in the loop), so time stays
// serial and the batch axis carries the parallelism.  One thread owns one
// dither bank (B channels that share one random word per tick, exactly like
// one MCU of the reference) and keeps every state word in registers for the
// whole run; 16 ticks of 8-bit duty are packed into one 128-bit store.
#include "common.cuh"

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
};

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

// Insert byte `q` (0..255) at byte position j of w.
template <int J>
__device__ __forceinline__ uint32_t put_byte(uint32_t w, uint32_t q) {
    // selector nibbles pick from {w.b0..b3 = 0..3, q.b0..b3 = 4..7}
    constexpr uint32_t sel = J == 0 ? 0x3214u : J == 1 ? 0x3240u : J == 2 ? 0x3410u : 0x4210u;
    return __byte_perm(w, q, sel);
}

template <int K, int B, bool TPB>
__global__ void __launch_bounds__(128) k_pdm_v2(const PdmV2Params p) {
    // TPB: thread == bank of B channels.  !TPB: thread == channel (B == 1),
    // bank = channel / bank_size, every thread of a bank replays the bank PRNG.
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t c0 = tid * B;
    if (c0 >= p.npad) return;
    const uint64_t bank = TPB ? tid : tid / p.bank_size;
    const bool rng_owner = TPB ? true : (tid % p.bank_size == 0);

    uint32_t sp[B], p0[B], v0[B], p1[B], v1[B], s[B][K];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const uint32_t *x = p.st + c0 + j;
        sp[j] = x[0]; p0[j] = x[p.npad]; v0[j] = x[2 * p.npad]; p1[j] = x[3 * p.npad]; v1[j] = x[4 * p.npad];
#pragma unroll
        for (int k = 0; k < K; ++k) s[j][k] = x[(5 + k) * p.npad];
    }
    uint32_t rng = p.prng[bank];
    const uint32_t L = p.ctl_div_log, div_mask = (1u << L) - 1u, sh = p.sh, dmask = p.dmask;
    uint32_t cnt = p.count0;
    uint64_t row = 0;
    const uint64_t groups = p.F >> 4;
    const uint32_t *dext = p.dither_ext ? p.dither_ext + bank * p.F : nullptr;

    for (uint64_t g = 0; g < groups; ++g) {
        if (cnt == 0) {
            // mod_pdm_pwm.c:129-137 + mod_controlrate.c:28-40
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if (p.setpoints && c0 + j < p.n) sp[j] = p.setpoints[row * p.n + c0 + j];
                p0[j] = p1[j]; v0[j] = v1[j];
                p1[j] += v1[j] << L;
                v1[j] = (uint32_t)((int32_t)(sp[j] - p1[j]) >> L);
            }
            ++row;
        }
        uint32_t w[B][4];
        uint32_t dbuf[16];
        if (dext) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 v = *reinterpret_cast<const uint4 *>(dext + g * 16 + i * 4);
                dbuf[4 * i] = v.x; dbuf[4 * i + 1] = v.y; dbuf[4 * i + 2] = v.z; dbuf[4 * i + 3] = v.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t d;
            if (dext) d = dbuf[i] & dmask;
            else { rng = xorshift32_step(rng); d = rng & dmask; }   // mod_pdm_pwm.c:127
#pragma unroll
            for (int j = 0; j < B; ++j) {
                p0[j] += v0[j];                                      // :101-104
                uint32_t q = pdm_step<K>(s[j], p0[j], sh, d);        // :108-116
                uint32_t &ww = w[j][i >> 2];
                switch (i & 3) {
                case 0: ww = q; break;
                case 1: ww = put_byte<1>(ww, q); break;
                case 2: ww = put_byte<2>(ww, q); break;
                default: ww = put_byte<3>(ww, q); break;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (c0 + j < p.n) {
                uint4 v = make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
                uint8_t *dst = p.layout == CPROC_CUDA_TILED
                    ? p.out + ((g * p.n + c0 + j) << 4)
                    : p.out + (c0 + j) * p.F + (g << 4);
                st_v4_stream(dst, v);
            }
        }
        cnt = (cnt + 16) & div_mask;
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
        uint32_t *x = p.st + c0 + j;
        x[0] = sp[j]; x[p.npad] = p0[j]; x[2 * p.npad] = v0[j]; x[3 * p.npad] = p1[j]; x[4 * p.npad] = v1[j];
#pragma unroll
        for (int k = 0; k < K; ++k) x[(5 + k) * p.npad] = s[j][k];
    }
    if (rng_owner && !dext) p.prng[bank] = rng;
}

// Any F, any count alignment, byte stores: the conformance path.
template <int K>
__global__ void k_pdm_v2_any(const PdmV2Params p) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.npad) return;
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

template <int K>
static void launch_v2_order(cproc_cuda_batch *b, const PdmV2Params &p, bool fast, bool tpb) {
    cproc_cuda_ctx *ctx = b->ctx;
    int blk = ctx->pdm_block;
    if (!fast) {
        k_pdm_v2_any<K><<<(unsigned)ceil_div_u64(p.npad, 128), 128, 0, ctx->stream>>>(p);
        return;
    }
    if (tpb) {
        unsigned grid = (unsigned)ceil_div_u64(p.n_banks, blk);
        switch (p.bank_size) {
        case 1: k_pdm_v2<K, 1, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        case 2: k_pdm_v2<K, 2, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        case 3: k_pdm_v2<K, 3, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        default: k_pdm_v2<K, 4, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        }
    } else {
        k_pdm_v2<K, 1, false><<<(unsigned)ceil_div_u64(p.npad, blk), blk, 0, ctx->stream>>>(p);
    }
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
        uint64_t first = b->count == 0 ? 0 : div - b->count;
        uint64_t rows = F > first ? 1 + (F - first - 1) / div : 0;
        if (rows > io->n_ctl) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v2: run crosses %llu control boundaries but ctl has %u rows", (unsigned long long)rows, io->n_ctl);
    }
    PdmV2Params p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n; p.n_banks = b->n_banks; p.bank_size = c.bank_size;
    p.prng = b->d_prng; p.dither_ext = (const uint32_t *)io->in2; p.setpoints = (const uint32_t *)io->ctl;
    p.out = (uint8_t *)io->out; p.F = F; p.count0 = b->count; p.ctl_div_log = c.ctl_div_log; p.sh = c.out_shift;
    p.dmask = c.dither_mask; p.layout = io->layout;
    bool aligned_ptr = ((uintptr_t)io->out & 15) == 0 && (!io->in2 || ((uintptr_t)io->in2 & 15) == 0);
    bool fast = (F & 15) == 0 && (b->count & 15) == 0 && c.ctl_div_log >= 4 && aligned_ptr &&
                (io->layout == CPROC_CUDA_TILED || io->layout == CPROC_CUDA_PLANAR);
    bool tpb = ctx->pdm_tpb && c.bank_size <= 4;
    switch (c.order) {
    case 1: launch_v2_order<1>(b, p, fast, tpb); break;
    case 2: launch_v2_order<2>(b, p, fast, tpb); break;
    case 3: launch_v2_order<3>(b, p, fast, tpb); break;
    default: launch_v2_order<4>(b, p, fast, tpb); break;
    }
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
};

// accu += x with carry out shifted into the LSB of bits (ARM: adds + adc;
// mod_pdm.c:236-240 uses rrx, i.e. the MSB: the final __brev gives the same
// time order LSB-first).
__device__ __forceinline__ void add_carry_shift(uint32_t &accu, uint32_t &bits, uint32_t x) {
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(accu), "+r"(bits) : "r"(x));
}

template <int B, bool TPB>
__global__ void __launch_bounds__(128) k_pdm_v1(const PdmV1Params p) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t c0 = tid * B;
    if (c0 >= p.npad) return;
    const uint64_t bank = TPB ? tid : tid / p.bank_size;
    const bool rng_owner = TPB ? true : (tid % p.bank_size == 0);
    uint32_t sp[B], acc[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { sp[j] = p.st[c0 + j]; acc[j] = p.st[p.npad + c0 + j]; }
    uint32_t rng = p.prng[bank];
    const uint32_t dmask = p.dmask;
    const uint32_t *dext = p.dither_ext ? p.dither_ext + bank * p.F : nullptr;
    const uint64_t words = p.F >> 5;
    const bool tiled = p.layout == CPROC_CUDA_TILED;
    uint32_t w0[B], w1[B], w2[B];
    for (uint64_t g = 0; g < words; ++g) {
        uint32_t bits[B];
#pragma unroll
        for (int j = 0; j < B; ++j) bits[j] = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            uint32_t d;
            if (dext) d = dext[g * 32 + i] & dmask;
            else { rng = xorshift32_step(rng); d = rng & dmask; }   // mod_pdm.c:261
#pragma unroll
            for (int j = 0; j < B; ++j) add_carry_shift(acc[j], bits[j], sp[j] + d);  // :235-240
        }
        const uint32_t q = (uint32_t)g & 3u;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            uint32_t wv = __brev(bits[j]);
            if (tiled) {
                if (q == 0) w0[j] = wv; else if (q == 1) w1[j] = wv; else if (q == 2) w2[j] = wv;
                else if (c0 + j < p.n)
                    st_v4_stream(p.out + (((g >> 2) * p.n + c0 + j) << 2), make_uint4(w0[j], w1[j], w2[j], wv));
            } else if (c0 + j < p.n) {
                uint64_t idx = p.layout == CPROC_CUDA_INTERLEAVED ? g * p.n + c0 + j : (c0 + j) * words + g;
                p.out[idx] = wv;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < B; ++j) p.st[p.npad + c0 + j] = acc[j];
    if (rng_owner && !dext) p.prng[bank] = rng;
}

int launch_pdm_v1(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    const cproc_cuda_config &c = b->cfg;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: out is NULL");
    if (F & 31) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: F must be a multiple of 32 (packed bit output)");
    if (io->layout == CPROC_CUDA_TILED && (F & 127)) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pdm_v1: TILED needs F %% 128 == 0");
    if (F == 0) return 0;
    PdmV1Params p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n; p.n_banks = b->n_banks; p.bank_size = c.bank_size;
    p.prng = b->d_prng; p.dither_ext = (const uint32_t *)io->in2; p.out = (uint32_t *)io->out; p.F = F;
    p.dmask = c.dither_mask; p.layout = io->layout;
    int blk = ctx->pdm_block;
    if (ctx->pdm_tpb && c.bank_size <= 4) {
        unsigned grid = (unsigned)ceil_div_u64(p.n_banks, blk);
        switch (c.bank_size) {
        case 1: k_pdm_v1<1, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        case 2: k_pdm_v1<2, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        case 3: k_pdm_v1<3, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        default: k_pdm_v1<4, true><<<grid, blk, 0, ctx->stream>>>(p); break;
        }
    } else {
        k_pdm_v1<1, false><<<(unsigned)ceil_div_u64(p.npad, blk), blk, 0, ctx->stream>>>(p);
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

int launch_pwm(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pwm: out is NULL");
    if (io->layout == CPROC_CUDA_TILED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "pwm: TILED layout not supported");
    if (F == 0) return 0;
    k_pwm<<<(unsigned)ceil_div_u64(b->n, 128), 128, 0, ctx->stream>>>(b->d_state, b->d_param, b->n, F, (uint8_t *)io->out, io->layout);
    CK_LAUNCH(ctx, "k_pwm");
    return 0;
}
