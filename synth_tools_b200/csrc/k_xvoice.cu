// k_xvoice.cu -- extension processors (NOT in the reference; defined by
// oracle/cproc_oracle.c: orc_xvoice_tick, orc_onepole_run).
//
//   voice = phasor_f (acc, cproc.h:140-142, read-then-advance like
//           synth.c:175-177) -> Chamberlin SVF low-pass -> linear AR envelope
//           -> pan.
//
// Every float operation is a single IEEE rounding in the same order as the
// oracle (fmaf == FFMA, explicit __fadd_rn/__fmul_rn so nvcc cannot contract
// differently), so the raw per-voice output is bit-exact against the C oracle.
// The mix is a float sum over voices and therefore order dependent: it is
// reduced in a fixed order (lane tree -> warp rows -> block partials -> final
// pass), deterministic run to run, and compared with tolerance.
#include "common.cuh"

struct XVoiceParams {
    uint32_t *st;            // SoA [5][npad]: phase, lp, bp, env, t
    const uint32_t *prm;     // SoA [8][npad]: inc, f, q, attack, release, gate_frames, gl, gr
    uint64_t npad, n, F;
    float *raw;              // PLANAR [inst][F][2] / TILED [F/2][inst][4] or null
    float *partial;          // [n_blocks][2][F] or null
    uint32_t layout;
};

struct XV {
    uint32_t phase, t, inc, gate;
    float lp, bp, env, f, q, att, rel, gl, gr;
};

// x = (float)(int)phase * 2^-31 is exact (a power-of-two scale of a 24-bit float, no
// underflow), so hp = x - lp (one rounding) is computed as fma(xi, 2^-31, -lp): the
// same bits, one instruction less.
__device__ __forceinline__ void xvoice_svf(XV &v, float &lp_out) {
    const float xi = __int2float_rn((int32_t)v.phase);
    v.phase += v.inc;
    const float lp = __fmaf_rn(v.f, v.bp, v.lp);
    float hp = __fmaf_rn(xi, 0x1p-31f, -lp);
    hp = __fmaf_rn(-v.q, v.bp, hp);
    v.bp = __fmaf_rn(v.f, hp, v.bp);
    v.lp = lp;
    lp_out = lp;
}

// Tick with the attack/release decision supplied by the caller (mix kernel: bit k of a
// per-chunk mask, so the frame counter is advanced once per chunk).  Same float
// operations as xvoice_tick except that a -0.0 envelope in release becomes +0.0
// (fmaxf); a -0.0 envelope can only come from an uploaded state.
__device__ __forceinline__ float xvoice_tick_flag(XV &v, bool attack) {
    float lp;
    xvoice_svf(v, lp);
    float e = v.env;
    if (attack) { e = __fadd_rn(e, v.att); if (e > 1.0f) e = 1.0f; }
    else { e = fmaxf(__fsub_rn(e, v.rel), 0.0f); }
    v.env = e;
    return __fmul_rn(lp, e);
}

__device__ __forceinline__ float xvoice_tick(XV &v) {
    const float xi = __int2float_rn((int32_t)v.phase);
    v.phase += v.inc;
    const float lp = __fmaf_rn(v.f, v.bp, v.lp);
    float hp = __fmaf_rn(xi, 0x1p-31f, -lp);
    hp = __fmaf_rn(-v.q, v.bp, hp);
    v.bp = __fmaf_rn(v.f, hp, v.bp);
    v.lp = lp;
    float e = v.env;
    if (v.t < v.gate) { e = __fadd_rn(e, v.att); if (e > 1.0f) e = 1.0f; }
    else { e = __fsub_rn(e, v.rel); if (e < 0.0f) e = 0.0f; }
    v.env = e;
    v.t += 1;
    return __fmul_rn(lp, e);
}

#define XV_BLOCK 128
#define XV_WARPS (XV_BLOCK / 32)
#define XV_CHUNK 32

template <bool RAW, bool MIX>
__global__ void __launch_bounds__(XV_BLOCK) k_xvoice(const XVoiceParams p) {
    __shared__ float red[XV_WARPS][2][XV_CHUNK][33];   // [warp][ch][frame][lane]
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < p.n;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    XV v = {};
    if (mine) {
        const uint32_t *s = p.st + i; const uint32_t *r = p.prm + i;
        v.phase = s[0]; v.lp = __uint_as_float(s[p.npad]); v.bp = __uint_as_float(s[2 * p.npad]);
        v.env = __uint_as_float(s[3 * p.npad]); v.t = s[4 * p.npad];
        v.inc = r[0]; v.f = __uint_as_float(r[p.npad]); v.q = __uint_as_float(r[2 * p.npad]);
        v.att = __uint_as_float(r[3 * p.npad]); v.rel = __uint_as_float(r[4 * p.npad]); v.gate = r[5 * p.npad];
        v.gl = __uint_as_float(r[6 * p.npad]); v.gr = __uint_as_float(r[7 * p.npad]);
    }
    for (uint64_t t0 = 0; t0 < p.F; t0 += XV_CHUNK) {
        const uint32_t cols = p.F - t0 < XV_CHUNK ? (uint32_t)(p.F - t0) : XV_CHUNK;
        float pl = 0.f, pr = 0.f;
        for (uint32_t k = 0; k < cols; ++k) {
            const float y = xvoice_tick(v);
            const float l = __fmul_rn(v.gl, y), r = __fmul_rn(v.gr, y);
            if (MIX) { red[warp][0][k][lane] = mine ? l : 0.f; red[warp][1][k][lane] = mine ? r : 0.f; }
            if (RAW && mine) {
                const uint64_t t = t0 + k;
                if (p.layout == CPROC_CUDA_TILED) {
                    if (k & 1) st_v4_stream(p.raw + (((t >> 1) * p.n + i) << 2),
                                            make_uint4(__float_as_uint(pl), __float_as_uint(pr), __float_as_uint(l), __float_as_uint(r)));
                    else { pl = l; pr = r; }
                } else {
                    *reinterpret_cast<float2 *>(p.raw + ((i * p.F + t) << 1)) = make_float2(l, r);
                }
            }
        }
        if (MIX) {
            __syncthreads();
            // lane f of warp w sums frame f over the 32 voices of warp w, in lane order
            float sl = 0.f, sr = 0.f;
            if (lane < cols) {
#pragma unroll 8
                for (int j = 0; j < 32; ++j) { sl = __fadd_rn(sl, red[warp][0][lane][j]); sr = __fadd_rn(sr, red[warp][1][lane][j]); }
            }
            __syncthreads();
            if (lane < cols) { red[warp][0][lane][0] = sl; red[warp][1][lane][0] = sr; }
            __syncthreads();
            if (threadIdx.x < 2 * cols) {
                const uint32_t ch = threadIdx.x / cols, f = threadIdx.x % cols;
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < XV_WARPS; ++w) s = __fadd_rn(s, red[w][ch][f][0]);
                p.partial[((uint64_t)blockIdx.x * 2 + ch) * p.F + t0 + f] = s;
            }
            __syncthreads();
        }
    }
    if (mine) {
        uint32_t *s = p.st + i;
        s[0] = v.phase; s[p.npad] = __float_as_uint(v.lp); s[2 * p.npad] = __float_as_uint(v.bp);
        s[3 * p.npad] = __float_as_uint(v.env); s[4 * p.npad] = v.t;
    }
}

// Mix-only render (C4: millions of voices -> one stereo bus).  The SVF is a
// recurrence in time, so a voice stays on one thread; the reduction over voices is
// taken out of the inner loop instead of being paid per voice-sample: a thread keeps
// 2 x 32 bus accumulators (32 frames, left/right) in registers and walks ITS voices
// (tid, tid + T, ...) through the same 32 frames one after the other, so the pan
// multiply and the mix add fuse into one FFMA per channel and nothing crosses lanes
// until all of the thread's voices are done.  Only then: one block reduction per 32
// frames (smem columns, fixed order) and one partial row per block; k_xvoice_final
// adds the block rows in a fixed order, so the result is deterministic run to run.
// Voice state travels through HBM once per 32 frames (13 words in, 5 out per voice:
// 2.25 B per voice-sample); the tick itself is the bit-exact xvoice_tick, so the
// downloaded state equals the oracle's.
#define XM_BLOCK 128
#define XM_CHUNK 32
__device__ __forceinline__ void xv_load(XV &v, const XVoiceParams &p, uint64_t i) {
    const uint32_t *s = p.st + i; const uint32_t *r = p.prm + i;
    v.phase = __ldcg(s); v.lp = __uint_as_float(__ldcg(s + p.npad)); v.bp = __uint_as_float(__ldcg(s + 2 * p.npad));
    v.env = __uint_as_float(__ldcg(s + 3 * p.npad)); v.t = __ldcg(s + 4 * p.npad);
    v.inc = __ldg(r); v.f = __uint_as_float(__ldg(r + p.npad)); v.q = __uint_as_float(__ldg(r + 2 * p.npad));
    v.att = __uint_as_float(__ldg(r + 3 * p.npad)); v.rel = __uint_as_float(__ldg(r + 4 * p.npad)); v.gate = __ldg(r + 5 * p.npad);
    v.gl = __uint_as_float(__ldg(r + 6 * p.npad)); v.gr = __uint_as_float(__ldg(r + 7 * p.npad));
}
__global__ void __launch_bounds__(XM_BLOCK, 4) k_xvoice_mix(const XVoiceParams p) {
    __shared__ float red[2 * XM_CHUNK][XM_BLOCK + 1];
    const uint64_t T = (uint64_t)gridDim.x * XM_BLOCK;
    const uint64_t tid = (uint64_t)blockIdx.x * XM_BLOCK + threadIdx.x;
    for (uint64_t t0 = 0; t0 < p.F; t0 += XM_CHUNK) {
        const uint32_t cols = p.F - t0 < XM_CHUNK ? (uint32_t)(p.F - t0) : XM_CHUNK;
        float aL[XM_CHUNK], aR[XM_CHUNK];
#pragma unroll
        for (int k = 0; k < XM_CHUNK; ++k) { aL[k] = 0.f; aR[k] = 0.f; }
        // software pipeline: the 13 words of the thread's next voice are in flight
        // while the current one renders its 32 frames
        XV nx = {};
        if (tid < p.n) xv_load(nx, p, tid);
        for (uint64_t i = tid; i < p.n; i += T) {
            XV v = nx;
            if (i + T < p.n) xv_load(nx, p, i + T);
            // bit k: tick k of this chunk is in the attack phase, (t + k) mod 2^32 < gate
            uint32_t amask;
            if (v.t <= 0xFFFFFFFFu - XM_CHUNK) {
                const uint32_t rem = v.t < v.gate ? v.gate - v.t : 0u;
                amask = rem >= 32u ? 0xFFFFFFFFu : (1u << rem) - 1u;
            } else {                                       // the frame counter wraps inside the chunk
                amask = 0;
                for (uint32_t k = 0; k < XM_CHUNK; ++k) amask |= (uint32_t)(v.t + k < v.gate) << k;
            }
            if (cols == XM_CHUNK) {
#pragma unroll
                for (int k = 0; k < XM_CHUNK; ++k) {
                    const float y = xvoice_tick_flag(v, (amask >> k) & 1u);
                    aL[k] = __fmaf_rn(v.gl, y, aL[k]);
                    aR[k] = __fmaf_rn(v.gr, y, aR[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < XM_CHUNK; ++k) {
                    if (k < (int)cols) {
                        const float y = xvoice_tick_flag(v, (amask >> k) & 1u);
                        aL[k] = __fmaf_rn(v.gl, y, aL[k]);
                        aR[k] = __fmaf_rn(v.gr, y, aR[k]);
                    }
                }
            }
            v.t += cols;
            uint32_t *w = p.st + i;
            w[0] = v.phase; w[p.npad] = __float_as_uint(v.lp); w[2 * p.npad] = __float_as_uint(v.bp);
            w[3 * p.npad] = __float_as_uint(v.env); w[4 * p.npad] = v.t;
        }
#pragma unroll
        for (int k = 0; k < XM_CHUNK; ++k) { red[k][threadIdx.x] = aL[k]; red[XM_CHUNK + k][threadIdx.x] = aR[k]; }
        __syncthreads();
        if (threadIdx.x < 2 * XM_CHUNK) {
            const uint32_t ch = threadIdx.x / XM_CHUNK, f = threadIdx.x % XM_CHUNK;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
            for (int j = 0; j < XM_BLOCK; j += 4) {
                s0 = __fadd_rn(s0, red[threadIdx.x][j]); s1 = __fadd_rn(s1, red[threadIdx.x][j + 1]);
                s2 = __fadd_rn(s2, red[threadIdx.x][j + 2]); s3 = __fadd_rn(s3, red[threadIdx.x][j + 3]);
            }
            if (f < cols) p.partial[((uint64_t)blockIdx.x * 2 + ch) * p.F + t0 + f] = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
        }
        __syncthreads();
    }
}

// mix[c][t] = SUM_b partial[b][c][t], fixed order, 4 independent chains
__global__ void k_xvoice_final(const float *partial, float *mix, uint64_t n_blocks, uint64_t cols) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cols) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    uint64_t b = 0;
    for (; b + 4 <= n_blocks; b += 4) {
        s0 = __fadd_rn(s0, partial[(b + 0) * cols + i]); s1 = __fadd_rn(s1, partial[(b + 1) * cols + i]);
        s2 = __fadd_rn(s2, partial[(b + 2) * cols + i]); s3 = __fadd_rn(s3, partial[(b + 3) * cols + i]);
    }
    for (; b < n_blocks; ++b) s0 = __fadd_rn(s0, partial[b * cols + i]);
    mix[i] = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

int launch_xvoice(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out && !io->mix) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "xvoice: out and mix are both NULL");
    if (io->out && io->layout == CPROC_CUDA_INTERLEAVED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "xvoice: INTERLEAVED layout not supported");
    if (io->out && io->layout == CPROC_CUDA_TILED && (F & 1)) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "xvoice: TILED needs even F");
    if (F == 0) return 0;
    const bool mix_only = io->mix && !io->out;
    const uint64_t n_blocks = mix_only ? (uint64_t)ctx->n_sm * 4 : ceil_div_u64(b->n, XV_BLOCK);
    XVoiceParams p;
    p.st = b->d_state; p.prm = b->d_param; p.npad = b->npad; p.n = b->n; p.F = F;
    p.raw = (float *)io->out; p.layout = io->layout; p.partial = nullptr;
    if (io->mix) {
        size_t need = sizeof(float) * n_blocks * 2 * F;
        if (b->cap_mix < need) {
            if (b->d_mix) cudaFree(b->d_mix);
            b->d_mix = nullptr; b->cap_mix = 0;
            CK(ctx, cudaMalloc(&b->d_mix, need));
            b->cap_mix = need;
        }
        p.partial = (float *)b->d_mix;
    }
    if (io->out && io->mix) k_xvoice<true, true><<<(unsigned)n_blocks, XV_BLOCK, 0, ctx->stream>>>(p);
    else if (io->out) k_xvoice<true, false><<<(unsigned)n_blocks, XV_BLOCK, 0, ctx->stream>>>(p);
    else k_xvoice_mix<<<(unsigned)n_blocks, XM_BLOCK, 0, ctx->stream>>>(p);
    CK_LAUNCH(ctx, "k_xvoice");
    if (io->mix) {
        k_xvoice_final<<<(unsigned)ceil_div_u64(2 * F, 128), 128, 0, ctx->stream>>>(p.partial, (float *)io->mix, n_blocks, 2 * F);
        CK_LAUNCH(ctx, "k_xvoice_final");
    }
    return 0;
}

// ---------------------------------------------------------------------------
// one-pole low-pass y = fma(a, x - y, y)
__global__ void k_onepole(float *y, const float *a, uint64_t n, uint64_t F, const float *in, float *out, uint32_t layout) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = y[i];
    const float c = a[i];
    for (uint64_t t = 0; t < F; ++t) {
        const uint64_t idx = layout == CPROC_CUDA_INTERLEAVED ? t * n + i : i * F + t;
        s = __fmaf_rn(c, __fsub_rn(in[idx], s), s);
        out[idx] = s;
    }
    y[i] = s;
}

int launch_onepole(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->in || !io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "onepole: in/out is NULL");
    if (io->layout == CPROC_CUDA_TILED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "onepole: TILED layout not supported");
    if (F == 0) return 0;
    k_onepole<<<(unsigned)ceil_div_u64(b->n, 128), 128, 0, ctx->stream>>>((float *)b->d_state, (const float *)b->d_param, b->n, F,
                                                                       (const float *)io->in, (float *)io->out, io->layout);
    CK_LAUNCH(ctx, "k_onepole");
    return 0;
}
