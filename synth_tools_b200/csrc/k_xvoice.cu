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

__device__ __forceinline__ float xvoice_tick(XV &v) {
    const float x = __fmul_rn(__int2float_rn((int32_t)v.phase), 0x1p-31f);
    v.phase += v.inc;
    const float lp = __fmaf_rn(v.f, v.bp, v.lp);
    float hp = __fsub_rn(x, lp);
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
    const uint64_t n_blocks = ceil_div_u64(b->n, XV_BLOCK);
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
    else k_xvoice<false, true><<<(unsigned)n_blocks, XV_BLOCK, 0, ctx->stream>>>(p);
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
