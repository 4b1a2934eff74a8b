// k_voice.cu -- the 32-bit phasor voice bank of linux/synth.c:169-202.
//
// Reference, per frame: sum = SUM_{v: inc_v != 0} ((int)state_v >> 4);
// state_v += inc_v; out = (float)sum * 2^-32   (sum_tick_saw, :169-181), or the
// OR of the phasor MSBs (sum_tick_square, :182-195).
//
// The phasor is a pure accumulator, so voice v at frame t is the closed form
// state_v + t*inc_v (mod 2^32).  That lets us put FRAMES on the lanes and
// walk the voices in the inner loop: every thread owns a private accumulator
// for its frames, voice parameters are staged once per block in shared memory
// and read as warp-wide broadcasts, and the mix needs no cross-lane traffic at
// all.  Integer wrap-around adds are associative, so any split of the voice
// axis (tiles, blocks, GPUs) reproduces the reference sum bit for bit.
#include "common.cuh"

struct VoiceParams {
    uint32_t *st;            // SoA [2][npad]: note_inc, note_state
    uint64_t npad, n;
    uint64_t G;              // voices per bus
    uint64_t n_bus;
    uint32_t tiles_per_bus;  // voice tiles per bus
    uint32_t tile;           // voices per tile
    uint64_t F;
    uint32_t mode;
    uint32_t advance;        // 1: the block also writes note_state += F * note_inc for its tile (one block per tile)
    int32_t *isum;           // [n_bus][F]
    float *vec;              // [n_bus][F] or null
};

template <int FPT, bool SQUARE>
__global__ void __launch_bounds__(256) k_voice_mix(const VoiceParams p) {
    extern __shared__ uint2 sv[];                 // (inc, state0) per voice of the tile
    const uint64_t bus = blockIdx.x / p.tiles_per_bus;
    const uint32_t tile = blockIdx.x % p.tiles_per_bus;
    const uint64_t v_lo = bus * p.G + (uint64_t)tile * p.tile;
    uint64_t v_hi = v_lo + p.tile;
    const uint64_t bus_end = (bus + 1) * p.G < p.n ? (bus + 1) * p.G : p.n;
    if (v_hi > bus_end) v_hi = bus_end;
    const uint32_t nv = v_hi > v_lo ? (uint32_t)(v_hi - v_lo) : 0u;
    for (uint32_t k = threadIdx.x; k < nv; k += blockDim.x) {
        uint32_t inc = p.st[v_lo + k], s0 = p.st[p.npad + v_lo + k];
        // voices that are off contribute nothing and do not advance (synth.c:173)
        sv[k] = inc ? make_uint2(inc, s0) : make_uint2(0u, 0u);
        if (p.advance) p.st[p.npad + v_lo + k] = s0 + (uint32_t)p.F * inc;      // synth.c:177 applied F times
    }
    __syncthreads();
    const uint64_t t0 = (uint64_t)blockIdx.y * (blockDim.x * FPT) + threadIdx.x;
    uint32_t tt[FPT], acc[FPT];
#pragma unroll
    for (int q = 0; q < FPT; ++q) { tt[q] = (uint32_t)(t0 + (uint64_t)q * blockDim.x); acc[q] = 0; }
#pragma unroll 8
    for (uint32_t k = 0; k < nv; ++k) {
        const uint2 v = sv[k];
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const uint32_t ph = v.y + tt[q] * v.x;           // state after tt ticks
            if (SQUARE) acc[q] |= ph & 0x80000000u;          // :188-190
            else acc[q] += (uint32_t)((int32_t)ph >> 4);     // :175-176
        }
    }
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        const uint64_t t = t0 + (uint64_t)q * blockDim.x;
        if (t >= p.F) continue;
        int32_t *dst = p.isum + bus * p.F + t;
        if (p.tiles_per_bus == 1) {
            *dst = (int32_t)acc[q];
            if (p.vec) p.vec[bus * p.F + t] = SQUARE ? __uint2float_rn(acc[q]) * 0x1p-32f
                                                    : __int2float_rn((int32_t)acc[q]) * 0x1p-32f;
        } else if (SQUARE) atomicOr((unsigned int *)dst, acc[q]);
        else atomicAdd((unsigned int *)dst, acc[q]);
    }
}

// isum -> float (synth.c:180 / :194) when the bus was split over several tiles
__global__ void k_voice_finish(const int32_t *isum, float *vec, uint64_t count, uint32_t mode) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    vec[i] = mode == CPROC_CUDA_MIX_SQUARE ? __uint2float_rn((uint32_t)isum[i]) * 0x1p-32f
                                           : __int2float_rn(isum[i]) * 0x1p-32f;
}

// note_state += F * note_inc for voices that are on (synth.c:177 applied F times)
__global__ void k_voice_advance(uint32_t *st, uint64_t npad, uint64_t n, uint32_t F) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st[npad + i] += F * st[i];
}

int launch_voice_bank(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out && !io->mix) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "voice_bank: out and mix are both NULL");
    if (F == 0) return 0;
    if (F > 0xFFFFFFFFull) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "voice_bank: F too large");
    VoiceParams p;
    p.st = b->d_state; p.npad = b->npad; p.n = b->n;
    p.G = b->cfg.voices_per_bus ? b->cfg.voices_per_bus : b->n;
    p.n_bus = b->n_bus; p.F = F; p.mode = b->cfg.mode;
    int blk, fpt;
    if (F <= 64) { blk = 64; fpt = 1; } else if (F <= 128) { blk = 128; fpt = 1; }
    else if (F <= 256) { blk = 256; fpt = 1; } else if (F <= 512) { blk = 256; fpt = 2; } else { blk = 256; fpt = 4; }
    const uint64_t gy = ceil_div_u64(F, (uint64_t)blk * fpt);
    // voices per tile: 4096 when there is enough work, smaller (down to 256) when that would leave
    // fewer than ~4 blocks per SM (a shard of a multi-GPU run: 512 Ki voices are only 128 tiles of 4096)
    {
        const uint64_t want_blocks = (uint64_t)ctx->n_sm * 4;
        const uint64_t tiles_needed = ceil_div_u64(want_blocks, p.n_bus * gy);
        uint64_t tile = ceil_div_u64(ceil_div_u64(p.G, tiles_needed), 256) * 256;
        if (tile > 4096) tile = 4096;
        if (tile > p.G) tile = p.G;
        p.tile = (uint32_t)tile;
    }
    p.tiles_per_bus = (uint32_t)ceil_div_u64(p.G, p.tile);
    p.advance = gy == 1;
    p.vec = (float *)io->out;
    // the integer mix always exists on the device: it is what gets reduced
    int32_t *isum = (int32_t *)io->mix;
    if (!isum) {
        size_t need = sizeof(int32_t) * p.n_bus * F;
        if (b->cap_mix < need) {
            if (b->d_mix) cudaFree(b->d_mix);
            b->d_mix = nullptr; b->cap_mix = 0;
            CK(ctx, cudaMalloc(&b->d_mix, need));
            b->cap_mix = need;
        }
        isum = (int32_t *)b->d_mix;
    }
    p.isum = isum;
    if (p.tiles_per_bus > 1) CK(ctx, cudaMemsetAsync(isum, 0, sizeof(int32_t) * p.n_bus * F, ctx->stream));
    const bool sq = p.mode == CPROC_CUDA_MIX_SQUARE;
    dim3 grid((unsigned)(p.n_bus * p.tiles_per_bus), (unsigned)gy);
    size_t smem = sizeof(uint2) * p.tile;
#define VOICE_LAUNCH(FPT) do { \
        if (sq) k_voice_mix<FPT, true><<<grid, blk, smem, ctx->stream>>>(p); \
        else k_voice_mix<FPT, false><<<grid, blk, smem, ctx->stream>>>(p); } while (0)
    if (fpt == 1) VOICE_LAUNCH(1); else if (fpt == 2) VOICE_LAUNCH(2); else VOICE_LAUNCH(4);
#undef VOICE_LAUNCH
    CK_LAUNCH(ctx, "k_voice_mix");
    if (p.tiles_per_bus > 1 && p.vec) {
        uint64_t cnt = p.n_bus * F;
        k_voice_finish<<<(unsigned)ceil_div_u64(cnt, 256), 256, 0, ctx->stream>>>(isum, p.vec, cnt, p.mode);
        CK_LAUNCH(ctx, "k_voice_finish");
    }
    if (!p.advance) {
        k_voice_advance<<<(unsigned)ceil_div_u64(b->n, 256), 256, 0, ctx->stream>>>(b->d_state, b->npad, b->n, (uint32_t)F);
        CK_LAUNCH(ctx, "k_voice_advance");
    }
    return 0;
}
