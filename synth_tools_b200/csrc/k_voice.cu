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
//
// One launch per frame block, whatever the shape:
//  * a bus that is split over several voice tiles is summed with integer atomics into an
//    internal accumulator row that is all zero between launches: the last tile of a
//    (bus, frame range) to arrive -- an atomic ticket -- reads the row back with
//    atomicExch(.., 0) (which also cleans it for the next launch), and writes the integer
//    mix and the float bus.  No memset, no conversion kernel.
//  * tiles are sized so that the grid is a whole number of blocks per SM (a 512 Ki-voice
//    shard of an 8-GPU render is 1,184 tiles of 443 voices: 8 blocks on every SM, where
//    512 tiles of 1,024 left 91 SMs with 4 and 57 with 3).
//  * with a mix bus attached (cproc_cuda_bus_attach) that last tile pushes its piece of the
//    mix into every peer's bus buffer instead, and the exchange completes in this launch or
//    beside the next one (bus_fused.cuh).
#include "common.cuh"
#include "bus_fused.cuh"

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
    int32_t *isum;           // [n_bus][F] or null
    float *vec;              // [n_bus][F] or null
    uint32_t *acc;           // [n_bus][F] accumulators (tiles_per_bus > 1), zero between launches
    uint32_t *tickets;       // [n_bus][gridDim.y] tiles arrived, zero between launches
    uint32_t n_tile_blocks;  // gridDim.x without the finisher block of a pipelined bus
};

__device__ __forceinline__ void voice_emit(const VoiceParams &p, const BusFused &bf, bool square, uint64_t idx, uint32_t v) {
    if (bf.world) { bus_emit_word(bf, idx, v); return; }
    if (p.isum) p.isum[idx] = (int32_t)v;
    if (p.vec) p.vec[idx] = square ? __uint2float_rn(v) * 0x1p-32f : __int2float_rn((int32_t)v) * 0x1p-32f;   // synth.c:194 / :180
}

template <int FPT, bool SQUARE>
__global__ void __launch_bounds__(256) k_voice_mix(const VoiceParams p, const BusFused bf) {
    extern __shared__ __align__(16) uint2 sv[];   // (inc, state0) per voice of the tile
    __shared__ uint32_t last_s;
    if (blockIdx.x >= p.n_tile_blocks) {          // pipelined bus: the block that runs the previous frame block's exchange
        if (blockIdx.y == 0) bus_exchange_block(bf);
        return;
    }
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
    if (threadIdx.x == 0 && (nv & 1)) sv[nv] = make_uint2(0u, 0u);             // silent partner of an odd last voice: (0 >> 4) adds nothing
    __syncthreads();
    const uint64_t t0 = (uint64_t)blockIdx.y * (blockDim.x * FPT) + threadIdx.x;
    uint32_t tt[FPT], acc[FPT];
#pragma unroll
    for (int q = 0; q < FPT; ++q) { tt[q] = (uint32_t)(t0 + (uint64_t)q * blockDim.x); acc[q] = 0; }
    // two voices per 128-bit broadcast load (the tile is padded to an even count with a silent voice)
    const uint4 *sv2 = reinterpret_cast<const uint4 *>(sv);
#pragma unroll 4
    for (uint32_t k = 0; k < (nv + 1) / 2; ++k) {
        const uint4 v = sv2[k];
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const uint32_t ph0 = v.y + tt[q] * v.x, ph1 = v.w + tt[q] * v.z;     // state after tt ticks
            if (SQUARE) acc[q] |= (ph0 | ph1) & 0x80000000u;                     // :188-190
            else acc[q] += (uint32_t)((int32_t)ph0 >> 4) + (uint32_t)((int32_t)ph1 >> 4);   // :175-176
        }
    }
    if (p.tiles_per_bus == 1) {                   // the block holds the whole bus: its registers are the mix
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const uint64_t t = t0 + (uint64_t)q * blockDim.x;
            if (t < p.F) voice_emit(p, bf, SQUARE, bus * p.F + t, acc[q]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const uint64_t t = t0 + (uint64_t)q * blockDim.x;
            if (t >= p.F) continue;
            if (SQUARE) atomicOr(p.acc + bus * p.F + t, acc[q]); else atomicAdd(p.acc + bus * p.F + t, acc[q]);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last_s = atomicAdd(p.tickets + bus * gridDim.y + blockIdx.y, 1u) == p.tiles_per_bus - 1u;
        __syncthreads();
        if (!last_s) return;                      // (uniform over the block)
        __threadfence();
#pragma unroll
        for (int q = 0; q < FPT; ++q) {           // last tile of this (bus, frame range): read the row back and leave it zero
            const uint64_t t = t0 + (uint64_t)q * blockDim.x;
            if (t < p.F) voice_emit(p, bf, SQUARE, bus * p.F + t, atomicExch(p.acc + bus * p.F + t, 0u));
        }
        if (threadIdx.x == 0) p.tickets[bus * gridDim.y + blockIdx.y] = 0;
    }
    if (bf.world) bus_participant_done(bf);
}

// isum -> float (synth.c:180 / :194): cproc_cuda_mix_to_float, after a host-side reduce of the integer mix
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
    // frames per thread: more frames per thread amortise the broadcast load of a voice (0.5 / FPT of the 3 instructions a
    // voice-frame costs), fewer threads per block cut the bus into more tiles (one atomic per frame per tile)
    int blk, fpt;
    if (F <= 64) { blk = 64; fpt = 1; } else if (F <= 128) { blk = 64; fpt = 2; }
    else if (F <= 256) { blk = 128; fpt = 2; } else if (F <= 512) { blk = 128; fpt = 4; } else if (F <= 1024) { blk = 256; fpt = 4; }
    else if (F <= 2048 || F > 4096) { blk = 256; fpt = 8; } else { blk = 256; fpt = 16; }
    if (ctx->voice_fpt) { fpt = ctx->voice_fpt; blk = (int)(ceil_div_u64(ceil_div_u64(F, fpt), 32) * 32); if (blk > 256) blk = 256; }
    const uint64_t gy = ceil_div_u64(F, (uint64_t)blk * fpt);
    // Voice tiles.  At most 2048 voices (16 KB of shared memory) per tile; when the buses and frame ranges alone are
    // fewer blocks than the chip holds at once (2048 threads per SM), a bus is cut into tiles so that the grid is a
    // whole number of blocks per SM -- but not below 64 voices per tile.
    {
        // (a pipelined bus keeps one slot for the block that completes the previous exchange, so that it runs from the start)
        const uint64_t per_sm = 2048 / (uint64_t)blk, slots = (uint64_t)ctx->n_sm * per_sm - (b->bus && b->bus_mode == 2 ? 1 : 0), rows = p.n_bus * gy;
        uint64_t tpb = ceil_div_u64(p.G, 2048);
        if (rows * tpb < slots || (rows * tpb) % slots) {
            const uint64_t waves = ceil_div_u64(rows * tpb, slots);
            uint64_t want = ceil_div_u64(waves * slots, rows);              // tiles per bus for `waves` full waves
            if (rows > slots) want = tpb;
            while (want > tpb && ceil_div_u64(p.G, want) < 64) --want;
            if (want > tpb) tpb = want;
        }
        p.tile = (uint32_t)ceil_div_u64(p.G, tpb);
        p.tiles_per_bus = (uint32_t)ceil_div_u64(p.G, p.tile);
    }
    p.advance = gy == 1;
    p.vec = (float *)io->out;
    p.isum = (int32_t *)io->mix;
    p.acc = nullptr; p.tickets = nullptr;
    if (p.tiles_per_bus > 1) {
        const size_t need = sizeof(uint32_t) * (p.n_bus * F + p.n_bus * gy);
        if (b->cap_acc < need) {
            if (b->d_acc) cudaFree(b->d_acc);
            b->d_acc = nullptr; b->cap_acc = 0;
            CK(ctx, cudaMalloc(&b->d_acc, need));
            CK(ctx, cudaMemsetAsync(b->d_acc, 0, need, ctx->stream));   // once: every launch leaves the rows and tickets zero
            b->cap_acc = need;
        }
        p.acc = b->d_acc; p.tickets = b->d_acc + p.n_bus * F;
    }
    const bool sq = p.mode == CPROC_CUDA_MIX_SQUARE;
    p.n_tile_blocks = (uint32_t)(p.n_bus * p.tiles_per_bus);
    BusFused bf;
    int rc = cproc_bus_fused_begin(b, &bf, p.n_bus * F, sq ? 1u : 0u, sq ? 2u : 1u, (uint32_t)(p.n_bus * gy), (int32_t *)io->mix, (float *)io->out);
    if (rc) return rc;
    dim3 grid(p.n_tile_blocks + (bf.world && bf.mode == 2 ? 1u : 0u), (unsigned)gy);
    size_t smem = sizeof(uint2) * (p.tile + 1);
#define VOICE_LAUNCH(FPT) do { \
        if (sq) k_voice_mix<FPT, true><<<grid, blk, smem, ctx->stream>>>(p, bf); \
        else k_voice_mix<FPT, false><<<grid, blk, smem, ctx->stream>>>(p, bf); } while (0)
    if (fpt == 1) VOICE_LAUNCH(1); else if (fpt == 2) VOICE_LAUNCH(2); else if (fpt == 4) VOICE_LAUNCH(4); else if (fpt == 8) VOICE_LAUNCH(8); else VOICE_LAUNCH(16);
#undef VOICE_LAUNCH
    CK_LAUNCH(ctx, "k_voice_mix");
    if (!p.advance) {
        k_voice_advance<<<(unsigned)ceil_div_u64(b->n, 256), 256, 0, ctx->stream>>>(b->d_state, b->npad, b->n, (uint32_t)F);
        CK_LAUNCH(ctx, "k_voice_advance");
    }
    return 0;
}
