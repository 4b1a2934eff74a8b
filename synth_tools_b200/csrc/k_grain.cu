// k_grain.cu -- square_grain~ (linux/synth_tools.c:85-100): a Schmitt-trigger
// squarer with a one-sample output delay.
//
//   out[i] = state;
//   if (state >= 0 && in[i] < -thresh) state = -0.5f;
//   else if (state < 0 && in[i] > thresh) state = +0.5f;
//
// Outputs are exactly {0, +0.5, -0.5}, so the result is bit-exact.  The
// hysteresis makes the recurrence a 3-state machine that is serial in time;
// the grain axis carries the parallelism (one thread per grain).
//
// k_grain_planar: hosts hand over planar [grain][F] vectors.  A warp owns 32
//   grains and walks the frame axis in 32-frame tiles: the tile is loaded with
//   32 coalesced 128-byte row reads into padded shared memory, each lane then
//   runs its own grain along its row IN PLACE (in may alias out, as in Pd),
//   and the tile is written back with 32 coalesced row stores.
// k_grain_interleaved: [F][grain] streams are already coalesced.
// k_grain_mix: config C3b -- input from a per-grain phasor, integer stereo mix.
#include "common.cuh"
#include <cuda.h>

struct GrainParams {
    float *state;            // SoA [1][npad]
    const float *thresh;     // SoA [1][npad]
    uint64_t n, F;
    const float *in;
    float *out;
};

__device__ __forceinline__ float grain_step(float &state, float val, float thresh) {
    const float o = state;                                   // :91
    if (state >= 0.0f && val < -thresh) state = -0.5f;       // :92-94
    else if (state < 0.0f && val > thresh) state = 0.5f;     // :95-97
    return o;
}

#define GRAIN_WARPS 4
__global__ void __launch_bounds__(GRAIN_WARPS * 32) k_grain_planar(const GrainParams p) {
    __shared__ float tile[GRAIN_WARPS][32][33];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t g0 = ((uint64_t)blockIdx.x * GRAIN_WARPS + warp) * 32;
    if (g0 >= p.n) return;
    const uint32_t rows = p.n - g0 < 32 ? (uint32_t)(p.n - g0) : 32u;
    float (*tl)[33] = tile[warp];
    const bool mine = lane < rows;
    float state = mine ? p.state[g0 + lane] : 0.0f;
    const float th = mine ? p.thresh[g0 + lane] : 0.0f;
    for (uint64_t t0 = 0; t0 < p.F; t0 += 32) {
        const uint32_t cols = p.F - t0 < 32 ? (uint32_t)(p.F - t0) : 32u;
        if (lane < cols)
            for (uint32_t r = 0; r < rows; ++r) tl[r][lane] = __ldcs(p.in + (g0 + r) * p.F + t0 + lane);
        __syncwarp();
        if (mine)
            for (uint32_t i = 0; i < cols; ++i) tl[lane][i] = grain_step(state, tl[lane][i], th);
        __syncwarp();
        if (lane < cols)
            for (uint32_t r = 0; r < rows; ++r) __stcs(p.out + (g0 + r) * p.F + t0 + lane, tl[r][lane]);
        __syncwarp();
    }
    if (mine) p.state[g0 + lane] = state;
}

// [F][grain] streams: a warp's 32 grains are 128 contiguous bytes of every frame row.
// in and out may be the same buffer, so the compiler cannot move loads across stores:
// the 16-frame batches make the independent loads explicit (all in flight, then the
// serial trigger steps, then the stores).  Dynamic shared memory is requested only to
// pin the number of resident blocks per SM (see grain_residency).
#define GI_BLOCK 128
#define GI_BATCH 16
__global__ void __launch_bounds__(GI_BLOCK) k_grain_interleaved(const GrainParams p) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n) return;
    float state = p.state[g];
    const float th = p.thresh[g];
    const float *src = p.in + g;
    float *dst = p.out + g;
    uint64_t t = 0;
    for (; t + GI_BATCH <= p.F; t += GI_BATCH) {
        float v[GI_BATCH];
#pragma unroll
        for (int i = 0; i < GI_BATCH; ++i) v[i] = __ldcs(src + (t + i) * p.n);
#pragma unroll
        for (int i = 0; i < GI_BATCH; ++i) v[i] = grain_step(state, v[i], th);
#pragma unroll
        for (int i = 0; i < GI_BATCH; ++i) __stcs(dst + (t + i) * p.n, v[i]);
    }
    for (; t < p.F; ++t) {
        const float v = __ldcs(src + t * p.n);
        __stcs(dst + t * p.n, grain_step(state, v, th));
    }
    p.state[g] = state;
}

// Four adjacent grains per thread (128-bit loads/stores): a block covers 2 KiB of every
// frame row, which quarters the number of distant row segments (and pages) a block walks.
#ifndef GI4_BATCH
#define GI4_BATCH 8
#endif
#ifndef GI4_MINB
#define GI4_MINB 1
#endif
__global__ void __launch_bounds__(GI_BLOCK, GI4_MINB) k_grain_interleaved4(const GrainParams p) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (g >= p.n) return;                                    // n % 4 == 0
    float4 st = *(const float4 *)(p.state + g);
    const float4 th = *(const float4 *)(p.thresh + g);
    const float *src = p.in + g;
    float *dst = p.out + g;
    auto step4 = [&](float4 v) {
        float4 o;
        o.x = grain_step(st.x, v.x, th.x); o.y = grain_step(st.y, v.y, th.y);
        o.z = grain_step(st.z, v.z, th.z); o.w = grain_step(st.w, v.w, th.w);
        return o;
    };
    uint64_t t = 0;
    for (; t + GI4_BATCH <= p.F; t += GI4_BATCH) {
        float4 v[GI4_BATCH];
#pragma unroll
        for (int i = 0; i < GI4_BATCH; ++i) v[i] = __ldcs((const float4 *)(src + (t + i) * p.n));
#pragma unroll
        for (int i = 0; i < GI4_BATCH; ++i) v[i] = step4(v[i]);
#pragma unroll
        for (int i = 0; i < GI4_BATCH; ++i) __stcs((float4 *)(dst + (t + i) * p.n), v[i]);
    }
    for (; t < p.F; ++t) __stcs((float4 *)(dst + t * p.n), step4(__ldcs((const float4 *)(src + t * p.n))));
    *(float4 *)(p.state + g) = st;
}

// Blocks of equal duration run in rounds of (SMs x resident blocks): pick the residency
// (within [r_min, r_max]) whose last round is fullest, and the shared-memory request that
// makes the hardware hold exactly that many blocks per SM.
static int grain_residency(uint64_t blocks, int n_sm, int r_min, int r_max, size_t *smem) {
    int best = r_max; double best_eff = 0.0;
    for (int r = r_max; r >= r_min; --r) {
        const uint64_t per_round = (uint64_t)n_sm * r;
        const double eff = (double)blocks / (double)(per_round * ceil_div_u64(blocks, per_round));
        if (eff > best_eff + 0.01) { best_eff = eff; best = r; }
    }
    *smem = (size_t)(227 * 1024) / best - 1024;          // 1 KiB per block is reserved by the system
    return best;
}

// ---------------------------------------------------------------------------
// k_grain_bulk: the planar layout without a register transpose.  Lane r of a warp
// owns grain r of the warp's 32.  Per tile of TF frames every lane issues ONE bulk
// copy (cp.async.bulk, the non-tensor TMA path) of its own row segment -- TF*4
// contiguous bytes of HBM -- into its row of a shared-memory stage; a per-stage
// mbarrier counts the bytes in.  The lane then walks its row with LDS.128 /
// STS.128 (row stride TF*4+16 bytes: the 16-byte units of 8 consecutive lanes fall
// in 8 different bank groups) and sends the row back with one bulk store.  Loads
// run one tile ahead of the compute, stores drain one tile behind it.  No thread
// touches another thread's data, so the only synchronisation is the mbarrier.
// `in` may alias `out` exactly (Pd in-place): a row segment is always read before
// it is written.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// One frame of the fast form: state is the predicate `neg`; valid for state = +-0.5 and
// thresh >= 0 (then at most one of the two flip conditions holds, synth_tools.c:92-97).
__device__ __forceinline__ float grain_step_p(bool &neg, float val, float th, float nth) {
    const float o = neg ? -0.5f : 0.5f;                      // :91
    neg = (val < nth) | (neg & !(val > th));
    return o;
}

#define GB_WARPS 2
template <int TF, int STAGES>
__global__ void __launch_bounds__(GB_WARPS * 32) k_grain_bulk(const GrainParams p) {
    constexpr uint32_t ROWB = TF * 4 + 16;
    constexpr uint32_t STAGEB = 32 * ROWB;
    extern __shared__ __align__(128) uint8_t gb_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t g0 = ((uint64_t)blockIdx.x * GB_WARPS + warp) * 32;
    if (g0 >= p.n) return;
    const uint32_t rows = p.n - g0 < 32 ? (uint32_t)(p.n - g0) : 32u;
    const bool mine = lane < rows;
    const uint32_t base = smem_u32(gb_smem) + warp * (STAGES * STAGEB);
    const uint32_t bar0 = smem_u32(gb_smem) + GB_WARPS * STAGES * STAGEB + warp * (STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float state = mine ? p.state[g0 + lane] : 0.5f;
    const float th = mine ? p.thresh[g0 + lane] : 0.0f;
    const float nth = -th;
    const float *src = p.in + (g0 + lane) * p.F;
    float *dst = p.out + (g0 + lane) * p.F;
    const uint32_t n_tiles = (uint32_t)((p.F + TF - 1) / TF);
    auto cols_of = [&](uint32_t k) { const uint64_t left = p.F - (uint64_t)k * TF; return left < TF ? (uint32_t)left : (uint32_t)TF; };
    auto issue = [&](uint32_t k) {
        const uint32_t s = k % STAGES, bytes = cols_of(k) * 4;
        if (lane == 0) mbar_expect_tx(bar0 + 8 * s, rows * bytes);
        if (mine) bulk_g2s(base + s * STAGEB + lane * ROWB, src + (uint64_t)k * TF, bytes, bar0 + 8 * s);
    };
#pragma unroll 1
    for (uint32_t k = 0; k < STAGES - 2 && k < n_tiles; ++k) issue(k);
#pragma unroll 1
    for (uint32_t k = 0; k < n_tiles; ++k) {
        // the stage tile k+STAGES-2 lands in last held tile k-2: its bulk store must have read it out
        if (k + STAGES - 2 < n_tiles) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            issue(k + STAGES - 2);
        }
        const uint32_t s = k % STAGES, cols = cols_of(k);
        mbar_wait(bar0 + 8 * s, (k / STAGES) & 1);
        const uint32_t row = base + s * STAGEB + lane * ROWB;
        if (mine) {
            if ((state == 0.5f || state == -0.5f) && th >= 0.0f && cols == TF) {
                bool neg = state < 0.0f;
#pragma unroll 4
                for (uint32_t c = 0; c < TF / 4; ++c) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(row + 16 * c));
                    v.x = grain_step_p(neg, v.x, th, nth); v.y = grain_step_p(neg, v.y, th, nth);
                    v.z = grain_step_p(neg, v.z, th, nth); v.w = grain_step_p(neg, v.w, th, nth);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + 16 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                state = neg ? -0.5f : 0.5f;
            } else {
                // initial 0.0 (or any other) state value, negative threshold, ragged last tile: literal form
                for (uint32_t c = 0; c < cols / 4; ++c) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(row + 16 * c));
                    v.x = grain_step(state, v.x, th); v.y = grain_step(state, v.y, th);
                    v.z = grain_step(state, v.z, th); v.w = grain_step(state, v.w, th);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + 16 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy row writes -> visible to the bulk store
            bulk_s2g(dst + (uint64_t)k * TF, row, cols * 4);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (mine) p.state[g0 + lane] = state;
}

template <int TF, int STAGES>
static int launch_grain_bulk(cproc_cuda_ctx *ctx, const GrainParams &p) {
    constexpr size_t smem = (size_t)GB_WARPS * STAGES * 32 * (TF * 4 + 16) + GB_WARPS * STAGES * 8;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        CK(ctx, cudaFuncSetAttribute(k_grain_bulk<TF, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    k_grain_bulk<TF, STAGES><<<(unsigned)ceil_div_u64(p.n, GB_WARPS * 32), GB_WARPS * 32, smem, ctx->stream>>>(p);
    return 0;
}

// ---------------------------------------------------------------------------
// k_grain_tma: the same lane-owns-a-row scheme with TENSOR TMA.  k_grain_bulk issues one bulk copy
// per lane, which the hardware's uniform datapath executes one lane at a time (the SASS loops over
// the 32 lanes around every UBLKCP: ~640 of the ~900 warp instructions per tile are copy issue).
// Here the [grain][frame] stream is described once per launch by a 2-D tensor map (box = 32 frames
// x 32 grains = one 128-byte row segment per lane, SWIZZLE_128B) and ONE elected lane issues two
// cp.async.bulk.tensor loads and two stores per 64-frame tile.  The 128-byte swizzle XORs the
// 16-byte chunk index with (row & 7): lane r finds chunk c of its row at r*128 + ((c ^ (r & 7)) << 4),
// so the LDS.128 / STS.128 of eight consecutive lanes fall in eight different bank groups.
// Out-of-range rows / frames are zero-filled on load and clipped on store by the TMA unit.
#define GT_WARPS 2
#define GT_STAGES 3
#define GT_BOXB 4096                       // 32 rows x 128 B
#define GT_STAGEB (2 * GT_BOXB)            // 64 frames
__global__ void __launch_bounds__(GT_WARPS * 32) k_grain_tma(const GrainParams p, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ __align__(1024) uint8_t gt_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t g0 = ((uint64_t)blockIdx.x * GT_WARPS + warp) * 32;
    if (g0 >= p.n) return;
    const uint32_t rows = p.n - g0 < 32 ? (uint32_t)(p.n - g0) : 32u;
    const bool mine = lane < rows;
    const uint32_t sm0 = (smem_u32(gt_smem) + 1023u) & ~1023u;                 // swizzle atoms are 1 KiB
    const uint32_t base = sm0 + warp * (GT_STAGES * GT_STAGEB);
    const uint32_t bar0 = sm0 + GT_WARPS * GT_STAGES * GT_STAGEB + warp * (GT_STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < GT_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float state = mine ? p.state[g0 + lane] : 0.5f;
    const float th = mine ? p.thresh[g0 + lane] : 0.0f;
    const float nth = -th;
    const uint32_t n_tiles = (uint32_t)((p.F + 63) / 64);
    const uint64_t tmi = reinterpret_cast<uint64_t>(&tm_in), tmo = reinterpret_cast<uint64_t>(&tm_out);
    auto cols_of = [&](uint32_t k) { const uint64_t left = p.F - (uint64_t)k * 64; return left < 64 ? (uint32_t)left : 64u; };
    auto issue = [&](uint32_t k) {                                             // lane 0 only
        const uint32_t s = k % GT_STAGES, nbox = cols_of(k) > 32 ? 2u : 1u;
        mbar_expect_tx(bar0 + 8 * s, nbox * GT_BOXB);
        for (uint32_t h = 0; h < nbox; ++h)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(base + s * GT_STAGEB + h * GT_BOXB), "l"(tmi), "r"((int32_t)(k * 64 + h * 32)), "r"((int32_t)g0), "r"(bar0 + 8 * s) : "memory");
    };
    if (lane == 0) for (uint32_t k = 0; k < GT_STAGES - 2 && k < n_tiles; ++k) issue(k);
#pragma unroll 1
    for (uint32_t k = 0; k < n_tiles; ++k) {
        if (k + GT_STAGES - 2 < n_tiles) {                                     // that stage last held tile k-2: its stores must have read it out
            if (lane == 0) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); issue(k + GT_STAGES - 2); }
        }
        const uint32_t s = k % GT_STAGES, cols = cols_of(k);
        mbar_wait(bar0 + 8 * s, (k / GT_STAGES) & 1);
        if (mine) {
            const bool fast = (state == 0.5f || state == -0.5f) && th >= 0.0f;
            bool neg = state < 0.0f;
            for (uint32_t c = 0; c < cols / 4; ++c) {
                const uint32_t a = base + s * GT_STAGEB + (c >> 3) * GT_BOXB + lane * 128 + (((c & 7) ^ (lane & 7)) << 4);
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                if (fast) {
                    v.x = grain_step_p(neg, v.x, th, nth); v.y = grain_step_p(neg, v.y, th, nth);
                    v.z = grain_step_p(neg, v.z, th, nth); v.w = grain_step_p(neg, v.w, th, nth);
                } else {                                                       // initial 0.0 (or any other) state value, negative threshold: literal form
                    v.x = grain_step(state, v.x, th); v.y = grain_step(state, v.y, th);
                    v.z = grain_step(state, v.z, th); v.w = grain_step(state, v.w, th);
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
            if (fast) state = neg ? -0.5f : 0.5f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            const uint32_t nbox = cols > 32 ? 2u : 1u;
            for (uint32_t h = 0; h < nbox; ++h)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(tmo), "r"((int32_t)(k * 64 + h * 32)), "r"((int32_t)g0), "r"(base + s * GT_STAGEB + h * GT_BOXB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (mine) p.state[g0 + lane] = state;
}

typedef CUresult (*cuTensorMapEncodeTiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cuTensorMapEncodeTiled_t grain_encode_fn() {
    static cuTensorMapEncodeTiled_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (cuTensorMapEncodeTiled_t)q;
        else cudaGetLastError();
    }
    return fn;
}

// returns 1 when the TMA kernel was launched, 0 when the caller should use another kernel, < 0 on error
static int launch_grain_tma(cproc_cuda_ctx *ctx, const GrainParams &p) {
    cuTensorMapEncodeTiled_t enc = grain_encode_fn();
    if (!enc) return 0;
    CUtensorMap tin, tout;
    const cuuint64_t dims[2] = {p.F, p.n}, strides[1] = {p.F * sizeof(float)};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    if (enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)p.in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 0;
    if (enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)p.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 0;
    constexpr size_t smem = (size_t)GT_WARPS * GT_STAGES * GT_STAGEB + GT_WARPS * GT_STAGES * 8 + 1024;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        CK(ctx, cudaFuncSetAttribute(k_grain_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    k_grain_tma<<<(unsigned)ceil_div_u64(p.n, GT_WARPS * 32), GT_WARPS * 32, smem, ctx->stream>>>(p, tin, tout);
    return 1;
}

int launch_square_grain(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->in || !io->out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "square_grain: in/out is NULL");
    if (io->layout == CPROC_CUDA_TILED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "square_grain: TILED layout not supported");
    if (F == 0) return 0;
    GrainParams p;
    p.state = (float *)b->d_state; p.thresh = (const float *)b->d_param; p.n = b->n; p.F = F;
    p.in = (const float *)io->in; p.out = (float *)io->out;
    // bulk copies move 16-byte units: rows must start and end on 16-byte boundaries
    const bool bulk_ok = ctx->grain_bulk && F % 4 == 0 && ((uintptr_t)p.in & 15) == 0 && ((uintptr_t)p.out & 15) == 0;
    if (io->layout == CPROC_CUDA_INTERLEAVED) {
        const bool vec4 = ctx->grain_vec4 && p.n % 4 == 0 && ((uintptr_t)p.in & 15) == 0 && ((uintptr_t)p.out & 15) == 0;
        const uint64_t blocks = ceil_div_u64(p.n, vec4 ? GI_BLOCK * 4 : GI_BLOCK);
        size_t smem = 0;
        grain_residency(blocks, ctx->n_sm, 6, 16, &smem);
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            CK(ctx, cudaFuncSetAttribute(k_grain_interleaved, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 6));
            CK(ctx, cudaFuncSetAttribute(k_grain_interleaved4, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 6));
            attr_set[ctx->device & 63] = true;
        }
        if (vec4) k_grain_interleaved4<<<(unsigned)blocks, GI_BLOCK, smem, ctx->stream>>>(p);
        else k_grain_interleaved<<<(unsigned)blocks, GI_BLOCK, smem, ctx->stream>>>(p);
    } else if (bulk_ok) {
        int rc = 0;
        // tensor TMA: rows of F * 4 bytes must be 16-byte multiples (F % 4, checked) and fewer than 2^32 elements per dimension
        if (ctx->grain_bulk == 5 && p.F < (1ull << 31) && p.n < (1ull << 31)) {
            rc = launch_grain_tma(ctx, p);
            if (rc < 0) return rc;
        }
        if (rc == 1) { /* launched */ }
        else switch (ctx->grain_bulk) {
        case 2: rc = launch_grain_bulk<128, 3>(ctx, p); break;
        case 3: rc = launch_grain_bulk<64, 4>(ctx, p); break;
        case 4: rc = launch_grain_bulk<32, 4>(ctx, p); break;
        default: rc = launch_grain_bulk<64, 3>(ctx, p); break;
        }
        if (rc < 0) return rc;
    } else
        k_grain_planar<<<(unsigned)ceil_div_u64(p.n, GRAIN_WARPS * 32), GRAIN_WARPS * 32, 0, ctx->stream>>>(p);
    CK_LAUNCH(ctx, "k_grain");
    return 0;
}

// ---------------------------------------------------------------------------
// C3b: per-grain phasor input, dyadic pan, integer stereo mix (units of 2^-7).
struct GrainMixParams {
    uint32_t *st;            // SoA [2][npad]: state (float bits), phase
    const uint32_t *prm;     // SoA [4][npad]: threshold (float bits), inc, gl, gr
    uint64_t npad, n, F;
    int32_t *imix;           // [2][F], zeroed by the launcher
};

#define GMIX_BLOCK 256
#define GMIX_CHUNK 32
__global__ void __launch_bounds__(GMIX_BLOCK) k_grain_mix(const GrainMixParams p) {
    // Each lane keeps one grain in registers.  Left and right gains (0..64) are
    // packed into one word so one REDUX.SUM per frame reduces both channels
    // over the warp; warp sums are accumulated per block in shared memory and
    // flushed to the global integer mix once per chunk of frames.
    __shared__ int32_t acc[2][GMIX_CHUNK];
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = g < p.n;
    const uint32_t lane = threadIdx.x & 31;
    float state = 0.0f, th = 0.0f;
    uint32_t ph = 0, inc = 0;
    int32_t pk = 0;
    if (mine) {
        state = __uint_as_float(p.st[g]); ph = p.st[p.npad + g];
        th = __uint_as_float(p.prm[g]); inc = p.prm[p.npad + g];
        pk = (int32_t)((p.prm[3 * p.npad + g] << 16) + p.prm[2 * p.npad + g]);
    }
    for (uint64_t t0 = 0; t0 < p.F; t0 += GMIX_CHUNK) {
        if (threadIdx.x < 2 * GMIX_CHUNK) (&acc[0][0])[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t cols = p.F - t0 < GMIX_CHUNK ? (uint32_t)(p.F - t0) : GMIX_CHUNK;
        int32_t keep = 0;
        for (uint32_t i = 0; i < cols; ++i) {
            const float val = __int2float_rn((int32_t)ph) * 0x1p-31f;   // acc -> signed saw
            ph += inc;                                                  // cproc.h:141
            const float o = grain_step(state, val, th);
            const int32_t c = o > 0.0f ? pk : (o < 0.0f ? -pk : 0);
            const int32_t s = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == (i & 31)) keep = s;
        }
        if (lane < cols) {
            const int32_t l = (int32_t)(int16_t)(keep & 0xFFFF);        // |sum L| <= 32*64
            const int32_t r = (keep - l) >> 16;
            atomicAdd(&acc[0][lane], l);
            atomicAdd(&acc[1][lane], r);
        }
        __syncthreads();
        if (threadIdx.x < cols) {
            atomicAdd(p.imix + t0 + threadIdx.x, acc[0][threadIdx.x]);
            atomicAdd(p.imix + p.F + t0 + threadIdx.x, acc[1][threadIdx.x]);
        }
        __syncthreads();
    }
    if (mine) { p.st[g] = __float_as_uint(state); p.st[p.npad + g] = ph; }
}

// ---------------------------------------------------------------------------
// C3b, second generation.  Two changes against k_grain_mix:
//  * the comparisons run on the integer phase.  val = (float)(int)ph * 2^-31 is a
//    monotone function of the signed phase, so `val < -thresh` is `x < lo` and
//    `val > thresh` is `x > hi` for two per-grain integers found once per parameter
//    upload by bisection with the very same I2F / FMUL / FSETP (k_grain_thresholds):
//    no conversion, no multiply in the loop, and still bit-exact.
//  * the mix leaves the inner loop: a thread keeps 32 packed (right << 16) + left
//    accumulators for a 64-frame chunk in registers and walks ITS grains through the
//    chunk one after the other; the state machine output m in {-1, 0, +1} enters the
//    bus as one IMAD (acc += m * pk).  One warp reduction per frame per chunk instead
//    of one per grain-sample.
// State on the device stays the reference float (0, +0.5, -0.5); other magnitudes are
// treated by their sign, as the integer mix always did, and come back as +-0.5.
__global__ void k_grain_thresholds(const uint32_t *prm, uint64_t n, uint64_t npad, int32_t *lo, int32_t *hi, int32_t *pk, uint32_t *weird) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float th = __uint_as_float(prm[g]);
    pk[g] = (int32_t)((prm[3 * npad + g] << 16) + prm[2 * npad + g]);
    // a negative threshold lets both flip conditions hold at once; only the literal
    // two-branch form (k_grain_mix / k_grain_mix2) orders them like the reference
    if (th < 0.0f) atomicOr(weird, 2u);
    auto val = [](int64_t x) { return __int2float_rn((int32_t)x) * 0x1p-31f; };
    // lo: smallest x with !(val(x) < -th); the predicate is true below, false from lo on
    int64_t a = INT32_MIN, b = INT32_MAX;
    if (!(val(a) < -th)) lo[g] = INT32_MIN;
    else if (val(b) < -th) { lo[g] = INT32_MAX; atomicOr(weird, 1u); }       // true for every phase: not representable
    else {
        while (b - a > 1) { const int64_t m = a + (b - a) / 2; if (val(m) < -th) a = m; else b = m; }
        lo[g] = (int32_t)b;
    }
    // hi: largest x with !(val(x) > th); the predicate is false up to hi, true above
    a = INT32_MIN; b = INT32_MAX;
    if (!(val(b) > th)) hi[g] = INT32_MAX;
    else if (val(a) > th) { hi[g] = INT32_MIN; atomicOr(weird, 1u); }
    else {
        while (b - a > 1) { const int64_t m = a + (b - a) / 2; if (val(m) > th) b = m; else a = m; }
        hi[g] = (int32_t)a;
    }
}

#define GM2_BLOCK 128
#define GM2_CHUNK 64
struct GrainMix2Params {
    GrainMixParams g;
    const int32_t *lo, *hi;
};

__device__ __forceinline__ void gm2_tick(int32_t &m, uint32_t &ph, uint32_t inc, int32_t pk, int32_t lo, int32_t hi, int32_t &acc) {
    const int32_t x = (int32_t)ph;
    ph += inc;                                              // cproc.h:141
    acc = m * pk + acc;                                     // out = state (synth_tools.c:91), panned into the bus
    // if (state >= 0 && in < -thresh) state = -0.5; else if (state < 0 && in > thresh) state = +0.5;   (:92-97)
    // three compares with the sign predicate folded in, two selects (ptxas makes 8+ of the C form)
    asm("{ .reg .pred pn, pd, pu;\n\t"
        "setp.lt.s32 pn, %0, 0;\n\t"
        "setp.lt.and.s32 pd, %1, %2, !pn;\n\t"
        "setp.gt.and.s32 pu, %1, %3, pn;\n\t"
        "selp.s32 %0, -1, %0, pd;\n\t"
        "selp.s32 %0, 1, %0, pu; }" : "+r"(m) : "r"(x), "r"(lo), "r"(hi));
}

struct GM2Grain { float st; uint32_t ph, inc; int32_t pk, lo, hi; };
__device__ __forceinline__ void gm2_load(GM2Grain &v, const GrainMix2Params &p, uint64_t g) {
    const uint64_t npad = p.g.npad;
    v.st = __uint_as_float(__ldcg(p.g.st + g));
    v.ph = __ldcg(p.g.st + npad + g);
    v.inc = __ldg(p.g.prm + npad + g);
    v.pk = (int32_t)((__ldg(p.g.prm + 3 * npad + g) << 16) + __ldg(p.g.prm + 2 * npad + g));
    v.lo = __ldg(p.lo + g); v.hi = __ldg(p.hi + g);
}

__global__ void __launch_bounds__(GM2_BLOCK) k_grain_mix2(const GrainMix2Params p) {
    __shared__ int32_t sacc[2][GM2_CHUNK];
    const uint64_t T = (uint64_t)gridDim.x * GM2_BLOCK;
    const uint64_t tid = (uint64_t)blockIdx.x * GM2_BLOCK + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n = p.g.n, npad = p.g.npad;
    for (uint64_t t0 = 0; t0 < p.g.F; t0 += GM2_CHUNK) {
        const uint32_t cols = p.g.F - t0 < GM2_CHUNK ? (uint32_t)(p.g.F - t0) : GM2_CHUNK;
        if (threadIdx.x < 2 * GM2_CHUNK) (&sacc[0][0])[threadIdx.x] = 0;
        __syncthreads();
        int32_t acc[GM2_CHUNK];
#pragma unroll
        for (int k = 0; k < GM2_CHUNK; ++k) acc[k] = 0;
        // software pipeline: the next grain's seven words are in flight while this one runs its 32 frames
        GM2Grain nx = {};
        if (tid < n) gm2_load(nx, p, tid);
        for (uint64_t g = tid; g < n; g += T) {
            const GM2Grain cur = nx;
            if (g + T < n) gm2_load(nx, p, g + T);
            int32_t m = cur.st > 0.0f ? 1 : (cur.st < 0.0f ? -1 : 0);
            uint32_t ph = cur.ph;
            const uint32_t inc = cur.inc;
            const int32_t pk = cur.pk, lo = cur.lo, hi = cur.hi;
            if (cols == GM2_CHUNK) {
#pragma unroll
                for (int k = 0; k < GM2_CHUNK; ++k) gm2_tick(m, ph, inc, pk, lo, hi, acc[k]);
            } else {
#pragma unroll
                for (int k = 0; k < GM2_CHUNK; ++k) if (k < (int)cols) gm2_tick(m, ph, inc, pk, lo, hi, acc[k]);
            }
            if (m != 0) p.g.st[g] = __float_as_uint(0.5f * (float)m);    // never flipped from 0.0: unchanged
            p.g.st[npad + g] = ph;
        }
#pragma unroll
        for (int h = 0; h < GM2_CHUNK / 32; ++h) {
            int32_t keepl = 0, keepr = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int32_t a = acc[h * 32 + k];
                const int32_t l = (int32_t)(int16_t)(a & 0xFFFF);        // |sum L| <= grains_per_thread * 64 < 2^15
                const int32_t r = (a - l) >> 16;
                const int32_t sl = __reduce_add_sync(0xFFFFFFFFu, l), sr = __reduce_add_sync(0xFFFFFFFFu, r);
                if (lane == k) { keepl = sl; keepr = sr; }
            }
            if (h * 32 + lane < cols) { atomicAdd(&sacc[0][h * 32 + lane], keepl); atomicAdd(&sacc[1][h * 32 + lane], keepr); }
        }
        __syncthreads();
        if (threadIdx.x < cols) {
            atomicAdd(p.g.imix + t0 + threadIdx.x, sacc[0][threadIdx.x]);
            atomicAdd(p.g.imix + p.g.F + t0 + threadIdx.x, sacc[1][threadIdx.x]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// C3b, third generation: the trigger state lives in a predicate.  For thresh >= 0
// (lo <= hi) the two-branch update of synth_tools.c:92-97 is the Schmitt trigger
//     neg' = (x < lo) | (neg & (x <= hi))
// which is two ISETP with a predicate operand (ISETP.LE.AND, ISETP.LT.OR); the bus
// term is `neg ? -pk : pk` (one SEL) added by an IMAD.IADD, the phasor a second
// IMAD.IADD: 5 instructions per grain-sample, 3 on the ALU pipe, 2 on the FMA pipe
// (k_grain_mix2: 7 and 5).  A grain whose state is still the initial 0.0 outputs 0
// until its first flip (synth_tools.c:91: out = state), so it walks its chunk with
// the literal form; once flipped it never returns to 0.
#define GM3_BLOCK 128
#define GM3_CHUNK 64
struct GrainMix3Params {
    GrainMixParams g;
    const int32_t *lo, *hi, *pk;
};
struct GM3Grain { float st; uint32_t ph, inc; int32_t pk, lo, hi; };
__device__ __forceinline__ void gm3_load(GM3Grain &v, const GrainMix3Params &p, uint64_t g) {
    v.st = __uint_as_float(__ldcg(p.g.st + g));
    v.ph = __ldcg(p.g.st + p.g.npad + g);
    v.inc = __ldg(p.g.prm + p.g.npad + g);
    v.pk = __ldg(p.pk + g); v.lo = __ldg(p.lo + g); v.hi = __ldg(p.hi + g);
}

__device__ __forceinline__ void gm3_chunk(const GM3Grain &cur, uint32_t cols, int32_t (&acc)[GM3_CHUNK], float &st_out, uint32_t &ph_out) {
    uint32_t ph = cur.ph;
    const uint32_t inc = cur.inc;
    const int32_t pk = cur.pk, npk = -cur.pk, lo = cur.lo, hi = cur.hi;
    if (cols == GM3_CHUNK && cur.st != 0.0f) {
        bool neg = cur.st < 0.0f;
#pragma unroll
        for (int k = 0; k < GM3_CHUNK; ++k) {
            const int32_t x = (int32_t)ph;
            ph += inc;                                       // cproc.h:141
            acc[k] += neg ? npk : pk;                        // out = state (:91), panned into the bus
            neg = (x < lo) | (neg & (x <= hi));              // :92-97
        }
        st_out = neg ? -0.5f : 0.5f;
    } else {
        // initial 0.0 state or a ragged last chunk: the literal form, tick by tick
        int32_t m = cur.st > 0.0f ? 1 : (cur.st < 0.0f ? -1 : 0);
#pragma unroll
        for (int k = 0; k < GM3_CHUNK; ++k)
            if (k < (int)cols) gm2_tick(m, ph, inc, pk, lo, hi, acc[k]);
        st_out = m != 0 ? 0.5f * (float)m : cur.st;
    }
    ph_out = ph;
}

__global__ void __launch_bounds__(GM3_BLOCK) k_grain_mix3(const GrainMix3Params p) {
    __shared__ int32_t sacc[2][GM3_CHUNK];
    const uint64_t T = (uint64_t)gridDim.x * GM3_BLOCK;
    const uint64_t tid = (uint64_t)blockIdx.x * GM3_BLOCK + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n = p.g.n, npad = p.g.npad;
    for (uint64_t t0 = 0; t0 < p.g.F; t0 += GM3_CHUNK) {
        const uint32_t cols = p.g.F - t0 < GM3_CHUNK ? (uint32_t)(p.g.F - t0) : GM3_CHUNK;
        if (threadIdx.x < 2 * GM3_CHUNK) (&sacc[0][0])[threadIdx.x] = 0;
        __syncthreads();
        int32_t acc[GM3_CHUNK];
#pragma unroll
        for (int k = 0; k < GM3_CHUNK; ++k) acc[k] = 0;
        GM3Grain nx = {};
        if (tid < n) gm3_load(nx, p, tid);
        for (uint64_t g = tid; g < n; g += T) {
            const GM3Grain cur = nx;
            if (g + T < n) gm3_load(nx, p, g + T);               // in flight while this grain runs its chunk
            float st; uint32_t ph;
            gm3_chunk(cur, cols, acc, st, ph);
            p.g.st[g] = __float_as_uint(st);
            p.g.st[npad + g] = ph;
        }
#pragma unroll
        for (int h = 0; h < GM3_CHUNK / 32; ++h) {
            int32_t keepl = 0, keepr = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int32_t a = acc[h * 32 + k];
                const int32_t l = (int32_t)(int16_t)(a & 0xFFFF);        // |sum L| <= grains_per_thread * 64 < 2^15
                const int32_t r = (a - l) >> 16;
                const int32_t sl = __reduce_add_sync(0xFFFFFFFFu, l), sr = __reduce_add_sync(0xFFFFFFFFu, r);
                if (lane == k) { keepl = sl; keepr = sr; }
            }
            if (h * 32 + lane < cols) { atomicAdd(&sacc[0][h * 32 + lane], keepl); atomicAdd(&sacc[1][h * 32 + lane], keepr); }
        }
        __syncthreads();
        if (threadIdx.x < cols) {
            atomicAdd(p.g.imix + t0 + threadIdx.x, sacc[0][threadIdx.x]);
            atomicAdd(p.g.imix + p.g.F + t0 + threadIdx.x, sacc[1][threadIdx.x]);
        }
        __syncthreads();
    }
}

__global__ void k_imix_to_float(const int32_t *imix, float *mix, uint64_t count, float scale) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) mix[i] = __int2float_rn(imix[i]) * scale;
}

int launch_square_grain_mix(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    if (!io->out && !io->mix) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "square_grain_mix: out and mix are both NULL");
    if (F == 0) return 0;
    int32_t *imix = (int32_t *)io->mix;
    if (!imix) {
        size_t need = sizeof(int32_t) * 2 * F;
        if (b->cap_mix < need) {
            if (b->d_mix) cudaFree(b->d_mix);
            b->d_mix = nullptr; b->cap_mix = 0;
            CK(ctx, cudaMalloc(&b->d_mix, need));
            b->cap_mix = need;
        }
        imix = (int32_t *)b->d_mix;
    }
    CK(ctx, cudaMemsetAsync(imix, 0, sizeof(int32_t) * 2 * F, ctx->stream));
    GrainMixParams p;
    p.st = b->d_state; p.prm = b->d_param; p.npad = b->npad; p.n = b->n; p.F = F; p.imix = imix;
    if (b->aux_dirty) {                                   // integer thresholds, once per parameter upload
        if (!b->d_aux) CK(ctx, cudaMalloc(&b->d_aux, sizeof(uint32_t) * (3 * b->npad + 4)));
        uint32_t *weird = b->d_aux + 3 * b->npad;
        CK(ctx, cudaMemsetAsync(weird, 0, sizeof(uint32_t), ctx->stream));
        k_grain_thresholds<<<(unsigned)ceil_div_u64(b->n, 128), 128, 0, ctx->stream>>>(b->d_param, b->n, b->npad, (int32_t *)b->d_aux, (int32_t *)b->d_aux + b->npad, (int32_t *)b->d_aux + 2 * b->npad, weird);
        CK_LAUNCH(ctx, "k_grain_thresholds");
        uint32_t h = 0;
        CK(ctx, cudaMemcpyAsync(&h, weird, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        b->aux_weird = h;
        b->aux_dirty = false;
    }
    if (ctx->grain_mix2 >= 2 && !b->aux_weird) {
        GrainMix3Params q;
        q.g = p; q.lo = (const int32_t *)b->d_aux; q.hi = q.lo + b->npad; q.pk = q.lo + 2 * b->npad;
        uint64_t blocks = ceil_div_u64(p.n, GM3_BLOCK);
        const uint64_t cap = (uint64_t)ctx->n_sm * (uint64_t)ctx->grain_blocks_per_sm;
        if (blocks > cap) blocks = cap;
        const uint64_t min_blocks = ceil_div_u64(p.n, (uint64_t)GM3_BLOCK * 500);   // packed 16-bit bus fields: <= 500 grains per thread
        if (blocks < min_blocks) blocks = min_blocks;
        k_grain_mix3<<<(unsigned)blocks, GM3_BLOCK, 0, ctx->stream>>>(q);
        CK_LAUNCH(ctx, "k_grain_mix3");
    } else if (ctx->grain_mix2 && !(b->aux_weird & 1u)) {
        GrainMix2Params q;
        q.g = p; q.lo = (const int32_t *)b->d_aux; q.hi = (const int32_t *)b->d_aux + b->npad;
        uint64_t blocks = ceil_div_u64(p.n, GM2_BLOCK);
        const uint64_t cap = (uint64_t)ctx->n_sm * (uint64_t)ctx->grain_blocks_per_sm;
        if (blocks > cap) blocks = cap;
        const uint64_t min_blocks = ceil_div_u64(p.n, (uint64_t)GM2_BLOCK * 500);   // packed 16-bit bus fields: <= 500 grains per thread
        if (blocks < min_blocks) blocks = min_blocks;
        k_grain_mix2<<<(unsigned)blocks, GM2_BLOCK, 0, ctx->stream>>>(q);
        CK_LAUNCH(ctx, "k_grain_mix2");
    } else {
        k_grain_mix<<<(unsigned)ceil_div_u64(p.n, GMIX_BLOCK), GMIX_BLOCK, 0, ctx->stream>>>(p);
        CK_LAUNCH(ctx, "k_grain_mix");
    }
    if (io->out) {
        k_imix_to_float<<<(unsigned)ceil_div_u64(2 * F, 256), 256, 0, ctx->stream>>>(imix, (float *)io->out, 2 * F, 0x1p-7f);
        CK_LAUNCH(ctx, "k_imix_to_float");
    }
    return 0;
}
