// k_graph_scan.cu -- time-parallel, bit-exact render of acc / edge graphs (few instances, long
// streams: the reference's own configuration is ONE voice, linux/test_cproc.c).
//
// A thread per instance walks time serially; with a handful of instances that leaves the chip
// idle.  Both reference processors are scannable over the time axis, exactly (integers):
//   acc   out_t = out_init + SUM of the inputs at the executed ticks <= t       (cproc.h:140-142)
//         chunk summary: the partial sum; composition: +
//   edge  out = (in != last); last = in, at executed ticks only                  (cproc.h:151-154)
//         a chunk acts on (out, last) as one of three maps -- IDENT (no executed tick),
//         ONE(X) (one executed tick: out = (X != last_in), last = X), CONST(O, X) (two or
//         more) -- and these compose associatively (ONE after anything that fixes `last`
//         becomes CONST).
// The graph is evaluated node by node in ANF order; per node: chunk summaries (k_gs_reduce),
// a block scan of the summaries over the chunk axis (k_gs_scan, one block per instance),
// and the chunk walk from the true chunk-start state that writes the node's output stream
// (k_gs_apply).  Streams of interior nodes live in scratch [node][inst][F]; the output node
// writes the caller's buffer.  Masks (`changed & cond_mask`) are honoured tick by tick.
#include "common.cuh"

struct GsNode {
    uint32_t kind, mask;          // CPROC_CUDA_NODE_ACC / _EDGE
    const uint32_t *x;            // source stream
    uint64_t xs_i, xs_t;          // source index = i * xs_i + t * xs_t
    uint32_t *y;                  // this node's output stream
    uint64_t ys_i, ys_t;
    const uint32_t *g;            // changed stream or null
    uint64_t gs_i, gs_t;
    uint32_t *st;                 // state row of word 0 (.out); word 1 (.last) is st + npad
    uint64_t npad, n, F, L, C;
    uint4 *part;                  // [C][n] chunk summaries: acc {sum,-,-,-}; edge {type, X, O, -}
    uint2 *start;                 // [C][n] state at the first tick of chunk c: {out, last}
};

enum { GS_IDENT = 0, GS_ONE = 1, GS_CONST = 2 };

__device__ __forceinline__ bool gs_exec(const GsNode &p, uint64_t i, uint64_t t) {
    return !p.g || (p.g[i * p.gs_i + t * p.gs_t] & p.mask) != 0;
}

// threads enumerate (chunk, instance) pairs, instance fastest
__global__ void __launch_bounds__(128) k_gs_reduce(const GsNode p) {
    const uint64_t gidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i = gidx % p.n, c = gidx / p.n;
    if (c >= p.C) return;
    const uint64_t t0 = c * p.L, t1 = t0 + p.L < p.F ? t0 + p.L : p.F;
    const bool always = !p.g && p.mask != 0;
    if (p.kind == CPROC_CUDA_NODE_ACC) {
        uint32_t sum = 0;
        if (p.g || p.mask == 0) { for (uint64_t t = t0; t < t1; ++t) if (p.mask && gs_exec(p, i, t)) sum += p.x[i * p.xs_i + t * p.xs_t]; }
        else for (uint64_t t = t0; t < t1; ++t) sum += p.x[i * p.xs_i + t * p.xs_t];
        p.part[c * p.n + i] = make_uint4(sum, 0, 0, 0);
    } else {
        uint32_t cnt = 0, xl = 0, xp = 0;                 // executed ticks, last and previous executed input
        for (uint64_t t = t0; t < t1; ++t) {
            if (always || (p.mask && gs_exec(p, i, t))) { xp = xl; xl = p.x[i * p.xs_i + t * p.xs_t]; ++cnt; }
        }
        p.part[c * p.n + i] = cnt == 0 ? make_uint4(GS_IDENT, 0, 0, 0) : (cnt == 1 ? make_uint4(GS_ONE, xl, 0, 0) : make_uint4(GS_CONST, xl, xl != xp, 0));
    }
}

// later map applied after the earlier one
__device__ __forceinline__ uint4 gs_compose_edge(uint4 later, uint4 earlier) {
    if (later.x == GS_IDENT) return earlier;
    if (later.x == GS_CONST) return later;
    if (earlier.x == GS_IDENT) return later;                                  // ONE after IDENT: still depends on last_in
    return make_uint4(GS_CONST, later.y, later.y != earlier.y, 0);           // ONE after a map that fixed last = earlier.y
}
__device__ __forceinline__ uint2 gs_apply_edge(uint4 m, uint2 s) {             // s = {out, last}
    if (m.x == GS_IDENT) return s;
    if (m.x == GS_ONE) return make_uint2(m.y != s.y, m.y);
    return make_uint2(m.z, m.y);
}

#define GS_BLOCK 256
__global__ void __launch_bounds__(GS_BLOCK) k_gs_scan(const GsNode p) {
    __shared__ uint4 sm[GS_BLOCK];
    const uint64_t i = blockIdx.x;
    const uint32_t t = threadIdx.x;
    const bool is_acc = p.kind == CPROC_CUDA_NODE_ACC;
    const uint64_t K = (p.C + GS_BLOCK - 1) / GS_BLOCK;
    const uint64_t c0 = (uint64_t)t * K < p.C ? (uint64_t)t * K : p.C, c1 = c0 + K < p.C ? c0 + K : p.C;
    uint4 run = make_uint4(is_acc ? 0u : (uint32_t)GS_IDENT, 0, 0, 0);        // composition of this thread's run of chunks
    for (uint64_t c = c0; c < c1; ++c) {
        const uint4 m = p.part[c * p.n + i];
        if (is_acc) run.x += m.x; else run = gs_compose_edge(m, run);
    }
    sm[t] = run;
    __syncthreads();
    for (uint32_t d = 1; d < GS_BLOCK; d <<= 1) {                             // inclusive Hillis-Steele
        uint4 q = make_uint4(is_acc ? 0u : (uint32_t)GS_IDENT, 0, 0, 0);
        if (t >= d) q = sm[t - d];
        __syncthreads();
        if (t >= d) { if (is_acc) sm[t].x += q.x; else sm[t] = gs_compose_edge(sm[t], q); }
        __syncthreads();
    }
    uint2 s = make_uint2(p.st[i], is_acc ? 0u : p.st[p.npad + i]);            // state before the run
    if (t) { if (is_acc) s.x += sm[t - 1].x; else s = gs_apply_edge(sm[t - 1], s); }
    for (uint64_t c = c0; c < c1; ++c) {
        p.start[c * p.n + i] = s;
        const uint4 m = p.part[c * p.n + i];
        if (is_acc) s.x += m.x; else s = gs_apply_edge(m, s);
    }
    if (c1 == p.C && c0 < c1) { p.st[i] = s.x; if (!is_acc) p.st[p.npad + i] = s.y; }   // the run that ends the stream leaves the node state
}

__global__ void __launch_bounds__(128) k_gs_apply(const GsNode p) {
    const uint64_t gidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i = gidx % p.n, c = gidx / p.n;
    if (c >= p.C) return;
    const uint64_t t0 = c * p.L, t1 = t0 + p.L < p.F ? t0 + p.L : p.F;
    const uint2 s0 = p.start[c * p.n + i];
    uint32_t out = s0.x, last = s0.y;
    const bool always = !p.g && p.mask != 0;
    for (uint64_t t = t0; t < t1; ++t) {
        if (always || (p.mask && gs_exec(p, i, t))) {
            const uint32_t x = p.x[i * p.xs_i + t * p.xs_t];
            if (p.kind == CPROC_CUDA_NODE_ACC) out += x;                      // cproc.h:141
            else { out = (x != last); last = x; }                            // cproc.h:152-153
        }
        p.y[i * p.ys_i + t * p.ys_t] = out;
    }
}

int launch_graph_scan(cproc_cuda_batch *b, uint64_t F, const cproc_cuda_io *io) {
    cproc_cuda_ctx *ctx = b->ctx;
    const uint64_t n = b->n, nn = b->nodes.size();
    for (const cproc_cuda_node &nd : b->nodes)
        if (CPROC_CUDA_NODE_KIND(nd.type) > CPROC_CUDA_NODE_EDGE) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "graph scan: only acc / edge nodes are scannable (glide and pdm have a quantiser in the loop)");
    if (io->out == io->in) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "graph scan: in place not supported");
    // chunk length: enough (instance, chunk) threads to fill the chip, at least 64 ticks
    uint64_t L = ceil_div_u64(F, ceil_div_u64((uint64_t)ctx->n_sm * 2048 * 2, n));
    if (ctx->xvoice_chunk > 0) L = (uint64_t)ctx->xvoice_chunk; else if (L < 64) L = 64;
    const uint64_t C = ceil_div_u64(F, L);
    const size_t bval = sizeof(uint32_t) * nn * n * F, bpart = sizeof(uint4) * C * n, bstart = sizeof(uint2) * C * n;
    const size_t need = ((bval + 15) & ~(size_t)15) + ((bpart + 15) & ~(size_t)15) + bstart + 64;
    if (b->cap_scratch < need) {
        if (b->d_scratch) cudaFree(b->d_scratch);
        b->d_scratch = nullptr; b->cap_scratch = 0;
        CK(ctx, cudaMalloc(&b->d_scratch, need));
        b->cap_scratch = need;
    }
    uint32_t *val = (uint32_t *)b->d_scratch;                                  // [node][inst][F], planar
    uint8_t *w = (uint8_t *)b->d_scratch + ((bval + 15) & ~(size_t)15);
    uint4 *part = (uint4 *)w; w += (bpart + 15) & ~(size_t)15;
    uint2 *start = (uint2 *)w;
    const bool il = io->layout == CPROC_CUDA_INTERLEAVED;
    std::vector<uint32_t> off(nn);
    uint32_t o = 0;
    for (uint64_t k = 0; k < nn; ++k) { off[k] = o; o += cproc_node_words(b->nodes[k].type); }
    for (uint64_t k = 0; k < nn; ++k) {
        const cproc_cuda_node &nd = b->nodes[k];
        GsNode p;
        p.kind = CPROC_CUDA_NODE_KIND(nd.type); p.mask = nd.cond_mask;
        if (nd.src >= 0) { p.x = val + (uint64_t)nd.src * n * F; p.xs_i = F; p.xs_t = 1; }
        else {
            const uint64_t j = (uint64_t)(-(nd.src + 1));
            if (il) { p.x = (const uint32_t *)io->in + j * n; p.xs_i = 1; p.xs_t = (uint64_t)b->cfg.n_inputs * n; }      // [F][n_inputs][inst]
            else { p.x = (const uint32_t *)io->in + j * F; p.xs_i = (uint64_t)b->cfg.n_inputs * F; p.xs_t = 1; }          // [inst][n_inputs][F]
        }
        p.y = val + k * n * F; p.ys_i = F; p.ys_t = 1;          // every node's stream stays in scratch for later nodes
        p.g = (const uint32_t *)io->in2; p.gs_i = il ? 1 : F; p.gs_t = il ? n : 1;
        p.st = b->d_state + (uint64_t)off[k] * b->npad; p.npad = b->npad;
        p.n = n; p.F = F; p.L = L; p.C = C; p.part = part; p.start = start;
        const unsigned grid = (unsigned)ceil_div_u64(n * C, 128);
        k_gs_reduce<<<grid, 128, 0, ctx->stream>>>(p); CK_LAUNCH(ctx, "k_gs_reduce");
        k_gs_scan<<<(unsigned)n, GS_BLOCK, 0, ctx->stream>>>(p); CK_LAUNCH(ctx, "k_gs_scan");
        k_gs_apply<<<grid, 128, 0, ctx->stream>>>(p); CK_LAUNCH(ctx, "k_gs_apply");
        // output streams: [inst][n_outputs][F] / [F][n_outputs][inst]; a node may feed several outputs
        const uint64_t no = b->outs.size();
        for (uint64_t q = 0; q < no; ++q) {
            if (b->outs[q] != k) continue;
            GsNode c2 = p;
            if (il) { c2.y = (uint32_t *)io->out + q * n; c2.ys_i = 1; c2.ys_t = no * n; }
            else { c2.y = (uint32_t *)io->out + q * F; c2.ys_i = no * F; c2.ys_t = 1; }
            k_gs_apply<<<grid, 128, 0, ctx->stream>>>(c2); CK_LAUNCH(ctx, "k_gs_apply");
        }
    }
    return 0;
}
