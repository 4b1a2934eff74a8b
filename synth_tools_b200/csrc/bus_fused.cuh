// bus_fused.cuh -- the mix-bus exchange as the tail of the render kernel itself.
//
// The only exchange step of the path is the shared mix bus: a few KB per frame block.  As separate launches
// (render, then k_bus_allreduce or ncclAllReduce + a conversion kernel) the exchange costs more than it moves:
// at 8 GPUs a 512 Ki-voice shard renders in ~28 us and every extra launch, event or NCCL call is 3..30 us.
// Here the render launch does it, in one of two ways:
//   mode 1: the block that completes a piece of the local mix (the last tile to arrive, by an atomic ticket) pushes that
//           piece straight into slot[rank] of every peer's bus buffer with stores through the NVLink peer mapping; the
//           last pusher of the launch publishes one flag per peer (release, system scope), waits for the peers' flags,
//           adds the slots in rank order and scales to float once (synth.c:180,194).  The bus is valid when the launch
//           ends; the exchange (~10 us of NVLink round trips) is exposed at its tail.
//   mode 2: pipelined.  The render only stages its mix in local memory.  ONE extra block of the NEXT launch, which runs
//           from that launch's start beside the render of block k+1 (the grid leaves it a free slot), pushes block k,
//           publishes, waits, reduces: the whole exchange is off the critical path, results trail by one block (the last
//           one is completed by cproc_cuda_bus_flush).
// Slots and flags are indexed by epoch & 3.  A rank can only push epoch e+2 after its launch that exchanged e+1 has
// ended, i.e. after every peer published e+1, which a peer does after the launch in which it read the slots of e.
#pragma once
#include "common.cuh"

#define BUS_MAX_WORLD 16
#define BUS_NPAR 4

struct BusFused {
    int world, rank;               // world == 0: no bus attached, the render writes its own mix
    uint32_t mode;                 // 1 exchange in this launch, 2 pipelined
    uint32_t epoch;                // mode 1: epoch of the mix this launch renders
    uint64_t cap;                  // words per slot
    int32_t *slots[BUS_MAX_WORLD]; // peer-mapped bases: [BUS_NPAR][world][cap]
    uint32_t *flags[BUS_MAX_WORLD];// [BUS_NPAR][world]
    uint32_t *status;              // local: epoch at which a peer failed to arrive (0 = fine)
    uint32_t *ticket;              // local, mode 1: pushers of this launch that are done (reset by the last)
    uint32_t participants;         // mode 1: blocks that push
    uint32_t *stage;               // local, mode 2: where this launch leaves its mix for the next launch's exchange block
    // the exchange this launch completes: mode 1 its own epoch, mode 2 the previous block's (fin_epoch == 0: none)
    uint32_t fin_epoch, fin_op, fin_scale;   // op 0 wrap-around sum, 1 OR, 2 float sum in rank order; scale 0 none, 1 saw, 2 square, 3 grain
    uint64_t fin_count;
    const uint32_t *fin_stage;     // mode 2: the staged mix of that block
    int32_t *fin_imix;             // reduced integer (or float-bit) words, may be null
    float *fin_out;                // scaled float bus, may be null
};

__device__ __forceinline__ void bus_st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t bus_ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// One finished word of this rank's mix.  mode 1: into slot[rank] of every rank (own copy included); mode 2: staged locally.
__device__ __forceinline__ void bus_emit_word(const BusFused &bf, uint64_t i, uint32_t v) {
    if (bf.mode == 2) { bf.stage[i] = v; return; }
    const uint64_t off = ((uint64_t)(bf.epoch & (BUS_NPAR - 1)) * bf.world + bf.rank) * bf.cap + i;
    for (int q = 0; q < bf.world; ++q) bf.slots[q][off] = (int32_t)v;
}

// Wait for every rank's flag of `epoch`, add the slots in rank order, write the bus.  One whole block.
__device__ __forceinline__ void bus_reduce_block(const BusFused &bf, uint32_t epoch, uint64_t count, uint32_t op, uint32_t scale, int32_t *imix, float *out) {
    __shared__ uint32_t bus_failed;
    const uint32_t par = epoch & (BUS_NPAR - 1);
    if (threadIdx.x == 0) bus_failed = 0;
    __syncthreads();
    if ((int)threadIdx.x < bf.world) {
        const uint32_t *f = bf.flags[bf.rank] + par * bf.world + threadIdx.x;
        const long long t0 = clock64();
        while (bus_ld_acquire_sys(f) != epoch) {
            if (clock64() - t0 > 20000000000ll) { bus_failed = 1; break; }       // ~10 s: a rank never arrived
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (bus_failed) { if (threadIdx.x == 0) *bf.status = epoch; return; }
    const int32_t *mine = bf.slots[bf.rank] + (uint64_t)par * bf.world * bf.cap;
    for (uint64_t i = threadIdx.x; i < count; i += blockDim.x) {
        uint32_t acc = 0;
        if (op == 2) {                                     // float bus (extension voices): fixed rank order, one rounding per add
            float fa = 0.0f;
            for (int q = 0; q < bf.world; ++q) fa = __fadd_rn(fa, __int_as_float(__ldcg(mine + (uint64_t)q * bf.cap + i)));
            acc = __float_as_uint(fa);
        } else {
            for (int q = 0; q < bf.world; ++q) {
                const uint32_t v = (uint32_t)__ldcg(mine + (uint64_t)q * bf.cap + i);
                acc = op ? (acc | v) : (acc + v);
            }
        }
        if (imix) imix[i] = (int32_t)acc;
        if (out) {
            float f;
            if (scale == 2) f = __uint2float_rn(acc) * 0x1p-32f;                 // synth.c:194
            else if (scale == 3) f = __int2float_rn((int32_t)acc) * 0x1p-7f;     // grain mix in units of 2^-7
            else if (scale == 1) f = __int2float_rn((int32_t)acc) * 0x1p-32f;    // synth.c:180
            else f = __uint_as_float(acc);
            out[i] = f;
        }
    }
}

// mode 1: called by every pushing block (all its threads) once its pushes are issued.  The last of the launch publishes
// the flags and performs the reduce.
__device__ __forceinline__ void bus_participant_done(const BusFused &bf) {
    __shared__ uint32_t bus_last;
    if (bf.mode != 1) return;
    __threadfence_system();                                // this block's pushes before its ticket
    __syncthreads();
    if (threadIdx.x == 0) bus_last = atomicAdd(bf.ticket, 1u) == bf.participants - 1u;
    __syncthreads();
    if (!bus_last) return;
    if (threadIdx.x == 0) *bf.ticket = 0;                  // ready for the next launch (stream order)
    __threadfence_system();
    if ((int)threadIdx.x < bf.world) bus_st_release_sys(bf.flags[threadIdx.x] + (bf.epoch & (BUS_NPAR - 1)) * bf.world + bf.rank, bf.epoch);
    bus_reduce_block(bf, bf.fin_epoch, bf.fin_count, bf.fin_op, bf.fin_scale, bf.fin_imix, bf.fin_out);
}

// mode 2: the extra block of a launch (or cproc_cuda_bus_flush): the whole exchange of the previous frame block.
__device__ __forceinline__ void bus_exchange_block(const BusFused &bf) {
    if (!bf.fin_epoch) return;
    const uint32_t par = bf.fin_epoch & (BUS_NPAR - 1);
    const uint64_t off = ((uint64_t)par * bf.world + bf.rank) * bf.cap;
    for (uint64_t i = threadIdx.x; i < bf.fin_count; i += blockDim.x) {
        const uint32_t v = __ldcg(bf.fin_stage + i);
        for (int q = 0; q < bf.world; ++q) bf.slots[q][off + i] = (int32_t)v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < bf.world) bus_st_release_sys(bf.flags[threadIdx.x] + par * bf.world + bf.rank, bf.fin_epoch);
    bus_reduce_block(bf, bf.fin_epoch, bf.fin_count, bf.fin_op, bf.fin_scale, bf.fin_imix, bf.fin_out);
}

// host side (bus.cu): fills `bf` for a launch that renders `count` words for the batch's attached bus (world = 0 when none);
// n_pushers = blocks that will call bus_participant_done after emitting.  Returns nonzero on error.
int cproc_bus_fused_begin(cproc_cuda_batch *b, BusFused *bf, uint64_t count, uint32_t op, uint32_t scale, uint32_t n_pushers,
                          int32_t *imix_out, float *float_out);
