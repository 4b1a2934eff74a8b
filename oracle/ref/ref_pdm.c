/* ref_pdm.c -- batched harness around the UNMODIFIED reference
 * stm32f103/pdm.h.  The v2 ISR body (mod_pdm_pwm.c:101-141) and the control
 * rate line update (mod_controlrate.c:28-40) depend on hw_stm32f103.h and are
 * restated here around the REAL pdmK_update calls.  Test infrastructure and
 * CPU baseline ("reference" kind); built only into oracle/_ref. */
#include <stdint.h>
#include "pdm.h"        /* <reference>/stm32f103/pdm.h via -I */

static inline uint32_t ref_pdm_step(uint32_t *s, uint32_t order, uint32_t in, uint32_t sh, uint32_t d) {
    switch (order) {
    case 1: return pdm1_update((struct pdm1 *)s, in, sh);
    case 2: return pdm2_update((struct pdm2 *)s, in, sh, d);
    case 3: return pdm3_update((struct pdm3 *)s, in, sh, d);
    default: return pdm4_update((struct pdm4 *)s, in, sh, d);
    }
}
uint32_t ref_pdm_sizeof(uint32_t order) {
    switch (order) { case 1: return sizeof(struct pdm1); case 2: return sizeof(struct pdm2);
                     case 3: return sizeof(struct pdm3); default: return sizeof(struct pdm4); }
}
void ref_pdm_run(uint32_t order, uint32_t *state, uint64_t N, uint64_t F,
                 const uint32_t *in, const uint32_t *in_const,
                 uint32_t out_shift, const uint32_t *dither, uint32_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++)
        for (uint64_t t = 0; t < F; t++)
            out[n * F + t] = ref_pdm_step(state + n * order, order,
                                          in ? in[n * F + t] : in_const[n], out_shift,
                                          dither ? dither[t] : 0);
}

/* xorshift32 (13,17,5): stands in for uc_tools random_u32 (parity unpinned). */
static inline uint32_t ref_rng(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }

/* Channel words: [setpoint, l0.pos, l0.vel, l1.pos, l1.vel, s1..sK]
 * (struct channel, mod_pdm_pwm.c:89-93).  Specialised per order so the
 * reference's always_inline pdmK_update is inlined in the tick loop, as in
 * the firmware ISR. */
#define REF_V2_BODY(STEP)                                                                   \
    for (uint64_t t = 0; t < F; t++) {                                                      \
        uint32_t d = (dither_ext ? dither_ext[b * F + t] : ref_rng(&rng)) & dither_mask;    \
        if (cnt == 0) {                                                                     \
            for (uint64_t c = c0; c < c1; c++) {                                            \
                uint32_t *x = chan + c * W;                                                 \
                if (setpoints) x[0] = setpoints[row * N + c];                               \
                x[1] = x[3]; x[2] = x[4];                                                   \
                x[3] += x[4] << ctl_div_log;                                                \
                int32_t span = x[0] - x[3];                                                 \
                x[4] = span >> ctl_div_log;                                                 \
            }                                                                               \
            row++;                                                                          \
        }                                                                                   \
        for (uint64_t c = c0; c < c1; c++) {                                                \
            uint32_t *x = chan + c * W;                                                     \
            x[1] += x[2];                                                                   \
            duty[c * F + t] = (uint8_t)STEP;                                                \
        }                                                                                   \
        cnt = (cnt + 1) % div;                                                              \
    }

void ref_pdm_v2_run(uint32_t *chan, uint32_t order, uint64_t N,
                    uint32_t bank_size, uint32_t *prng,
                    const uint32_t *dither_ext, uint32_t dither_mask,
                    uint32_t *count, uint32_t ctl_div_log, uint32_t out_shift,
                    const uint32_t *setpoints, uint64_t F, uint8_t *duty) {
    uint64_t n_banks = (N + bank_size - 1) / bank_size;
    uint32_t W = 5 + order, div = 1u << ctl_div_log, count0 = *count;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)n_banks; b++) {
        uint64_t c0 = (uint64_t)b * bank_size, c1 = c0 + bank_size < N ? c0 + bank_size : N;
        uint32_t rng = prng ? prng[b] : 0, cnt = count0;
        uint64_t row = 0;
        switch (order) {
        case 1: REF_V2_BODY(pdm1_update((struct pdm1 *)(x + 5), x[1], out_shift)) break;
        case 2: REF_V2_BODY(pdm2_update((struct pdm2 *)(x + 5), x[1], out_shift, d)) break;
        case 3: REF_V2_BODY(pdm3_update((struct pdm3 *)(x + 5), x[1], out_shift, d)) break;
        default: REF_V2_BODY(pdm4_update((struct pdm4 *)(x + 5), x[1], out_shift, d)) break;
        }
        if (prng) prng[b] = rng;
    }
    *count = (uint32_t)((count0 + F) % div);
}
