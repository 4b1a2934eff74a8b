/* Stand-in for uc_tools xorshift.h (mod_pdm_pwm.c:8-9, XORSHIFT_STATIC): the repository
 * (github:zwizwa/uc_tools) is not in the reference tree, so random_u32() is RESTATED here as
 * Marsaglia xorshift32 (13,17,5) -- "parity unpinned" for the generator itself; everything the
 * ISR does with its output is the reference's own code.  Test infrastructure. */
#ifndef STUB_XORSHIFT_H
#define STUB_XORSHIFT_H
#include <stdint.h>
static uint32_t ref_xorshift_state = 2463534242u;
static inline uint32_t random_u32(void) {
    uint32_t x = ref_xorshift_state;
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    return ref_xorshift_state = x;
}
#endif
