/* arm_v1_model.c -- an instruction-level model of the two ARM instructions the reference's v1 carry-bit PDM is written in
 * (stm32f103/mod_pdm.c:214-244: `adds %0, %0, %2` then `rrx %1, %1`), executed the way pdm_channels_update (:254-259) and
 * pdm_update (:267-275) sequence them: ONE shift register per tick, every channel of the MCU rotating its carry into the MSB
 * in channel order, then the shift that lines the word up with the GPIO pins.  TEST INFRASTRUCTURE: it exists to check the
 * add-with-carry restatement in cproc_oracle.c (orc_v1_channel_update: `carry = a < x`) against something that shares none of
 * its code -- here the carry is bit 32 of the 33-bit sum and the rotate is modelled on the CPSR C flag.  No ARM toolchain or
 * emulator is in this image, so the v1 path stays "restated, parity unpinned by the reference's own binary"; what this pins
 * is that two independent readings of the ARM ARM agree.
 *
 * ARMv7-M ARM: ADDS Rd, Rn, Rm: (result, carry, overflow) = AddWithCarry(R[n], R[m], '0'); APSR.C = carry (A7.7.4).
 *              RRX  Rd, Rm:     (result, carry_out) = RRX_C(R[m], APSR.C): result = C:R[m]<31:1>; without the S suffix the
 *                               flags are not updated (A7.7.116; the reference's rrx has no S).
 */
#include <stdint.h>

typedef struct { uint32_t r[16]; unsigned c; } arm_cpu;

static void arm_adds(arm_cpu *cpu, int rd, int rn, int rm) {
    const uint64_t wide = (uint64_t)cpu->r[rn] + (uint64_t)cpu->r[rm];      /* unsigned_sum of AddWithCarry, carry_in = 0 */
    cpu->r[rd] = (uint32_t)wide;
    cpu->c = (unsigned)(wide >> 32) & 1u;                                   /* carry_out = (UInt(result) != unsigned_sum) */
}
static void arm_rrx(arm_cpu *cpu, int rd, int rm) {
    cpu->r[rd] = ((uint32_t)cpu->c << 31) | (cpu->r[rm] >> 1);              /* C flag unchanged */
}

static uint32_t model_xorshift32(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }

/* One MCU: n_ch channels {setpoint, accu} (mod_pdm.c:196-199), F timer interrupts.  gpio[t] = `set_bits` of pdm_update (:270)
 * for pin_chan0: channel c drives bit pin_chan0 + c.  rng: the dither generator state (in/out); dither = rnd & dmask (:256). */
void arm_v1_mcu_run(uint32_t *chan, uint32_t n_ch, uint32_t pin_chan0, uint32_t *rng, uint32_t dmask, uint64_t F, uint32_t *gpio) {
    arm_cpu cpu = {{0}, 0};
    for (uint64_t t = 0; t < F; t++) {
        cpu.r[1] = 0;                                                       /* uint32_t shiftreg = 0 */
        const uint32_t dither = model_xorshift32(rng) & dmask;
        for (uint32_t c = 0; c < n_ch; c++) {                               /* PDM_FOR_CHANNELS(PDM_CHANNEL_UPDATE) */
            cpu.r[0] = chan[2 * c + 1];                                     /* %0 = channel->accu */
            cpu.r[2] = chan[2 * c] + dither;                                /* %2 = dither_setpoint (:230) */
            arm_adds(&cpu, 0, 0, 2);
            arm_rrx(&cpu, 1, 1);
            chan[2 * c + 1] = cpu.r[0];
        }
        gpio[t] = cpu.r[1] >> (32 - n_ch - pin_chan0);                      /* :270 */
    }
}
