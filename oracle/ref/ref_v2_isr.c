/* ref_v2_isr.c -- the reference's PDM v2 firmware, UNMODIFIED, on the host: stm32f103/mod_pdm_pwm.c
 * (struct channel, pdm_update_glide, PDM_COPY_LINE, PDM_UPDATE_CHANNEL, the TIM3 ISR :80-141) and
 * stm32f103/mod_controlrate.c (pdm_update_line, control_update, the EXTI1 software interrupt :28-61)
 * are #included as they lie in the reference tree, compiled against oracle/shim/stm32 (a stand-in
 * for the hardware layer) with the configuration mod_synth.c:29-31 gives them.  The harness loads
 * the globals, calls the timer ISR once per tick and collects the three duty values it hands to
 * hw_multi_pwm_duty().  This pins the oracle's v2 channel arithmetic (glide, line copy, control
 * divider, line update, dither mask, out_shift) to the reference's own statements; only
 * random_u32() (uc_tools, absent) stays restated.  Test infrastructure; built only into oracle/_ref. */
#include <stdint.h>
#include <string.h>
#include "hw_stub.h"
uint32_t ref_v2_duty[4];
#define PDM_DIV_LOG 8                        /* mod_synth.c:29-31 */
#define PDM_DIV (1 << PDM_DIV_LOG)
#define PWM_HZ (72000000 / PDM_DIV)
#include "mod_pdm_pwm.c"                     /* <reference>/stm32f103 via -I */
#include "mod_controlrate.c"

uint32_t ref_v2_isr_nb_channels(void) { return PDM_NB_CHANNELS; }
uint32_t ref_v2_isr_sizeof_channel(void) { return sizeof(struct channel); }
uint32_t ref_v2_isr_control_div_log(void) { return CONTROL_DIV_LOG; }

/* One MCU: PDM_NB_CHANNELS channels sharing one dither word per tick.
 *   chan       [nb][7] words = struct channel {setpoint; line[2]{position, velocity}; pdm2{s1, s2}}, in/out
 *   prng       xorshift state, in/out;  count: control_div_count, in/out
 *   setpoints  [n_rows][nb] or NULL: row k is written to channel.setpoint just before the k-th control
 *              boundary met during the run (the firmware's main loop writes setpoints asynchronously; the
 *              ISR reads them in control_update, triggered at the boundary)
 *   duty       [nb][F] bytes: what the ISR hands to hw_multi_pwm_duty() at every tick */
void ref_v2_isr_run(uint32_t *chan, uint32_t *prng, uint32_t *count, const uint32_t *setpoints, uint64_t n_rows, uint64_t F, uint8_t *duty) {
    const uint32_t nb = PDM_NB_CHANNELS;
    memcpy(pdm_channel, chan, sizeof(pdm_channel));
    ref_xorshift_state = *prng;
    control_div_count = *count;
    uint64_t row = 0;
    for (uint64_t t = 0; t < F; t++) {
        if (control_div_count == 0 && setpoints && row < n_rows) {
            for (uint32_t i = 0; i < nb; i++) pdm_channel[i].setpoint = pdm_safe_setpoint(setpoints[row * nb + i]);
            row++;
        }
        HW_TIM_ISR(TIM_PDM)();
        for (uint32_t i = 0; i < nb; i++) duty[(uint64_t)i * F + t] = (uint8_t)ref_v2_duty[i];
    }
    memcpy(chan, pdm_channel, sizeof(pdm_channel));
    *prng = ref_xorshift_state;
    *count = control_div_count;
}
