/* ref_pwm_tail.c -- appended (same translation unit) to lines 159-175 of the reference
 * stm32f103/mod_pdm.c (pwm_phase, pwm_speed, PHASE_MASK, pwm_update), which oracle/build_ref.sh pipes
 * in after <stdint.h>.  The rest of that file is ARM inline assembly and hardware glue. */
void ref_pwm_run(uint32_t *phase, const uint32_t *speed, uint64_t N, uint64_t F, uint8_t *duty) {
    for (uint64_t n = 0; n < N; n++) {
        pwm_phase = phase[n]; pwm_speed = speed[n];
        for (uint64_t t = 0; t < F; t++) duty[n * F + t] = (uint8_t)pwm_update();
        phase[n] = pwm_phase;
    }
}
uint32_t ref_pwm_defaults(int what) { return what ? pwm_speed : pwm_phase; }
