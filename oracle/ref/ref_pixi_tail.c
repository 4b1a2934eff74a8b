}
/* ref_pixi_tail.c -- closes ref_pixi_lfo_tick (the reference lines above) and runs it for `ticks` timer
 * interrupts: dac[NB_DAC] in/out, trace [ticks][NB_DAC] = the values written to the DAC after each tick. */
void ref_pixi_lfo_run(uint16_t *dac, uint16_t adc0, uint64_t ticks, uint16_t *trace) {
    struct app a;
    memset(&a, 0, sizeof(a));
    memcpy(a.dac_vals, dac, sizeof(a.dac_vals));
    a.adc_vals[0] = adc0;
    for (uint64_t t = 0; t < ticks; t++) {
        ref_pixi_lfo_tick(&a);
        memcpy(trace + t * NB_DAC, a.dac_vals, sizeof(a.dac_vals));
    }
    memcpy(dac, a.dac_vals, sizeof(a.dac_vals));
}
