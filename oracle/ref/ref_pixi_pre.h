/* ref_pixi_pre.h -- what stm32f103/pixi.c:279,282-285 (the demo LFO bank of dac_adc_update) need around them:
 * the two arrays of struct app (pixi.c:63-64,77-78).  oracle/build_ref.sh pipes the reference lines in between
 * this header and ref_pixi_tail.c; the SPI register traffic around them is hardware glue and stays out. */
#include <stdint.h>
#include <string.h>
#define NB_DAC 12
#define NB_ADC 6
struct app { uint16_t dac_vals[NB_DAC]; uint16_t adc_vals[NB_ADC]; };
static inline void ref_pixi_lfo_tick(struct app *app) {
