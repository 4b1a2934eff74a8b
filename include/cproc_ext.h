/* cproc_ext.h -- extension processors in DEF_PROC form (SURVEY 8 a-X): what the reference's
 * processor library (generic/cproc.h:128-154 has acc and edge) lacks for its own stated
 * configurations -- an oscillator -> filter chain and a polyphonic voice (osc + SVF + envelope).
 *
 * Include AFTER the reference's generic/cproc.h: the definitions below use its DEF_PROC
 * (cproc.h:89-95), DEF_PROC_STRUCTS (:99-103) and `w` (:127) unchanged, so a graph text that
 * names these processors in PROC / PROC_COND statements compiles as C for a CPU host exactly like
 * one that names acc / edge -- and renders on the GPU through CPROC_CUDA_GRAPH (cproc_cuda.h: node
 * kinds CPROC_CUDA_NODE_PHASOR_F .. _ASFLOAT), bit for bit the same when this header is compiled
 * with -ffp-contract=off (each statement below is ONE IEEE rounding; fmaf is a fused multiply-add).
 *
 * "All atomic values are machine words ... is a property of the individual processors, and is not
 * essential for the composition mechanism" (cproc.h:123-126): these processors carry binary32
 * floats in their 32-bit fields.
 */
#ifndef CPROC_EXT_H
#define CPROC_EXT_H
#include <math.h>
#include <stdint.h>
#include <string.h>

/* Phase accumulator read as a signed saw in [-1, 1): acc (cproc.h:140-142) with the float read-out
 * of linux/synth.c:175-177 (read, then advance).  .mod is added to the increment (FM / sync input). */
#define for_phasor_f_state(m)  m(float,out) m(w,phase)
#define for_phasor_f_input(m)  m(w,mod)
#define for_phasor_f_config(m)
#define for_phasor_f_param(m)  m(w,inc)
DEF_PROC(phasor_f, s, c, p, i) {
    s->out = (float)(int32_t)s->phase * (1.0f / 2147483648.0f);
    s->phase += p->inc + i->mod;
}

/* Chamberlin state-variable filter, low-pass output. */
#define for_svf_state(m)  m(float,out) m(float,bp)
#define for_svf_input(m)  m(float,in)
#define for_svf_config(m)
#define for_svf_param(m)  m(float,f) m(float,q)
DEF_PROC(svf, s, c, p, i) {
    float lp = fmaf(p->f, s->bp, s->out);
    float hp = i->in - lp;
    hp = fmaf(-p->q, s->bp, hp);
    s->bp = fmaf(p->f, hp, s->bp);
    s->out = lp;
}

/* Linear attack / release envelope applied to the input: attack for the first gate_frames ticks. */
#define for_env_state(m)  m(float,out) m(float,env) m(w,t)
#define for_env_input(m)  m(float,in)
#define for_env_config(m)
#define for_env_param(m)  m(float,attack) m(float,release) m(w,gate_frames)
DEF_PROC(env, s, c, p, i) {
    float e = s->env;
    if (s->t < p->gate_frames) { e = e + p->attack; if (e > 1.0f) e = 1.0f; }
    else { e = e - p->release; if (e < 0.0f) e = 0.0f; }
    s->env = e;
    s->t += 1;
    s->out = i->in * e;
}

/* One-pole low-pass. */
#define for_onepole_state(m)  m(float,out)
#define for_onepole_input(m)  m(float,in)
#define for_onepole_config(m)
#define for_onepole_param(m)  m(float,a)
DEF_PROC(onepole, s, c, p, i) {
    s->out = fmaf(p->a, i->in - s->out, s->out);
}

/* Gain (two of them pan a voice). */
#define for_gain_state(m)  m(float,out)
#define for_gain_input(m)  m(float,in)
#define for_gain_config(m)
#define for_gain_param(m)  m(float,g)
DEF_PROC(gain, s, c, p, i) {
    s->out = p->g * i->in;
}

/* The float whose bits arrive in a `w` (cproc_update's inputs are `w *input`, test_cproc.c:13):
 * how an external float stream enters a graph. */
#define for_asfloat_state(m)  m(float,out)
#define for_asfloat_input(m)  m(w,in)
#define for_asfloat_config(m)
#define for_asfloat_param(m)
DEF_PROC(asfloat, s, c, p, i) {
    uint32_t bits = i->in;
    memcpy(&s->out, &bits, sizeof(bits));
}

/* Control rate -> audio rate interpolation of a float parameter (doc/combinators.org:28-34; the float counterpart of the
 * firmware's struct line / pdm_update_line, mod_controlrate.c:28-40): .in is read once per 2^div_log ticks, .out ramps towards
 * it.  The scale by a power of two is exact, so a tick is one rounding (the add) and a boundary two (the subtract, the add). */
#define for_glide_f_state(m)  m(float,out) m(float,step) m(w,count)
#define for_glide_f_input(m)  m(float,in)
#define for_glide_f_config(m) m(w,div_log)
#define for_glide_f_param(m)
DEF_PROC(glide_f, s, c, p, i) {
    if (s->count == 0) s->step = (i->in - s->out) * (1.0f / (float)(1u << c->div_log));
    s->out = s->out + s->step;
    s->count = (s->count + 1) & ((1u << c->div_log) - 1u);
}

/* Two audio-rate signals multiplied: glide_f -> mul applies a control-rate gain amount to an audio-rate signal. */
#define for_mul_state(m)  m(float,out)
#define for_mul_input(m)  m(float,in) m(float,gain)
#define for_mul_config(m)
#define for_mul_param(m)
DEF_PROC(mul, s, c, p, i) {
    s->out = i->in * i->gain;
}

/* Float results leave through their own upcall: cproc_output(uint32_t index, w value)
 * (test_cproc.c:5-7, mod_cproc_plugin.c:40-43) would convert a float by value. */
void cproc_output_f(uint32_t index, float value);

#endif /* CPROC_EXT_H */
