/* fakepd.c -- scripted Pd stand-in (test infrastructure): loads the external's setup routine, creates
 * GRAINS objects with different thresholds, builds the DSP chain by calling each object's "dsp"
 * method, then runs TICKS DSP ticks of BLOCK samples -- objects 0 and 1 in place (inlet vector ==
 * outlet vector, as Pd does) -- with inputs from xorshift32, sends a "threshold" message half way,
 * and prints every outlet block as hex floats:   out <tick> <grain> <f0> ...   A second scene switches one object off for three
 * ticks (out2 lines). */
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "m_pd.h"

#define GRAINS 5
#define BLOCK 64
#define TICKS 6
struct fake_class { t_newmethod newm; size_t size; t_method dsp, threshold; };
static t_class the_class;
static struct { t_perfroutine f; t_int w[8]; } chain[64];
static int chain_len;

t_symbol *gensym(const char *s) { t_symbol *y = malloc(sizeof(*y)); y->s_name = strdup(s); return y; }
t_class *class_new(t_symbol *name, t_newmethod newmethod, t_method freemethod, size_t size, int flags, t_atomtype arg1, ...) {
    (void)name; (void)freemethod; (void)flags; (void)arg1;
    the_class.newm = newmethod; the_class.size = size;
    return &the_class;
}
void class_addmethod(t_class *c, t_method fn, t_symbol *sel, t_atomtype arg1, ...) {
    (void)arg1;
    if (!strcmp(sel->s_name, "dsp")) c->dsp = fn;
    if (!strcmp(sel->s_name, "threshold")) c->threshold = fn;
}
void class_domainsignalin(t_class *c, int onset) { (void)c; (void)onset; }
t_pd *pd_new(t_class *cls) { t_object *o = calloc(1, cls->size); o->ob_pd = cls; return (t_pd *)o; }
t_inlet *inlet_new(t_object *owner, t_pd *dest, t_symbol *s1, t_symbol *s2) { (void)owner; (void)dest; (void)s1; (void)s2; return NULL; }
t_outlet *outlet_new(t_object *owner, t_symbol *s) { (void)owner; (void)s; return NULL; }
void dsp_add(t_perfroutine f, int n, ...) {
    va_list ap; va_start(ap, n);
    chain[chain_len].f = f;
    chain[chain_len].w[0] = (t_int)f;
    for (int k = 0; k < n; k++) chain[chain_len].w[1 + k] = va_arg(ap, t_int);
    va_end(ap);
    chain_len++;
}
void post(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fprintf(stderr, "post: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap); }

void square_grain_b200_tilde_setup(void);

int main(void) {
    square_grain_b200_tilde_setup();
    void *obj[GRAINS];
    static float vin[GRAINS][BLOCK], vout[GRAINS][BLOCK];
    for (int g = 0; g < GRAINS; g++) {
        obj[g] = ((void *(*)(t_floatarg))the_class.newm)(0.05f + 0.1f * (float)g);
        t_signal in = {BLOCK, vin[g]}, out = {BLOCK, g < 2 ? vin[g] : vout[g]};     /* grains 0, 1: in place */
        t_signal *sp[2] = {&in, &out};
        /* dsp_add reads the values, so the stack signals may go out of scope */
        ((void (*)(void *, t_signal **))the_class.dsp)(obj[g], sp);
    }
    uint32_t s = 2463534242u;
    for (int tick = 0; tick < TICKS; tick++) {
        if (tick == 3) ((void (*)(void *, t_floatarg))the_class.threshold)(obj[2], -0.4f);    /* fabs() -> 0.4 */
        for (int g = 0; g < GRAINS; g++)
            for (int t = 0; t < BLOCK; t++) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; vin[g][t] = (float)(int32_t)s * (1.0f / 2147483648.0f); }
        for (int k = 0; k < chain_len; k++) chain[k].f(chain[k].w);
        for (int g = 0; g < GRAINS; g++) {
            const float *o = g < 2 ? vin[g] : vout[g];
            printf("out %d %d", tick, g);
            for (int t = 0; t < BLOCK; t++) printf(" %a", o[t]);
            printf("\n");
        }
    }
    /* second scene: object 1 sits in a switched-off subpatch for ticks TICKS .. TICKS+2 (its perform routine is not called), then
       comes back; the others carry on.  Printed as out2 <tick> <grain>. */
    for (int tick = TICKS; tick < TICKS + 5; tick++) {
        const int off = tick < TICKS + 3;
        for (int g = 0; g < GRAINS; g++)
            for (int t = 0; t < BLOCK; t++) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; vin[g][t] = (float)(int32_t)s * (1.0f / 2147483648.0f); }
        for (int k = 0; k < chain_len; k++) if (!(off && k == 1)) chain[k].f(chain[k].w);
        for (int g = 0; g < GRAINS; g++) {
            if (off && g == 1) continue;
            const float *o = g < 2 ? vin[g] : vout[g];
            printf("out2 %d %d", tick, g);
            for (int t = 0; t < BLOCK; t++) printf(" %a", o[t]);
            printf("\n");
        }
    }
    return 0;
}
