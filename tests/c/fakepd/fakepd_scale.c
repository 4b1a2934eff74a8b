/* fakepd_scale.c -- scripted Pd stand-in for synth_tools_b200/host/pd/scale.c (test infrastructure):
 * creates [scale exp 20 20000], [scale lin -1 3] and a [scale foo ..] that must be refused, sends the
 * MIDI values 0..127 to both and prints every outlet value as a hex float:  out <obj> <value> <f> */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "m_pd.h"

struct fake_class { t_newmethod newm; size_t size; t_method flt; };
struct fake_outlet { int id; };
static t_class the_class;
static int n_outlets;
static t_float last_value;
static int last_outlet = -1;

t_symbol *gensym(const char *s) { t_symbol *y = malloc(sizeof(*y)); y->s_name = strdup(s); return y; }
t_class *class_new(t_symbol *name, t_newmethod newmethod, t_method freemethod, size_t size, int flags, t_atomtype arg1, ...) {
    (void)name; (void)freemethod; (void)flags; (void)arg1;
    the_class.newm = newmethod; the_class.size = size;
    return &the_class;
}
void class_addfloat(t_class *c, t_method fn) { c->flt = fn; }
t_pd *pd_new(t_class *cls) { t_object *o = calloc(1, cls->size); o->ob_pd = cls; return (t_pd *)o; }
t_outlet *outlet_new(t_object *owner, t_symbol *s) { (void)owner; (void)s; t_outlet *o = malloc(sizeof(*o)); o->id = n_outlets++; return o; }
void outlet_float(t_outlet *x, t_float f) { last_outlet = x->id; last_value = f; }

void scale_setup(void);

int main(void) {
    scale_setup();
    void *(*mk)(t_symbol *, t_floatarg, t_floatarg) = (void *(*)(t_symbol *, t_floatarg, t_floatarg))the_class.newm;
    void *e = mk(gensym("exp"), 20.0f, 20000.0f), *l = mk(gensym("lin"), -1.0f, 3.0f);
    printf("refused %d\n", mk(gensym("foo"), 1.0f, 2.0f) == NULL);
    void *obj[2] = {e, l};
    for (int k = 0; k < 2; k++)
        for (int v = 0; v < 128; v++) {
            ((void (*)(void *, t_floatarg))the_class.flt)(obj[k], (t_float)v);
            printf("out %d %d %a %d\n", k, v, last_value, last_outlet);
        }
    return 0;
}
