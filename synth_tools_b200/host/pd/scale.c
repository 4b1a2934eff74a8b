/* scale.c -- the [scale exp|lin min max] control object of linux/synth_tools.c:147-194: maps a MIDI
 * value 0..127 to a parameter range, exponentially (base * (max/min)^frac) or linearly
 * (base + (max-min) * frac).  SURVEY 8 f-4 lists it with the host adapters: it is what feeds
 * thresholds and increments to the batched objects from controller data.  Pure host code (one powf
 * per message): nothing here belongs on the GPU.  Same float operations in the same order as the
 * reference, so the outlet values are bit-identical with the same libm. */
#include <math.h>
#include <string.h>
#include "m_pd.h"

static t_class *scale_class;
struct scale {
    t_object x_obj;
    t_outlet *out;
    t_float base, diff;
    int exponential;
};

/* the two maps, usable without Pd (synth_tools.c:155-164) */
t_float cproc_scale_exp(t_float base, t_float diff, t_float mfrac) {
    t_float frac = mfrac * (1.0f / 127.0f);
    return base * powf(diff, frac);
}
t_float cproc_scale_lin(t_float base, t_float diff, t_float mfrac) {
    t_float frac = mfrac * (1.0f / 127.0f);
    return base + diff * frac;
}

static void scale_float(struct scale *x, t_floatarg mfrac) {              /* :165-167 */
    outlet_float(x->out, x->exponential ? cproc_scale_exp(x->base, x->diff, mfrac) : cproc_scale_lin(x->base, x->diff, mfrac));
}

static void *scale_new(t_symbol *type, t_floatarg min, t_floatarg max) {  /* :168-188 */
    int exponential;
    if (!strcmp(type->s_name, "exp")) exponential = 1;
    else if (!strcmp(type->s_name, "lin")) exponential = 0;
    else return NULL;
    struct scale *x = (struct scale *)pd_new(scale_class);
    x->base = min;
    x->diff = exponential ? max / min : max - min;
    x->exponential = exponential;
    x->out = outlet_new(&x->x_obj, gensym("float"));
    return x;
}

void scale_setup(void) {                                                  /* :191-194 */
    scale_class = class_new(gensym("scale"), (t_newmethod)scale_new, 0, sizeof(struct scale), CLASS_DEFAULT, A_DEFSYMBOL, A_DEFFLOAT, A_DEFFLOAT, 0);
    class_addfloat(scale_class, (t_method)scale_float);
}
