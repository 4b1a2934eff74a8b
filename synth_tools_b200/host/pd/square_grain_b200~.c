/* square_grain_b200~.c -- a Pd external with the interface of square_grain~ (linux/synth_tools.c:78-150:
 * one signal inlet, one signal outlet, a float inlet / "threshold" and "brightness" messages) whose
 * Schmitt-trigger loop runs on the B200: SURVEY 8 f-4.
 *
 * One object is one grain.  Pd calls the perform routines of a patch one after the other inside a
 * DSP tick, so a per-object GPU call would cost a launch per grain per 64 samples.  Instead all
 * objects of the library share ONE batch: a perform routine copies its inlet block into its row of
 * the shared input buffer and hands out its row of the previous tick's result; the first perform of
 * a tick renders the rows collected during the previous tick in one cproc_cuda_run (N grains x n
 * frames).  Cost: one block of latency; state and threshold changes stay exact (the batch is the
 * reference loop, bit for bit, delayed by one block).  Objects in switched-off subpatches and objects
 * running at another block size do not disturb the others (see struct sg_pipe).
 *
 * Build (with Pd headers):  gcc -std=gnu99 -O2 -fPIC -shared square_grain_b200~.c -I<repo>/include \
 *     -L<repo>/synth_tools_b200 -lcproc_cuda -o square_grain_b200~.pd_linux
 * tests/test_gpu_dropin.py builds it against tests/c/fakepd (a scripted stand-in for m_pd.h). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "m_pd.h"
#include "cproc_cuda.h"

#define SG_MAX_GRAINS 4096
#define SG_MAX_BLOCK 4096
#define SG_MAX_PIPES 4

struct square_grain_b200 {
    t_object x_obj;
    t_float x_f;
    t_float brightness;
    t_float threshold;
    int slot;                 /* row in the shared batches */
};

static t_class *square_grain_b200_class;

/* One pipeline per block size in use (objects under a [block~] run at another size): rows of its size collected in the
 * running tick, the rendered rows of the previous one.  A tick ends for a pipeline when an object that has already
 * performed in it performs again -- not when a count is reached, so objects in a switched-off subpatch (whose perform
 * routine is not called) neither stall the others nor advance: a row that was not collected is not rendered, its state
 * stays, and the object gets silence for the first block after it comes back. */
struct sg_pipe {
    int block;                              /* frames per row; 0 = free */
    float *in, *out;                        /* [slot][block] */
    unsigned char ran[SG_MAX_GRAINS];       /* collected in the running tick */
    unsigned char have[SG_MAX_GRAINS];      /* out row holds the render of the previous tick */
    int any;                                /* rows collected in the running tick */
    cproc_cuda_batch *batch;
    int capacity;
};
static struct sg_pipe sg_pipes[SG_MAX_PIPES];
static cproc_cuda_ctx *sg_ctx;
static int sg_count;                    /* objects alive (slots 0..sg_count-1) */
static struct square_grain_b200 *sg_obj[SG_MAX_GRAINS];
static float sg_state[SG_MAX_GRAINS];   /* grain state carried on the host between renders (the struct's `state`) */
static float sg_th[SG_MAX_GRAINS];      /* threshold in force when the row was collected (messages arrive between ticks) */
static int sg_failed;

static struct sg_pipe *sg_pipe_for(int n) {
    struct sg_pipe *free_pipe = NULL;
    for (int k = 0; k < SG_MAX_PIPES; k++) {
        if (sg_pipes[k].block == n) return &sg_pipes[k];
        if (!sg_pipes[k].block && !free_pipe) free_pipe = &sg_pipes[k];
    }
    if (!free_pipe) {                                        /* every pipeline is taken: recycle one with nothing pending */
        for (int k = 0; k < SG_MAX_PIPES && !free_pipe; k++) if (!sg_pipes[k].any) free_pipe = &sg_pipes[k];
        if (!free_pipe) return NULL;
        free(free_pipe->in); free(free_pipe->out);
        if (free_pipe->batch) cproc_cuda_free(free_pipe->batch);
        memset(free_pipe, 0, sizeof(*free_pipe));
    }
    free_pipe->block = n;
    free_pipe->in = (float *)calloc((size_t)SG_MAX_GRAINS * (size_t)n, sizeof(float));
    free_pipe->out = (float *)calloc((size_t)SG_MAX_GRAINS * (size_t)n, sizeof(float));
    if (!free_pipe->in || !free_pipe->out) { free(free_pipe->in); free(free_pipe->out); memset(free_pipe, 0, sizeof(*free_pipe)); return NULL; }
    return free_pipe;
}

/* renders the rows the pipeline collected during the tick that just ended */
static void sg_render(struct sg_pipe *p) {
    memset(p->have, 0, sizeof(p->have));
    if (sg_failed || sg_count == 0 || !p->any) { memset(p->ran, 0, sizeof(p->ran)); p->any = 0; return; }
    int rc = 0;
    if (!sg_ctx) rc = cproc_cuda_open(0, NULL, &sg_ctx);
    if (!rc && (!p->batch || p->capacity != sg_count)) {
        if (p->batch) cproc_cuda_free(p->batch);
        p->batch = NULL;
        cproc_cuda_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.proc = CPROC_CUDA_SQUARE_GRAIN; cfg.layout = CPROC_CUDA_PLANAR;
        rc = cproc_cuda_alloc(sg_ctx, &cfg, (uint64_t)sg_count, &p->batch);
        p->capacity = sg_count;
    }
    if (!rc) {
        static float before[SG_MAX_GRAINS];
        memcpy(before, sg_state, sizeof(float) * (size_t)sg_count);
        rc = cproc_cuda_upload_param(p->batch, sg_th, sizeof(float));
        cproc_cuda_io io;
        memset(&io, 0, sizeof(io));
        io.in = p->in; io.out = p->out; io.layout = CPROC_CUDA_PLANAR;
        /* the objects' states travel with the tick: records in, render, records out, one synchronisation */
        if (!rc) rc = cproc_cuda_run_period(p->batch, (uint64_t)p->block, &io, sg_state, sizeof(float));
        if (!rc) for (int g = 0; g < sg_count; g++) { if (p->ran[g]) p->have[g] = 1; else sg_state[g] = before[g]; }   /* rows that were not collected did not run */
    }
    memset(p->ran, 0, sizeof(p->ran));
    p->any = 0;
    if (rc) { sg_failed = rc; memset(p->have, 0, sizeof(p->have)); post("square_grain_b200~: render failed (%d): %s", rc, cproc_cuda_last_error(sg_ctx)); }
}

static t_int *square_grain_b200_perform(t_int *w) {
    struct square_grain_b200 *x = (struct square_grain_b200 *)(w[1]);
    const int n = (int)(w[2]);
    t_float *in = (t_float *)(w[3]);
    t_float *out = (t_float *)(w[4]);
    struct sg_pipe *p = n > 0 && n <= SG_MAX_BLOCK ? sg_pipe_for(n) : NULL;
    if (!p) { memset(out, 0, sizeof(t_float) * (size_t)(n > 0 ? n : 0)); return w + 5; }
    if (p->ran[x->slot]) sg_render(p);                       /* this object again: the pipeline's tick is over */
    /* in may alias out (Pd in-place DSP): take the inlet block before handing out the previous tick's result */
    float *row_in = p->in + (size_t)x->slot * n, *row_out = p->out + (size_t)x->slot * n;
    memcpy(row_in, in, sizeof(float) * (size_t)n);
    sg_th[x->slot] = x->threshold;
    if (p->have[x->slot]) memcpy(out, row_out, sizeof(float) * (size_t)n);
    else memset(out, 0, sizeof(float) * (size_t)n);
    p->ran[x->slot] = 1; p->any = 1;
    return w + 5;
}

static void square_grain_b200_dsp(struct square_grain_b200 *x, t_signal **sp) {
    dsp_add(square_grain_b200_perform, 4, x, sp[0]->s_n, sp[0]->s_vec, sp[1]->s_vec);
}
static void square_grain_b200_brightness(struct square_grain_b200 *x, t_floatarg val) { x->brightness = val; }   /* :101-104 */
static void square_grain_b200_threshold(struct square_grain_b200 *x, t_floatarg val) { x->threshold = fabsf(val); } /* :105-109 */

static void *square_grain_b200_new(t_floatarg threshold) {
    if (sg_count >= SG_MAX_GRAINS) return NULL;
    struct square_grain_b200 *x = (struct square_grain_b200 *)pd_new(square_grain_b200_class);
    x->threshold = threshold;                                /* :130 (the reference does not take fabs here either) */
    x->brightness = 1.0f;
    x->slot = sg_count;
    sg_state[sg_count] = 0.0f;                               /* :129 */
    for (int k = 0; k < SG_MAX_PIPES; k++) { sg_pipes[k].ran[sg_count] = 0; sg_pipes[k].have[sg_count] = 0; }
    sg_obj[sg_count++] = x;
    inlet_new(&x->x_obj, &x->x_obj.ob_pd, gensym("float"), gensym("threshold"));
    outlet_new(&x->x_obj, gensym("signal"));
    return x;
}

static void square_grain_b200_free(struct square_grain_b200 *x) {
    /* compact the slots: the last object moves into the freed row */
    const int last = sg_count - 1;
    if (x->slot != last) {
        sg_obj[x->slot] = sg_obj[last]; sg_obj[x->slot]->slot = x->slot;
        sg_state[x->slot] = sg_state[last]; sg_th[x->slot] = sg_th[last];
        for (int k = 0; k < SG_MAX_PIPES; k++) {
            struct sg_pipe *p = &sg_pipes[k];
            if (!p->block) continue;
            memcpy(p->in + (size_t)x->slot * p->block, p->in + (size_t)last * p->block, sizeof(float) * (size_t)p->block);
            memcpy(p->out + (size_t)x->slot * p->block, p->out + (size_t)last * p->block, sizeof(float) * (size_t)p->block);
            p->ran[x->slot] = p->ran[last]; p->have[x->slot] = p->have[last];
        }
    }
    for (int k = 0; k < SG_MAX_PIPES; k++) { sg_pipes[k].ran[last] = 0; sg_pipes[k].have[last] = 0; }
    sg_count = last;
}

void square_grain_b200_tilde_setup(void) {
    square_grain_b200_class = class_new(gensym("square_grain_b200~"), (t_newmethod)square_grain_b200_new, (t_method)square_grain_b200_free,
                                        sizeof(struct square_grain_b200), CLASS_DEFAULT, A_DEFFLOAT, 0);
    CLASS_MAINSIGNALIN(square_grain_b200_class, struct square_grain_b200, x_f);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_dsp, gensym("dsp"), A_CANT, 0);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_threshold, gensym("threshold"), A_FLOAT, 0);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_brightness, gensym("brightness"), A_FLOAT, 0);
}
