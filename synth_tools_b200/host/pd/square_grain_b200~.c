/* square_grain_b200~.c -- a Pd external with the interface of square_grain~ (linux/synth_tools.c:78-150:
 * one signal inlet, one signal outlet, a float inlet / "threshold" and "brightness" messages) whose
 * Schmitt-trigger loop runs on the B200: SURVEY 8 f-4.
 *
 * One object is one grain.  Pd calls the perform routines of a patch one after the other inside a
 * DSP tick, so a per-object GPU call would cost a launch per grain per 64 samples.  Instead all
 * objects of the library share ONE batch: a perform routine copies its inlet block into its row of
 * the shared input buffer and hands out its row of the previous tick's result; the first perform of
 * a tick renders the rows collected during the previous tick in one cproc_cuda_run (N grains x n
 * frames).  Cost: one block of latency; state, threshold changes and block-size changes stay exact
 * (the batch is the reference loop, bit for bit, delayed by one block).
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

struct square_grain_b200 {
    t_object x_obj;
    t_float x_f;
    t_float brightness;
    t_float threshold;
    int slot;                 /* row in the shared batch */
};

static t_class *square_grain_b200_class;

/* the shared batch */
static cproc_cuda_ctx *sg_ctx;
static cproc_cuda_batch *sg_batch;
static int sg_capacity;                 /* grains the batch was allocated for */
static int sg_count;                    /* objects alive (slots 0..sg_count-1) */
static struct square_grain_b200 *sg_obj[SG_MAX_GRAINS];
static float *sg_in, *sg_out;           /* [slot][block] rows collected in this tick / rendered for the previous one */
static float sg_state[SG_MAX_GRAINS];   /* grain state carried on the host between renders (the struct's `state`) */
static float sg_th[SG_MAX_GRAINS];      /* threshold in force when the row was collected (messages arrive between ticks) */
static int sg_block;                    /* frames per row */
static int sg_filled;                   /* rows written since the last render */
static int sg_have_out;                 /* sg_out holds a rendered tick */
static int sg_failed;

static void sg_render(void) {
    /* renders the rows collected during the previous tick */
    if (sg_failed || sg_count == 0 || sg_block == 0) return;
    int rc = 0;
    if (!sg_ctx) rc = cproc_cuda_open(0, NULL, &sg_ctx);
    if (!rc && (!sg_batch || sg_capacity != sg_count)) {
        if (sg_batch) cproc_cuda_free(sg_batch);
        sg_batch = NULL;
        cproc_cuda_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.proc = CPROC_CUDA_SQUARE_GRAIN; cfg.layout = CPROC_CUDA_PLANAR;
        rc = cproc_cuda_alloc(sg_ctx, &cfg, (uint64_t)sg_count, &sg_batch);
        sg_capacity = sg_count;
    }
    if (!rc) {
        rc = cproc_cuda_upload_state(sg_batch, sg_state, sizeof(float));
        if (!rc) rc = cproc_cuda_upload_param(sg_batch, sg_th, sizeof(float));
        cproc_cuda_io io;
        memset(&io, 0, sizeof(io));
        io.in = sg_in; io.out = sg_out; io.layout = CPROC_CUDA_PLANAR;
        if (!rc) rc = cproc_cuda_run(sg_batch, (uint64_t)sg_block, &io);
        if (!rc) rc = cproc_cuda_download_state(sg_batch, sg_state, sizeof(float));
    }
    if (rc) { sg_failed = rc; post("square_grain_b200~: render failed (%d): %s", rc, cproc_cuda_last_error(sg_ctx)); return; }
    sg_have_out = 1;
}

static t_int *square_grain_b200_perform(t_int *w) {
    struct square_grain_b200 *x = (struct square_grain_b200 *)(w[1]);
    const int n = (int)(w[2]);
    t_float *in = (t_float *)(w[3]);
    t_float *out = (t_float *)(w[4]);
    if (n != sg_block || n > SG_MAX_BLOCK) {                 /* block size changed: restart the pipeline */
        sg_block = n <= SG_MAX_BLOCK ? n : 0; sg_filled = 0; sg_have_out = 0;
        free(sg_in); free(sg_out);
        sg_in = (float *)calloc((size_t)SG_MAX_GRAINS * (size_t)(sg_block ? sg_block : 1), sizeof(float));
        sg_out = (float *)calloc((size_t)SG_MAX_GRAINS * (size_t)(sg_block ? sg_block : 1), sizeof(float));
    }
    if (sg_filled >= sg_count) {                             /* first perform of a new tick: render the previous one */
        sg_render();
        sg_filled = 0;
    }
    /* in may alias out (Pd in-place DSP): hand out the previous tick first into a scratch copy */
    float prev[SG_MAX_BLOCK];
    if (sg_have_out && sg_block) memcpy(prev, sg_out + (size_t)x->slot * sg_block, sizeof(float) * n);
    else memset(prev, 0, sizeof(float) * n);
    if (sg_block) memcpy(sg_in + (size_t)x->slot * sg_block, in, sizeof(float) * n);
    sg_th[x->slot] = x->threshold;
    memcpy(out, prev, sizeof(float) * n);
    sg_filled++;
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
    sg_obj[sg_count++] = x;
    sg_filled = sg_count;                                    /* the next perform starts a fresh tick */
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
        if (sg_block) {
            memcpy(sg_in + (size_t)x->slot * sg_block, sg_in + (size_t)last * sg_block, sizeof(float) * sg_block);
            memcpy(sg_out + (size_t)x->slot * sg_block, sg_out + (size_t)last * sg_block, sizeof(float) * sg_block);
        }
    }
    sg_count = last;
    sg_filled = sg_count;
}

void square_grain_b200_tilde_setup(void) {
    square_grain_b200_class = class_new(gensym("square_grain_b200~"), (t_newmethod)square_grain_b200_new, (t_method)square_grain_b200_free,
                                        sizeof(struct square_grain_b200), CLASS_DEFAULT, A_DEFFLOAT, 0);
    CLASS_MAINSIGNALIN(square_grain_b200_class, struct square_grain_b200, x_f);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_dsp, gensym("dsp"), A_CANT, 0);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_threshold, gensym("threshold"), A_FLOAT, 0);
    class_addmethod(square_grain_b200_class, (t_method)square_grain_b200_brightness, gensym("brightness"), A_FLOAT, 0);
}
