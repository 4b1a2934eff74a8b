/* test_dropin.c -- a plain-C host that uses the reference's own entry points
 * (synth_run, cproc_update + cproc_output, square_grain) through
 * libcproc_dropin.so, and the batched C-ABI directly.  Prints results as text;
 * tests/test_gpu_dropin.py compares them with the oracle. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cproc_cuda.h"

typedef uint32_t w;
struct voice { uint32_t note_inc, note_state; };
struct synth { int note2voice[128]; struct voice voice[64]; };
void synth_run(struct synth *, float *vec, int n);
void cproc_update(w *input, w changed);
void cproc_dropin_square_grain_proc(float *state, float threshold, intptr_t n, const float *in, float *out);
int cproc_dropin_last_status(void);
void cproc_dropin_shutdown(void);

/* the host's upcall, as in linux/test_cproc.c:5-7 */
void cproc_output(uint32_t index, w value) { printf("output %u %u\n", index, value); }

static uint32_t xs(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }

int main(void) {
    /* (1) JACK host path: 4 notes on, two 64-frame periods */
    struct synth sy;
    memset(&sy, 0, sizeof(sy));
    const uint32_t inc[4] = {39370533u, 23409859u, 1122405051u, 731558u};   /* notes 69, 60, 127, 0 */
    for (int i = 0; i < 4; i++) sy.voice[i].note_inc = inc[i];
    float vec[64];
    for (int period = 0; period < 2; period++) {
        synth_run(&sy, vec, 64);
        for (int i = 0; i < 64; i++) printf("synth %a\n", vec[i]);
    }
    for (int i = 0; i < 4; i++) printf("phase %u\n", sy.voice[i].note_state);
    /* (2) generated cproc graph, tick by tick (SURVEY 8c anchor sequence) */
    const w seq[11] = {0, 1, 1, 0, 0, 1, 0, 1, 1, 1, 0};
    for (int i = 0; i < 11; i++) { w in[1] = {seq[i]}; cproc_update(in, 1); }
    { w in[1] = {1}; cproc_update(in, 0); cproc_update(in, 3); }
    /* (3) Pd host path: square_grain, in place, two 64-sample blocks */
    float state = 0.0f, buf[64];
    uint32_t s = 12345;
    for (int blk = 0; blk < 2; blk++) {
        for (int i = 0; i < 64; i++) buf[i] = (float)(int32_t)xs(&s) * (1.0f / 2147483648.0f);
        cproc_dropin_square_grain_proc(&state, 0.25f, 64, buf, buf);
        for (int i = 0; i < 64; i++) printf("grain %a\n", buf[i]);
    }
    printf("status %d\n", cproc_dropin_last_status());
    /* (4) the batched ABI from C: 6 PDM v2 channels (2 banks of 3), 4096+32 ticks */
    cproc_cuda_ctx *ctx;
    cproc_cuda_batch *b;
    if (cproc_cuda_open(0, NULL, &ctx)) { printf("open failed: %s\n", cproc_cuda_last_error(NULL)); return 1; }
    cproc_cuda_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.proc = CPROC_CUDA_PDM_V2; cfg.order = 2; cfg.out_shift = 24; cfg.bank_size = 3; cfg.dither_mask = 0x3FF; cfg.ctl_div_log = 12;
    if (cproc_cuda_alloc(ctx, &cfg, 6, &b)) { printf("alloc failed: %s\n", cproc_cuda_last_error(ctx)); return 1; }
    uint32_t chan[6][7];
    memset(chan, 0, sizeof(chan));
    for (int c = 0; c < 6; c++) chan[c][0] = 0x40000000u + 0x10000000u * (uint32_t)c;   /* setpoints */
    const uint32_t prng[2] = {2463534242u, 7u};
    cproc_cuda_upload_state(b, chan, 0);
    cproc_cuda_upload_bank(b, prng, 0);
    const uint64_t F = 4096 + 32;
    uint8_t *duty = malloc(6 * F);
    cproc_cuda_io io;
    memset(&io, 0, sizeof(io));
    io.out = duty; io.layout = CPROC_CUDA_PLANAR;
    int rc = cproc_cuda_run(b, F, &io);
    printf("pdm rc %d\n", rc);
    uint64_t sum[6] = {0};
    for (int c = 0; c < 6; c++) for (uint64_t t = 0; t < F; t++) sum[c] += duty[c * F + t] * (t + 1);
    for (int c = 0; c < 6; c++) printf("pdm %d %llu\n", c, (unsigned long long)sum[c]);
    cproc_cuda_download_state(b, chan, 0);
    for (int c = 0; c < 6; c++) printf("pdmstate %u %u %u %u %u %u %u\n", chan[c][0], chan[c][1], chan[c][2], chan[c][3], chan[c][4], chan[c][5], chan[c][6]);
    /* error path: never aborts */
    rc = cproc_cuda_run(b, 16, NULL);
    printf("err rc %d\n", rc);
    free(duty);
    cproc_cuda_free(b);
    cproc_cuda_close(ctx);
    cproc_dropin_shutdown();
    return 0;
}
