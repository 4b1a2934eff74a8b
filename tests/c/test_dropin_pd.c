/* test_dropin_pd.c -- built TOGETHER with synth_tools_b200/host/dropin.c under -DCPROC_HAVE_PD against
 * tests/c/fakepd/m_pd.h (this image has no Pd): the Pd host's DSP-chain entry calls square_grain_proc by the
 * reference's own name and prototype (linux/synth_tools.c:85-86, called as at :110-119), and the firmware
 * plugin's message handler drives the generated graph one TAG_U32 message at a time
 * (stm32f103/mod_cproc_plugin.c:24-38).  Prints results as text; tests/test_gpu_dropin.py checks them. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "m_pd.h"
#include "cproc_cuda.h"

/* synth_tools.c:78-84 */
struct square_grain {
    t_object x_obj;
    t_float x_f;
    t_float brightness;
    t_float threshold;
    t_float state;
};
void square_grain_proc(struct square_grain *s, t_int n, t_float *in, t_float *out);

/* the Pd perform routine: w[1] object, w[2] block size, w[3] inlet vector, w[4] outlet vector */
static t_int *perform(t_int *w) {
    square_grain_proc((struct square_grain *)w[1], (t_int)w[2], (t_float *)w[3], (t_float *)w[4]);
    return w + 5;
}

typedef uint32_t w_t;
int cproc_dropin_handle_tag_u32(const uint32_t *args, uint32_t nb_args, uint32_t nb_bytes);
void cproc_dropin_shutdown(void);
int cproc_dropin_last_status(void);
void cproc_output(uint32_t index, w_t value) { printf("output %u %u\n", index, value); }

static uint32_t xs(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }

int main(void) {
    /* (1) two objects with their own state and threshold, three 64-sample blocks each, in place; the threshold
           message (synth_tools.c:105-109) lands between blocks 1 and 2 */
    struct square_grain obj[2];
    memset(obj, 0, sizeof(obj));
    obj[0].threshold = 0.25f; obj[1].threshold = 0.05f;
    uint32_t s = 4242;
    for (int blk = 0; blk < 3; blk++) {
        for (int o = 0; o < 2; o++) {
            t_float buf[64];
            for (int i = 0; i < 64; i++) buf[i] = (t_float)(int32_t)xs(&s) * (1.0f / 2147483648.0f);
            t_int prog[5] = {0, (t_int)&obj[o], 64, (t_int)buf, (t_int)buf};
            if (perform(prog) != prog + 5) printf("fail perform\n");
            for (int i = 0; i < 64; i++) printf("grain%d %a\n", o, buf[i]);
        }
        if (blk == 1) obj[0].threshold = 0.4f;
    }
    printf("state %a %a\n", obj[0].state, obj[1].state);
    printf("offset %zu\n", (size_t)((char *)&obj[0].state - (char *)&obj[0]) - sizeof(t_object));
    /* (2) TAG_U32 messages [i, v]: cproc_input[i] = v; cproc_update(cproc_input, -1) */
    const uint32_t seq[9] = {0, 1, 1, 0, 1, 0, 0, 1, 0};
    for (int k = 0; k < 9; k++) {
        uint32_t args[2] = {0, seq[k]};
        printf("msg %d\n", cproc_dropin_handle_tag_u32(args, 2, 0));
    }
    { uint32_t bad[2] = {1, 5}; printf("msg %d\n", cproc_dropin_handle_tag_u32(bad, 2, 0)); }      /* i >= CPROC_NB_INPUTS */
    { uint32_t bad[3] = {0, 5, 6}; printf("msg %d\n", cproc_dropin_handle_tag_u32(bad, 3, 0)); }   /* not a two-word message */
    { uint32_t bad[2] = {0, 5}; printf("msg %d\n", cproc_dropin_handle_tag_u32(bad, 2, 4)); }      /* carries bytes */
    printf("status %d\n", cproc_dropin_last_status());
    /* (3) the same through the batched ABI: 3 graph instances, events addressed to one of them */
    cproc_cuda_ctx *ctx;
    cproc_cuda_batch *b;
    if (cproc_cuda_open(0, NULL, &ctx)) { printf("fail open\n"); return 1; }
    static const cproc_cuda_node nodes[2] = { { CPROC_CUDA_NODE_EDGE, -1, 1, 0 }, { CPROC_CUDA_NODE_ACC, 0, 1, 0 } };
    cproc_cuda_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.proc = CPROC_CUDA_GRAPH; cfg.nodes = nodes; cfg.n_nodes = 2; cfg.n_inputs = 1; cfg.out_node = 1;
    if (cproc_cuda_alloc(ctx, &cfg, 3, &b)) { printf("fail alloc\n"); return 1; }
    for (int k = 0; k < 12; k++) {
        uint32_t out = 99;
        int inst = k % 3 == 2 ? 2 : k & 1;
        int rc = cproc_cuda_graph_event(b, (uint64_t)inst, 0, (uint32_t)(k >> 1) & 1, &out);
        printf("event %d %d %u\n", rc, inst, out);
    }
    uint32_t all[3];
    printf("tick %d", cproc_cuda_graph_tick(b, 0xFFFFFFFFu, all));
    printf(" %u %u %u\n", all[0], all[1], all[2]);
    printf("bad %d %d\n", cproc_cuda_graph_set_input(b, 0, 1, 7), cproc_cuda_graph_event(b, 3, 0, 1, NULL));
    cproc_cuda_free(b);
    cproc_cuda_close(ctx);
    cproc_dropin_shutdown();
    return 0;
}
