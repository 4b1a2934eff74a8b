/* ref_cproc.c -- batched harness around the UNMODIFIED reference
 * generic/cproc.h (acc_update, edge_update) and linux/test_cproc.c
 * (cproc_update).  Test infrastructure; built only into oracle/_ref. */
#include <stdarg.h>
#include <string.h>
#include <stdint.h>
#include "cproc.h"      /* <reference>/generic/cproc.h via -I */
#include "macros.h"     /* oracle/shim */

static uint32_t cap_index, cap_value, cap_count;
void ref_log_capture(const char *fmt, ...) {
    if (strncmp(fmt, "output", 6) == 0) {
        va_list ap; va_start(ap, fmt);
        cap_index = va_arg(ap, uint32_t);
        cap_value = va_arg(ap, uint32_t);
        va_end(ap);
        cap_count++;
    }
}

/* linux/test_cproc.c compiled whole (main renamed on the command line). */
void cproc_update(w *input, w g);
/* One tick of the reference's own generated graph.  State is function-static
 * inside cproc_update, i.e. ONE instance per loaded copy of this library. */
uint32_t ref_test_cproc_tick(uint32_t in, uint32_t g, uint32_t *index) {
    w input[1] = { in };
    cproc_update(input, g);
    if (index) *index = cap_index;
    return cap_value;
}

uint32_t ref_sizeof(int what) {
    switch (what) {
    case 0: return sizeof(acc_state);
    case 1: return sizeof(acc_input);
    case 2: return sizeof(acc_config);
    case 3: return sizeof(acc_param);
    case 4: return sizeof(edge_state);
    case 5: return sizeof(edge_input);
    case 6: return sizeof(w);
    }
    return 0xFFFFFFFFu;
}

/* Table-driven graphs calling the real DEF_PROC bodies; same table format as
 * the oracle (type 0 = acc, 1 = edge; the reference has no two-input processor, src2 is unused). */
typedef struct { uint32_t type; int32_t src; uint32_t cond_mask; int32_t src2; } ref_node;
void ref_graph_run(const ref_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                   uint32_t out_node, uint32_t *state, uint64_t N, uint64_t F,
                   const uint32_t *in, const uint32_t *changed, uint32_t *out) {
    uint32_t off[64], sw = 0;
    for (uint32_t i = 0; i < n_nodes; i++) { off[i] = sw; sw += nodes[i].type ? sizeof(edge_state) / 4 : sizeof(acc_state) / 4; }
    for (uint64_t n = 0; n < N; n++) {
        uint32_t *st = state + n * sw;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t g = changed ? changed[n * F + t] : (w)-1;
            for (uint32_t i = 0; i < n_nodes; i++) {
                if (!(g & nodes[i].cond_mask)) continue;
                w x = nodes[i].src >= 0 ? st[off[nodes[i].src]]
                                        : in[(n * n_inputs + (uint32_t)(-(nodes[i].src + 1))) * F + t];
                if (nodes[i].type) { const edge_input ei = { .in = x }; edge_update((edge_state *)(st + off[i]), NULL, NULL, &ei); }
                else               { const acc_input  ai = { .in = x }; acc_update((acc_state *)(st + off[i]), NULL, NULL, &ai); }
            }
            out[n * F + t] = st[off[out_node]];
        }
    }
}
