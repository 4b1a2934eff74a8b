/* ref_cproc.c -- batched harness around the UNMODIFIED reference
 * generic/cproc.h (acc_update, edge_update) and linux/test_cproc.c
 * (cproc_update).  Test infrastructure; built only into oracle/_ref. */
#include <stdarg.h>
#include <string.h>
#include <stdint.h>
#include "cproc.h"      /* <reference>/generic/cproc.h via -I */
#include "macros.h"     /* oracle/shim */
#include "cproc_ext.h"  /* <repo>/include: the extension processors, DEF_PROC against the reference's own cproc.h */

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

/* Graphs that use the extension processors: the node table drives the real DEF_PROC-generated NAME_update functions
 * (cproc_ext.h compiled against the reference's cproc.h) on NAME_state / NAME_param / NAME_input structs laid over the
 * word arrays -- the proof that they sit behind the reference's processor API (struct split, update signature). */
uint32_t ref_ext_sizeof(int what) {
    switch (what) {
    case 0: return sizeof(phasor_f_state); case 1: return sizeof(phasor_f_param); case 2: return sizeof(phasor_f_input);
    case 3: return sizeof(svf_state);      case 4: return sizeof(svf_param);      case 5: return sizeof(svf_input);
    case 6: return sizeof(env_state);      case 7: return sizeof(env_param);      case 8: return sizeof(env_input);
    case 9: return sizeof(onepole_state);  case 10: return sizeof(onepole_param); case 11: return sizeof(onepole_input);
    case 12: return sizeof(gain_state);    case 13: return sizeof(gain_param);    case 14: return sizeof(gain_input);
    case 15: return sizeof(asfloat_state); case 16: return sizeof(asfloat_param); case 17: return sizeof(asfloat_input);
    case 18: return sizeof(glide_f_state); case 19: return sizeof(glide_f_param); case 20: return sizeof(glide_f_input);
    case 21: return sizeof(mul_state);     case 22: return sizeof(mul_param);     case 23: return sizeof(mul_input);
    case 24: return sizeof(glide_f_config);
    }
    return 0xFFFFFFFFu;
}
static const uint8_t ext_sw[12] = {1, 2, 0, 0, sizeof(phasor_f_state) / 4, sizeof(svf_state) / 4, sizeof(env_state) / 4, sizeof(onepole_state) / 4, sizeof(gain_state) / 4, sizeof(asfloat_state) / 4,
                                   sizeof(glide_f_state) / 4, sizeof(mul_state) / 4};
static const uint8_t ext_pw[12] = {0, 0, 0, 0, sizeof(phasor_f_param) / 4, sizeof(svf_param) / 4, sizeof(env_param) / 4, sizeof(onepole_param) / 4, sizeof(gain_param) / 4, 0, 0, 0};
static float w_as_f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
void ref_graph_run_ext(const ref_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                       const uint32_t *out_nodes, uint32_t n_out, uint32_t *state, const uint32_t *param, uint64_t N, uint64_t F,
                       const uint32_t *in, const uint32_t *changed, uint32_t *out) {
    uint32_t off[64], poff[64], sw = 0, pw = 0;
    for (uint32_t i = 0; i < n_nodes; i++) { off[i] = sw; sw += ext_sw[nodes[i].type & 0xFF]; poff[i] = pw; pw += ext_pw[nodes[i].type & 0xFF]; }
    for (uint64_t n = 0; n < N; n++) {
        uint32_t *st = state + n * sw;
        const uint32_t *pr = param ? param + n * pw : NULL;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t g = changed ? changed[n * F + t] : (w)-1;
            for (uint32_t i = 0; i < n_nodes; i++) {
                if (!(g & nodes[i].cond_mask)) continue;
                const uint32_t kind = nodes[i].type & 0xFF;
                const int32_t src = nodes[i].src;
                /* the source as the C expression would see it: a `w` lvalue (input[k], an acc / edge .out) or a float .out */
                int src_f = 0; w xw = 0; float xf = 0.0f;
                if (src != (int32_t)0x80000000) {
                    if (src >= 0) { xw = st[off[src]]; src_f = (nodes[src].type & 0xFF) >= 4; if (src_f) xf = w_as_f(xw); }
                    else xw = in[(n * n_inputs + (uint32_t)(-(src + 1))) * F + t];
                }
                void *s = st + off[i];
                const void *p = pr ? pr + poff[i] : NULL;
                switch (kind) {
                case 0: { const acc_input ai = { .in = xw }; acc_update(s, NULL, NULL, &ai); break; }
                case 1: { const edge_input ei = { .in = xw }; edge_update(s, NULL, NULL, &ei); break; }
                case 4: { const phasor_f_input pi = { .mod = xw }; phasor_f_update(s, NULL, p, &pi); break; }
                /* `.in = <expr>`: a w expression converts to float by value, a float .out is assigned as is */
                case 5: { const svf_input vi = { .in = src_f ? xf : (float)xw }; svf_update(s, NULL, p, &vi); break; }
                case 6: { const env_input vi = { .in = src_f ? xf : (float)xw }; env_update(s, NULL, p, &vi); break; }
                case 7: { const onepole_input vi = { .in = src_f ? xf : (float)xw }; onepole_update(s, NULL, p, &vi); break; }
                case 8: { const gain_input vi = { .in = src_f ? xf : (float)xw }; gain_update(s, NULL, p, &vi); break; }
                case 9: { const asfloat_input vi = { .in = xw }; asfloat_update(s, NULL, NULL, &vi); break; }
                case 10: { const glide_f_config gc = { .div_log = (nodes[i].type >> 8) & 0xFF }; const glide_f_input vi = { .in = src_f ? xf : (float)xw }; glide_f_update(s, &gc, NULL, &vi); break; }
                case 11: {                                       /* second input: the same conversion rule */
                    const int32_t s2 = nodes[i].src2; int s2_f = 0; w yw = 0; float yf = 0.0f;
                    if (s2 != (int32_t)0x80000000) {
                        if (s2 >= 0) { yw = st[off[s2]]; s2_f = (nodes[s2].type & 0xFF) >= 4; if (s2_f) yf = w_as_f(yw); }
                        else yw = in[(n * n_inputs + (uint32_t)(-(s2 + 1))) * F + t];
                    }
                    const mul_input vi = { .in = src_f ? xf : (float)xw, .gain = s2_f ? yf : (float)yw };
                    mul_update(s, NULL, NULL, &vi); break; }
                }
            }
            for (uint32_t k = 0; k < n_out; k++) out[(n * n_out + k) * F + t] = st[off[out_nodes[k]]];
        }
    }
}
