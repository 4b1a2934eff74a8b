/* cproc_oracle.c -- CPU restatement of the synth_tools per-sample hot path.
 * TEST INFRASTRUCTURE ONLY (see cproc_oracle.h).  Plain C99, no threads
 * except the optional OpenMP loops over independent instances.
 * Build: gcc -std=gnu99 -O2 -fwrapv -ffp-contract=off -mfma -fopenmp -shared -fPIC
 */
#include "cproc_oracle.h"
#include <math.h>
#include <string.h>

/* ======================================================================= */
/* generic/cproc.h                                                          */

/* cproc.h:140-142  "s->out += i->in;" on w = uint32_t (wraps mod 2^32). */
void orc_acc_update(orc_acc_state *s, orc_w in) { s->out += in; }

/* cproc.h:151-154  out = (in != last); last = in. */
void orc_edge_update(orc_edge_state *s, orc_w in) {
    s->out = (in != s->last);
    s->last = in;
}

/* mod_pdm_pwm.c:129-143 + mod_controlrate.c:28-40 on one parameter (see cproc_oracle.h) */
void orc_glide_update(orc_glide_state *s, orc_w in, uint32_t div_log) {
    if (s->count == 0) {
        s->out = s->pos1; s->vel0 = s->vel1;                      /* PDM_COPY_LINE */
        s->pos1 += s->vel1 << div_log;                            /* pdm_update_line */
        int32_t span = (int32_t)(in - s->pos1);
        s->vel1 = (uint32_t)(span >> div_log);
    }
    s->out += s->vel0;                                            /* pdm_update_glide */
    s->count = (s->count + 1) & ((1u << div_log) - 1u);
}

uint32_t orc_node_state_words(uint32_t type) {
    /* sizeof(acc_state)=4, sizeof(edge_state)=8 (cproc.h:134,145); glide: 5 words */
    switch (type & 0xFF) {
    case ORC_NODE_EDGE: return 2u;
    case ORC_NODE_GLIDE: return 5u;
    case ORC_NODE_PDM: return 1u + ((type >> 8) & 7u);
    case ORC_NODE_PHASOR_F: case ORC_NODE_SVF: return 2u;   /* cproc_ext.h: {out, phase}, {out, bp} */
    case ORC_NODE_ENV: case ORC_NODE_GLIDE_F: return 3u;    /* {out, env, t}, {out, step, count} */
    default: return 1u;
    }
}
uint32_t orc_node_param_words(uint32_t type) {
    switch (type & 0xFF) {
    case ORC_NODE_PHASOR_F: case ORC_NODE_ONEPOLE: case ORC_NODE_GAIN: return 1u;   /* {inc}, {a}, {g} */
    case ORC_NODE_SVF: return 2u;                                                   /* {f, q} */
    case ORC_NODE_ENV: return 3u;                                                   /* {attack, release, gate_frames} */
    default: return 0u;
    }
}
uint32_t orc_graph_param_words(const orc_node *nodes, uint32_t n_nodes) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < n_nodes; i++) w += orc_node_param_words(nodes[i].type);
    return w;
}
uint32_t orc_graph_state_words(const orc_node *nodes, uint32_t n_nodes) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < n_nodes; i++) w += orc_node_state_words(nodes[i].type);
    return w;
}

/* One cproc_update() call = one tick of every node whose subgraph condition
 * holds (cproc.h:72-77).  State is persistent and zero-initialised by the
 * caller (cproc.h:65-66,73); a node that is skipped keeps its .out, which
 * downstream nodes still read (test_cproc.c:15 reads n1.out). */
void orc_graph_run(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                   uint32_t out_node, uint32_t *state, uint64_t N, uint64_t F,
                   const uint32_t *in, const uint32_t *changed, uint32_t *out) {
    uint32_t sw = orc_graph_state_words(nodes, n_nodes);
    uint32_t off[64];
    uint32_t o = 0;
    for (uint32_t i = 0; i < n_nodes && i < 64; i++) { off[i] = o; o += orc_node_state_words(nodes[i].type); }
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t *st = state + (uint64_t)n * sw;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t g = changed ? changed[(uint64_t)n * F + t] : 0xFFFFFFFFu;
            for (uint32_t i = 0; i < n_nodes; i++) {
                if (!(g & nodes[i].cond_mask)) continue;
                uint32_t x = nodes[i].src >= 0
                    ? st[off[nodes[i].src]] /* .out is the first state word */
                    : in[((uint64_t)n * n_inputs + (uint32_t)(-(nodes[i].src + 1))) * F + t];
                if ((nodes[i].type & 0xFF) == ORC_NODE_PDM) {
                    uint32_t d = nodes[i].src2 >= 0 ? st[off[nodes[i].src2]]
                        : in[((uint64_t)n * n_inputs + (uint32_t)(-(nodes[i].src2 + 1))) * F + t];
                    uint32_t order = (nodes[i].type >> 8) & 7u, sh = (nodes[i].type >> 11) & 31u;
                    st[off[i]] = orc_pdm_update(st + off[i] + 1, order, x, sh, d);   /* pdm.h:13-77 */
                    continue;
                }
                switch (nodes[i].type & 0xFF) {
                case ORC_NODE_EDGE: orc_edge_update((orc_edge_state *)(st + off[i]), x); break;
                case ORC_NODE_GLIDE: orc_glide_update((orc_glide_state *)(st + off[i]), x, (nodes[i].type >> 8) & 0xFF); break;
                default: orc_acc_update((orc_acc_state *)(st + off[i]), x); break;
                }
            }
            out[(uint64_t)n * F + t] = st[off[out_node]];
        }
    }
}

void orc_graph_run_multi(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                         const uint32_t *out_nodes, uint32_t n_out, uint32_t *state, uint64_t N, uint64_t F,
                         const uint32_t *in, const uint32_t *changed, uint32_t *out) {
    uint32_t sw = orc_graph_state_words(nodes, n_nodes);
    uint32_t off[64];
    uint32_t o = 0;
    for (uint32_t i = 0; i < n_nodes && i < 64; i++) { off[i] = o; o += orc_node_state_words(nodes[i].type); }
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t *st = state + (uint64_t)n * sw;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t g = changed ? changed[(uint64_t)n * F + t] : 0xFFFFFFFFu;
            for (uint32_t i = 0; i < n_nodes; i++) {
                if (!(g & nodes[i].cond_mask)) continue;
                uint32_t x = nodes[i].src >= 0
                    ? st[off[nodes[i].src]] /* .out is the first state word */
                    : in[((uint64_t)n * n_inputs + (uint32_t)(-(nodes[i].src + 1))) * F + t];
                if ((nodes[i].type & 0xFF) == ORC_NODE_PDM) {
                    uint32_t d = nodes[i].src2 >= 0 ? st[off[nodes[i].src2]]
                        : in[((uint64_t)n * n_inputs + (uint32_t)(-(nodes[i].src2 + 1))) * F + t];
                    uint32_t order = (nodes[i].type >> 8) & 7u, sh = (nodes[i].type >> 11) & 31u;
                    st[off[i]] = orc_pdm_update(st + off[i] + 1, order, x, sh, d);   /* pdm.h:13-77 */
                    continue;
                }
                switch (nodes[i].type & 0xFF) {
                case ORC_NODE_EDGE: orc_edge_update((orc_edge_state *)(st + off[i]), x); break;
                case ORC_NODE_GLIDE: orc_glide_update((orc_glide_state *)(st + off[i]), x, (nodes[i].type >> 8) & 0xFF); break;
                default: orc_acc_update((orc_acc_state *)(st + off[i]), x); break;
                }
            }
            for (uint32_t q = 0; q < n_out; q++) out[((uint64_t)n * n_out + q) * F + t] = st[off[out_nodes[q]]];
        }
    }
}


/* Extension processors as graph nodes: the update bodies of include/cproc_ext.h restated on word arrays
 * (floats as bit patterns); each statement one IEEE rounding, fmaf fused. */
static float orc_bits_f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t orc_f_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
static int orc_kind_in_float(uint32_t kind) { return kind == ORC_NODE_SVF || kind == ORC_NODE_ENV || kind == ORC_NODE_ONEPOLE || kind == ORC_NODE_GAIN || kind == ORC_NODE_GLIDE_F || kind == ORC_NODE_MUL; }
static int orc_kind_out_float(uint32_t kind) { return kind >= ORC_NODE_PHASOR_F && kind < ORC_NODE_KINDS; }
static void orc_ext_update(uint32_t type, uint32_t *s, const uint32_t *p, uint32_t x, uint32_t x2) {
    const uint32_t kind = type & 0xFF;
    switch (kind) {
    case ORC_NODE_GLIDE_F: {                                 /* cproc_ext.h glide_f: the scale by 2^-L is exact */
        uint32_t L = (type >> 8) & 0xFF;
        if (s[2] == 0) s[1] = orc_f_bits((orc_bits_f(x) - orc_bits_f(s[0])) * (1.0f / (float)(1u << L)));
        s[0] = orc_f_bits(orc_bits_f(s[0]) + orc_bits_f(s[1]));
        s[2] = (s[2] + 1) & ((1u << L) - 1u);
        break; }
    case ORC_NODE_MUL: s[0] = orc_f_bits(orc_bits_f(x) * orc_bits_f(x2)); break;
    case ORC_NODE_PHASOR_F:                                  /* cproc_ext.h phasor_f: read, then advance */
        s[0] = orc_f_bits((float)(int32_t)s[1] * (1.0f / 2147483648.0f));
        s[1] += p[0] + x;
        break;
    case ORC_NODE_SVF: {                                     /* cproc_ext.h svf */
        float f = orc_bits_f(p[0]), q = orc_bits_f(p[1]), bp = orc_bits_f(s[1]);
        float lp = fmaf(f, bp, orc_bits_f(s[0]));
        float hp = orc_bits_f(x) - lp;
        hp = fmaf(-q, bp, hp);
        s[1] = orc_f_bits(fmaf(f, hp, bp));
        s[0] = orc_f_bits(lp);
        break; }
    case ORC_NODE_ENV: {                                     /* cproc_ext.h env */
        float e = orc_bits_f(s[1]);
        if (s[2] < p[2]) { e = e + orc_bits_f(p[0]); if (e > 1.0f) e = 1.0f; }
        else { e = e - orc_bits_f(p[1]); if (e < 0.0f) e = 0.0f; }
        s[1] = orc_f_bits(e);
        s[2] += 1;
        s[0] = orc_f_bits(orc_bits_f(x) * e);
        break; }
    case ORC_NODE_ONEPOLE: { float y = orc_bits_f(s[0]); s[0] = orc_f_bits(fmaf(orc_bits_f(p[0]), orc_bits_f(x) - y, y)); break; }
    case ORC_NODE_GAIN: s[0] = orc_f_bits(orc_bits_f(p[0]) * orc_bits_f(x)); break;
    case ORC_NODE_ASFLOAT: s[0] = x; break;
    }
}

void orc_graph_run_ext(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                       const uint32_t *out_nodes, uint32_t n_out, uint32_t *state, const uint32_t *param, uint64_t N, uint64_t F,
                       const uint32_t *in, const uint32_t *changed, uint32_t *out) {
    uint32_t sw = orc_graph_state_words(nodes, n_nodes), pw = orc_graph_param_words(nodes, n_nodes);
    uint32_t off[64], poff[64];
    uint32_t o = 0, q = 0;
    for (uint32_t i = 0; i < n_nodes && i < 64; i++) { off[i] = o; o += orc_node_state_words(nodes[i].type); poff[i] = q; q += orc_node_param_words(nodes[i].type); }
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t *st = state + (uint64_t)n * sw;
        const uint32_t *pr = param ? param + (uint64_t)n * pw : NULL;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t g = changed ? changed[(uint64_t)n * F + t] : 0xFFFFFFFFu;
            for (uint32_t i = 0; i < n_nodes; i++) {
                if (!(g & nodes[i].cond_mask)) continue;
                uint32_t kind = nodes[i].type & 0xFF;
                uint32_t xs[2] = {0, 0};
                for (int j = 0; j < (kind == ORC_NODE_PDM || kind == ORC_NODE_MUL ? 2 : 1); j++) {
                    int32_t src = j ? nodes[i].src2 : nodes[i].src;
                    if (src == ORC_SRC_ZERO) continue;                       /* 0 / +0.0f */
                    uint32_t v; int is_f = 0;
                    if (src >= 0) { v = st[off[src]]; is_f = orc_kind_out_float(nodes[src].type & 0xFF); }   /* .out is the first state word */
                    else v = in[((uint64_t)n * n_inputs + (uint32_t)(-(src + 1))) * F + t];
                    /* `.in = <w expression>` into a float member converts by value (cproc.h:75) */
                    xs[j] = (orc_kind_in_float(kind) && !is_f) ? orc_f_bits((float)v) : v;
                }
                switch (kind) {
                case ORC_NODE_PDM: {
                    uint32_t order = (nodes[i].type >> 8) & 7u, sh = (nodes[i].type >> 11) & 31u;
                    st[off[i]] = orc_pdm_update(st + off[i] + 1, order, xs[0], sh, xs[1]);   /* pdm.h:13-77 */
                    break; }
                case ORC_NODE_EDGE: orc_edge_update((orc_edge_state *)(st + off[i]), xs[0]); break;
                case ORC_NODE_GLIDE: orc_glide_update((orc_glide_state *)(st + off[i]), xs[0], (nodes[i].type >> 8) & 0xFF); break;
                case ORC_NODE_ACC: orc_acc_update((orc_acc_state *)(st + off[i]), xs[0]); break;
                default: orc_ext_update(nodes[i].type, st + off[i], pr ? pr + poff[i] : NULL, xs[0], xs[1]); break;
                }
            }
            for (uint32_t k = 0; k < n_out; k++) out[((uint64_t)n * n_out + k) * F + t] = st[off[out_nodes[k]]];
        }
    }
}

/* ======================================================================= */
/* stm32f103/pdm.h                                                          */

/* pdm.h:13-24 (order 1), :32-40 (2), :48-57 (3), :67-77 (4).
 * out_q = sK >> sh; out_a = (out_q << sh) + dither (order 1: no dither);
 * s1 += input - out_a; s(k) += s(k-1) - out_a; return out_q. */
uint32_t orc_pdm_update(uint32_t *s, uint32_t order, uint32_t input,
                        uint32_t out_shift, uint32_t dither) {
    uint32_t out_q = s[order - 1] >> out_shift;
    uint32_t out_a = (out_q << out_shift) + (order == 1 ? 0u : dither);
    s[0] += input - out_a;
    for (uint32_t k = 1; k < order; k++) s[k] += s[k - 1] - out_a;
    return out_q;
}

void orc_pdm_run(uint32_t order, uint32_t *state, uint64_t N, uint64_t F,
                 const uint32_t *in, const uint32_t *in_const,
                 uint32_t out_shift, const uint32_t *dither, uint32_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t *s = state + (uint64_t)n * order;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t x = in ? in[(uint64_t)n * F + t] : in_const[n];
            out[(uint64_t)n * F + t] = orc_pdm_update(s, order, x, out_shift, dither ? dither[t] : 0u);
        }
    }
}

/* ======================================================================= */
/* dither PRNG: uc_tools xorshift.h random_u32() -- PARITY UNPINNED.
 * Marsaglia xorshift32 with the (13,17,5) triple; returns the new state. */
uint32_t orc_xorshift32(uint32_t *state) {
    uint32_t x = *state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    *state = x;
    return x;
}

/* ======================================================================= */
/* stm32f103/mod_pdm.c  (v1)                                                */

/* mod_pdm.c:230-244: dither_setpoint = setpoint + dither (wraps);
 * "adds accu, accu, dither_setpoint" sets C to the unsigned carry out;
 * "rrx" rotates C into the shift register MSB.  We return C. */
uint32_t orc_v1_channel_update(orc_v1_channel *c, uint32_t dither) {
    uint32_t x = c->setpoint + dither;
    uint32_t a = c->accu + x;
    uint32_t carry = a < x;
    c->accu = a;
    return carry;
}

void orc_pdm_v1_run(orc_v1_channel *ch, uint64_t N, uint32_t bank_size,
                    uint32_t *prng, const uint32_t *dither_ext,
                    uint32_t dither_mask, uint64_t F, uint8_t *bits) {
    uint64_t n_banks = (N + bank_size - 1) / bank_size;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)n_banks; b++) {
        uint64_t c0 = (uint64_t)b * bank_size;
        uint64_t c1 = c0 + bank_size < N ? c0 + bank_size : N;
        uint32_t rng = prng ? prng[b] : 0;
        for (uint64_t t = 0; t < F; t++) {
            /* mod_pdm.c:261: one random word per tick for all channels */
            uint32_t d = (dither_ext ? dither_ext[(uint64_t)b * F + t] : orc_xorshift32(&rng)) & dither_mask;
            for (uint64_t c = c0; c < c1; c++)
                bits[c * F + t] = (uint8_t)orc_v1_channel_update(&ch[c], d);
        }
        if (prng) prng[b] = rng;
    }
}

/* mod_pdm.c:167-175 */
uint32_t orc_pwm_update(uint32_t *pwm_phase, uint32_t pwm_speed) {
    uint32_t phase = *pwm_phase;
    uint32_t duty = phase >> 16;
    phase = (phase + pwm_speed + (phase >> 9)) & 0xFFFFFFu; /* PHASE_MASK :162 */
    *pwm_phase = phase;
    return duty;
}
void orc_pwm_run(uint32_t *phase, const uint32_t *speed, uint64_t N, uint64_t F,
                 uint8_t *duty) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++)
        for (uint64_t t = 0; t < F; t++)
            duty[(uint64_t)n * F + t] = (uint8_t)orc_pwm_update(&phase[n], speed[n]);
}

/* ======================================================================= */
/* stm32f103/mod_pdm_pwm.c + mod_controlrate.c  (v2)                        */

/* mod_controlrate.c:28-40 on line[1] = words 3,4; setpoint = word 0. */
void orc_v2_update_line(uint32_t *c, uint32_t ctl_div_log) {
    uint32_t pos = c[3];
    int32_t vel = (int32_t)c[4];
    pos += (uint32_t)vel << ctl_div_log;             /* :32 */
    int32_t span = (int32_t)(c[0] - pos);            /* :33 */
    c[3] = pos;
    c[4] = (uint32_t)(span >> ctl_div_log);          /* :34 arithmetic shift */
}

void orc_pdm_v2_run(uint32_t *chan, uint32_t order, uint64_t N,
                    uint32_t bank_size, uint32_t *prng,
                    const uint32_t *dither_ext, uint32_t dither_mask,
                    uint32_t *count, uint32_t ctl_div_log, uint32_t out_shift,
                    const uint32_t *setpoints, uint64_t F, uint8_t *duty) {
    uint64_t n_banks = (N + bank_size - 1) / bank_size;
    uint32_t W = ORC_V2_WORDS(order);
    uint32_t div = 1u << ctl_div_log;
    uint32_t count0 = *count;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)n_banks; b++) {
        uint64_t c0 = (uint64_t)b * bank_size;
        uint64_t c1 = c0 + bank_size < N ? c0 + bank_size : N;
        uint32_t rng = prng ? prng[b] : 0;
        uint32_t cnt = count0;
        uint64_t row = 0;
        for (uint64_t t = 0; t < F; t++) {
            /* mod_pdm_pwm.c:127 */
            uint32_t d = (dither_ext ? dither_ext[(uint64_t)b * F + t] : orc_xorshift32(&rng)) & dither_mask;
            if (cnt == 0) {                           /* :129 */
                for (uint64_t c = c0; c < c1; c++) {
                    uint32_t *x = chan + c * W;
                    if (setpoints) x[0] = setpoints[row * N + c]; /* mod_synth.c:104-111 */
                    x[1] = x[3]; x[2] = x[4];         /* PDM_COPY_LINE :118-119 */
                    orc_v2_update_line(x, ctl_div_log); /* control_trigger :136 -> mod_controlrate.c:46-57 */
                }
                row++;
            }
            for (uint64_t c = c0; c < c1; c++) {
                uint32_t *x = chan + c * W;
                x[1] += x[2];                         /* pdm_update_glide :101-104 */
                duty[c * F + t] = (uint8_t)orc_pdm_update(x + 5, order, x[1], out_shift, d); /* :108-116 */
            }
            cnt = (cnt + 1) % div;                    /* :141 */
        }
        if (prng) prng[b] = rng;
    }
    *count = (uint32_t)((count0 + F) % div);
}

/* ======================================================================= */
/* linux/synth.c                                                            */

/* synth.c:94-98 note_tab as evaluated by gcc at compile time (double ->
 * uint32 truncation); values re-derived from the compiled reference in
 * tests/test_oracle_vs_ref.py. */
static const uint32_t orc_note_tab[12] = {
    594573364u, 629928536u, 667386036u, 707070875u,
    749115497u, 793660223u, 840853716u, 890853479u,
    943826384u, 999949221u, 1059409296u, 1122405051u,
};
/* synth.c:108-125: midi_tab maps note -> (octave shift, table index):
 * notes 0..7 are NOTE(10,4..11); then OCTAVE(9) ... OCTAVE(0). */
uint32_t orc_note_to_inc(int note) {
    note &= 127;
    int octave, n;
    if (note < 8) { octave = 10; n = note + 4; }
    else { octave = 9 - (note - 8) / 12; n = (note - 8) % 12; }
    return orc_note_tab[n] >> octave;
}

void orc_voice_bank_run(orc_voice *v, uint64_t N, uint64_t voices_per_bus,
                        int mode, uint64_t F, int32_t *isum, float *vec) {
    uint64_t n_bus = (N + voices_per_bus - 1) / voices_per_bus;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)n_bus; b++) {
        uint64_t v0 = (uint64_t)b * voices_per_bus;
        uint64_t v1 = v0 + voices_per_bus < N ? v0 + voices_per_bus : N;
        for (uint64_t t = 0; t < F; t++) {
            uint32_t sum = 0; /* wraps like the reference's int under -fwrapv */
            for (uint64_t i = v0; i < v1; i++) {
                if (!v[i].note_inc) continue;              /* :173 / :186 */
                if (mode == ORC_MIX_SAW) {
                    int32_t p = (int32_t)v[i].note_state;  /* :175 */
                    sum += (uint32_t)(p >> 4);             /* :176 arithmetic */
                } else {
                    sum |= v[i].note_state & 0x80000000u;  /* :188-190 */
                }
                v[i].note_state += v[i].note_inc;          /* :177 / :191 */
            }
            if (isum) isum[(uint64_t)b * F + t] = (int32_t)sum;
            if (vec) {
                /* :180 (1.0/2^32) * (float)(int)sum ; :194 (float)(unsigned)accu */
                float f = mode == ORC_MIX_SAW ? (float)(int32_t)sum : (float)sum;
                vec[(uint64_t)b * F + t] = (float)((1.0 / 4294967296.0) * f);
            }
        }
    }
}

/* ======================================================================= */
/* linux/synth_tools.c:85-100                                               */
void orc_square_grain_run(float *state_v, const float *threshold, uint64_t N,
                          uint64_t F, const float *in, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        float state = state_v[n];
        float thresh = threshold[n];
        for (uint64_t i = 0; i < F; i++) {
            float val = in[(uint64_t)n * F + i];      /* :90 read before write: alias-safe */
            out[(uint64_t)n * F + i] = state;         /* :91 */
            if ((state >= 0) && (val < -thresh)) state = -0.5f;      /* :92-94 */
            else if ((state < 0) && (val > thresh)) state = 0.5f;    /* :95-97 */
        }
        state_v[n] = state;
    }
}

void orc_square_grain_mix_run(float *state_v, const float *threshold,
                              uint32_t *phase, const uint32_t *inc,
                              const uint8_t *gl, const uint8_t *gr, uint64_t N,
                              uint64_t F, int32_t *imix, float *mix) {
    for (uint64_t t = 0; t < 2 * F; t++) imix[t] = 0;
    for (uint64_t n = 0; n < N; n++) {
        float state = state_v[n], thresh = threshold[n];
        uint32_t ph = phase[n];
        for (uint64_t i = 0; i < F; i++) {
            float val = (float)(int32_t)ph * (1.0f / 2147483648.0f); /* acc -> signed saw */
            ph += inc[n];                                            /* cproc.h:141 */
            /* out = state in {0,+-0.5}: in units of 2^-7 with gain k/64 ->
             * state*2 in {0,+-1} times k */
            int32_t s2 = state > 0 ? 1 : (state < 0 ? -1 : 0);
            imix[i]     = (int32_t)((uint32_t)imix[i]     + (uint32_t)(s2 * (int32_t)gl[n]));
            imix[F + i] = (int32_t)((uint32_t)imix[F + i] + (uint32_t)(s2 * (int32_t)gr[n]));
            if ((state >= 0) && (val < -thresh)) state = -0.5f;
            else if ((state < 0) && (val > thresh)) state = 0.5f;
        }
        state_v[n] = state; phase[n] = ph;
    }
    if (mix) for (uint64_t t = 0; t < 2 * F; t++) mix[t] = (float)imix[t] * (1.0f / 128.0f);
}

/* ======================================================================= */
/* Extension processors: the definition (no reference counterpart).         */

float orc_xvoice_tick(orc_xvoice_state *s, const orc_xvoice_param *p) {
    /* phasor_f: acc (cproc.h:141) read before update like synth.c:175-177 */
    float x = (float)(int32_t)s->phase * (1.0f / 2147483648.0f);
    s->phase += p->inc;
    /* Chamberlin SVF, low-pass out; each line is ONE rounding (fmaf). */
    float lp = fmaf(p->f, s->bp, s->lp);
    float hp = x - lp;
    hp = fmaf(-p->q, s->bp, hp);
    float bp = fmaf(p->f, hp, s->bp);
    s->lp = lp; s->bp = bp;
    /* linear attack/release envelope */
    float e = s->env;
    if (s->t < p->gate_frames) { e = e + p->env_attack; if (e > 1.0f) e = 1.0f; }
    else { e = e - p->env_release; if (e < 0.0f) e = 0.0f; }
    s->env = e;
    s->t += 1;
    return lp * e;
}

void orc_xvoice_run(orc_xvoice_state *s, const orc_xvoice_param *p, uint64_t N,
                    uint64_t F, float *raw, float *mix) {
    double *acc = NULL;
    if (mix) { acc = (double *)__builtin_malloc(sizeof(double) * 2 * F); memset(acc, 0, sizeof(double) * 2 * F); }
    for (uint64_t n = 0; n < N; n++) {
        for (uint64_t t = 0; t < F; t++) {
            float y = orc_xvoice_tick(&s[n], &p[n]);
            float l = p[n].gl * y, r = p[n].gr * y;
            if (raw) { raw[((uint64_t)n * F + t) * 2] = l; raw[((uint64_t)n * F + t) * 2 + 1] = r; }
            if (acc) { acc[t] += l; acc[F + t] += r; }
        }
    }
    if (mix) { for (uint64_t t = 0; t < 2 * F; t++) mix[t] = (float)acc[t]; __builtin_free(acc); }
}

void orc_onepole_run(float *y, const float *a, uint64_t N, uint64_t F,
                     const float *in, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        float s = y[n];
        for (uint64_t t = 0; t < F; t++) {
            float x = in[(uint64_t)n * F + t];
            s = fmaf(a[n], x - s, s);
            out[(uint64_t)n * F + t] = s;
        }
        y[n] = s;
    }
}

/* linux/clock.c:109-120 (clock_phase, clock_pol, clock_hperiod :42,61-62) */
void orc_word_clock_run(int32_t *state, const int32_t *hperiod, uint64_t N, uint64_t F, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        int32_t phase = state[2 * n], pol = state[2 * n + 1];
        const int32_t h = hperiod[n];
        for (uint64_t t = 0; t < F; t++) {
            if (phase >= h) { phase = (int32_t)((uint32_t)phase - (uint32_t)h); pol ^= 1; }
            out[(uint64_t)n * F + t] = (float)pol;
            phase = (int32_t)((uint32_t)phase + 1u);
        }
        state[2 * n] = phase; state[2 * n + 1] = pol;
    }
}
