/* cproc_oracle.h -- CPU restatement of the synth_tools per-sample DSP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Every function restates one reference function and cites it
 * (paths relative to the synth_tools tree).  The restatement is pinned
 * against the reference's own sources compiled unmodified (oracle/_ref,
 * built by oracle/build_ref.sh) in tests/test_oracle_vs_ref.py, and against
 * the committed vectors under tests/golden/.
 *
 * Parity status:
 *   - acc, edge, edge->acc graph, pdm1..4, voice bank, note table,
 *     square_grain: PINNED (reference source compiled here).
 *   - v2 channel (ISR body, glide, PDM_COPY_LINE, pdm_update_line, control
 *     divider): PINNED -- mod_pdm_pwm.c and mod_controlrate.c are compiled
 *     whole against a hosted stand-in for the hardware layer and the timer
 *     ISR is called per tick (oracle/ref/ref_v2_isr.c).
 *   - word clock (linux/clock.c:108-120): PINNED (the loop is piped from the
 *     reference file into the harness).
 *   - pwm_update (mod_pdm.c:159-175): PINNED (lines piped from the reference).
 *   - v1 carry-bit channel (ARM inline asm, no ARM toolchain here): restated
 *     arithmetic, checked against an independent 64-bit formulation.
 *   - dither PRNG random_u32(): PARITY UNPINNED.  uc_tools xorshift.h
 *     (github:zwizwa/uc_tools rev c0853b29811c5d184d39b630b3c848a86d5d5e9e)
 *     is not vendored in the reference tree.  We restate Marsaglia's
 *     xorshift32 (13,17,5) and additionally accept an external dither
 *     stream so that PDM parity never depends on the PRNG choice.
 *   - onepole, svf, env, phasor_f extension processors: not in the
 *     reference at all; this file is their definition.
 */
#ifndef CPROC_ORACLE_H
#define CPROC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- generic/cproc.h:128-154 ------------------------------------------ */
typedef uint32_t orc_w;
typedef struct { orc_w out; } orc_acc_state;            /* cproc.h:134 */
typedef struct { orc_w out; orc_w last; } orc_edge_state; /* cproc.h:145 */
void orc_acc_update(orc_acc_state *s, orc_w in);        /* cproc.h:140-142 */
void orc_edge_update(orc_edge_state *s, orc_w in);      /* cproc.h:151-154 */

/* A generated cproc graph (linux/test_cproc.c:12-17, stm32f103/bp5_plugin.c:4-9)
 * as a table: one row per PROC_COND statement, in ANF order. */
enum { ORC_NODE_ACC = 0, ORC_NODE_EDGE = 1, ORC_NODE_GLIDE = 2, ORC_NODE_PDM = 3,
       /* extension processors (include/cproc_ext.h is the definition; restated in cproc_oracle.c, parity unpinned by the reference) */
       ORC_NODE_PHASOR_F = 4, ORC_NODE_SVF = 5, ORC_NODE_ENV = 6, ORC_NODE_ONEPOLE = 7, ORC_NODE_GAIN = 8, ORC_NODE_ASFLOAT = 9, ORC_NODE_GLIDE_F = 10, ORC_NODE_MUL = 11, ORC_NODE_KINDS = 12 };
#define ORC_SRC_ZERO ((int32_t)0x80000000)   /* an input the PROC statement does not name: 0 (C initialiser) */
/* pdm node: pdmK_update (pdm.h:13-77) as a processor, state {out, s1..sK}; .in = src, .dither = src2;
 * type = 3 | (K | out_shift << 3) << 8. */
/* glide: the control-rate -> audio-rate parameter interpolation of the firmware
 * (doc/combinators.org:28-34 "representative example") as a processor.  State is the
 * two line segments of struct channel (mod_pdm_pwm.c:80-93) plus the divider count
 * (:78); .in is the control-rate value (the setpoint), read once per 2^L ticks;
 * .out is line[0].position.  One tick = the ISR order of mod_pdm_pwm.c:129-143:
 *   if (count == 0) { line[0] = line[1];                      (:108-109, :133)
 *                     line[1].position += line[1].velocity << L;     (mod_controlrate.c:32)
 *                     line[1].velocity = (int32)(in - line[1].position) >> L; }  (:33-34)
 *   line[0].position += line[0].velocity;                     (:97-100)
 *   count = (count + 1) % 2^L                                 (:141)
 * L = bits 8..15 of the node type. */
typedef struct { orc_w out; orc_w vel0; orc_w pos1; orc_w vel1; orc_w count; } orc_glide_state;
void orc_glide_update(orc_glide_state *s, orc_w in, uint32_t div_log);
typedef struct {
    uint32_t type;      /* ORC_NODE_* | (argument << 8) */
    int32_t  src;       /* >=0: .in = n<src>.out ; <0: .in = input[-(src+1)] */
    uint32_t cond_mask; /* executed iff (changed & cond_mask) != 0 */
    int32_t  src2;      /* second input (pdm: dither), same encoding */
} orc_node;
uint32_t orc_node_state_words(uint32_t type);
uint32_t orc_graph_state_words(const orc_node *nodes, uint32_t n_nodes);
/* state [N][state_words]; in [N][n_inputs][F]; changed [N][F] or NULL (= -1,
 * mod_cproc_plugin.c:32); out [N][F] = value handed to cproc_output() at
 * each tick (test_cproc.c:16), i.e. nodes[out_node].out after the tick. */
void orc_graph_run(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                   uint32_t out_node, uint32_t *state, uint64_t N, uint64_t F,
                   const uint32_t *in, const uint32_t *changed, uint32_t *out);

/* Several cproc_output() statements: out [N][n_out][F], out_nodes[q] = node behind output q. */
void orc_graph_run_multi(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                         const uint32_t *out_nodes, uint32_t n_out, uint32_t *state, uint64_t N, uint64_t F,
                         const uint32_t *in, const uint32_t *changed, uint32_t *out);

/* Graphs with extension processors: param [N][param_words] = the nodes' param structs in ANF order (NULL if none);
 * n_inputs may be 0 (in == NULL).  Float fields travel as their bit patterns. */
uint32_t orc_node_param_words(uint32_t type);
uint32_t orc_graph_param_words(const orc_node *nodes, uint32_t n_nodes);
void orc_graph_run_ext(const orc_node *nodes, uint32_t n_nodes, uint32_t n_inputs,
                       const uint32_t *out_nodes, uint32_t n_out, uint32_t *state, const uint32_t *param, uint64_t N, uint64_t F,
                       const uint32_t *in, const uint32_t *changed, uint32_t *out);

/* ---- stm32f103/pdm.h:10-77 -------------------------------------------- */
/* s[0..order-1] = s1..sK.  order 1 ignores dither (pdm.h:13). */
uint32_t orc_pdm_update(uint32_t *s, uint32_t order, uint32_t input,
                        uint32_t out_shift, uint32_t dither);
/* state [N][order]; in [N][F] or NULL -> in_const [N]; dither [F] or NULL;
 * out [N][F] (full uint32 quantiser output). */
void orc_pdm_run(uint32_t order, uint32_t *state, uint64_t N, uint64_t F,
                 const uint32_t *in, const uint32_t *in_const,
                 uint32_t out_shift, const uint32_t *dither, uint32_t *out);

/* ---- dither PRNG (uc_tools xorshift.h; parity unpinned) ---------------- */
uint32_t orc_xorshift32(uint32_t *state);

/* ---- stm32f103/mod_pdm.c:198-264 (v1 carry-bit PDM) --------------------- */
typedef struct { uint32_t setpoint; uint32_t accu; } orc_v1_channel; /* :198-201 */
/* One tick of one channel: returns the carry out of accu += setpoint+dither
 * (ARM adds/rrx, mod_pdm.c:230-244). */
uint32_t orc_v1_channel_update(orc_v1_channel *c, uint32_t dither);
/* Channels [N] grouped in banks of bank_size (last bank may be short); each
 * bank draws ONE dither word per tick for all its channels (mod_pdm.c:261):
 *   dither = (dither_ext ? dither_ext[bank][t] : xorshift32(&prng[bank])) & dither_mask
 * bits [N][F], one byte per sample (0/1). */
void orc_pdm_v1_run(orc_v1_channel *ch, uint64_t N, uint32_t bank_size,
                    uint32_t *prng, const uint32_t *dither_ext,
                    uint32_t dither_mask, uint64_t F, uint8_t *bits);

/* mod_pdm.c:160-175: 24-bit saw with phase>>9 feedback. */
uint32_t orc_pwm_update(uint32_t *phase, uint32_t speed);
void orc_pwm_run(uint32_t *phase, const uint32_t *speed, uint64_t N, uint64_t F,
                 uint8_t *duty);

/* ---- stm32f103/mod_pdm_pwm.c:80-143 + mod_controlrate.c:28-40 (v2) ------ */
/* Channel record as words: [setpoint, line0.position, line0.velocity,
 * line1.position, line1.velocity, s1..sK]  == struct channel (:89-93) with
 * PDM_ORDER = K. */
#define ORC_V2_WORDS(order) (5u + (order))
void orc_v2_update_line(uint32_t *chan, uint32_t ctl_div_log); /* mod_controlrate.c:28-40 */
/* Runs F ticks of the TIM_PDM ISR (mod_pdm_pwm.c:123-143) for N channels in
 * banks of bank_size sharing dither and the control divider *count.
 * At every tick where *count == 0: (optional) setpoints row is latched into
 * channel.setpoint, then line[0] = line[1] (:133), then the control-rate SWI
 * recomputes line[1] (mod_controlrate.c:46-57) -- it only touches line[1], so
 * running it before this tick's channel updates is equivalent.
 * setpoints [n_ctl][N] (row consumed per control boundary hit) or NULL.
 * duty [N][F] (low 8 bits of the quantiser output; out_shift >= 24). */
void orc_pdm_v2_run(uint32_t *chan, uint32_t order, uint64_t N,
                    uint32_t bank_size, uint32_t *prng,
                    const uint32_t *dither_ext, uint32_t dither_mask,
                    uint32_t *count, uint32_t ctl_div_log, uint32_t out_shift,
                    const uint32_t *setpoints, uint64_t F, uint8_t *duty);

/* ---- linux/synth.c:33-202 (phasor voice bank) --------------------------- */
typedef struct { uint32_t note_inc; uint32_t note_state; } orc_voice; /* :33-36 */
uint32_t orc_note_to_inc(int note);                    /* :118-125, table :94-115 */
enum { ORC_MIX_SAW = 0, ORC_MIX_SQUARE = 1 };
/* N voices, consecutive groups of voices_per_bus voices feed one bus
 * (reference: 64 voices -> 1 bus, synth.c:39).  isum [n_bus][F] raw integer
 * mix (int sum :171 / unsigned accu :184), vec [n_bus][F] float output
 * (:180 / :194).  Either output may be NULL. */
void orc_voice_bank_run(orc_voice *v, uint64_t N, uint64_t voices_per_bus,
                        int mode, uint64_t F, int32_t *isum, float *vec);

/* ---- linux/synth_tools.c:78-100 (square_grain~) ------------------------- */
/* state/threshold [N]; in/out [N][F]; in may alias out. */
void orc_square_grain_run(float *state, const float *threshold, uint64_t N,
                          uint64_t F, const float *in, float *out);
/* Config C3b: input generated by a per-grain phasor (acc + signed saw
 * (int)phase * 2^-31), dyadic pan gains gl,gr = k/64, mix in int32 units of
 * 2^-7 (mirrors synth.c:169-181 integer mix), mix [2][F] float. */
void orc_square_grain_mix_run(float *state, const float *threshold,
                              uint32_t *phase, const uint32_t *inc,
                              const uint8_t *gl, const uint8_t *gr, uint64_t N,
                              uint64_t F, int32_t *imix, float *mix);

/* ---- extension processors (NOT in the reference; defined here) ---------- */
/* All float math is single precision, one rounding per operation
 * (compile with -ffp-contract=off). */
typedef struct {          /* per-voice parameters */
    uint32_t inc;         /* phasor increment (note_to_inc) */
    float f;              /* SVF frequency coefficient 2*sin(pi*fc/fs) */
    float q;              /* SVF damping 1/Q */
    float env_attack;     /* linear attack increment per sample */
    float env_release;    /* linear release decrement per sample */
    uint32_t gate_frames; /* gate is high for frames [0,gate_frames) of the run counter */
    float gl, gr;         /* pan gains */
} orc_xvoice_param;
typedef struct {
    uint32_t phase;       /* acc */
    float lp, bp;         /* Chamberlin SVF */
    float env;            /* linear AR */
    uint32_t t;           /* frames since note-on */
} orc_xvoice_state;
/* One sample of one voice: returns the mono voice sample (before pan). */
float orc_xvoice_tick(orc_xvoice_state *s, const orc_xvoice_param *p);
/* raw: out [N][F][2] (gl*y, gr*y) or NULL; mix: [2][F] double-accumulated
 * reference mix then rounded to float, or NULL. */
void orc_xvoice_run(orc_xvoice_state *s, const orc_xvoice_param *p, uint64_t N,
                    uint64_t F, float *raw, float *mix);
/* linux/clock.c:109-120: integer-divisor word clock.  state[n] = {phase, pol}; out [N][F] = pol as
 * float.  Restated (the loop sits inside a JACK process callback); int arithmetic wraps. */
void orc_word_clock_run(int32_t *state, const int32_t *hperiod, uint64_t N, uint64_t F, float *out);
/* one-pole lowpass y += a*(x-y) as a stand-alone processor */
void orc_onepole_run(float *y, const float *a, uint64_t N, uint64_t F,
                     const float *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
