/* ref_ext_voice.c -- the graph TEXTS under tests/golden/*.cproc compiled as C, unchanged, against the reference's own
 * generic/cproc.h (PROC / PROC_COND, cproc.h:72-81) plus include/cproc_ext.h: what a CPU host of the reference would
 * build from the same text the GPU front end parses.  Node state is function-static (cproc.h:73): ONE instance per
 * loaded copy of this library, ticked from zero state.  Test infrastructure; built only into oracle/_ref. */
#include <stdint.h>
#include <string.h>
#include "cproc.h"
#include "cproc_ext.h"

static uint32_t cap[8];
void cproc_output(uint32_t index, w value) { if (index < 8) cap[index] = value; }
void cproc_output_f(uint32_t index, float value) { if (index < 8) memcpy(&cap[index], &value, 4); }

#define cproc_update ref_ext_voice_update
#include "ext_voice.cproc"
#undef cproc_update
#undef CPROC_NB_INPUTS
#define cproc_update ref_ext_chain_update
#include "ext_chain.cproc"
#undef cproc_update
#undef CPROC_NB_INPUTS
#define cproc_update ref_ext_gain_update
#include "ext_gain.cproc"
#undef cproc_update

/* F ticks of graph `which` (0 voice, 1 chain, 2 gain): in [n_in][F], changed [F], out [n_out][F] = the captured upcall values by
 * index (voice: 0, 1; chain: 2, 3, 4; gain: 0). */
void ref_ext_text_run(int which, const uint32_t *in, const uint32_t *changed, uint64_t F, uint32_t *out) {
    for (uint64_t t = 0; t < F; t++) {
        if (which == 0) {
            w input[1] = { in[t] };
            ref_ext_voice_update(input, changed ? changed[t] : (w)-1);
            out[t] = cap[0]; out[F + t] = cap[1];
        } else if (which == 2) {
            w input[1] = { in[t] };
            ref_ext_gain_update(input, changed ? changed[t] : (w)-1);
            out[t] = cap[0];
        } else {
            w input[2] = { in[t], in[F + t] };
            ref_ext_chain_update(input, changed ? changed[t] : (w)-1);
            out[t] = cap[2]; out[F + t] = cap[3]; out[2 * F + t] = cap[4];
        }
    }
}
