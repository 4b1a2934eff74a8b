/* ref_synth_tail.c -- appended (same translation unit) to lines 29-206 of the
 * reference linux/synth.c, which are piped in by oracle/build_ref.sh.  Exposes
 * the reference voice bank through a flat C interface. */
#include <unistd.h>
#include <fcntl.h>

uint32_t ref_synth_sizeof(int what) { return what == 0 ? sizeof(struct voice) : sizeof(struct synth); }

static int ref_mute_stderr(void) {
    fflush(stderr);
    int keep = dup(2), nul = open("/dev/null", O_WRONLY);
    dup2(nul, 2); close(nul);
    return keep;
}
static void ref_unmute_stderr(int keep) { fflush(stderr); dup2(keep, 2); close(keep); }

/* note_to_inc LOGs to stderr (synth.c:123); silence it while building the table */
void ref_note_table(uint32_t *inc128) {
    int keep = ref_mute_stderr();
    for (int n = 0; n < 128; n++) inc128[n] = note_to_inc(n);
    ref_unmute_stderr(keep);
}
void ref_note_tab12(uint32_t *t12) { for (int i = 0; i < 12; i++) t12[i] = note_tab[i]; }

/* voices: [n_synth][64]{inc,state}; vec [n_synth][F].  mode 0 = synth_run
 * (saw, :196-202), 1 = per-frame sum_tick_square (:182-195). */
void ref_voice_bank_run(uint32_t *voices, uint64_t n_synth, int mode, uint64_t F, float *vec) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < (int64_t)n_synth; s++) {
        struct synth x;
        synth_init(&x);
        for (int v = 0; v < 64; v++) { x.voice[v].note_inc = voices[(s * 64 + v) * 2]; x.voice[v].note_state = voices[(s * 64 + v) * 2 + 1]; }
        if (mode == 0) synth_run(&x, vec + s * F, (int)F);
        else for (uint64_t i = 0; i < F; i++) vec[s * F + i] = sum_tick_square(&x);
        for (int v = 0; v < 64; v++) { voices[(s * 64 + v) * 2] = x.voice[v].note_inc; voices[(s * 64 + v) * 2 + 1] = x.voice[v].note_state; }
    }
}
/* Drive the reference through its own note API: notes[] >= 0 note_on, < 0
 * note_off(-n-1); then render F frames.  Returns the voice table. */
void ref_synth_play(const int *notes, int n_notes, uint64_t F, float *vec, uint32_t *voices_out) {
    struct synth x;
    int keep = ref_mute_stderr();
    synth_init(&x);
    for (int i = 0; i < n_notes; i++) { if (notes[i] >= 0) synth_note_on(&x, notes[i]); else synth_note_off(&x, -notes[i] - 1); }
    ref_unmute_stderr(keep);
    synth_run(&x, vec, (int)F);
    for (int v = 0; v < 64; v++) { voices_out[v * 2] = x.voice[v].note_inc; voices_out[v * 2 + 1] = x.voice[v].note_state; }
}
