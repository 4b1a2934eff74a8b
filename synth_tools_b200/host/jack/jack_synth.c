/* jack_synth.c -- a JACK client that renders the reference's phasor synth (linux/synth.c) on the
 * B200: SURVEY 8 f-4, the host adapter around cproc_cuda.h.
 *
 * Where the reference runs ONE 64-voice synth on MIDI channel 0 with the loop body inside the JACK
 * process callback (synth.c:208-312), this client runs one 64-voice synth per MIDI channel (16
 * parts, 1,024 voices) and renders all of them in ONE batched call per period:
 *   - MIDI in: note on / note off per channel, voice allocation as synth.c:143-165 (first free
 *     voice, else voice 0; note2voice per part);
 *   - audio out: one port per part (part_00 .. part_15), planar float as JACK hands it over;
 *   - per period: upload the {note_inc, note_state} records of all parts (the reference's
 *     struct voice, 8 bytes), render nframes of every part (CPROC_CUDA_VOICE_BANK, 64 voices per
 *     bus), download the phases.  The order "MIDI first, then audio" is the reference's (:291-296).
 * Exits when stdin reaches EOF, like the reference (:305-310).
 *
 * Build (on a machine with JACK):  gcc -std=gnu99 -O2 jack_synth.c -I../../../include \
 *     -L../.. -lcproc_cuda -ljack -Wl,-rpath,'$ORIGIN/../..' -o jack_synth_b200
 * tests/test_gpu_dropin.py builds it against tests/c/fakejack (a scripted JACK stand-in). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <jack/jack.h>
#include <jack/midiport.h>
#include "cproc_cuda.h"
#include "note_table.h"

#define N_PARTS 16
#define N_VOICES 64

struct voice { uint32_t note_inc, note_state; };                         /* synth.c:33-36 */
struct part { int note2voice[128]; struct voice voice[N_VOICES]; };      /* struct synth, synth.c:37-40 */

static struct part parts[N_PARTS];
static jack_client_t *client;
static jack_port_t *midi_in, *audio_out[N_PARTS];
static cproc_cuda_ctx *ctx;
static cproc_cuda_batch *bank;
static float *mixbuf;                     /* [N_PARTS][max_frames] */
static jack_nframes_t max_frames;
static struct voice flat[N_PARTS * N_VOICES];
static int failed;

static int voice_alloc(struct part *p) {                                  /* synth.c:143-152 */
    for (int v = 0; v < N_VOICES; v++) if (p->voice[v].note_inc == 0) return v;
    return 0;
}
static void note_on(struct part *p, int note) {                           /* :154-158 */
    const int v = voice_alloc(p);
    p->note2voice[note & 127] = v;
    p->voice[v].note_inc = cproc_note_to_inc(note & 127);
}
static void note_off(struct part *p, int note) {                          /* :159-163 */
    const int v = p->note2voice[note & 127];
    p->note2voice[note & 127] = 0;
    p->voice[v].note_inc = 0;
}

static void process_midi(jack_nframes_t nframes) {
    void *buf = jack_port_get_buffer(midi_in, nframes);
    const jack_nframes_t n = jack_midi_get_event_count(buf);
    for (jack_nframes_t i = 0; i < n; i++) {
        jack_midi_event_t ev;
        if (jack_midi_event_get(&ev, buf, i) || ev.size != 3) continue;
        const uint8_t *m = ev.buffer;
        struct part *p = &parts[m[0] & 0x0F];
        if ((m[0] & 0xF0) == 0x90) { if (m[2] == 0) note_off(p, m[1]); else note_on(p, m[1]); }   /* :247-256 */
        else if ((m[0] & 0xF0) == 0x80) note_off(p, m[1]);                                        /* :258-262 */
    }
}

static void process_audio(jack_nframes_t nframes) {
    float *dst[N_PARTS];
    for (int k = 0; k < N_PARTS; k++) {
        dst[k] = (float *)jack_port_get_buffer(audio_out[k], nframes);
        memset(dst[k], 0, sizeof(float) * nframes);
    }
    if (failed || nframes == 0 || nframes > max_frames) return;
    for (int k = 0; k < N_PARTS; k++) memcpy(flat + k * N_VOICES, parts[k].voice, sizeof(parts[k].voice));
    cproc_cuda_io io;
    memset(&io, 0, sizeof(io));
    io.out = mixbuf; io.layout = CPROC_CUDA_PLANAR;                       /* float [part][nframes] */
    /* voice records in, render, records out: one stream sequence, one synchronisation, no allocation (cproc_cuda_run_period) */
    int rc = cproc_cuda_run_period(bank, nframes, &io, flat, sizeof(struct voice));
    if (rc) { failed = rc; fprintf(stderr, "jack_synth_b200: render failed (%d): %s\n", rc, cproc_cuda_last_error(ctx)); return; }
    for (int k = 0; k < N_PARTS; k++) {
        memcpy(parts[k].voice, flat + k * N_VOICES, sizeof(parts[k].voice));
        memcpy(dst[k], mixbuf + (size_t)k * nframes, sizeof(float) * nframes);
    }
}

static int process(jack_nframes_t nframes, void *arg) {
    (void)arg;
    process_midi(nframes);                /* order is important (synth.c:293) */
    process_audio(nframes);
    return 0;
}

int main(void) {
    jack_status_t status = 0;
    client = jack_client_open("synth_b200", JackNullOption, &status);
    if (!client) { fprintf(stderr, "jack_synth_b200: no JACK server\n"); return 1; }
    if (cproc_cuda_open(0, NULL, &ctx)) { fprintf(stderr, "jack_synth_b200: %s\n", cproc_cuda_last_error(NULL)); return 1; }
    cproc_cuda_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.proc = CPROC_CUDA_VOICE_BANK; cfg.layout = CPROC_CUDA_PLANAR;
    cfg.mode = CPROC_CUDA_MIX_SAW; cfg.voices_per_bus = N_VOICES;         /* synth_run -> sum_tick_saw, synth.c:196-202 */
    if (cproc_cuda_alloc(ctx, &cfg, N_PARTS * N_VOICES, &bank)) { fprintf(stderr, "jack_synth_b200: %s\n", cproc_cuda_last_error(ctx)); return 1; }
    max_frames = jack_get_buffer_size(client);
    if (max_frames < 1024) max_frames = 1024;
    mixbuf = (float *)malloc(sizeof(float) * N_PARTS * max_frames);
    midi_in = jack_port_register(client, "midi_in", JACK_DEFAULT_MIDI_TYPE, JackPortIsInput, 0);
    for (int k = 0; k < N_PARTS; k++) {
        char name[32];
        snprintf(name, sizeof(name), "part_%02d", k);
        audio_out[k] = jack_port_register(client, name, JACK_DEFAULT_AUDIO_TYPE, JackPortIsOutput, 0);
    }
    /* Warm the period path before the RT callback exists: the first call with a shape allocates the staging buffers (the pinned record staging included), the second
       captures the period's CUDA graph (cproc_cuda_run, run_graph), and neither belongs on the JACK thread.  The voice table is
       silent (note_inc == 0 everywhere), so the warm-up leaves every phase where it was. */
    {
        const jack_nframes_t period = jack_get_buffer_size(client);
        cproc_cuda_io io;
        memset(&io, 0, sizeof(io));
        io.out = mixbuf; io.layout = CPROC_CUDA_PLANAR;
        memset(flat, 0, sizeof(flat));
        for (int k = 0; k < 3 && period && period <= max_frames; k++) {
            int rc = cproc_cuda_run_period(bank, period, &io, flat, sizeof(struct voice));
            if (rc) { fprintf(stderr, "jack_synth_b200: warm-up failed (%d): %s\n", rc, cproc_cuda_last_error(ctx)); return 1; }
        }
    }
    jack_set_process_callback(client, process, 0);
    if (jack_activate(client)) { fprintf(stderr, "jack_synth_b200: cannot activate\n"); return 1; }
    for (;;) {                            /* stdin is only used to signal exit (synth.c:305-310) */
        uint8_t b[4];
        if (read(0, b, sizeof(b)) <= 0) break;
    }
    jack_client_close(client);
    cproc_cuda_free(bank);
    cproc_cuda_close(ctx);
    free(mixbuf);
    return failed ? 1 : 0;
}
