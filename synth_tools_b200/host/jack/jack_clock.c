/* jack_clock.c -- the reference's JACK word-clock / MIDI-clock master (linux/clock.c) with the
 * square-wave loop (clock.c:109-120) rendered on the B200: SURVEY 8 f-4.
 *
 * The reference runs ONE clock; this client runs one clock per BPM given on the command line
 * (default: 120) and renders all of them with one batched call per period (CPROC_CUDA_WORD_CLOCK):
 *   - audio out: one port per clock (clock_00 ..), the integer-divisor square wave, planar float;
 *   - MIDI out: start / continue / stop received on midi_in go out first at time 0 (clock.c:83-95:
 *     JACK wants events sorted), then a MIDI clock byte 0xF8 at every sample where clock 0 turns
 *     positive (clock.c:113-116), found by reading the rendered block;
 *   - half period = sr*5 / (bpm*4) samples (BPM_TO_HPERIOD, clock.c:58), latched on the first period
 *     because the sample rate is only valid inside the process thread (clock.c:99-106);
 *   - initial state: phase 0, polarity 1 (clock.c:61-62).
 * The Erlang command queue of the reference (clock.c:73-82, 125-160) is control plane and not part
 * of this adapter.  Exits when stdin reaches EOF.
 *
 * Build (on a machine with JACK):  gcc -std=gnu99 -O2 jack_clock.c -I../../../include \
 *     -L../.. -lcproc_cuda -ljack -Wl,-rpath,'$ORIGIN/../..' -o jack_clock_b200
 * tests/test_gpu_dropin.py builds it against tests/c/fakejack (a scripted JACK stand-in). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <jack/jack.h>
#include <jack/midiport.h>
#include "cproc_cuda.h"

#define MAX_CLOCKS 64
#define BPM_TO_HPERIOD(sr, bpm) (((sr) * 5) / ((bpm) * 4))            /* clock.c:58 */

struct clock_state { int32_t phase, pol; };                            /* clock_phase, clock_pol (clock.c:61-62) */

static jack_client_t *client;
static jack_port_t *midi_in, *midi_out, *audio_out[MAX_CLOCKS];
static cproc_cuda_ctx *ctx;
static cproc_cuda_batch *bank;
static int n_clocks;
static jack_nframes_t bpm[MAX_CLOCKS];
static int32_t hperiod[MAX_CLOCKS];
static int hperiod_set;
static int32_t last_pol0 = 1;
static float *block;                      /* [n_clocks][max_frames] */
static jack_nframes_t max_frames;
static int failed;

static void send_midi(void *out_buf, jack_nframes_t time, const uint8_t *data, size_t n) {   /* clock.c:45-56 */
    jack_midi_data_t *buf = jack_midi_event_reserve(out_buf, time, n);
    if (buf) memcpy(buf, data, n);
}

static int process(jack_nframes_t nframes, void *arg) {
    (void)arg;
    void *midi_out_buf = jack_port_get_buffer(midi_out, nframes);
    jack_midi_clear_buffer(midi_out_buf);
    void *midi_in_buf = jack_port_get_buffer(midi_in, nframes);
    const jack_nframes_t n_ev = jack_midi_get_event_count(midi_in_buf);
    for (jack_nframes_t i = 0; i < n_ev; i++) {                          /* clock.c:83-95 */
        jack_midi_event_t ev;
        if (jack_midi_event_get(&ev, midi_in_buf, i) || ev.size != 1) continue;
        if (ev.buffer[0] == 0xFA || ev.buffer[0] == 0xFB || ev.buffer[0] == 0xFC) send_midi(midi_out_buf, 0, ev.buffer, 1);
    }
    float *dst[MAX_CLOCKS];
    for (int k = 0; k < n_clocks; k++) {
        dst[k] = (float *)jack_port_get_buffer(audio_out[k], nframes);
        memset(dst[k], 0, sizeof(float) * nframes);
    }
    if (failed || nframes == 0 || nframes > max_frames) return 0;
    if (!hperiod_set) {                                                  /* clock.c:99-106 */
        const jack_nframes_t sr = jack_get_sample_rate(client);
        for (int k = 0; k < n_clocks; k++) {
            hperiod[k] = (int32_t)BPM_TO_HPERIOD(sr, bpm[k]);
            fprintf(stderr, "clock %d: bpm_set = %u, sr = %u -> clock_hperiod = %d, bpm_actual = %3.6f\n", k, (unsigned)bpm[k], (unsigned)sr,
                    (int)hperiod[k], (double)((((float)sr) * 1.25f) / ((float)hperiod[k])));
        }
        if (cproc_cuda_upload_param(bank, hperiod, sizeof(int32_t))) failed = 1;
        hperiod_set = 1;
    }
    cproc_cuda_io io;
    memset(&io, 0, sizeof(io));
    io.out = block; io.layout = CPROC_CUDA_PLANAR;
    const int rc = failed ? failed : cproc_cuda_run(bank, nframes, &io);
    if (rc) { failed = rc; fprintf(stderr, "jack_clock_b200: render failed (%d): %s\n", rc, cproc_cuda_last_error(ctx)); return 0; }
    for (int k = 0; k < n_clocks; k++) memcpy(dst[k], block + (size_t)k * nframes, sizeof(float) * nframes);
    for (jack_nframes_t t = 0; t < nframes; t++) {                       /* positive edges of clock 0 (clock.c:112-116) */
        const int32_t pol = (int32_t)dst[0][t];
        if (pol == 1 && last_pol0 != 1) { const uint8_t clk[] = {0xF8}; send_midi(midi_out_buf, t, clk, sizeof(clk)); }
        last_pol0 = pol;
    }
    return 0;
}

int main(int argc, char **argv) {
    n_clocks = argc > 1 ? argc - 1 : 1;
    if (n_clocks > MAX_CLOCKS) n_clocks = MAX_CLOCKS;
    for (int k = 0; k < n_clocks; k++) {
        bpm[k] = argc > 1 ? (jack_nframes_t)strtoul(argv[1 + k], NULL, 0) : 120;          /* clock.c:60 */
        if (bpm[k] == 0) { fprintf(stderr, "jack_clock_b200: bpm must be > 0\n"); return 1; }
    }
    jack_status_t status = 0;
    client = jack_client_open("clock_b200", JackNullOption, &status);
    if (!client) { fprintf(stderr, "jack_clock_b200: no JACK server\n"); return 1; }
    if (cproc_cuda_open(0, NULL, &ctx)) { fprintf(stderr, "jack_clock_b200: %s\n", cproc_cuda_last_error(NULL)); return 1; }
    cproc_cuda_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.proc = CPROC_CUDA_WORD_CLOCK; cfg.layout = CPROC_CUDA_PLANAR;
    if (cproc_cuda_alloc(ctx, &cfg, (uint64_t)n_clocks, &bank)) { fprintf(stderr, "jack_clock_b200: %s\n", cproc_cuda_last_error(ctx)); return 1; }
    struct clock_state init[MAX_CLOCKS];
    for (int k = 0; k < n_clocks; k++) { init[k].phase = 0; init[k].pol = 1; }
    if (cproc_cuda_upload_state(bank, init, sizeof(struct clock_state))) { fprintf(stderr, "jack_clock_b200: %s\n", cproc_cuda_last_error(ctx)); return 1; }
    max_frames = jack_get_buffer_size(client);
    if (max_frames < 1024) max_frames = 1024;
    block = (float *)malloc(sizeof(float) * (size_t)n_clocks * max_frames);
    midi_in = jack_port_register(client, "midi_in", JACK_DEFAULT_MIDI_TYPE, JackPortIsInput, 0);
    midi_out = jack_port_register(client, "midi_out", JACK_DEFAULT_MIDI_TYPE, JackPortIsOutput, 0);
    for (int k = 0; k < n_clocks; k++) {
        char name[32];
        snprintf(name, sizeof(name), "clock_%02d", k);
        audio_out[k] = jack_port_register(client, name, JACK_DEFAULT_AUDIO_TYPE, JackPortIsOutput, 0);
    }
    jack_set_process_callback(client, process, 0);
    if (jack_activate(client)) { fprintf(stderr, "jack_clock_b200: cannot activate\n"); return 1; }
    for (;;) {
        uint8_t b[4];
        if (read(0, b, sizeof(b)) <= 0) break;
    }
    jack_client_close(client);
    cproc_cuda_free(bank);
    cproc_cuda_close(ctx);
    free(block);
    return failed ? 1 : 0;
}
