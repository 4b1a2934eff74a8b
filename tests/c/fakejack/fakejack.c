/* fakejack.c -- scripted JACK stand-in (test infrastructure).  jack_activate() plays a fixed script
 * through the client's process callback: PERIODS periods of FRAMES frames with MIDI events at
 * chosen periods, and prints every audio port's samples as hex floats and every MIDI-out event:
 *     audio <period> <port name> <f0> <f1> ...
 *     midi <period> <port name> <time> <byte> ...
 * FAKEJACK_SCRIPT=clock selects the realtime-message script (start / stop / a stray clock byte),
 * FAKEJACK_PERIODS overrides the number of periods.
 * tests/test_gpu_dropin.py replays the same scripts through the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jack/jack.h"
#include "jack/midiport.h"

#define FRAMES 64
#define PERIODS 6
struct fake_port { char name[32]; int is_midi, is_out; float audio[1024]; int n_out; struct { jack_nframes_t time; size_t size; unsigned char b[8]; } out[64]; };
struct fake_client { JackProcessCallback cb; void *arg; struct fake_port ports[32]; int n_ports; };
static struct fake_client the_client;

/* the script: {period, status, data1, data2} */
static const unsigned char script[][4] = {
    {0, 0x90, 69, 100}, {0, 0x90, 60, 100}, {0, 0x93, 127, 1}, {0, 0x9F, 0, 64},
    {1, 0x90, 72, 90},  {2, 0x80, 60, 0},   {2, 0x93, 127, 0}, {3, 0x9F, 12, 80},
    {4, 0x90, 69, 0},   {4, 0x91, 40, 127}, {5, 0x8F, 0, 0},
};
/* realtime messages for jack_clock.c: {period, status}; only 0xFA / 0xFB / 0xFC may pass */
static const unsigned char script_clock[][4] = { {1, 0xFA, 0, 0}, {3, 0xF8, 0, 0}, {3, 0xFC, 0, 0}, {5, 0xFB, 0, 0}, {5, 0xFE, 0, 0} };
static unsigned char ev_bytes[16][3];
static int ev_count;
static int ev_size = 3;

jack_client_t *jack_client_open(const char *name, jack_options_t options, jack_status_t *status, ...) {
    (void)name; (void)options; if (status) *status = 0;
    memset(&the_client, 0, sizeof(the_client));
    return &the_client;
}
int jack_client_close(jack_client_t *c) { (void)c; return 0; }
jack_nframes_t jack_get_buffer_size(jack_client_t *c) { (void)c; return FRAMES; }
jack_nframes_t jack_get_sample_rate(jack_client_t *c) { (void)c; return 48000; }
void jack_midi_clear_buffer(void *port_buffer) { ((struct fake_port *)port_buffer)->n_out = 0; }
jack_midi_data_t *jack_midi_event_reserve(void *port_buffer, jack_nframes_t time, size_t data_size) {
    struct fake_port *p = (struct fake_port *)port_buffer;
    if (p->n_out >= 64 || data_size > 8) return NULL;
    p->out[p->n_out].time = time; p->out[p->n_out].size = data_size;
    return p->out[p->n_out++].b;
}
jack_port_t *jack_port_register(jack_client_t *c, const char *name, const char *type, unsigned long flags, unsigned long bufsize) {
    (void)bufsize;
    struct fake_port *p = &c->ports[c->n_ports++];
    snprintf(p->name, sizeof(p->name), "%s", name);
    p->is_midi = strstr(type, "midi") != NULL; p->is_out = (flags & JackPortIsOutput) != 0;
    return p;
}
void *jack_port_get_buffer(jack_port_t *p, jack_nframes_t nframes) { (void)nframes; return p->is_midi ? (void *)p : (void *)p->audio; }
int jack_set_process_callback(jack_client_t *c, JackProcessCallback cb, void *arg) { c->cb = cb; c->arg = arg; return 0; }
jack_nframes_t jack_midi_get_event_count(void *port_buffer) { (void)port_buffer; return (jack_nframes_t)ev_count; }
int jack_midi_event_get(jack_midi_event_t *event, void *port_buffer, uint32_t i) {
    (void)port_buffer;
    if ((int)i >= ev_count) return -1;
    event->time = 0; event->size = (size_t)ev_size; event->buffer = ev_bytes[i];
    return 0;
}
int jack_activate(jack_client_t *c) {
    const char *which = getenv("FAKEJACK_SCRIPT");
    const int clock = which && !strcmp(which, "clock");
    const unsigned char (*sc)[4] = clock ? script_clock : script;
    const size_t n_sc = clock ? sizeof(script_clock) / sizeof(script_clock[0]) : sizeof(script) / sizeof(script[0]);
    const int periods = getenv("FAKEJACK_PERIODS") ? atoi(getenv("FAKEJACK_PERIODS")) : PERIODS;
    ev_size = clock ? 1 : 3;
    for (int period = 0; period < periods; period++) {
        ev_count = 0;
        for (size_t k = 0; k < n_sc; k++)
            if (sc[k][0] == period) { memcpy(ev_bytes[ev_count], &sc[k][1], 3); ev_count++; }
        if (c->cb(FRAMES, c->arg)) return -1;
        for (int q = 0; q < c->n_ports; q++) {
            struct fake_port *p = &c->ports[q];
            if (!p->is_out || !p->is_midi) continue;
            for (int e = 0; e < p->n_out; e++) {
                printf("midi %d %s %u", period, p->name, (unsigned)p->out[e].time);
                for (size_t k = 0; k < p->out[e].size; k++) printf(" %02x", p->out[e].b[k]);
                printf("\n");
            }
        }
        for (int q = 0; q < c->n_ports; q++) {
            struct fake_port *p = &c->ports[q];
            if (!p->is_out || p->is_midi) continue;
            printf("audio %d %s", period, p->name);
            for (int t = 0; t < FRAMES; t++) printf(" %a", p->audio[t]);
            printf("\n");
        }
    }
    fflush(stdout);
    return 0;
}
