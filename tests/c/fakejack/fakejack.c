/* fakejack.c -- scripted JACK stand-in (test infrastructure).  jack_activate() plays a fixed script
 * through the client's process callback: PERIODS periods of FRAMES frames with MIDI events at
 * chosen periods, and prints every audio port's samples as hex floats:
 *     audio <period> <port name> <f0> <f1> ...
 * tests/test_gpu_dropin.py replays the same script through the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jack/jack.h"
#include "jack/midiport.h"

#define FRAMES 64
#define PERIODS 6
struct fake_port { char name[32]; int is_midi, is_out; float audio[1024]; };
struct fake_client { JackProcessCallback cb; void *arg; struct fake_port ports[32]; int n_ports; };
static struct fake_client the_client;

/* the script: {period, status, data1, data2} */
static const unsigned char script[][4] = {
    {0, 0x90, 69, 100}, {0, 0x90, 60, 100}, {0, 0x93, 127, 1}, {0, 0x9F, 0, 64},
    {1, 0x90, 72, 90},  {2, 0x80, 60, 0},   {2, 0x93, 127, 0}, {3, 0x9F, 12, 80},
    {4, 0x90, 69, 0},   {4, 0x91, 40, 127}, {5, 0x8F, 0, 0},
};
static unsigned char ev_bytes[16][3];
static int ev_count;

jack_client_t *jack_client_open(const char *name, jack_options_t options, jack_status_t *status, ...) {
    (void)name; (void)options; if (status) *status = 0;
    memset(&the_client, 0, sizeof(the_client));
    return &the_client;
}
int jack_client_close(jack_client_t *c) { (void)c; return 0; }
jack_nframes_t jack_get_buffer_size(jack_client_t *c) { (void)c; return FRAMES; }
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
    event->time = 0; event->size = 3; event->buffer = ev_bytes[i];
    return 0;
}
int jack_activate(jack_client_t *c) {
    for (int period = 0; period < PERIODS; period++) {
        ev_count = 0;
        for (size_t k = 0; k < sizeof(script) / sizeof(script[0]); k++)
            if (script[k][0] == period) { memcpy(ev_bytes[ev_count], &script[k][1], 3); ev_count++; }
        if (c->cb(FRAMES, c->arg)) return -1;
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
