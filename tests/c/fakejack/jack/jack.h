/* A scripted stand-in for the part of the JACK API that synth_tools_b200/host/jack/jack_synth.c
 * and jack_clock.c use (this image has no libjack).  Test infrastructure: see fakejack.c. */
#ifndef FAKE_JACK_H
#define FAKE_JACK_H
#include <stdint.h>
typedef uint32_t jack_nframes_t;
typedef struct fake_client jack_client_t;
typedef struct fake_port jack_port_t;
typedef int jack_status_t;
typedef int jack_options_t;
typedef float jack_default_audio_sample_t;
typedef int (*JackProcessCallback)(jack_nframes_t nframes, void *arg);
#define JackNullOption 0
#define JackPortIsInput 1
#define JackPortIsOutput 2
#define JACK_DEFAULT_AUDIO_TYPE "32 bit float mono audio"
#define JACK_DEFAULT_MIDI_TYPE "8 bit raw midi"
jack_client_t *jack_client_open(const char *name, jack_options_t options, jack_status_t *status, ...);
int jack_client_close(jack_client_t *c);
jack_nframes_t jack_get_buffer_size(jack_client_t *c);
jack_nframes_t jack_get_sample_rate(jack_client_t *c);
jack_port_t *jack_port_register(jack_client_t *c, const char *name, const char *type, unsigned long flags, unsigned long bufsize);
void *jack_port_get_buffer(jack_port_t *p, jack_nframes_t nframes);
int jack_set_process_callback(jack_client_t *c, JackProcessCallback cb, void *arg);
int jack_activate(jack_client_t *c);
#endif
