/* ref_clock_pre.h -- what lines 108-120 of the reference linux/clock.c need around them (they sit inside
 * a JACK process callback): the three file-scope variables (:42, :61-62) as harness-visible statics, the
 * buffers the loop writes, and send_midi() (:45-56) as a capture.  oracle/build_ref.sh pipes
 * this file, `void ref_clock_loop(...) {`, the reference's loop, and ref_clock_tail.c into gcc. */
#include <stdint.h>
#include <stddef.h>
typedef uint32_t jack_nframes_t;
static jack_nframes_t clock_hperiod;
static int clock_phase, clock_pol;
static uint32_t *ref_clock_ev_time; static uint32_t ref_clock_ev_count, ref_clock_ev_cap;
static inline void send_midi(void *out_buf, jack_nframes_t time, const void *data_buf, size_t nb_bytes) {
    (void)out_buf; (void)nb_bytes;
    if (((const uint8_t *)data_buf)[0] == 0xF8 && ref_clock_ev_count < ref_clock_ev_cap) ref_clock_ev_time[ref_clock_ev_count] = time;
    ref_clock_ev_count++;
}
static void ref_clock_loop(int nframes, float *audio_out_buf, void *midi_out_buf) {
