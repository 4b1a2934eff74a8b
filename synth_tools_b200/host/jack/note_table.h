/* note_table.h -- MIDI note -> 32-bit phasor increment at 48 kHz, the table of linux/synth.c:78-125:
 * twelve increments for the top octave (MIDI 116..127), computed by the reference at compile time in
 * double precision and truncated to uint32 (values recorded in SURVEY.md 8a-9 from the reference
 * compiled with gcc), lower octaves by a right shift.  Never recomputed in float. */
#ifndef CPROC_NOTE_TABLE_H
#define CPROC_NOTE_TABLE_H
#include <stdint.h>
static const uint32_t cproc_note_top_octave[12] = {
    594573364u, 629928536u, 667386036u, 707070875u, 749115497u, 793660223u,
    840853716u, 890853479u, 943826384u, 999949221u, 1059409296u, 1122405051u,
};
static inline uint32_t cproc_note_to_inc(int note) {
    note &= 127;
    /* the table covers 116..127; note n sits (127 - n) / 12 octaves below its entry */
    const int below = (127 - note) / 12;
    return cproc_note_top_octave[11 - (127 - note) % 12] >> below;
}
#endif
