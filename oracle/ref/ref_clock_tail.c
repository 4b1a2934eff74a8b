}
/* ref_clock_tail.c -- closes ref_clock_loop() around the reference's loop and exposes it.
 * state[n] = {phase, pol}; out [N][F]; ev_time[0..ev_cap) receives the sample times of the MIDI clock
 * bytes of clock `ev_clock` (the count is returned even beyond the capacity). */
uint32_t ref_word_clock_run(int32_t *state, const int32_t *hperiod, uint64_t N, uint64_t F, float *out,
                            uint64_t ev_clock, uint32_t *ev_time, uint32_t ev_cap) {
    uint32_t n_ev = 0;
    for (uint64_t n = 0; n < N; n++) {
        clock_phase = state[2 * n]; clock_pol = state[2 * n + 1]; clock_hperiod = (jack_nframes_t)hperiod[n];
        ref_clock_ev_time = n == ev_clock ? ev_time : NULL; ref_clock_ev_cap = n == ev_clock ? ev_cap : 0; ref_clock_ev_count = 0;
        ref_clock_loop((int)F, out + n * F, NULL);
        if (n == ev_clock) n_ev = ref_clock_ev_count;
        state[2 * n] = clock_phase; state[2 * n + 1] = clock_pol;
    }
    return n_ev;
}
