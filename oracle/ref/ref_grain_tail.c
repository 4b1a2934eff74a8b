/* ref_grain_tail.c -- appended (same translation unit) to lines 78-100 of the
 * reference linux/synth_tools.c (struct square_grain + square_grain_proc),
 * which are piped in by oracle/build_ref.sh after ref_pre_pd.h. */
uint32_t ref_grain_sizeof(void) { return sizeof(struct square_grain); }
uint32_t ref_grain_offsetof(int what) { return what == 0 ? offsetof(struct square_grain, threshold) : offsetof(struct square_grain, state); }
/* state/threshold [N]; in/out [N][F] planar; in may alias out. */
void ref_square_grain_run(float *state, const float *threshold, uint64_t N, uint64_t F, const float *in, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        struct square_grain g;
        g.state = state[n]; g.threshold = threshold[n]; g.brightness = 1.0f;
        square_grain_proc(&g, (t_int)F, (t_float *)(in + n * F), out + n * F);
        state[n] = g.state;
    }
}
