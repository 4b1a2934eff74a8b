#!/bin/bash
# Build oracle/_ref/libref.so from the UNMODIFIED reference sources where they
# lie under $REF (default /root/reference).  Nothing is copied: headers come in
# through -I, and the two line ranges that live inside larger, unbuildable
# files (JACK / Pd glue around them) are piped from sed straight into gcc.
# The reference's own build system (uc_tools build.sh, nix) is NOT run.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present, keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$HERE/_ref"
# Two builds of the same sources: libref.so with -O2 (the primary CPU baseline: the reference's own optimisation flags live in
# uc_tools' build.sh, not in the tree) and libref_o3.so with -O3 for the x86-64-v3 level (AVX2, BMI2, FMA).  BASELINE.md 3 asks
# for -march=native, but the library is built HERE (the reference sources do not travel to the GPU box) and must run on the
# box's host CPU, whatever it is: v3 is what every host of a B200 has.
build_one() {
OUT="$1"; OPT="$2"
OBJ="$(mktemp -d /tmp/cproc_ref_obj.XXXXXX)"
CFLAGS="-std=gnu99 $OPT -fwrapv -ffp-contract=off -fPIC -fopenmp -w"
# (1) generic/cproc.h + linux/test_cproc.c (whole file; main renamed)
gcc $CFLAGS -I"$HERE/shim" -I"$REF/generic" -I"$HERE/../include" -c "$HERE/ref/ref_cproc.c" -o "$OBJ/ref_cproc.o"
gcc $CFLAGS -I"$HERE/shim" -I"$REF/generic" -Dmain=ref_test_cproc_main -c "$REF/linux/test_cproc.c" -o "$OBJ/test_cproc.o"
# (1b) the graph texts of tests/golden/*.cproc compiled as C against the real cproc.h + include/cproc_ext.h
gcc $CFLAGS -I"$HERE/shim" -I"$REF/generic" -I"$HERE/../include" -I"$HERE/../tests/golden" -c "$HERE/ref/ref_ext_voice.c" -o "$OBJ/ref_ext_voice.o"
# (2) stm32f103/pdm.h
gcc $CFLAGS -I"$REF/stm32f103" -c "$HERE/ref/ref_pdm.c" -o "$OBJ/ref_pdm.o"
# (3) linux/synth.c:29-206 (voice bank DSP; JACK glue excluded)
( echo '#include <stdint.h>'; echo '#include <strings.h>'; sed -n 29,206p "$REF/linux/synth.c"; cat "$HERE/ref/ref_synth_tail.c" ) \
  | gcc $CFLAGS -x c -c - -o "$OBJ/ref_synth.o"
# (4) linux/synth_tools.c:78-100 (struct square_grain + square_grain_proc)
( cat "$HERE/ref/ref_pre_pd.h"; sed -n 78,100p "$REF/linux/synth_tools.c"; cat "$HERE/ref/ref_grain_tail.c" ) \
  | gcc $CFLAGS -x c -c - -o "$OBJ/ref_grain.o"
# (5) stm32f103/mod_pdm_pwm.c + mod_controlrate.c, whole files, against a hosted stand-in for the hardware layer
# (pdm.h's trailing always_inline attribute lands on the next external function: allow inlining it under -fPIC)
gcc $CFLAGS -fno-semantic-interposition -I"$HERE/shim/stm32" -I"$REF/stm32f103" -c "$HERE/ref/ref_v2_isr.c" -o "$OBJ/ref_v2_isr.o"
# (6) linux/clock.c:108-120 (the word-clock loop; JACK glue excluded)
( cat "$HERE/ref/ref_clock_pre.h"; sed -n 108,120p "$REF/linux/clock.c"; cat "$HERE/ref/ref_clock_tail.c" ) \
  | gcc $CFLAGS -x c -c - -o "$OBJ/ref_clock.o"
# (7) stm32f103/mod_pdm.c:159-175 (pwm_update and the globals it works on)
( echo '#include <stdint.h>'; sed -n 159,175p "$REF/stm32f103/mod_pdm.c"; cat "$HERE/ref/ref_pwm_tail.c" ) \
  | gcc $CFLAGS -Dcontrol_div_count=ref_pwm_control_div_count -x c -c - -o "$OBJ/ref_pwm.o"   # (:165 also defines the v1 divider; mod_pdm_pwm.c has its own)
# (8) stm32f103/pixi.c:279,282-285 (the PIXI demo LFO bank: inc = adc[0] >> 5; dac = (dac + inc) & 0xFFF)
( cat "$HERE/ref/ref_pixi_pre.h"; sed -n '279p;282,285p' "$REF/stm32f103/pixi.c"; cat "$HERE/ref/ref_pixi_tail.c" ) \
  | gcc $CFLAGS -x c -c - -o "$OBJ/ref_pixi.o"
gcc -shared -fopenmp -o "$HERE/_ref/$OUT" "$OBJ/ref_cproc.o" "$OBJ/ref_ext_voice.o" "$OBJ/test_cproc.o" "$OBJ/ref_pdm.o" "$OBJ/ref_synth.o" "$OBJ/ref_grain.o" \
    "$OBJ/ref_v2_isr.o" "$OBJ/ref_clock.o" "$OBJ/ref_pwm.o" "$OBJ/ref_pixi.o" -lm
rm -rf "$OBJ"
echo "built $HERE/_ref/$OUT ($OPT)"
}
build_one libref.so "-O2"
build_one libref_o3.so "-O3 -march=x86-64-v3"
