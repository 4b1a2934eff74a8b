#!/bin/bash
# Build oracle/liboracle.so (the CPU restatement; test infrastructure).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
gcc -std=gnu99 -O2 -fwrapv -ffp-contract=off -mfma -fopenmp -fPIC -shared -Wall \
    "$HERE/cproc_oracle.c" "$HERE/arm_v1_model.c" -o "$HERE/liboracle.so" -lm
echo "built $HERE/liboracle.so"
