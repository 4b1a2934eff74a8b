#!/bin/bash
# Tile-shape sweep of the generated PLANAR graph kernel on the C4 voice graph (run under gpurun).
for hint in 0 1; do for tf in 32 64; do for w in 1 2; do
  echo -n "HINT=$hint TF=$tf STAGES=3 WARPS=$w  "
  CPROC_GRAPH_ST_HINT=$hint CPROC_GRAPH_TF=$tf CPROC_GRAPH_WARPS=$w timeout 120 python tools/prof_one.py gvoice 3 2>&1 | tail -1
done; done; done
for tf in 32 64; do echo -n "per-lane bulk copies (planar_bulk=1) TF=$tf  "; CPROC_GRAPH_TF=$tf timeout 120 python tools/prof_one.py gvoice 3 planar_bulk=1 2>&1 | tail -1; done
