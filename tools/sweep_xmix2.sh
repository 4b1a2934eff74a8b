#!/bin/bash
# k_xvoice_mix2 against k_xvoice_mix (run under gpurun): the block after note-on (gates crossed inside the block) and the steady
# state (every voice released).
for m in 1 0; do
  echo -n "xvoice_mix2=$m first block:  "; python tools/prof_one.py xvoice 4 xvoice_mix2=$m | tail -1
  echo -n "xvoice_mix2=$m steady:       "; XV_STEADY=1 python tools/prof_one.py xvoice 4 xvoice_mix2=$m | tail -1
done
