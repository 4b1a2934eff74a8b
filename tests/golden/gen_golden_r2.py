"""Round-2 golden vectors (tests/golden/golden_r2.npz), like gen_golden.py from the reference's OWN
code compiled unmodified (oracle/_ref/libref.so, oracle/build_ref.sh).  Run in the build container:

    python tests/golden/gen_golden_r2.py

pixi_*: the PIXI demo LFO bank, stm32f103/pixi.c:279,282-285 (`inc = adc[0] >> 5; dac = (dac + inc) & 0xFFF`,
12 DAC channels), for several knob positions and start values.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(20261018)
ref = po.Ref()
G = {}

adc = np.array([0, 31, 32, 1000, 4095, 65535, 40000], np.uint16)        # 12-bit ADC readings and beyond (uint16 >> 5)
dac0 = rng.integers(0, 0x1000, (len(adc), 12)).astype(np.uint16)
dac0[0] = 0
dac0[1] = 0xFFF
T = 300
trace = np.zeros((len(adc), T, 12), np.uint16)
dac1 = dac0.copy()
for k, a in enumerate(adc):
    trace[k] = ref.pixi_lfo_run(dac1[k], int(a), T)
G["pixi_adc0"], G["pixi_dac0"], G["pixi_trace"], G["pixi_dac1"] = adc, dac0, trace, dac1

np.savez_compressed(os.path.join(HERE, "golden_r2.npz"), **G)
print("wrote golden_r2.npz,", len(G), "arrays")
