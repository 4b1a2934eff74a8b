import sys; sys.path.insert(0,'/root/repo')
import numpy as np, synth_tools_b200 as st
from tools import bench_configs as bc
ctx = st.Context(0)
for ch in (1, 2):
    ctx.set_option("pdm_v1_chains", ch)
    for persist in (0, 2):
        ctx.set_option("pdm_persist", persist)
        r = bc.c2_v1(st, ctx)
        print("chains", ch, "persist", persist, round(r["ms"], 3), "ms", "%.3e" % r["value"], round(r["frac"], 3))
