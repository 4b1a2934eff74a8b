"""Instruction budgets of the hot loops, read from the SASS of the built objects with cuobjdump (no GPU needed).

DESIGN.md quotes instructions per tick for the issue-bound kernels (PDM v2: 6 per channel-tick in the consumer warps and 7 per
bank-tick in the dither producer; PDM v1: 13 per bank-tick at banks of 2; the extension voice pair: packed FFMA2 ticks).  Their
speed follows the number of issued instructions (profiles/r2_pdm_v2_mix_experiments.txt), so a change that lets the compiler
spill an add onto another pipe or re-derive an address inside a tick loop costs percent at once -- this test pins the loop
bodies to the counts the measurements were taken with."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "synth_tools_b200", "build")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")

_LINE = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)")
_BRA = re.compile(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s+)?0x([0-9a-f]+)")


def sass(obj, fun):
    if not os.path.exists(os.path.join(OBJ, obj)):
        pytest.skip("%s not built here (python synth_tools_b200/build.py keeps its objects under synth_tools_b200/build/)" % obj)
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = _LINE.search(line)
        if m and "/* 0x" not in line.split("*/")[0]:
            ins.append((int(m.group(1), 16), m.group(2), line))
    assert ins, "kernel %s not found in %s (build the library first: python synth_tools_b200/build.py)" % (fun, obj)
    return ins


def loops(ins, min_len=200):
    """Bodies of the backward branches, longest first: (start, end, Counter of opcodes with their first modifier)."""
    res = []
    for addr, op, line in ins:
        if op.startswith("BRA"):
            m = _BRA.search(line)
            if m and int(m.group(1), 16) < addr:
                t = int(m.group(1), 16)
                body = [o for a, o, _ in ins if t <= a <= addr]
                if len(body) >= min_len:
                    res.append((t, addr, collections.Counter(".".join(o.split(".")[:2]) if o.startswith("IMAD") else o.split(".")[0] for o in body)))
    return sorted(res, key=lambda r: r[0] - r[1])


def test_pdm_v2_consumer_batch_is_six_instructions_per_tick_plus_overhead():
    """k_pdm_v2_ws4<K=2, CW=3, T=128, TILED>: the consumer's batch loop (128 ticks of one channel per lane)."""
    ins = sass("k_pdm_v2.o", "_Z12k_pdm_v2_ws4ILi2ELi3ELi7ELi0EEv11PdmV2Params9PdmV2Work14CUtensorMap_st")
    body = min((l for l in loops(ins) if l[2]["PRMT"] >= 90), key=lambda l: l[1] - l[0])          # innermost loop that packs bytes: one batch
    c = body[2]
    total = sum(c.values())
    # per tick: LOP3 (quantiser), IMAD (in - out), IADD3 (3 registers), 2 x IMAD.IADD (glide, s0 + p), 0.75 PRMT, 0.25 LDS.128
    assert c["LOP3"] in range(128, 134) and c["PRMT"] == 96 and c["LDS"] == 32 and c["STG"] == 8
    assert 128 <= c["IADD3"] <= 140 and 128 <= c["IMAD"] <= 140 and 256 <= c["IMAD.IADD"] <= 272
    assert total <= 880, "consumer batch loop grew: %d instructions for 128 ticks (6.0 per tick + %d)" % (total, total - 768)


def test_pdm_v2_producer_is_seven_instructions_per_bank_tick():
    ins = sass("k_pdm_v2.o", "_Z12k_pdm_v2_ws4ILi2ELi3ELi7ELi0EEv11PdmV2Params9PdmV2Work14CUtensorMap_st")
    prod = [l for l in loops(ins, 150) if l[2]["STS"] >= 8 and l[2]["PRMT"] == 0]
    assert prod, "producer loop not found"
    c = min(prod, key=lambda l: l[1] - l[0])[2]          # the innermost generator loop
    ticks = c["SHF"]                                     # one right shift per xorshift step
    assert ticks in (32, 64)
    # xorshift32: 2 x IMAD.SHL + SHF + 3 x LOP3 (xor), + 1 LOP3 (dither mask) = 7 per bank-tick; one STS.128 per 4
    assert c["LOP3"] == 4 * ticks and c["IMAD.SHL"] == 2 * ticks and c["STS"] == ticks // 4
    assert sum(c.values()) <= 7.6 * ticks


def test_pdm_v1_tiled_group_loop():
    """k_pdm_v1_simple<B=2, thread per bank>: two unrolled groups of four words = 256 ticks of two channels."""
    ins = sass("k_pdm.o", "_Z15k_pdm_v1_simpleILi2ELb1ELb0EEv11PdmV1Params")
    body = loops(ins, 2000)[0][2]
    ticks = 256
    assert body["BREV"] == 16 and body["STG"] == 2
    # per bank-tick: generator 4 LOP3 + SHF + 2 IMAD.SHL; per channel-tick sp + d (IMAD.IADD), add.cc (IADD3), addc (IMAD.X)
    assert body["IADD3"] <= 2 * ticks + 8 and body["IMAD.X"] <= 2 * ticks and body["IMAD.IADD"] <= 2 * ticks + 8
    assert body["LOP3"] <= 4.1 * ticks and body["IMAD.SHL"] <= 2.05 * ticks
    assert sum(body.values()) <= 13.4 * ticks, sum(body.values())


def test_xvoice_mix2_tick_loops_are_packed():
    """k_xvoice_mix2: both unrolled 32-tick loops (uniform envelope phase, per-tick selection) run the SVF, the envelope add and the
    output multiply of a voice PAIR as packed fp32 instructions."""
    ins = sass("k_xvoice.o", "_Z13k_xvoice_mix212XVoiceParams8BusFused")
    c = collections.Counter(o.split(".")[0] for _, o, _ in ins)
    assert c["FFMA2"] == 2 * 32 * 4 and c["FADD2"] == 2 * 32 and c["FMUL2"] == 2 * 32
    assert c["LDGSTS"] >= 5                              # the tile's state comes in with cp.async
