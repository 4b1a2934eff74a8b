"""Golden vectors (tests/golden/golden_r1.npz, generated from the reference's own
code compiled unmodified -- tests/golden/gen_golden.py).  CPU: the oracle
reproduces every vector.  GPU: the CUDA path, through the C-ABI, does too."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r1.npz"))


class OracleBackend:
    name = "oracle"

    def __init__(self):
        self.o = po.Oracle()

    def graph(self, rows, state, inp, changed):
        N, n_in, F = inp.shape
        return self.o.graph_run(rows, n_in, len(rows) - 1, state, N, F, np.ascontiguousarray(inp), changed)

    def graph_ext(self, rows, n_in, outs, state, prm, inp, changed):
        N, F = state.shape[0], inp.shape[-1]
        return self.o.graph_run_ext(rows, n_in, outs, state, prm, N, F, inp, changed)

    def pdm(self, order, state, inp, dither):
        N, F = inp.shape
        return self.o.pdm_run(order, state, N, F, inp, None, 24, dither)

    def v2(self, chan, order, bank, prng, count, ctl, sp, F):
        N = chan.shape[0]
        duty, cnt = self.o.pdm_v2_run(chan, order, N, bank, prng, None, 0x3FF, count, ctl, 24, sp, F)
        return duty, cnt

    def v1(self, ch, bank, prng, F):
        return np.packbits(self.o.pdm_v1_run(ch, ch.shape[0], bank, prng, None, 0x0FFFFFFF, F), axis=1, bitorder="little")

    def pwm(self, ph, spd, F):
        return self.o.pwm_run(ph, spd, len(ph), F)

    def voices(self, v, G_, mode, F):
        return self.o.voice_bank_run(v, v.shape[0], G_, mode, F)[1]

    def grain(self, state, th, inp):
        N, F = inp.shape
        return self.o.square_grain_run(state, th, N, F, inp)


class CudaBackend:
    name = "cuda"

    def __init__(self):
        import synth_tools_b200 as st
        self.st = st
        self.ctx = st.Context(0)

    def graph(self, rows, state, inp, changed):
        st = self.st
        N, n_in, F = inp.shape
        b = self.ctx.batch(st.GRAPH, N, nodes=rows, n_inputs=n_in)
        b.upload_state(state)
        out = np.zeros((N, F), np.uint32)
        b.run(F, inp=np.ascontiguousarray(inp), in2=changed, out=out)
        state[:] = b.download_state()
        b.free()
        return out

    def graph_ext(self, rows, n_in, outs, state, prm, inp, changed):
        st = self.st
        N, F = state.shape[0], inp.shape[-1]
        b = self.ctx.batch(st.GRAPH, N, nodes=rows, n_inputs=n_in, out_node=list(outs))
        b.upload_state(state)
        if prm is not None and prm.shape[1]:
            b.upload_param(prm)
        out = np.zeros((N, len(outs), F), np.uint32)
        b.run(F, inp=np.ascontiguousarray(inp), in2=changed, out=out)
        state[:] = b.download_state()
        b.free()
        return out

    def pdm(self, order, state, inp, dither):
        st = self.st
        N, F = inp.shape
        b = self.ctx.batch(st.PDM, N, order=order, out_shift=24)
        b.upload_state(state)
        out = np.zeros((N, F), np.uint32)
        b.run(F, inp=inp, in2=dither, out=out)
        state[:] = b.download_state()
        b.free()
        return out

    def v2(self, chan, order, bank, prng, count, ctl, sp, F):
        st = self.st
        N = chan.shape[0]
        b = self.ctx.batch(st.PDM_V2, N, order=order, bank_size=bank, ctl_div_log=ctl)
        b.upload_state(chan); b.upload_bank(prng, count)
        out = np.zeros((N, F), np.uint8)
        b.run(F, ctl=sp, out=out)
        chan[:] = b.download_state()
        p, cnt = b.download_bank()
        prng[:] = p
        b.free()
        return out, cnt

    def v1(self, ch, bank, prng, F):
        st = self.st
        N = ch.shape[0]
        b = self.ctx.batch(st.PDM_V1, N, bank_size=bank, dither_mask=0x0FFFFFFF)
        b.upload_state(ch); b.upload_bank(prng)
        out = np.zeros((N, F // 32), np.uint32)
        b.run(F, out=out)
        ch[:] = b.download_state()
        prng[:] = b.download_bank()[0]
        b.free()
        return out.view(np.uint8)

    def pwm(self, ph, spd, F):
        st = self.st
        b = self.ctx.batch(st.PWM, len(ph))
        b.upload_state(ph.reshape(-1, 1).copy()); b.upload_param(spd.reshape(-1, 1).copy())
        out = np.zeros((len(ph), F), np.uint8)
        b.run(F, out=out)
        ph[:] = b.download_state()[:, 0]
        b.free()
        return out

    def voices(self, v, G_, mode, F):
        st = self.st
        b = self.ctx.batch(st.VOICE_BANK, v.shape[0], mode=mode, voices_per_bus=G_)
        b.upload_state(v)
        out = np.zeros((v.shape[0] // G_, F), np.float32)
        b.run(F, out=out)
        v[:] = b.download_state()
        b.free()
        return out

    def grain(self, state, th, inp):
        st = self.st
        N, F = inp.shape
        b = self.ctx.batch(st.SQUARE_GRAIN, N)
        b.upload_state(state.reshape(-1, 1).copy()); b.upload_param(th.reshape(-1, 1).copy())
        out = np.zeros((N, F), np.float32)
        b.run(F, inp=inp, out=out)
        state[:] = b.download_state().view(np.float32)[:, 0]
        b.free()
        return out


@pytest.fixture(scope="module", params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request):
    return OracleBackend() if request.param == "oracle" else CudaBackend()


def test_golden_test_cproc_graph(be):
    st = np.zeros((1, 3), np.uint32)
    out = be.graph(po.GRAPH_TEST_CPROC, st, G["cproc_in"][None, None, :], G["cproc_mask"][None, :].copy())
    assert np.array_equal(out[0], G["cproc_out"])


def test_golden_bp5_graph(be):
    st = np.zeros((8, 4), np.uint32)
    out = be.graph(po.GRAPH_BP5, st, G["bp5_in"], G["bp5_changed"].copy())
    assert np.array_equal(out, G["bp5_out"]) and np.array_equal(st, G["bp5_state"])


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_golden_pdm_update(be, k):
    st = np.zeros((4, k), np.uint32)
    out = be.pdm(k, st, G["pdm_in"].copy(), G["pdm_dither"].copy())
    assert np.array_equal(out, G["pdm%d_out" % k]) and np.array_equal(st, G["pdm%d_state" % k])


def test_golden_v2_firmware_config(be):
    F = 2 * 4096 + 64
    chan = np.zeros((3, 7), np.uint32)
    chan[:, 0] = [2000000000, 0x40000000, 0x40000000]
    prng = np.array([2463534242], np.uint32)
    duty, cnt = be.v2(chan, 2, 3, prng, 0, 12, G["v2fw_setpoints"].copy(), F)
    assert np.array_equal(duty[:, -256:], G["v2fw_duty_tail"])
    wsum = (duty.astype(np.uint64) * (np.arange(F, dtype=np.uint64) + 1)).sum(axis=1)
    assert np.array_equal(wsum, G["v2fw_duty_wsum"])
    assert np.array_equal(chan, G["v2fw_state"]) and np.array_equal(prng, G["v2fw_prng"]) and cnt == G["v2fw_count"][0]


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_golden_v2_orders(be, k):
    chan = G["v2k%d_chan0" % k].copy()
    prng = np.array([1, 2, 3, 4], np.uint32)
    sp = po.pdm_setpoints(12, 512 // 64 + 1)
    duty, _ = be.v2(chan, k, 3, prng, 16, 6, sp, 512)
    assert np.array_equal(duty, G["v2k%d_duty" % k])
    assert np.array_equal(chan, G["v2k%d_state" % k]) and np.array_equal(prng, G["v2k%d_prng" % k])


def test_golden_note_table():
    o = po.Oracle()
    assert [o.note_to_inc(n) for n in range(128)] == G["note_table"].tolist()


def test_golden_voice_bank(be):
    for mode, key in ((0, "saw"), (1, "square")):
        v = G["voices0"].copy()
        vec = be.voices(v, 64, mode, 64)
        assert np.array_equal(vec.view(np.uint32), G[key + "_vec"].view(np.uint32))
        assert np.array_equal(v, G[key + "_voices"])
    v = np.zeros((64, 2), np.uint32)
    v[:4, 0] = G["note_table"][[69, 60, 127, 0]]
    vec = be.voices(v, 64, 0, 8)
    assert np.array_equal(vec[0].view(np.uint32), G["chord_vec"].view(np.uint32)) and np.array_equal(v, G["chord_voices"])


def test_golden_square_grain(be):
    s = G["grain_state0"].copy()
    out = be.grain(s, G["grain_thresh"].copy(), G["grain_in"].copy())
    assert np.array_equal(out.view(np.uint32), G["grain_out"].view(np.uint32)) and np.array_equal(s, G["grain_state"])


def test_golden_restated_v1_and_pwm(be):
    ch = np.zeros((6, 2), np.uint32)
    ch[:, 0] = [2000000000, 0x40000000, 0x60000000, 0x80000000, 0xA0000000, 0xC0000000]
    prng = np.array([2463534242, 5, 9], np.uint32)
    bits = be.v1(ch, 2, prng, 256)
    assert np.array_equal(bits, G["restated_v1_bits"]) and np.array_equal(ch, G["restated_v1_state"])
    assert np.array_equal(prng, G["restated_v1_prng"])
    ph = np.array([0, 0x123456], np.uint32)
    duty = be.pwm(ph, np.array([256 * 13, 5000], np.uint32), 256)
    assert np.array_equal(duty, G["restated_pwm_duty"]) and np.array_equal(ph, G["restated_pwm_phase"])


G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r2.npz"))


def test_golden_pixi_lfo_bank_is_acc_under_a_12_bit_mask(be):
    """the PIXI demo LFO bank (stm32f103/pixi.c:279,282-285), `dac = (dac + inc) & 0xFFF` with
    inc = adc[0] >> 5 on 12 uint16 channels, is the `acc` processor (cproc.h:134-142) read through a 12-bit
    mask: (acc.out & 0xFFF) equals the reference's DAC values tick for tick, for every knob position."""
    adc, dac0, trace, dac1 = G2["pixi_adc0"], G2["pixi_dac0"], G2["pixi_trace"], G2["pixi_dac1"]
    K, T, D = trace.shape
    N = K * D                                                # one acc instance per (knob position, DAC channel)
    inc = np.repeat(adc.astype(np.uint32) >> 5, D)
    inp = np.ascontiguousarray(np.broadcast_to(inc[:, None, None], (N, 1, T))).astype(np.uint32)
    state = dac0.astype(np.uint32).reshape(N, 1).copy()
    out = be.graph([(po.NODE_ACC, -1, 0xFFFFFFFF)], state, inp, None)
    assert np.array_equal((out & 0xFFF).reshape(K, D, T).transpose(0, 2, 1), trace.astype(np.uint32))
    assert np.array_equal((state[:, 0] & 0xFFF).reshape(K, D), dac1.astype(np.uint32))


@pytest.mark.parametrize("k", [0, 1, 2])
def test_golden_extension_processor_graphs(be, k):
    """include/cproc_ext.h compiled against the reference's cproc.h (DEF_PROC bodies), on node tables that mix the
    extension processors with acc / edge."""
    rows = [tuple(int(x) for x in r) for r in G2["ext%d_rows" % k]]
    prm, inp, chg = G2["ext%d_param" % k].copy(), G2["ext%d_in" % k].copy(), G2["ext%d_changed" % k]
    state = np.zeros_like(G2["ext%d_state" % k])
    outs = list(range(len(rows)))[-3:]
    out = be.graph_ext(rows, 2, outs, state, prm if prm.shape[1] else None, inp, chg.copy() if chg.size else None)
    assert np.array_equal(out, G2["ext%d_out" % k]) and np.array_equal(state, G2["ext%d_state" % k])


def test_golden_graph_texts_compiled_as_c(be):
    """tests/golden/ext_voice.cproc / ext_chain.cproc compiled as C with the reference's PROC / PROC_COND macros."""
    from synth_tools_b200 import abi
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, n_in in (("voice", 1), ("chain", 2), ("gain", 1)):
        g = abi.graph_parse_ex(open(os.path.join(here, "ext_%s.cproc" % name)).read())
        inp = G2["ext%s_in" % name].reshape(1, n_in, -1).copy()
        chg = G2["extchain_changed"].reshape(1, -1).copy() if name == "chain" else None
        state = np.zeros((1, sum(po.node_words(r[0]) for r in g["rows"])), np.uint32)
        out = be.graph_ext(g["rows"], n_in, g["out_nodes"], state, g["param_init"].reshape(1, -1).copy(), inp, chg)
        assert np.array_equal(out[0], G2["ext%s_out" % name]), name
