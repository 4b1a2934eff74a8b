"""GPU: a plain-C host calls the reference's own entry points (synth_run,
cproc_update/cproc_output, square_grain) through libcproc_dropin.so and the
batched C-ABI directly; results are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def xs(s):
    s ^= (s << 13) & 0xFFFFFFFF
    s ^= s >> 17
    s ^= (s << 5) & 0xFFFFFFFF
    return s


def test_c_host_through_dropin_and_abi(tmp_path, oracle):
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "test_dropin.c"), "-o", exe, "-L", pkg,
                           "-lcproc_dropin", "-lcproc_cuda", "-Wl,-rpath," + pkg])
    out = subprocess.check_output([exe], text=True).splitlines()
    tag = lambda t: [l.split()[1:] for l in out if l.split()[0] == t]
    # (1) synth_run: two 64-frame periods of the 4-note chord
    v = np.zeros((64, 2), np.uint32)
    v[:4, 0] = [39370533, 23409859, 1122405051, 731558]
    _, want = oracle.voice_bank_run(v, 64, 64, po.MIX_SAW, 128)
    got = np.array([float.fromhex(x[0]) for x in tag("synth")], np.float32)
    assert np.array_equal(got.view(np.uint32), want[0].view(np.uint32))
    assert [int(x[0]) for x in tag("phase")] == v[:4, 1].tolist()
    # (2) cproc_update upcalls cproc_output(2, n2.out)
    assert [(int(a), int(b)) for a, b in tag("output")] == [(2, x) for x in [0, 1, 1, 2, 2, 3, 4, 5, 5, 5, 6, 6, 7]]
    # (3) square_grain in place
    s = 12345
    inp = np.zeros((1, 128), np.float32)
    for i in range(128):
        s = xs(s)
        inp[0, i] = np.float32(np.int32(np.uint32(s))) * np.float32(1.0 / 2147483648.0)
    st = np.zeros(1, np.float32)
    wantg = oracle.square_grain_run(st, np.array([0.25], np.float32), 1, 128, inp)
    gotg = np.array([float.fromhex(x[0]) for x in tag("grain")], np.float32)
    assert np.array_equal(gotg.view(np.uint32), wantg[0].view(np.uint32))
    assert tag("status") == [["0"]]
    # (4) batched ABI from C
    F = 4096 + 32
    chan = np.zeros((6, 7), np.uint32)
    chan[:, 0] = 0x40000000 + 0x10000000 * np.arange(6)
    prng = np.array([2463534242, 7], np.uint32)
    duty, _ = oracle.pdm_v2_run(chan, 2, 6, 3, prng, None, 0x3FF, 0, 12, 24, None, F)
    wsum = (duty.astype(np.uint64) * (np.arange(F, dtype=np.uint64) + 1)).sum(axis=1)
    assert tag("pdm") == [["rc", "0"]] + [[str(c), str(int(wsum[c]))] for c in range(6)]
    assert [[int(x) for x in r] for r in tag("pdmstate")] == chan.tolist()
    assert tag("err") == [["rc", "-1"]]


def test_c_host_square_grain_proc_by_name_and_tag_u32_events(tmp_path, oracle):
    """tests/c/test_dropin_pd.c + host/dropin.c under -DCPROC_HAVE_PD: square_grain_proc(struct square_grain *,
    t_int, t_float *, t_float *) by the reference's name (synth_tools.c:85-86) from a perform routine, the
    TAG_U32 message handler (mod_cproc_plugin.c:24-38) and cproc_cuda_graph_event / _tick."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "test_dropin_pd")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-DCPROC_HAVE_PD", "-I", os.path.join(ROOT, "tests", "c", "fakepd"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "test_dropin_pd.c"),
                           os.path.join(pkg, "host", "dropin.c"), "-o", exe, "-L", pkg, "-lcproc_cuda", "-Wl,-rpath," + pkg])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = res.stdout.splitlines()
    tag = lambda t: [l.split()[1:] for l in out if l.split() and l.split()[0] == t]
    assert not tag("fail"), out
    # (1) square_grain_proc by name: two objects, three in-place blocks, threshold message after block 1
    s = 4242
    state = np.zeros(2, np.float32)
    th = np.array([0.25, 0.05], np.float32)
    want = [[], []]
    for blk in range(3):
        for o in range(2):
            inp = np.zeros((1, 64), np.float32)
            for i in range(64):
                s = xs(s)
                inp[0, i] = np.float32(np.int32(np.uint32(s))) * np.float32(1.0 / 2147483648.0)
            st1 = state[o:o + 1].copy()
            want[o].append(oracle.square_grain_run(st1, th[o:o + 1].copy(), 1, 64, inp)[0])
            state[o] = st1[0]
        if blk == 1:
            th[0] = np.float32(0.4)
    for o in range(2):
        got = np.array([float.fromhex(x[0]) for x in tag("grain%d" % o)], np.float32)
        assert np.array_equal(got.view(np.uint32), np.concatenate(want[o]).view(np.uint32)), o
    assert [float.fromhex(x) for x in tag("state")[0]] == state.tolist()
    assert tag("offset") == [["12"]]                           # x_f, brightness, threshold in front of state (synth_tools.c:78-84)
    # (2) TAG_U32 messages: test_cproc.c's graph, one tick per message under changed = -1
    rows = [(po.NODE_EDGE, -1, 1), (po.NODE_ACC, 0, 1)]
    seq = [0, 1, 1, 0, 1, 0, 0, 1, 0]
    wantc = oracle.graph_run(rows, 1, 1, np.zeros((1, 3), np.uint32), 1, len(seq), np.array(seq, np.uint32).reshape(1, 1, -1))
    assert [(int(a), int(b)) for a, b in tag("output")] == [(2, int(x)) for x in wantc[0]]
    assert [int(x[0]) for x in tag("msg")] == [0] * 9 + [-1, -1, -1]
    assert tag("status") == [["0"]]
    # (3) cproc_cuda_graph_event on a 3-instance batch: only the addressed instance ticks
    hist = {0: [], 1: [], 2: []}
    ev = []
    for k in range(12):
        inst = 2 if k % 3 == 2 else k & 1
        hist[inst].append((k >> 1) & 1)
        w = oracle.graph_run(rows, 1, 1, np.zeros((1, 3), np.uint32), 1, len(hist[inst]), np.array(hist[inst], np.uint32).reshape(1, 1, -1))
        ev.append(["0", str(inst), str(int(w[0, -1]))])
    assert tag("event") == ev
    last = []
    for inst in range(3):
        h = hist[inst] + [hist[inst][-1]]                      # the final tick re-reads the persistent cproc_input
        last.append(str(int(oracle.graph_run(rows, 1, 1, np.zeros((1, 3), np.uint32), 1, len(h), np.array(h, np.uint32).reshape(1, 1, -1))[0, -1])))
    assert tag("tick") == [["0"] + last]
    assert tag("bad") == [["-1", "-1"]]


def test_c_host_generated_text_patcher_bus(tmp_path, oracle):
    """tests/c/test_graph_text.c: the generated graph text, the patcher and the mix bus from plain C."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "test_graph_text")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "test_graph_text.c"), "-o", exe, "-L", pkg, "-lcproc_cuda", "-Wl,-rpath," + pkg])
    out = subprocess.check_output([exe], text=True).splitlines()
    tag = lambda t: [l.split()[1:] for l in out if l.split() and l.split()[0] == t]
    assert not tag("fail"), out
    assert tag("parse") == [["0", "5", "2", "2"]] and tag("out") == [["2", "3"], ["4", "4"]]
    rows = [(po.NODE_EDGE, -1, 1), (po.NODE_ACC, 0, 1), (po.NODE_ACC, 1, 1), (po.node_glide(3), 2, 0xFFFFFFFF), (po.node_pdm(2, 24), 3, 0xFFFFFFFF, -2)]
    assert [[int(x, 0) for x in r] for r in tag("node")] == [[r[0], r[1], r[2], r[3] if len(r) > 3 else 0] for r in rows]
    N, F = 3, 40
    inp = np.zeros((N, 2, F), np.uint32); chg = np.zeros((N, F), np.uint32)
    s = 99
    for i in range(N):
        for t in range(F):
            s = xs(s)
            inp[i, 0, t] = (s >> 7) & 1; inp[i, 1, t] = s & 0x3FF; chg[i, t] = (s >> 20) & 3
    st = np.zeros((N, 12), np.uint32)
    want = oracle.graph_run_multi(rows, 2, [2, 4], st, N, F, inp, chg)
    got = {(int(r[0]), int(r[1])): [int(x) for x in r[2:]] for r in tag("stream")}
    for i in range(N):
        for q in range(2):
            assert got[(i, q)] == want[i, q].tolist()
    assert tag("jitlog") == [["[]"]]                       # compiled by NVRTC without a diagnostic
    # patcher: bp5 graph, every instance runs (mask -1)
    assert tag("patch") == [["0", "1", "2", "3", "edge"]]
    prow = [(po.NODE_EDGE, -1, 0xFFFFFFFF), (po.NODE_ACC, 0, 0xFFFFFFFF), (po.NODE_ACC, 1, 0xFFFFFFFF)]
    pst = np.zeros((N, 4), np.uint32)
    pw = oracle.graph_run(prow, 1, 2, pst, N, F, np.ascontiguousarray(inp[:, :1, :]))
    assert [[int(x) for x in r[1:]] for r in tag("pstream")] == pw.tolist()
    assert tag("pget") == [[str(int(pst[2, 3]))]] and tag("pbad") == [["-1"]]
    assert tag("bus") == [["1", "0", "64"]] and out[-1] == "done"


def test_jack_host_adapter_with_scripted_backend(tmp_path, oracle):
    """synth_tools_b200/host/jack/jack_synth.c (16 parts x 64 voices, one batched render per
    JACK period) built against tests/c/fakejack, which plays a fixed MIDI script through the process
    callback.  Every part's audio must equal the oracle's sum_tick_saw replay bit for bit."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "jack_synth_fake")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "c", "fakejack"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host", "jack"),
                           os.path.join(pkg, "host", "jack", "jack_synth.c"), os.path.join(ROOT, "tests", "c", "fakejack", "fakejack.c"),
                           "-o", exe, "-L", pkg, "-lcproc_cuda", "-Wl,-rpath," + pkg])
    res = subprocess.run([exe], stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = {}
    for l in res.stdout.splitlines():
        f = l.split()
        if f and f[0] == "audio":
            got[(int(f[1]), f[2])] = np.array([float.fromhex(x) for x in f[3:]], np.float32)
    script = [(0, 0x90, 69, 100), (0, 0x90, 60, 100), (0, 0x93, 127, 1), (0, 0x9F, 0, 64), (1, 0x90, 72, 90), (2, 0x80, 60, 0),
              (2, 0x93, 127, 0), (3, 0x9F, 12, 80), (4, 0x90, 69, 0), (4, 0x91, 40, 127), (5, 0x8F, 0, 0)]
    F, P, V = 64, 16, 64
    voices = np.zeros((P * V, 2), np.uint32)                  # {note_inc, note_state} per part, as struct synth
    note2voice = np.zeros((P, 128), np.int64)
    for period in range(6):
        for (per, status, d1, d2) in script:
            if per != period:
                continue
            part, kind = status & 0x0F, status & 0xF0
            pv = voices[part * V:(part + 1) * V]
            if kind == 0x90 and d2 != 0:                      # note on: first free voice, else voice 0 (synth.c:143-158)
                free = np.flatnonzero(pv[:, 0] == 0)
                v = int(free[0]) if len(free) else 0
                note2voice[part, d1] = v
                pv[v, 0] = oracle.note_to_inc(d1)
            else:                                             # note off (:159-163)
                v = note2voice[part, d1]
                note2voice[part, d1] = 0
                pv[v, 0] = 0
        _, want = oracle.voice_bank_run(voices, P * V, V, po.MIX_SAW, F)      # advances the phases in place
        for part in range(P):
            g = got[(period, "part_%02d" % part)]
            assert np.array_equal(g.view(np.uint32), want[part].view(np.uint32)), (period, part)
    assert len(got) == 6 * 16
    assert any(np.abs(got[(1, "part_00")]).max() > 0 for _ in [0]) and np.abs(got[(0, "part_07")]).max() == 0


def test_pd_external_with_scripted_backend(tmp_path, oracle):
    """synth_tools_b200/host/pd/square_grain_b200~.c built against tests/c/fakepd.  Five
    objects share one batch; a perform routine hands out the previous tick's result, so outlet block k
    must equal the oracle's square_grain_proc output for inlet block k-1 (block 0 is silence), bit for
    bit, including the in-place objects and the threshold message that arrives before tick 3."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "pd_fake")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "c", "fakepd"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(pkg, "host", "pd", "square_grain_b200~.c"),
                           os.path.join(ROOT, "tests", "c", "fakepd", "fakepd.c"), "-o", exe, "-L", pkg, "-lcproc_cuda", "-lm", "-Wl,-rpath," + pkg])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = {}
    for l in res.stdout.splitlines():
        f = l.split()
        if f and f[0] == "out":
            got[(int(f[1]), int(f[2]))] = np.array([float.fromhex(x) for x in f[3:]], np.float32)
    G, B, T = 5, 64, 6
    s = 2463534242
    inp = np.zeros((T, G, B), np.float32)
    for tick in range(T):
        for g in range(G):
            for t in range(B):
                s = xs(s)
                inp[tick, g, t] = np.float32(np.int32(np.uint32(s))) * np.float32(1.0 / 2147483648.0)
    state = np.zeros(G, np.float32)
    th = (np.float32(0.05) + np.float32(0.1) * np.arange(G, dtype=np.float32)).astype(np.float32)
    for g in range(G):
        assert np.all(got[(0, g)] == 0.0)
    for tick in range(T - 1):                                 # inlet block `tick` comes out at tick + 1
        if tick == 3:
            th[2] = np.float32(0.4)                           # the message before tick 3: fabs(-0.4)
        want = oracle.square_grain_run(state, th.copy(), G, B, np.ascontiguousarray(inp[tick]))
        for g in range(G):
            assert np.array_equal(got[(tick + 1, g)].view(np.uint32), want[g].view(np.uint32)), (tick, g)
    # second scene: object 1 switched off for ticks T .. T+2 (a global count would stall every object)
    got2 = {}
    for l in res.stdout.splitlines():
        f = l.split()
        if f and f[0] == "out2":
            got2[(int(f[1]), int(f[2]))] = np.array([float.fromhex(x) for x in f[3:]], np.float32)
    # the last block of scene one (tick T-1) is rendered when scene two starts
    want = oracle.square_grain_run(state, th.copy(), G, B, np.ascontiguousarray(inp[T - 1]))
    inp2 = np.zeros((5, G, B), np.float32)
    for tick in range(5):
        for g in range(G):
            for t in range(B):
                s = xs(s)
                inp2[tick, g, t] = np.float32(np.int32(np.uint32(s))) * np.float32(1.0 / 2147483648.0)
    for tick in range(5):
        off = tick < 3
        for g in range(G):
            if off and g == 1:
                assert (T + tick, g) not in got2
                continue
            if g == 1 and tick == 3:
                assert np.all(got2[(T + tick, g)] == 0.0)        # not collected in the previous tick: silence, state kept
            else:
                assert np.array_equal(got2[(T + tick, g)].view(np.uint32), want[g].view(np.uint32)), (tick, g)
        active = [g for g in range(G) if not (off and g == 1)]
        sub_state = state[active].copy()
        w = oracle.square_grain_run(sub_state, th[active].copy(), len(active), B, np.ascontiguousarray(inp2[tick][active]))
        state[active] = sub_state
        want = np.zeros((G, B), np.float32)
        want[active] = w


def test_jack_clock_adapter_with_scripted_backend(tmp_path, oracle):
    """synth_tools_b200/host/jack/jack_clock.c (the reference's linux/clock.c word-clock / MIDI-clock
    master, one clock per BPM) against tests/c/fakejack.  Audio must equal the restated clock.c:109-120 loop bit for
    bit over 24 periods; MIDI out must carry start / stop / continue first at time 0 (clock.c:83-95, other
    realtime bytes filtered) and 0xF8 at exactly the samples where clock 0 turns positive (clock.c:113-116)."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    exe = str(tmp_path / "jack_clock_fake")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "c", "fakejack"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(pkg, "host", "jack", "jack_clock.c"),
                           os.path.join(ROOT, "tests", "c", "fakejack", "fakejack.c"), "-o", exe, "-L", pkg, "-lcproc_cuda", "-Wl,-rpath," + pkg])
    bpms = [960, 120, 7500]
    env = dict(os.environ, FAKEJACK_SCRIPT="clock", FAKEJACK_PERIODS="24")
    res = subprocess.run([exe] + [str(b) for b in bpms], stdin=subprocess.DEVNULL, capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr
    audio, midi = {}, {}
    for l in res.stdout.splitlines():
        f = l.split()
        if f and f[0] == "audio":
            audio[(int(f[1]), f[2])] = np.array([float.fromhex(x) for x in f[3:]], np.float32)
        elif f and f[0] == "midi":
            midi.setdefault(int(f[1]), []).append((int(f[3]), [int(x, 16) for x in f[4:]]))
    F, P = 64, 24
    hp = np.array([(48000 * 5) // (b * 4) for b in bpms], np.int32)            # BPM_TO_HPERIOD, clock.c:58
    assert list(hp) == [62, 500, 8]
    state = np.zeros((3, 2), np.int32); state[:, 1] = 1                       # clock_phase = 0, clock_pol = 1 (clock.c:61-62)
    passed = {1: 0xFA, 3: 0xFC, 5: 0xFB}
    last = 1
    n_clock_bytes = 0
    for period in range(P):
        want = oracle.word_clock_run(state, hp, 3, F)
        for k in range(3):
            assert np.array_equal(audio[(period, "clock_%02d" % k)].view(np.uint32), want[k].view(np.uint32)), (period, k)
        expect = [(0, [passed[period]])] if period in passed else []
        for t in range(F):
            pol = int(want[0, t])
            if pol == 1 and last != 1:
                expect.append((t, [0xF8]))
            last = pol
        n_clock_bytes += sum(1 for e in expect if e[1] == [0xF8])
        assert midi.get(period, []) == expect, (period, midi.get(period), expect)
    assert n_clock_bytes == (P * F) // 124                                   # one positive edge per 2 * 62 samples
