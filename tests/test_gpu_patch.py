"""GPU: the dynamic patcher boundary (stm32f103/mod_bpmodular.c RPC tree:
class/<c>/apply, inst/<node>/state/<k>/get|set, patch/reset, patch/tick) on device state,
checked against the CPU oracle's graph interpreter."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(11)


@pytest.fixture(scope="module")
def st():
    import synth_tools_b200 as st
    return st


@pytest.fixture(scope="module")
def ctx(st):
    c = st.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return po.Oracle()


def test_patch_builds_the_bp5_graph_incrementally(st, ctx, oracle):
    """apply edge, tick; apply acc, tick; apply acc, tick: every rebuild keeps the state of the
    existing instances, and the whole history equals the oracle run on the growing table."""
    N, F = 300, 64
    p = st.Patch(ctx, N, n_inputs=1)
    n_in = p.apply("input", config=0)
    n1 = p.apply("edge", n_in)
    assert (n_in, n1) == (0, 1)
    p.output(n1)
    rows = [(po.NODE_EDGE, -1, 0xFFFFFFFF)]
    state = np.zeros((N, 2), np.uint32)                    # apply zero-initialises (mod_bpmodular.c:101)
    out = np.zeros((N, F), np.uint32)
    for step, cls in enumerate([None, "acc", "acc"]):
        if cls:
            last = p.apply(cls, p.node_count - 1)
            p.output(last)
            rows.append((po.NODE_ACC, len(rows) - 1, 0xFFFFFFFF))
            state = np.concatenate([state, np.zeros((N, 1), np.uint32)], axis=1)
        inp = rng.integers(0, 2, (N, 1, F), dtype=np.uint32)
        want = oracle.graph_run(rows, 1, len(rows) - 1, state, N, F, inp)
        p.tick(F, inp, out)
        assert np.array_equal(out, want), step
    # inst/<node>/state/<k>/get
    for inst in (0, 17, N - 1):
        assert p.get(n1, 0, inst) == state[inst, 0] and p.get(n1, 1, inst) == state[inst, 1]
        assert p.get(3, 0, inst) == state[inst, 3]
    # set: overwrite an accumulator of one instance, the next tick continues from it
    p.set(2, 0, 1000, instance=5)
    state[5, 2] = 1000
    inp = rng.integers(0, 2, (N, 1, F), dtype=np.uint32)
    want = oracle.graph_run(rows, 1, 2, state, N, F, inp)
    p.tick(F, inp, out)
    assert np.array_equal(out, want)
    p.close()


def test_patch_glide_portamento_and_reset(st, ctx, oracle):
    N, F, L = 64, 512, 5
    p = st.Patch(ctx, N, n_inputs=2, layout=st.INTERLEAVED)
    a = p.apply("input", config=1)
    g = p.apply("glide", a, config=L)
    osc = p.apply("acc", g)
    p.output(osc)
    rows = [(po.node_glide(L), -2, 0xFFFFFFFF), (po.NODE_ACC, 0, 0xFFFFFFFF)]
    state = np.zeros((N, 6), np.uint32)
    inp = rng.integers(0, 2**24, (N, 2, F), dtype=np.uint32)
    want = oracle.graph_run(rows, 2, 1, state, N, F, inp)
    out = np.zeros((F, N), np.uint32)
    p.tick(F, np.ascontiguousarray(inp.transpose(2, 1, 0)), out)
    assert np.array_equal(out.T, want)
    assert p.get(g, 4, 3) == F % (1 << L)                  # glide.count
    p.reset()                                              # patch/reset: balloci_clear
    assert p.node_count == 0
    with pytest.raises(st.CprocCudaError):
        p.tick(F, inp, out)                                # nothing to run
    p.close()


def test_patch_pdm_channel(st, ctx, oracle):
    """apply(glide), apply(pdm): the v2 channel patched together at run time."""
    N, F, L = 40, 640, 5
    p = st.Patch(ctx, N, n_inputs=2)
    sp_in, di_in = p.apply("input", config=0), p.apply("input", config=1)
    line = p.apply("glide", sp_in, config=L)
    mod = p.apply("pdm", line, di_in, config=2 | (24 << 3))
    p.output(mod)
    rows = [(po.node_glide(L), -1, 0xFFFFFFFF), (po.node_pdm(2, 24), 0, 0xFFFFFFFF, -2)]
    inp = rng.integers(0, 2**32, (N, 2, F), dtype=np.uint32)
    inp[:, 1] &= 0x3FF
    state = np.zeros((N, 8), np.uint32)
    want = oracle.graph_run(rows, 2, 1, state, N, F, inp)
    out = np.zeros((N, F), np.uint32)
    p.tick(F, inp, out)
    assert np.array_equal(out, want)
    assert p.get(mod, 1, 7) == state[7, 6] and p.get(mod, 2, 7) == state[7, 7]      # s1, s2 of instance 7
    with pytest.raises(st.CprocCudaError):
        p.get(mod, 3, 0)                                   # an order-2 instance has fields out, s1, s2
    with pytest.raises(st.CprocCudaError):
        p.apply("pdm", line, di_in, config=5 | (24 << 3))
    with pytest.raises(st.CprocCudaError):
        p.apply("pdm", line)
    p.close()


def test_patch_errors(st, ctx):
    p = st.Patch(ctx, 8, n_inputs=1)
    with pytest.raises(st.CprocCudaError):                 # bad_node (mod_bpmodular.c:106-110)
        p.apply("acc", 0)
    i = p.apply("input", config=0)
    with pytest.raises(st.CprocCudaError):                 # input stream out of range
        p.apply("input", config=1)
    with pytest.raises(st.CprocCudaError):                 # wrong number of inputs (:263)
        p.apply("acc")
    with pytest.raises(st.CprocCudaError):
        p.apply("glide", i, config=0)
    with pytest.raises(st.CprocCudaError):                 # an input node has no state to output
        p.output(i)
    a = p.apply("acc", i)
    with pytest.raises(st.CprocCudaError):                 # no output node yet
        p.tick(4, np.zeros((8, 1, 4), np.uint32), np.zeros((8, 4), np.uint32))
    p.output(a)
    with pytest.raises(st.CprocCudaError):                 # params are not stored (as in the reference)
        p.get(a, 0, 0, kind=0)
    with pytest.raises(st.CprocCudaError):
        p.get(a, 1, 0)                                     # acc has one state field
    with pytest.raises(st.CprocCudaError):
        p.get(a, 0, 8)                                     # instance out of range
    for _ in range(63):
        a = p.apply("acc", a)
    with pytest.raises(st.CprocCudaError) as e:            # alloc_fail (:94-97)
        p.apply("acc", a)
    assert e.value.code == st.abi.ENOMEM
    p.close()
