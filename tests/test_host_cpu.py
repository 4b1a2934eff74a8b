"""CPU tests of host-side pieces that need no GPU: the restated word clock against its closed form and the
Pd [scale] object (synth_tools.c:147-194) through the scripted Pd stand-in."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def oracle():
    import sys
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    return po.Oracle()


def test_word_clock_oracle_is_an_integer_divisor_square_wave(oracle):
    """clock.c:109-120 from phase 0, pol 1: out[t] = 1 ^ ((t // hperiod) & 1) for hperiod >= 1; the state carries
    over block boundaries; BPM_TO_HPERIOD(48000, 120) = 500 samples = 24 ppqn at 2 beats per second."""
    hp = np.array([1, 2, 3, 8, 62, 500, 1000], np.int32)
    N, F = len(hp), 4096
    state = np.zeros((N, 2), np.int32); state[:, 1] = 1
    a = oracle.word_clock_run(state, hp, N, F // 2)
    b = oracle.word_clock_run(state, hp, N, F // 2)
    got = np.concatenate([a, b], axis=1)
    t = np.arange(F)
    for k in range(N):
        assert np.array_equal(got[k], (1 ^ ((t // hp[k]) & 1)).astype(np.float32)), k
    assert (48000 * 5) // (120 * 4) == 500
    # hperiod 0: toggles every sample while the counter runs away, exactly as the C loop does
    state = np.array([[0, 1]], np.int32)
    z = oracle.word_clock_run(state, np.array([0], np.int32), 1, 8)
    assert list(z[0]) == [0, 1, 0, 1, 0, 1, 0, 1] and state[0, 0] == 8


def test_pd_scale_object(tmp_path):
    """[scale exp 20 20000] and [scale lin -1 3]: every MIDI value 0..127 gives base * powf(max/min, v/127) resp.
    base + (max-min) * v/127 in float arithmetic (bit-exact against libm's powf); unknown types are refused."""
    exe = str(tmp_path / "pd_scale_fake")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "c", "fakepd"),
                           os.path.join(ROOT, "synth_tools_b200", "host", "pd", "scale.c"),
                           os.path.join(ROOT, "tests", "c", "fakepd", "fakepd_scale.c"), "-o", exe, "-lm"])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = [l.split() for l in res.stdout.splitlines()]
    assert lines[0] == ["refused", "1"]
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float; libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    f32 = np.float32
    got = {(int(l[1]), int(l[2])): (f32(float.fromhex(l[3])), int(l[4])) for l in lines[1:]}
    assert len(got) == 256
    for v in range(128):
        frac = f32(v) * f32(1.0 / 127.0)
        want_exp = f32(20.0) * f32(libm.powf(f32(20000.0) / f32(20.0), frac))
        want_lin = f32(-1.0) + (f32(3.0) - f32(-1.0)) * frac
        assert got[(0, v)] == (want_exp, 0), v
        assert got[(1, v)] == (want_lin, 1), v
    assert got[(0, 0)][0] == 20.0 and got[(0, 127)][0] == 20000.0 and got[(1, 127)][0] == 3.0


@pytest.mark.parametrize("name,srcs,incs,libs", [
    ("jack_synth", ["synth_tools_b200/host/jack/jack_synth.c", "tests/c/fakejack/fakejack.c"], ["tests/c/fakejack", "synth_tools_b200/host/jack"], ["-lcproc_cuda"]),
    ("jack_clock", ["synth_tools_b200/host/jack/jack_clock.c", "tests/c/fakejack/fakejack.c"], ["tests/c/fakejack"], ["-lcproc_cuda"]),
    ("pd_square_grain", ["synth_tools_b200/host/pd/square_grain_b200~.c", "tests/c/fakepd/fakepd.c"], ["tests/c/fakepd"], ["-lcproc_cuda", "-lm"]),
    ("dropin_pd", ["tests/c/test_dropin_pd.c", "synth_tools_b200/host/dropin.c"], ["tests/c/fakepd"], ["-DCPROC_HAVE_PD", "-lcproc_cuda"]),
])
def test_host_adapters_build_warning_free(tmp_path, name, srcs, incs, libs):
    """The JACK / Pd adapters compile with -Wall -Werror and link against the in-tree C-ABI library (running them
    needs a GPU: tests/test_gpu_dropin.py)."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    if not os.path.exists(os.path.join(pkg, "libcproc_cuda.so")):
        pytest.skip("libcproc_cuda.so not built")
    cmd = ["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include")]
    for i in incs:
        cmd += ["-I", os.path.join(ROOT, i)]
    cmd += [os.path.join(ROOT, s) for s in srcs] + ["-o", str(tmp_path / name), "-L", pkg, "-Wl,-rpath," + pkg] + libs
    subprocess.check_call(cmd)


def test_dropin_exports_reference_names(tmp_path):
    """host/dropin.c built as a shared object exports the reference's own entry points: synth_run (synth.c:45),
    cproc_update (test_cproc.c:13), cproc_input (mod_cproc_plugin.c:20) and -- with m_pd.h on the include path --
    square_grain_proc (synth_tools.c:85-86).  Symbols only: calling them needs a GPU."""
    import ctypes
    pkg = os.path.join(ROOT, "synth_tools_b200")
    if not os.path.exists(os.path.join(pkg, "libcproc_cuda.so")):
        pytest.skip("libcproc_cuda.so not built")
    so = str(tmp_path / "libdropin_pd.so")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", "-DCPROC_HAVE_PD", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "c", "fakepd"), os.path.join(pkg, "host", "dropin.c"), "-o", so,
                           "-L", pkg, "-lcproc_cuda", "-Wl,-rpath," + pkg])
    lib = ctypes.CDLL(so)
    for name in ["synth_run", "cproc_update", "cproc_input", "square_grain_proc", "cproc_dropin_square_grain_proc", "cproc_dropin_handle_tag_u32",
                 "cproc_dropin_last_status", "cproc_dropin_shutdown"]:
        assert hasattr(lib, name), name
    shipped = ctypes.CDLL(os.path.join(pkg, "libcproc_dropin.so"))
    for name in ["synth_run", "cproc_update", "cproc_input", "cproc_dropin_square_grain_proc", "cproc_dropin_handle_tag_u32"]:
        assert hasattr(shipped, name), name


def test_closed_form_zero_state_response_math():
    """The derivation behind k_sweep_pre / k_sweep_zsr_closed, restated in numpy fp64: the zero-state
    response of the SVF to the wrapped phase ramp, advanced per phase WRAP (gaps of k or k+1 ticks from an integer
    remainder walk, the last stretch from the table of 2^i-tick maps), equals the ticked recurrence on the exact
    integer input to 1e-9 of full scale -- for increments with none, few and many wraps, powers of two and >= 2^31."""
    rng = np.random.default_rng(11)
    L = 544
    nlev = L.bit_length()

    def seg_then(x, y):                      # x ticks, then y ticks: (A, P, Q, m)
        return (y[0] @ x[0], y[0] @ x[1] + y[1], y[0] @ x[2] + x[3] * y[1] + y[2], x[3] + y[3])

    for inc in [0, 1, 3, 1 << 20, (1 << 24) + 1, 39370533, 715827883, 0x55555556, 1 << 31, (1 << 31) + 5, 0xFFFFFFFF, 1 << 28, 3 << 28]:
        f, q = float(np.float32(rng.uniform(0.01, 0.3))), float(np.float32(rng.uniform(0.5, 2.0)))
        A = np.array([[1.0, f], [-f, 1.0 - f * q - f * f]]); b = np.array([0.0, f])
        lev = (A, b, np.zeros(2), 1.0)
        ident = (np.eye(2), np.zeros(2), np.zeros(2), 0.0)
        table, sl, sk = [], ident, ident
        k = rho = 0
        if inc:
            k = 0xFFFFFFFF // inc; rho = (-k * inc) & 0xFFFFFFFF
            if rho == inc:
                k += 1; rho = 0
        for l in range(nlev):
            table.append(lev)
            if (L >> l) & 1:
                sl = seg_then(sl, lev)
            if inc and k < L and (k >> l) & 1:
                sk = seg_then(sk, lev)
            lev = seg_then(lev, lev)
        for ph in [0, 0x7FFFFFF0, 0x80000000, int(rng.integers(0, 2**32)), 0xFFFFFFFF]:
            # ticked reference on the exact integer input
            s = np.zeros(2); p = ph
            for _ in range(L):
                x = float(p - (1 << 32) if p >= (1 << 31) else p)
                lp = s[0] + f * s[1]; s = np.array([lp, s[1] + f * (x - q * s[1] - lp)]); p = (p + inc) & 0xFFFFFFFF
            # closed form
            u0 = ph ^ 0x80000000
            W = (u0 + (L - 1) * inc) >> 32
            y = np.zeros(2)
            if W:
                at = (~u0 & 0xFFFFFFFF) // inc + 1
                if W > 1:
                    r = (u0 + at * inc) & 0xFFFFFFFF
                    for j in range(1, W):
                        y = sk[0] @ y + j * sk[1]; at += k
                        if r >= rho:
                            r -= rho
                        else:
                            r = r - rho + inc; at += 1
                            y = A @ y + j * b
                e, l = L - at, 0
                assert e >= 1
                while e:
                    if e & 1:
                        y = table[l][0] @ y + W * table[l][1]
                    e >>= 1; l += 1
            x0 = float(ph - (1 << 32) if ph >= (1 << 31) else ph)
            got = sl[1] * x0 + sl[2] * inc - 4294967296.0 * y
            assert np.abs(got - s).max() <= 1e-9 * 2.0**31 * max(1.0, np.abs(s).max() / 2.0**31), (inc, ph, got, s)
