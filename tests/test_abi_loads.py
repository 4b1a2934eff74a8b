"""CPU: the C-ABI library loads and exports every symbol include/cproc_cuda.h
declares; without a GPU the product refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cproc_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cproc_cuda_[a-z0-9_]+)\s*\(", src)) - {"cproc_cuda_chunk_fn"})


def test_every_declared_symbol_is_exported_and_bound():
    from synth_tools_b200 import abi
    names = _declared()
    assert len(names) >= 30
    raw = ctypes.CDLL(abi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libcproc_cuda.so does not export %s" % n
        assert n in abi.SYMBOLS, "abi.py does not bind %s" % n
    assert sorted(abi.SYMBOLS) == names
    assert abi.lib.cproc_cuda_abi_version() == 3


def test_header_is_plain_c99(tmp_path):
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "cproc_cuda.h"\nint main(void){ cproc_cuda_config c = {0}; cproc_cuda_io io = {0}; (void)c; (void)io; return sizeof(cproc_cuda_node) == 16 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=gnu99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_struct_mirrors_match_header(tmp_path):
    """ctypes Config / IO / Node must have the C sizes and offsets."""
    import subprocess
    from synth_tools_b200 import abi
    c = tmp_path / "s.c"
    c.write_text('#include <stdio.h>\n#include "cproc_cuda.h"\nint main(void){ printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(cproc_cuda_config), sizeof(cproc_cuda_io), sizeof(cproc_cuda_node), offsetof(cproc_cuda_config, voices_per_bus), offsetof(cproc_cuda_config, nodes), offsetof(cproc_cuda_io, layout)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(abi.Config), ctypes.sizeof(abi.IO), ctypes.sizeof(abi.Node), abi.Config.voices_per_bus.offset,
            abi.Config.nodes.offset, abi.IO.layout.offset]
    assert got == want


def test_no_cpu_fallback():
    """On a box without a GPU open() must fail with ENODEV, never compute on the CPU."""
    import torch
    from synth_tools_b200 import abi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert abi.lib.cproc_cuda_device_count() == 0
    with pytest.raises(abi.CprocCudaError) as e:
        abi.Context(0)
    assert e.value.code == abi.ENODEV
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The package must not reference oracle/ in any way."""
    pkg = os.path.join(ROOT, "synth_tools_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "libref" not in txt, f
