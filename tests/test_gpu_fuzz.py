"""GPU: randomised parity sweep (tools/fuzz_parity.py) with fixed seeds -- random shapes, counters, layouts
and kernel generations for PDM v2 / v1, square_grain (+ mix), the voice bank, the xvoice mix and generated
graphs, every case against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_parity(seed):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), str(seed), "40"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "failures: 0" in res.stdout
