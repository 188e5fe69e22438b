"""Randomised differential test (tools/stress.py): random shapes x data regimes x packers, GPU stream ==
oracle stream byte for byte, decode with the encoder's index and with a rebuilt one, verify."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_shapes_and_regimes(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress.py"), "80", str(seed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "stress ok: 80 cases" in r.stdout
