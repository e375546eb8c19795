"""CPU, build container only: the committed tests/golden/reference_python.npz is exactly what
tests/golden/make_reference_golden.py produces from the reference's source tree (skipped where /root/reference
does not exist, e.g. on the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/lib"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_python_fixture_is_reproducible(tmp_path):
    out = str(tmp_path / "regen.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_reference_golden.py"), out],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    a = np.load(os.path.join(ROOT, "tests", "golden", "reference_python.npz"))
    b = np.load(out)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), k
