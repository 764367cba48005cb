"""GPU: the C++ host example (examples/heat_equation.cpp -- the reference's
examples/heat_equation.cr over the C-ABI) reproduces the oracle's replay of the example."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_example_matches_oracle():
    from oracle import ph_oracle as O
    exe = os.path.join(ROOT, "examples", "heat_equation")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    body = out.stdout[out.stdout.index("[") + 1: out.stdout.index("]")]
    got = np.array([float(v) for v in body.split(",")])
    want = O.heat_simulate_1d_example()
    assert got.tobytes() == want.tobytes()                     # %.17g round-trips f64: bit-exact
    assert "COEFF = 0x1.9912f7d0247d5p-12" in out.stdout
    assert re.search(r"launches = [1-9]", out.stdout) and "flags = 0" in out.stdout
    assert abs(got.sum() - 480.0) < 1e-9
