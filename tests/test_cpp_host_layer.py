"""The C++ host layer above the C-ABI (include/ph_narray.hpp: Phase::DeviceNArray / DeviceView with
the reference's method names and exception classes) and its spec program tests/cpp/device_narray_spec.cpp,
which replays the reference's own specs for the path (spec/n_array_spec.cr, spec/multi_writable_spec.cr,
spec/index_region_spec.cr, README.md) with the goldens typed in -- no oracle is linked into it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "device_narray_spec")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)


def test_host_only_specs_pass_without_a_gpu():
    """Region literal table at bound 10, error classes, trim!/translate!/reverse!: pure ph_host.h code."""
    _build()
    out = subprocess.run([EXE, "--host-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout and "FAIL" not in out.stdout


def test_spec_program_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=60)
    assert out.returncode == 3 and "no CPU fallback" in out.stderr


def test_spec_program_links_no_oracle():
    _build()
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libphgpu.so" in out and "oracle" not in out
    src = open(os.path.join(ROOT, "include", "ph_narray.hpp")).read()
    assert "oracle" not in src and "#include <cuda" not in src       # host code over the C-ABI only


@pytest.mark.gpu
def test_reference_specs_replayed_through_the_cpp_host_layer():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert " 0 failed" in out.stdout and "FAIL" not in out.stdout
    launches = int(out.stdout.rsplit(",", 1)[1].split()[0])
    assert launches > 100                                            # the specs ran on the device


def test_binary_dump_is_interchangeable_between_the_cpp_and_python_hosts(tmp_path):
    """f-4: the C++ host layer (include/ph_narray_io.hpp) and the Python mirror (ph_core_b200/io.py) read each
    other's binary checkpoints and agree on the JSON form (host halves only: no GPU needed)."""
    import json
    import numpy as np
    from ph_core_b200 import io
    _build()
    # C++ writes, Python reads
    p1 = str(tmp_path / "from_cpp.phbin")
    assert subprocess.run([EXE, "--write-dump", p1]).returncode == 0
    got = io.read_dump(p1)
    want = (0.25 * np.arange(12, dtype=np.float32) - 1.0).reshape(3, 4)
    assert got.dtype == np.float32 and got.shape == (3, 4) and got.tobytes() == want.tobytes()
    # Python writes, C++ reads and prints the reference's JSON form
    rs = np.random.RandomState(3)
    edge = [0.1, 2.0, -1.5e-7, 1e22, 5e-324, 0.30000000000000004, 1e14, 1e15, 1e16, 1e-4, 1e-5, -0.0, 123456.789e3]
    cases = [("f64", rs.rand(4, 3, 2)), ("f32", rs.rand(5, 7).astype(np.float32)),
             ("f64", np.array(edge)), ("f32", np.array([0.1, 3.4e38, 1e-38, 16777216.0, 1e16, -2.5e-7], np.float32)),
             ("i16", rs.randint(-30000, 30000, size=(6,)).astype(np.int16)),
             ("u8", (rs.rand(2, 9) < 0.5)), ("f32", np.zeros((3, 0, 2), np.float32))]
    for tag, arr in cases:
        p2 = str(tmp_path / f"from_py_{tag}.phbin")
        io.write_dump(arr, p2)
        out = subprocess.run([EXE, "--read-dump", p2, tag], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        obj = json.loads(out.stdout)
        assert obj["shape"] == list(arr.shape)
        if tag != "u8":                                   # (Bool arrays are read back as UInt8 by the CLI)
            assert out.stdout.strip() == io.host_to_json(arr), "the two hosts write different text"
        back = np.array(obj["elements"], dtype=arr.dtype if arr.dtype != np.bool_ else np.uint8).reshape(arr.shape)
        assert back.tobytes() == np.ascontiguousarray(arr).view(back.dtype).tobytes()      # shortest round-trip text is exact
    # a dump of the wrong element type, a truncated one and a foreign file are refused
    p3 = str(tmp_path / "f64.phbin")
    io.write_dump(rs.rand(3), p3)
    assert subprocess.run([EXE, "--read-dump", p3, "f32"], capture_output=True).returncode == 4
    raw = open(p3, "rb").read()
    open(p3, "wb").write(raw[:-5])
    assert subprocess.run([EXE, "--read-dump", p3, "f64"], capture_output=True).returncode == 4
    open(p3, "wb").write(b"not a dump at all\\n")
    assert subprocess.run([EXE, "--read-dump", p3, "f64"], capture_output=True).returncode == 4
