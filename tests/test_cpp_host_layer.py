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
