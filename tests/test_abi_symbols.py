"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; computing
without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

import ph_core_b200 as ph
from ph_core_b200 import _lib


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _lib.header_functions()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_enum_constants_parsed_from_header():
    assert ph.K["PH_OK"] == 0 and ph.K["PH_MAX"] == 2 and ph.K["PH_FLAG_ARGUMENT"] == 8
    assert ph.K["PH_HOST_NEEDS_COPY"] == 110
    assert C.sizeof(ph.PhDesc) == 8 + 8 + 8 * 8 + 8 * 8


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if ".cubin" in l}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert lib.ph_init(0) != 0                                  # loud failure
    d = ph.PhDesc.contiguous([4])
    assert lib.ph_ewise_binary(0, 0, 1, C.byref(d), 1, C.byref(d), 1, C.byref(d)) == ph.K["PH_ERR_NOT_INIT"]
    with pytest.raises(ph.PhError):
        ph.DeviceNArray.fill([4], 1.0, "float32")


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ph-core_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle's own R/Step objects", ""), f"{f} mentions the oracle"
