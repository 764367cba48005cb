"""CPU: the SASS of the built sm_100a objects carries the instructions DESIGN.md claims (read with cuobjdump,
no GPU needed) -- TMA + mbarrier in the stencil pipeline, 256-bit global accesses in the streaming kernels,
warp shuffles in the reductions and the shuffle-based stencil, and NO fused multiply-add wherever the
reference's arithmetic is `a * b` then `+ c` with two roundings (the float MUL / ADD / SUB / mul-add
elementwise kernels and every stencil kernel)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ph-core_b200", "build")


def _functions(obj):
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        pytest.skip(f"{obj} has not been built")
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, timeout=600).stdout
    out, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and "/*" in line:
            ins = re.sub(r"/\*\s*[0-9a-fx]+\s*\*/", "", line).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            if ins:
                out[name].append(ins.split()[0])
    return out


def _count(fn_ops, pattern):
    return sum(1 for op in fn_ops if re.match(pattern, op))


def test_stencil_pipeline_uses_tma_and_mbarriers_and_never_contracts():
    fns = _functions("heat_tma.o")
    tma = {n: ops for n, ops in fns.items() if "heat_tma" in n}
    assert len(tma) >= 4
    for name, ops in tma.items():
        assert _count(ops, r"UTMALDG\.3D") >= 1, name                       # cp.async.bulk.tensor.3d
        assert _count(ops, r"SYNCS\.ARRIVE\.TRANS64") >= 1, name            # mbarrier.arrive.expect_tx
        assert _count(ops, r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT") >= 1, name  # mbarrier.try_wait.parity
    for obj in ("heat_tma.o", "heat.o"):
        for name, ops in _functions(obj).items():
            assert _count(ops, r"[FD]FMA\b") == 0, f"{obj}:{name} contracts a multiply-add"
            assert _count(ops, r"[FD]MUL\b|[FD]ADD\b|FADD2|FMUL2") > 0 or "heat" not in name
    assert any(_count(ops, r"SHFL\.(UP|DOWN)") for ops in _functions("heat.o").values())   # x-neighbours by shuffle


@pytest.mark.parametrize("obj,tag", [("ewise_f32.o", "If"), ("ewise_f64.o", "Id")])
def test_float_add_sub_mul_and_muladd_kernels_have_two_roundings(obj, tag):
    fns = _functions(obj)
    picked = {n: ops for n, ops in fns.items()
              if re.search(rf"BinaryOp{tag}Li[012]E", n) or f"MulAddOp{tag}" in n}        # PH_ADD, PH_SUB, PH_MUL, mul_add
    assert len(picked) >= 16
    for name, ops in picked.items():
        assert _count(ops, r"[FD]FMA\b") == 0, f"{name} contracts a multiply-add"
    muladd = [ops for n, ops in picked.items() if "MulAddOp" in n]
    mul, add = ("FMUL", "FADD") if tag == "If" else ("DMUL", "DADD")
    assert all(_count(ops, mul + r"\b") and _count(ops, add + r"\b") for ops in muladd)    # one rounding each


def test_streaming_kernels_move_256_bit_groups():
    flat = {n: ops for n, ops in _functions("ewise_f32.o").items() if "map_flat_kernel" in n and "Li8ELi2E" in n}
    assert flat
    for name, ops in flat.items():
        assert _count(ops, r"LDG\.E\..*256") >= 1, name
        if "CompareOp" not in name:                                     # a comparison stores 8 Bool bytes per group
            assert _count(ops, r"STG\.E\..*256") >= 1, name
    copy = _functions("copy.o")
    assert sum(_count(ops, r"LDG\.E\..*256") for ops in copy.values()) > 50
    red = _functions("reduce_f32.o")
    assert sum(_count(ops, r"LDG\.E\..*256") for ops in red.values()) > 50
    assert any(_count(ops, r"SHFL\.(DOWN|BFLY)") for n, ops in red.items() if "partial_kernel" in n)   # warp-level combine
