#!/usr/bin/env python
"""Trimmed SASS listings of the kernels tests/test_sass_claims.py makes claims about, readable without
rebuilding: per kernel the opcode histogram and the instruction stream (addresses and encodings stripped;
long listings cut to their first `--lines` instructions, which always include the hot loop's first body).
    python scripts/sass_listing.py            # writes profiles/r02_sass_*.txt from ph-core_b200/build/*.o"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ph-core_b200", "build")
WANT = [
    ("ewise_f32.o", r"map_flat_kernelINS_8BinaryOpIfLi0EEELi8ELi2E", "r02_sass_map_flat_add_f32.txt", 400,
     "out = a + c, contiguous f32, 32-byte groups x 2 in flight: 256-bit LDG / STG, FADD, no FFMA"),
    ("ewise_f32.o", r"map_flat_kernelINS_8MulAddOpIfEELi8ELi2E", "r02_sass_map_flat_muladd_f32.txt", 400,
     "fused (a * b) + c with TWO roundings: FMUL then FADD, never FFMA"),
    ("copy.o", r"transpose_kernelImE", "r02_sass_transpose_u64.txt", 600,
     "32x33 shared-memory tile transpose of 8-byte elements"),
    ("heat_tma.o", r"heat_tma2_kernelIfLi16ELi6ELb0ELb1E", "r02_sass_heat_tma2_f32.txt", 1400,
     "two time steps per pass: UTMALDG.3D + mbarrier (SYNCS) ring, FADD / FMUL only (no FFMA), one BAR.SYNC per plane"),
    ("reduce_f64.o", r"axis_strip_staged_kernelIdLi0ELi16ELi8E", "r02_sass_axis_strip_staged_sum_f64.txt", 500,
     "few-column ordered fold: LDGSTS.E.BYPASS.128 (cp.async) ring, LDS + DADD in k order"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles"))
    args = ap.parse_args()
    for obj, pat, name, limit, what in WANT:
        text = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        blocks = re.split(r"\n\s*Function : ", text)
        hit = [b for b in blocks if re.match(r"\S*" + pat, b)]
        if not hit:
            print("not found:", obj, pat)
            continue
        b = hit[0]
        fn = b.split("\n", 1)[0].strip()
        ins = []
        for line in b.splitlines():
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
            if m:
                ins.append(m.group(2).strip())
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins)
        with open(os.path.join(args.out, name), "w") as f:
            f.write(f"# {fn}\n# {what}\n# object: ph-core_b200/build/{obj} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false)\n")
            f.write(f"# {len(ins)} instructions; opcode histogram:\n")
            for op, n in hist.most_common():
                f.write(f"#   {n:6d}  {op}\n")
            f.write(f"# ---- instruction stream (first {min(limit, len(ins))} of {len(ins)})\n")
            for i in ins[:limit]:
                f.write(i + "\n")
        print(name, len(ins), "instructions")


if __name__ == "__main__":
    main()
