"""ctypes binding of libphgpu.so (include/ph_gpu.h).  Plumbing only: every compute call
goes to the CUDA library; if the library is missing or fails, this raises -- there is
no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libphgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ph_gpu.h")

PH_MAX_RANK = 8


class PhDesc(C.Structure):
    """struct ph_desc (include/ph_gpu.h)"""
    _fields_ = [("rank", C.c_int32), ("_pad", C.c_int32), ("offset", C.c_int64),
                ("extent", C.c_int64 * PH_MAX_RANK), ("stride", C.c_int64 * PH_MAX_RANK)]

    @classmethod
    def make(cls, extents, strides, offset=0) -> "PhDesc":
        d = cls()
        if len(extents) > PH_MAX_RANK:
            raise ValueError(f"rank {len(extents)} exceeds PH_MAX_RANK")
        d.rank = len(extents)
        d.offset = int(offset)
        for i, (e, s) in enumerate(zip(extents, strides)):
            d.extent[i] = int(e)
            d.stride[i] = int(s)
        return d

    @classmethod
    def contiguous(cls, shape) -> "PhDesc":
        strides, acc = [], 1
        for e in reversed(list(shape)):
            strides.append(acc)
            acc *= int(e)
        return cls.make(list(shape), list(reversed(strides)))


class PhRangeLit(C.Structure):
    """struct ph_range_lit (include/ph_host.h)"""
    _fields_ = [("is_index", C.c_int32), ("has_first", C.c_int32), ("has_last", C.c_int32),
                ("has_step", C.c_int32), ("exclusive", C.c_int32),
                ("first", C.c_int64), ("last", C.c_int64), ("step", C.c_int64)]


class PhRegion(C.Structure):
    """struct ph_region (include/ph_host.h)"""
    _fields_ = [("rank", C.c_int32), ("drop", C.c_int32),
                ("first", C.c_int64 * PH_MAX_RANK), ("step", C.c_int64 * PH_MAX_RANK),
                ("last", C.c_int64 * PH_MAX_RANK), ("proper_shape", C.c_int64 * PH_MAX_RANK),
                ("degeneracy", C.c_int32 * PH_MAX_RANK), ("reduced_rank", C.c_int32),
                ("reduced_shape", C.c_int64 * PH_MAX_RANK)]

    @property
    def shape(self):
        return [int(self.reduced_shape[i]) for i in range(self.reduced_rank)]


class PhSlicePeer(C.Structure):
    """struct ph_slice_peer (include/ph_host.h)"""
    _fields_ = [("send", PhDesc), ("land", PhDesc), ("recv0", C.c_int64), ("recv1", C.c_int64)]


class PhSlicePlan(C.Structure):
    """struct ph_slice_plan (include/ph_host.h)"""
    _fields_ = [("local", C.c_int32), ("dims", C.c_int32), ("new_shape", C.c_int64 * PH_MAX_RANK),
                ("my_new_rows", C.c_int64 * 2)]


class PhSlab(C.Structure):
    """struct ph_slab (include/ph_host.h)"""
    _fields_ = [("start", C.c_int64), ("stop", C.c_int64), ("count", C.c_int64), ("local_planes", C.c_int64),
                ("ghost", C.c_int32), ("lo_rank", C.c_int32), ("hi_rank", C.c_int32), ("_pad", C.c_int32)]


class PhTransposePeer(C.Structure):
    """struct ph_transpose_peer (include/ph_host.h)"""
    _fields_ = [("send0", C.c_int64), ("send1", C.c_int64), ("recv0", C.c_int64), ("recv1", C.c_int64),
                ("send_shape", C.c_int64 * PH_MAX_RANK), ("recv_shape", C.c_int64 * PH_MAX_RANK)]


class PhTransposePlan(C.Structure):
    """struct ph_transpose_plan (include/ph_host.h)"""
    _fields_ = [("local", C.c_int32), ("dims", C.c_int32), ("k", C.c_int32), ("j", C.c_int32),
                ("new_shape", C.c_int64 * PH_MAX_RANK), ("my_rows", C.c_int64 * 2), ("my_new_rows", C.c_int64 * 2)]


HOST_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ph_host.h")


def header_constants() -> dict:
    """Parse the enum constants out of ph_gpu.h so Python never re-declares them."""
    text = open(HEADER_PATH).read() + open(HOST_HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    consts = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", text, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [x.strip() for x in item.split("=")]
                nxt = int(val, 0)
            else:
                name = item
            consts[name] = nxt
            nxt += 1
    return consts


def header_functions() -> list:
    """Names of every function ph_gpu.h and ph_host.h declare."""
    text = open(HEADER_PATH).read() + open(HOST_HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ph_[a-z0-9_]+)\s*\(", text)))


K = header_constants()
globals().update(K)


class PhError(RuntimeError):
    """A non-zero status from libphgpu (CUDA / NCCL / argument failure)."""


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PhError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u8p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    dp = C.POINTER(PhDesc)
    sig = {
        "ph_init": [i32], "ph_shutdown": [], "ph_device_count": [C.POINTER(i32)], "ph_sm_count": [C.POINTER(i32)],
        "ph_alloc": [C.c_size_t, C.POINTER(vp)], "ph_free": [vp],
        "ph_h2d": [vp, vp, C.c_size_t], "ph_d2h": [vp, vp, C.c_size_t], "ph_d2h_async": [vp, vp, C.c_size_t],
        "ph_d2h_flags": [vp, vp, C.c_size_t, C.POINTER(C.c_uint32)], "ph_d2d": [vp, vp, C.c_size_t],
        "ph_host_alloc": [C.c_size_t, C.POINTER(vp)], "ph_host_free": [vp],
        "ph_sync": [], "ph_set_stream": [vp], "ph_stream_create": [C.POINTER(vp)], "ph_stream_destroy": [vp],
        "ph_stream_wait": [vp, vp], "ph_stream_sync": [vp], "ph_free_on": [vp, vp],
        "ph_checksum64": [vp, C.c_size_t, C.c_uint64, C.POINTER(C.c_uint64)],
        "ph_compare3": [i32, vp, dp, vp, dp, vp, dp], "ph_compare3_scalar": [i32, vp, dp, vp, i32, vp, dp], "ph_take_arith_flags": [C.POINTER(C.c_uint32)],
        "ph_timer_start": [], "ph_timer_stop": [C.POINTER(C.c_float)],
        "ph_ewise_binary": [i32, i32, vp, dp, vp, dp, vp, dp],
        "ph_ewise_scalar": [i32, i32, vp, dp, vp, i32, vp, dp],
        "ph_ewise_unary": [i32, i32, vp, dp, vp, dp],
        "ph_ewise_mul_add": [i32, vp, dp, vp, dp, vp, dp, vp, dp],
        "ph_compare": [i32, i32, vp, dp, vp, dp, u8p, dp],
        "ph_compare_scalar": [i32, i32, vp, dp, vp, i32, u8p, dp],
        "ph_mask_set_scalar": [i32, vp, dp, u8p, dp, vp],
        "ph_mask_set_array": [i32, vp, dp, u8p, dp, vp, dp],
        "ph_copy_strided": [i32, vp, dp, vp, dp],
        "ph_fill_region": [i32, vp, dp, vp],
        "ph_reduce_full": [i32, i32, vp, dp, vp, C.POINTER(i64)],
        "ph_reduce_full_dev": [i32, i32, vp, dp, vp, vp],
        "ph_reduce_axis": [i32, i32, vp, dp, i32, vp, dp],
        "ph_heat_step": [i32, i32, C.POINTER(i64), vp, i32, vp, vp],
        "ph_heat_run": [i32, i32, C.POINTER(i64), vp, i32, vp, vp, i64, C.POINTER(i32)],
        "ph_heat_step_slab": [i32, i32, C.POINTER(i64), vp, i32, i32, i64, i64, vp, vp, vp],
        "ph_heat_pass_slab": [i32, i32, C.POINTER(i64), vp, i32, i32, i32, i32, i64, i64, vp, vp, vp],
        "ph_comm_unique_id": [vp], "ph_comm_init": [i32, i32, vp], "ph_comm_destroy": [],
        "ph_allreduce": [i32, i32, vp, i64], "ph_allgather": [vp, vp, i64],
        "ph_comm_p2p_ready": [C.POINTER(i32)], "ph_symm_alloc": [C.c_size_t, C.POINTER(vp)], "ph_symm_free": [vp],
        "ph_symm_peer": [vp, i32, C.POINTER(vp)],
        "ph_reduce_full_sharded": [i32, i32, vp, dp, i64, vp, C.POINTER(i64), C.POINTER(C.c_uint32)],
        "ph_alltoallv": [C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64)],
        "ph_alltoall_strided": [i32, vp, dp, vp, dp],
        "ph_halo_exchange": [vp, vp, i32, vp, vp, i32, i64, vp],
        "ph_heat_run_sharded": [i32, i32, C.POINTER(i64), vp, i32, vp, vp, i64, C.POINTER(i32)],
    }
    i64p, i32p, rp, lp = C.POINTER(i64), C.POINTER(i32), C.POINTER(PhRegion), C.POINTER(PhRangeLit)
    sig.update({
        "ph_canonicalize_range": [lp, i64, i64p, i64p, i64p, i64p],
        "ph_canonicalize_coord": [i64p, i32, i64p, i32, i64p],
        "ph_region_new": [lp, i32, i64p, i32, i32, rp],
        "ph_region_cover": [i64p, i32, i32, rp],
        "ph_region_new_trimmed": [lp, i32, i64p, i64p, i32, i32, rp],
        "ph_region_fits_in": [rp, i64p, i32, i32p],
        "ph_region_trim": [rp, i64p, i32], "ph_region_reverse": [rp], "ph_region_translate": [rp, i64p, i32],
        "ph_shapes_compatible": [i64p, i32, i64p, i32, i32p],
        "ph_broadcast_shapes": [i64p, i64p, i32, i64p],
        "ph_concat_shape": [i64p, i32p, i32, i32, i64p, i32p],
        "ph_row_chunks": [i64, i64, i32, i32, i64p, i32, i32p],
        "ph_desc_contiguous": [i64p, i32, dp], "ph_desc_region": [dp, rp, dp],
        "ph_desc_permute": [dp, i32p, i32, dp], "ph_desc_reverse": [dp, dp],
        "ph_desc_reshape": [dp, i64p, i32, dp], "ph_desc_broadcast": [dp, i64p, i32, dp],
        "ph_desc_offset_of": [dp, i64p, i32, i64p],
        "ph_shard_range": [i64, i32, i32, i64p, i64p],
        "ph_slab_layout": [i64, i32, i32, i32, C.POINTER(PhSlab)],
        "ph_transpose_plan_of": [i64p, i32, i32p, i32, i32, C.POINTER(PhTransposePlan), C.POINTER(PhTransposePeer)],
        "ph_slice_plan_of": [i64p, i32, rp, i32, i32, C.POINTER(PhSlicePlan), C.POINTER(PhSlicePeer)],
        "ph_combine_extremum_records": [vp, i32, i32, i32, i32p, i64p],
    })
    for name, args in sig.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = i32
    if hasattr(lib, "ph_stream"):
        lib.ph_stream.restype = vp
        lib.ph_stream.argtypes = []
    if hasattr(lib, "ph_last_error_string"):
        lib.ph_last_error_string.restype = C.c_char_p
        lib.ph_last_error_string.argtypes = []
    if hasattr(lib, "ph_host_last_error"):
        lib.ph_host_last_error.restype = C.c_char_p
        lib.ph_host_last_error.argtypes = []
    for counter in ("ph_launch_count", "ph_nccl_call_count"):
        if hasattr(lib, counter):
            getattr(lib, counter).restype = i64
            getattr(lib, counter).argtypes = []
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ph_last_error_string()
        raise PhError(f"libphgpu status {status}: {msg.decode() if msg else ''}")


_inited_device = None


def init(device: int | None = None) -> int:
    """ph_init on `device` (default: LOCAL_RANK or 0).  Raises if there is no GPU."""
    global _inited_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device == device:
        return device
    check(load().ph_init(device))
    _inited_device = device
    return device
