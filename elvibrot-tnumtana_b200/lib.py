"""ctypes binding of libevr_sg4.so (the C-ABI declared in include/evr_sg4.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded from the package
directory.  There is no fallback of any kind: if the shared object is missing or a CUDA call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("EVR_SG4_LIB") or os.path.join(_HERE, "libevr_sg4.so")   # EVR_SG4_LIB: experiment builds
_lib = None

# enum values of include/evr_sg4.h
TAB_NB_SG, TAB_NB, TAB_S, TAB_NQ, TAB_COUNT0, TAB_LMIN = 0, 1, 2, 3, 4, 5
TAB_TAB_L, TAB_WEIGHT, TAB_TAB_NQ, TAB_TAB_NB, TAB_SUM_NQ, TAB_SUM_NB, TAB_PACKEDB, TAB_MAP = 10, 11, 12, 13, 14, 15, 16, 17
INFO_LAUNCHES, INFO_ALG_BYTES_NPSI1, INFO_ALG_BYTES_PER_RHS_EXTRA, INFO_NQ_LOCAL, INFO_S_LOCAL = 0, 1, 2, 3, 4
FLAG_WORDS = 2 * 16 + 2      # EVR_SG4_FLAG_WORDS (include/evr_sg4_comm.h)
INFO_SMEM_BYTES, INFO_GRID_CTAS, INFO_PATH, INFO_FLOPS_NPSI1, INFO_ISO, INFO_DEVICES, INFO_GENERIC_TERMS = 5, 6, 7, 8, 9, 10, 11

EXPORTS = [
    "evr_sg4_version", "evr_sg4_last_error",
    "evr_sg4_tables_build", "evr_sg4_tables_destroy", "evr_sg4_tables_size", "evr_sg4_tables_get",
    "evr_sg4_ini_iGs", "evr_sg4_balanced_iGs",
    "evr_sg4_plan_create", "evr_sg4_plan_create_ex", "evr_sg4_device_count", "evr_sg4_plan_set_op", "evr_sg4_plan_set_op10", "evr_sg4_apply", "evr_sg4_apply_device", "evr_sg4_apply_device_scaled",
    "evr_sg4_plan_info", "evr_sg4_plan_destroy", "evr_sg4_model_grid",
    "evr_sg4_BtoG", "evr_sg4_GtoB", "evr_sg4_DerivOp_G", "evr_sg4_BtoG_device", "evr_sg4_GtoB_device", "evr_sg4_DerivOp_G_device",
    "evr_sg4_allreduce_slices", "evr_sg4_allreduce_fused", "evr_sg4_allgather_slices", "evr_sg4_reduce_slice", "evr_sg4_reduce_to", "evr_sg4_slice_bounds",
    "evr_sg4_set_devices", "evr_sg4_get_devices", "evr_sg4_host_register", "evr_sg4_host_unregister",
    "evr_sg4_vec_alloc", "evr_sg4_vec_free", "evr_sg4_vec_upload", "evr_sg4_vec_download", "evr_sg4_vec_gram", "evr_sg4_vec_lincomb",
    "evr_sg4_vec_scale", "evr_sg4_vec_precond", "evr_sg4_vec_schmidt",
]


class EvrSg4Error(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ into libevr_sg4.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcdir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(srcdir, f) for f in os.listdir(srcdir)] + [os.path.join(_HERE, "..", "include", "evr_sg4.h")]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-j", "8", "-C", srcdir] + (["-B"] if force else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
        if out.returncode != 0:
            raise EvrSg4Error("nvcc build of libevr_sg4.so failed")
    return SO_PATH


def lib():
    """Load the shared object; it is built first only when it is missing (``build()`` is the staleness-aware entry:
    __graft_entry__.build() calls it; a GPU box gets the prebuilt library with the snapshot and must not rebuild)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        build()
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.evr_sg4_version.restype = i32
    L.evr_sg4_last_error.restype = C.c_char_p
    L.evr_sg4_tables_build.restype = i32
    L.evr_sg4_tables_build.argtypes = [C.POINTER(vp), i32, i32, i32, vp, vp]
    L.evr_sg4_tables_destroy.argtypes = [C.POINTER(vp)]
    L.evr_sg4_tables_size.restype = i64
    L.evr_sg4_tables_size.argtypes = [vp, i32]
    L.evr_sg4_tables_get.restype = i32
    L.evr_sg4_tables_get.argtypes = [vp, i32, vp]
    L.evr_sg4_ini_iGs.restype = i32
    L.evr_sg4_ini_iGs.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.evr_sg4_balanced_iGs.restype = i32
    L.evr_sg4_balanced_iGs.argtypes = [i32, vp, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.evr_sg4_plan_create.restype = i32
    L.evr_sg4_plan_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i64, i32] + [vp] * 11 + [i32, i32]
    L.evr_sg4_plan_create_ex.restype = i32
    L.evr_sg4_plan_create_ex.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i64, i32] + [vp] * 5 + [i64, i64] + [vp] * 6 + [i32, i32]
    L.evr_sg4_device_count.restype = i32
    L.evr_sg4_plan_set_op.restype = i32
    L.evr_sg4_plan_set_op.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.evr_sg4_plan_set_op10.restype = i32
    L.evr_sg4_plan_set_op10.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.evr_sg4_apply.restype = i32
    L.evr_sg4_apply.argtypes = [vp, i32, vp, vp]
    L.evr_sg4_apply_device.restype = i32
    L.evr_sg4_apply_device.argtypes = [vp, i32, vp, vp, vp]
    L.evr_sg4_apply_device_scaled.restype = i32
    L.evr_sg4_apply_device_scaled.argtypes = [vp, i32, vp, vp, C.c_double, C.c_double, vp]
    for name, at in (("evr_sg4_BtoG", [vp, i32, vp, vp]), ("evr_sg4_GtoB", [vp, i32, vp, vp]),
                     ("evr_sg4_DerivOp_G", [vp, i32, vp, vp, i32, i32]), ("evr_sg4_BtoG_device", [vp, i32, vp, vp, vp]),
                     ("evr_sg4_GtoB_device", [vp, i32, vp, vp, vp]), ("evr_sg4_DerivOp_G_device", [vp, i32, vp, i32, i32, vp])):
        getattr(L, name).restype = i32
        getattr(L, name).argtypes = at
    L.evr_sg4_model_grid.restype = i32
    L.evr_sg4_model_grid.argtypes = [i32, i32, i32, i32, vp, vp, vp, i32, vp, i32, i32, vp]
    L.evr_sg4_plan_info.restype = i64
    L.evr_sg4_plan_info.argtypes = [vp, i32]
    L.evr_sg4_plan_destroy.restype = i32
    L.evr_sg4_plan_destroy.argtypes = [C.POINTER(vp)]
    L.evr_sg4_allreduce_slices.restype = i32
    L.evr_sg4_allreduce_slices.argtypes = [vp, i32, i32, i64, vp]
    L.evr_sg4_allreduce_fused.restype = i32
    L.evr_sg4_allreduce_fused.argtypes = [vp, vp, i32, i32, i64, C.c_uint64, vp]
    L.evr_sg4_allgather_slices.restype = i32
    L.evr_sg4_allgather_slices.argtypes = [vp, i32, i32, i64, vp]
    L.evr_sg4_reduce_slice.restype = i32
    L.evr_sg4_reduce_slice.argtypes = [vp, i32, i32, i64, vp]
    L.evr_sg4_reduce_to.restype = i32
    L.evr_sg4_reduce_to.argtypes = [vp, i32, i64, vp, vp]
    L.evr_sg4_slice_bounds.restype = i32
    L.evr_sg4_slice_bounds.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    f64 = C.c_double
    L.evr_sg4_vec_alloc.restype = i32
    L.evr_sg4_vec_alloc.argtypes = [C.POINTER(vp), i64, i32]
    L.evr_sg4_vec_free.restype = i32
    L.evr_sg4_vec_free.argtypes = [vp]
    L.evr_sg4_vec_upload.restype = i32
    L.evr_sg4_vec_upload.argtypes = [vp, vp, i64, vp]
    L.evr_sg4_vec_download.restype = i32
    L.evr_sg4_vec_download.argtypes = [vp, vp, i64, vp]
    L.evr_sg4_vec_gram.restype = i32
    L.evr_sg4_vec_gram.argtypes = [i64, i32, vp, i64, i32, vp, i64, vp, vp]
    L.evr_sg4_vec_lincomb.restype = i32
    L.evr_sg4_vec_lincomb.argtypes = [i64, i32, vp, i64, i32, vp, f64, vp, i64, vp]
    L.evr_sg4_vec_scale.restype = i32
    L.evr_sg4_vec_scale.argtypes = [i64, f64, vp, vp]
    L.evr_sg4_vec_precond.restype = i32
    L.evr_sg4_vec_precond.argtypes = [i64, vp, vp, f64, f64, vp]
    L.evr_sg4_vec_schmidt.restype = i32
    L.evr_sg4_vec_schmidt.argtypes = [i64, i32, vp, i64, vp, C.POINTER(f64), vp]
    L.evr_sg4_set_devices.restype = i32
    L.evr_sg4_set_devices.argtypes = [i32]
    L.evr_sg4_get_devices.restype = i32
    L.evr_sg4_host_register.restype = i32
    L.evr_sg4_host_register.argtypes = [vp, i64]
    L.evr_sg4_host_unregister.restype = i32
    L.evr_sg4_host_unregister.argtypes = [vp]
    _lib = L
    return L


def set_devices(ndev: int):
    """Plans created afterwards on the default device span the first ``ndev`` GPUs of the node (one process, NVLink peer
    memory; C-ABI evr_sg4_set_devices)."""
    check(lib().evr_sg4_set_devices(int(ndev)), "evr_sg4_set_devices")


def slice_bounds(n: int, np_: int, rank: int):
    lo, hi = C.c_int64(), C.c_int64()
    check(lib().evr_sg4_slice_bounds(int(n), int(np_), int(rank), C.byref(lo), C.byref(hi)), "evr_sg4_slice_bounds")
    return lo.value, hi.value


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().evr_sg4_last_error().decode(errors="replace")
        raise EvrSg4Error(f"{what}: {msg}" if what else msg)
