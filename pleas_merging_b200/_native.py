"""ctypes binding of libpleas_b200.so (include/pleas_b200.h).

No torch types cross this boundary: tensors are passed as ``data_ptr()`` integers, the
stream as ``torch.cuda.current_stream().cuda_stream``.  There is no CPU fallback: if the
library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpleas_b200.so")

c_i32, c_i64, c_f32, c_f64, c_ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class GemmProblem(ctypes.Structure):
    """Mirror of PlbGemmProblem (include/pleas_b200.h)."""
    _fields_ = [
        ("a_hi", c_ptr), ("a_lo", c_ptr), ("b_hi", c_ptr), ("b_lo", c_ptr), ("partial", c_ptr),
        ("a_row_groups", c_i32), ("b_row_groups", c_i32), ("k_blocks", c_i32),
        ("m_tiles", c_i32), ("n_tiles", c_i32), ("splits", c_i32), ("cta_begin", c_i32), ("symmetric", c_i32),
    ]


class FinalizeTap(ctypes.Structure):
    """Mirror of PlbFinalizeTap (include/pleas_b200.h)."""
    _fields_ = [("partial", c_ptr), ("qa", c_ptr), ("qb", c_ptr), ("sa", c_ptr), ("sb", c_ptr),
                ("ld_m", c_i64), ("ld_n", c_i64), ("K", c_i64), ("splits", c_i32), ("n_affine", c_i32),
                ("affine", c_ptr)]


class FinalizeGroup(ctypes.Structure):
    """Mirror of PlbFinalizeGroup (include/pleas_b200.h)."""
    _fields_ = [("cost", c_ptr), ("ldc", c_i64), ("n", c_i32), ("tap_begin", c_i32), ("tap_end", c_i32),
                ("block_begin", c_i32)]


_SIGNATURES = {
    "plb_version": (ctypes.c_int, []),
    "plb_last_error_string": (ctypes.c_char_p, []),
    "plb_plane_bytes": (c_i64, [c_i64, c_i64, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "plb_pack_split": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_i32,
                                      c_ptr, c_ptr, c_ptr]),
    "plb_pack_split_pair": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32,
                                           c_ptr, c_ptr, c_ptr]),
    "plb_gram_direct": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "plb_gram_tma": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr,
                                    c_ptr]),
    "plb_pack_split_pair_sums": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i32,
                                                c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "plb_cross_finalize_grouped": (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "plb_cross_finalize_corr": (ctypes.c_int, [c_ptr, c_i32, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                               c_i64, c_ptr, c_i64, c_i32, c_ptr]),
    "plb_gram_tma_geometry": (ctypes.c_int, [c_i64, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32),
                                             ctypes.POINTER(c_i32), ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "plb_debug_set_trace": (ctypes.c_int, [c_ptr]),
    "plb_pack_im2col": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                       c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64, c_i64, c_i32,
                                       c_ptr, c_ptr, c_i32, c_i32, c_ptr]),
    "plb_gemm_grouped": (ctypes.c_int, [c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "plb_cross_finalize": (ctypes.c_int, [c_ptr, c_i32, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i32, c_ptr,
                                          c_ptr, c_i64, c_i32, c_i32, c_ptr]),
    "plb_lap_solve_batched": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr]),
    "plb_lap_solve_batched_warm": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr,
                                                  c_f64, c_ptr, c_ptr]),
    "plb_get_blocks": (ctypes.c_int, [c_ptr, c_i64, c_ptr, c_i32, c_f32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                      c_ptr]),
    "plb_block_merge": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i32, c_ptr, c_ptr]),
    "plb_gather_axis": (ctypes.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "plb_compose_perm": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "plb_wm_progress": (ctypes.c_int, [c_ptr, c_i64, c_ptr, c_i32, c_ptr, c_ptr, c_ptr]),
    "plb_chol_solve": (ctypes.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_f64, c_ptr, c_ptr]),
    "plb_conv_debug_set_trace": (ctypes.c_int, [c_ptr]),
    "plb_conv_packed_floats": (c_i64, [c_i64, c_i64, c_i32, c_i32]),
    "plb_conv_pack_weights": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "plb_conv2d_forward": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i64, c_i64, c_i64, c_i64,
                                          c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "plb_conv2d_affine_forward": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i64, c_i64,
                                                 c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def lib():
    """Loads the shared library on first use; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"pleas_merging_b200: CUDA library not found at {LIB_PATH}; build it with "
                "`python -m pleas_merging_b200.build` (or __graft_entry__.build()). There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


import collections

LAUNCH_COUNTS = collections.Counter()  # C-ABI calls by entry point (each launches >= 1 kernel)


def check(rc, what):
    LAUNCH_COUNTS[what] += 1
    if rc != 0:
        msg = lib().plb_last_error_string().decode(errors="replace")
        raise RuntimeError(f"pleas_merging_b200: {what} failed with status {rc}: {msg}")


def call(name, device, *args):
    """Invokes the entry point ``name`` for tensors living on ``device``: the launch happens with
    that device current and on ITS current stream (the C ABI takes raw pointers, so nothing else
    ties a launch to the tensors' device; a model on cuda:1 while cuda:0 is current must not run on
    device 0's stream).  The stream handle is appended as the last argument."""
    import torch

    fn = getattr(lib(), name)
    if device.index is None or device.index == torch.cuda.current_device():
        rc = fn(*args, torch.cuda.current_stream().cuda_stream)
    else:
        with torch.cuda.device(device):
            rc = fn(*args, torch.cuda.current_stream(device).cuda_stream)
    check(rc, name)


def stream_ptr(device=None):
    import torch

    return torch.cuda.current_stream(device).cuda_stream


_SM_COUNT = {}


def sm_count(device=None):
    """Streaming multiprocessors of ``device`` (148 on a B200), cached per device ordinal."""
    import torch

    idx = torch.cuda.current_device() if device is None or getattr(device, "index", None) is None else device.index
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def ptr(t):
    """data_ptr of a tensor or None -> NULL."""
    return None if t is None else t.data_ptr()
