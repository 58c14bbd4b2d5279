"""ctypes binding of libagb200.so (include/agb200.h).

This is the Python twin of the `ccall` stubs shown in INTEGRATION.md: one thin wrapper per C entry
point, status codes mapped to the exception the reference would raise.  There is NO fallback: if the
shared library is missing the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGB200_LIB") or os.path.join(_HERE, "libagb200.so")   # AGB200_LIB: an instrumented build (experiments)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libagb200.so is not built (%s). Run `python __graft_entry__.py` (nvcc, sm_100a). "
        "There is no CPU fallback for the device path." % LIB_PATH)

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

F32, F64 = 0, 1
OK, EINVAL, ENOMEM, ECUDA, ENCCL, EUNSUPPORTED, ESTATE = range(7)

i32, i64, u64, dbl, vp = C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_void_p
pi64 = C.POINTER(C.c_int64)

class GemmDesc(C.Structure):
    """ag_gemm_desc (include/agb200.h)."""
    _fields_ = [("transA", C.c_int32), ("transB", C.c_int32), ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64), ("bias", C.c_void_p),
                ("act", C.c_int32), ("C", C.c_void_p), ("C2", C.c_void_p), ("ldc", C.c_int64)]


_SIGS = {
    "ag_init": [i32],
    "ag_shutdown": [],
    "ag_sync": [],
    "ag_last_error": [C.c_char_p, i64],
    "ag_device_info": [pi64],
    "ag_stream": [C.POINTER(u64)],
    "ag_launch_count": [pi64],
    "ag_alloc": [i64, C.POINTER(vp)],
    "ag_free": [vp],
    "ag_pool_trim": [],
    "ag_host_alloc": [i64, C.POINTER(vp)],
    "ag_host_free": [vp],
    "ag_copy_h2d": [vp, vp, i64],
    "ag_copy_h2d_prefetch": [vp, vp, i64],
    "ag_prefetch_slot": [i32],
    "ag_prefetch_wait": [],
    "ag_prefetch_release": [],
    "ag_copy_d2h": [vp, vp, i64],
    "ag_copy_d2h_async": [vp, vp, i64],
    "ag_copy_d2d": [vp, vp, i64],
    "ag_fill": [vp, i64, i32, dbl],
    "ag_event_create": [C.POINTER(u64)],
    "ag_event_record": [u64],
    "ag_event_elapsed_ms": [u64, u64, C.POINTER(C.c_float)],
    "ag_event_destroy": [u64],
    "ag_range_push": [C.c_char_p],
    "ag_range_pop": [],
    "ag_graph_begin": [],
    "ag_graph_end": [C.POINTER(u64)],
    "ag_graph_launch": [u64],
    "ag_graph_destroy": [u64],
    "ag_unary_fwd": [i32, i32, vp, vp, i64],
    "ag_special_param": [i32, i32, i32, vp, vp, i64],
    "ag_inv_det": [i32, vp, i64, vp, vp],
    "ag_argext": [i32, vp, i64, i32, vp],
    "ag_cast": [i32, i32, vp, vp, i64],
    "ag_unary_bwd": [i32, i32, vp, i32, dbl, vp, vp, vp, i64],
    "ag_binary_fwd": [i32, i32, vp, dbl, pi64, vp, dbl, pi64, vp, i32, pi64],
    "ag_binary_bwd": [i32, i32, i32, vp, pi64, vp, dbl, pi64, vp, dbl, pi64, vp, pi64, vp, i32, pi64],
    "ag_ewise_vm": [i32, C.POINTER(i32), i32, C.POINTER(dbl), i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(dbl), vp, i64],
    "ag_ewise_fused": [i32, C.POINTER(i32), i32, C.POINTER(dbl), i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(dbl), i32, C.POINTER(vp), i64, i64],
    "ag_ewise_fused_reduce": [i32, C.POINTER(i32), i32, C.POINTER(dbl), i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(dbl), i32, C.POINTER(vp), i64, i64, i64, i64, i64, vp],
    "ag_ewise_colprog": [i32, C.POINTER(i32), i32, C.POINTER(dbl), i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(dbl), i32, C.POINTER(vp), i64, i64, vp],
    "ag_ewise_fused_set_static": [i32],
    "ag_ewise_fused_query": [C.POINTER(i32), i32, C.POINTER(i32)],
    "ag_ewise_fused_stats": [pi64],
    "ag_sum_all": [i32, i32, vp, i64, vp],
    "ag_sum_dim": [i32, i32, vp, i64, i64, i64, vp],
    "ag_sum_dim_batch": [i32, i32, i32, C.POINTER(vp), i64, i64, i64, C.POINTER(vp)],
    "ag_reduce_dim": [i32, i32, vp, i64, i64, i64, vp],
    "ag_permute": [i32, vp, i32, pi64, C.POINTER(i32), vp],
    "ag_add": [i32, vp, vp, vp, i64],
    "ag_axpy": [i32, dbl, vp, vp, i64],
    "ag_scale": [i32, dbl, vp, i64],
    "ag_gemm": [i32, i32, i32, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64],
    "ag_gemm_epilogue": [i32, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i32, vp, vp, i64],
    "ag_gemm_epilogue_query": [i32, i32, i32, i64, i64, i64, vp, i64, vp, i64, C.POINTER(i32)],
    "ag_gemm_group": [i32, i32, C.POINTER(GemmDesc)],
    "ag_gemm_set_engine": [i32],
    "ag_gemm_last_engine": [C.c_char_p, i64],
    "ag_ubyte_to_float": [i32, vp, vp, i64, dbl],
    "ag_onehot_labels": [i32, vp, i64, i32, i32, vp],
    "ag_gather": [i32, vp, vp, vp, i64],
    "ag_gather_cols": [i32, vp, i64, vp, vp, i64],
    "ag_scatter_add": [i32, vp, i64, vp, vp, i64, i64],
    "ag_slab_copy": [i32, vp, i64, i64, vp, i64, i64, i64, i64, i64],
    "ag_cat": [i32, i32, C.POINTER(vp), pi64, vp, i64, i64],
    "ag_transpose": [i32, vp, i64, i64, vp],
    "ag_comm_unique_id": [C.POINTER(C.c_uint8)],
    "ag_comm_init": [i32, i32, C.POINTER(C.c_uint8)],
    "ag_comm_destroy": [],
    "ag_allreduce_sum": [i32, vp, i64],
    "ag_p2p_export": [i64, i32, C.POINTER(C.c_uint8)],
    "ag_p2p_connect": [i32, i32, C.POINTER(C.c_uint8)],
    "ag_p2p_allreduce_sum": [i32, vp, i64],
    "ag_p2p_allreduce_multi": [i32, i32, C.POINTER(vp), C.POINTER(vp), pi64, dbl],
    "ag_multi_axpy": [i32, i32, dbl, C.POINTER(vp), C.POINTER(vp), pi64],
    "ag_multi_sumabs2": [i32, i32, C.POINTER(vp), pi64, vp],
    "ag_multi_axpy_scaled": [i32, i32, dbl, vp, C.POINTER(vp), C.POINTER(vp), pi64],
    "ag_multi_copy": [i32, i32, C.POINTER(vp), C.POINTER(vp), pi64],
    "ag_p2p_destroy": [],
    "ag_tape_new": [C.POINTER(u64)],
    "ag_tape_free": [u64],
    "ag_tape_leaf": [u64, i32, vp, i32, pi64, i32, C.POINTER(i32)],
    "ag_tape_record": [u64, i32, i32, i32, C.POINTER(i32), dbl, C.POINTER(i32)],
    "ag_tape_backward": [u64, i32],
    "ag_tape_grad": [u64, i32, C.POINTER(vp)],
    "ag_tape_value": [u64, i32, C.POINTER(vp), C.POINTER(i32), pi64],
}

EXPORTS = sorted(_SIGS)

for _name, _args in _SIGS.items():
    _f = getattr(lib, _name)   # AttributeError here == header and library disagree
    _f.argtypes = _args
    _f.restype = i32


class DeviceError(RuntimeError):
    pass


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (src/unbroadcast.jl:20)."""


class UnsupportedOnDevice(TypeError):
    """No device method and no CPU fallback (Julia would raise MethodError)."""


def last_error():
    buf = C.create_string_buffer(1024)
    lib.ag_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(status):
    if status == OK:
        return
    msg = last_error()
    if status == EINVAL:
        if "BoundsError" in msg:
            raise IndexError(msg)          # the sticky device-side index check (Julia: BoundsError)
        raise DimensionMismatch(msg)
    if status == ENOMEM:
        raise MemoryError(msg)
    if status == EUNSUPPORTED:
        raise UnsupportedOnDevice(msg)
    raise DeviceError("libagb200 status %d: %s" % (status, msg))
