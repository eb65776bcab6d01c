"""ctypes binding of libdynam3d_b200.so (the C ABI declared in include/dynam3d_b200.h).

There is NO fallback: if the shared library is missing or the device is not sm_100 the import of any
compute entry fails loudly (`D3DLibraryError`).  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynam3d_b200.so")

D3D_F16, D3D_BF16, D3D_OUT_F32 = 0, 1, 2
ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SILU, ACT_SWIGLU = 0, 1, 2, 3, 4


class D3DLibraryError(RuntimeError):
    pass


class D3DError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64),
        ("W", ctypes.c_void_p), ("ldw", ctypes.c_int64),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int64),
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("in_kind", ctypes.c_int), ("out_kind", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("act", ctypes.c_int),
        ("residual", ctypes.c_void_p), ("ldres", ctypes.c_int64),
    ]


_lib = None


def lib():
    """Load the library once; raise if it is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise D3DLibraryError(f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.d3d_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise D3DError(f"libdynam3d_b200 error {rc}: {lib().d3d_last_error().decode()}")


def require_device(dev=0):
    import torch
    if not torch.cuda.is_available():
        raise D3DLibraryError("no CUDA device: dynam3d_b200 has no CPU fallback")
    check(lib().d3d_check_device(int(dev)))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def kind_of(dtype):
    import torch
    if dtype == torch.float16:
        return D3D_F16
    if dtype == torch.bfloat16:
        return D3D_BF16
    if dtype == torch.float32:
        return D3D_OUT_F32
    raise TypeError(dtype)
