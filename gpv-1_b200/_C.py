"""ctypes binding of the C-ABI in include/gpvb200.h.

There is deliberately no fallback: if the shared library is missing, or a call fails (wrong architecture, bad
shape), a RuntimeError is raised with the library's own message.
"""
import ctypes
import os
from ctypes import c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libgpvb200.so")
_lib = None


class GemmDesc(ctypes.Structure):
    """Mirror of `gpvb200_gemm_desc` (field order and types must match include/gpvb200.h exactly)."""
    _fields_ = [
        ("mode", c_int32), ("M", c_int32), ("N", c_int32), ("K", c_int32), ("batch", c_int32),
        ("a_mn", c_int32), ("b_mn", c_int32), ("act", c_int32), ("aux_mode", c_int32),
        ("d_fp32", c_int32), ("d_atomic", c_int32), ("splits", c_int32),
        ("n_img", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("Ho", c_int32), ("Wo", c_int32), ("stride", c_int32),
        ("ntaps", c_int32),
        ("tap_dh", c_int32 * 9), ("tap_dw", c_int32 * 9), ("tap_w", c_int32 * 9),
        ("OH", c_int32), ("OW", c_int32), ("out_stride", c_int32), ("out_off_h", c_int32), ("out_off_w", c_int32),
        ("alpha", c_float), ("res_fp32", c_int32),
        ("A", c_void_p), ("B", c_void_p), ("D", c_void_p), ("D2", c_void_p),
        ("bias", c_void_p), ("rowscale", c_void_p), ("residual", c_void_p), ("aux", c_void_p),
        ("lda", c_int64), ("ldb", c_int64), ("ldd", c_int64), ("ldr", c_int64), ("ldaux", c_int64),
        ("a_batch_stride", c_int64), ("b_batch_stride", c_int64), ("d_batch_stride", c_int64),
        ("drop_seed", c_void_p), ("drop_mode", c_int32), ("drop_site", ctypes.c_uint32), ("drop_p", c_float), ("max_ctas", c_int32),
    ]


def lib():
    """Load lib/libgpvb200.so (built by gpv-1_b200/build.py). Raises if it is absent -- no CPU path exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} not found: build it with `python gpv-1_b200/build.py` (the product path has no fallback)")
        L = ctypes.CDLL(SO_PATH)
        L.gpvb200_gemm_desc_size.restype = c_size_t
        if L.gpvb200_gemm_desc_size() != ctypes.sizeof(GemmDesc):
            raise RuntimeError("gpvb200_gemm_desc layout mismatch between _C.py and libgpvb200.so")
        L.gpvb200_last_error.argtypes = [ctypes.c_char_p, c_size_t]
        L.gpvb200_pack_item_size.restype = c_size_t
        L.gpvb200_gemm_pair_launches.restype = c_int64
        L.gpvb200_optim_item_size.restype = c_size_t
        _lib = _Counting(L)
    return _lib


class _Counting:
    """Thin proxy over the CDLL: counts kernel-launching C-ABI calls (bench.py reports them as `gpu_launches`) and, when
    `trace` is a list, brackets every call with CUDA events on the launching stream (per-kernel time breakdown)."""

    def __init__(self, L):
        object.__setattr__(self, "_L", L)
        object.__setattr__(self, "launches", 0)
        object.__setattr__(self, "trace", None)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._L, name)
            if name in ("gpvb200_version", "gpvb200_last_error", "gpvb200_gemm_desc_size", "gpvb200_gemm_pair_launches", "gpvb200_pack_item_size",
                        "gpvb200_pack_chunk",
                        "gpvb200_optim_item_size", "gpvb200_optim_chunk"):
                fn = raw
            else:
                def fn(*a, _raw=raw, _name=name):
                    object.__setattr__(self, "launches", self.launches + 1)
                    tr = self.trace
                    if tr is None:
                        return _raw(*a)
                    import torch
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = _raw(*a)
                    e1.record()
                    tr.append((_name, a, e0, e1))
                    return rc
            self._cache[name] = fn
        return fn


def counters():
    return lib()


def last_error() -> str:
    buf = ctypes.create_string_buffer(1024)
    lib().gpvb200_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"gpvb200 {what} failed (code {rc}): {last_error()}")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
