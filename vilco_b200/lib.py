"""ctypes binding of the C-ABI library ``libvilco_b200.so`` (declared in ``include/vilco_b200.h``).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used only for device memory and streams; every compute call goes through the C ABI with
raw device pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvilco_b200.so")

ACT_NONE, ACT_RELU, ACT_GELU, ACT_EXP2 = 0, 1, 2, 3
F32, BF16, F16 = 0, 1, 2


class VilcoError(RuntimeError):
    pass


class VilcoGemm(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_ld", C.c_int64), ("a_s1", C.c_int64), ("a_s2", C.c_int64), ("a_lo", C.c_int64), ("a_rows", C.c_int32),
        ("B", C.c_void_p), ("b_ld", C.c_int64), ("b_s1", C.c_int64), ("b_s2", C.c_int64), ("b_lo", C.c_int64),
        ("b_major", C.c_int32), ("b_batched", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("taps", C.c_int32), ("Z1", C.c_int32), ("Z2", C.c_int32),
        ("D", C.c_void_p), ("d_dtype", C.c_int32), ("d_ld", C.c_int64), ("d_s1", C.c_int64), ("d_s2", C.c_int64), ("d_lo", C.c_int64),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("rowmul", C.c_void_p), ("rowmul_zs", C.c_int64),
        ("act", C.c_int32),
        ("colscale", C.c_void_p),
        ("resid", C.c_void_p), ("resid_masked", C.c_int32),
        ("impl", C.c_int32),
        ("band_lo", C.c_int32), ("band_hi", C.c_int32), ("a_major", C.c_int32),
        ("a_fmt", C.c_int32), ("b_fmt", C.c_int32),
        ("rowsub", C.c_void_p), ("rowsub_s1", C.c_int64), ("rowsub_s2", C.c_int64), ("colscale_zs", C.c_int64), ("emul", C.c_void_p),
    ]


_lib = None


def lib():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VilcoError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vilco_b200 has no CPU / PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vilco_last_error.restype = C.c_char_p
        _lib.vilco_launch_count.restype = C.c_uint64
        _lib.vilco_version.restype = C.c_int
        _lib.vilco_nms_workspace_bytes.restype = C.c_size_t
        if _lib.vilco_version() < 3:
            raise VilcoError(f"{LIB_PATH} is stale (ABI version {_lib.vilco_version()} < 3): rebuild it")
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise VilcoError(f"{what} failed (code {rc}): {lib().vilco_last_error().decode()}")


def launch_count():
    return int(lib().vilco_launch_count())


_dev_index = None


def stream_ptr():
    """raw handle of torch's current CUDA stream (the cheap private accessor: this is called once per kernel launch, ~1500
    times per training step).  One process drives one GPU, so the device index is looked up once."""
    global _dev_index
    if _dev_index is None:
        _dev_index = torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(_dev_index))


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def default_gemm_impl():
    return 1 if os.environ.get("VILCO_GEMM", "tc") == "simt" else 0


_FMT = {torch.bfloat16: BF16, torch.float16: F16}
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}
_gemm_cache = {}     # descriptor without the pointers -> prepared VilcoGemm (a training step issues ~570 GEMMs from ~60 shapes)
_gemm_fn = None


def gemm(A, B, D, *, M, N, K, a_rows, a_ld, b_ld, d_ld, a_s=(0, 0), b_s=(0, 0), d_s=(0, 0), Z=(1, 1), taps=1,
         b_major=0, b_batched=False, alpha=1.0, bias=None, rowmul=None, rowmul_zs=0, act=ACT_NONE,
         colscale=None, resid=None, resid_masked=False, impl=None, a_lo=0, b_lo=0, d_lo=0, band=(0, 0), a_major=0,
         rowsub=None, rowsub_s=(0, 0), colscale_zs=0, emul=None):
    """Raw descriptor-level call of ``vilco_gemm`` (see include/vilco_b200.h for the contract)."""
    global _gemm_fn
    key = (M, N, K, a_rows, a_ld, b_ld, d_ld, a_s, b_s, d_s, Z, taps, b_major, b_batched, alpha, rowmul_zs, act, resid_masked,
           impl, a_lo, b_lo, d_lo, band, a_major, A.dtype, B.dtype, D.dtype, rowsub_s, colscale_zs)
    g = _gemm_cache.get(key)
    if g is None:
        assert A.dtype in _FMT and B.dtype in _FMT and A.is_cuda and B.is_cuda, "operands must be fp16 / bf16 planes"
        assert D.dtype in _DT
        g = VilcoGemm()
        g.a_ld, g.a_s1, g.a_s2, g.a_rows = a_ld, a_s[0], a_s[1], a_rows
        g.a_lo, g.b_lo, g.d_lo = a_lo, b_lo, d_lo
        g.b_ld, g.b_s1, g.b_s2 = b_ld, b_s[0], b_s[1]
        g.b_major, g.b_batched = b_major, int(b_batched)
        g.M, g.N, g.K, g.taps, g.Z1, g.Z2 = M, N, K, taps, Z[0], Z[1]
        g.d_dtype, g.d_ld, g.d_s1, g.d_s2 = _DT[D.dtype], d_ld, d_s[0], d_s[1]
        g.a_fmt, g.b_fmt = _FMT[A.dtype], _FMT[B.dtype]
        g.alpha = alpha
        g.rowmul_zs = rowmul_zs
        g.act = act
        g.resid_masked = int(resid_masked)
        g.impl = default_gemm_impl() if impl is None else impl
        g.band_lo, g.band_hi = band
        g.a_major = a_major
        g.rowsub_s1, g.rowsub_s2, g.colscale_zs = rowsub_s[0], rowsub_s[1], colscale_zs
        _gemm_cache[key] = g
        if _gemm_fn is None:
            _gemm_fn = lib().vilco_gemm
    g.A, g.B, g.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.rowmul = rowmul.data_ptr() if rowmul is not None else None
    g.colscale = colscale.data_ptr() if colscale is not None else None
    g.resid = resid.data_ptr() if resid is not None else None
    g.rowsub = rowsub.data_ptr() if rowsub is not None else None
    g.emul = emul.data_ptr() if emul is not None else None
    rc = _gemm_fn(C.byref(g), stream_ptr())
    if rc != 0:
        check(rc, "vilco_gemm")
    return D
