"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/softnms.c (built by oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_nms.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_softnms_1d.restype = C.c_int64
        _lib.oracle_softnms_1d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int,
                                           C.c_void_p, C.c_void_p]
    return _lib


def softnms_1d(segs, scores, iou_threshold, sigma, min_score, method):
    """Same signature as oracle.mq_oracle.softnms_1d -> (dets (n,3) f32, inds (n,) i64)."""
    segs = np.ascontiguousarray(segs, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = segs.shape[0]
    dets = np.zeros((n, 3), np.float32)
    inds = np.zeros((n,), np.int64)
    k = _load().oracle_softnms_1d(segs.ctypes.data, scores.ctypes.data, n, iou_threshold, sigma, min_score, method,
                                  dets.ctypes.data, inds.ctypes.data)
    return dets[:k], inds[:k]
