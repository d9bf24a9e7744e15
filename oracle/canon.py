"""ctypes front-end for oracle/canon.c (canonical kNN / FPS).  TEST INFRASTRUCTURE ONLY.

``build()`` compiles the C file with gcc into oracle/libvcr_canon.so (git-ignored, travels
to the GPU box with the snapshot).  ``-ffp-contract=off`` keeps the separately rounded
multiplies and adds that the reference's FPS uses (hard part 5).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "canon.c")
_SO = os.path.join(_HERE, "libvcr_canon.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", _SRC,
                               "-o", _SO, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.canon_knn.restype = ctypes.c_int
        _lib.canon_fps.restype = ctypes.c_int
    return _lib


def knn(x: np.ndarray, k: int) -> np.ndarray:
    """x [B,D,N] float32 -> int64 [B,N,k]; canonical order (see canon.c)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, D, N = x.shape
    out = np.empty((B, N, k), dtype=np.int32)
    rc = _load().canon_knn(x.ctypes.data_as(ctypes.c_void_p), B, D, N, k,
                           out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"canon_knn failed: {rc}")
    return out.astype(np.int64)


def fps(xyz: np.ndarray, npoint: int) -> np.ndarray:
    """xyz [B,3,N] float32 -> int64 [B,npoint]."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, C, N = xyz.shape
    assert C == 3
    out = np.empty((B, npoint), dtype=np.int32)
    rc = _load().canon_fps(xyz.ctypes.data_as(ctypes.c_void_p), B, N, npoint,
                           out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"canon_fps failed: {rc}")
    return out.astype(np.int64)
