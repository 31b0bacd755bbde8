"""ORACLE (test infrastructure): ctypes front-end of `oracle/visibility_ref.c`."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "visibility_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_seg_counts.restype = ctypes.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def visibility_view(points, depth, inv_pose, K, threshold=0.05, want_pixels=False):
    """One view: returns mask (N,) int64 [, pixels (N,2) int64, zdepth (N,) fp64]."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    dep = np.ascontiguousarray(depth, dtype=np.float32)
    inv = np.ascontiguousarray(inv_pose, dtype=np.float64)  # fp32 inverses widen exactly, fp64 ones stay
    Kc = np.ascontiguousarray(K, dtype=np.float64)
    n = pts.shape[0]
    mask = np.empty(n, dtype=np.int64)
    pix = np.empty((n, 2), dtype=np.int64) if want_pixels else None
    zd = np.empty(n, dtype=np.float64) if want_pixels else None
    lib().oracle_visibility(_p(pts), ctypes.c_int64(n), _p(dep), ctypes.c_int64(dep.shape[0]),
                            ctypes.c_int64(dep.shape[1]), _p(inv), _p(Kc), ctypes.c_double(threshold),
                            _p(mask), _p(pix), _p(zd))
    return (mask, pix, zd) if want_pixels else mask


def visibility_mask(points, depths, poses, K, threshold=0.05, inv_poses=None):
    """(V,N) int64. `poses` are camera->world matrices (fp32 in production, fp64 allowed); the inverse is taken with
    np.linalg.inv in the pose dtype exactly like utils/transforms.py:54 unless `inv_poses`
    (already inverted, e.g. stored in a golden file) is given."""
    rows = []
    for v, depth in enumerate(depths):
        inv = inv_poses[v] if inv_poses is not None else np.linalg.inv(poses[v])
        rows.append(visibility_view(points, depth, inv, K, threshold))
    return np.stack(rows)


def transform(points, matrix):
    """(matrix . [p;1])[:3] per point in fp64 (utils/transforms.py:43-61); the matrix is widened to fp64 like np.dot does."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    M = np.ascontiguousarray(matrix, dtype=np.float64).reshape(16)
    out = np.empty_like(pts)
    lib().oracle_transform(_p(pts), ctypes.c_int64(pts.shape[0]), _p(M), _p(out))
    return out


def seg_counts(seg, nbins):
    s = np.ascontiguousarray(seg, dtype=np.int64).reshape(-1)
    counts = np.empty(nbins, dtype=np.int64)
    outside = lib().oracle_seg_counts(_p(s), ctypes.c_int64(s.size), ctypes.c_int64(nbins), _p(counts))
    return counts, int(outside)


def quantize(xyz32, size):
    x = np.ascontiguousarray(xyz32, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.int32)
    lib().oracle_quantize(_p(x), ctypes.c_int64(x.size), ctypes.c_float(size), _p(out))
    return out
