"""ctypes loader for the plain-C oracle (oracle/nm_oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnm_oracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "nm_oracle_c.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = ctypes.CDLL(_SO)
        f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        L.nmo_episodic_normalization_f32.argtypes = [f32p, ctypes.c_size_t, ctypes.c_double, ctypes.c_double,
                                                     ctypes.c_double, f64p]
        L.nmo_voxelize_f64.argtypes = [f64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, f32p]
        L.nmo_normalize_voxelize_clip.argtypes = [f32p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_double,
                                                  ctypes.c_double, ctypes.c_double, f64p, f32p]
        L.nmo_voxel_chamfer_frame.argtypes = [f32p, f32p, ctypes.c_int, f64p]
        L.nmo_semantic_nearest_frame.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                 ctypes.POINTER(ctypes.c_int)]
        L.nmo_semantic_nearest_frame.restype = None
        for fn in (L.nmo_episodic_normalization_f32, L.nmo_voxelize_f64, L.nmo_normalize_voxelize_clip,
                   L.nmo_voxel_chamfer_frame):
            fn.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def episodic_normalization(seq: np.ndarray, scale=1.0, x_trans=0.0, z_trans=0.0) -> np.ndarray:
    seq = np.ascontiguousarray(seq, dtype=np.float32)
    out = np.empty(seq.shape, dtype=np.float64)
    rc = lib().nmo_episodic_normalization_f32(_p(seq, ctypes.c_float), seq.size // 3, scale, x_trans, z_trans,
                                              _p(out, ctypes.c_double))
    if rc:
        raise ValueError("empty clip")
    return out


def voxelize(points: np.ndarray, G: int) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.float64)
    grid = np.empty((1, G, G, G), dtype=np.float32)
    rc = lib().nmo_voxelize_f64(_p(pts, ctypes.c_double), pts.shape[0], pts.shape[1] if pts.ndim == 2 else 3, G,
                                _p(grid, ctypes.c_float))
    if rc:
        raise IndexError("point outside [-1, 1)")
    return grid


def normalize_voxelize_clip(raw: np.ndarray, G: int, scale=1.0, x_trans=0.0, z_trans=0.0) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    T, N, _ = raw.shape
    scratch = np.empty((T, N, 3), dtype=np.float64)
    grids = np.empty((T, 1, G, G, G), dtype=np.float32)
    rc = lib().nmo_normalize_voxelize_clip(_p(raw, ctypes.c_float), T, N, G, scale, x_trans, z_trans,
                                           _p(scratch, ctypes.c_double), _p(grids, ctypes.c_float))
    if rc:
        raise IndexError("point outside [-1, 1)" if rc == 2 else "empty clip")
    return grids


def voxel_chamfer_frame(gt: np.ndarray, recon: np.ndarray) -> float:
    gt = np.ascontiguousarray(gt, dtype=np.float32)
    recon = np.ascontiguousarray(recon, dtype=np.float32)
    out = ctypes.c_double()
    if lib().nmo_voxel_chamfer_frame(_p(gt, ctypes.c_float), _p(recon, ctypes.c_float), gt.shape[-1], ctypes.byref(out)):
        raise IndexError("empty gt or recon")
    return out.value


def semantic_nearest_frame(kypt: np.ndarray, gt: np.ndarray, threshold: float = 0.2) -> np.ndarray:
    kypt = np.ascontiguousarray(kypt, dtype=np.float32)
    gt = np.ascontiguousarray(gt, dtype=np.float32)
    idx = np.empty(gt.shape[0], dtype=np.int32)
    lib().nmo_semantic_nearest_frame(_p(kypt, ctypes.c_float), _p(gt, ctypes.c_float), kypt.shape[0], gt.shape[0], threshold,
                                     _p(idx, ctypes.c_int))
    return idx
