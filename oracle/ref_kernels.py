"""ctypes loader for oracle/_ref/libref_kernels_{nofma,fma}.so: the reference's OWN OpenCL kernel source
(Watermark_GPU/kernels/*.hpp) compiled for the CPU through oracle/clshim/.  TEST INFRASTRUCTURE ONLY.

Used to pin the C restatement (oracle/wm_oracle.c) against the reference's kernels: the restatement must agree
bit for bit with what the kernel text computes.  Inputs/outputs here are logical (rows, cols) numpy arrays; the
ArrayFire column-major buffers the kernels see are built inside these wrappers.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_kernels_nofma.so"))


def lib(fma=False):
    key = "fma" if fma else "nofma"
    if key not in _libs:
        path = os.path.join(_HERE, "_ref", "libref_kernels_%s.so" % key)
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
        L = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        L.ref_nvf.argtypes = [fp, C.c_int, C.c_int, fp]
        L.ref_nvf_p.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.ref_scaled_neighbors.argtypes = [fp, C.c_int, C.c_int, fp, fp]
        L.ref_me.argtypes = [fp, C.c_int, C.c_int, fp, fp]
        _libs[key] = L
    return _libs[key]


def _colmajor(img):
    return np.ascontiguousarray(np.asarray(img, np.float32).T)  # (cols, rows) C-order == (rows, cols) column-major


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def nvf(img, fma=False, p=3):
    """the reference's nvf kernel compiled with -Dp=<p> (kernels/nvf.hpp:14-17; p in 3, 5, 7, 9)"""
    rows, cols = img.shape
    buf = _colmajor(img)
    out = np.zeros_like(buf)
    if p == 3:
        lib(fma).ref_nvf(_p(buf), rows, cols, _p(out))
    elif lib(fma).ref_nvf_p(_p(buf), rows, cols, p, _p(out)) != 0:
        raise ValueError("p must be 3, 5, 7 or 9")
    return np.ascontiguousarray(out.T)


def scaled_neighbors(img, coeffs, fma=False):
    rows, cols = img.shape
    buf = _colmajor(img)
    c = np.ascontiguousarray(coeffs, np.float32)
    out = np.zeros_like(buf)
    lib(fma).ref_scaled_neighbors(_p(buf), rows, cols, _p(c), _p(out))
    return np.ascontiguousarray(out.T)


def me_partials(img, fma=False):
    """-> (RxPartial [ngroups, 64], rxPartial [ngroups, 8]) exactly as the `me` kernel writes them, grouped the
    way transformCorrelationArrays (Watermark.cpp:140-151) reshapes them before af::sum."""
    rows, cols = img.shape
    pc = (cols + 63) & ~63
    buf = _colmajor(img)
    Rxp = np.zeros(rows * pc, np.float32)
    rxp = np.zeros(rows * pc // 8, np.float32)
    lib(fma).ref_me(_p(buf), rows, cols, _p(Rxp), _p(rxp))
    return Rxp.reshape(-1, 64), rxp.reshape(-1, 8)


def rx(img, fma=False):
    """Rx (8x8), rx (8): group partials summed in f64 (the oracle's canonical stand-in for af::sum)."""
    Rxp, rxp = me_partials(img, fma)
    return Rxp.astype(np.float64).sum(0).reshape(8, 8), rxp.astype(np.float64).sum(0)
