"""Builds oracle/_build/libcpu_a16w4.so (gcc + OpenMP) and binds it with ctypes.  Test / baseline infrastructure only:
the product never imports this (tests/test_cabi_cpu.py::test_product_does_not_import_oracle)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_a16w4.c")
LIB = os.path.join(HERE, "_build", "libcpu_a16w4.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -march=x86-64-v3 (AVX2 + FMA), not -march=native: the .so is built in the dev container and travels to the GPU box
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.a16w4_forward_f32.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4
        _lib.a16w4_forward_f32.restype = None
        _lib.a16w4_threads.restype = ctypes.c_int
    return _lib


def forward(Wp_i32, scales_f32, zeros_f32, x_f32):
    """Wp (Kp, N) int32, scales / zeros (G, N) float32, x (M, K) float32 numpy arrays -> y (M, N) float32."""
    lib = load()
    Wp = np.ascontiguousarray(Wp_i32, np.int32)
    s = np.ascontiguousarray(scales_f32, np.float32)
    z = np.ascontiguousarray(zeros_f32, np.float32)
    x = np.ascontiguousarray(x_f32, np.float32)
    Kp, N = Wp.shape
    M = x.shape[0]
    assert x.shape[1] == Kp * 8 and N % 64 == 0 and Kp % s.shape[0] == 0
    y = np.empty((M, N), np.float32)
    lib.a16w4_forward_f32(Wp.ctypes.data, s.ctypes.data, z.ctypes.data, x.ctypes.data, y.ctypes.data, M, Kp, N, s.shape[0])
    return y


def threads():
    return int(load().a16w4_threads())
