"""The spline math the kernels run (sbi_b200/csrc/rqs.cuh is __host__ __device__) compiled with g++
and checked on the CPU against the oracle: forward, inverse and the hand-derived backward."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle.nflows_port.transforms.splines.rational_quadratic import (
    unconstrained_rational_quadratic_spline as urqs,
)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "rqs.cuh"
static sbi::RqsConst mk(int K, float B, float isq) {
  sbi::RqsConst c{K, B, isq, 1e-3f, 1e-3f, 1e-3f, (float)log(exp(1.0 - 1e-3) - 1.0)};
  return c;
}
extern "C" {
void h_fwd(const float* p, int K, float B, float isq, float x, float* y, float* ld) {
  sbi::rqs_forward(p, 1, mk(K, B, isq), x, *y, *ld);
}
void h_inv(const float* p, int K, float B, float isq, float x, float* y, float* ld) {
  sbi::rqs_inverse(p, 1, mk(K, B, isq), x, *y, *ld);
}
float h_bwd(const float* p, int K, float B, float isq, float x, float gy, float gl, float* g) {
  return sbi::rqs_backward(p, 1, mk(K, B, isq), x, gy, gl, g, 1);
}
// the register-array variant of the tensor-core training kernel (K = 10 instantiated there)
float h_bwd_reg10(const float* p, float B, float isq, float x, float gy, float gl, float* g) {
  float pp[32] = {0}, gg[32];
  for (int i = 0; i < 29; ++i) pp[i] = p[i];
  const float gx = sbi::rqs_backward_reg<10>(pp, mk(10, B, isq), x, gy, gl, gg);
  for (int i = 0; i < 32; ++i) g[i] = gg[i];
  return gx;
}
}
'''


@pytest.fixture(scope="module")
def hostlib():
    td = tempfile.mkdtemp()
    src = os.path.join(td, "rqs_host.cpp")
    open(src, "w").write(SRC)
    so = os.path.join(td, "rqs_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "sbi_b200", "csrc"),
                           "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.h_bwd.restype = ctypes.c_float
    lib.h_bwd_reg10.restype = ctypes.c_float
    return lib


FP = ctypes.POINTER(ctypes.c_float)
cf = ctypes.c_float


@pytest.mark.parametrize("K", [10, 4, 16])
def test_rqs_forward_inverse_backward(hostlib, K):
    B, H = 3.0, 50
    isq = 1 / np.sqrt(H)
    rng = np.random.default_rng(K)
    worst = dict(y=0.0, ld=0.0, inv=0.0, gx=0.0, gp=0.0)
    edge = [-3.0, 3.0, 0.0, -3.0000002, 2.9999998, -3.5, 3.5, float("nan")]
    for t in range(600):
        p = (rng.standard_normal(3 * K - 1) * 3).astype(np.float32)
        x = np.float32(edge[t]) if t < len(edge) else np.float32(rng.uniform(-3.3, 3.3))
        y, ld = cf(), cf()
        hostlib.h_fwd(p.ctypes.data_as(FP), K, cf(B), cf(isq), cf(x), ctypes.byref(y), ctypes.byref(ld))
        if np.isnan(x):
            assert np.isnan(y.value) and ld.value == 0.0     # NaN is "outside" -> identity
            continue
        pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        xt = torch.tensor([float(x)], dtype=torch.float64, requires_grad=True)
        yo, ldo = urqs(xt, pt[None, :K] / np.sqrt(H), pt[None, K:2 * K] / np.sqrt(H), pt[None, 2 * K:],
                       tails="linear", tail_bound=B)
        worst["y"] = max(worst["y"], abs(yo.item() - y.value))
        worst["ld"] = max(worst["ld"], abs(ldo.item() - ld.value))
        gy, gl = rng.standard_normal(2)
        (yo * gy + ldo * gl).sum().backward()
        g = np.zeros(3 * K - 1, np.float32)
        gx = hostlib.h_bwd(p.ctypes.data_as(FP), K, cf(B), cf(isq), cf(x), cf(gy), cf(gl), g.ctypes.data_as(FP))
        pg = pt.grad.numpy() if pt.grad is not None else np.zeros(3 * K - 1)
        sc = max(1.0, np.abs(pg).max(), abs(xt.grad.item()))
        worst["gx"] = max(worst["gx"], abs(xt.grad.item() - gx) / sc)
        worst["gp"] = max(worst["gp"], np.abs(pg - g).max() / sc)
        xi, ldi = cf(), cf()
        hostlib.h_inv(p.ctypes.data_as(FP), K, cf(B), cf(isq), cf(y.value), ctypes.byref(xi), ctypes.byref(ldi))
        if abs(ld.value) < 6:   # well-conditioned segments only: fp32 inverse of a e^-6 slope is noise
            worst["inv"] = max(worst["inv"], abs(xi.value - x) + abs(ldi.value + ld.value))
    print(K, worst)
    assert worst["y"] < 5e-5 and worst["ld"] < 5e-4
    assert worst["gx"] < 2e-3 and worst["gp"] < 2e-3
    assert worst["inv"] < 5e-3


def test_rqs_backward_register_variant_equals_strided(hostlib):
    """rqs_backward_reg<10> (tensor-core training kernel) == rqs_backward bit for bit."""
    K, B, isq = 10, 3.0, 1 / np.sqrt(50)
    rng = np.random.default_rng(3)
    for t in range(400):
        p = (rng.standard_normal(3 * K - 1) * 3).astype(np.float32)
        x = np.float32(rng.uniform(-3.4, 3.4)) if t > 3 else np.float32([-3.0, 3.0, 0.0, 5.0][t])
        gy, gl = rng.standard_normal(2)
        g1, g2 = np.zeros(3 * K - 1, np.float32), np.zeros(32, np.float32)
        a = hostlib.h_bwd(p.ctypes.data_as(FP), K, cf(B), cf(isq), cf(x), cf(gy), cf(gl), g1.ctypes.data_as(FP))
        b = hostlib.h_bwd_reg10(p.ctypes.data_as(FP), cf(B), cf(isq), cf(x), cf(gy), cf(gl), g2.ctypes.data_as(FP))
        assert a == b and np.array_equal(g1, g2[:29]) and not g2[29:].any()
