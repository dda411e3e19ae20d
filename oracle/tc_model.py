"""TEST INFRASTRUCTURE -- CPU model of the tensor-core arithmetic of the rollout kernels (not the product, not a fallback).

The tensor-core kernels (cadm_b200/csrc/rollout_tcs.cu, rollout_tc.cu) do not multiply fp32 numbers: every operand is
pre-scaled by a power of two (activations x8, weights x64; csrc/ptx.cuh kXScale / kWScale), split into
hi = fp16(x), lo = fp16(x - hi) (ptx.cuh split2) and a product is three fp16 MMAs accumulated in fp32
(X_hi W_hi + X_hi W_lo + X_lo W_hi; "tc3x"), the scale 512 removed exactly in the epilogue.  "tc1x" keeps the first MMA only.
This module restates that arithmetic with NumPy so the precision argument of DESIGN.md section 4.2 -- dropped terms are
<~2^-21 relative, so a whole 5-iteration decision stays two orders of magnitude inside the 1e-4 parity bar, while a single
fp16 pass does not -- can be checked on the CPU against the float64 oracle (tests/test_tc_precision_model.py).  What it
does not model: the summation order inside the tensor core and the MUFU ex2 / rcp approximations of the epilogue (each
<= 2 ulp); the GPU tests measure those on the device.
"""
import numpy as np

X_SCALE, W_SCALE = np.float32(8.0), np.float32(64.0)


def split_f16(x):
    """fp32 -> (hi, lo) fp16 pair, returned as the exactly representable float32 values (ptx.cuh split2, saturating)."""
    lim = np.float32(65504.0)
    x = np.asarray(x, np.float32)
    hi = np.clip(x, -lim, lim).astype(np.float16).astype(np.float32)
    lo = np.clip(x - hi, -lim, lim).astype(np.float16).astype(np.float32)
    return hi, lo


def matmul_split(x, W, terms=3):
    """x [E, R, K] @ W [E, K, N] the way the kernels compute it: fp16 operand halves, fp32 accumulation.  A product of two
    fp16 values is exact in fp32 (22 significant bits), so float32 matmuls of the halves model the MMAs up to summation
    order."""
    xh, xl = split_f16(np.asarray(x, np.float32) * X_SCALE)
    wh, wl = split_f16(np.asarray(W, np.float32) * W_SCALE)
    acc = np.matmul(xh, wh)
    if terms == 3:
        acc = acc + np.matmul(xh, wl) + np.matmul(xl, wh)
    return acc * np.float32(1.0 / (X_SCALE * W_SCALE))


def dense_split(terms=3):
    """A drop-in for oracle.cadm_oracle.dense with the contraction replaced by the split product."""
    def dense(x, W, b, act=None):
        out = matmul_split(x, W, terms) + np.asarray(b, np.float32)
        return act(out) if act is not None else out
    return dense
