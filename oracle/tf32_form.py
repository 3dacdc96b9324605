"""NumPy emulation of the GEMM-form ADD-S variant -- TEST INFRASTRUCTURE, NOT PRODUCT.

Mirrors 6d-pose-estimation_b200/csrc/p6d_tf32.cu step by step (same re-centring, same TF32
splits, same K = 16 operand rows), so that the tensor-core question of BASELINE.json's north_star --
"tensor cores are excluded unless a 3xTF32 variant passes the stated tolerance" -- has an answer
that can be reproduced without a GPU.  The reference op is models/add_loss.py:185-190.

The one thing NumPy cannot know is how the tensor core rounds inside its accumulation; two
bracketing models are provided:
    accumulate="exact"     the 16 products summed exactly (float64), rounded to float32 once
                           (optimistic: no hardware can do better with an FP32 accumulator)
    accumulate="f32_seq"   every product added in float32 round-to-nearest, k = 0..15 in order
    accumulate="f32_rz"    the same with every addition truncated toward zero (tensor-core adders
                           are commonly described as truncating; truncation is a BIAS, which the
                           mean over the points does not average out)
The GPU test (tests/test_tf32_variant.py, -m gpu) measures the real kernel.
"""
from __future__ import annotations

import numpy as np


def tf32_rn(x):
    """Round float32 to the 10-bit TF32 mantissa, nearest / ties away (cvt.rna.tf32.f32)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def _fma_sq3(v):
    """|v|^2 with the kernels' rounding: fma(z,z, fma(y,y, x*x))."""
    d = v.astype(np.float64)
    s = np.float32(d[..., 0] * d[..., 0]).astype(np.float64)
    s = np.float32(d[..., 1] * d[..., 1] + s).astype(np.float64)
    return np.float32(d[..., 2] * d[..., 2] + s)


def operands(pred_cloud, gt_cloud, centre, split_terms=3):
    """A [n,16], B [n,16] operand rows and |p|^2 [n] exactly as the kernel builds them."""
    f = np.float32
    p = (pred_cloud.astype(f) - centre.astype(f)).astype(f)
    g = (gt_cloud.astype(f) - centre.astype(f)).astype(f)
    n = p.shape[0]
    A, B = np.zeros((n, 16), f), np.zeros((n, 16), f)
    ph = tf32_rn(p)
    gh = tf32_rn(g)
    A[:, 0:3] = ph
    B[:, 0:3] = -2.0 * gh
    if split_terms >= 3:
        A[:, 3:6] = ph
        A[:, 6:9] = tf32_rn((p - ph).astype(f))
        B[:, 3:6] = -2.0 * tf32_rn((g - gh).astype(f))
        B[:, 6:9] = -2.0 * gh
    A[:, 9:12] = 1.0
    g2 = _fma_sq3(g)
    h = tf32_rn(g2)
    m = tf32_rn((g2 - h).astype(f))
    B[:, 9], B[:, 10], B[:, 11] = h, m, tf32_rn(((g2 - h).astype(f) - m).astype(f))
    return A, B, _fma_sq3(p)


def adds_gemm_form(pred_cloud, gt_cloud, gt_translation, split_terms=3, accumulate="f32_seq", block=256):
    """mean_i min_j |pred_i - gt_j| through  d^2 = |p|^2 + (|g|^2 - 2 p.g)  with TF32 operands."""
    centre = (np.rint(np.asarray(gt_translation, np.float32) * 256.0) / 256.0).astype(np.float32)
    A, B, pn = operands(pred_cloud, gt_cloud, centre, split_terms)
    n = A.shape[0]
    mins = np.full(n, np.inf, np.float32)
    for lo in range(0, n, block):
        a = A[lo:lo + block]
        if accumulate == "exact":
            S = (a.astype(np.float64) @ B.astype(np.float64).T).astype(np.float32)
        elif accumulate == "f32_rz":
            S = np.zeros((a.shape[0], n), np.float64)
            for k in range(16):
                t = S + a[:, k:k + 1].astype(np.float64) * B[None, :, k].astype(np.float64)   # exact in float64
                r = t.astype(np.float32)
                # round-to-nearest overshot in magnitude -> step back one ulp toward zero
                over = np.abs(r.astype(np.float64)) > np.abs(t)
                r = np.where(over, np.nextafter(r, np.float32(0)), r).astype(np.float32)
                S = r.astype(np.float64)
            S = S.astype(np.float32)
        else:
            S = np.zeros((a.shape[0], n), np.float32)
            for k in range(16):
                # TF32 x TF32 products are exact in float32; the running sum rounds every step
                S = (S + (a[:, k:k + 1] * B[None, :, k]).astype(np.float32)).astype(np.float32)
        mins[lo:lo + block] = S.min(axis=1)
    d2 = (pn + mins).astype(np.float32)
    d = np.sqrt(np.maximum(d2, 0).astype(np.float64)).astype(np.float32)
    return np.float32(d.astype(np.float64).mean())
