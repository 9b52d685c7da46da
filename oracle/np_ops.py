"""ORACLE -- test infrastructure, not product code.

Independent numpy (float64) restatement of the primitives on the hot path, written as plain loops over
kernel taps so that nothing is shared with either torch's conv3d or the CUDA kernels:

  conv3d / conv3d_dgrad / conv3d_wgrad   nn.Conv3d(bias=False, dilation=1) and its autograd, the call at
                                         /root/reference/src/models/R2Plus1D.py:44-51,57
  bn_lrelu_fwd / bn_lrelu_bwd            nn.BatchNorm3d in training mode (biased variance to normalise,
                                         unbiased variance into running_var, momentum 0.1, eps 1e-5) followed
                                         by nn.LeakyReLU(slope): R2Plus1D.py:53-54,57
  loss_and_grad                          CELoss / FocalLoss / LDAMLoss values and logit gradients,
                                         /root/reference/src/loss.py:25-34, 52-69, 80-81

Only tests/ import this file (tests/test_oracle_golden.py checks it against the torch-based port, and the
GPU kernel tests use it for tiny known-answer cases).  Sizes must stay small: these are O(taps) numpy passes.
"""
from __future__ import annotations

import numpy as np


def _out_dim(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def conv3d(x, w, stride, padding):
    """x (B,C,T,H,W), w (K,C,kt,kh,kw) -> (B,K,To,Ho,Wo)."""
    B, C, T, H, W = x.shape
    K, _, kt, kh, kw = w.shape
    st, sh, sw = stride
    pt, ph, pw = padding
    To, Ho, Wo = _out_dim(T, kt, st, pt), _out_dim(H, kh, sh, ph), _out_dim(W, kw, sw, pw)
    xp = np.zeros((B, C, T + 2 * pt, H + 2 * ph, W + 2 * pw), dtype=np.float64)
    xp[:, :, pt:pt + T, ph:ph + H, pw:pw + W] = x
    y = np.zeros((B, K, To, Ho, Wo), dtype=np.float64)
    for dt in range(kt):
        for dh in range(kh):
            for dw in range(kw):
                patch = xp[:, :, dt:dt + st * (To - 1) + 1:st, dh:dh + sh * (Ho - 1) + 1:sh, dw:dw + sw * (Wo - 1) + 1:sw]
                y += np.einsum("bcthw,kc->bkthw", patch, w[:, :, dt, dh, dw])
    return y


def conv3d_dgrad(dy, w, x_shape, stride, padding):
    B, C, T, H, W = x_shape
    K, _, kt, kh, kw = w.shape
    st, sh, sw = stride
    pt, ph, pw = padding
    _, _, To, Ho, Wo = dy.shape
    dxp = np.zeros((B, C, T + 2 * pt, H + 2 * ph, W + 2 * pw), dtype=np.float64)
    for dt in range(kt):
        for dh in range(kh):
            for dw in range(kw):
                dxp[:, :, dt:dt + st * (To - 1) + 1:st, dh:dh + sh * (Ho - 1) + 1:sh, dw:dw + sw * (Wo - 1) + 1:sw] += \
                    np.einsum("bkthw,kc->bcthw", dy, w[:, :, dt, dh, dw])
    return dxp[:, :, pt:pt + T, ph:ph + H, pw:pw + W]


def conv3d_wgrad(x, dy, w_shape, stride, padding):
    B, C, T, H, W = x.shape
    K, _, kt, kh, kw = w_shape
    st, sh, sw = stride
    pt, ph, pw = padding
    _, _, To, Ho, Wo = dy.shape
    xp = np.zeros((B, C, T + 2 * pt, H + 2 * ph, W + 2 * pw), dtype=np.float64)
    xp[:, :, pt:pt + T, ph:ph + H, pw:pw + W] = x
    dw_ = np.zeros(w_shape, dtype=np.float64)
    for dt in range(kt):
        for dh in range(kh):
            for dw in range(kw):
                patch = xp[:, :, dt:dt + st * (To - 1) + 1:st, dh:dh + sh * (Ho - 1) + 1:sh, dw:dw + sw * (Wo - 1) + 1:sw]
                dw_[:, :, dt, dh, dw] = np.einsum("bkthw,bcthw->kc", dy, patch)
    return dw_


def bn_lrelu_fwd(y, gamma, beta, slope, eps=1e-5, momentum=0.1, running_mean=None, running_var=None):
    """Train-mode BatchNorm over (B,T,H,W) per channel, then LeakyReLU.  Returns (z, cache)."""
    K = y.shape[1]
    axes = (0, 2, 3, 4)
    n = y.size // K
    mean = y.mean(axis=axes)
    var = y.var(axis=axes)                      # biased: used to normalise
    rstd = 1.0 / np.sqrt(var + eps)
    sh = (1, K, 1, 1, 1)
    xhat = (y - mean.reshape(sh)) * rstd.reshape(sh)
    u = xhat * gamma.reshape(sh) + beta.reshape(sh)
    z = np.where(u > 0, u, u * slope)
    rm = np.zeros(K) if running_mean is None else running_mean
    rv = np.ones(K) if running_var is None else running_var
    cache = dict(xhat=xhat, rstd=rstd, gamma=gamma, u=u, slope=slope, n=n,
                 running_mean=(1 - momentum) * rm + momentum * mean,
                 running_var=(1 - momentum) * rv + momentum * var * n / max(n - 1, 1))
    return z, cache


def bn_lrelu_bwd(dz, cache):
    """Returns (dy, dgamma, dbeta)."""
    xhat, rstd, gamma, u, slope = cache["xhat"], cache["rstd"], cache["gamma"], cache["u"], cache["slope"]
    K = xhat.shape[1]
    sh = (1, K, 1, 1, 1)
    axes = (0, 2, 3, 4)
    g = dz * np.where(u > 0, 1.0, slope)
    dbeta = g.sum(axis=axes)
    dgamma = (g * xhat).sum(axis=axes)
    n = cache["n"]
    dy = (gamma * rstd).reshape(sh) * (g - dbeta.reshape(sh) / n - xhat * dgamma.reshape(sh) / n)
    return dy, dgamma, dbeta


def loss_and_grad(kind, logits, target, weight=None, gamma=2.0, s=1.0, margins=None):
    """kind in {'ce','focal','ldam'}; float64.  Returns (loss, dloss/dlogits) with the reference's reductions:
    CE and Focal are sums over the batch, LDAM is the weighted mean."""
    z = np.asarray(logits, dtype=np.float64).copy()
    n, C = z.shape
    y = np.asarray(target)
    rows = np.arange(n)
    w = np.ones(C) if weight is None else np.asarray(weight, dtype=np.float64)
    wy = w[y]
    scale = 1.0
    if kind == "ldam":
        z[rows, y] -= np.asarray(margins, dtype=np.float64)[y]
        z *= s
        scale = s
    zmax = z.max(axis=1, keepdims=True)
    lse = zmax[:, 0] + np.log(np.exp(z - zmax).sum(axis=1))
    ce = lse - z[rows, y]
    sm = np.exp(z - lse[:, None])
    dce_dz = sm.copy()
    dce_dz[rows, y] -= 1.0
    if kind == "ce":
        return float((wy * ce).sum()), wy[:, None] * dce_dz
    if kind == "focal":
        p = np.exp(-ce)
        one_m = 1.0 - p
        if gamma == 0:
            f, df = np.ones_like(p), np.zeros_like(p)
        else:
            f = one_m ** gamma
            df = gamma * one_m ** (gamma - 1.0) * p       # d(1-p)^gamma / dCE
        return float((wy * f * ce).sum()), (wy * (f + df * ce))[:, None] * dce_dz
    if kind == "ldam":
        if weight is None:
            return float(ce.mean()), scale * dce_dz / n
        return float((wy * ce).sum() / wy.sum()), scale * (wy / wy.sum())[:, None] * dce_dz
    raise ValueError(kind)
