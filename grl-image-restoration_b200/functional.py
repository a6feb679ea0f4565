"""Thin functional layer over the C ABI (capi.py): allocates outputs with torch, passes raw device pointers +
the current CUDA stream, raises RuntimeError on a non-zero status.  Activations are channels-last fp32."""
import ctypes

import torch

from . import capi

ACT_NONE, ACT_GELU, ACT_LEAKY = 0, 1, 2


class KernelTimer:
    """Optional CUDA-event timer around the attention launches (bench.py's roofline leg).  Events are recorded on
    the stream the kernels are launched on (torch's current stream)."""

    def __init__(self):
        self.pairs = {}

    def wrap(self, tag, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.pairs.setdefault(tag, []).append((e0, e1))
        return out

    def totals_ms(self):
        """{tag: (total ms, launches)} -- call after a synchronize."""
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.pairs.items()}


timer = None  # set to a KernelTimer to time attention kernels


def _timed(tag, fn):
    return fn() if timer is None else timer.wrap(tag, fn)


def _f32c(t, name):
    capi.require_device(t)
    if t.dtype != torch.float32:
        raise RuntimeError(f"grl_b200: {name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def linear(x, weight, bias=None, act=ACT_NONE, slope=0.0, res=None, out=None):
    """x (..., K) -> (..., N): y = act(x W^T + b) (+ res)."""
    x = _f32c(x, "x")
    K = x.shape[-1]
    N = weight.shape[0]
    M = x.numel() // K
    y = out if out is not None else torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
    if res is not None:
        res = _f32c(res, "res")
    capi.check(capi.lib().grl_linear_f32(capi.ptr(x), K, capi.ptr(_f32c(weight, "weight")), capi.ptr(bias),
                                         capi.ptr(res), N, capi.ptr(y), N, M, N, K, act, slope, capi.stream()))
    return y


def pack_conv_weight(weight):
    """(Cout, Cin, 3, 3) -> (Cout, 9*Cin) with k = (ky*3+kx)*Cin + c (the im2col order of the kernels)."""
    co, ci, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise RuntimeError("grl_b200: only 3x3 convolutions are on this path")
    return weight.detach().permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous()


def conv3x3(x, wpacked, bias=None, act=ACT_NONE, slope=0.0, res=None):
    """x (B, H, W, Cin) channels-last -> (B, H, W, Cout); stride 1, zero pad 1."""
    x = _f32c(x, "x")
    B, H, W, Cin = x.shape
    Cout = wpacked.shape[0]
    y = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float32)
    if res is not None:
        res = _f32c(res, "res")
    capi.check(capi.lib().grl_conv3x3_f32(capi.ptr(x), capi.ptr(wpacked), capi.ptr(bias), capi.ptr(res), capi.ptr(y),
                                          B, H, W, Cin, Cout, act, slope, capi.stream()))
    return y


def avgpool(x, df):
    x = _f32c(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, H // df, W // df, C, device=x.device, dtype=torch.float32)
    capi.check(capi.lib().grl_avgpool_f32(capi.ptr(x), capi.ptr(y), B, H, W, C, df, capi.stream()))
    return y


def ln_residual(x, u, gamma, beta, eps=1e-5, res_scale=1.0, cab_y=None, cab_gate=None):
    """(x or 0) + res_scale * LN(u) (+ cab_y * gate[b]); x, u (B, L, C)."""
    u = _f32c(u, "u")
    B, L, C = u.shape
    out = torch.empty_like(u)
    capi.check(capi.lib().grl_ln_residual_f32(
        capi.ptr(_f32c(x, "x")) if x is not None else None, capi.ptr(u), capi.ptr(gamma), capi.ptr(beta), eps,
        res_scale, capi.ptr(_f32c(cab_y, "cab_y")) if cab_y is not None else None,
        capi.ptr(cab_gate) if cab_gate is not None else None, L, capi.ptr(out), B * L, C, capi.stream()))
    return out


def channel_gate(y, w1, b1, w2, b2):
    """y (B, L, C) -> gate (B, C) = sigmoid(W2 relu(W1 mean_L(y) + b1) + b2)."""
    y = _f32c(y, "y")
    B, L, C = y.shape
    R = w1.shape[0]
    nbytes = capi.lib().grl_channel_gate_workspace(B, L, C)
    ws = torch.empty(max(nbytes, 4) // 4, device=y.device, dtype=torch.float32)
    gate = torch.empty(B, C, device=y.device, dtype=torch.float32)
    capi.check(capi.lib().grl_channel_gate_f32(capi.ptr(y), B, L, C, capi.ptr(w1), capi.ptr(b1), capi.ptr(w2),
                                               capi.ptr(b2), R, capi.ptr(gate), capi.ptr(ws), nbytes, capi.stream()))
    return gate


def bias_table(table, w1, b1, w2):
    """table (..., 2) -> activated bias (heads, rows) = 16*sigmoid(cpb_mlp(table))."""
    t = _f32c(table, "table").reshape(-1, 2)
    heads, hidden = w2.shape
    out = torch.empty(heads, t.shape[0], device=t.device, dtype=torch.float32)
    capi.check(capi.lib().grl_bias_table_f32(capi.ptr(t), t.shape[0], capi.ptr(_f32c(w1, "w1")), capi.ptr(b1),
                                             capi.ptr(_f32c(w2, "w2")), hidden, heads, capi.ptr(out), capi.stream()))
    return out


def affine_(attn, logit_scale, bias, index, mask):
    """In-place AffineTransform on a materialised (B_, heads, n1, n2) map."""
    attn = _f32c(attn, "attn")
    B_, H, n1, n2 = attn.shape
    nW = mask.shape[0] if mask is not None else 0
    capi.check(capi.lib().grl_affine_f32(capi.ptr(attn), B_, H, n1, n2, capi.ptr(logit_scale.reshape(-1)),
                                         capi.ptr(bias), bias.shape[1], capi.ptr(index.contiguous()),
                                         capi.ptr(_f32c(mask, "mask")) if mask is not None else None, nW,
                                         capi.stream()))
    return attn


def _token_rows(t, name):
    capi.require_device(t)
    if t.dtype != torch.float32 or t.dim() != 3 or t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
        raise RuntimeError(f"grl_b200: {name} must be float32 (B, L, c) with unit channel stride and packed rows")
    return ctypes.c_void_p(t.data_ptr()), t.stride(1)


def window_attention(qkv, B, grid, heads, logit_scale, bias, use_mask, out=None):
    """qkv (B, L, 3c) view (window half) -> (B, L, c)."""
    qp, ldq = _token_rows(qkv, "qkv")
    c = qkv.shape[2] // 3
    if out is None:
        out = torch.empty(B, qkv.shape[1], c, device=qkv.device, dtype=torch.float32)
    op, ldo = _token_rows(out, "out")
    _timed("window_attn", lambda: capi.check(capi.lib().grl_window_attn_f32(
        qp, ldq, op, ldo, B, grid, heads, c // heads, capi.ptr(logit_scale.reshape(-1)), capi.ptr(bias),
        int(use_mask), capi.stream())))
    return out


def stripe_attention(qkv, anchor, B, tok_grid, anc_grid, heads, scale1, bias1, scale2, bias2, use_mask, out=None):
    """qkv (B, L, 3c) view (stripe half), anchor (B, Ha, Wa, c) -> (B, L, c)."""
    qp, ldq = _token_rows(qkv, "qkv")
    c = qkv.shape[2] // 3
    anchor = _f32c(anchor, "anchor")
    if out is None:
        out = torch.empty(B, qkv.shape[1], c, device=qkv.device, dtype=torch.float32)
    op, ldo = _token_rows(out, "out")
    d = c // heads
    nbytes = capi.lib().grl_stripe_attn_workspace(B, tok_grid, anc_grid, heads, d)
    ws = torch.empty(max(nbytes, 4) // 4, device=qkv.device, dtype=torch.float32)
    _timed("stripe_attn", lambda: capi.check(capi.lib().grl_stripe_attn_f32(
        qp, ldq, capi.ptr(anchor), c, op, ldo, B, tok_grid, anc_grid, heads, d, capi.ptr(scale1.reshape(-1)),
        capi.ptr(bias1), capi.ptr(scale2.reshape(-1)), capi.ptr(bias2), int(use_mask), capi.ptr(ws), nbytes,
        capi.stream())))
    return out
