"""Host-side geometry of the attention levels: stripe sizes/shifts, GrlGrid construction, and the
reference-shaped tables / indices / masks expanded from the closed forms in csrc/grl_geometry.h.

The product path only ever needs `coords_table` (the input of the CPB MLP); `position_index` and
`shift_mask` exist so callers that follow the reference API (WindowAttention.forward(qkv, x_size, table,
index, mask)) can still obtain those tensors, and so the tests can prove the closed forms bit-exact.
Reference: models/common/ops.py:76-157,:225-271,:308-375; mixed_attn_block_efficient.py:61-70.
"""
import ctypes

import torch

from . import capi


def to_2tuple(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v)


def stripe_info(stripe_size_in, stripe_groups_in, stripe_shift, input_resolution):
    """(stripe_size, shift_size) for one block; mixed_attn_block_efficient.py:61-70."""
    size, shift = [], []
    for s, g, d in zip(stripe_size_in, stripe_groups_in, input_resolution):
        if g is None:
            size.append(s)
            shift.append(s // 2 if stripe_shift else 0)
        else:
            size.append(d // g)
            shift.append(0 if g == 1 else d // (g * 2))
    return size, shift


_get_stripe_info = stripe_info  # the reference's name


def coords_table(window_size, df=1):
    """get_relative_coords_table_all (pretrained size 0): (1, wh+awh-1, ww+aww-1, 2) fp32 on the CPU."""
    wh, ww = window_size
    nh, nw = wh + wh // df - 1, ww + ww // df - 1
    out = torch.empty(nh * nw, 2, dtype=torch.float32)
    capi.check(capi.lib().grl_coords_table_host(wh, ww, df, ctypes.c_void_p(out.data_ptr())))
    return out.view(1, nh, nw, 2)


def position_index(window_size, df=1, window_to_anchor=True):
    """get_relative_position_index_simple: (n1, n2) int64 on the CPU."""
    wh, ww = window_size
    n_w, n_a = wh * ww, (wh // df) * (ww // df)
    n1, n2 = (n_w, n_a) if window_to_anchor else (n_a, n_w)
    out = torch.empty(n1, n2, dtype=torch.int64)
    capi.check(capi.lib().grl_rel_index_host(wh, ww, df, int(window_to_anchor), ctypes.c_void_p(out.data_ptr())))
    return out


def shift_mask(input_resolution, window_size, shift_size, df=1, window_to_anchor=True):
    """calculate_mask / calculate_mask_all: (nW, n1, n2) fp32 of 0 / -100 on the CPU."""
    H, W = input_resolution
    wh, ww = window_size
    sh, sw = to_2tuple(shift_size)
    n_w, n_a = wh * ww, (wh // df) * (ww // df)
    n1, n2 = (n_w, n_a) if window_to_anchor else (n_a, n_w)
    out = torch.empty((H // wh) * (W // ww), n1, n2, dtype=torch.float32)
    capi.check(capi.lib().grl_shift_mask_host(H, W, wh, ww, sh, sw, df, int(window_to_anchor),
                                              ctypes.c_void_p(out.data_ptr())))
    return out


def token_grid(x_size, window_size, shift):
    return capi.grid(x_size[0], x_size[1], window_size[0], window_size[1], shift[0], shift[1])


def anchor_grid(x_size, window_size, shift, df):
    return capi.grid(x_size[0] // df, x_size[1] // df, window_size[0] // df, window_size[1] // df,
                     shift[0] // df, shift[1] // df)
