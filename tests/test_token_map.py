"""Roll + window-partition addressing (csrc/grl_geometry.h: locate) against the oracle's tensor construction, and the box
condition of the experimental TMA-producer attention kernel (attn_tc_tma.cu): CPU only, through the C ABI."""
import ctypes

import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st


def token_map(pkg, H, W, wh, ww, sh, sw):
    from grl_image_restoration_b200 import capi

    g = capi.GrlGrid(H, W, wh, ww, sh, sw)
    out = torch.empty((H // wh) * (W // ww), wh * ww, dtype=torch.int32)
    capi.check(capi.lib().grl_token_map_host(g, ctypes.c_void_p(out.data_ptr())))
    return g, out


@st.composite
def _grid(draw):
    wh, ww = draw(st.integers(1, 12)), draw(st.sampled_from([2, 4, 6, 8, 12, 16, 24, 32, 36, 64]))
    nh, nw = draw(st.integers(1, 3)), draw(st.integers(1, 3))
    sh = draw(st.integers(0, wh - 1))
    sw = draw(st.sampled_from([0, ww // 2, ww // 4, 1 if ww > 1 else 0, draw(st.integers(0, ww - 1))]))
    return nh * wh, nw * ww, wh, ww, sh, sw


@settings(max_examples=150, deadline=None)
@given(_grid())
def test_token_map_is_roll_then_partition(pkg, oracle, grid):
    """torch.roll(x, (-sh, -sw)) followed by window_partition (ops.py:36-53) -- as the oracle does it with tensors."""
    H, W, wh, ww, sh, sw = grid
    _, got = token_map(pkg, H, W, wh, ww, sh, sw)
    flat = torch.arange(H * W, dtype=torch.float32).view(1, H, W, 1)
    want = oracle.partition(torch.roll(flat, (-sh, -sw), (1, 2)), [wh, ww]).reshape(-1, wh * ww).to(torch.int32)
    assert torch.equal(got, want)


@settings(max_examples=300, deadline=None)
@given(_grid())
def test_tma_boxes_are_contiguous_runs(pkg, grid):
    """Every aligned run of `box` tokens of every window is `box` consecutive pixels of one image row (no wrap inside)."""
    from grl_image_restoration_b200 import capi

    H, W, wh, ww, sh, sw = grid
    g, tm = token_map(pkg, H, W, wh, ww, sh, sw)
    box = capi.lib().grl_tc_attn_box_tokens(g)
    if box == 0:
        return
    assert box >= 8 and box & (box - 1) == 0 and box <= 64 and ww % box == 0 and 64 % box == 0
    runs = tm.view(tm.shape[0], -1, box)  # (window, run, token in run)
    first = runs[..., :1]
    assert torch.equal(runs, first + torch.arange(box, dtype=torch.int32)), "a box must be consecutive pixels"
    assert torch.equal(first // W, (first + box - 1) // W), "a box must stay inside one image row"


@pytest.mark.parametrize("grid,box", [((256, 256, 32, 32, 16, 16), 16), ((256, 256, 32, 32, 0, 0), 32), ((256, 256, 64, 64, 32, 32), 32),
                                      ((128, 128, 32, 32, 16, 16), 16), ((256, 256, 64, 128, 0, 0), 64), ((96, 96, 12, 12, 6, 6), 0),
                                      ((96, 192, 48, 96, 24, 48), 16), ((56, 40, 7, 5, 3, 2), 0), ((64, 64, 8, 8, 4, 4), 0), ((64, 64, 16, 16, 8, 8), 8)])
def test_box_tokens_of_released_geometries(pkg, grid, box):
    from grl_image_restoration_b200 import capi

    assert capi.lib().grl_tc_attn_box_tokens(capi.GrlGrid(*grid)) == box
