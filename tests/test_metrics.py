"""Validation metrics against values produced by the reference's own functions
(tests/golden/metrics.npz, written by oracle/make_golden_metrics.py from utils/metrics/{psnr,ssim}.py and
utils/utils_image.py with the recipe of engines/base.py:255-268)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


@pytest.mark.parametrize("case,is_sr", [("sr_x4", True), ("dn", False)])
def test_validation_metrics_match_reference(pkg, case, is_sr):
    from grl_image_restoration_b200 import metrics

    g = np.load(GOLDEN)
    restored, target = torch.from_numpy(g[f"{case}_restored"]), torch.from_numpy(g[f"{case}_target"])
    keep = restored.clone()
    got = metrics.validation_metrics(restored, target, scale=int(g[f"{case}_border"]) or 1, is_sr=is_sr)
    assert torch.equal(restored, keep), "metrics must not modify the images they are given"
    for name, tol in (("psnr", 1e-4), ("psnr_y", 1e-4), ("ssim", 2e-6), ("ssim_y", 2e-6)):
        want = torch.from_numpy(g[f"{case}_{name}"])
        assert got[name].shape == want.shape
        assert (got[name] - want).abs().max().item() <= tol, (name, got[name], want)


def test_luma_is_on_the_8bit_grid(pkg):
    from grl_image_restoration_b200 import metrics

    x = metrics.tensor_round(torch.rand(2, 3, 9, 7, generator=torch.Generator().manual_seed(0)))
    y = metrics.rgb_to_y(x) * 255.0
    assert y.shape == (2, 1, 9, 7)
    assert torch.equal(y.round(), y) or (y - y.round()).abs().max().item() < 1e-4
    assert y.min().item() >= 16 and y.max().item() <= 235


def test_identical_images(pkg):
    from grl_image_restoration_b200 import metrics

    x = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(1))
    m = metrics.validation_metrics(x, x.clone())
    assert torch.isinf(m["psnr"]).all() and torch.isinf(m["psnr_y"]).all()
    assert (m["ssim"] - 1).abs().max().item() < 1e-6 and (m["ssim_y"] - 1).abs().max().item() < 1e-6
