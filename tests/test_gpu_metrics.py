"""The fused PSNR kernel (csrc/metric.cu behind grl_psnr_f32) against values produced by the reference's own functions
(tests/golden/metrics.npz) and against the torch-op definitions in metrics.py on random images."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


@pytest.mark.parametrize("case", ["sr_x4", "dn"])
def test_fused_psnr_matches_reference_functions(pkg, device, case):
    from grl_image_restoration_b200 import metrics

    g = np.load(GOLDEN)
    restored, target = torch.from_numpy(g[f"{case}_restored"]).to(device), torch.from_numpy(g[f"{case}_target"]).to(device)
    keep = restored.clone()
    border = int(g[f"{case}_border"])
    p, py = metrics.psnr_fused(restored, target, border)
    assert torch.equal(restored, keep)
    assert (p.cpu() - torch.from_numpy(g[f"{case}_psnr"])).abs().max().item() <= 1e-4
    assert (py.cpu() - torch.from_numpy(g[f"{case}_psnr_y"])).abs().max().item() <= 1e-4


@pytest.mark.parametrize("shape,border", [((3, 3, 67, 45), 0), ((2, 3, 256, 256), 4), ((1, 1, 40, 33), 2), ((16, 3, 1024, 1024), 4)])
def test_fused_psnr_vs_torch_ops(pkg, device, shape, border):
    from grl_image_restoration_b200 import metrics

    g = torch.Generator().manual_seed(5)
    a = (torch.rand(shape, generator=g) * 1.3 - 0.15).to(device)   # values outside [0, 1] exercise the clamp
    b = torch.rand(shape, generator=g).to(device)
    p, py = metrics.psnr_fused(a, b, border)
    want = metrics.psnr(a, b, border)
    assert (p - want).abs().max().item() <= 2e-4
    if shape[1] == 3:
        assert (py - metrics.psnr(a, b, border, "y")).abs().max().item() <= 2e-3  # luma: rare round-to-8-bit ties
    p2, _ = metrics.psnr_fused(a, b, border)
    assert torch.equal(p, p2)  # integer accumulation: bit-identical run to run
    same, _ = metrics.psnr_fused(b, b.clone(), border)
    assert torch.isinf(same).all()
