"""TEST INFRASTRUCTURE ONLY. Writes tests/golden/metrics.npz from the UNMODIFIED reference metric functions.

Runs only in the build container (needs /root/reference).  The reference's metric files import torchmetrics (absent
here) for their Metric base class only; a stand-in module with an empty `Metric` class is registered so that
utils/metrics/psnr.py:44 (psnr), utils/metrics/ssim.py:73 (ssim) and utils/utils_image.py:8,30,43 (shave,
tensor_round, rgb2ycbcr) can be imported and called as they are.  The recipe mirrors engines/base.py:255-268.

    python oracle/make_golden_metrics.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import REF_ROOT, reference_available  # noqa: E402


def main():
    if not reference_available():
        raise SystemExit("reference not mounted: golden metrics can only be regenerated in the build container")
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.Metric = type("Metric", (torch.nn.Module,), {})
        sys.modules["torchmetrics"] = tm
    sys.path.insert(0, REF_ROOT)
    from utils.metrics.psnr import psnr as ref_psnr
    from utils.metrics.ssim import ssim as ref_ssim
    from utils.utils_image import rgb2ycbcr, shave, tensor_round

    g = torch.Generator().manual_seed(1234)
    out = {}
    for name, (b, h, w, border, noise) in {"sr_x4": (3, 48, 40, 4, 0.08), "dn": (2, 37, 53, 0, 0.2)}.items():
        target = torch.rand(b, 3, h, w, generator=g)
        restored = target + noise * torch.randn(b, 3, h, w, generator=g)  # leaves [0, 1]: exercises the clamp
        out[f"{name}_restored"], out[f"{name}_target"] = restored.numpy(), target.numpy()
        r, t = tensor_round(restored.clone(), 1.0), tensor_round(target.clone(), 1.0)
        r, t = shave(r, border), shave(t, border)
        ry, ty = rgb2ycbcr(r, 1.0), rgb2ycbcr(t, 1.0)
        out[f"{name}_psnr"] = ref_psnr(r, t).numpy()
        out[f"{name}_psnr_y"] = ref_psnr(ry, ty).numpy()
        out[f"{name}_ssim"] = np.array([ref_ssim(p.unsqueeze(0), q.unsqueeze(0)).item() for p, q in zip(r, t)], np.float32)
        out[f"{name}_ssim_y"] = np.array([ref_ssim(p.unsqueeze(0), q.unsqueeze(0)).item() for p, q in zip(ry, ty)],
                                         np.float32)
        out[f"{name}_border"] = np.array(border)
    path = os.path.join(HERE, "..", "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), {k: v.tolist() for k, v in out.items() if "psnr" in k or "ssim" in k})


if __name__ == "__main__":
    main()
