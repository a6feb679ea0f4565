"""The image-quality numbers the reference's validation step reports, evaluated on the device the images live on.

Reference: engines/base.py:255-268 (round -> shave for SR -> metrics), utils/utils_image.py:8-11 (shave), :30-33
(tensor_round), :43-80 (rgb2ycbcr, MATLAB coefficients, rounded to 8 bit), utils/metrics/psnr.py:44-48 (psnr),
utils/metrics/ssim.py:17-82 (Gaussian-window SSIM, 11 taps, sigma 1.5, taps rounded to 6 decimals, zero padding).
psnr_fused() is the hand-written kernel (csrc/metric.cu, grl_psnr_f32) the benchmarked validation step uses for PSNR
(RGB and luma); the torch-op functions below define the same quantities on any device (SSIM stays torch ops) and are
what the CPU tests pin against the reference's own functions.
"""
import math

import torch
import torch.nn.functional as F


def tensor_round(img, data_range=1.0):
    img = img.clamp(0.0, 1.0 * data_range)
    return (img * 255.0 / data_range).round() * data_range / 255.0


def shave(img, border):
    return img[..., border:-border, border:-border] if border > 0 else img


def rgb_to_y(img, data_range=1.0):
    """Luma of MATLAB's rgb2ycbcr on (B, 3, H, W), rounded to the 8-bit grid, returned as (B, 1, H, W)."""
    scale = 255.0 if data_range == 1.0 else 1.0
    coeff = torch.tensor([65.481, 128.553, 24.966], device=img.device, dtype=img.dtype) / 255.0
    y = (img * scale).permute(0, 2, 3, 1) @ coeff + 16.0
    y = y.round().unsqueeze(1)
    return y / 255.0 if data_range == 1 else y


def psnr(restored, target, border=0, channel="rgb"):
    """Per-image PSNR (B,) of 8-bit-rounded tensors, `border` pixels shaved (SR: border = scale), RGB or luma."""
    a, b = shave(tensor_round(restored), border), shave(tensor_round(target), border)
    if channel == "y":
        a, b = rgb_to_y(a), rgb_to_y(b)
    return -10 * (a - b).pow(2).mean([-3, -2, -1]).log10()


def psnr_fused(restored, target, border=0):
    """(psnr_rgb, psnr_y), each (B,), from ONE fused kernel over CUDA fp32 (B, C, H, W) images: tensor_round + shave +
    exact integer squared-error reduction (csrc/metric.cu).  No torch math on the images."""
    from . import capi

    capi.require_device(restored)
    capi.require_device(target)
    if restored.shape != target.shape or restored.dim() != 4:
        raise RuntimeError(f"grl_b200: psnr_fused needs two (B, C, H, W) tensors of one shape, got {tuple(restored.shape)} / {tuple(target.shape)}")
    a = restored if (restored.dtype == torch.float32 and restored.is_contiguous()) else restored.float().contiguous()
    b = target if (target.dtype == torch.float32 and target.is_contiguous()) else target.float().contiguous()
    B, C, H, W = a.shape
    ws = torch.empty(2 * max(B, 1), device=a.device, dtype=torch.int64)
    out = torch.empty(2, B, device=a.device, dtype=torch.float32)
    capi.check(capi.lib().grl_psnr_f32(capi.ptr(a), capi.ptr(b), B, C, H, W, int(border), capi.ptr(ws), ws.numel() * 8,
                                       capi.ptr(out[0]), capi.ptr(out[1]), capi.stream()))
    return out[0], out[1]


def _gaussian_window(channels, size, sigma, like):
    taps = torch.tensor([round(math.exp(-((i - size // 2) ** 2) / (2.0 * sigma**2)), 6) for i in range(size)],
                        dtype=torch.float64)
    taps = taps / taps.sum()
    win = torch.outer(taps, taps).float()
    return win.expand(channels, 1, size, size).contiguous().to(device=like.device, dtype=like.dtype)


def ssim(restored, target, border=0, channel="rgb", window_size=11, sigma=1.5):
    """Per-image SSIM (B,): mean of the SSIM map over channels and pixels (zero-padded Gaussian local statistics)."""
    a, b = shave(tensor_round(restored), border), shave(tensor_round(target), border)
    if channel == "y":
        a, b = rgb_to_y(a), rgb_to_y(b)
    c = a.shape[1]
    win = _gaussian_window(c, window_size, sigma, a)

    def blur(t):
        return F.conv2d(t, win, padding=window_size // 2, groups=c)

    mu_a, mu_b = blur(a), blur(b)
    var_a, var_b, cov = blur(a * a) - mu_a.pow(2), blur(b * b) - mu_b.pow(2), blur(a * b) - mu_a * mu_b
    c1, c2 = 0.01**2, 0.03**2
    ssim_map = ((2 * mu_a * mu_b + c1) * (2 * cov + c2)) / ((mu_a.pow(2) + mu_b.pow(2) + c1) * (var_a + var_b + c2))
    return ssim_map.mean([-3, -2, -1])


def validation_metrics(restored, target, scale=1, is_sr=False):
    """dict of per-image (B,) tensors: psnr, psnr_y, ssim, ssim_y, as validation_step + the metric collection yield them."""
    border = scale if is_sr else 0
    return {
        "psnr": psnr(restored, target, border, "rgb"),
        "psnr_y": psnr(restored, target, border, "y"),
        "ssim": ssim(restored, target, border, "rgb"),
        "ssim_y": ssim(restored, target, border, "y"),
    }
