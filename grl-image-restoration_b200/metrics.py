"""The PSNR the reference's validation step reports, evaluated on the device.
utils/utils_image.py:30-33 (tensor_round), utils/metrics/psnr.py:44-48 (psnr), engines/base.py:265-267 (SR shave)."""
import torch


def tensor_round(img, data_range=1.0):
    img = img.clamp(0.0, 1.0 * data_range)
    return (img * 255.0 / data_range).round() * data_range / 255.0


def psnr(restored, target, border=0):
    """Per-image PSNR (B,) of 8-bit-rounded tensors, with `border` pixels shaved (SR uses border = scale)."""
    a, b = tensor_round(restored), tensor_round(target)
    if border > 0:
        a, b = a[..., border:-border, border:-border], b[..., border:-border, border:-border]
    return -10 * (a - b).pow(2).mean([-3, -2, -1]).log10()
