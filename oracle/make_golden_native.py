"""TEST INFRASTRUCTURE.  Golden outputs of the UNMODIFIED reference at BASELINE.json's native shapes
(cfg2 Small x4 256^2, cfg3 Base DN sigma50 256^2, cfg4 Base x4 256^2, cfg5 Base deblur 480^2 tile and the whole
1280x720 frame through the engine's forward_tile loop).  Runs only where /root/reference exists:

    python oracle/make_golden_native.py [--cases cfg4_init,...] [--oracle] [--frame]

Each case runs `models.networks.grl.GRL` (reference, fp32, CPU, all cores) once on seeded synthetic weights
(oracle.synth_state_dict, styles "init" = the reference constructor's own distribution and "spread" = the harsh
fixture weights) and the seeded synthetic input of SURVEY.md 8d, and stores compactly in tests/golden/native_<case>.npz:

  sub          fp32 output[..., ::s, ::s]   (s coprime with the pixel-shuffle factor, so every phase is sampled)
  stride       s
  psnr_ref_gt  PSNR(reference output, GT) per image with the reference's own definition over the FULL output
               (utils/utils_image.py:30-33 tensor_round, engines/base.py:265-267 border shave,
               utils/metrics/psnr.py:44-48), GT = torch.rand(shape, seed 9)
  sha256       digest of the full fp32 output (informational)
  oracle_err   max-abs |oracle - reference| over the full output when --oracle was given (pins the restatement
               at native size), else -1
The GPU tests (tests/test_gpu_native_shapes.py) regenerate weights/inputs/GT from the same seeds.
"""
import argparse
import hashlib
import importlib.util
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import grl_oracle as orc  # noqa: E402
from _pkgload import load_package  # noqa: E402
from _ref_import import import_reference  # noqa: E402

pkg = load_package()
configs = pkg.configs
GOLD = os.path.join(ROOT, "tests", "golden")

# name: (variant, task, scale, img_size, (H, W), sigma, stride)
SHAPES = {
    "cfg2": ("small", "sr", 4, 256, (256, 256), 0.0, 5),
    "cfg3": ("base", "dn", 1, 256, (256, 256), 50.0, 1),
    "cfg4": ("base", "sr", 4, 256, (256, 256), 0.0, 5),
    "cfg5": ("base", "deblur", 1, 480, (480, 480), 0.0, 2),
}
FRAME = dict(name="cfg5_frame", model=("base", "deblur", 1, 480), hw=(720, 1280), tile=480, overlap=48, stride=3)
GT_SEED = 9


def load_ref_utils():
    spec = importlib.util.spec_from_file_location("ref_utils_image", "/root/reference/utils/utils_image.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_psnr(ui, y, gt, border):
    a, b = ui.tensor_round(y.clone()), ui.tensor_round(gt.clone())
    if border > 0:
        a, b = ui.shave(a, border), ui.shave(b, border)
    return -10 * (a - b).pow(2).mean([-3, -2, -1]).log10()  # utils/metrics/psnr.py:44-48


def build_reference(grl_mod, cfg, sd):
    torch.manual_seed(0)
    m = grl_mod.GRL(**cfg).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.split("_")[0] in ("table", "index", "mask") for k in missing), missing
    return m


def engine_forward_tile(model, x, tile, overlap, scale):
    """engines/base.py:90-116 (the engine itself needs pytorch_lightning, absent here): same loop, same order."""
    b, c, h, w = x.shape
    tile = min(tile, h, w)
    stride = tile - overlap
    h_idx = list(range(0, h - tile, stride)) + [h - tile]
    w_idx = list(range(0, w - tile, stride)) + [w - tile]
    E = torch.zeros(b, c, h * scale, w * scale)
    W = torch.zeros_like(E)
    for hi in h_idx:
        for wi in w_idx:
            out = model(x[..., hi:hi + tile, wi:wi + tile])
            E[..., hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale].add_(out)
            W[..., hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale].add_(torch.ones_like(out))
    return E.div_(W)


def save(name, y, stride, psnr_val, oracle_err, extra=None):
    arrs = dict(sub=y[..., ::stride, ::stride].contiguous().numpy(), stride=np.int64(stride),
                psnr_ref_gt=psnr_val.numpy().astype(np.float64), oracle_err=np.float64(oracle_err),
                sha256=np.array(hashlib.sha256(np.ascontiguousarray(y.numpy()).tobytes()).hexdigest()),
                shape=np.array(y.shape, dtype=np.int64))
    if extra:
        arrs.update(extra)
    np.savez_compressed(os.path.join(GOLD, f"native_{name}.npz"), **arrs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg4_init,cfg3_init,cfg2_init,cfg5_init,cfg4_spread,cfg3_spread,cfg2_spread,cfg5_spread")
    ap.add_argument("--oracle", action="store_true", help="also run the oracle restatement and record |oracle-ref|")
    ap.add_argument("--frame", action="store_true", help="also the 1280x720 frame through the engine's tile loop")
    args = ap.parse_args()
    grl_mod, _, _, _ = import_reference()
    ui = load_ref_utils()
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    for case in [c for c in args.cases.split(",") if c]:
        shape_name, style = case.split("_")
        variant, task, scale, img_size, hw, sigma, stride = SHAPES[shape_name]
        cfg = configs.grl_config(variant, task, scale, img_size)
        sd = orc.synth_state_dict(cfg, seed=0, style=style)
        x = orc.synth_input((1, 3, *hw), seed=1234, noise_sigma=sigma)
        t0 = time.time()
        ref = build_reference(grl_mod, cfg, sd)
        with torch.no_grad():
            y = ref(x.clone())
        del ref
        t1 = time.time()
        gt = torch.rand(y.shape, generator=torch.Generator().manual_seed(GT_SEED))
        p = ref_psnr(ui, y, gt, scale if scale > 1 else 0)
        assert torch.equal(p, orc.psnr(y, gt, scale if scale > 1 else 0))
        err = -1.0
        if args.oracle:
            with torch.no_grad():
                yo = orc.grl_forward(sd, cfg, x.clone())
            err = (yo - y).abs().max().item()
            assert err <= 1e-5 * max(1.0, y.abs().max().item()), (case, err)
        save(case, y, stride, p, err)
        print(f"[native] {case}: out {tuple(y.shape)} range {y.min():.3f}..{y.max():.3f} psnr_ref_gt {p.tolist()} "
              f"oracle_err {err:.3e} ref {t1 - t0:.0f}s total {time.time() - t0:.0f}s", flush=True)
    if args.frame:
        variant, task, scale, img_size = FRAME["model"]
        cfg = configs.grl_config(variant, task, scale, img_size)
        sd = orc.synth_state_dict(cfg, seed=0, style="init")
        x = orc.synth_input((1, 3, *FRAME["hw"]), seed=1234)
        t0 = time.time()
        ref = build_reference(grl_mod, cfg, sd)
        with torch.no_grad():
            y = engine_forward_tile(ref, x, FRAME["tile"], FRAME["overlap"], 1)
        gt = torch.rand(y.shape, generator=torch.Generator().manual_seed(GT_SEED))
        p = ref_psnr(ui, y, gt, 0)
        save(FRAME["name"], y, FRAME["stride"], p, -1.0,
             dict(tile=np.int64(FRAME["tile"]), overlap=np.int64(FRAME["overlap"])))
        print(f"[native] {FRAME['name']}: out {tuple(y.shape)} psnr_ref_gt {p.tolist()} {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    main()
