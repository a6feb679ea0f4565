"""TEST INFRASTRUCTURE.  Pins oracle/grl_oracle.py against the UNMODIFIED reference and writes the
golden fixtures under tests/golden/.  Runs only where /root/reference exists (the build container):

    python oracle/make_golden.py            # validate + (re)write fixtures
    python oracle/make_golden.py --check    # validate only

Fixtures (all produced by the reference's own code, never by the oracle):
  tests/golden/geometry.json   sha256 of every table / index / mask the reference builds for the
                               geometries of SURVEY.md Appendix B (+ small ones stored in full in
                               geometry_small.npz)
  tests/golden/model_<name>.npz  reference GRL outputs (and per-module taps of one block) for seeded
                               synthetic weights/inputs (oracle.synth_state_dict / synth_input)
"""
import argparse
import hashlib
import json
import os
import sys

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

# (x_size, window, stripe(H-type), groups, df) -- Appendix B geometries + awkward ones
GEOMETRIES = {
    "sr_base_256": ((256, 256), 32, [64, 64], [None, None], 2),
    "sr_small_128": ((128, 128), 32, [64, 64], [None, None], 4),
    "dn_base_128x256": ((128, 256), 32, [64, 128], [None, None], 2),
    "dn_small_128": ((128, 128), 16, [64, 128], [None, None], 4),
    "deblur_96x192": ((96, 192), 12, [48, 96], [None, None], 4),
    "jpeg_144": ((144, 144), 36, [72, 144], [None, None], 4),
    "dm_64": ((64, 64), 8, [32, 32], [None, None], 4),
    "yaml_default_64": ((64, 64), 8, [8, None], [None, 4], 4),
    "groups_g1_32": ((32, 32), 8, [None, 8], [1, None], 2),
    "micro_16x32": ((16, 32), 8, [8, 16], [None, None], 2),
    "micro_32_df1": ((32, 32), 4, [4, 8], [None, None], 1),
}
SMALL_FULL = ("micro_16x32", "micro_32_df1")


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()


def geometry_cfg(window, stripe, groups, df):
    return dict(window_size=window, stripe_size=stripe, stripe_groups=groups, anchor_window_down_factor=df)


def ref_geometry(grl_mod, x_size, window, stripe, groups, df):
    """Runs the reference's GRL.set_table_index_mask without building a network."""
    from timm.models.layers import to_2tuple

    stub = type("Stub", (), {})()
    stub.stripe_size, stub.stripe_groups = stripe, groups
    stub.anchor_window_down_factor = df
    stub.window_size = to_2tuple(window)
    stub.shift_size = [w // 2 for w in stub.window_size]
    stub.pretrained_window_size = [0, 0]
    stub.pretrained_stripe_size = [0, 0]
    return grl_mod.GRL.set_table_index_mask(stub, x_size)


MODEL_CASES = {
    # name: (cfg, batch, (H, W), noise_sigma)
    "cfg1_tiny_x2_64": (configs.grl_config("tiny", "sr", 2, 64), 1, (64, 64), 0.0),
    "micro_cab_x2": (configs.micro_config(), 2, (32, 32), 0.0),
    "micro_pad_dn": (configs.micro_config(embed_dim=36, stripe=(8, 16), df=2, upsampler="", upscale=1, img_size=32),
                     1, (24, 40), 50.0),
    "micro_groups": (configs.micro_config(embed_dim=32, heads=2, window=4, stripe=(4, None), groups=(None, 2), df=2,
                                          local_connection=False, upsampler="pixelshuffledirect", upscale=3,
                                          img_size=16), 1, (16, 16), 0.0),
    "micro_odd_d": (configs.micro_config(embed_dim=60, heads=3, window=8, stripe=(16, 8), df=4, local_connection=True,
                                         upsampler="nearest+conv", upscale=4, img_size=32, depth=2, stages=2),
                    1, (32, 32), 0.0),
    "micro_gray": (configs.micro_config(embed_dim=32, heads=1, window=6, stripe=(6, 12), df=3, local_connection=False,
                                        upsampler="", upscale=1, img_size=24, in_channels=1), 1, (24, 24), 25.0),
}


def build_reference(grl_mod, cfg, sd):
    torch.manual_seed(0)
    m = grl_mod.GRL(**cfg).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.split("_")[0] in ("table", "index", "mask") for k in missing), missing
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    grl_mod, eff, mab, ops = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    # ---- 1. geometry: oracle == reference bit-exact; digests committed
    digests, small = {}, {}
    for name, (x_size, window, stripe, groups, df) in GEOMETRIES.items():
        ref = ref_geometry(grl_mod, x_size, window, stripe, groups, df)
        mine = orc.table_index_mask(geometry_cfg(window, stripe, groups, df), x_size)
        digests[name] = dict(x_size=list(x_size), window=window, stripe=stripe, groups=groups, df=df, sha256={})
        for k, v in ref.items():
            assert mine[k].dtype == v.dtype and mine[k].shape == v.shape, (name, k)
            assert torch.equal(mine[k], v), f"oracle geometry mismatch {name}:{k}"
            digests[name]["sha256"][k] = sha(v)
            digests[name].setdefault("shape", {})[k] = list(v.shape)
            if name in SMALL_FULL:
                small[f"{name}/{k}"] = v.numpy()
        print(f"[geometry] {name}: oracle == reference (bit-exact, {len(ref)} tensors)")
    # the reference's own __main__ known answers (ops.py:472-551 / SURVEY.md section 4)
    for ws, df, rows in (([4, 86], 1, 1197), ([4, 86], 2, 640), ([8, 8], 1, 225), ([8, 8], 2, 121)):
        t = ops.get_relative_coords_table_all(ws, [0, 0], df)
        assert t.shape[1] * t.shape[2] == rows, (ws, df, t.shape)
        assert torch.equal(orc.coords_table(ws, df), t)
        for w2a in (True, False):
            i_all = ops.get_relative_position_index_all(ws, df, w2a)
            i_simple = ops.get_relative_position_index_simple(ws, df, w2a)
            assert torch.equal(i_all, i_simple) and int(i_simple.max()) == rows - 1
            assert torch.equal(orc.position_index(ws, df, w2a), i_simple)
    print("[geometry] reference __main__ known answers reproduced (rows 1197/640/225/121)")

    # ---- 2. models: oracle vs reference, outputs committed
    out_files = {}
    for name, (cfg, batch, hw, sigma) in MODEL_CASES.items():
        sd = orc.synth_state_dict(cfg, seed=0)
        ref = build_reference(grl_mod, cfg, sd)
        x = orc.synth_input((batch, cfg["in_channels"], *hw), seed=1234, noise_sigma=sigma)
        with torch.no_grad():
            y_ref = ref(x.clone())
            y_orc = orc.grl_forward(sd, cfg, x.clone())
        err = (y_ref - y_orc).abs().max().item()
        print(f"[model] {name}: out {tuple(y_ref.shape)} |oracle-ref|max = {err:.3e}  (range {y_ref.min():.3f}..{y_ref.max():.3f})")
        assert err <= 2e-6 * max(1.0, y_ref.abs().max().item()), name
        out_files[name] = dict(output=y_ref.numpy())

    # per-module taps of block 2 (H, window shift + stripe shift) and block 3 (W, stripe shift) of micro_cab_x2
    cfg, batch, hw, _ = MODEL_CASES["micro_cab_x2"]
    sd = orc.synth_state_dict(cfg, seed=0)
    ref = build_reference(grl_mod, cfg, sd)
    g = torch.Generator().manual_seed(99)
    hw = (16, 32)  # non-square, 2x2 windows of 8x8, stripes 8x16 / 16x8
    xb = torch.randn(1, hw[0] * hw[1], cfg["embed_dim"], generator=g)
    tim_ref = ref.get_table_index_mask(None, tuple(hw))
    tim_orc = orc.table_index_mask(cfg, tuple(hw))
    for bi in (0, 1, 2, 3):
        blk = ref.layers[0].blocks[bi]
        taps = {}
        hooks = [
            blk.attn.qkv.register_forward_hook(lambda m, i, o: taps.__setitem__("qkv", o)),
            blk.attn.anchor.register_forward_hook(lambda m, i, o: taps.__setitem__("anchor", o)),
            blk.attn.window_attn.register_forward_hook(lambda m, i, o: taps.__setitem__("x_window", o)),
            blk.attn.stripe_attn.register_forward_hook(lambda m, i, o: taps.__setitem__("x_stripe", o)),
            blk.attn.register_forward_hook(lambda m, i, o: taps.__setitem__("attn_out", o)),
            blk.conv.register_forward_hook(lambda m, i, o: taps.__setitem__("cab", o)),
            blk.mlp.register_forward_hook(lambda m, i, o: taps.__setitem__("mlp", o)),
        ]
        with torch.no_grad():
            y = blk(xb, tuple(hw), tim_ref)
        for h in hooks:
            h.remove()
        mine = {}
        with torch.no_grad():
            y2 = orc.transformer_block(sd, f"layers.0.blocks.{bi}.", xb, tuple(hw), orc.block_settings(cfg, 0, bi),
                                       tim_orc, mine)
        mine["cab"] = orc.cab(sd, f"layers.0.blocks.{bi}.conv.", xb, tuple(hw))
        for k in ("qkv", "anchor", "x_window", "x_stripe", "attn_out", "cab"):
            e = (taps[k] - mine[k]).abs().max().item()
            assert e <= 5e-6, (bi, k, e)
        assert (y - y2).abs().max().item() <= 1e-5
        for k, v in taps.items():
            if k not in ("qkv", "mlp"):  # keep the fixture small
                out_files["micro_cab_x2"][f"block{bi}/{k}"] = v.numpy()
        out_files["micro_cab_x2"][f"block{bi}/out"] = y.numpy()
        print(f"[taps] block {bi}: qkv/anchor/window/stripe/proj/cab/out match (<=5e-6)")
    out_files["micro_cab_x2"]["block_input"] = xb.numpy()
    # stage-level
    with torch.no_grad():
        ys = ref.layers[0](xb, tuple(hw), tim_ref)
        ys2 = orc.transformer_stage(sd, cfg, 0, xb, tuple(hw), tim_orc)
    assert (ys - ys2).abs().max().item() <= 2e-5
    out_files["micro_cab_x2"]["stage0/out"] = ys.numpy()

    # metric known-answers from the reference's own psnr/tensor_round (imported by file path: utils/ has no deps we lack)
    import importlib.util

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ui = load("/root/reference/utils/utils_image.py", "ref_utils_image")
    g = torch.Generator().manual_seed(5)
    a = torch.rand(2, 3, 40, 40, generator=g) * 1.2 - 0.1
    b = torch.rand(2, 3, 40, 40, generator=g)
    ar, br = ui.tensor_round(a.clone()), ui.tensor_round(b.clone())
    sh = lambda t: ui.shave(t, 4)
    d = sh(ar) - sh(br)
    ref_psnr = -10 * d.pow(2).mean([-3, -2, -1]).log10()  # utils/metrics/psnr.py:44-48
    assert torch.allclose(orc.psnr(a, b, 4), ref_psnr, atol=0, rtol=0)
    out_files["micro_cab_x2"]["psnr/a"] = a.numpy()
    out_files["micro_cab_x2"]["psnr/b"] = b.numpy()
    out_files["micro_cab_x2"]["psnr/value_border4"] = ref_psnr.numpy()
    print("[metric] psnr/tensor_round match the reference")

    # parameter counts (paper table 5 / SURVEY.md section 4)
    counts = {}
    for v, task, s in (("tiny", "sr", 2), ("tiny", "sr", 4), ("small", "sr", 4), ("base", "sr", 4), ("base", "dn", 1)):
        n = sum(int(np.prod(sh_)) for sh_ in orc.param_shapes(configs.grl_config(v, task, s, 64 if task == "sr" else 128)).values())
        counts[f"{v}_{task}_x{s}"] = n
    print("[params]", counts)
    assert counts["tiny_sr_x2"] == 885420 or abs(counts["tiny_sr_x2"] / 1e6 - 0.885) < 1e-3
    assert abs(counts["base_sr_x4"] / 1e6 - 20.201) < 2e-3 and abs(counts["small_sr_x4"] / 1e6 - 3.488) < 1e-3

    if args.check:
        print("check OK (fixtures not rewritten)")
        return
    with open(os.path.join(GOLD, "geometry.json"), "w") as f:
        json.dump(dict(geometries=digests, param_counts=counts), f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(GOLD, "geometry_small.npz"), **small)
    for name, arrs in out_files.items():
        np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **arrs)
    with open(os.path.join(GOLD, "cases.json"), "w") as f:
        json.dump({k: dict(cfg=v[0], batch=v[1], hw=list(v[2]), sigma=v[3]) for k, v in MODEL_CASES.items()}, f, indent=1)
    print("fixtures written to", GOLD)


if __name__ == "__main__":
    main()
