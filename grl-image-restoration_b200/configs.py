"""Effective ("released") GRL hyper-parameters: the model yamls overridden by the experiment yamls
and the evaluation commands.

Cites (reference): config/model/grl/grl_{tiny,small,base}.yaml, config/experiment/sr/grl/grl_p256.yaml:31-41,
config/experiment/dn/grl/grl_p256.yaml:34-45, config/experiment/db_motion/grl_p480.yaml:33-44,
scripts/grl/grl_test.md:35-128.
"""
import copy

_COMMON = dict(
    in_channels=3, img_range=1.0, stripe_shift=True, mlp_ratio=2, qkv_proj_type="linear",
    anchor_proj_type="avgpool", anchor_one_stage=True, out_proj_type="linear", conv_type="1conv",
    init_method="n", fairscale_checkpoint=False, offload_to_cpu=False, euclidean_dist=False,
)

_VARIANT = {
    "tiny": dict(embed_dim=64, depths=[4, 4, 4, 4], num_heads_window=[2] * 4, num_heads_stripe=[2] * 4,
                 local_connection=False, upsampler="pixelshuffledirect"),
    "small": dict(embed_dim=128, depths=[4, 4, 4, 4], num_heads_window=[2] * 4, num_heads_stripe=[2] * 4,
                  local_connection=False, upsampler="pixelshuffle"),
    "base": dict(embed_dim=180, depths=[4, 4, 8, 8, 8, 4, 4], num_heads_window=[3] * 7, num_heads_stripe=[3] * 7,
                 local_connection=True, upsampler="pixelshuffle"),
}


def grl_config(variant, task="sr", upscale=4, img_size=256, yaml_default=False):
    """Constructor kwargs for GRL(**cfg).  task in {sr, dn, deblur, jpeg, dm}."""
    cfg = dict(_COMMON)
    cfg.update(copy.deepcopy(_VARIANT[variant]))
    cfg["img_size"] = img_size
    if yaml_default:  # bare config/model/grl/*.yaml: window 8, stripe [8, W/4], df 4
        cfg.update(window_size=8, stripe_size=[8, None], stripe_groups=[None, 4], anchor_window_down_factor=4,
                   upscale=upscale)
        return cfg
    big = variant == "base"
    if task == "sr":
        cfg.update(window_size=32, stripe_size=[64, 64], stripe_groups=[None, None],
                   anchor_window_down_factor=2 if big else 4, upscale=upscale)
    elif task == "dn":
        cfg.update(window_size=32 if big else 16, stripe_size=[64, 128], stripe_groups=[None, None],
                   anchor_window_down_factor=2 if big else 4, upscale=1, upsampler="")
    elif task == "deblur":
        cfg.update(window_size=12, stripe_size=[48, 96], stripe_groups=[None, None],
                   anchor_window_down_factor=4, upscale=1, upsampler="")
    elif task == "jpeg":
        cfg.update(window_size=36, stripe_size=[72, 144], stripe_groups=[None, None],
                   anchor_window_down_factor=4, upscale=1, upsampler="")
    elif task == "dm":
        cfg.update(window_size=8, stripe_size=[32, 32], stripe_groups=[None, None],
                   anchor_window_down_factor=4, upscale=1, upsampler="")
    else:
        raise ValueError(task)
    return cfg


def micro_config(embed_dim=36, depth=4, stages=1, heads=2, window=8, stripe=(8, 16), groups=(None, None), df=2,
                 local_connection=True, upsampler="pixelshuffle", upscale=2, img_size=32, in_channels=3):
    """Small test architecture that still walks all four block personalities (i % 4)."""
    cfg = dict(_COMMON)
    cfg.update(embed_dim=embed_dim, depths=[depth] * stages, num_heads_window=[heads] * stages,
               num_heads_stripe=[heads] * stages, window_size=window, stripe_size=list(stripe),
               stripe_groups=list(groups), anchor_window_down_factor=df, local_connection=local_connection,
               upsampler=upsampler, upscale=upscale, img_size=img_size, in_channels=in_channels)
    return cfg


# BASELINE.json configs
BASELINE_CONFIGS = {
    "cfg1": dict(model=("tiny", "sr", 2, 64), batch=1, size=(64, 64)),
    "cfg2": dict(model=("small", "sr", 4, 256), batch=16, size=(256, 256)),
    "cfg3": dict(model=("base", "dn", 1, 256), batch=8, size=(256, 256)),
    "cfg4": dict(model=("base", "sr", 4, 256), batch=128, size=(256, 256)),
    "cfg5": dict(model=("base", "deblur", 1, 480), batch=1, size=(720, 1280)),
}
