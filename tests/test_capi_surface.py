"""C-ABI library + nn.Module surface checks that need no GPU."""
import ctypes
import os

import pytest
import torch


def test_library_exports_every_header_symbol(pkg):
    from grl_image_restoration_b200 import capi

    names = capi.header_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/grl_b200.h but not exported"
    assert set(names) == set(capi._SIGNATURES), "ctypes signatures out of sync with the header"
    assert capi.lib().grl_abi_version() == capi.ABI_VERSION


def test_library_is_sm100a_native(pkg):
    import shutil
    import subprocess
    from grl_image_restoration_b200 import capi

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(pkg):
    m = pkg.GRL(**pkg.configs.micro_config())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        m.layers[0].blocks[0].mlp(torch.rand(1, 16, 36))


def test_error_reporting(pkg):
    from grl_image_restoration_b200 import capi

    rc = capi.lib().grl_rel_index_host(0, 4, 1, 1, None)
    assert rc == -1
    assert b"rel_index" in capi.lib().grl_last_error()
    with pytest.raises(RuntimeError, match="rel_index"):
        capi.check(rc)


@pytest.mark.parametrize("v,t,s,sz", [("tiny", "sr", 2, 64), ("small", "sr", 4, 64), ("base", "sr", 4, 64),
                                      ("base", "dn", 1, 128), ("base", "deblur", 1, 96)])
def test_parameter_names_and_shapes(pkg, oracle, v, t, s, sz):
    cfg = pkg.configs.grl_config(v, t, s, sz)
    m = pkg.GRL(**cfg)
    mine = {k: tuple(p.shape) for k, p in m.state_dict().items() if not k.startswith("table_")}
    assert mine == oracle.param_shapes(cfg)
    assert {k for k in m.state_dict() if k.startswith("table_")} == {"table_w", "table_sh", "table_sv"}


def test_state_dict_roundtrip_and_reference_buffers(pkg, oracle):
    cfg = pkg.configs.micro_config()
    m = pkg.GRL(**cfg)
    sd = oracle.synth_state_dict(cfg)
    full = dict(sd)
    full.update(oracle.table_index_mask(cfg, (32, 32)))  # a reference state_dict also has index_*/mask_* buffers
    m.load_state_dict(full, strict=True)
    for k, v in sd.items():
        assert torch.equal(m.state_dict()[k], v)
    # tools/trainer.py:93-115 flow: keys prefixed "model.", convert_checkpoint drops buffers
    ck = {"model." + k: v for k, v in full.items()}
    ck = m.convert_checkpoint(ck)
    assert not any(k.startswith("model.index_") or k.startswith("model.mask_") or k.startswith("model.table_") for k in ck)
    cur = m.state_dict()
    cur.update({k[len("model."):]: v for k, v in ck.items()})
    m.load_state_dict(cur, strict=True)


def test_constructor_contract(pkg):
    # yaml-only keys are swallowed by **kwargs (grl.py:255); fairscale flags are accepted no-ops
    cfg = pkg.configs.micro_config()
    m = pkg.GRL(name="grl_base", double_window=False, stripe_square=False, separable_conv_act=True, use_buffer=True,
                use_efficient_buffer=True, fairscale_checkpoint=True, offload_to_cpu=True, **{k: v for k, v in cfg.items()
                                                                                             if k not in ("fairscale_checkpoint", "offload_to_cpu")})
    assert m.pad_size == 16 and len(m.layers[0].blocks) == 4
    b = m.layers[0].blocks
    assert [x.stripe_type for x in b] == ["H", "W", "H", "W"]
    assert [x.window_shift for x in b] == [True, False, True, False]  # shift on EVEN blocks (grl.py:112)
    assert [x.stripe_shift for x in b] == [False, False, True, True]
    assert b[1].stripe_size == [16, 8]
    with pytest.raises(RuntimeError):  # img_size must be a multiple of the stripe size
        pkg.GRL(**dict(cfg, img_size=36))
    t = m.get_table_index_mask(None, (32, 32))
    assert t["table_w"] is m.table_w and t["mask_w"] is not None and t["mask_w"].numel() == 0


def test_tables_match_oracle(pkg, oracle):
    cfg = pkg.configs.grl_config("base", "dn", 1, 128)
    m = pkg.GRL(**cfg)
    tim = oracle.table_index_mask(cfg, (128, 128))
    for k in ("table_w", "table_sh", "table_sv"):
        assert torch.equal(getattr(m, k), tim[k])


def test_reference_checkpoint_ingestion(pkg, oracle, tmp_path):
    """tools/trainer.py:93-115: Lightning checkpoint with `model.`-prefixed keys, engine buffers and the reference's
    table / index / mask buffers -> strict load."""
    from grl_image_restoration_b200 import checkpoint

    cfg = pkg.configs.micro_config()
    sd = oracle.synth_state_dict(cfg, seed=4)
    full = {"model." + k: v for k, v in sd.items()}
    full.update({"model." + k: v for k, v in oracle.table_index_mask(cfg, (32, 32)).items()})
    full.update(current_val_metric=torch.tensor(0.0), best_val_metric=torch.tensor(31.2), best_iter=torch.tensor(5))
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": full, "epoch": 3}, path)
    m = pkg.GRL(**cfg)
    res = checkpoint.load_reference_checkpoint(m, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in sd.items():
        assert torch.equal(m.state_dict()[k], v)
