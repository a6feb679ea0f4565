"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference model from /root/reference.

Used by oracle/make_golden*.py (build container, /root/reference) to pin the oracle restatement and to write
tests/golden/*.npz, and by bench.py's reference arm / cpu_baseline leg (GPU box: the copy staged under oracle/_ref by
oracle/make_ref.py).

The reference's model files need three absent third-party packages for a handful of helpers
(SURVEY.md section 8c / Appendix E); we register minimal stand-ins in sys.modules:
  timm.models.layers.{to_2tuple, trunc_normal_, DropPath}  (grl.py:28, mixed_attn_block_efficient.py:20)
  fairscale.nn.checkpoint_wrapper                          (grl.py:10)
  omegaconf.OmegaConf.create                               (grl.py:11,302)
"""
import collections.abc
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pick_root():
    """/root/reference in the build container; oracle/_ref (staged byte for byte by oracle/make_ref.py, git-ignored,
    shipped with the gpurun snapshot) on the GPU box."""
    for cand in (os.environ.get("GRL_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "networks", "grl.py")):
            return cand
    return os.environ.get("GRL_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _pick_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "networks", "grl.py"))


def _install_standins():
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        def to_2tuple(x):
            if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
                return tuple(x)
            return (x, x)

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        class DropPath(torch.nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                shape = (x.shape[0],) + (1,) * (x.ndim - 1)
                return x * x.new_empty(shape).bernoulli_(keep) / keep

        layers.to_2tuple = to_2tuple
        layers.trunc_normal_ = trunc_normal_
        layers.DropPath = DropPath
        timm.models = timm_models
        timm_models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = timm_models
        sys.modules["timm.models.layers"] = layers
    if "fairscale.nn" not in sys.modules:
        fairscale = types.ModuleType("fairscale")
        fnn = types.ModuleType("fairscale.nn")
        fnn.checkpoint_wrapper = lambda m, offload_to_cpu=False: m
        fairscale.nn = fnn
        sys.modules["fairscale"] = fairscale
        sys.modules["fairscale.nn"] = fnn
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def create(d):
                return types.SimpleNamespace(**d)

        oc.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = oc


def import_reference():
    """Returns the reference's `models` package modules: (grl, efficient, mab, ops)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    _install_standins()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import models.networks.grl as grl
    import models.common.mixed_attn_block_efficient as eff
    import models.common.mixed_attn_block as mab
    import models.common.ops as ops

    return grl, eff, mab, ops
