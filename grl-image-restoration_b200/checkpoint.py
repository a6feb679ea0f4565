"""Checkpoint ingestion exactly as the reference's launcher does it (tools/trainer.py:93-115): take the Lightning
checkpoint's `state_dict`, drop the engine's three buffers, strip the `model.` attribute prefix, run
GRL.convert_checkpoint (drops every table / index / mask buffer, grl.py:556-569), merge into the current state dict
and load strictly."""
import torch

_ENGINE_BUFFERS = ("current_val_metric", "best_val_metric", "best_iter")


def load_reference_checkpoint(model, ckpt, strict=True):
    """`ckpt`: path to a .ckpt / .pth file, or an already loaded dict (Lightning checkpoint or bare state dict)."""
    if isinstance(ckpt, (str, bytes)):
        ckpt = torch.load(ckpt, map_location="cpu")
    sd = dict(ckpt.get("state_dict", ckpt))
    for k in _ENGINE_BUFFERS:
        sd.pop(k, None)
    sd = model.convert_checkpoint(sd)
    sd = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}
    current = model.state_dict()
    current.update(sd)
    return model.load_state_dict(current, strict=strict)
