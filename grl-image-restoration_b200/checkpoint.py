"""Checkpoint ingestion as the reference's launcher does it (tools/trainer.py:93-115): take the Lightning
checkpoint's `state_dict`, drop the engine's three buffers, run GRL.convert_checkpoint (drops every table / index /
mask buffer, grl.py:556-569), unwrap a `params` sub-dict, merge into the current state dict and load strictly.

The reference loads into the ENGINE (whose attribute `model` is the network); this function loads into the network
itself, so when the checkpoint carries engine-level keys (`model.*` next to loss / EMA / metric states) only the
`model.*` entries are kept and the prefix is stripped -- the keys a strict load into the engine would have put on
`engine.model`."""
import torch

_ENGINE_BUFFERS = ("current_val_metric", "best_val_metric", "best_iter")


def load_reference_checkpoint(model, ckpt, strict=True):
    """`ckpt`: path to a .ckpt / .pth file, or an already loaded dict (Lightning checkpoint or bare state dict).
    Lightning checkpoints pickle omegaconf hyper-parameters next to the tensors, so files are read with
    weights_only=False exactly like the reference's torch.load call: only load checkpoints you trust."""
    if isinstance(ckpt, (str, bytes)):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd = ckpt
    if "state_dict" in sd:
        sd = dict(sd["state_dict"])
        for k in _ENGINE_BUFFERS:
            sd.pop(k, None)
    else:
        sd = dict(sd)
    sd = model.convert_checkpoint(sd)
    if "params" in sd:
        sd = dict(sd["params"])
    if any(k.startswith("model.") for k in sd):
        sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    current = model.state_dict()
    current.update(sd)
    return model.load_state_dict(current, strict=strict)
