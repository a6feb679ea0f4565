"""Tiled inference with overlap averaging -- the engine's `forward_tile` (engines/base.py:90-116) with all tiles of the
image(s) batched into a few forwards instead of a Python double loop of single-tile forwards.

Semantics are the reference's exactly: tile = min(tile, h, w); origins range(0, h - tile, stride) + [h - tile] with
stride = tile - overlap (same for w); every tile is restored independently (so per-tile operators such as the CAB
global pool see the same data as in the reference), outputs are summed into E, a ones mask into W, result E / W.
"""
import torch


def tile_origins(size, tile, overlap):
    stride = tile - overlap
    return list(range(0, size - tile, stride)) + [size - tile]


@torch.no_grad()
def forward_tile(model, x, tile, tile_overlap, scale=None, max_batch=16):
    """x (B, C, H, W) on the GPU -> (B, C_out, H*scale, W*scale)."""
    b, _, h, w = x.shape
    scale = model.upscale if scale is None else scale
    tile = min(tile, h, w)
    hs, ws = tile_origins(h, tile, tile_overlap), tile_origins(w, tile, tile_overlap)
    origins = [(bi, hi, wi) for bi in range(b) for hi in hs for wi in ws]
    E = W = None
    for i in range(0, len(origins), max_batch):
        chunk = origins[i:i + max_batch]
        patches = torch.stack([x[bi, :, hi:hi + tile, wi:wi + tile] for bi, hi, wi in chunk])
        out = model(patches)
        if E is None:
            E = torch.zeros(b, out.shape[1], h * scale, w * scale, device=x.device, dtype=out.dtype)
            W = torch.zeros_like(E)
        for (bi, hi, wi), o in zip(chunk, out):
            E[bi, :, hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale].add_(o)
            W[bi, :, hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale].add_(1.0)
    return E.div_(W)
