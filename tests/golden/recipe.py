"""Deterministic weights for the golden-vector tests: every tensor of a state_dict is filled from a
generator seeded by the CRC32 of its (reference) name, so the container-side script that runs the
UNMODIFIED reference and the GPU-side test that runs pasco_b200.net3d agree on 130 M parameters
without committing them."""
import zlib

import torch


def _canonical(name: str) -> str:
    """The reference registers its one transformer under three names (Net.transformer_predictor,
    unet3d.transformer_predictor, unet3d.decoder_generative.transformer_predictor): seed them identically."""
    for alias in ("unet3d.decoder_generative.transformer_predictor.", "unet3d.transformer_predictor."):
        if name.startswith(alias):
            return "transformer_predictor." + name[len(alias):]
    return name


def fill_state_dict(sd):
    out = {}
    for name, t in sd.items():
        g = torch.Generator().manual_seed(zlib.crc32(_canonical(name).encode()))
        if not t.dtype.is_floating_point:
            out[name] = torch.zeros_like(t)
            continue
        shape = tuple(t.shape)
        if name.endswith("running_var"):
            v = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            v = 0.1 * torch.randn(shape, generator=g)
        elif t.ndim <= 1 or (t.ndim == 2 and shape[0] == 1 and name.endswith("bias")):
            v = 0.1 * torch.randn(shape, generator=g)
            if name.endswith("weight"):
                v = v + 1.0                       # norm-layer gains
        else:
            fan_in = t.numel() // shape[-1] if name.endswith("kernel") else t.numel() // shape[0]
            if "query_feat" in name or "query_embed" in name:
                fan_in = 1
            v = torch.randn(shape, generator=g) / (fan_in ** 0.5)
        out[name] = v.to(t.dtype)
    return out
