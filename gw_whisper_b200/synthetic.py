"""Synthetic inputs and random-init weights for tests and benchmarks (no checkpoints are
downloadable offline): the named Whisper geometries instantiated through `transformers`' own
WhisperEncoder class, seeded DoRA adapters with the shipped geometry (r=8, alpha=32, use_dora), and
seeded classifier weights.  These are *inputs*; no algorithm of the hot path lives here."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

SIZES = {
    "tiny": dict(d_model=384, encoder_layers=4, encoder_attention_heads=6, encoder_ffn_dim=1536),
    "base": dict(d_model=512, encoder_layers=6, encoder_attention_heads=8, encoder_ffn_dim=2048),
    "small": dict(d_model=768, encoder_layers=12, encoder_attention_heads=12, encoder_ffn_dim=3072),
}


# Multipliers (q/k projections, other layer matrices, conv stem) of the "spread" weight set per size
# (default init gives logits with a 1e-4 spread across windows, which makes parity vacuous).
# tiny/base: probed in round 1 -- lifts the std of the last-token representation across Gaussian-noise
# windows from 1.2e-4 to ~4e-2 while PyTorch's own bf16 autocast of the same model still agrees with fp32 to
# ~1e-2 (x20 on q,k gives a 0.4 spread but is chaotic: torch bf16 autocast itself is then off by O(1)).
# base / small (round 2, tools/condition_search.py on a B200, 48 windows, profiles/r2_condition_search.txt): with
# (6, 3, 3) the twelve layers of whisper-small are a chaotic map (torch's own bf16 autocast is off by ~1.0 on the
# logits, fp16 autocast by 8e-2); the sets below keep the logit spread across windows >= 0.12 while fp16-operand
# arithmetic stays within 1e-2 of fp32 (spread / max error ~ 17-18; bf16 operands: ~2).
CONDITIONED = {"tiny": (6.0, 3.0, 3.0), "base": (6.0, 1.0, 3.0), "small": (4.0, 1.0, 3.0)}


def scale_encoder_(enc: nn.Module, qk: float, layer: float, conv: float) -> nn.Module:
    """In-place rescale of a random-init HF WhisperEncoder: q/k projection weights by `qk`, every other 2-D
    layer matrix by `layer`, the conv-stem weights by `conv`."""
    with torch.no_grad():
        for name, p in enc.named_parameters():
            if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):
                p.mul_(qk)
            elif name.startswith("layers.") and name.endswith("weight") and p.dim() == 2:
                p.mul_(layer)
            elif name.startswith("conv") and name.endswith("weight"):
                p.mul_(conv)
    return enc


def make_encoder(size: str = "tiny", seed: int = 0, init_std: Optional[float] = None,
                 spread: bool = False) -> nn.Module:
    """Random-init HF WhisperEncoder (fp32, eval).  `spread=True` rescales the projection / MLP
    weights and LayerNorm affine so the last-token representation varies O(1) across inputs
    (default init gives logits with a 1e-4 spread, which makes parity vacuous)."""
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder

    kw = dict(SIZES[size])
    if init_std is not None:
        kw["init_std"] = init_std
    cfg = WhisperConfig(decoder_layers=1, decoder_attention_heads=kw["encoder_attention_heads"],
                        decoder_ffn_dim=64, **kw)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    enc = WhisperEncoder(cfg).float().eval()
    if spread:
        scale_encoder_(enc, *CONDITIONED[size])
    return enc



def synthetic_dora(size: str, seed: int = 7, r: int = 8, alpha: float = 32.0,
                   targets=("k_proj", "v_proj")) -> Dict[str, object]:
    """Seeded adapter with the shipped geometry (r=8, alpha=32, use_dora) for any Whisper size:
    A ~ kaiming-uniform-ish, small random B, m = ||W0||_row * U(0.8, 1.2) is emulated with U(0.3,1.1)
    (shipped lora_magnitude_vector values span 0.09..1.12)."""
    d = SIZES[size]["d_model"]
    L = SIZES[size]["encoder_layers"]
    g = torch.Generator().manual_seed(seed)
    t = {}
    for i in range(L):
        for proj in targets:
            base = f"base_model.model.layers.{i}.self_attn.{proj}."
            t[base + "lora_A.weight"] = ((torch.rand(r, d, generator=g) * 2 - 1) / d ** 0.5).numpy()
            t[base + "lora_B.weight"] = (0.02 * torch.randn(d, r, generator=g)).numpy()
            t[base + "lora_magnitude_vector"] = (0.3 + 0.8 * torch.rand(d, generator=g)).numpy()
    return {"tensors": t, "r": r, "lora_alpha": alpha, "use_dora": True}



def seeded_head(head: nn.Sequential, seed: int = 3, gain: float = 1.0) -> nn.Sequential:
    """Deterministic head weights (default nn.Linear init under a fixed seed, optional gain so the
    logits have O(1) spread with random-init encoders)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in head:
            if isinstance(m, nn.Linear):
                bound = gain / m.in_features ** 0.5
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
    return head.eval()
