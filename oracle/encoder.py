"""CPU oracle for the encoder + DoRA + heads (fp32 PyTorch).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

* The encoder arithmetic lives in a third-party dependency of the reference: `transformers`
  (pinned 4.37.2 in /root/reference/requirements.txt:296; 5.5.0 is installed in this image, same
  maths: pre-LN blocks, erf-GELU, q scaled by head_dim^-0.5 after bias, k_proj without bias;
  modeling_whisper.py:593-648).  `make_encoder()` instantiates that class directly, so the oracle
  *is* the reference's encoder code with random-init weights (no checkpoint download offline).
* DoRA lives in `peft==0.12.0` (requirements.txt:177), absent from this image.  `DoraLinear`
  restates `peft/tuners/lora/dora.py::DoraLinearLayer.forward` (eval mode, dropout off):
      y = base(x) + (m/||W0 + s B A||_row - 1) * (x W0^T) + (m/||W0 + s B A||_row) * s * (x A^T B^T)
  Parity unpinned for DoRA (no peft here, reference has no tests): anchored instead on the shipped
  adapter tensors via merged == unmerged (tests/test_oracle.py).
* Heads restate the nn.Sequential stacks of Signal_vs_Noise/src/model.py:9-20,35-47,
  Glitch_classification/src/model.py:10-21, MLGWSC-1/inference.py:371-382 and are checked against
  the reference classes imported from /root/reference when that tree is present.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from gw_whisper_b200.synthetic import SIZES, make_encoder, seeded_head, synthetic_dora  # noqa: F401  (weight factories)



class DoraLinear(nn.Module):
    """Unmerged DoRA forward of one adapted nn.Linear (PEFT 0.12 semantics, eval mode)."""

    def __init__(self, base: nn.Linear, A: torch.Tensor, B: torch.Tensor, m: torch.Tensor, scale: float):
        super().__init__()
        self.base, self.scale = base, float(scale)
        self.register_buffer("A", A.float().clone())
        self.register_buffer("B", B.float().clone())
        self.register_buffer("m", m.float().clone())

    def forward(self, x):
        W0 = self.base.weight
        lora_w = self.B @ self.A
        wnorm = torch.linalg.norm(W0 + self.scale * lora_w, dim=1)
        mag = (self.m / wnorm).view(1, -1)
        base_out = self.base(x)
        xw = torch.nn.functional.linear(x, W0)            # without bias
        lora_out = torch.nn.functional.linear(torch.nn.functional.linear(x, self.A), self.B)
        return base_out + (mag - 1) * xw + mag * self.scale * lora_out


def merged_dora_weight(W0, A, B, m, scale):
    """W' = diag(m / ||W0 + s B A||_row) (W0 + s B A)   (SURVEY.md section 8a, row E2)."""
    V = W0.float() + float(scale) * (B.float() @ A.float())
    return (m.float() / torch.linalg.norm(V, dim=1)).unsqueeze(1) * V


def attach_dora(encoder: nn.Module, dora: Dict[str, object]) -> nn.Module:
    """Wrap the adapted projections of an HF WhisperEncoder with unmerged DoRA (what
    PeftModel.from_pretrained does at MLGWSC-1/inference.py:411)."""
    scale = float(dora["lora_alpha"]) / float(dora["r"])
    t = dora["tensors"]
    for i, layer in enumerate(encoder.layers):
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            base = f"base_model.model.layers.{i}.self_attn.{proj}."
            if base + "lora_A.weight" not in t:
                continue
            mkey = base + "lora_magnitude_vector"
            m = t[mkey] if mkey in t else t[mkey + ".weight"]
            lin = getattr(layer.self_attn, proj)
            setattr(layer.self_attn, proj,
                    DoraLinear(lin, torch.as_tensor(t[base + "lora_A.weight"]),
                               torch.as_tensor(t[base + "lora_B.weight"]), torch.as_tensor(m), scale))
    return encoder


# ------------------------------------------------------------------------------------------------
# heads / model classes (restated)
# ------------------------------------------------------------------------------------------------
def head_two_channel(d: int, num_classes: int = 1) -> nn.Sequential:       # model.py:9-20
    return nn.Sequential(nn.Linear(d * 2, 1024), nn.ReLU(), nn.Linear(1024, 512), nn.ReLU(),
                         nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, num_classes))


def head_one_channel(d: int, num_classes: int = 1, softmax: bool = False) -> nn.Sequential:  # model.py:35-47
    layers = [nn.Linear(d, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 128),
              nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, num_classes)]
    if softmax:
        layers.append(nn.Softmax(dim=1))
    return nn.Sequential(*layers)


def head_glitch(d: int, num_classes: int = 10) -> nn.Sequential:           # Glitch model.py:10-21
    return nn.Sequential(nn.Linear(d, 512), nn.ReLU(), nn.Dropout(0.3), nn.Linear(512, 256), nn.ReLU(),
                         nn.Dropout(0.3), nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.3),
                         nn.Linear(128, num_classes))


def head_mlgwsc(d: int, n_detectors: int = 2, num_classes: int = 2, softmax: bool = True) -> nn.Sequential:
    layers = [nn.Linear(d * n_detectors, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
              nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, num_classes)]
    if softmax:                                                             # inference.py:371-382
        layers.append(nn.Softmax(dim=1))
    return nn.Sequential(*layers)


class TwoChannelOracle(nn.Module):
    """two_channel_ligo_binary_classifier.forward (Signal_vs_Noise/src/model.py:22-29)."""

    def __init__(self, encoder, num_classes=1):
        super().__init__()
        self.encoder = encoder
        self.classifier = head_two_channel(encoder.config.d_model, num_classes)

    def forward(self, mel0, mel1):
        a = self.encoder(mel0).last_hidden_state[:, -1, :]
        b = self.encoder(mel1).last_hidden_state[:, -1, :]
        return self.classifier(torch.cat((a, b), dim=1))


class OneChannelOracle(nn.Module):
    def __init__(self, encoder, num_classes=1, head=None):
        super().__init__()
        self.encoder = encoder
        self.classifier = head if head is not None else head_one_channel(encoder.config.d_model, num_classes)

    def forward(self, mel):
        return self.classifier(self.encoder(mel).last_hidden_state[:, -1, :])


