"""The reference's model classes, same constructor/forward signatures and state_dict key names,
running on the B200 path when given a `B200WhisperEncoder`.

  two_channel_ligo_binary_classifier   Signal_vs_Noise/src/model.py:4-29
  one_channel_ligo_binary_classifier   Signal_vs_Noise/src/model.py:31-52 (also Efficiency_test/src/network.py:69-90)
  glitch_one_channel_classifier        Glitch_classification/src/model.py:4-39 (named
                                       one_channel_ligo_binary_classifier there; Dropout indices kept)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .encoder import B200WhisperEncoder


def _linears(seq: nn.Sequential):
    """(weight, bias) of the Linear layers of a Linear-ReLU-...-Linear head for the C ABI (which applies ReLU
    between consecutive entries).  Two Linear layers with NO activation between them -- the Efficiency_test tail
    Linear(64, 2) -> Linear(2, 2, bias=False), test_network.py:89-99 -- are composed into one (W2 W1, W2 b1 + b2)."""
    out = []
    prev_was_linear = False
    for m in seq:
        if isinstance(m, nn.Linear):
            w = m.weight.detach().float().cpu()
            b = m.bias.detach().float().cpu() if m.bias is not None else torch.zeros(m.out_features)
            if prev_was_linear:
                w0, b0 = out.pop()
                w, b = w @ w0, w @ b0 + b
            out.append((w, b))
            prev_was_linear = True
        elif isinstance(m, (nn.ReLU,)):
            prev_was_linear = False
        elif isinstance(m, (nn.Dropout, nn.Softmax, nn.Identity)):
            pass                                   # inert in eval / handled by the softmax flag
        else:
            raise TypeError(f"unsupported layer in a classifier head: {type(m).__name__}")
    return out


def replace_softmax_by_mutual_subtraction(network) -> None:
    """Efficiency_test's "unbounded softmax replacement" (Signal_vs_Noise/Efficiency_test/src/test_network.py:89-99):
    the trailing nn.Softmax of the 2-class head is replaced by a fixed bias-free Linear(2, 2) with weight
    [[1, -1], [-1, 1]], i.e. the outputs become (x0 - x1, x1 - x0).  Raises ValueError like the reference when the
    last layer is not a Softmax."""
    layers = list(network.classifier.children())
    if not isinstance(layers[-1], nn.Softmax):
        raise ValueError("The last layer of the classifier is not a Softmax layer.")
    new_layer = nn.Linear(2, 2, bias=False)
    new_layer.weight = nn.Parameter(torch.tensor([[1.0, -1.0], [-1.0, 1.0]]), requires_grad=False)
    layers[-1] = new_layer
    network.classifier = nn.Sequential(*layers)
    if hasattr(network, "refresh"):
        network.refresh()


class _B200Classifier(nn.Module):
    """Common plumbing: the classifier is a real nn.Sequential (state_dict compatible with the
    reference's .pth heads); its weights are mirrored into the C-ABI model handle lazily."""

    def __init__(self, encoder):
        super().__init__()
        if not isinstance(encoder, B200WhisperEncoder):
            raise TypeError("gw_whisper_b200 models need a B200WhisperEncoder (no PyTorch fallback path)")
        object.__setattr__(self, "encoder", encoder)  # not an nn.Module
        self._head_dirty = True

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._head_dirty = True
        return r

    def refresh(self):
        """Re-upload the classifier after editing its parameters in place."""
        self._head_dirty = True

    def _sync_head(self):
        if self._head_dirty or self.encoder._head_key != id(self):
            softmax = any(isinstance(m, nn.Softmax) for m in self.classifier)
            self.encoder.set_head(_linears(self.classifier), softmax=softmax)
            self.encoder._head_key = id(self)
            self._head_dirty = False

    def _from_feats(self, *mels):
        self._sync_head()
        reps = [self.encoder.pooled(m, use_last_token=True) for m in mels]
        return self.encoder.head(torch.cat(reps, dim=1))

    @torch.no_grad()
    def forward_strain(self, strain: torch.Tensor) -> torch.Tensor:
        """Fused path: whitened strain [B, D, 2048] -> logits [B, C] (front end + encoder + head)."""
        self._sync_head()
        return self.encoder.forward_windows_logmel(strain)


class two_channel_ligo_binary_classifier(_B200Classifier):
    def __init__(self, encoder, num_classes=1):
        super().__init__(encoder)
        d = encoder.config.d_model
        self.classifier = nn.Sequential(
            nn.Linear(d * 2, 1024), nn.ReLU(), nn.Linear(1024, 512), nn.ReLU(),
            nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, num_classes))

    @torch.no_grad()
    def forward(self, mel_tensor_0, mel_tensor_1):
        return self._from_feats(mel_tensor_0, mel_tensor_1)


class one_channel_ligo_binary_classifier(_B200Classifier):
    def __init__(self, encoder, num_classes=1, softmax=False):
        super().__init__(encoder)
        d = encoder.config.d_model
        layers = [nn.Linear(d, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 128),
                  nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, num_classes)]
        if softmax:  # Efficiency_test/src/network.py:73-85
            layers.append(nn.Softmax(dim=1))
        self.classifier = nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, mel_tensor_0):
        return self._from_feats(mel_tensor_0)


class glitch_one_channel_classifier(_B200Classifier):
    def __init__(self, encoder, num_classes=10):
        super().__init__(encoder)
        d = encoder.config.d_model
        self.classifier = nn.Sequential(
            nn.Linear(d, 512), nn.ReLU(), nn.Dropout(0.3), nn.Linear(512, 256), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.3), nn.Linear(128, num_classes))

    @torch.no_grad()
    def forward(self, mel_tensor):
        return self._from_feats(mel_tensor)
