"""Front end A on the GPU: resample 2048 Hz -> 16 kHz + Whisper log-mel, one fused kernel.

Mirrors `resample_timeseries` (Signal_vs_Noise/utils/preprocess.py:44-51) and
`WhisperFeatureExtractor(audio, sampling_rate=16000, return_tensors="pt").input_features`
(Signal_vs_Noise/src/dataset.py:20-24).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib


def logmel_features(strain: torch.Tensor) -> torch.Tensor:
    """[..., 2048] f32 CUDA strain at 2048 Hz -> [..., 80, 3000] f32 log-mel features."""
    if strain.shape[-1] != 2048:
        raise ValueError(f"expected 2048-sample windows, got {strain.shape[-1]}")
    if not strain.is_cuda:
        raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
    lib = _lib.load()
    s = strain.contiguous().float()
    lead = s.shape[:-1]
    n = s.numel() // 2048
    out = torch.empty((n, 80, 3000), dtype=torch.float32, device=s.device)
    _lib.check(lib.gww_logmel_frontend(s.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, 80, 3000)


def resample_timeseries(data: torch.Tensor) -> torch.Tensor:
    """Kept for signature parity with preprocess.py:44; the fused kernel resamples internally, so
    this returns the 2048 Hz window unchanged for `LogMelFeatureExtractor` to consume."""
    return data


class LogMelFeatureExtractor:
    """`fe(audio_2048hz, sampling_rate=16000, return_tensors="pt").input_features` -> [B,80,3000]."""

    def __call__(self, audio, sampling_rate: int = 16000, return_tensors: str = "pt", **_):
        if sampling_rate != 16000:
            raise ValueError("Whisper features are defined at 16 kHz (reference passes sampling_rate=16000)")
        a = torch.as_tensor(audio)
        if a.dim() == 1:
            a = a[None]
        return SimpleNamespace(input_features=logmel_features(a.cuda()))
