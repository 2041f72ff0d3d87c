"""Front end A on the GPU: resample 2048 Hz -> 16 kHz + Whisper log-mel.

Mirrors, with the same names and argument meaning,
  * `resample_timeseries(data)`                     Signal_vs_Noise/utils/preprocess.py:44-51 (copies at
    Real_events/preprocess_real_events.py:19-23, Efficiency_test/src/test_network.py:42-46,
    Glitch_classification/utils/preprocess_data.py:10-17): scipy.signal.resample(x, len*16000//2048)
  * `WhisperFeatureExtractor(audio, sampling_rate=16000, return_tensors="pt").input_features`
    as the datasets call it per item (Signal_vs_Noise/src/dataset.py:20-24,40-44,
    Glitch_classification/src/dataset.py:46-47)
  * `two_channel_LigoBinaryData` / `one_channel_LigoBinaryData` (Signal_vs_Noise/src/dataset.py:8-48)
and adds the fused form `logmel_features(strain_2048)` the sliding-window path uses (one kernel, the
16 kHz audio never leaves shared memory).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import _lib

WINDOW = 2048
AUDIO = 16000


def _as_cuda_f32(x, what: str) -> torch.Tensor:
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError(f"gw_whisper_b200 has no CPU path: {what} needs a CUDA device")
        t = t.cuda()
    return t.contiguous().float()


def logmel_features(strain: torch.Tensor) -> torch.Tensor:
    """[..., 2048] f32 CUDA strain at 2048 Hz -> [..., 80, 3000] f32 log-mel features (fused kernel)."""
    if strain.shape[-1] != WINDOW:
        raise ValueError(f"expected 2048-sample windows, got {strain.shape[-1]}")
    if not strain.is_cuda:
        raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
    lib = _lib.load()
    s = strain.contiguous().float()
    lead = s.shape[:-1]
    n = s.numel() // WINDOW
    out = torch.empty((n, 80, 3000), dtype=torch.float32, device=s.device)
    _lib.check(lib.gww_logmel_frontend(s.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, 80, 3000)


def resample_timeseries(data):
    """scipy.signal.resample(data, len(data) * 16000 // 2048) for 1 s windows (preprocess.py:44-51), on the
    GPU: [..., 2048] -> [..., 16000] f32 (the reference stores the result as float32, preprocess.py:95).
    numpy in -> numpy out, tensor in -> CUDA tensor out."""
    is_np = not isinstance(data, torch.Tensor)
    s = _as_cuda_f32(data, "resample_timeseries")
    if s.shape[-1] != WINDOW:
        raise ValueError(f"resample_timeseries: the B200 path handles 1 s windows of 2048 samples, got {s.shape[-1]}")
    lead = s.shape[:-1]
    n = s.numel() // WINDOW
    out = torch.empty((n, AUDIO), dtype=torch.float32, device=s.device)
    _lib.check(_lib.load().gww_resample_16k(s.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    out = out.reshape(*lead, AUDIO)
    return out.cpu().numpy() if is_np else out


def logmel_from_16k(audio: torch.Tensor) -> torch.Tensor:
    """[..., 16000] f32 CUDA audio at 16 kHz (a resampled 1 s window) -> [..., 80, 3000] f32."""
    if audio.shape[-1] != AUDIO:
        raise ValueError(f"expected 16000-sample audio (1 s at 16 kHz), got {audio.shape[-1]}")
    if not audio.is_cuda:
        raise RuntimeError("gw_whisper_b200 has no CPU path: audio must be a CUDA tensor")
    a = audio.contiguous().float()
    lead = a.shape[:-1]
    n = a.numel() // AUDIO
    out = torch.empty((n, 80, 3000), dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().gww_logmel_from_16k(a.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, 80, 3000)


class LogMelFeatureExtractor:
    """`WhisperFeatureExtractor` stand-in for the calls the reference makes:
    `fe(audio_16k, sampling_rate=16000, return_tensors="pt").input_features -> [B, 80, 3000]` with the
    reference's 16 000-sample (1 s) resampled audio.  A 2048-sample window is also accepted (the fused
    resample + log-mel kernel), which is what a caller holding raw 2048 Hz strain should pass."""

    sampling_rate = 16000
    n_samples = 480000
    feature_size = 80

    @classmethod
    def from_pretrained(cls, *_, **__):     # WhisperFeatureExtractor.from_pretrained(f"openai/whisper-{size}")
        return cls()                         # tiny/base/small share one extractor configuration

    def __call__(self, audio, sampling_rate: int = 16000, return_tensors: str = "pt", **_):
        if sampling_rate != 16000:
            # same condition and exception type as HF feature_extraction_whisper.py:239-245
            raise ValueError(
                f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a "
                f"sampling rate of 16000. Please make sure that the provided `raw_speech` input was sampled with "
                f"16000 and not {sampling_rate}.")
        a = _as_cuda_f32(audio, "LogMelFeatureExtractor")
        if a.dim() == 1:
            a = a[None]
        if a.shape[-1] == AUDIO:
            feats = logmel_from_16k(a)
        elif a.shape[-1] == WINDOW:
            feats = logmel_features(a)
        else:
            raise ValueError(f"the B200 front end handles 1 s windows: 16000 samples at 16 kHz (or 2048 at 2048 Hz), "
                             f"got {a.shape[-1]}")
        if return_tensors == "np":
            feats = feats.cpu().numpy()
        return SimpleNamespace(input_features=feats)


class two_channel_LigoBinaryData(torch.utils.data.Dataset):
    """Signal_vs_Noise/src/dataset.py:8-26 with the feature extraction on the GPU."""

    def __init__(self, ds, device, encoder):
        self.ds = ds
        self.device = device
        self.feature_extractor = LogMelFeatureExtractor.from_pretrained(f"openai/whisper-{encoder}")

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        it = self.ds[idx]
        h1 = self.feature_extractor(it["h1_timeseries"], sampling_rate=16000, return_tensors="pt").input_features.squeeze(0)
        l1 = self.feature_extractor(it["l1_timeseries"], sampling_rate=16000, return_tensors="pt").input_features.squeeze(0)
        return h1, l1, it["labels"], it["injection_snr"]


class one_channel_LigoBinaryData(torch.utils.data.Dataset):
    """Signal_vs_Noise/src/dataset.py:28-48."""

    def __init__(self, ds, device, encoder):
        self.ds = ds
        self.device = device
        self.feature_extractor = LogMelFeatureExtractor.from_pretrained(f"openai/whisper-{encoder}")

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        it = self.ds[idx]
        l1 = self.feature_extractor(it["l1_timeseries"], sampling_rate=16000, return_tensors="pt").input_features.squeeze(0)
        return l1, torch.tensor(it["labels"]).float(), it["injection_snr"]
