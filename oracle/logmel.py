"""CPU oracle for front end A (resample 2048 Hz -> 16 kHz, Whisper log-mel).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product path (gw_whisper_b200) never imports this module.

Restates, in numpy/scipy, what the reference computes on the CPU for every 1 s window:
  * Signal_vs_Noise/utils/preprocess.py:44-51  resample_timeseries -> scipy.signal.resample(x, 16000)
    (result stored as float32, preprocess.py:95)
  * Signal_vs_Noise/src/dataset.py:20-24       WhisperFeatureExtractor(audio, sampling_rate=16000)
    -> HF feature_extraction_whisper.py:104-133 (_np_extract_fbank_features, the only path in the
    reference's pinned transformers 4.37.2): zero-pad to 480000, reflect-padded STFT 400/160 with a
    periodic Hann window, |.|^2, 80 slaney mel filters, log10(max(.,1e-10)), clamp to max-8, (x+4)/4.

Parity pinning: the reference ships no tests/golden vectors for this path, so
`logmel_reference()` below *is* the reference implementation executed in this image (scipy + the
installed transformers), and `logmel_restated()` is our independent f64 restatement checked against
it in tests/test_oracle.py; golden vectors generated from it live in tests/golden/.
"""
from __future__ import annotations

import numpy as np

SR_IN = 2048
SR_OUT = 16000
N_FFT = 400
HOP = 160
N_MELS = 80
N_FRAMES = 3000
N_LIVE_FRAMES = 102  # frames 0..101 overlap the 16000 resampled samples; the rest see only zeros


def resample_reference(x: np.ndarray) -> np.ndarray:
    """scipy.signal.resample(x, 16000) along the last axis, cast to float32 (preprocess.py:44-51,95)."""
    from scipy.signal import resample

    n_out = x.shape[-1] * SR_OUT // SR_IN
    return resample(np.asarray(x, dtype=np.float64), n_out, axis=-1).astype(np.float32)


def logmel_reference(x: np.ndarray, path: str = "np") -> np.ndarray:
    """Reference front end: [..., 2048] strain -> [..., 80, 3000] float32 via the installed HF
    WhisperFeatureExtractor (default ctor == openai/whisper-{tiny,base,small} settings)."""
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor()
    y = resample_reference(x)
    flat = y.reshape(-1, y.shape[-1])
    if path == "np":
        padded = np.zeros((flat.shape[0], fe.n_samples), dtype=np.float32)
        padded[:, : flat.shape[1]] = flat
        out = fe._np_extract_fbank_features(padded, "cpu").astype(np.float32)
    else:  # the torch/f32 path transformers >= 4.4x takes when torch is importable
        out = np.stack(
            [fe(a, sampling_rate=SR_OUT, return_tensors="np").input_features[0] for a in flat]
        ).astype(np.float32)
    return out.reshape(*x.shape[:-1], N_MELS, N_FRAMES)


# --------------------------------------------------------------------------------------------
# Independent restatement (no scipy.signal / transformers calls).
# --------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    log_region = f >= 1000.0
    mels = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), mels)
    return mels


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    log_region = m >= 15.0
    return np.where(log_region, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filter_bank_slaney(n_freq: int = 201, n_mels: int = N_MELS, fmin=0.0, fmax=8000.0,
                           sr: int = SR_OUT) -> np.ndarray:
    """[n_freq, n_mels] float64 triangular filters, slaney scale + slaney area norm
    (HF audio_utils.mel_filter_bank(norm='slaney', mel_scale='slaney'), called at
    feature_extraction_whisper.py:94-102)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    fdiff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2 : n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def resample_restated(x: np.ndarray) -> np.ndarray:
    """Fourier-domain resample 2048 -> 16000 as scipy does it for real even-length input:
    X=rfft(x); Y[:1025]=X; Y[1024]*=0.5; y=irfft(Y,16000)*(16000/2048)."""
    x = np.asarray(x, dtype=np.float64)
    X = np.fft.rfft(x, axis=-1)
    Y = np.zeros(x.shape[:-1] + (SR_OUT // 2 + 1,), dtype=np.complex128)
    Y[..., : X.shape[-1]] = X
    Y[..., X.shape[-1] - 1] *= 0.5
    y = np.fft.irfft(Y, SR_OUT, axis=-1) * (SR_OUT / x.shape[-1])
    return y.astype(np.float32)


def logmel_restated(x: np.ndarray) -> np.ndarray:
    """[..., 2048] -> [..., 80, 3000] float32, f64 arithmetic on the float32 resampled audio."""
    y = resample_restated(x).astype(np.float64)
    lead = y.shape[:-1]
    y = y.reshape(-1, y.shape[-1])
    fb = mel_filter_bank_slaney()                      # [201, 80]
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)   # periodic Hann
    out = np.empty((y.shape[0], N_MELS, N_FRAMES), dtype=np.float32)
    for i, a in enumerate(y):
        # only the first 102 frames can be non-zero; build just that part of the padded signal
        need = (N_LIVE_FRAMES - 1) * HOP + N_FFT       # samples of the centre-padded signal
        sig = np.zeros(need, dtype=np.float64)
        sig[: N_FFT // 2] = a[N_FFT // 2 : 0 : -1]     # reflect pad: padded[-k] = a[k]
        ncopy = min(need - N_FFT // 2, a.shape[0])
        sig[N_FFT // 2 : N_FFT // 2 + ncopy] = a[:ncopy]
        idx = np.arange(N_LIVE_FRAMES)[:, None] * HOP + np.arange(N_FFT)[None, :]
        spec = np.fft.rfft(sig[idx] * win[None, :], axis=-1)            # [102, 201]
        power = spec.real ** 2 + spec.imag ** 2
        mel = np.maximum(power @ fb, 1e-10)                              # [102, 80]
        logm = np.full((N_FRAMES, N_MELS), -10.0)
        logm[:N_LIVE_FRAMES] = np.log10(mel)
        logm = np.maximum(logm, logm.max() - 8.0)
        out[i] = ((logm + 4.0) / 4.0).T.astype(np.float32)
    return out.reshape(*lead, N_MELS, N_FRAMES)


def feature_error(a: np.ndarray, b: np.ndarray) -> float:
    """Scale-normalised parity metric used for the 1e-4 front-end gate:
    max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
