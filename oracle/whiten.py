"""CPU oracle for whitening: the reference's `whiten`
(MLGWSC-1/inference.py:56-137) with the pycbc calls it makes restated in numpy float64.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

PARITY UNPINNED: the arithmetic lives in `pycbc==2.4.0` (requirements.txt:204), which is not installed in this
image, and the reference ships no tests or golden vectors.  Restated from the published pycbc 2.4.0 sources:

  TimeSeries.psd(segment_duration)      pycbc/types/timeseries.py   seg_len = round(dur*fs), stride = seg_len//2
  pycbc.psd.welch(window='hann', avg_method='median')   pycbc/psd/estimate.py
        segments trimmed symmetrically to a whole number of strides, numpy.hanning window,
        |fft(seg*w)*delta_t|^2 with DC and Nyquist halved, median / median_bias(n), * 2*delta_f*seg_len/sum(w^2)
  pycbc.psd.interpolate(psd, delta_f)   pycbc/psd/estimate.py       numpy.interp onto k*delta_f, k <= N/2
  pycbc.psd.inverse_spectrum_truncation(psd, max_filter_len, low_frequency_cutoff, trunc_method='hann')
        inv_asd[kmin:N/2] = psd^-1/2, q = irfft, q[0:L/2] *= hanning(L)[-L/2:], q[N-L/2:] *= hanning(L)[:L/2],
        q[L/2:N-L/2] = 0, psd_out = 1 / |rfft(q)|^2
  white = irfft(rfft(x) * psd_out^-1/2)[L/2 : N-L/2]                inference.py:94-98
(the delta_t / delta_f factors pycbc's fft/ifft wrappers apply cancel pairwise; they are kept out).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def median_bias(n: int) -> float:
    """pycbc.psd.estimate.median_bias"""
    if n >= 1000:
        return float(np.log(2))
    ans = 1.0
    for i in range(1, int((n - 1) / 2 + 1)):
        ans += 1.0 / (2 * i + 1) - 1.0 / (2 * i)
    return ans


def welch_psd(x: np.ndarray, delta_t: float, seg_len: int, seg_stride: int) -> np.ndarray:
    """pycbc.psd.welch(ts, seg_len, seg_stride, window='hann', avg_method='median') -> psd [seg_len/2+1]."""
    x = np.asarray(x, dtype=np.float64)
    num_samples = len(x)
    num_segments = int(num_samples // seg_stride)
    if (num_segments - 1) * seg_stride + seg_len > num_samples:
        num_segments -= 1
    if num_segments < 1:
        raise ValueError("time series too short for one Welch segment")
    data_len = (num_segments - 1) * seg_stride + seg_len
    if data_len < num_samples:
        diff = num_samples - data_len
        start = diff // 2
        end = num_samples - diff // 2
        if diff % 2:
            start = start + 1
        x = x[start:end]
        num_samples = len(x)
    assert num_samples == data_len
    w = np.hanning(seg_len).astype(np.float64)
    delta_f = 1.0 / delta_t / seg_len
    idx = np.arange(num_segments)[:, None] * seg_stride + np.arange(seg_len)[None, :]
    tilde = np.fft.rfft(x[idx] * w[None, :], axis=1) * delta_t
    seg_psd = np.abs(tilde * tilde.conj())
    seg_psd[:, 0] /= 2
    seg_psd[:, -1] /= 2
    psd = np.median(seg_psd, axis=0) / median_bias(num_segments)
    psd *= 2 * delta_f * seg_len / (w * w).sum()
    return psd


def interpolate_psd(psd: np.ndarray, psd_delta_f: float, delta_f: float) -> np.ndarray:
    """pycbc.psd.interpolate"""
    new_n = (len(psd) - 1) * psd_delta_f / delta_f + 1
    samples = np.arange(0, np.rint(new_n)) * delta_f
    return np.interp(samples, np.arange(len(psd)) * psd_delta_f, psd)


def inverse_spectrum_truncation(psd: np.ndarray, delta_f: float, max_filter_len: int,
                                low_frequency_cutoff: Optional[float] = None,
                                trunc_method: Optional[str] = None) -> np.ndarray:
    """pycbc.psd.inverse_spectrum_truncation -> truncated PSD (same length)."""
    N = (len(psd) - 1) * 2
    inv_asd = np.zeros(len(psd), dtype=np.complex128)
    kmin = 1
    if low_frequency_cutoff:
        kmin = int(low_frequency_cutoff / delta_f)
    inv_asd[kmin:N // 2] = (1.0 / psd[kmin:N // 2]) ** 0.5
    q = np.fft.irfft(inv_asd, N)
    trunc_start = max_filter_len // 2
    trunc_end = N - max_filter_len // 2
    if trunc_end < trunc_start:
        raise ValueError("Invalid value in inverse_spectrum_truncation")
    if trunc_method == "hann":
        tw = np.hanning(max_filter_len)
        q[0:trunc_start] *= tw[-trunc_start:]
        q[trunc_end:N] *= tw[0:max_filter_len // 2]
    if trunc_start < trunc_end:
        q[trunc_start:trunc_end] = 0
    pt = np.fft.rfft(q)
    pt = pt * pt.conj()
    return 1.0 / np.abs(pt)


def whiten(strain: np.ndarray, delta_t: float = 1.0 / 2048.0, segment_duration: float = 0.5,
           max_filter_duration: float = 0.25, trunc_method: Optional[str] = "hann",
           remove_corrupted: bool = True, low_frequency_cutoff: Optional[float] = None,
           return_psd: bool = False, return_filter: bool = False):
    """MLGWSC-1/inference.py:56-137 for a 1-D or 2-D strain array (psd=None branch: PSD estimated from the data)."""
    strain = np.asarray(strain, dtype=np.float64)
    if strain.ndim == 2:
        res = [whiten(s, delta_t, segment_duration, max_filter_duration, trunc_method, remove_corrupted,
                      low_frequency_cutoff, return_psd) for s in strain]
        if return_psd:
            return np.stack([r[0] for r in res], axis=0), [r[1] for r in res]
        return np.stack(res, axis=0)
    if strain.ndim != 1:
        raise ValueError("Strain must be 1D or 2D.")
    n = len(strain)
    if n % 2:
        raise ValueError("whitening needs an even number of samples")
    sample_rate = 1.0 / delta_t
    seg_len = int(round(segment_duration * sample_rate))
    seg_stride = int(seg_len / 2)
    psd0 = welch_psd(strain, delta_t, seg_len, seg_stride)
    ts_delta_f = 1.0 / (n * delta_t)
    psd = interpolate_psd(psd0, 1.0 / delta_t / seg_len, ts_delta_f)
    max_filter_len = int(max_filter_duration * sample_rate)
    psd = inverse_spectrum_truncation(psd, ts_delta_f, max_filter_len, low_frequency_cutoff, trunc_method)
    inv_psd = 1.0 / psd
    white = np.fft.irfft(np.fft.rfft(strain) * inv_psd ** 0.5, n)
    if remove_corrupted:
        white = white[max_filter_len // 2:(n - max_filter_len // 2)]
    if return_filter:
        return white, psd0, inv_psd ** 0.5
    if return_psd:
        return white, psd0
    return white


def colored_noise(n: int, seed: int, delta_t: float = 1.0 / 2048.0) -> np.ndarray:
    """Synthetic detector-like noise for tests: Gaussian noise shaped by an aLIGO-like analytic amplitude
    spectrum (steep seismic wall below ~20 Hz, bucket near 200 Hz, rising shot noise) plus two narrow lines."""
    rng = np.random.default_rng(seed)
    f = np.fft.rfftfreq(n, delta_t)
    ff = np.maximum(f, 10.0)          # real strain is high-passed: keep the seismic wall finite (Hann leakage)
    asd = 1e-23 * ((20.0 / ff) ** 4.1 * 30 + 1.0 + (ff / 300.0) ** 2)
    asd = asd * (1 + 40 * np.exp(-0.5 * ((f - 60.0) / 0.05) ** 2) + 25 * np.exp(-0.5 * ((f - 500.0) / 0.1) ** 2))
    spec = (rng.standard_normal(len(f)) + 1j * rng.standard_normal(len(f))) * asd * np.sqrt(n / (4 * delta_t))
    spec[0] = 0
    return np.fft.irfft(spec, n)
