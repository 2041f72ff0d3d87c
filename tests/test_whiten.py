"""Whitening: the oracle restatement of the reference's `whiten`
(MLGWSC-1/inference.py:56-137 through pycbc 2.4.0 -- PARITY UNPINNED, pycbc is not installed) checked for the
properties the algorithm must have, and the GPU implementation (gww_whiten) against that oracle."""
import numpy as np
import pytest
import torch

from oracle import whiten as W

FS = 2048


def test_oracle_welch_segmentation_and_median_bias():
    # pycbc.psd.welch: n // stride segments minus the ones that do not fit; symmetric trim
    x = np.random.default_rng(0).standard_normal(10 * FS + 137)
    psd = W.welch_psd(x, 1 / FS, 1024, 512)
    assert psd.shape == (513,)
    # white unit-variance noise: one-sided PSD = 2 / fs
    assert abs(np.median(psd[5:500]) * FS / 2 - 1.0) < 0.1
    assert W.median_bias(1) == 1.0 and abs(W.median_bias(3) - (1 + 1 / 3 - 1 / 2)) < 1e-15
    assert W.median_bias(1000) == float(np.log(2))


def test_oracle_whiten_properties():
    n = 32 * FS
    x = W.colored_noise(n, 3)
    w = W.whiten(x, low_frequency_cutoff=20.0)
    assert w.shape == (n - 512,) and w.dtype == np.float64
    # scale invariance: the PSD is estimated from the data itself
    w2 = W.whiten(1e3 * x, low_frequency_cutoff=20.0)
    assert np.allclose(w, w2, rtol=1e-9, atol=1e-9 * np.abs(w).max())
    # the result is white above the cut-off: Welch PSD of the output is flat (ratio of band medians ~ 1)
    p = W.welch_psd(w, 1 / FS, 1024, 512)
    band = lambda lo, hi: np.median(p[lo // 2:hi // 2])      # noqa: E731   (2 Hz bins)
    flat = [band(30, 60), band(100, 200), band(300, 500), band(700, 1000)]
    assert max(flat) / min(flat) < 1.25
    assert band(2, 14) < 1e-3 * band(100, 200)               # below the cut-off the filter is zero
    # 2-D input == per-channel; return_psd returns the un-interpolated Welch estimate
    both, psds = W.whiten(np.stack([x, x[::-1].copy()]), low_frequency_cutoff=20.0, return_psd=True)
    assert np.array_equal(both[0], w) and psds[0].shape == (513,)
    with pytest.raises(ValueError):
        W.whiten(x[:-1])


def test_filter_is_effectively_compact():
    """The design fact whiten.cuh relies on: irfft(psd_trunc^-1/2) decays fast beyond the 512 retained taps."""
    n = 64 * FS
    x = W.colored_noise(n, 1)
    _, _, Q = W.whiten(x, low_frequency_cutoff=20.0, return_filter=True)
    imp = np.roll(np.fft.irfft(Q, n), n // 2)
    c = n // 2
    tail = lambda L: np.abs(np.concatenate([imp[:c - L], imp[c + L + 1:]])).max() / np.abs(imp).max()   # noqa: E731
    assert tail(256) < 1e-5 and tail(8192) < 1e-8
    # applying the filter truncated to +-8192 taps instead of the N-point product: < 1e-4 of the output rms, and
    # the difference lives below the cut-off (above 30 Hz: < 1e-6)
    w = W.whiten(x, low_frequency_cutoff=20.0, remove_corrupted=False)
    imp[:c - 8192] = 0
    imp[c + 8193:] = 0
    y = np.fft.irfft(np.fft.rfft(x) * np.fft.rfft(np.roll(imp, -c)), n)
    rms = np.sqrt(np.mean(w * w))
    assert np.abs(y - w).max() / rms < 1e-4
    D = np.fft.rfft(y - w)
    D[:30 * 64] = 0                                          # bins below 30 Hz (delta_f = 1/64 Hz)
    assert np.abs(np.fft.irfft(D, n)).max() / rms < 1e-6


# ---------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    return float(np.abs(a - b).max() / np.sqrt(np.mean(b * b)))


def _rel_above(a, b, f_lo, fs=FS):
    """same, after removing everything below f_lo Hz from the difference"""
    n = len(a)
    D = np.fft.rfft(a - b)
    D[:int(f_lo * n / fs)] = 0
    return float(np.abs(np.fft.irfft(D, n)).max() / np.sqrt(np.mean(b * b)))


@pytest.mark.gpu
@pytest.mark.parametrize("seconds,seed,flow", [(8, 1, 20.0), (32, 2, 20.0), (33, 3, None), (300, 4, 20.0)])
def test_gpu_whiten_vs_oracle(seconds, seed, flow):
    from gw_whisper_b200.whiten import whiten
    n = seconds * FS + (2 * seed if seconds == 33 else 0)     # one case with a length that is not a multiple of 512
    x = W.colored_noise(n, seed) * 1e20                       # O(1e-3..1) numbers, like dyn-range scaled strain
    want, psd_want = W.whiten(x, low_frequency_cutoff=flow, return_psd=True)
    got, psd_got = whiten(x, low_frequency_cutoff=flow, return_psd=True)
    assert got.shape == want.shape == (n - 512,) and got.dtype == np.float64
    e_psd = float(np.abs(psd_got - psd_want).max() / psd_want.max())
    e = _rel(got, want)
    print(f"whiten {seconds} s (n={n}, f_low={flow}): Welch PSD rel err {e_psd:.3e}; whitened strain max err / rms {e:.3e}")
    e_band = _rel_above(got, want, 30.0)
    print(f"   ... above 30 Hz: {e_band:.3e}")
    assert e_psd < 1e-9
    assert e < 1e-4          # FIR truncation of psd_trunc^-1/2 at +-8192 taps (exact when the segment is shorter)
    assert e_band < 2e-5     # most of the truncation error lives below the low-frequency cut-off (the rest sits on
                             # the two narrow 40x / 25x spectral lines of the synthetic noise)
    if n <= 2 * 8192:
        assert e < 1e-7


@pytest.mark.gpu
def test_gpu_whiten_two_channels_scale_invariance_and_f32():
    from gw_whisper_b200.whiten import whiten, whiten_device
    n = 64 * FS
    x = np.stack([W.colored_noise(n, 5), W.colored_noise(n, 6)]) * 1e20
    a = whiten(x, low_frequency_cutoff=20.0)
    b = whiten(7.5 * x, low_frequency_cutoff=20.0)
    assert a.shape == (2, n - 512)
    assert _rel(b, a) < 1e-9                                   # size-independent property: scale invariance
    f32 = whiten_device(torch.from_numpy(x[1]).cuda(), low_frequency_cutoff=20.0, out_dtype=torch.float32)
    assert f32.dtype == torch.float32 and np.array_equal(f32.cpu().numpy(), a[1].astype(np.float32))
    with pytest.raises(RuntimeError):
        whiten(x[0, :-1])                                      # odd length


@pytest.mark.gpu
def test_segment_slicer_whitens_when_not_white():
    """SegmentSlicer(white=False): whitened on the GPU, start time shifted by 0.125 s, 512 fewer samples
    (inference.py:218-252)."""
    from gw_whisper_b200 import inference as I
    n = 40 * FS
    seg = np.stack([W.colored_noise(n, 7), W.colored_noise(n, 8)]) * 1e20
    f = I.ArrayFile.from_segments({"H1": {"100": seg[0]}, "L1": {"100": seg[1]}}, {"100": np.float64(100.0)})
    s = I.TorchSegmentSlicer(f, "100", white=False, low_frequency_cutoff=20.0)
    assert s.white and s.start_time == 100.125 and s.dss.shape == (2, n - 512)
    assert len(s) == 1 + (n - 512 - 2048) // 204
    want = W.whiten(seg, low_frequency_cutoff=20.0)
    assert _rel(s.dss, want) < 1e-4
