"""CPU oracle for front end B: the Q-transform (`ml4gw.transforms.QScan`) and the Q-Adapter CNN.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

PARITY UNPINNED for QScan: the arithmetic lives in the third-party package `ml4gw`, which is neither
pinned in /root/reference/requirements.txt nor installed in this image, and the reference ships no
tests or golden vectors.  This file restates the published algorithm of
ml4gw's `QTile` / `SingleQTransform` / `QScan` (itself a batched port of GWpy v3.0.8
`gwpy/signal/qtransform.py`, gwpy==3.0.8 is pinned at requirements.txt:84) and is FROZEN as this
project's specification of Q1; it is anchored on the reference's call sites
(MLGWSC-1/inference.py:316-321 constructor arguments, :345 call with a 2-D [B, 2048] tensor).

  X = rfft(x, norm="forward"); X[1:] *= 2
  per plane q = qmin * exp(sqrt2 * dq * (i + 1/2)):
    frequencies f = fmin * base^(j+1/2), floored to multiples of 1/duration, unique
    per row (q, f): qprime = q/sqrt(11); windowsize = 2*int(f/qprime*dur)+1;
        indices = round(arange(-half, half+1) + 1 + f*dur)
        window  = (1 - (k/dur * qprime/f)^2)^2 * ntiles/(dur*fs) * sqrt(315 qprime/(128 f))
        ntiles  = 2^ceil(log2(2 pi f dur / q / deltam)),  deltam = 2 sqrt(mismatch/3)
        zero-pad (int((pad-1)/2), int((pad+1)/2)), ifftshift, ifft, |.|^2, / median (torch.quantile)
  plane choice = arg-max over planes of the max tile energy over the WHOLE call (batch-coupled)
  interpolation: bicubic (cubic convolution A=-0.75, align_corners=False) along time to 512 per
  row, rows stacked as if uniformly spaced, then bicubic along frequency to 512.

The Q-Adapter (MLGWSC-1/inference.py:303-351) is restated verbatim in structure (same parameter
names, so reference state_dicts load) and checked against the reference class in tests.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class QTile(nn.Module):
    def __init__(self, q: float, frequency: float, duration: float, sample_rate: float, mismatch: float):
        super().__init__()
        self.mismatch = mismatch
        self.q = q
        self.deltam = 2 * (mismatch / 3.0) ** 0.5
        self.qprime = q / 11 ** 0.5
        self.frequency = float(frequency)
        self.duration = duration
        self.sample_rate = sample_rate
        self.windowsize = 2 * int(self.frequency / self.qprime * self.duration) + 1
        pad = self.ntiles() - self.windowsize
        self.register_buffer("padding", torch.Tensor((int((pad - 1) / 2.0), int((pad + 1) / 2.0))))
        self.register_buffer("indices", self.get_data_indices())
        self.register_buffer("window", self.get_window())

    def ntiles(self) -> int:
        tcum_mismatch = self.duration * 2 * math.pi * self.frequency / self.q
        return int(2 ** math.ceil(math.log2(tcum_mismatch / self.deltam)))

    def _get_indices(self) -> torch.Tensor:
        half = int((self.windowsize - 1) / 2)
        return torch.arange(-half, half + 1)

    def get_window(self) -> torch.Tensor:
        wfrequencies = self._get_indices() / self.duration
        xfrequencies = wfrequencies * self.qprime / self.frequency
        norm = (self.ntiles() / (self.duration * self.sample_rate)
                * (315 * self.qprime / (128 * self.frequency)) ** 0.5)
        return torch.Tensor((1 - xfrequencies ** 2) ** 2 * norm)

    def get_data_indices(self) -> torch.Tensor:
        return torch.round(self._get_indices() + 1 + self.frequency * self.duration).type(torch.long)

    def forward(self, fseries: torch.Tensor, norm: Optional[str] = "median") -> torch.Tensor:
        windowed = fseries[..., self.indices] * self.window
        left, right = self.padding
        padded = F.pad(windowed, (int(left), int(right)), mode="constant")
        wenergy = torch.fft.ifftshift(padded, dim=-1)
        tdenergy = torch.fft.ifft(wenergy)
        energy = tdenergy.real ** 2.0 + tdenergy.imag ** 2.0
        if norm == "median":
            med = torch.quantile(energy, q=0.5, dim=-1, keepdim=True)
            energy = energy / med
        elif norm == "mean":
            energy = energy / energy.mean(dim=-1, keepdim=True)
        return energy


class SingleQTransform(nn.Module):
    def __init__(self, duration: float, sample_rate: float, spectrogram_shape: Sequence[int], q: float = 12,
                 frange: Optional[List[float]] = None, mismatch: float = 0.2):
        super().__init__()
        self.q = q
        self.spectrogram_shape = list(spectrogram_shape)
        self.frange = list(frange) if frange is not None else [0, math.inf]
        self.duration = duration
        self.mismatch = mismatch
        qprime = self.q / 11 ** 0.5
        if self.frange[0] <= 0:
            self.frange[0] = 50 * self.q / (2 * math.pi * duration)
        if math.isinf(self.frange[1]):
            self.frange[1] = sample_rate / 2 / (1 + 1 / qprime)
        self.freqs = self.get_freqs()
        self.qtile_transforms = nn.ModuleList(
            [QTile(self.q, float(f), self.duration, sample_rate, self.mismatch) for f in self.freqs])
        self.qtiles: Optional[List[torch.Tensor]] = None

    def get_freqs(self) -> torch.Tensor:
        minf, maxf = self.frange
        fcum_mismatch = math.log(maxf / minf) * (2 + self.q ** 2) ** 0.5 / 2.0
        deltam = 2 * (self.mismatch / 3.0) ** 0.5
        nfreq = int(max(1, math.ceil(fcum_mismatch / deltam)))
        fstep = fcum_mismatch / nfreq
        fstepmin = 1 / self.duration
        freq_base = math.exp(2 / ((2 + self.q ** 2) ** 0.5) * fstep)
        freqs = torch.Tensor([minf * freq_base ** (i + 0.5) for i in range(nfreq)])
        freqs = torch.div(freqs, fstepmin, rounding_mode="floor") * fstepmin
        return torch.unique(freqs)

    def get_max_energy(self) -> torch.Tensor:
        return torch.stack([t.max() for t in self.qtiles]).max()

    def compute_qtiles(self, X: torch.Tensor, norm: Optional[str] = "median") -> None:
        X = torch.fft.rfft(X, norm="forward")
        X[..., 1:] *= 2
        self.qtiles = [qt(X, norm) for qt in self.qtile_transforms]

    def interpolate(self) -> torch.Tensor:
        num_f, num_t = self.spectrogram_shape
        # each qtile is [1, B, ntiles] (1 batch x B channels, see QScan.forward); bicubic over
        # (channel, time) with the channel extent unchanged == 1-D cubic convolution along time
        res = [F.interpolate(qt[None], (qt.shape[-2], num_t), mode="bicubic") for qt in self.qtiles]
        res = torch.cat(res, dim=0)                 # [nfreq, 1, B, T]
        res = torch.transpose(res, 0, 2)            # [B, 1, nfreq, T]
        res = F.interpolate(res, (num_f, num_t), mode="bicubic")
        return torch.squeeze(res)


class QScan(nn.Module):
    def __init__(self, duration: float, sample_rate: float, spectrogram_shape: Sequence[int],
                 qrange: Sequence[float] = (4, 64), frange: Optional[List[float]] = None, mismatch: float = 0.2):
        super().__init__()
        self.qrange = list(qrange)
        self.mismatch = mismatch
        self.qs = self.get_qs()
        self.frange = list(frange) if frange is not None else [0, math.inf]
        self.spectrogram_shape = list(spectrogram_shape)
        self.q_transforms = nn.ModuleList(
            [SingleQTransform(duration, sample_rate, spectrogram_shape, q, list(self.frange), mismatch)
             for q in self.qs])

    def get_qs(self) -> List[float]:
        deltam = 2 * (self.mismatch / 3.0) ** 0.5
        cumum = math.log(self.qrange[1] / self.qrange[0]) / 2 ** 0.5
        nplanes = int(max(math.ceil(cumum / deltam), 1))
        dq = cumum / nplanes
        return [self.qrange[0] * math.exp(2 ** 0.5 * dq * (i + 0.5)) for i in range(nplanes)]

    @torch.no_grad()
    def forward(self, X: torch.Tensor, norm: Optional[str] = "median", return_plane: bool = False):
        # the reference passes x[:, i] of shape [B, 2048] (inference.py:345): ml4gw promotes it to
        # 3-D as 1 batch x B channels, so the plane choice below is coupled across the batch
        while X.dim() < 3:
            X = X[None]
        for t in self.q_transforms:
            t.compute_qtiles(X, norm)
        idx = int(torch.argmax(torch.Tensor([float(t.get_max_energy()) for t in self.q_transforms])))
        out = self.q_transforms[idx].interpolate()
        return (out, idx) if return_plane else out

    def tiling_plan(self) -> List[dict]:
        """(q, rows, ntiles per row) -- used by tests to pin SURVEY.md's 148 rows / 49 664 tiles."""
        plan = []
        for t in self.q_transforms:
            plan.append({"q": t.q, "freqs": [float(f) for f in t.freqs],
                         "ntiles": [qt.ntiles() for qt in t.qtile_transforms],
                         "windowsize": [qt.windowsize for qt in t.qtile_transforms]})
        return plan


class QTransformAdapter(nn.Module):
    """MLGWSC-1/inference.py:303-351, same parameter names (freq_adapter.{0,3,6,8}, scale, bias,
    film_gamma, film_beta) so reference adapter checkpoints load with strict=True (QScan buffers
    are registered under q_transform.* exactly like ml4gw's)."""

    def __init__(self, kernel_length: float = 1.0, sample_rate: int = 2048, q_range: Sequence[int] = (4, 128),
                 spectrogram_shape: Sequence[int] = (512, 512), target_shape: Tuple[int, int] = (80, 3000),
                 n_detectors: int = 2, channels: Sequence[int] = (16, 32, 64)):
        # channels=(32, 64, 128) with spectrogram_shape=(128, 128) is the adapter of MLGWSC-1/train.py:104,118-123
        super().__init__()
        self.n_detectors = n_detectors
        c1, c2, c3 = channels
        self.q_transform = QScan(duration=kernel_length, sample_rate=sample_rate,
                                 spectrogram_shape=list(spectrogram_shape), qrange=list(q_range))
        self.freq_adapter = nn.Sequential(
            nn.Conv2d(1, c1, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(c1, c2, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(c2, c3, 3, padding=1), nn.ReLU(), nn.Conv2d(c3, 1, 1))
        self.final_pool = nn.AdaptiveAvgPool2d(target_shape)
        self.scale = nn.Parameter(torch.ones(1))
        self.bias = nn.Parameter(torch.zeros(1))
        self.film_gamma = nn.Parameter(torch.ones(n_detectors))
        self.film_beta = nn.Parameter(torch.zeros(n_detectors))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, D, _ = x.shape
        outs = []
        for i in range(D):
            with torch.no_grad():
                qspec = self.q_transform(x[:, i]).unsqueeze(1)
            y = self.freq_adapter(qspec)
            y = self.final_pool(y).squeeze(1)
            y = self.scale * y + self.bias
            y = y * self.film_gamma[i] + self.film_beta[i]
            outs.append(y)
        return torch.stack(outs, dim=1)
