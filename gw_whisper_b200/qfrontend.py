"""Front end B and the MLGWSC-1 model on the B200 path: the reference's `QTransformAdapter`,
`GWWhisperClassifier` and `remove_softmax_from_classifier` (MLGWSC-1/inference.py:303-400) with the
same constructor arguments, attribute names, forward signatures and state_dict keys.

The Q-transform (`ml4gw.transforms.QScan`, inference.py:316-321,345) and the adapter CNN run in
hand-written CUDA (csrc/qfront.cuh) behind `gww_qscan` / `gww_qadapter` / `gww_forward_windows_qscan`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .encoder import B200WhisperEncoder


def _f32(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())


def _p(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_float_p)


class QScanB200:
    """`ml4gw.transforms.QScan(duration, sample_rate, spectrogram_shape, qrange)` on the GPU.
    `__call__(x[B, 2048]) -> [B, F, T]`; like ml4gw the plane is chosen once per call from the peak
    normalised tile energy over the whole batch."""

    def __init__(self, duration: float = 1.0, sample_rate: float = 2048, spectrogram_shape: Sequence[int] = (512, 512),
                 qrange: Sequence[float] = (4, 128), mismatch: float = 0.2):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.spectrogram_shape = [int(spectrogram_shape[0]), int(spectrogram_shape[1])]
        _lib.check(self._lib.gww_qfront_create(float(duration), float(sample_rate), float(qrange[0]), float(qrange[1]),
                                               float(mismatch), self.spectrogram_shape[0], self.spectrogram_shape[1],
                                               C.byref(self._h)))
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._lib.gww_qfront_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_planes, self.n_rows, self.n_tiles = a.value, b.value, c.value
        self._ws: Optional[torch.Tensor] = None
        self._ws_n = 0
        self._keep: List[np.ndarray] = []

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.gww_qfront_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def tiling_plan(self) -> dict:
        q = np.zeros(self.n_planes, dtype=np.float64)
        plane = np.zeros(self.n_rows, dtype=np.int32)
        freq = np.zeros(self.n_rows, dtype=np.float32)
        nt = np.zeros(self.n_rows, dtype=np.int32)
        wsz = np.zeros(self.n_rows, dtype=np.int32)
        off = np.zeros(self.n_rows, dtype=np.int32)
        _lib.check(self._lib.gww_qfront_plan(self._h, q.ctypes.data, plane.ctypes.data, freq.ctypes.data,
                                             nt.ctypes.data, wsz.ctypes.data, off.ctypes.data))
        return {"q": q, "plane": plane, "freq": freq, "ntiles": nt, "windowsize": wsz, "offset": off}

    def workspace(self, n: int) -> torch.Tensor:
        if self._ws is None or self._ws_n < n:
            nbytes = self._lib.gww_qfront_workspace_bytes(self._h, n)
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            self._ws_n = n
        return self._ws

    def __call__(self, x: torch.Tensor, return_tiles: bool = False, return_plane: bool = False):
        if x.dim() != 2 or x.shape[-1] != 2048:
            raise ValueError(f"expected strain [B, 2048], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
        x = x.contiguous().float()
        n = x.shape[0]
        F, T = self.spectrogram_shape
        spec = torch.empty((n, F, T), dtype=torch.float32, device=x.device)
        tiles = torch.empty((n, self.n_tiles), dtype=torch.float32, device=x.device) if return_tiles else None
        plane = torch.zeros(1, dtype=torch.int32, device=x.device)
        ws = self.workspace(n)
        _lib.check(self._lib.gww_qscan(self._h, x.data_ptr(), n, 2048, spec.data_ptr(), _lib.ptr(tiles),
                                       plane.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        out = [spec]
        if return_tiles:
            out.append(tiles)
        if return_plane:
            out.append(int(plane.item()))
        return out[0] if len(out) == 1 else tuple(out)


class QTransformAdapter(nn.Module):
    """Convert raw strain [B, D, T] into Whisper-like features [B, D, 80, 3000] (inference.py:303-351).
    Parameters keep the reference names (`freq_adapter.{0,3,6,8}.{weight,bias}`, `scale`, `bias`,
    `film_gamma`, `film_beta`); `q_transform.*` buffers of ml4gw checkpoints are accepted and ignored
    (the tiling plan is rebuilt from the constructor arguments)."""

    def __init__(self, kernel_length: float = 1.0, sample_rate: int = 2048, q_range: Sequence[int] = (4, 128),
                 spectrogram_shape: Sequence[int] = (512, 512), target_shape: Tuple[int, int] = (80, 3000),
                 n_detectors: int = 2, channels: Sequence[int] = (16, 32, 64)) -> None:
        super().__init__()
        if tuple(target_shape) != (80, 3000):
            raise ValueError("the Whisper encoder needs target_shape == (80, 3000)")
        self.n_detectors = n_detectors
        self.channels = tuple(int(c) for c in channels)
        c1, c2, c3 = self.channels
        object.__setattr__(self, "q_transform", QScanB200(kernel_length, sample_rate, spectrogram_shape, q_range))
        self.freq_adapter = nn.Sequential(
            nn.Conv2d(1, c1, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(c1, c2, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(c2, c3, 3, padding=1), nn.ReLU(), nn.Conv2d(c3, 1, 1))
        self.final_pool = nn.AdaptiveAvgPool2d(target_shape)
        self.scale = nn.Parameter(torch.ones(1))
        self.bias = nn.Parameter(torch.zeros(1))
        self.film_gamma = nn.Parameter(torch.ones(n_detectors))
        self.film_beta = nn.Parameter(torch.zeros(n_detectors))
        self._dirty = True

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.startswith("q_transform.")}
        r = super().load_state_dict(sd, strict=strict, **kw)
        self._dirty = True
        return r

    def refresh(self) -> None:
        self._dirty = True

    def _sync(self) -> None:
        if not self._dirty:
            return
        fa = self.freq_adapter
        arrs = [_f32(fa[0].weight), _f32(fa[0].bias), _f32(fa[3].weight), _f32(fa[3].bias),
                _f32(fa[6].weight), _f32(fa[6].bias), _f32(fa[8].weight), _f32(fa[8].bias),
                _f32(self.film_gamma), _f32(self.film_beta)]
        w = _lib.QAdapterWeights()
        (w.conv1_w, w.conv1_b, w.conv2_w, w.conv2_b, w.conv3_w, w.conv3_b, w.conv4_w, w.conv4_b) = [_p(a) for a in arrs[:8]]
        w.scale, w.bias = float(self.scale.item()), float(self.bias.item())
        w.n_detectors = int(self.n_detectors)
        w.film_gamma, w.film_beta = _p(arrs[8]), _p(arrs[9])
        w.c1, w.c2, w.c3 = self.channels
        qt = self.q_transform
        _lib.check(qt._lib.gww_qfront_set_adapter(qt._h, C.byref(w)))
        qt._ws, qt._ws_n = None, 0       # the activation buffers are sized by the CNN widths: re-query
        self._dirty = False

    @torch.no_grad()
    def adapt(self, qspec: torch.Tensor, det: int) -> torch.Tensor:
        """freq_adapter + final_pool + affine + FiLM[det] on a Q-spectrogram batch [B, F, T]."""
        self._sync()
        qt = self.q_transform
        n = qspec.shape[0]
        out = torch.empty((n, 80, 3000), dtype=torch.float32, device=qspec.device)
        ws = qt.workspace(n)
        _lib.check(qt._lib.gww_qadapter(qt._h, qspec.contiguous().data_ptr(), n, det, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, D, _ = x.shape
        self._sync()
        outs = [self.adapt(self.q_transform(x[:, i]), i) for i in range(D)]
        return torch.stack(outs, dim=1)


class TrainQTransformAdapter(QTransformAdapter):
    """The adapter geometry MLGWSC-1/train.py trains (`QTransformAdapter` there, :78-160): a 128x128 Q-spectrogram
    (:104) and a 32 / 64 / 128 CNN (:118-123); same parameter names, so its checkpoints
    (`save_model_components`, train.py:723) load here.  Runs on the generic fp32 convolution kernels."""

    def __init__(self, kernel_length: float = 1.0, sample_rate: int = 2048, q_range: Sequence[int] = (4, 128),
                 spectrogram_shape: Sequence[int] = (128, 128), target_shape: Tuple[int, int] = (80, 3000),
                 n_detectors: int = 2) -> None:
        super().__init__(kernel_length, sample_rate, q_range, spectrogram_shape, target_shape, n_detectors,
                         channels=(32, 64, 128))


class GWWhisperClassifier(nn.Module):
    """Q-Adapter -> Whisper encoder (per detector) -> MLP classifier (inference.py:354-392)."""

    def __init__(self, whisper_encoder, n_detectors: int, num_classes: int = 2,
                 q_adapter: Optional[QTransformAdapter] = None, use_last_token: bool = True) -> None:
        super().__init__()
        if not isinstance(whisper_encoder, B200WhisperEncoder):
            raise TypeError("gw_whisper_b200 models need a B200WhisperEncoder (no PyTorch fallback path)")
        self.n_detectors = n_detectors
        object.__setattr__(self, "encoder", whisper_encoder)
        self.adapter = q_adapter if q_adapter is not None else QTransformAdapter(n_detectors=n_detectors)
        self.use_last_token = use_last_token
        hidden = whisper_encoder.config.d_model
        self.classifier = nn.Sequential(
            nn.Linear(hidden * n_detectors, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
            nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, num_classes),
            nn.Softmax(dim=1))  # removed in USR mode
        self._head_dirty = True

    def __setattr__(self, name, value):
        if name == "classifier":
            object.__setattr__(self, "_head_dirty", True)
        super().__setattr__(name, value)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._head_dirty = True
        return r

    def refresh(self) -> None:
        self._head_dirty = True
        self.adapter.refresh()

    def _sync_head(self) -> None:
        if self._head_dirty or self.encoder._head_key != id(self):
            from .models import _linears
            lin = _linears(self.classifier)
            softmax = any(isinstance(m, nn.Softmax) for m in self.classifier)
            self.encoder.set_head(lin, softmax=softmax)
            self.encoder._head_key = id(self)
            self._head_dirty = False
        self.adapter._sync()

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, D, 2048] f32 CUDA -> [B, num_classes]; the B windows form one QScan call per detector."""
        if x.dim() != 3 or x.shape[-1] != 2048 or x.shape[1] != self.n_detectors:
            raise ValueError(f"expected strain [B, {self.n_detectors}, 2048], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
        self._sync_head()
        enc, qt = self.encoder, self.adapter.q_transform
        x = x.contiguous().float()
        B, D, _ = x.shape
        out = torch.empty((B, enc._head_out), dtype=torch.float32, device=x.device)
        ws, qws = enc.workspace(B * D), qt.workspace(B)
        _lib.check(enc._lib.gww_forward_windows_qscan(
            enc._handle, qt._h, x.data_ptr(), B, D, int(self.use_last_token), out.data_ptr(),
            ws.data_ptr(), ws.numel(), qws.data_ptr(), qws.numel(), _lib.stream_ptr()))
        return out

    def stream_search(self, strain: torch.Tensor, hop: int, n_windows: int, thr: float, first_window: int = 0,
                      batch: int = 256):
        """Sliding-window search used by `inference.evaluate_slices`: batches of 256 consecutive windows
        (the reference's DataLoader batch, inference.py:465), score = out[:, 0] (:481)."""
        if not strain.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: stream_search needs the segment on the GPU")
        self._sync_head()
        enc, qt = self.encoder, self.adapter.q_transform
        D, N = strain.shape
        dev = strain.device
        scores = torch.empty(n_windows, dtype=torch.float32, device=dev)
        tidx = torch.empty(max(n_windows, 1), dtype=torch.long, device=dev)
        tsc = torch.empty(max(n_windows, 1), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        batch = max(1, min(batch, n_windows))
        ws, qws = enc.workspace(batch * D), qt.workspace(batch)
        _lib.check(enc._lib.gww_stream_search_qscan(
            enc._handle, qt._h, strain.contiguous().float().data_ptr(), D, N, hop, first_window, n_windows, batch,
            float(thr), scores.data_ptr(), tidx.data_ptr(), tsc.data_ptr(), cnt.data_ptr(), n_windows,
            ws.data_ptr(), ws.numel(), qws.data_ptr(), qws.numel(), _lib.stream_ptr()))
        c = int(cnt.item())
        return scores, tidx[:c], tsc[:c]


def remove_softmax_from_classifier(model: GWWhisperClassifier) -> None:
    """Switch to USR mode (raw logits), inference.py:395-400."""
    if isinstance(model.classifier, nn.Sequential) and len(model.classifier) > 0:
        layers = list(model.classifier.children())
        if isinstance(layers[-1], nn.Softmax):
            model.classifier = nn.Sequential(*layers[:-1])
