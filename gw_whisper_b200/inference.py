"""Sliding-window search driver: the reference's `inference.py` functions with the same names,
arguments and return values, running the per-window work on the GPU.

  SegmentSlicer / TorchSegmentSlicer   MLGWSC-1/inference.py:173-296   (framing + window times)
  evaluate_slices                      MLGWSC-1/inference.py:454-489   (batch loop, threshold)
  get_clusters                         MLGWSC-1/inference.py:140-166   (0.35 s max-clustering)
  extract_segments                     Signal_vs_Noise/Real_events/preprocess_real_events.py:12-17

Differences that are deliberate:
  * the whole segment is uploaded once and windows are cut on the device (no per-window tensors,
    no DataLoader), scores come back in one copy and triggers are compacted on the device instead
    of a python loop with `.item()` per window (inference.py:482-487);
  * whitening (inference.py:56-137, pycbc) is upstream of this path: segments must be whitened
    (`white=True`), otherwise NotImplementedError.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

SLICE_LENGTH = 2048
BATCH_SIZE = 256          # DataLoader(batch_size=256) at inference.py:465


def extract_segments(data, window_size: int = 2048, step_size: int = 204):
    """List of overlapping windows (views), preprocess_real_events.py:12-17."""
    return [data[s:s + window_size] for s in range(0, len(data) - window_size + 1, step_size)]


def get_clusters(triggers: Dict[str, List[List[float]]], cluster_threshold: float = 0.35
                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Cluster per-key triggers (time-ordered); a new cluster starts when the gap to the previous
    trigger exceeds `cluster_threshold`; each cluster is represented by its max-score trigger
    (first one on ties, like np.argmax); var is the constant 0.2 (inference.py:140-166)."""
    times: List[float] = []
    vals: List[float] = []
    for trig_list in triggers.values():
        if len(trig_list) == 0:
            continue
        arr = np.asarray(trig_list, dtype=np.float64).reshape(-1, 2)
        t, s = arr[:, 0], arr[:, 1]
        new = np.ones(len(t), dtype=bool)
        new[1:] = (t[1:] - t[:-1]) > cluster_threshold
        starts = np.flatnonzero(new)
        ends = np.append(starts[1:], len(t))
        for a, b in zip(starts, ends):
            k = a + int(np.argmax(s[a:b]))
            times.append(t[k])
            vals.append(s[k])
    return np.array(times), np.array(vals), np.array([0.2] * len(times))


class _ArrayDataset:
    """Minimal stand-in for an h5py dataset: `ds[()]`, `ds.attrs`, `ds.dtype`, `ds.shape`."""

    def __init__(self, data: np.ndarray, start_time: float, delta_t: float):
        self._d = np.asarray(data)
        self.attrs = {"start_time": start_time, "delta_t": delta_t}
        self.dtype, self.shape = self._d.dtype, self._d.shape

    def __getitem__(self, idx):
        return self._d[idx]

    def __len__(self):
        return len(self._d)


class ArrayFile(dict):
    """In-memory strain file with the MLGWSC-1 layout `file[det][str(int(start))]`
    (MLGWSC-1/generate_data.py:197-216) for machines without h5py."""

    @classmethod
    def from_segments(cls, segments: Dict[str, Dict[str, np.ndarray]], start_times: Dict[str, float],
                      delta_t: float = 1.0 / 2048):
        f = cls()
        for det, segs in segments.items():
            f[det] = {k: _ArrayDataset(v, start_times[k], delta_t) for k, v in segs.items()}
        return f


class SegmentSlicer:
    """Slice multi-detector whitened strain for inference (inference.py:173-264)."""

    def __init__(self, infile, key: str, step_size: float = 0.1, peak_offset: float = 0.6,
                 slice_length: int = 2048, detectors: Optional[List[str]] = None, white: bool = False,
                 whitened_file: Optional[str] = None, save_psd: bool = False,
                 low_frequency_cutoff: Optional[float] = None, segment_duration: float = 0.5,
                 max_filter_duration: float = 0.25) -> None:
        self.step_size = step_size
        self.peak_offset = peak_offset
        self.slice_length = slice_length
        self.detectors = detectors or ["H1", "L1"]
        self.white = white
        self.key = key
        dss = [infile[det][key] for det in self.detectors]
        # sampling interval is stored inverted in attrs (inference.py:196-197)
        self.delta_t = 1.0 / (1.0 / dss[0].attrs["delta_t"])
        self.index_step_size = int(self.step_size / self.delta_t)
        self.time_step_size = self.delta_t * self.index_step_size
        self.start_time = dss[0].attrs["start_time"]
        for ds in dss:
            assert ds.attrs["start_time"] == self.start_time
        if not self.white:
            raise NotImplementedError(
                "whitening (MLGWSC-1/inference.py:56-137, pycbc) is upstream of the B200 path; "
                "pass already-whitened strain with white=True")
        self.dss = np.stack([np.asarray(ds[()]) for ds in dss], axis=0)
        self.psds: List = []

    def __len__(self) -> int:
        return 1 + (self.dss.shape[1] - self.slice_length) // self.index_step_size

    def __iter__(self):
        self.current_index = 0
        self.current_time = self.start_time
        return self

    def get_next_slice(self):
        if self.current_index + self.slice_length > self.dss.shape[1]:
            raise StopIteration
        sl = self.dss[:, self.current_index:self.current_index + self.slice_length]
        ts = self.current_time + self.peak_offset
        self.current_index += self.index_step_size
        self.current_time += self.time_step_size
        return sl, ts

    def __next__(self):
        return self.get_next_slice()

    def window_times(self, reference_float32: bool = True) -> np.ndarray:
        """Time stamp of every window exactly as the reference produces it: a float64 running sum
        `current_time += time_step_size` (inference.py:262) plus peak_offset, then -- because the
        DataLoader collates `torch.tensor(ts)` as float32 (inference.py:296) -- rounded
        to float32 when `reference_float32` is set."""
        n = len(self)
        steps = np.full(n, self.time_step_size, dtype=np.float64)
        steps[0] = self.start_time
        t = np.cumsum(steps) + self.peak_offset        # cumsum accumulates sequentially in f64
        return t.astype(np.float32).astype(np.float64) if reference_float32 else t


class TorchSegmentSlicer(SegmentSlicer):
    def __next__(self):
        sl, ts = self.get_next_slice()
        return torch.from_numpy(np.ascontiguousarray(sl)), torch.tensor(ts)


def evaluate_slices(slicer: SegmentSlicer, network, device: str = "cuda", trigger_threshold: float = 0.2,
                    verbose: bool = False, reference_float32_times: bool = True
                    ) -> Tuple[List[List[float]], List[np.ndarray]]:
    """Run `network` over all slices; return triggers [[time, score], ...] and raw scores (one array
    per 256-window batch, like the reference's `all_vals`)."""
    n = len(slicer)
    if n <= 0:
        return [], []
    times = slicer.window_times(reference_float32_times)
    strain = torch.from_numpy(np.ascontiguousarray(slicer.dss, dtype=np.float32)).to(device)
    if hasattr(network, "stream_search"):
        scores, trig_idx, trig_sc = network.stream_search(strain, slicer.index_step_size, n, trigger_threshold)
    else:  # generic module on [B, D, 2048] batches, same batching as the reference
        outs = []
        with torch.no_grad():
            for k0 in range(0, n, BATCH_SIZE):
                idx = torch.arange(k0, min(k0 + BATCH_SIZE, n), device=strain.device) * slicer.index_step_size
                win = strain[:, (idx[:, None] + torch.arange(SLICE_LENGTH, device=strain.device)[None, :])]
                outs.append(network(win.permute(1, 0, 2).contiguous())[:, 0])
        scores = torch.cat(outs)
        keep = (scores > trigger_threshold).nonzero().flatten()
        trig_idx, trig_sc = keep, scores[keep]
    scores = scores.float().cpu().numpy()
    trig_idx = trig_idx.cpu().numpy()
    trig_sc = trig_sc.float().cpu().numpy()
    triggers = [[float(times[i]), float(s)] for i, s in zip(trig_idx, trig_sc)]
    all_vals = [scores[k0:k0 + BATCH_SIZE] for k0 in range(0, n, BATCH_SIZE)]
    return triggers, all_vals


class LogMelStreamNetwork:
    """Sliding-window network for the log-mel models (Real_events variant,
    evaluation_real_events.py:29-64): wraps a gw_whisper_b200 classifier so `evaluate_slices` can use
    the fused device-side path (window gather + front end + encoder + head + compaction)."""

    def __init__(self, model, sigmoid: bool = False):
        self.model = model
        self.sigmoid = sigmoid

    def stream_search(self, strain: torch.Tensor, hop: int, n_windows: int, thr: float,
                      first_window: int = 0):
        m = self.model
        m._sync_head()
        enc = m.encoder
        lib = _lib.load()
        D, N = strain.shape
        dev = strain.device
        scores = torch.empty(n_windows, dtype=torch.float32, device=dev)
        tidx = torch.empty(max(n_windows, 1), dtype=torch.long, device=dev)
        tsc = torch.empty(max(n_windows, 1), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        chunk = max(D, min(enc.chunk, n_windows * D))
        ws = enc.workspace(chunk)
        thr_raw = float(thr)
        if self.sigmoid:   # threshold on sigmoid(logit) == threshold on the logit itself
            thr_raw = float(np.log(thr / (1.0 - thr)))
        _lib.check(lib.gww_stream_search_logmel(
            enc._handle, strain.contiguous().data_ptr(), D, N, hop, first_window, n_windows, thr_raw,
            scores.data_ptr(), tidx.data_ptr(), tsc.data_ptr(), cnt.data_ptr(), n_windows,
            ws.data_ptr(), ws.numel(), chunk, _lib.stream_ptr()))
        c = int(cnt.item())
        if self.sigmoid:
            return torch.sigmoid(scores), tidx[:c], torch.sigmoid(tsc[:c])
        return scores, tidx[:c], tsc[:c]


def get_triggers_from_file(network, infile, step_size: float = 0.1, trigger_threshold: float = 0.2,
                           device: str = "cuda", verbose: bool = False, detectors: Sequence[str] = ("H1", "L1"),
                           ) -> Tuple[Dict[str, List[List[float]]], List[np.ndarray]]:
    """Body of the reference's get_triggers loop (inference.py:532-589) for an already-built network
    and an open strain file (h5py.File or ArrayFile): segments sorted by length (desc), sliced with
    white=True, evaluated one after the other."""
    keys = sorted(infile[detectors[0]].keys(), key=lambda k: -len(infile[detectors[0]][k]))
    triggers: Dict[str, List[List[float]]] = {}
    all_vals: List[np.ndarray] = []
    for key in keys:
        slicer = TorchSegmentSlicer(infile, key, step_size=step_size, detectors=list(detectors), white=True)
        trig, vals = evaluate_slices(slicer, network, device=device, trigger_threshold=trigger_threshold,
                                     verbose=verbose)
        triggers[key] = trig
        all_vals.extend(vals)
    return triggers, all_vals
