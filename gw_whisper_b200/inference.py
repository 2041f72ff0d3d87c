"""Sliding-window search driver: the reference's `inference.py` functions with the same names,
arguments and return values, running the per-window work on the GPU.

  whiten                               MLGWSC-1/inference.py:56-137    (-> whiten.py, GPU)
  get_clusters                         MLGWSC-1/inference.py:140-166   (0.35 s max-clustering)
  SegmentSlicer / TorchSegmentSlicer   MLGWSC-1/inference.py:173-296   (framing + window times)
  build_encoder_with_lora, build_model MLGWSC-1/inference.py:407-434   (the three weight artefacts)
  worker, evaluate_slices              MLGWSC-1/inference.py:437-489   (batch loop, threshold)
  get_triggers                         MLGWSC-1/inference.py:492-589   (all segments of an input file)
  parse_args, main                     MLGWSC-1/inference.py:596-675   (CLI: `python -m gw_whisper_b200.inference`)
  extract_segments                     Signal_vs_Noise/Real_events/preprocess_real_events.py:12-17

Differences that are deliberate:
  * the whole segment is uploaded once and windows are cut on the device (no per-window tensors,
    no DataLoader), scores come back in one copy and triggers are compacted on the device instead
    of a python loop with `.item()` per window (inference.py:482-487);
  * whitening runs on the GPU, so the reference's CPU process pool (`num_workers`, forkserver,
    Manager().dict() chunk hand-over, inference.py:548-575) has nothing left to do: the argument is
    accepted and ignored;
  * files are read through h5py when it is installed and through `hdf5io` (pure Python) otherwise.
"""
from __future__ import annotations

import logging
import os
import sys
import time as _time
from argparse import ArgumentParser
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .hdf5io import open_file

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
        self.segment_duration = segment_duration
        self.max_filter_duration = max_filter_duration
        self.whitened_file = whitened_file
        self.low_frequency_cutoff = low_frequency_cutoff
        dss = [infile[det][key] for det in self.detectors]
        # sampling interval is stored inverted in attrs (inference.py:196-197)
        self.delta_t = 1.0 / (1.0 / dss[0].attrs["delta_t"])
        self.index_step_size = int(self.step_size / self.delta_t)
        self.time_step_size = self.delta_t * self.index_step_size
        self.start_time = dss[0].attrs["start_time"]
        for ds in dss:
            assert ds.attrs["start_time"] == self.start_time
        self.dss = dss
        self.process(save_psd)

    def process(self, save_psd: bool) -> None:
        """inference.py:218-245: whiten every detector unless `white`, optionally store the whitened data,
        stack, and shift the start time by the 0.125 s the whitening crop removes."""
        from .whiten import whiten
        out: List[np.ndarray] = []
        self.psds: List[Any] = []
        for ds, det in zip(self.dss, self.detectors):
            if self.white:
                new_ds = np.asarray(ds[()])
            else:
                new_ds = whiten(np.asarray(ds[()]), delta_t=self.delta_t,
                                low_frequency_cutoff=self.low_frequency_cutoff,
                                segment_duration=self.segment_duration,
                                max_filter_duration=self.max_filter_duration, return_psd=save_psd)
                if save_psd:
                    new_ds, psd = new_ds
                    self.psds.append(psd)
            out.append(new_ds)
            if self.whitened_file is not None:
                with open_file(self.whitened_file, "a") as wfile:
                    wfile.require_group(det).create_dataset(self.key, data=new_ds)
        self.dss = np.stack(out, axis=0)
        if not self.white:
            self.start_time += 0.125
        self.white = True

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

    def times_are_float32(self) -> bool:
        """Precision of the reference's trigger times for this segment.  `TorchSegmentSlicer.__next__` wraps
        the time in `torch.tensor(ts)` (inference.py:296) and the DataLoader collates those tensors: with the
        np.float64 `start_time` an HDF5 attribute yields (h5py, hdf5io) `ts` is np.float64 and the tensor is
        float64 -- full precision; only a plain Python float start time (in-memory files) makes a float32
        tensor, which at GPS ~1.24e9 quantises to 128 s."""
        return not isinstance(self.start_time, np.generic)     # numpy scalar + float -> np.float64 -> f64 tensor

    def window_times(self, reference_float32: Optional[bool] = None) -> np.ndarray:
        """Time stamp of every window exactly as the reference produces it: a float64 running sum
        `current_time += time_step_size` (inference.py:262) plus peak_offset; rounded to float32 only where the
        reference itself would (`times_are_float32`, or when forced with reference_float32=True)."""
        n = len(self)
        steps = np.full(n, self.time_step_size, dtype=np.float64)
        steps[0] = self.start_time
        t = np.cumsum(steps) + self.peak_offset        # cumsum accumulates sequentially in f64
        if reference_float32 is None:
            reference_float32 = self.times_are_float32()
        return t.astype(np.float32).astype(np.float64) if reference_float32 else t


class TorchSegmentSlicer(SegmentSlicer):
    def __next__(self):
        sl, ts = self.get_next_slice()
        return torch.from_numpy(np.ascontiguousarray(sl)), torch.tensor(ts)


def evaluate_slices(slicer: SegmentSlicer, network, device: str = "cuda", trigger_threshold: float = 0.2,
                    verbose: bool = False, reference_float32_times: Optional[bool] = None
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
        if not strain.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: stream_search needs the segment on the GPU")
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
            if thr <= 0.0:
                thr_raw = -float("inf")        # sigmoid(x) > 0 for every finite logit: everything triggers
            elif thr >= 1.0:
                thr_raw = float("inf")         # sigmoid(x) < 1: nothing does
            else:
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


# =============================================================================
# Builders (MLGWSC-1/inference.py:407-434)
# =============================================================================
WHISPER_BASE_ENV = "GWW_WHISPER_BASE"


def _load_base_encoder(name: str = "openai/whisper-tiny"):
    """The frozen base encoder the reference downloads with `WhisperModel.from_pretrained(name)`
    (inference.py:408).  Offline, in this order: $GWW_WHISPER_BASE (a directory `from_pretrained` accepts, or a
    torch-saved encoder state_dict next to a geometry name `tiny|base|small` in the file name), then the local
    HF cache.  Returns (state_dict with HF WhisperEncoder keys, WhisperGeometry)."""
    from .encoder import WhisperGeometry
    src = os.environ.get(WHISPER_BASE_ENV)
    if src and os.path.isfile(src):
        sd = torch.load(src, map_location="cpu")
        sd = {k[len("encoder."):] if k.startswith("encoder.") else k: v for k, v in sd.items()}
        d = int(sd["conv1.weight"].shape[0])
        size = {384: "tiny", 512: "base", 768: "small"}.get(d)
        if size is None:
            raise RuntimeError(f"{src}: unsupported Whisper width d_model={d}")
        return sd, WhisperGeometry.named(size)
    from transformers import WhisperModel
    try:
        model = WhisperModel.from_pretrained(src or name, local_files_only=True)
    except Exception as e:   # noqa: BLE001
        raise RuntimeError(
            f"cannot load the base Whisper encoder '{src or name}' offline ({type(e).__name__}); set "
            f"${WHISPER_BASE_ENV} to a local checkpoint directory or a torch-saved encoder state_dict") from e
    c = model.config
    geo = WhisperGeometry(c.d_model, c.encoder_layers, c.encoder_attention_heads, c.encoder_ffn_dim)
    return model.encoder.state_dict(), geo


def build_encoder_with_lora(lora_weights_path: str, device: str = "cuda", chunk: int = 2 * BATCH_SIZE):
    """`PeftModel.from_pretrained(WhisperModel.from_pretrained("openai/whisper-tiny").encoder, lora_dir)`
    (inference.py:407-412): here the adapter (LoRA or DoRA, adapter_config.json + adapter_model.safetensors)
    is merged into the base weights when the encoder handle is created."""
    from .encoder import B200WhisperEncoder, load_dora_adapter
    sd, geo = _load_base_encoder()
    dora = load_dora_adapter(lora_weights_path) if lora_weights_path else None
    return B200WhisperEncoder(sd, geo, dora=dora, chunk=chunk)


def build_model(lora_weights_path: str, dense_weights_path: str, adapter_weights_path: str, device: str = "cuda",
                n_detectors: int = 2, usr: bool = False):
    """inference.py:415-434: Q-Adapter (.pt state_dict, `q_transform.*` buffers accepted) + encoder with the
    PEFT adapter + dense head (.pth state_dict of `model.classifier`); `usr` removes the trailing softmax."""
    from .qfrontend import GWWhisperClassifier, QTransformAdapter, remove_softmax_from_classifier
    _set_device(device)
    adapter = QTransformAdapter(n_detectors=n_detectors)
    adapter.load_state_dict(torch.load(adapter_weights_path, map_location="cpu"))
    encoder = build_encoder_with_lora(lora_weights_path, device)
    model = GWWhisperClassifier(whisper_encoder=encoder, n_detectors=n_detectors, q_adapter=adapter)
    model.classifier.load_state_dict(torch.load(dense_weights_path, map_location="cpu"))
    model.refresh()
    if usr:
        remove_softmax_from_classifier(model)
    return model


def _set_device(device: str) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"gw_whisper_b200 has no CPU path (device={device!r}); use 'cuda' or 'cuda:N'")
    if not torch.cuda.is_available():
        raise RuntimeError("gw_whisper_b200 needs a CUDA device (sm_100); none is visible")
    torch.cuda.set_device(dev.index or 0)
    return torch.device("cuda", dev.index or 0)


# =============================================================================
# Evaluation over a file (MLGWSC-1/inference.py:437-451, 492-589)
# =============================================================================
def worker(inp: Dict[str, Any]) -> TorchSegmentSlicer:
    """Prepare one slicer (inference.py:437-451).  The shared-dict hand-over of the reference's process pool
    (`wdata_dict`, split_and_pop / stack_and_load) is not needed in-process and is ignored."""
    inp = dict(inp)
    fpath = inp.pop("fpath")
    key = inp.pop("key")
    inp.pop("wdata_dict", None)
    with open_file(fpath, "r") as infile:
        return TorchSegmentSlicer(infile, key, **inp)


def get_triggers(lora_weights_path: str, dense_weights_path: str, adapter_weights_path: str, inputfile: str,
                 step_size: float = 0.1, trigger_threshold: float = 0.2, device: str = "cuda",
                 verbose: bool = False, white: bool = False, whitened_file: Optional[str] = None,
                 low_frequency_cutoff: float = 20.0, num_workers: int = -1, usr: bool = False,
                 network=None) -> Tuple[Dict[str, List[List[float]]], List[np.ndarray]]:
    """Compute triggers for all segments in `inputfile` (inference.py:492-589): same arguments, same return
    value (triggers per segment key, sorted by key; raw scores per 256-window batch in processing order).
    `network` (extension) reuses an already-built model instead of the three artefact paths."""
    if network is None:
        network = build_model(lora_weights_path=lora_weights_path, dense_weights_path=dense_weights_path,
                              adapter_weights_path=adapter_weights_path, device=device, n_detectors=2, usr=usr)
    network.eval()
    detectors = ["H1", "L1"]
    if whitened_file is not None:
        with open_file(whitened_file, "w") as wfile:
            for d in detectors:
                wfile.create_group(d)
    triggers: Dict[str, List[List[float]]] = {}
    all_vals_all: List[np.ndarray] = []
    arguments: List[Dict[str, Any]] = []
    with open_file(inputfile, "r") as infile:
        det_grp = next(iter(infile.values()))
        for key in list(det_grp.keys()):
            arguments.append(dict(fpath=inputfile, key=key, step_size=step_size,
                                  low_frequency_cutoff=low_frequency_cutoff, white=white,
                                  whitened_file=whitened_file, detectors=detectors))
        arguments.sort(key=(lambda x: len(infile[x["detectors"][0]][x["key"]])), reverse=True)
    # the reference fans the (CPU, pycbc) whitening out to `num_workers` processes; here it runs on the GPU
    for kwargs in arguments:
        slicer = worker(kwargs)
        logging.info("Evaluating %s (%d slices)", slicer.key, len(slicer))
        sub_trigs, sub_vals = evaluate_slices(slicer, network, device=device, trigger_threshold=trigger_threshold,
                                              verbose=verbose)
        triggers[slicer.key] = sub_trigs
        all_vals_all.extend(sub_vals)
    return dict(sorted(triggers.items(), key=lambda x: x[0])), all_vals_all


# =============================================================================
# CLI (MLGWSC-1/inference.py:42-49, 596-675)
# =============================================================================
def configure_logging(verbose: bool = False, debug: bool = False) -> None:
    level = logging.DEBUG if debug else (logging.INFO if verbose else logging.WARNING)
    logging.basicConfig(format="%(levelname)s | %(asctime)s: %(message)s", level=level,
                        datefmt="%d-%m-%Y %H:%M:%S", handlers=[logging.StreamHandler(sys.stdout)], force=True)


def parse_args(argv: Optional[Sequence[str]] = None) -> Any:
    """The reference's flags, names and defaults (inference.py:596-618)."""
    parser = ArgumentParser(description="Apply a trained two-detector GW-Whisper model and save triggers.")
    parser.add_argument("--verbose", action="store_true", help="Print update messages.")
    parser.add_argument("--debug", action="store_true", help="Show debug messages.")
    parser.add_argument("--force", action="store_true", help="Overwrite existing output file.")
    parser.add_argument("inputfile", type=str, help="Path to input HDF5.")
    parser.add_argument("outputfile", type=str, help="Path to output HDF5 (must not exist unless --force).")
    parser.add_argument("--white", action="store_true", help="Input is already whitened (skip whitening).")
    parser.add_argument("--softmax", action="store_true", help="Use Softmax outputs (default is USR logits).")
    parser.add_argument("--coinc-window", type=float, default=0.1, help="(Reserved) coincidence window; not used.")
    parser.add_argument("--lora-weights", type=str, required=True, help="Path to LoRA weights dir.")
    parser.add_argument("--dense-weights", type=str, required=True, help="Path to dense head weights (.pth).")
    parser.add_argument("--adapter-weights", type=str, required=True, help="Path to Q-Adapter weights (.pt).")
    parser.add_argument("-t", "--trigger-threshold", type=float, default=-0.5, help="Trigger threshold on signal score.")
    parser.add_argument("--step-size", type=float, default=0.1, help="Sliding window step (s).")
    parser.add_argument("--cluster-threshold", type=float, default=0.35, help="Time gap for clustering (s).")
    parser.add_argument("--device", type=str, default="cuda", help="Device, e.g. 'cuda', 'cuda:1'.")
    parser.add_argument("--debug-triggers-file", type=str, default=None, help="Save pre-cluster triggers here (optional).")
    parser.add_argument("--debug-whitened-file", type=str, default=None, help="Save whitened inputs to this HDF5 (optional).")
    parser.add_argument("--num-workers", type=int, default=8, help="Accepted for compatibility (whitening runs on the GPU).")
    return parser.parse_args(argv)


def write_trigger_file(path: str, time_arr, stat_arr, var_arr, all_vals_flat) -> None:
    """The four datasets of the reference's output file (inference.py:667-672)."""
    with open_file(path, "w") as outfile:
        outfile.create_dataset("time", data=np.asarray(time_arr, dtype=np.float64))
        outfile.create_dataset("stat", data=np.asarray(stat_arr, dtype=np.float64))
        outfile.create_dataset("var", data=np.asarray(var_arr, dtype=np.float64))
        outfile.create_dataset("all_vals", data=np.asarray(all_vals_flat, dtype=np.float32))


def main(argv: Optional[Sequence[str]] = None) -> None:
    start_time = _time.time()
    args = parse_args(argv)
    configure_logging(verbose=args.verbose, debug=args.debug)
    if os.path.isfile(args.outputfile) and not args.force:
        raise RuntimeError("Output file exists. Use --force to overwrite.")
    if args.debug_whitened_file is not None and os.path.isfile(args.debug_whitened_file) and not args.force:
        raise RuntimeError("Whitened file exists. Use --force to overwrite.")
    if args.debug_triggers_file is not None and os.path.isfile(args.debug_triggers_file) and not args.force:
        raise RuntimeError("Triggers file exists. Use --force to overwrite.")
    triggers, all_vals = get_triggers(
        lora_weights_path=args.lora_weights, dense_weights_path=args.dense_weights,
        adapter_weights_path=args.adapter_weights, inputfile=args.inputfile, step_size=args.step_size,
        trigger_threshold=args.trigger_threshold, device=args.device, verbose=args.verbose, white=args.white,
        whitened_file=args.debug_whitened_file, low_frequency_cutoff=20.0, num_workers=args.num_workers,
        usr=not args.softmax)
    logging.info("Total slices above threshold %.3f: %d", args.trigger_threshold,
                 sum(len(v) for v in triggers.values()))
    if args.debug_triggers_file is not None:
        with open_file(args.debug_triggers_file, "w") as dbg:
            for key, trig_list in triggers.items():
                dbg.create_dataset(key, data=np.array(trig_list, dtype=np.float32).reshape(-1, 2))
    time_arr, stat_arr, var_arr = get_clusters(triggers, args.cluster_threshold)
    all_vals_flat = np.concatenate(all_vals).astype("float32") if len(all_vals) else np.array([], dtype="float32")
    write_trigger_file(args.outputfile, time_arr, stat_arr, var_arr, all_vals_flat)
    print(f"Total execution time: {_time.time() - start_time:.2f} seconds")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
