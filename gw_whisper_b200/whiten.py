"""`whiten` of the reference (MLGWSC-1/inference.py:56-137) on the GPU: same name, arguments and return values.

The reference whitens each detector of a segment through pycbc (Welch-median PSD over 0.5 s segments, PSD
interpolation, inverse-spectrum truncation to 0.25 s with a Hann window above `low_frequency_cutoff`, FFT
division, 0.125 s cropped at both ends) in a CPU process pool; here one C-ABI call per detector
(`gww_whiten`, csrc/whiten.cuh) does the same in f64 on the device.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib


def whiten_device(strain: torch.Tensor, delta_t: float = 1.0 / 2048.0, segment_duration: float = 0.5,
                  max_filter_duration: float = 0.25, trunc_method: Optional[str] = "hann",
                  remove_corrupted: bool = True, low_frequency_cutoff: Optional[float] = None,
                  out_dtype: torch.dtype = torch.float64, return_psd: bool = False, fir_half: int = 0):
    """One detector: strain CUDA f64 [n] -> whitened CUDA tensor [n - max_filter_len] (f64 or f32)."""
    if strain.dim() != 1:
        raise ValueError("whiten_device takes one channel [n]")
    if not strain.is_cuda:
        raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
    if trunc_method not in ("hann", None):
        raise ValueError(f"unknown trunc_method {trunc_method!r}")
    if out_dtype not in (torch.float64, torch.float32):
        raise ValueError("out_dtype must be float64 or float32")
    lib = _lib.load()
    x = strain.contiguous().to(torch.float64)
    n = x.numel()
    sample_rate = 1.0 / delta_t
    seg_len = int(round(segment_duration * sample_rate))          # TimeSeries.psd
    seg_stride = int(seg_len / 2)
    max_filter_len = int(max_filter_duration * sample_rate)       # inference.py:86
    nbytes = lib.gww_whiten_workspace_bytes(n, seg_len, seg_stride, max_filter_len, fir_half)
    if nbytes == 0:
        raise RuntimeError(f"gww error: {lib.gww_last_error().decode()}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    n_out = n - max_filter_len if remove_corrupted else n
    out = torch.empty(max(n_out, 0), dtype=out_dtype, device=x.device)
    psd = torch.empty(seg_len // 2 + 1, dtype=torch.float64, device=x.device) if return_psd else None
    _lib.check(lib.gww_whiten(
        x.data_ptr(), n, float(delta_t), seg_len, seg_stride, max_filter_len,
        float(low_frequency_cutoff) if low_frequency_cutoff else 0.0, int(trunc_method == "hann"),
        int(bool(remove_corrupted)), int(fir_half),
        out.data_ptr() if out_dtype == torch.float64 else None,
        out.data_ptr() if out_dtype == torch.float32 else None,
        _lib.ptr(psd), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return (out, psd) if return_psd else out


def whiten(strain: np.ndarray, delta_t: float = 1.0 / 2048.0, segment_duration: float = 0.5,
           max_filter_duration: float = 0.25, trunc_method: Optional[str] = "hann", remove_corrupted: bool = True,
           low_frequency_cutoff: Optional[float] = None, psd: Optional[np.ndarray] = None, return_psd: bool = False,
           **kwargs: Any) -> Union[np.ndarray, Tuple[np.ndarray, Any]]:
    """Whiten a 1D or 2D strain array (inference.py:56-137).  numpy in, numpy float64 out; the returned PSD
    (return_psd) is the un-interpolated Welch estimate as a float64 array with delta_f = 1/segment_duration."""
    if psd is not None:
        raise NotImplementedError("whiten(psd=...): only the psd=None branch (PSD estimated from the data, the one "
                                  "SegmentSlicer uses, inference.py:224-231) is implemented on the GPU")
    if kwargs:
        raise TypeError(f"unsupported welch options: {sorted(kwargs)}")
    strain = np.asarray(strain)
    if strain.ndim == 1:
        x = torch.from_numpy(np.ascontiguousarray(strain, dtype=np.float64)).cuda()
        res = whiten_device(x, delta_t, segment_duration, max_filter_duration, trunc_method, remove_corrupted,
                            low_frequency_cutoff, torch.float64, return_psd)
        if return_psd:
            return res[0].cpu().numpy(), res[1].cpu().numpy()
        return res.cpu().numpy()
    if strain.ndim == 2:
        results = [whiten(sd, delta_t=delta_t, segment_duration=segment_duration,
                          max_filter_duration=max_filter_duration, trunc_method=trunc_method,
                          remove_corrupted=remove_corrupted, low_frequency_cutoff=low_frequency_cutoff,
                          return_psd=return_psd) for sd in strain]
        if return_psd:
            return np.stack([r[0] for r in results], axis=0), [r[1] for r in results]
        return np.stack(results, axis=0)
    raise ValueError("Strain must be 1D or 2D.")
