"""Whisper encoder with DoRA merged at load, running on hand-written sm_100a kernels.

Drop-in for the `encoder` argument of the reference model classes
(Signal_vs_Noise/src/model.py:4-52, Glitch_classification/src/model.py:4-39,
MLGWSC-1/inference.py:354-392): exposes `.config.d_model` and
`__call__(input_features[B,80,3000]).last_hidden_state` like HF `WhisperEncoder`
(modeling_whisper.py:593-648) / `PeftModel(encoder)` (MLGWSC-1/inference.py:407-412).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib

N_CTX = 1500
N_MELS = 80
N_FRAMES = 3000


@dataclass
class WhisperGeometry:
    d_model: int
    encoder_layers: int
    encoder_attention_heads: int
    encoder_ffn_dim: int

    @staticmethod
    def named(name: str) -> "WhisperGeometry":
        table = {"tiny": (384, 4, 6, 1536), "base": (512, 6, 8, 2048), "small": (768, 12, 12, 3072)}
        return WhisperGeometry(*table[name])


def load_dora_adapter(adapter_dir: str) -> Dict[str, object]:
    """Reads a PEFT adapter directory (adapter_config.json + adapter_model.safetensors), e.g.
    Signal_vs_Noise/results/*/models/best_lora_weights*/ -- the artefact
    `PeftModel.from_pretrained(encoder, lora_dir)` consumes at MLGWSC-1/inference.py:411."""
    from safetensors.numpy import load_file

    with open(os.path.join(adapter_dir, "adapter_config.json")) as fh:
        cfg = json.load(fh)
    tensors = load_file(os.path.join(adapter_dir, "adapter_model.safetensors"))
    return {"tensors": tensors, "r": int(cfg["r"]), "lora_alpha": float(cfg["lora_alpha"]),
            "use_dora": bool(cfg.get("use_dora", False))}


def _f32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_lib.c_float_p)


class B200WhisperEncoder:
    """encoder(feats).last_hidden_state on B200.  `state_dict` uses HF WhisperEncoder key names."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], geometry: WhisperGeometry,
                 dora: Optional[Dict[str, object]] = None, chunk: int = 256):
        self.config = SimpleNamespace(d_model=geometry.d_model, encoder_layers=geometry.encoder_layers,
                                      encoder_attention_heads=geometry.encoder_attention_heads,
                                      encoder_ffn_dim=geometry.encoder_ffn_dim,
                                      num_mel_bins=N_MELS, max_source_positions=N_CTX)
        self.chunk = int(chunk)
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        self._ws: Optional[torch.Tensor] = None
        self._ws_chunk = 0
        self._head_key = None
        keep = []  # host arrays referenced by the ctypes structs during create

        def g(name):
            a = _f32(state_dict[name])
            keep.append(a)
            return a

        L = geometry.encoder_layers
        layers = (_lib.LayerWeights * L)()
        for i in range(L):
            p = f"layers.{i}."
            lw = layers[i]
            lw.ln1_g, lw.ln1_b = _p(g(p + "self_attn_layer_norm.weight")), _p(g(p + "self_attn_layer_norm.bias"))
            lw.q_w, lw.q_b = _p(g(p + "self_attn.q_proj.weight")), _p(g(p + "self_attn.q_proj.bias"))
            lw.k_w = _p(g(p + "self_attn.k_proj.weight"))
            lw.v_w, lw.v_b = _p(g(p + "self_attn.v_proj.weight")), _p(g(p + "self_attn.v_proj.bias"))
            lw.o_w, lw.o_b = _p(g(p + "self_attn.out_proj.weight")), _p(g(p + "self_attn.out_proj.bias"))
            lw.ln2_g, lw.ln2_b = _p(g(p + "final_layer_norm.weight")), _p(g(p + "final_layer_norm.bias"))
            lw.fc1_w, lw.fc1_b = _p(g(p + "fc1.weight")), _p(g(p + "fc1.bias"))
            lw.fc2_w, lw.fc2_b = _p(g(p + "fc2.weight")), _p(g(p + "fc2.bias"))
            if dora is not None:
                scale = float(dora["lora_alpha"]) / float(dora["r"])
                for proj, field in (("q_proj", "dora_q"), ("k_proj", "dora_k"), ("v_proj", "dora_v"),
                                    ("out_proj", "dora_o")):
                    base = f"base_model.model.layers.{i}.self_attn.{proj}."
                    t = dora["tensors"]
                    if base + "lora_A.weight" not in t:
                        continue
                    A, B = _f32(t[base + "lora_A.weight"]), _f32(t[base + "lora_B.weight"])
                    mkey = base + "lora_magnitude_vector"
                    if mkey in t:
                        m = _f32(t[mkey])
                    elif mkey + ".weight" in t:
                        m = _f32(t[mkey + ".weight"])
                    else:  # plain LoRA: magnitude == row norm of the updated weight => no renormalisation
                        W0 = _f32(state_dict[f"layers.{i}.self_attn.{proj}.weight"])
                        m = np.linalg.norm(W0 + scale * (B @ A), axis=1).astype(np.float32)
                    keep.extend([A, B, m])
                    dd = getattr(lw, field)
                    dd.lora_A, dd.lora_B, dd.magnitude = _p(A), _p(B), _p(m)
                    dd.r, dd.scale = int(dora["r"]), scale
        ew = _lib.EncoderWeights()
        ew.conv1_w, ew.conv1_b = _p(g("conv1.weight")), _p(g("conv1.bias"))
        ew.conv2_w, ew.conv2_b = _p(g("conv2.weight")), _p(g("conv2.bias"))
        ew.pos_emb = _p(g("embed_positions.weight"))
        ew.ln_post_g, ew.ln_post_b = _p(g("layer_norm.weight")), _p(g("layer_norm.bias"))
        ew.layers = layers
        cfg = _lib.EncoderConfig(geometry.d_model, L, geometry.encoder_attention_heads,
                                 geometry.encoder_ffn_dim)
        _lib.check(self._lib.gww_model_create(C.byref(cfg), C.byref(ew), C.byref(self._handle)))

    # -- construction helpers ---------------------------------------------------------------
    @classmethod
    def from_hf(cls, hf_encoder, dora: Optional[Dict[str, object]] = None, chunk: int = 256):
        """Build from a `transformers` WhisperEncoder (e.g. WhisperModel(...).encoder)."""
        c = hf_encoder.config
        geo = WhisperGeometry(c.d_model, c.encoder_layers, c.encoder_attention_heads, c.encoder_ffn_dim)
        return cls(hf_encoder.state_dict(), geo, dora=dora, chunk=chunk)

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self._lib.gww_model_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # -- nn.Module-ish surface used by the reference scripts ---------------------------------
    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self

    def parameters(self):
        return iter(())

    # -- plumbing ----------------------------------------------------------------------------
    def workspace(self, chunk: Optional[int] = None) -> torch.Tensor:
        chunk = int(chunk or self.chunk)
        if self._ws is None or self._ws_chunk < chunk:
            nbytes = self._lib.gww_workspace_bytes(self._handle, chunk)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            self._ws_chunk = chunk
        return self._ws

    def set_head(self, linears, softmax: bool = False) -> None:
        """Upload the classifier (list of (weight[out,in], bias[out]) tensors)."""
        hw = _lib.HeadWeights()
        keep = []
        hw.n_layers = len(linears)
        hw.softmax = int(softmax)
        hw.dims[0] = int(linears[0][0].shape[1])
        for i, (w, b) in enumerate(linears):
            wa, ba = _f32(w), _f32(b)
            keep.extend([wa, ba])
            hw.dims[i + 1] = int(wa.shape[0])
            hw.w[i], hw.b[i] = _p(wa), _p(ba)
        _lib.check(self._lib.gww_model_set_head(self._handle, C.byref(hw)))
        self._head_out = int(hw.dims[len(linears)])

    # -- forward -----------------------------------------------------------------------------
    def __call__(self, input_features: torch.Tensor, **_):
        return self.forward(input_features)

    def forward(self, input_features: torch.Tensor, **_):
        x = self._check_feats(input_features)
        n = x.shape[0]
        out = torch.empty((n, N_CTX, self.config.d_model), dtype=torch.float32, device=x.device)
        chunk = min(self.chunk, max(n, 1))
        ws = self.workspace(chunk)
        _lib.check(self._lib.gww_encoder_forward(self._handle, x.data_ptr(), n, out.data_ptr(), None, 1,
                                                 ws.data_ptr(), ws.numel(), chunk, _lib.stream_ptr()))
        return SimpleNamespace(last_hidden_state=out)

    def pooled(self, input_features: torch.Tensor, use_last_token: bool = True) -> torch.Tensor:
        """== encoder(feats).last_hidden_state[:, -1, :] (or .mean(1)) without materialising it."""
        x = self._check_feats(input_features)
        n = x.shape[0]
        out = torch.empty((n, self.config.d_model), dtype=torch.float32, device=x.device)
        chunk = min(self.chunk, max(n, 1))
        ws = self.workspace(chunk)
        _lib.check(self._lib.gww_encoder_forward(self._handle, x.data_ptr(), n, None, out.data_ptr(),
                                                 int(use_last_token), ws.data_ptr(), ws.numel(), chunk,
                                                 _lib.stream_ptr()))
        return out

    def head(self, reps: torch.Tensor) -> torch.Tensor:
        reps = reps.contiguous().float()
        out = torch.empty((reps.shape[0], self._head_out), dtype=torch.float32, device=reps.device)
        _lib.check(self._lib.gww_head_forward(self._handle, reps.data_ptr(), reps.shape[0], out.data_ptr(),
                                              _lib.stream_ptr()))
        return out

    def forward_windows_logmel(self, strain: torch.Tensor, return_pooled: bool = False):
        """strain [B, D, 2048] f32 (cuda) -> head output [B, C] through the fused path."""
        if strain.dim() != 3 or strain.shape[-1] != 2048:
            raise ValueError(f"expected strain [B, D, 2048], got {tuple(strain.shape)}")
        if not strain.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: strain must be a CUDA tensor")
        s = strain.contiguous().float()
        B, D, _ = s.shape
        out = torch.empty((B, self._head_out), dtype=torch.float32, device=s.device)
        pooled = torch.empty((B * D, self.config.d_model), dtype=torch.float32, device=s.device) \
            if return_pooled else None
        chunk = max(D, min(self.chunk, B * D))
        ws = self.workspace(chunk)
        _lib.check(self._lib.gww_forward_windows_logmel(
            self._handle, s.data_ptr(), B, D, out.data_ptr(), _lib.ptr(pooled), ws.data_ptr(), ws.numel(),
            chunk, _lib.stream_ptr()))
        return (out, pooled) if return_pooled else out

    def _check_feats(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 3 or x.shape[1] != N_MELS:
            raise ValueError(f"expected input_features [B, 80, 3000], got {tuple(x.shape)}")
        if x.shape[-1] != N_FRAMES:
            # same check and wording as HF modeling_whisper.py:613-617
            raise ValueError(
                f"Whisper expects the mel input features to be of length {N_FRAMES}, but found "
                f"{x.shape[-1]}. Make sure to pad the input mel features to {N_FRAMES}.")
        if not x.is_cuda:
            raise RuntimeError("gw_whisper_b200 has no CPU path: input_features must be a CUDA tensor")
        return x.contiguous().float()
