"""Host side of the drop-in boundary that needs no GPU: the CLI flag surface equals the reference's, artefact
loaders accept the reference's on-disk formats, error behaviour without a device."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
have_ref = os.path.isdir(REF)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_cli_flags_match_reference_parse_args(monkeypatch):
    from test_inference_host import _load_reference_inference
    from gw_whisper_b200 import inference as I
    ref = _load_reference_inference()
    argv = ["in.hdf", "out.hdf", "--lora-weights", "l", "--dense-weights", "d", "--adapter-weights", "a"]
    variants = [[], ["--white", "--softmax", "-t", "0.3", "--step-size", "0.2", "--cluster-threshold", "0.5",
                     "--device", "cuda:1", "--num-workers", "0", "--force", "--verbose", "--debug",
                     "--debug-triggers-file", "x", "--debug-whitened-file", "y", "--coinc-window", "0.2"]]
    for extra in variants:
        monkeypatch.setattr(sys, "argv", ["inference.py"] + argv + extra)
        want = vars(ref.parse_args())
        got = vars(I.parse_args(argv + extra))
        assert got == want


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_shipped_dora_adapters_and_heads_load():
    """The PEFT adapter directories and dense heads the reference ships load through the product loaders."""
    from gw_whisper_b200.encoder import load_dora_adapter
    import glob
    dirs = sorted(set(os.path.dirname(p) for p in glob.glob(os.path.join(REF, "**", "adapter_model.safetensors"), recursive=True)))
    assert len(dirs) >= 3
    for d in dirs:
        a = load_dora_adapter(d)
        assert a["r"] == 8 and a["lora_alpha"] == 32 and a["use_dora"]
        keys = list(a["tensors"].keys())
        assert any(k.endswith("lora_magnitude_vector") for k in keys) and len(keys) == 24
    sd = torch.load(os.path.join(REF, "Signal_vs_Noise/results/Single_detector/models/best_dense_layers.pth"),
                    map_location="cpu")
    assert sorted(sd.keys())[:2] == ["0.bias", "0.weight"]


def test_no_cpu_path_errors_are_loud():
    from gw_whisper_b200 import inference as I
    with pytest.raises(RuntimeError):
        I._set_device("cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            I._set_device("cuda")
        from gw_whisper_b200 import resample_timeseries
        with pytest.raises(RuntimeError):
            resample_timeseries(np.zeros(2048, dtype=np.float32))


def test_base_encoder_loader_from_state_dict(tmp_path, monkeypatch):
    from gw_whisper_b200 import inference as I
    from gw_whisper_b200 import synthetic as S
    enc = S.make_encoder("tiny", 0)
    p = str(tmp_path / "enc.pt")
    torch.save({"encoder." + k: v for k, v in enc.state_dict().items()}, p)     # WhisperModel-style prefix accepted
    monkeypatch.setenv(I.WHISPER_BASE_ENV, p)
    sd, geo = I._load_base_encoder()
    assert geo.d_model == 384 and geo.encoder_layers == 4 and "conv1.weight" in sd
    monkeypatch.setenv(I.WHISPER_BASE_ENV, str(tmp_path / "missing-dir"))
    with pytest.raises(RuntimeError):
        I._load_base_encoder()
