"""Golden vectors of the MLGWSC-1 model classes, produced by the REFERENCE's own code: imports
/root/reference/MLGWSC-1/inference.py (its missing third-party modules stubbed) with
`oracle.qscan.QScan` injected as `ml4gw.transforms.QScan`, and runs the reference's `QTransformAdapter`,
`GWWhisperClassifier`, `remove_softmax_from_classifier` and `evaluate_slices` on seeded inputs.

This pins everything of front end B / the MLGWSC-1 model EXCEPT the Q-transform itself (ml4gw is absent and
unpinned upstream: the QScan inside is this project's frozen restatement).  Re-run only where /root/reference
exists:   python tests/golden/make_mlgwsc_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference_inference():
    from oracle import qscan as OQ
    stubs = {}
    for name in ("h5py", "peft", "ml4gw", "ml4gw.transforms", "pycbc", "pycbc.waveform", "pycbc.noise",
                 "pycbc.psd", "pycbc.distributions", "pycbc.detector", "pycbc.types", "pycbc.filter"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
    if "h5py" in stubs:
        stubs["h5py"].File = object
    if "peft" in stubs:
        stubs["peft"].PeftModel = object
    if "ml4gw.transforms" in stubs:
        stubs["ml4gw.transforms"].QScan = OQ.QScan
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_inference_golden", os.path.join(REF, "MLGWSC-1/inference.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in stubs:
            sys.modules.pop(k, None)
    return mod


def strain(B, D, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, 2048, generator=g)
    t = torch.arange(2048) / 2048.0
    x[0, 0] += 8.0 * torch.exp(-(t - 0.5) ** 2 / (2 * 0.01 ** 2)) * torch.sin(2 * np.pi * 180.0 * t)
    x[1, 1] += 5.0 * torch.exp(-(t - 0.3) ** 2 / (2 * 0.03 ** 2)) * torch.sin(2 * np.pi * 420.0 * t)
    return x


def main():
    from oracle import encoder as E
    ref = load_reference_inference()
    torch.manual_seed(11)
    adapter = ref.QTransformAdapter(n_detectors=2).eval()
    with torch.no_grad():
        adapter.scale.fill_(1.3)
        adapter.bias.fill_(0.1)
        adapter.film_gamma.copy_(torch.tensor([0.9, 1.2]))
        adapter.film_beta.copy_(torch.tensor([0.05, -0.1]))
    x = strain(4, 2, 8)
    with torch.no_grad():
        feats = adapter(x)                                           # [4, 2, 80, 3000]
    out = {"strain": x.numpy()}
    for k, v in adapter.state_dict().items():
        if not k.startswith("q_transform."):
            out["adapter." + k] = v.numpy()
    # full resolution is 7.7 MB: keep every 25th time column of all rows + two full rows per (window, detector)
    out["feats_cols"] = np.arange(0, 3000, 25)
    out["feats_sub"] = feats[..., ::25].numpy()
    out["feats_rows"] = np.array([0, 41, 79])
    out["feats_fullrows"] = feats[:, :, [0, 41, 79], :].numpy()
    # the reference classifier class on a seeded tiny encoder (DoRA attached the way PeftModel wraps it)
    base = E.make_encoder("tiny", 0, spread=True)
    dora = E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    enc = E.attach_dora(base, dora)
    for ult in (True, False):
        model = ref.GWWhisperClassifier(whisper_encoder=enc, n_detectors=2, q_adapter=adapter, use_last_token=ult).eval()
        E.seeded_head(model.classifier, seed=3, gain=3.0)
        with torch.no_grad():
            prob = model(x)
            ref.remove_softmax_from_classifier(model)
            logit = model(x)
        tag = "last" if ult else "mean"
        out[f"softmax_{tag}"] = prob.numpy()
        out[f"usr_{tag}"] = logit.numpy()
    # the reference's own evaluate_slices on a short segment (batch loop, score = out[:, 0], threshold)
    model = ref.GWWhisperClassifier(whisper_encoder=enc, n_detectors=2, q_adapter=adapter).eval()
    E.seeded_head(model.classifier, seed=3, gain=3.0)
    ref.remove_softmax_from_classifier(model)
    hop, n_win = 204, 7
    g = torch.Generator().manual_seed(77)
    seg = torch.randn(2, 2048 + hop * (n_win - 1), generator=g).numpy().astype(np.float64)

    class DS:
        def __init__(self, d):
            self.d = d
            self.attrs = {"start_time": np.float64(1238166018.0), "delta_t": np.float64(1.0 / 2048)}
            self.dtype, self.shape = d.dtype, d.shape
        def __getitem__(self, i):
            return self.d[i]
        def __len__(self):
            return len(self.d)
    f = {"H1": {"1238166018": DS(seg[0])}, "L1": {"1238166018": DS(seg[1])}}
    slicer = ref.TorchSegmentSlicer(f, "1238166018", white=True)
    thr = 4.16
    trig, vals = ref.evaluate_slices(slicer, model, device="cpu", trigger_threshold=thr)
    out["seg"] = seg
    out["seg_thr"] = np.float64(thr)
    out["seg_scores"] = np.concatenate(vals)
    out["seg_triggers"] = np.array(trig, dtype=np.float64).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "mlgwsc_golden.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items() if hasattr(v, "shape")})
    print("triggers:", out["seg_triggers"], "scores:", out["seg_scores"])


if __name__ == "__main__":
    main()
