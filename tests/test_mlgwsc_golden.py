"""Pins front end B's Q-Adapter, the MLGWSC-1 classifier (softmax and USR) and the batch driver to the
REFERENCE's own classes: tests/golden/mlgwsc_golden.npz was produced by importing
/root/reference/MLGWSC-1/inference.py (QTransformAdapter :303-351, GWWhisperClassifier :354-392,
remove_softmax_from_classifier :395-400, evaluate_slices :454-489) with oracle.qscan.QScan injected for the absent
ml4gw (tests/golden/make_mlgwsc_golden.py).  CPU part: the restated oracle classes reproduce those outputs, so they
are a faithful checker.  GPU part (-m gpu): the CUDA path against the same fixtures, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as E
from oracle import qscan as OQ

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mlgwsc_golden.npz"))


def _adapter_state():
    return {k[len("adapter."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("adapter.")}


def _nerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _check_feats(feats, tol):
    feats = feats.detach().cpu().numpy() if isinstance(feats, torch.Tensor) else feats
    e1 = _nerr(feats[..., ::25], G["feats_sub"])
    e2 = _nerr(feats[:, :, G["feats_rows"], :], G["feats_fullrows"])
    print(f"adapter features vs reference-class golden: normalised error {e1:.3e} (column subset), {e2:.3e} (full rows)")
    assert e1 <= tol and e2 <= tol


def test_restated_adapter_equals_reference_class():
    ref = OQ.QTransformAdapter(n_detectors=2).eval()
    ref.load_state_dict(_adapter_state(), strict=False)
    with torch.no_grad():
        feats = ref(torch.from_numpy(G["strain"]))
    assert feats.shape == (4, 2, 80, 3000)
    _check_feats(feats, 1e-6)


def test_restated_classifier_equals_reference_class():
    adapter = OQ.QTransformAdapter(n_detectors=2).eval()
    adapter.load_state_dict(_adapter_state(), strict=False)
    base = E.make_encoder("tiny", 0, spread=True)
    enc = E.attach_dora(base, E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj")))
    head = E.seeded_head(E.head_mlgwsc(384, 2, 2, softmax=True), seed=3, gain=3.0)
    x = torch.from_numpy(G["strain"])
    with torch.no_grad():
        feats = adapter(x)
        reps = torch.cat([enc(feats[:, i]).last_hidden_state[:, -1, :] for i in range(2)], dim=1)
        prob, logit = head(reps), head[:-1](reps)
    assert np.abs(prob.numpy() - G["softmax_last"]).max() < 1e-5
    assert np.abs(logit.numpy() - G["usr_last"]).max() < 1e-4


def test_trigger_times_are_float64_for_hdf5_start_times():
    """with the np.float64 start_time an HDF5 attribute yields, the reference's trigger times
    are float64 (consecutive windows 0.0996 s apart at GPS 1.24e9), not float32-quantised to 128 s."""
    from gw_whisper_b200 import inference as I
    seg = G["seg"]
    st = np.float64(1238166018.0)
    f = I.ArrayFile.from_segments({"H1": {"1238166018": seg[0]}, "L1": {"1238166018": seg[1]}}, {"1238166018": st})
    slicer = I.TorchSegmentSlicer(f, "1238166018", white=True)
    assert not slicer.times_are_float32()
    t = slicer.window_times()
    assert t.dtype == np.float64 and len(t) == 7
    assert np.allclose(np.diff(t), 204 / 2048, rtol=0, atol=1e-6)
    keep = G["seg_scores"] > G["seg_thr"]
    assert np.array_equal(t[keep], G["seg_triggers"][:, 0])              # bit-equal to the reference's times
    # a plain python-float start time is what makes the reference collate float32 times
    f32 = I.ArrayFile.from_segments({"H1": {"k": seg[0]}, "L1": {"k": seg[1]}}, {"k": 1238166018.0})
    s32 = I.TorchSegmentSlicer(f32, "k", white=True)
    assert s32.times_are_float32()
    assert np.all(s32.window_times() == np.float64(np.float32(1238166018.6)))


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_adapter_features_vs_reference_class_golden():
    from gw_whisper_b200 import QTransformAdapter
    ours = QTransformAdapter(n_detectors=2)
    ours.load_state_dict(_adapter_state())
    feats = ours(torch.from_numpy(G["strain"]).cuda())
    _check_feats(feats, 1e-4)                                            # north_star: front-end features within 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("tag,use_last", [("last", True), ("mean", False)])
def test_cuda_classifier_vs_reference_class_golden(tag, use_last):
    from gw_whisper_b200 import (B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter,
                                 remove_softmax_from_classifier)
    base = E.make_encoder("tiny", 0, spread=True)
    dora = E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=8)
    adapter = QTransformAdapter(n_detectors=2)
    adapter.load_state_dict(_adapter_state())
    model = GWWhisperClassifier(enc, 2, num_classes=2, q_adapter=adapter, use_last_token=use_last)
    E.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    x = torch.from_numpy(G["strain"]).cuda()
    prob = model(x).cpu().numpy()
    remove_softmax_from_classifier(model)
    logit = model(x).cpu().numpy()
    e_p = np.abs(prob - G[f"softmax_{tag}"]).max()
    e_l = np.abs(logit - G[f"usr_{tag}"]).max()
    print(f"GWWhisperClassifier[{tag}] vs reference-class golden: softmax err {e_p:.3e}, USR logit err {e_l:.3e}")
    assert e_p < 2e-2 and e_l < 2e-2


@pytest.mark.gpu
def test_cuda_evaluate_slices_vs_reference_evaluate_slices_golden():
    from gw_whisper_b200 import B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter, remove_softmax_from_classifier
    from gw_whisper_b200 import inference as I
    base = E.make_encoder("tiny", 0, spread=True)
    dora = E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=16)
    adapter = QTransformAdapter(n_detectors=2)
    adapter.load_state_dict(_adapter_state())
    model = GWWhisperClassifier(enc, 2, q_adapter=adapter)
    E.seeded_head(model.classifier, seed=3, gain=3.0)
    remove_softmax_from_classifier(model)
    seg = G["seg"]
    st = np.float64(1238166018.0)
    f = I.ArrayFile.from_segments({"H1": {"1238166018": seg[0]}, "L1": {"1238166018": seg[1]}}, {"1238166018": st})
    slicer = I.TorchSegmentSlicer(f, "1238166018", white=True)
    thr = float(G["seg_thr"])
    trig, vals = I.evaluate_slices(slicer, model, device="cuda", trigger_threshold=thr)
    scores = np.concatenate(vals)
    e = np.abs(scores - G["seg_scores"]).max()
    print(f"evaluate_slices scores vs the reference's evaluate_slices: max err {e:.3e}")
    assert e < 2e-2
    ref_t = {float(t): float(s) for t, s in G["seg_triggers"]}
    got_t = {float(t): float(s) for t, s in trig}
    times = slicer.window_times()
    for k, t in enumerate(times):                                       # decisive windows must agree; times bit-equal
        if abs(G["seg_scores"][k] - thr) > 2e-2:
            assert (float(t) in got_t) == (float(t) in ref_t)
    assert set(got_t) <= set(float(t) for t in times)
