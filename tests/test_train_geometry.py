"""the adapter geometry MLGWSC-1/train.py trains (128x128 Q-spectrogram, 32/64/128 CNN,
train.py:78-160) and the Efficiency_test "mutual subtraction" replacement of the softmax
(Signal_vs_Noise/Efficiency_test/src/test_network.py:89-99).  Golden vectors come from the reference's own train.py
class (tests/golden/make_train_adapter_golden.py, oracle QScan injected for the absent ml4gw)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import qscan as OQ

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_adapter_golden.npz"))


def _state():
    return {k[len("adapter."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("adapter.")}


def _nerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _check(feats, tol):
    feats = feats.detach().cpu().numpy()
    e1 = _nerr(feats[..., ::25], G["feats_sub"])
    e2 = _nerr(feats[:, :, G["feats_rows"], :], G["feats_fullrows"])
    print(f"train-geometry adapter features vs reference train.py class: {e1:.3e} (columns), {e2:.3e} (rows)")
    assert e1 <= tol and e2 <= tol


def test_restated_train_adapter_equals_reference_class():
    a = OQ.QTransformAdapter(spectrogram_shape=(128, 128), channels=(32, 64, 128), n_detectors=2).eval()
    a.load_state_dict(_state(), strict=False)
    with torch.no_grad():
        _check(a(torch.from_numpy(G["strain"])), 1e-6)


def test_tiling_plan_is_independent_of_the_spectrogram_shape():
    from gw_whisper_b200.qfrontend import QScanB200
    a, b = QScanB200(1.0, 2048, [512, 512], [4, 128]), QScanB200(1.0, 2048, [128, 128], [4, 128])
    assert (a.n_planes, a.n_rows, a.n_tiles) == (b.n_planes, b.n_rows, b.n_tiles) == (5, 148, 49664)
    with pytest.raises(RuntimeError):
        QScanB200(1.0, 2048, [100, 128], [4, 128])


def test_mutual_subtraction_tail_structure():
    """CPU part: the Sequential surgery equals the reference's (last layer Linear(2,2,bias=False), fixed weight)."""
    from gw_whisper_b200.models import replace_softmax_by_mutual_subtraction

    class Dummy(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Sequential(nn.Linear(8, 2), nn.Softmax(dim=1))
    d = Dummy()
    replace_softmax_by_mutual_subtraction(d)
    last = list(d.classifier.children())[-1]
    assert isinstance(last, nn.Linear) and last.bias is None
    assert torch.equal(last.weight, torch.tensor([[1.0, -1.0], [-1.0, 1.0]]))
    x = torch.randn(5, 8)
    y = d.classifier(x)
    z = d.classifier[0](x)
    assert torch.allclose(y, torch.stack([z[:, 0] - z[:, 1], z[:, 1] - z[:, 0]], dim=1))
    with pytest.raises(ValueError):
        replace_softmax_by_mutual_subtraction(d)                     # no softmax left (test_network.py:98-99)


@pytest.mark.gpu
def test_cuda_train_geometry_adapter_vs_reference_class_golden():
    from gw_whisper_b200 import TrainQTransformAdapter
    a = TrainQTransformAdapter(n_detectors=2)
    a.load_state_dict(_state())
    _check(a(torch.from_numpy(G["strain"]).cuda()), 1e-4)


@pytest.mark.gpu
def test_cuda_efficiency_test_network_softmax_and_mutual_subtraction():
    """Efficiency_test/src/network.py:69-90 (1-detector, 2 classes, Softmax) and its USR form on the B200 path."""
    from oracle import encoder as E, logmel as L
    from gw_whisper_b200 import (B200WhisperEncoder, logmel_features, one_channel_ligo_binary_classifier,
                                 replace_softmax_by_mutual_subtraction)
    dev = torch.device("cuda")
    base = E.make_encoder("tiny", 0, spread=True)
    dora = E.synthetic_dora("tiny", targets=("k_proj", "v_proj"))                 # Efficiency_test/src/train.py:58
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=8)
    model = one_channel_ligo_binary_classifier(enc, num_classes=2, softmax=True)
    E.seeded_head(model.classifier, seed=4, gain=3.0)
    model.refresh()
    ref = E.OneChannelOracle(E.attach_dora(base, dora), head=E.head_one_channel(384, 2, softmax=True)).eval()
    ref.classifier.load_state_dict(model.classifier.state_dict())
    g = torch.Generator().manual_seed(5)
    strain = torch.randn(16, 2048, generator=g)
    feats = logmel_features(strain.to(dev))
    with torch.no_grad():
        want_p = ref.to(dev)(feats).cpu()
        want_u = ref.classifier[:-1](ref.encoder(feats).last_hidden_state[:, -1, :]).cpu()
    got_p = model(feats).cpu()
    replace_softmax_by_mutual_subtraction(model)
    got_u = model(feats).cpu()
    want_u = torch.stack([want_u[:, 0] - want_u[:, 1], want_u[:, 1] - want_u[:, 0]], dim=1)
    ep, eu = (got_p - want_p).abs().max().item(), (got_u - want_u).abs().max().item()
    print(f"efficiency-test network: softmax err {ep:.3e}, mutual-subtraction (USR) err {eu:.3e}")
    assert ep < 2e-2 and eu < 2e-2 and torch.allclose(got_u[:, 0], -got_u[:, 1], atol=1e-6)
