"""Front end B (QScan + Q-Adapter) and the MLGWSC-1 model on a real B200 against the oracle
(oracle/qscan.py restatement of ml4gw QScan -- parity unpinned upstream, frozen as the spec -- and the
reference's QTransformAdapter / GWWhisperClassifier structure) on identical seeded inputs.
Gates: features within 1e-4 (max|a-b| / max|b|), logits within 2e-2 absolute."""
import numpy as np
import pytest
import torch

from oracle import encoder as E
from oracle import qscan as OQ

pytestmark = pytest.mark.gpu


def _nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max())


def _strain(B, D=None, seed=5, inject=True):
    g = torch.Generator().manual_seed(seed)
    shape = (B, 2048) if D is None else (B, D, 2048)
    x = torch.randn(*shape, generator=g)
    if inject:                      # a sine-Gaussian so the planes do not all look like noise
        t = torch.arange(2048) / 2048.0
        sg = 8.0 * torch.exp(-(t - 0.5) ** 2 / (2 * 0.01 ** 2)) * torch.sin(2 * np.pi * 180.0 * t)
        x[0] = x[0] + sg
    return x


def _seeded_adapter(seed=11):
    torch.manual_seed(seed)
    ref = OQ.QTransformAdapter(n_detectors=2).eval()
    with torch.no_grad():
        ref.scale.fill_(1.3)
        ref.bias.fill_(0.1)
        ref.film_gamma.copy_(torch.tensor([0.9, 1.2]))
        ref.film_beta.copy_(torch.tensor([0.05, -0.1]))
    return ref


@pytest.mark.parametrize("B,seed,inject", [(5, 5, True), (3, 6, False), (1, 7, True)])
def test_qscan_tiles_plane_and_spectrogram(B, seed, inject):
    from gw_whisper_b200 import QScanB200
    x = _strain(B, seed=seed, inject=inject)
    ref = OQ.QScan(1.0, 2048, [512, 512], qrange=[4, 128])
    ref_spec, ref_plane = ref(x, return_plane=True)
    ref_spec = ref_spec.reshape(B, 512, 512)
    q = QScanB200()
    spec, tiles, plane = q(x.cuda(), return_tiles=True, return_plane=True)
    pl = q.tiling_plan()
    # tile energies of every plane / row
    i = 0
    worst = 0.0
    for t in ref.q_transforms:
        for qt in t.qtiles:
            r = qt[0]                                              # [B, ntiles]
            got = tiles[:, pl["offset"][i]: pl["offset"][i] + pl["ntiles"][i]].cpu()
            worst = max(worst, _nerr(got, r))
            i += 1
    print(f"qscan B={B}: worst per-row normalised tile error {worst:.3e}; plane ours={plane} oracle={ref_plane}")
    assert worst <= 1e-4
    assert plane == ref_plane
    e = _nerr(spec.cpu(), ref_spec)
    print(f"qscan B={B}: spectrogram normalised error {e:.3e}")
    assert e <= 1e-4


def test_qscan_plane_choice_is_batch_coupled():
    """The same window lands on different planes depending on its batch (ml4gw takes the arg-max over
    the whole call); our kernel must reproduce the oracle's choice in both batches."""
    from gw_whisper_b200 import QScanB200
    t = torch.arange(2048) / 2048.0
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(4, 2048, generator=g)
    lowq = noise.clone()
    lowq[1] += 30.0 * torch.exp(-(t - 0.4) ** 2 / (2 * 0.004 ** 2)) * torch.sin(2 * np.pi * 120.0 * t)   # short burst
    highq = noise.clone()
    highq[2] += 3.0 * torch.sin(2 * np.pi * 800.0 * t) * torch.exp(-(t - 0.5) ** 2 / (2 * 0.2 ** 2))     # long ring
    ref = OQ.QScan(1.0, 2048, [512, 512], qrange=[4, 128])
    q = QScanB200()
    planes = []
    for x in (lowq, highq):
        rs, rp = ref(x, return_plane=True)
        spec, p = q(x.cuda(), return_plane=True)
        assert p == rp
        assert _nerr(spec.cpu(), rs) <= 1e-4
        planes.append(p)
    print("planes chosen:", planes)
    assert planes[0] != planes[1], "test inputs should exercise two different planes"


def test_qadapter_features_vs_oracle():
    from gw_whisper_b200 import QTransformAdapter
    ref = _seeded_adapter()
    x = _strain(4, 2, seed=8)
    with torch.no_grad():
        want = ref(x)                                              # [4, 2, 80, 3000]
    ours = QTransformAdapter(n_detectors=2)
    ours.load_state_dict(ref.state_dict())
    got = ours(x.cuda()).cpu()
    assert got.shape == want.shape
    for i in range(2):
        e = _nerr(got[:, i], want[:, i])
        print(f"qadapter detector {i}: normalised feature error {e:.3e}")
        assert e <= 1e-4
    # adapter stage alone on the oracle's own spectrogram
    with torch.no_grad():
        spec = ref.q_transform(x[:, 1]).reshape(4, 512, 512)
        y = ref.freq_adapter(spec.unsqueeze(1))
        y = ref.final_pool(y).squeeze(1)
        y = (ref.scale * y + ref.bias) * ref.film_gamma[1] + ref.film_beta[1]
    got2 = ours.adapt(spec.cuda(), 1).cpu()
    e2 = _nerr(got2, y)
    print(f"qadapter CNN alone (tensor-core convolutions, bf16 hi/lo split): normalised error {e2:.3e}")
    assert e2 <= 4e-5


def test_qadapter_fp32_cuda_core_path_still_matches(tmp_path):
    """GWW_QADAPTER_TC=0 selects the round-1 fp32 CUDA-core convolutions (read once per process -> child process);
    they must agree with the oracle to fp32 accuracy and with the tensor-core path to its split precision."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = _seeded_adapter()
    sd = str(tmp_path / "adapter.pt")
    torch.save({k: v for k, v in ref.state_dict().items() if not k.startswith("q_transform.")}, sd)
    x = _strain(3, 2, seed=8)
    with torch.no_grad():
        spec = ref.q_transform(x[:, 1]).reshape(3, 512, 512)
        y = ref.freq_adapter(spec.unsqueeze(1))
        y = ref.final_pool(y).squeeze(1)
        want = (ref.scale * y + ref.bias) * ref.film_gamma[1] + ref.film_beta[1]
    torch.save(spec, str(tmp_path / "spec.pt"))
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from gw_whisper_b200 import QTransformAdapter\n"
        "a = QTransformAdapter(n_detectors=2); a.load_state_dict(torch.load(sys.argv[1]))\n"
        "torch.save(a.adapt(torch.load(sys.argv[2]).cuda(), 1).cpu(), sys.argv[3])\n")
    outs = {}
    for tc, env in (("0", {"GWW_QADAPTER_TC": "0"}), ("1", {"GWW_QADAPTER_TC": "1"}),
                    ("1-conv1-plain", {"GWW_QADAPTER_TC": "1", "GWW_QA_CONV1_TMA": "0"})):
        out = str(tmp_path / f"out{tc}.pt")
        subprocess.run([sys.executable, "-c", code, sd, str(tmp_path / "spec.pt"), out], check=True,
                       env={**os.environ, **env}, timeout=600)
        outs[tc] = torch.load(out)
    e0, e1, e01 = _nerr(outs["0"], want), _nerr(outs["1"], want), _nerr(outs["1"], outs["0"])
    print(f"adapter CNN vs oracle: fp32 CUDA-core path {e0:.3e}, tensor-core path {e1:.3e}; between them {e01:.3e}")
    assert e0 <= 5e-6 and e1 <= 4e-5 and e01 <= 4e-5
    # the persistent TMA-fed conv1 (default) and the one-CTA-per-tile conv1 run the same arithmetic in the same order
    assert torch.equal(outs["1"], outs["1-conv1-plain"])


def _reference_model(base, dora, adapter, num_classes=2, use_last_token=True):
    enc = E.attach_dora(base, dora)
    head = E.seeded_head(E.head_mlgwsc(base.config.d_model, 2, num_classes, softmax=True), seed=3, gain=3.0)

    def fwd(x):
        with torch.no_grad():
            feats = adapter(x)
            reps = []
            for i in range(feats.size(1)):
                seq = enc(feats[:, i]).last_hidden_state
                reps.append(seq[:, -1, :] if use_last_token else seq.mean(dim=1))
            return head(torch.cat(reps, dim=1)), head[:-1](torch.cat(reps, dim=1))
    return fwd, head


@pytest.mark.parametrize("use_last_token", [True, False])
def test_gwwhisper_classifier_logits(use_last_token):
    from gw_whisper_b200 import (B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter,
                                 remove_softmax_from_classifier)
    B = 4
    x = _strain(B, 2, seed=31)
    base = E.make_encoder("tiny", 0, spread=True)
    dora = E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    ref_adapter = _seeded_adapter()
    enc_b200 = B200WhisperEncoder.from_hf(base, dora=dora, chunk=2 * B)
    fwd, head = _reference_model(base, dora, ref_adapter, use_last_token=use_last_token)
    want_prob, want_logit = fwd(x)
    adapter = QTransformAdapter(n_detectors=2)
    adapter.load_state_dict(ref_adapter.state_dict())
    model = GWWhisperClassifier(enc_b200, 2, num_classes=2, q_adapter=adapter, use_last_token=use_last_token)
    model.classifier.load_state_dict(head.state_dict())
    got_prob = model(x.cuda()).cpu()
    e = (got_prob - want_prob).abs().max().item()
    print(f"GWWhisperClassifier softmax outputs (last_token={use_last_token}): max_abs_err {e:.3e} "
          f"spread {want_prob.std(0).mean().item():.3e}")
    assert e < 2e-2
    assert torch.allclose(got_prob.sum(1), torch.ones(B), atol=1e-5)
    remove_softmax_from_classifier(model)                           # USR mode
    got_logit = model(x.cuda()).cpu()
    e2 = (got_logit - want_logit).abs().max().item()
    print(f"GWWhisperClassifier USR logits: max_abs_err {e2:.3e} spread {want_logit.std(0).mean().item():.3e}")
    assert e2 < 2e-2


def test_stream_search_qscan_matches_batched_forward():
    """evaluate_slices over a segment == the model applied to the reference's 256-window batches;
    ragged last batch (here 1 window, the reference's torch.squeeze hazard)."""
    from gw_whisper_b200 import B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter
    from gw_whisper_b200 import inference as I
    hop, batch = 204, 4
    n_win = 2 * batch + 1
    g = torch.Generator().manual_seed(77)
    seg = torch.randn(2, 2048 + hop * (n_win - 1), generator=g)
    base = E.make_encoder("tiny", 0, spread=True)
    enc = B200WhisperEncoder.from_hf(base, chunk=2 * batch)
    adapter = QTransformAdapter(n_detectors=2)
    adapter.load_state_dict(_seeded_adapter().state_dict())
    model = GWWhisperClassifier(enc, 2, q_adapter=adapter)
    E.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    segc = seg.cuda()
    scores, tidx, tsc = model.stream_search(segc, hop, n_win, 0.5, batch=batch)
    want = []
    for k0 in range(0, n_win, batch):
        ks = range(k0, min(k0 + batch, n_win))
        xb = torch.stack([segc[:, k * hop: k * hop + 2048] for k in ks])      # [b, 2, 2048]
        want.append(model(xb)[:, 0])
    want = torch.cat(want)
    assert torch.allclose(scores, want, atol=1e-6, rtol=0)
    keep = (want > 0.5).nonzero().flatten()
    assert torch.equal(tidx, keep) and torch.allclose(tsc, want[keep], atol=1e-6, rtol=0)
    # through the reference-shaped driver
    f = I.ArrayFile.from_segments({"H1": {"100": seg[0].numpy()}, "L1": {"100": seg[1].numpy()}}, {"100": 100.0})
    slicer = I.TorchSegmentSlicer(f, "100", white=True)
    assert len(slicer) == n_win
