"""Size-independent properties at BASELINE.json's full size (configs[1]: whisper-base, 1024 windows x 2
detectors per step, chunks of 256 det-windows), where the fp32 oracle is too slow to be the checker:

  * windows are independent units of the log-mel path, so a window's logit must not
    depend on which batch / chunk / batch position it is computed in -- bit for bit;
  * the streaming search over a resident segment (device-side hop-204 window gather + compaction) must
    return exactly the scores of the explicitly cut windows, and its compacted trigger list must equal
    thresholding those scores on the host (indices in order, scores bit-equal);
  * linearity of the framing: window k of a segment is window 0 of the segment shifted by k * 204.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HOP = 204


def _model(chunk=256):
    from gw_whisper_b200 import B200WhisperEncoder, two_channel_ligo_binary_classifier
    from gw_whisper_b200 import synthetic as S
    base = S.make_encoder("base", 0, spread=True)
    dora = S.synthetic_dora("base", targets=("q_proj", "k_proj", "v_proj"))
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=chunk)
    model = two_channel_ligo_binary_classifier(enc, num_classes=1)
    S.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    return model


def test_full_batch_is_composition_invariant():
    dev = torch.device("cuda")
    model = _model(256)
    g = torch.Generator().manual_seed(1234)
    strain = torch.randn(1024, 2, 2048, generator=g).to(dev)
    full = model.forward_strain(strain)                      # 4 chunks of 256 det-windows... x2 detectors = 8
    assert full.shape == (1024, 1) and bool(torch.isfinite(full).all())
    assert full.std().item() > 1e-3, "spread-scaled weights should give logits that move with the input"
    # the same windows in another order, another batch size (ragged last chunk), another chunk size
    pick = torch.tensor([0, 1, 255, 256, 511, 777, 1023, 3, 640], device=dev)
    sub = model.forward_strain(strain[pick])
    assert torch.equal(sub, full[pick]), f"max diff {(sub - full[pick]).abs().max().item():.3e}"
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(7)).to(dev)
    shuffled = model.forward_strain(strain[perm][:300])      # 300 windows -> chunks of 256 + 44 det-windows... ragged
    assert torch.equal(shuffled, full[perm][:300])
    small_chunks = _model(64).forward_strain(strain[:128])
    assert torch.equal(small_chunks, full[:128])


def test_stream_search_equals_explicit_windows_and_host_threshold():
    from gw_whisper_b200.inference import LogMelStreamNetwork
    dev = torch.device("cuda")
    model = _model(256)
    n = 1024
    g = torch.Generator().manual_seed(99)
    seg = torch.randn(2, 2048 + HOP * (n - 1), generator=g).to(dev)
    windows = seg.unfold(1, 2048, HOP).permute(1, 0, 2).contiguous()      # [n, 2, 2048]
    assert windows.shape[0] == n
    ref = model.forward_strain(windows)[:, 0]
    thr = ref.median().item()
    scores, idx, sc = LogMelStreamNetwork(model).stream_search(seg, HOP, n, thr)
    assert torch.equal(scores, ref), f"max diff {(scores - ref).abs().max().item():.3e}"
    keep = torch.nonzero(ref > thr).flatten()
    assert torch.equal(idx, keep)                            # ordered compaction, same strict '>' as inference.py:477
    assert torch.equal(sc, ref[keep])
    assert 0 < keep.numel() < n
    # a sub-range of the same segment (what a time shard computes) gives the same scores
    s2, i2, _ = LogMelStreamNetwork(model).stream_search(seg, HOP, 300, thr, first_window=500)
    assert torch.equal(s2, ref[500:800])
    assert torch.equal(i2, keep[(keep >= 500) & (keep < 800)])   # trigger indices are global window indices
