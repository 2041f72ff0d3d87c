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


def test_sharded_search_world1_and_emulated_ranks_equal_the_single_stream():
    """sharding.sharded_search on the GPU: with world=1 it must return exactly what one
    stream_search per segment returns; and the union of the pieces every rank of a 3-rank plan would compute
    (each piece handed only its sample range + halo) must reproduce the same scores bit for bit --
    including ragged segments and, for the MLGWSC-1 model, the batch-coupled QScan plane choice (whole
    256-window batches per shard)."""
    from gw_whisper_b200 import B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter, sharding
    from gw_whisper_b200 import synthetic as S
    dev = torch.device("cuda")
    base = S.make_encoder("tiny", 0, spread=True)
    enc = B200WhisperEncoder.from_hf(base, dora=S.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj")),
                                     chunk=512)
    torch.manual_seed(11)
    model = GWWhisperClassifier(enc, 2, num_classes=2, q_adapter=QTransformAdapter(n_detectors=2))
    S.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    g = torch.Generator().manual_seed(5)
    lens = [2048 + HOP * 700, 2048 + HOP * 255 + 17, 2048 + HOP * 256, 3000, 1000]      # ragged, one too short
    segs = [torch.randn(2, n, generator=g).to(dev) for n in lens]
    nws = [sharding.n_windows(n, HOP) for n in lens]
    thr = 0.5
    (seg, idx, sc), scores = sharding.sharded_search(model, segs, HOP, thr, 0, 1)
    assert [int(s.numel()) for s in scores] == nws
    for i, s in enumerate(segs):
        if nws[i] == 0:
            continue
        ref, ridx, rsc = model.stream_search(s, HOP, nws[i], thr)
        assert torch.equal(scores[i], ref)
        assert torch.equal(idx[seg == i], ridx) and torch.equal(sc[seg == i], rsc)
    # emulate a 3-rank run in this process: every rank's pieces, merged by hand
    plan = sharding.plan_shards(nws, 3)
    assert all(len(p) > 0 for p in plan)
    merged = [torch.full((n,), float("nan"), device=dev) for n in nws]
    for pieces in plan:
        for p in pieces:
            lo, hi = p.sample_range(HOP)
            part, _, _ = model.stream_search(segs[p.segment][:, lo:hi].contiguous(), HOP, p.n_windows, thr)
            merged[p.segment][p.first_window:p.first_window + p.n_windows] = part
    for a, b in zip(merged, scores):
        assert torch.equal(a, b)
