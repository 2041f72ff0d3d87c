"""Kernel-level parity on a real B200, through the C ABI (ctypes), against fp32 PyTorch references
of the same op (floating-point kernels) on identical seeded inputs."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch.device("cuda")


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _report(name, got, ref, tol_abs, tol_rel):
    got, ref = got.double(), ref.double()
    err = (got - ref).abs()
    lim = tol_abs + tol_rel * ref.abs()
    bad = err > lim
    idx = int(err.argmax())
    msg = (f"{name}: max_abs_err={err.max().item():.4e} at flat {idx} got={got.flatten()[idx].item():.5f} "
           f"ref={ref.flatten()[idx].item():.5f} n_bad={int(bad.sum())}/{bad.numel()} "
           f"ref_rms={ref.pow(2).mean().sqrt().item():.4e}")
    print(msg)
    assert not bool(bad.any()), msg


GEMM_CASES = [
    # M, N, K, epilogue, block_n
    (128, 128, 64, 0, 128),
    (256, 256, 128, 0, 256),
    (1000, 384, 384, 0, 192),
    (1000, 384, 384, 0, 128),
    (3000, 1152, 384, 0, 192),
    (4096, 2048, 512, 1, 256),
    (3001, 512, 2048, 2, 256),
    (1500, 384, 1152, 3, 192),
    (20000, 1536, 512, 0, 256),
    (777, 768, 3072, 2, 256),
    (5, 512, 512, 2, 256),        # tiny-M GEMMs of the last-token-only final layer
    (8, 2048, 512, 1, 256),
    (256, 512, 2048, 2, 256),
]


@pytest.mark.parametrize("M,N,K,epi,bn", GEMM_CASES)
def test_gemm_tcgen05(lib, M, N, K, epi, bn):
    from gw_whisper_b200 import _lib
    dev = _cuda()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    A = torch.randn(M, K, generator=g).to(dev).to(_lib.operand_dtype())
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).to(_lib.operand_dtype())
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev) if epi == 2 else None
    pos = torch.randn(M, N, generator=g).to(dev) if epi == 3 else None
    acc = A.float() @ W.float().t() + bias
    if epi == 0:
        ref = acc
    elif epi == 1:
        ref = _gelu(acc)
    elif epi == 2:
        ref = resid + acc
    else:
        ref = _gelu(acc) + pos
    out_f32 = epi in (2, 3)
    C = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32 if out_f32 else _lib.operand_dtype())
    rc = lib.gww_gemm_bf16(A.data_ptr(), W.data_ptr(), C.data_ptr(), bias.data_ptr(), _lib.ptr(resid),
                           _lib.ptr(pos), M, N, K, epi, bn, _lib.stream_ptr())
    _lib.check(rc)
    torch.cuda.synchronize()
    if out_f32:
        _report(f"gemm{(M, N, K, epi, bn)}", C, ref, 2e-4, 2e-5)
    else:
        _report(f"gemm{(M, N, K, epi, bn)}", C.float(), ref, 1e-2, 1e-2)   # bf16 output rounding


def test_gemm_inplace_residual(lib):
    """out_proj / fc2 write the residual stream in place (resid == C)."""
    from gw_whisper_b200 import _lib
    dev = _cuda()
    g = torch.Generator().manual_seed(5)
    M, N, K = 2500, 512, 512
    A = torch.randn(M, K, generator=g).to(dev).to(_lib.operand_dtype())
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).to(_lib.operand_dtype())
    bias = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    ref = x + A.float() @ W.float().t() + bias
    _lib.check(lib.gww_gemm_bf16(A.data_ptr(), W.data_ptr(), x.data_ptr(), bias.data_ptr(), x.data_ptr(),
                                 None, M, N, K, 2, 256, _lib.stream_ptr()))
    torch.cuda.synchronize()
    _report("gemm_inplace", x, ref, 2e-4, 2e-5)


@pytest.mark.parametrize("d", [384, 512, 768])
@pytest.mark.parametrize("out_bf16", [1, 0])
def test_layernorm(lib, d, out_bf16):
    from gw_whisper_b200 import _lib
    dev = _cuda()
    g = torch.Generator().manual_seed(d)
    rows = 1003
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(dev)
    gamma = torch.randn(d, generator=g).to(dev)
    beta = torch.randn(d, generator=g).to(dev)
    ref = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    out = torch.empty(rows, d, device=dev, dtype=_lib.operand_dtype() if out_bf16 else torch.float32)
    _lib.check(lib.gww_layernorm(x.data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, d,
                                 out_bf16, _lib.stream_ptr()))
    torch.cuda.synchronize()
    if out_bf16:
        _report(f"ln{d}", out.float(), ref, 1e-2, 1e-2)
    else:
        _report(f"ln{d}", out, ref, 2e-5, 2e-5)


def _attn_ref(qkv, d):
    n, T, _ = qkv.shape
    H = d // 64
    q, k, v = qkv.float().split(d, dim=2)
    q = q.view(n, T, H, 64).transpose(1, 2)
    k = k.view(n, T, H, 64).transpose(1, 2)
    v = v.view(n, T, H, 64).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (p @ v).transpose(1, 2).reshape(n, T, d)


@pytest.mark.parametrize("n,T,d,scale", [(1, 256, 64, 1.0), (2, 1500, 384, 1.0), (3, 1500, 512, 4.0),
                                          (1, 700, 128, 12.0)])
def test_attention_tcgen05(lib, n, T, d, scale):
    """softmax(QK^T)V per head; `scale` sharpens the scores (large scale exercises the lazy
    rescale path and peaky rows)."""
    from gw_whisper_b200 import _lib
    dev = _cuda()
    g = torch.Generator().manual_seed(n * 100 + T + d)
    qkv = torch.randn(n, T, 3 * d, generator=g)
    qkv[:, :, :d] *= scale / 8.0
    qkv = qkv.to(dev).to(_lib.operand_dtype())
    # make later keys progressively larger for one case so the running max keeps growing
    ref = _attn_ref(qkv, d)
    out = torch.full((n, T, d), float("nan"), device=dev, dtype=_lib.operand_dtype())
    _lib.check(lib.gww_attention(qkv.data_ptr(), out.data_ptr(), n, T, d, _lib.stream_ptr()))
    torch.cuda.synchronize()
    _report(f"attn{(n, T, d, scale)}", out.float(), ref, 2e-2, 2e-2)


def test_attention_growing_max(lib):
    """Keys whose scores grow along the sequence force repeated rescales of O in TMEM."""
    from gw_whisper_b200 import _lib
    dev = _cuda()
    n, T, d = 1, 1500, 64
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(n, T, 3 * d, generator=g)
    ramp = torch.linspace(0.2, 6.0, T).view(1, T, 1)
    qkv[:, :, d:2 * d] = qkv[:, :, d:2 * d].abs() * ramp        # k grows with position
    qkv[:, :, :d] = qkv[:, :, :d].abs() * 0.5                    # q positive -> scores grow
    qkv = qkv.to(dev).to(_lib.operand_dtype())
    ref = _attn_ref(qkv, d)
    out = torch.full((n, T, d), float("nan"), device=dev, dtype=_lib.operand_dtype())
    _lib.check(lib.gww_attention(qkv.data_ptr(), out.data_ptr(), n, T, d, _lib.stream_ptr()))
    torch.cuda.synchronize()
    _report("attn_growing", out.float(), ref, 2e-2, 2e-2)


def test_logmel_frontend_vs_oracle(lib):
    """Front-end gate from BASELINE.json north_star: features within 1e-4 (scale-normalised,
    max|a-b|/max|b|) of the reference CPU path on identical Gaussian-noise strain."""
    from oracle import logmel as L
    from gw_whisper_b200 import logmel_features
    dev = _cuda()
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((5, 2048)).astype(np.float32)
    x[3] *= 1e-3                      # quiet window
    x[4] = 0.0                        # silent window: everything sits on the 1e-10 floor
    ref = L.logmel_restated(x)
    got = logmel_features(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert got.shape == (5, 80, 3000)
    for i in range(5):
        e = L.feature_error(got[i], ref[i])
        print(f"logmel window {i}: normalised err {e:.3e}  tail const {np.ptp(got[i][:, 102:]):.1e}")
        assert e <= 1e-4
    assert np.array_equal(got[:, :, 102:], np.broadcast_to(got[:, :, 102:103], got[:, :, 102:].shape))


def test_resample_and_logmel_from_16k_match_the_reference_calls(lib):
    """The two halves of front end A as the reference's datasets use them: the stored audio is
    scipy.signal.resample(x, 16000) (preprocess.py:44-51, f32 at :95) and the features are
    WhisperFeatureExtractor(audio, sampling_rate=16000) per item (dataset.py:20-24)."""
    from scipy.signal import resample
    from transformers import WhisperFeatureExtractor
    from oracle import logmel as L
    from gw_whisper_b200 import LogMelFeatureExtractor, logmel_features, resample_timeseries
    dev = _cuda()
    rng = np.random.default_rng(77)
    x = rng.standard_normal((4, 2048)).astype(np.float32)
    x[2] *= 1e-3
    want_audio = np.stack([resample(xi.astype(np.float64), 16000) for xi in x]).astype(np.float32)
    got_audio = resample_timeseries(x)                       # numpy in -> numpy out, like the reference function
    assert got_audio.shape == (4, 16000) and got_audio.dtype == np.float32
    e_a = np.abs(got_audio - want_audio).max() / np.abs(want_audio).max()
    print(f"resample_timeseries vs scipy.signal.resample: normalised err {e_a:.3e}")
    assert e_a <= 2e-7                                       # f32 storage rounding
    t_audio = resample_timeseries(torch.from_numpy(x).to(dev))
    assert t_audio.is_cuda and np.array_equal(t_audio.cpu().numpy(), got_audio)
    # features from the REFERENCE's stored audio through the HF extractor vs ours from the same audio
    hf = WhisperFeatureExtractor()
    want = np.stack([hf(a, sampling_rate=16000, return_tensors="np").input_features[0] for a in want_audio])
    fe = LogMelFeatureExtractor.from_pretrained("openai/whisper-tiny")
    got = fe(want_audio, sampling_rate=16000, return_tensors="pt").input_features
    assert got.shape == (4, 80, 3000) and got.is_cuda
    for i in range(4):
        e = L.feature_error(got[i].cpu().numpy(), want[i])
        print(f"logmel_from_16k window {i}: normalised err vs HF WhisperFeatureExtractor {e:.3e}")
        assert e <= 1e-4
    one = fe(want_audio[0], sampling_rate=16000).input_features   # 1-D item, as Dataset.__getitem__ passes it
    assert one.shape == (1, 80, 3000) and torch.equal(one[0], got[0])
    # fused kernel == the two halves chained (same arithmetic, same f32 rounding of the audio)
    fused = logmel_features(torch.from_numpy(x).to(dev))
    chained = fe(torch.from_numpy(got_audio).to(dev), sampling_rate=16000).input_features
    assert torch.equal(fused, chained)
    with pytest.raises(ValueError):
        fe(want_audio, sampling_rate=8000)
    with pytest.raises(ValueError):
        fe(np.zeros(12345, dtype=np.float32), sampling_rate=16000)


def test_head_and_compaction(lib):
    from gw_whisper_b200 import _lib
    from gw_whisper_b200.encoder import B200WhisperEncoder
    from oracle import encoder as E
    dev = _cuda()
    enc_ref = E.make_encoder("tiny", 0)
    enc = B200WhisperEncoder.from_hf(enc_ref, chunk=4)
    for head, B in ((E.head_two_channel(384, 1), 37), (E.head_mlgwsc(384, 2, 2, True), 9),
                    (E.head_glitch(384, 11), 64), (E.head_one_channel(384, 2, True), 5)):
        head = E.seeded_head(head, gain=2.0)
        lin = [(m.weight, m.bias) for m in head if isinstance(m, torch.nn.Linear)]
        enc.set_head(lin, softmax=any(isinstance(m, torch.nn.Softmax) for m in head))
        x = torch.randn(B, lin[0][0].shape[1], generator=torch.Generator().manual_seed(B))
        ref = head(x)
        got = enc.head(x.to(dev))
        _report(f"head B={B}", got.cpu(), ref.detach(), 1e-5, 1e-5)
    # ordered compaction (strict >, window order)
    n, Cc = 5000, 2
    out = torch.rand(n, Cc, generator=torch.Generator().manual_seed(1)).to(dev)
    thr = 0.9
    idx = torch.zeros(n, dtype=torch.long, device=dev)
    sc = torch.zeros(n, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.gww_threshold_compact(out.data_ptr(), Cc, n, thr, 1000, idx.data_ptr(), sc.data_ptr(),
                                         cnt.data_ptr(), n, _lib.stream_ptr()))
    torch.cuda.synchronize()
    keep = (out[:, 0] > thr).nonzero().flatten()
    c = int(cnt.item())
    assert c == keep.numel()
    assert torch.equal(idx[:c], keep + 1000)
    assert torch.equal(sc[:c], out[keep, 0])
