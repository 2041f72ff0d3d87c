"""Shared helpers of the GPU parity tests: the fp32 oracle run on the GPU in eager fp32 (TF32 off), error
statistics, and the trigger / argmax agreement report with its stated guard band."""
import torch

LOGIT_TOL = 2e-2          # BASELINE.json north_star: encoder logits within 2e-2 absolute (16-bit operands vs fp32)
GUARD = LOGIT_TOL         # guard band around a threshold: a window is "decisive" when |oracle score - thr| > GUARD


def fp32_strict():
    """The oracle runs in true fp32: no TF32 in matmuls or cuDNN convolutions."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def batched(fn, *tensors, bs=32):
    outs = []
    with torch.no_grad():
        for i in range(0, tensors[0].shape[0], bs):
            outs.append(fn(*[t[i:i + bs] for t in tensors]))
    return torch.cat(outs)


def stats(name, got, ref):
    err = (got - ref).abs()
    spread = ref.std(0).mean().item() if ref.shape[0] > 1 else float("nan")
    e = err.max().item()
    print(f"{name}: n={ref.shape[0]} max_abs_err={e:.4e} mean_abs_err={err.mean().item():.4e} "
          f"oracle_spread(std over windows)={spread:.4e} max_err/spread={e / spread if spread == spread else float('nan'):.4f}")
    return e, spread


def search_threshold(ref_scores: torch.Tensor, top_frac: float = 0.05) -> float:
    """A search places its threshold in the upper tail: midpoint of the widest gap between consecutive sorted
    oracle scores within the top `top_frac` of windows (deterministic in the oracle scores alone)."""
    s = torch.sort(ref_scores.flatten().double(), descending=True).values
    k = max(3, int(round(top_frac * s.numel())))
    gaps = s[:k - 1] - s[1:k]
    i = int(torch.argmax(gaps))
    return float((s[i] + s[i + 1]) / 2)


def trigger_agreement(name, got_scores, ref_scores, thr, guard=GUARD):
    """Thresholded decisions `score > thr` (inference.py:484) of ours vs the oracle.  Returns (overall agreement
    fraction, number of decisive windows, number of marginal windows, decisive disagreements)."""
    got_scores, ref_scores = got_scores.flatten(), ref_scores.flatten()
    dec = (ref_scores - thr).abs() > guard
    same = (got_scores > thr) == (ref_scores > thr)
    n = ref_scores.numel()
    agree = same.float().mean().item()
    bad_dec = int((~same & dec).sum())
    print(f"{name}: thr={thr:+.4f} triggers(oracle)={int((ref_scores > thr).sum())} agreement={100 * agree:.3f}% of {n} "
          f"windows; decisive={int(dec.sum())} marginal(|score-thr|<={guard:g})={int((~dec).sum())} "
          f"decisive disagreements={bad_dec} marginal disagreements={int((~same & ~dec).sum())}")
    return agree, int(dec.sum()), int((~dec).sum()), bad_dec


def argmax_agreement(name, got, ref, guard=GUARD):
    top2 = ref.topk(2, dim=1).values
    dec = (top2[:, 0] - top2[:, 1]) > 2 * guard          # both logits may move by the tolerance
    same = got.argmax(1) == ref.argmax(1)
    agree = same.float().mean().item()
    bad_dec = int((~same & dec).sum())
    print(f"{name}: argmax agreement={100 * agree:.3f}% of {ref.shape[0]} windows; decisive(margin>{2 * guard:g})={int(dec.sum())} "
          f"marginal={int((~dec).sum())} decisive disagreements={bad_dec} marginal disagreements={int((~same & ~dec).sum())}")
    return agree, int(dec.sum()), int((~dec).sum()), bad_dec
