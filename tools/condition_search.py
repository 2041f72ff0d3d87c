"""Offline search (run on the GPU box) for a *conditioned* random-init weight set per Whisper size:
multipliers on (q,k), (other layer matrices), (conv stem) such that PyTorch's own bf16 autocast of the
fp32 reference stays within ~1e-2 of fp32 on the logits while the logits move >= 10x that across windows
(VERDICT r1 item 1b).  Prints one JSON line per candidate; the chosen table goes to
gw_whisper_b200/synthetic.py::CONDITIONED.

    python tools/condition_search.py small 48
"""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "small"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    from gw_whisper_b200 import B200WhisperEncoder, logmel_features
    from gw_whisper_b200 import synthetic as S
    from oracle import encoder as E

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(4321)
    strain = torch.randn(n, 2048, generator=g).to(dev)
    feats = logmel_features(strain)
    d = S.SIZES[size]["d_model"]
    head = E.seeded_head(E.head_glitch(d, 11), seed=5, gain=3.0).to(dev)
    grid = list(itertools.product([1.0, 3.0, 4.0, 6.0, 8.0], [1.0, 2.0, 3.0], [3.0]))
    if os.environ.get("COND_GRID") == "full":
        grid = list(itertools.product([1.0, 2.0, 3.0, 4.0, 6.0], [1.0, 1.5, 2.0, 3.0], [1.0, 3.0]))
    for qk, lay, conv in grid:
        enc = S.make_encoder(size, 0, spread=False)
        S.scale_encoder_(enc, qk, lay, conv)
        enc = enc.to(dev)
        with torch.no_grad():
            ref = torch.cat([enc(feats[i:i + 16]).last_hidden_state[:, -1] for i in range(0, n, 16)])
            with torch.autocast("cuda", dtype=torch.bfloat16):
                b16 = torch.cat([enc(feats[i:i + 16]).last_hidden_state[:, -1] for i in range(0, n, 16)]).float()
            with torch.autocast("cuda", dtype=torch.float16):
                h16 = torch.cat([enc(feats[i:i + 16]).last_hidden_state[:, -1] for i in range(0, n, 16)]).float()
            ours = B200WhisperEncoder.from_hf(enc.cpu(), chunk=min(n, 64)).pooled(feats)
            lr, lb, lo, lh = head(ref), head(b16), head(ours), head(h16)
        rec = {"size": size, "qk": qk, "layer": lay, "conv": conv,
               "rep_spread": ref.std(0).mean().item(),
               "rep_err_bf16": (b16 - ref).abs().max().item(), "rep_err_ours": (ours - ref).abs().max().item(),
               "logit_spread": lr.std(0).mean().item(),
               "logit_err_bf16": (lb - lr).abs().max().item(), "logit_err_ours": (lo - lr).abs().max().item(),
               "logit_mean_err_ours": (lo - lr).abs().mean().item(),
               "rep_err_f16": (h16 - ref).abs().max().item(), "logit_err_f16": (lh - lr).abs().max().item()}
        rec["ratio_ours"] = rec["logit_spread"] / max(rec["logit_err_ours"], 1e-12)
        print(json.dumps(rec), flush=True)
        del enc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
