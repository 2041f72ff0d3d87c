"""Generates the golden vectors under tests/golden/ by running the REFERENCE code in this container
(/root/reference on sys.path + the installed transformers/scipy).  Committed together with its
outputs; re-run only where /root/reference exists:

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle import encoder as E, logmel as L
    from scipy.signal import resample
    from transformers import WhisperFeatureExtractor

    rng = np.random.default_rng(20261018)
    x = rng.standard_normal((3, 2048)).astype(np.float32)
    x[2] *= 0.05
    # --- front end A exactly as the reference runs it: preprocess.py:44-51 then dataset.py:20-24
    fe = WhisperFeatureExtractor()
    # strain files written by pycbc are float64 (the dtype scipy.signal.resample then computes in)
    y = np.stack([resample(xi.astype(np.float64), len(xi) * 16000 // 2048) for xi in x]).astype(np.float32)
    padded = np.zeros((3, fe.n_samples), np.float32)
    padded[:, :16000] = y
    feats_np = fe._np_extract_fbank_features(padded, "cpu").astype(np.float32)       # transformers 4.37 path
    feats_pt = np.stack([fe(a, sampling_rate=16000, return_tensors="np").input_features[0] for a in y])
    np.savez_compressed(os.path.join(HERE, "logmel_golden.npz"), strain=x,
                        resampled_head=y[:, :512], feats_np_head=feats_np[:, :, :104],
                        feats_np_tail=feats_np[:, :, -1], feats_pt_head=feats_pt[:, :, :104].astype(np.float32))

    # --- reference model classes (imported from the reference tree) on a seeded tiny encoder
    svn = _load(os.path.join(REF, "Signal_vs_Noise/src/model.py"), "ref_svn_model")
    gl = _load(os.path.join(REF, "Glitch_classification/src/model.py"), "ref_glitch_model")
    enc = E.make_encoder("tiny", 0, spread=True)
    feats = torch.from_numpy(feats_np[:2])
    torch.manual_seed(11)
    m2 = svn.two_channel_ligo_binary_classifier(enc, 1).eval()
    E.seeded_head(m2.classifier, seed=3, gain=3.0)
    m1 = svn.one_channel_ligo_binary_classifier(enc, 1).eval()
    E.seeded_head(m1.classifier, seed=4, gain=3.0)
    mg = gl.one_channel_ligo_binary_classifier(enc, 11).eval()
    E.seeded_head(mg.classifier, seed=5, gain=3.0)
    with torch.no_grad():
        hs = enc(feats).last_hidden_state
        out2 = m2(feats, feats.flip(0))
        out1 = m1(feats)
        outg = mg(feats)
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), last_token=hs[:, -1].numpy(),
                        token0=hs[:, 0].numpy(), two_channel=out2.numpy(), one_channel=out1.numpy(),
                        glitch=outg.numpy())

    # --- shipped single-detector dense head (real trained weights) on seeded inputs
    sd = torch.load(os.path.join(REF, "Signal_vs_Noise/results/Single_detector/models/best_dense_layers.pth"),
                    map_location="cpu")
    head = E.head_one_channel(384, 1)
    head.load_state_dict(sd)
    xin = torch.randn(4, 384, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        yout = head(xin)
    np.savez_compressed(os.path.join(HERE, "shipped_head_golden.npz"), x=xin.numpy(), y=yout.numpy())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
