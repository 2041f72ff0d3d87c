"""The reference's driver entry points end to end: `build_model` from the three weight
artefacts (PEFT adapter dir, dense .pth, Q-Adapter .pt -- MLGWSC-1/inference.py:415-434), `get_triggers` over an
HDF5 strain file in the MLGWSC-1 layout (:492-589) with and without whitening, and `main` with the reference's
flags writing `time/stat/var/all_vals` (:596-675).  Everything runs through the C ABI on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import encoder as E
from oracle import qscan as OQ
from oracle import whiten as W

pytestmark = pytest.mark.gpu
FS = 2048


def _artefacts(tmp):
    """Writes the three artefacts the reference loads + a base-encoder state_dict (no checkpoint download offline)."""
    from safetensors.numpy import save_file
    base = E.make_encoder("tiny", 0, spread=True)
    torch.save(base.state_dict(), os.path.join(tmp, "whisper-tiny-encoder.pt"))
    dora = E.synthetic_dora("tiny", targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    lora_dir = os.path.join(tmp, "best_lora_weights")
    os.makedirs(lora_dir)
    save_file({k: np.ascontiguousarray(v, dtype=np.float32) for k, v in dora["tensors"].items()},
              os.path.join(lora_dir, "adapter_model.safetensors"))
    with open(os.path.join(lora_dir, "adapter_config.json"), "w") as fh:
        json.dump({"r": dora["r"], "lora_alpha": dora["lora_alpha"], "use_dora": True, "peft_type": "LORA",
                   "target_modules": ["q_proj", "k_proj", "v_proj", "out_proj"]}, fh)
    torch.manual_seed(11)
    adapter = OQ.QTransformAdapter(n_detectors=2)
    torch.save(adapter.state_dict(), os.path.join(tmp, "adapter.pt"))          # includes q_transform.* buffers
    head = E.seeded_head(E.head_mlgwsc(384, 2, 2, softmax=True), seed=3, gain=3.0)
    torch.save(head.state_dict(), os.path.join(tmp, "dense.pth"))
    return base, dora, adapter, head, lora_dir


def _strain_file(path, whitened: bool):
    from gw_whisper_b200 import hdf5io as H
    segs = {}
    with H.File(path, "w") as f:
        for det, off in (("H1", 0), ("L1", 100)):
            g = f.create_group(det)
            for st, n, seed in ((1238166018, 6 * FS, 1), (1238170000, 9 * FS, 2)):
                x = np.random.default_rng(seed + off).standard_normal(n) if whitened else W.colored_noise(n, seed + off) * 1e20
                ds = g.create_dataset(str(st), data=x, compression="gzip", compression_opts=4, shuffle=True)
                ds.attrs["start_time"] = float(st)
                ds.attrs["delta_t"] = 1.0 / FS
                segs[(det, str(st))] = x
    return segs


def test_get_triggers_and_main_from_artefacts(tmp_path, monkeypatch):
    from gw_whisper_b200 import hdf5io as H
    from gw_whisper_b200 import inference as I
    tmp = str(tmp_path)
    base, dora, adapter, head, lora_dir = _artefacts(tmp)
    monkeypatch.setenv(I.WHISPER_BASE_ENV, os.path.join(tmp, "whisper-tiny-encoder.pt"))
    infile = os.path.join(tmp, "white.hdf")
    segs = _strain_file(infile, whitened=True)
    thr = 0.0
    triggers, all_vals = I.get_triggers(lora_dir, os.path.join(tmp, "dense.pth"), os.path.join(tmp, "adapter.pt"),
                                        infile, trigger_threshold=thr, white=True, usr=True)
    assert list(triggers.keys()) == ["1238166018", "1238170000"]              # sorted by key (inference.py:589)
    n_long, n_short = 1 + (9 * FS - 2048) // 204, 1 + (6 * FS - 2048) // 204
    assert sum(len(v) for v in all_vals) == n_long + n_short
    assert len(all_vals[0]) == min(256, n_long)                                # longest segment first (:546)
    # spot-check one batch against the fp32 oracle model on the same windows (first 6 windows of the long segment)
    x = np.stack([segs[("H1", "1238170000")], segs[("L1", "1238170000")]])
    nb = min(256, n_long)
    win = torch.from_numpy(np.stack([x[:, k * 204:k * 204 + 2048] for k in range(nb)]).astype(np.float32))
    enc = E.attach_dora(base, dora)
    with torch.no_grad():
        feats = adapter.eval()(win)                                           # QScan plane choice needs the whole batch
        reps = torch.cat([enc(feats[:6, i]).last_hidden_state[:, -1, :] for i in range(2)], dim=1)
        want = head[:-1](reps)[:, 0].numpy()
    got = all_vals[0][:6]
    print("get_triggers scores vs fp32 oracle (6 windows):", np.abs(got - want).max())
    assert np.abs(got - want).max() < 2e-2
    # trigger times are float64 and follow start + k*204/2048 + 0.6
    for key, n in (("1238170000", n_long), ("1238166018", n_short)):
        for t, s in triggers[key]:
            k = round((t - float(key) - 0.6) / (204 / 2048))
            assert abs(t - (float(key) + 0.6 + k * 204 / 2048)) < 1e-6 and 0 <= k < n and s > thr
    # --- main(): same flags as the reference CLI, HDF5 outputs
    out = os.path.join(tmp, "out.hdf")
    dbg = os.path.join(tmp, "dbg.hdf")
    argv = [infile, out, "--white", "--lora-weights", lora_dir, "--dense-weights", os.path.join(tmp, "dense.pth"),
            "--adapter-weights", os.path.join(tmp, "adapter.pt"), "-t", str(thr), "--debug-triggers-file", dbg]
    I.main(argv)
    with H.File(out) as f:
        t, s, v, av = f["time"][()], f["stat"][()], f["var"][()], f["all_vals"][()]
    tt, ss, vv = I.get_clusters(triggers, 0.35)
    assert np.array_equal(t, tt) and np.array_equal(s, ss) and np.array_equal(v, vv)
    assert av.dtype == np.float32 and np.array_equal(av, np.concatenate(all_vals).astype(np.float32))
    with H.File(dbg) as f:
        assert sorted(f.keys()) == ["1238166018", "1238170000"]
    with pytest.raises(RuntimeError):
        I.main(argv)                                                          # output exists, no --force (:627-632)
    I.main(argv + ["--force"])


def test_get_triggers_whitens_unwhitened_input(tmp_path, monkeypatch):
    from gw_whisper_b200 import hdf5io as H
    from gw_whisper_b200 import inference as I
    tmp = str(tmp_path)
    base, dora, adapter, head, lora_dir = _artefacts(tmp)
    monkeypatch.setenv(I.WHISPER_BASE_ENV, os.path.join(tmp, "whisper-tiny-encoder.pt"))
    infile = os.path.join(tmp, "raw.hdf")
    segs = _strain_file(infile, whitened=False)
    wfile = os.path.join(tmp, "whitened.hdf")
    network = I.build_model(lora_dir, os.path.join(tmp, "dense.pth"), os.path.join(tmp, "adapter.pt"), "cuda", usr=True)
    trig_raw, vals_raw = I.get_triggers(None, None, None, infile, trigger_threshold=0.0, white=False,
                                        whitened_file=wfile, low_frequency_cutoff=20.0, network=network)
    # the whitened debug file holds what the oracle's whiten produces, and searching it with --white gives the
    # same scores (the 0.125 s start shift only moves the trigger times)
    with H.File(wfile) as f:
        for (det, key), x in segs.items():
            got = f[det][key][()]
            want = W.whiten(x, low_frequency_cutoff=20.0)
            assert got.shape == want.shape
            assert np.abs(got - want).max() / np.sqrt(np.mean(want ** 2)) < 1e-4
    # rebuild a strain file from the whitened data and run with white=True
    wfile2 = os.path.join(tmp, "whitened_in.hdf")
    with H.File(wfile) as f, H.File(wfile2, "w") as o:
        for det in ("H1", "L1"):
            g = o.create_group(det)
            for key in f[det].keys():
                ds = g.create_dataset(key, data=f[det][key][()])
                ds.attrs["start_time"] = float(key) + 0.125
                ds.attrs["delta_t"] = 1.0 / FS
    trig_w, vals_w = I.get_triggers(None, None, None, wfile2, trigger_threshold=0.0, white=True, network=network)
    assert np.array_equal(np.concatenate(vals_raw), np.concatenate(vals_w))
    for key in trig_raw:
        assert np.allclose(np.array(trig_raw[key]), np.array(trig_w[key]), rtol=0, atol=1e-6)
