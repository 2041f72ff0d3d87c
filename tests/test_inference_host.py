"""Host-side logic of the sliding-window driver against the reference's own functions
(MLGWSC-1/inference.py imported with its missing third-party modules stubbed)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
have_ref = os.path.isdir(REF)


def _load_reference_inference():
    """Import MLGWSC-1/inference.py with h5py / peft / pycbc / ml4gw stubbed."""
    stubs = {}
    for name in ("h5py", "peft", "ml4gw", "ml4gw.transforms", "pycbc", "pycbc.waveform", "pycbc.noise",
                 "pycbc.psd", "pycbc.distributions", "pycbc.detector", "pycbc.types", "pycbc.filter"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
    stubs.get("h5py", sys.modules.get("h5py")).File = object
    if "peft" in stubs:
        stubs["peft"].PeftModel = object
    if "ml4gw.transforms" in stubs:
        stubs["ml4gw.transforms"].QScan = object
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_inference", os.path.join(REF, "MLGWSC-1/inference.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in stubs:
            sys.modules.pop(k, None)
    return mod


def _rand_triggers(seed, n_keys=3):
    rng = np.random.default_rng(seed)
    trig = {}
    for k in range(n_keys):
        n = int(rng.integers(0, 60))
        t = np.sort(rng.uniform(0, 30, n)) + 1000.0 * k
        if n > 4:
            t[3] = t[2] + 0.35            # gap exactly at the threshold: stays in the cluster (strict >)
            t = np.sort(t)
        s = rng.uniform(0, 1, n)
        if n > 6:
            s[5] = s[4]                   # tie inside a cluster: first one wins
        trig[str(k)] = [[float(a), float(b)] for a, b in zip(t, s)]
    return trig


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_get_clusters_matches_reference(seed):
    from gw_whisper_b200.inference import get_clusters
    ref = _load_reference_inference()
    trig = _rand_triggers(seed)
    a = ref.get_clusters(trig, 0.35)
    b = get_clusters(trig, 0.35)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y))


def test_get_clusters_edge_cases():
    from gw_whisper_b200.inference import get_clusters
    t, s, v = get_clusters({}, 0.35)
    assert len(t) == len(s) == len(v) == 0
    t, s, v = get_clusters({"a": [], "b": [[1.0, 0.3]]}, 0.35)
    assert t.tolist() == [1.0] and s.tolist() == [0.3] and v.tolist() == [0.2]
    t, s, v = get_clusters({"a": [[0.0, 0.1], [0.3, 0.9], [0.6, 0.5], [1.0, 0.2]]}, 0.35)
    assert t.tolist() == [0.3, 1.0] and s.tolist() == [0.9, 0.2]


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
@pytest.mark.parametrize("n_samples", [2048, 2048 + 203, 2048 + 204, 50000, 123457])
def test_segment_slicer_matches_reference(n_samples):
    from gw_whisper_b200.inference import ArrayFile, SegmentSlicer
    ref = _load_reference_inference()
    rng = np.random.default_rng(n_samples)
    h1 = rng.standard_normal(n_samples)
    l1 = rng.standard_normal(n_samples)
    start = 1238.25
    f = ArrayFile.from_segments({"H1": {"1238": h1}, "L1": {"1238": l1}}, {"1238": start})
    ours = SegmentSlicer(f, "1238", white=True)
    theirs = ref.SegmentSlicer(f, "1238", white=True)
    assert len(ours) == len(theirs)
    assert ours.index_step_size == theirs.index_step_size == 204
    ref_times, ref_first = [], None
    for i, (sl, ts) in enumerate(iter(theirs)):
        ref_times.append(ts)
        if i == 0:
            ref_first = sl.copy()
    assert len(ref_times) == len(ours)
    assert np.array_equal(ours.window_times(reference_float32=False), np.array(ref_times))
    # what the reference's DataLoader hands to the trigger loop: float32 times (H8)
    f32 = np.array([torch.tensor(t).item() for t in ref_times])
    assert np.array_equal(ours.window_times(reference_float32=True), f32)
    sl0, t0 = next(iter(ours))
    assert np.array_equal(sl0, ref_first) and t0 == ref_times[0]
    # extract_segments framing (Real_events) == slicer framing
    from gw_whisper_b200.inference import extract_segments
    assert len(extract_segments(h1)) == len(ours)


def test_slicer_whitening_needs_the_gpu_and_start_times_must_agree():
    from gw_whisper_b200.inference import ArrayFile, SegmentSlicer
    f = ArrayFile.from_segments({"H1": {"0": np.zeros(4096)}, "L1": {"0": np.zeros(4096)}}, {"0": 0.0})
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):        # white=False whitens on the GPU: no device, no silent CPU path
            SegmentSlicer(f, "0", white=False)
    with pytest.raises(AssertionError):
        g = ArrayFile.from_segments({"H1": {"0": np.zeros(4096)}}, {"0": 0.0})
        g["L1"] = ArrayFile.from_segments({"L1": {"0": np.zeros(4096)}}, {"0": 1.0})["L1"]
        SegmentSlicer(g, "0", white=True)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_evaluate_slices_generic_network_matches_reference():
    """With a plain torch module as `network`, our evaluate_slices reproduces the reference loop
    (CPU, tiny fake network) including batch boundaries and the strict > threshold."""
    from gw_whisper_b200.inference import ArrayFile, TorchSegmentSlicer, evaluate_slices
    ref = _load_reference_inference()

    class Net(torch.nn.Module):
        def forward(self, x):                    # [B, 2, 2048] -> [B, 2]
            s = torch.tanh(x[:, 0, :64].mean(1) * 4 + x[:, 1, 100:164].mean(1) * 4)
            return torch.stack([s, -s], dim=1)

    rng = np.random.default_rng(7)
    n = 2048 + 204 * 700 + 17
    f = ArrayFile.from_segments({"H1": {"5": rng.standard_normal(n).astype(np.float32)},
                                 "L1": {"5": rng.standard_normal(n).astype(np.float32)}}, {"5": 5.0})
    net = Net().eval()
    theirs_trig, theirs_vals = ref.evaluate_slices(ref.TorchSegmentSlicer(f, "5", white=True), net,
                                                   device="cpu", trigger_threshold=0.2)
    ours_trig, ours_vals = evaluate_slices(TorchSegmentSlicer(f, "5", white=True), net, device="cpu",
                                           trigger_threshold=0.2)
    assert len(theirs_vals) == len(ours_vals) == 3
    for a, b in zip(theirs_vals, ours_vals):
        assert np.allclose(a, b, atol=1e-6)
    assert len(theirs_trig) == len(ours_trig) > 0
    assert np.allclose(np.array(theirs_trig), np.array(ours_trig), atol=1e-6)
