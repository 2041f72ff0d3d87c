"""FAR / sensitive-distance scoring (gw_whisper_b200/evaluate.py) against the reference's own
`MLGWSC-1/evaluate.py` (imported with `h5py` stubbed) on seeded synthetic event lists, and against golden
outputs of that reference committed under tests/golden/ (made by tests/golden/make_evaluate_golden.py), which
is what runs where /root/reference is not mounted.  Integer / index outputs must be identical, float outputs
equal to 1e-12 relative (the reference's unstable argsort may order tied statistics differently inside a
cumulative sum)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
have_ref = os.path.isdir(REF)
INT_KEYS = ("found-indices", "missed-indices", "true-positive-event-indices", "false-positive-event-indices",
            "sorting-indices")


def load_reference_evaluate():
    stub = None
    if "h5py" not in sys.modules:
        stub = types.ModuleType("h5py")
        stub.File = object
        sys.modules["h5py"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_evaluate", os.path.join(REF, "MLGWSC-1/evaluate.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if stub is not None:
            sys.modules.pop("h5py", None)
    return mod


def make_case(seed, n_inj=400, n_fg=900, n_bg=700, ties=True):
    """A month-like toy: injections every ~24 s with jitter, foreground events = recovered injections with
    timing error + false alarms, background = false alarms only."""
    rng = np.random.default_rng(seed)
    tc = np.sort(np.cumsum(rng.uniform(20.0, 30.0, n_inj)) + 1.0e9)
    inj = {"tc": tc, "distance": rng.uniform(100.0, 7000.0, n_inj),
           "mass1": rng.uniform(10.0, 50.0, n_inj), "mass2": rng.uniform(7.0, 40.0, n_inj)}
    n_found = n_fg // 2
    which = rng.choice(n_inj, n_found, replace=True)            # some injections picked up twice
    t_found = tc[which] + rng.normal(0.0, 0.15, n_found)         # some outside the 0.3 s window
    t_false = rng.uniform(tc[0] - 50.0, tc[-1] + 50.0, n_fg - n_found)
    t = np.r_[t_found, t_false]
    stat = np.r_[rng.uniform(0.3, 1.0, n_found), rng.uniform(0.0, 0.8, n_fg - n_found)]
    if ties:
        stat[:40] = np.round(stat[:40], 1)                       # tied ranking statistics
    perm = rng.permutation(n_fg)
    fg = np.vstack([t[perm], stat[perm], np.full(n_fg, 0.3)])
    bg_stat = rng.uniform(0.0, 0.9, n_bg)
    if ties:
        bg_stat[:30] = np.round(bg_stat[:30], 1)
    bg = np.vstack([rng.uniform(tc[0], tc[-1], n_bg), bg_stat, np.full(n_bg, 0.3)])
    return fg, bg, inj


CASES = [(0, dict()), (1, dict(n_inj=50, n_fg=40, n_bg=10)), (2, dict(n_inj=1000, n_fg=5000, n_bg=4000)),
         (3, dict(n_inj=30, n_fg=300, n_bg=0)), (4, dict(ties=False))]


def compare(got, ref):
    assert set(got) == set(ref)
    for k in ref:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        assert a.shape == b.shape, k
        if k in INT_KEYS:
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=0.0, err_msg=k)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
@pytest.mark.parametrize("seed,kw", CASES)
@pytest.mark.parametrize("chirp", [False, True])
@pytest.mark.parametrize("dur", [None, 2.5e4])
def test_get_stats_matches_reference(seed, kw, chirp, dur):
    from gw_whisper_b200 import evaluate as E
    ref_mod = load_reference_evaluate()
    fg, bg, inj = make_case(seed, **kw)
    ref = ref_mod.get_stats(fg.copy(), bg.copy(), {k: v.copy() for k, v in inj.items()}, duration=dur, chirp_distance=chirp)
    got = E.get_stats(fg, bg, inj, duration=dur, chirp_distance=chirp)
    compare(got, ref)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_find_closest_index_matches_reference():
    from gw_whisper_b200 import evaluate as E
    ref_mod = load_reference_evaluate()
    rng = np.random.default_rng(5)
    arr = rng.uniform(0, 100, 200)
    arr[10] = arr[11]                                 # duplicates
    val = np.r_[rng.uniform(-10, 110, 500), arr[:20], (np.sort(arr)[:-1] + np.sort(arr)[1:]) / 2]   # incl. exact hits and midpoints
    assert np.array_equal(E.find_closest_index(arr, val), ref_mod.find_closest_index(arr, val.copy()))
    assert np.array_equal(E.find_closest_index(np.sort(arr), val, assume_sorted=True),
                          ref_mod.find_closest_index(np.sort(arr), val.copy(), assume_sorted=True))
    with pytest.raises(ValueError):
        E.find_closest_index(np.array([]), val)
    assert np.isclose(E.mchirp(30.0, 20.0), ref_mod.mchirp(30.0, 20.0), rtol=0, atol=0)


def test_get_stats_matches_golden():
    from gw_whisper_b200 import evaluate as E
    gold = np.load(os.path.join(HERE, "golden", "evaluate_golden.npz"))
    n = 0
    for seed, kw in CASES:
        for chirp in (False, True):
            fg, bg, inj = make_case(seed, **kw)
            got = E.get_stats(fg, bg, inj, duration=None, chirp_distance=chirp)
            ref = {k.split("|", 2)[2]: gold[k] for k in gold.files if k.startswith(f"{seed}|{int(chirp)}|")}
            compare(got, ref)
            n += 1
    assert n == 2 * len(CASES)


def test_no_recovered_injection_raises_like_the_reference():
    from gw_whisper_b200 import evaluate as E
    fg, bg, inj = make_case(7, n_inj=20, n_fg=10, n_bg=5)
    fg[0] += 1.0e6                                     # every event far from every injection
    with pytest.raises(IndexError):
        E.get_stats(fg, bg, inj)


def test_properties():
    """Scale / permutation properties that hold at any size: event order does not matter, FAR is a
    non-increasing step function of the threshold, the sensitive fraction lies in [0, 1]."""
    from gw_whisper_b200 import evaluate as E
    fg, bg, inj = make_case(11, n_inj=3000, n_fg=20000, n_bg=20000)
    a = E.get_stats(fg, bg, inj)
    perm = np.random.default_rng(0).permutation(fg.shape[1])
    b = E.get_stats(fg[:, perm], bg[:, ::-1], inj)
    for k in ("far", "sensitive-distance", "sensitive-volume", "fg-far"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12)
    assert np.all(np.diff(a["far"]) <= 0) and a["far"][-1] == 0
    assert np.all((a["sensitive-fraction"] >= 0) & (a["sensitive-fraction"] <= 1))
    assert np.all(np.diff(a["sensitive-distance"]) <= 1e-9)      # louder threshold -> smaller reach
