"""Golden outputs of the REFERENCE `MLGWSC-1/evaluate.py::get_stats` for the seeded cases of
tests/test_evaluate_host.py.  Run where /root/reference exists:  python tests/golden/make_evaluate_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_evaluate_host as T  # noqa: E402

ref = T.load_reference_evaluate()
out = {}
for seed, kw in T.CASES:
    for chirp in (False, True):
        fg, bg, inj = T.make_case(seed, **kw)
        res = ref.get_stats(fg, bg, inj, duration=None, chirp_distance=chirp)
        for k, v in res.items():
            out[f"{seed}|{int(chirp)}|{k}"] = np.asarray(v)
np.savez_compressed(os.path.join(HERE, "evaluate_golden.npz"), **out)
print("wrote", len(out), "arrays")
