"""False-alarm rate and sensitive distance of a trigger list (the scoring step after the search).

Drop-in for the numerical core of the reference's `MLGWSC-1/evaluate.py` (`find_closest_index` :62-97,
`mchirp` :100-101, `get_stats` :104-278): same signatures, same dictionary keys, same results, written
from scratch as array operations (the reference walks the injections in a Python loop).  This is SURVEY.md
section 8(f) row 3: with it the month-long time-sharded search scores itself from the gathered
`(time, stat, var)` triggers.  It is host code by nature -- O(events log events) on a few 10^5 events --
and stays numpy; the HDF5 reading / writing of the reference's `main` is not reproduced (h5py is not in
this image), callers pass arrays.

Reference behaviours kept on purpose (the parity tests pin them):
  * `find_closest_index` returns indices into the SORTED copy of `array`; `get_stats` then uses them on
    the caller's `tc` as given (evaluate.py:160-163), so `tc` is expected in ascending order, as
    `generate_data.py` writes it;
  * ties between the two neighbours go to the right one (strict `<` at :92-93);
  * an event list without a single recovered injection makes the reference fail with an IndexError at
    :236 (`found_injections[1]` of an empty array); the same exception type is raised here.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

__all__ = ["find_closest_index", "mchirp", "get_stats"]


def find_closest_index(array, value, assume_sorted: bool = False):
    """Index (into the ascending-sorted `array`) of the element closest to each `value`
    (evaluate.py:62-97)."""
    array = np.asarray(array)
    if array.size == 0:
        raise ValueError("Cannot find closest index for empty input array.")
    srt = array if assume_sorted else np.sort(array)
    value = np.asarray(value)
    right = np.searchsorted(srt, value, side="right")
    left = np.maximum(right - 1, 0)
    right_c = np.minimum(right, srt.size - 1)
    take_left = (right == srt.size) | (np.fabs(srt[left] - value) < np.fabs(srt[right_c] - value))
    return np.where(take_left, right - 1, right)


def mchirp(mass1, mass2):
    """Chirp mass (evaluate.py:100-101)."""
    return (mass1 * mass2) ** (3.0 / 5.0) / (mass1 + mass2) ** (1.0 / 5.0)


def _descending_rank_rate(stats: np.ndarray, duration: float) -> np.ndarray:
    """FAR attached to each sorted noise statistic: number of louder noise events per unit time."""
    n = stats.size
    return (n - np.arange(n) - 1) / duration


def get_stats(fgevents, bgevents, injparams: Dict[str, np.ndarray], duration: Optional[float] = None,
              chirp_distance: bool = False) -> Dict[str, np.ndarray]:
    """`fgevents` / `bgevents`: arrays [3, n] of (time, ranking statistic, matching window);
    `injparams`: at least `tc` and `distance` (plus `mass1`, `mass2` for `chirp_distance`).
    Returns the reference's dictionary (evaluate.py:104-278)."""
    fgevents = np.asarray(fgevents)
    bgevents = np.asarray(bgevents)
    injtimes = np.asarray(injparams["tc"])
    dist = np.asarray(injparams["distance"])
    massc = mchirp(np.asarray(injparams["mass1"]), np.asarray(injparams["mass2"])) if chirp_distance else None
    if duration is None:
        duration = injtimes.max() - injtimes.min()

    order = fgevents[0].argsort()
    fg = fgevents[:, order]
    nearest = find_closest_index(injtimes, fg[0])
    diff = np.abs(injtimes[nearest] - fg[0])
    is_tp = diff <= fg[2]
    tp_idx = np.flatnonzero(is_tp)
    fp_idx = np.flatnonzero(~is_tp)

    ret: Dict[str, np.ndarray] = {
        "fg-events": fg,
        "found-indices": np.arange(injtimes.size)[nearest],
        "true-positive-event-indices": tp_idx,
        "false-positive-event-indices": fp_idx,
        "sorting-indices": order,
        "true-positive-diffs": diff[tp_idx],
        "false-positive-diffs": diff[fp_idx],
        "true-positives": fg[:, tp_idx],
        "false-positives": fg[:, fp_idx],
    }
    ret["missed-indices"] = np.setdiff1d(np.arange(injtimes.size), ret["found-indices"])

    ret["fg-far"] = _descending_rank_rate(np.sort(fg[1, fp_idx]), duration)
    noise_stats = np.sort(bgevents[1])
    ret["far"] = _descending_rank_rate(noise_stats, duration)

    # loudest true positive of every injection that has one, injections in ascending index order
    inj_of_tp = nearest[tp_idx]
    stat_of_tp = fg[1, tp_idx]
    if inj_of_tp.size == 0:
        raise IndexError("no injection was recovered: the sensitive volume is undefined "
                         "(the reference fails at evaluate.py:236 on the empty found-injection table)")
    by_inj = np.lexsort((stat_of_tp, inj_of_tp))            # primary key injection, then statistic
    inj_sorted = inj_of_tp[by_inj]
    last_of_group = np.flatnonzero(np.r_[inj_sorted[1:] != inj_sorted[:-1], True])
    found_inj = inj_sorted[last_of_group]
    found_stat = stat_of_tp[by_inj][last_of_group]           # the maximum of each group

    by_stat = np.argsort(found_stat, kind="stable")
    found_inj = found_inj[by_stat]
    found_stat = found_stat[by_stat]

    max_distance = dist.max()
    vtot = (4.0 / 3.0) * np.pi * max_distance ** 3.0
    ninj = dist.size
    above = np.searchsorted(found_stat, noise_stats, side="right")   # found injections at or below each threshold
    nfound = found_stat.size - above
    if chirp_distance:
        mchirp_max = massc.max()
        prefactor = vtot / (mchirp_max ** (5.0 / 2.0) * massc.size)
        loudest_first = massc[found_inj.astype(int)][::-1]
        tail = np.r_[np.cumsum(loudest_first ** (5.0 / 2.0))[::-1], 0.0]
        tail_sq = np.r_[np.cumsum(loudest_first ** 5)[::-1], 0.0]
        mc_sum = tail[above]
        ninj = np.sum((mchirp_max / massc) ** (5.0 / 2.0))
        sample_variance = tail_sq[above] / ninj - (mc_sum / ninj) ** 2
    else:
        prefactor = vtot / ninj
        mc_sum = nfound
        sample_variance = nfound / ninj - (nfound / ninj) ** 2
    vol = prefactor * mc_sum
    ret["sensitive-volume"] = vol
    ret["sensitive-distance"] = (3 * vol / (4 * np.pi)) ** (1.0 / 3.0)
    ret["sensitive-volume-error"] = prefactor * (ninj * sample_variance) ** 0.5
    ret["sensitive-fraction"] = nfound / ninj
    return ret
