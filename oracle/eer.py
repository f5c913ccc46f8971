"""CPU restatement of the reference's EER computation -- TEST INFRASTRUCTURE ONLY (imported by tests/ only).

Follows /root/reference/eval_metrics_DF.py:21-48 (``compute_det_curve`` :21-39, ``compute_eer`` :42-48) as called by
/root/reference/evaluate_2021_DF.py:21-39: scores are concatenated targets-first, sorted ascending with a STABLE
mergesort (ties keep targets before non-targets), false-rejection / false-acceptance rates come from cumulative label
sums, and the EER is the mean of the two rates at the index where they are closest.
Pinned: tests/test_oracle.py::test_eer_matches_reference_code runs the reference module itself where
/root/reference is mounted, and against tests/golden/eer_cases.npz everywhere.
"""
import numpy as np


def compute_det_curve(target_scores: np.ndarray, nontarget_scores: np.ndarray):
    n_t, n_n = target_scores.size, nontarget_scores.size
    scores = np.concatenate((target_scores, nontarget_scores))
    labels = np.concatenate((np.ones(n_t), np.zeros(n_n)))
    order = np.argsort(scores, kind="mergesort")                       # eval_metrics_DF.py:28
    labels = labels[order]
    tar = np.cumsum(labels)                                            # :32
    non = n_n - (np.arange(1, n_t + n_n + 1) - tar)                    # :33
    frr = np.concatenate((np.atleast_1d(0), tar / n_t))                # :35
    far = np.concatenate((np.atleast_1d(1), non / n_n))                # :36
    thr = np.concatenate((np.atleast_1d(scores[order[0]] - 0.001), scores[order]))   # :37
    return frr, far, thr


def compute_eer(target_scores: np.ndarray, nontarget_scores: np.ndarray):
    frr, far, thr = compute_det_curve(target_scores, nontarget_scores)
    i = int(np.argmin(np.abs(frr - far)))                              # :46
    return float(np.mean((frr[i], far[i]))), float(thr[i])             # :47
