#!/usr/bin/env python
"""Same command line, checks and output as the reference's ``evaluate_2021_DF.py`` (score file, key directory, phase ->
"eer: xx.xx"), computed by ``sls_b200.compute_eer`` (the device restatement of eval_metrics_DF.compute_eer; runs on the GPU
when there is one, on the CPU otherwise):

    python tools/evaluate_2021_DF.py Score_DF.txt ./keys eval

 -Score_DF.txt: "<utt> <score>" rows as written by produce_evaluation_file / tools/score_files.py
 -keys: directory holding CM/trial_metadata.txt (column 1 = utterance, 5 = bonafide | spoof, 7 = progress | eval | hidden_track)
 -phase: progress | eval | hidden_track
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def eval_to_score_file(score_file, cm_key_file, phase):
    """evaluate_2021_DF.py:21-39."""
    import torch
    import sls_b200
    with open(cm_key_file) as f:
        keys = [ln.split(" ") for ln in f.read().splitlines() if ln]
    with open(score_file) as f:
        rows = [ln.split() for ln in f.read().splitlines() if ln.strip()]
    if len(rows) != len(keys):
        print("CHECK: submission has %d of %d expected trials." % (len(rows), len(keys)))
        sys.exit(1)
    if any(len(r) > 2 for r in rows):
        print("CHECK: submission has more columns (%d) than expected (2). Check for leading/ending blank spaces." % max(len(r) for r in rows))
        sys.exit(1)
    label = {k[1]: k[5] for k in keys if k[7] == phase}            # :32 inner join on the utterance id, rows of this phase only
    kept = [(float(s), label[u]) for u, s in rows if u in label]
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    scores = torch.tensor([s for s, _ in kept], dtype=torch.float64, device=dev)
    bona = torch.tensor([lab == "bonafide" for _, lab in kept], device=dev)
    known = torch.tensor([lab in ("bonafide", "spoof") for _, lab in kept], device=dev)
    eer_cm = sls_b200.compute_eer(scores[known], bona[known])[0]
    print("eer: %.2f\n" % (100 * eer_cm))
    return eer_cm


if __name__ == "__main__":
    if len(sys.argv) != 4:
        print("CHECK: invalid input arguments. Please read the instruction below:")
        print(__doc__)
        sys.exit(1)
    submit_file, truth_dir, phase = sys.argv[1:4]
    if not os.path.isfile(submit_file):
        print("%s doesn't exist" % submit_file)
        sys.exit(1)
    if not os.path.isdir(truth_dir):
        print("%s doesn't exist" % truth_dir)
        sys.exit(1)
    if phase not in ("progress", "eval", "hidden_track"):
        print("phase must be either progress, eval, or hidden_track")
        sys.exit(1)
    eval_to_score_file(submit_file, os.path.join(truth_dir, "CM/trial_metadata.txt"), phase)
