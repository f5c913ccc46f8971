#!/usr/bin/env python
"""Replay of the reference's evaluation branch on this package (``python main.py --is_eval ...``, /root/reference/main.py).

The reference's ``main.py`` itself runs unchanged on top of the shim modules of INTEGRATION.md (``tests/test_host.py::
test_reference_main_py_runs_on_the_shim`` executes it verbatim where ``/root/reference`` is mounted); this file is the same
sequence of observable steps for boxes that do not have the reference tree (the GPU box), written from its behaviour:

    main.py:403-455   the command line (same flag names / defaults; training-only flags are accepted and ignored)
    main.py:493-517   ModelWindowTopK / ModelTopK constructed with exactly these keyword arguments
    main.py:518-521   nn.DataParallel(model).to(device), parameter count printed
    main.py:531-592   checkpoint: locate the state_dict, fix the ``module.`` prefix, strict load with non-strict fallback
    main.py:630-653   genSpoof_list -> Dataset_*_eval -> delete a stale score file -> produce_evaluation_file (batch 20)

    python tools/main_eval.py --is_eval --track DF --cp_path xlsr2_300m.pt --model_path best.pth \
        --database_path /data/ASVspoof2021_DF_eval --protocols_path /data/DF.cm.eval.trl.txt --eval_output scores/scores_DF.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(argv=None):
    ap = argparse.ArgumentParser(description="ASVspoof2021 SSL+SAE scoring (evaluation branch of main.py)")
    ap.add_argument("--database_path", type=str, default="/root/autodl-tmp/CLAD/Datasets/LA/")
    ap.add_argument("--protocols_path", type=str, default="/root/autodl-tmp/CLAD/Datasets/LA/")
    ap.add_argument("--track", type=str, default="DF", choices=["LA", "In-the-Wild", "DF"])
    ap.add_argument("--batch_size", type=int, default=14)
    ap.add_argument("--cp_path", type=str, default="/root/autodl-tmp/SLSforASVspoof-2021-DF/xlsr2_300m.pt")
    ap.add_argument("--sae_weight", type=float, default=0.1)
    ap.add_argument("--sae_dict_size", type=int, default=4096)
    ap.add_argument("--sae_k", type=int, default=128)
    ap.add_argument("--use_window_topk", action="store_true", default=False)
    ap.add_argument("--sae_window_size", type=int, default=8)
    ap.add_argument("--use_sparse_features", action="store_true", default=True)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--comment", type=str, default=None)
    ap.add_argument("--quick_test", action="store_true", default=False)
    ap.add_argument("--model_path", type=str, default=None)
    ap.add_argument("--is_eval", action="store_true", default=False)
    ap.add_argument("--eval_output", type=str, default=None)
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"], help="this package only: bf16 tensor-core path or fp32 verification path")
    # training-only flags of main.py: accepted so that an existing command line keeps working, never read
    for name, typ in (("num_epochs", int), ("lr", float), ("weight_decay", float), ("algo", int), ("nBands", int), ("minF", int), ("maxF", int),
                      ("minBW", int), ("maxBW", int), ("minCoeff", int), ("maxCoeff", int), ("minG", int), ("maxG", int),
                      ("minBiasLinNonLin", int), ("maxBiasLinNonLin", int), ("N_f", int), ("P", int), ("g_sd", int), ("SNRmin", int), ("SNRmax", int)):
        ap.add_argument("--" + name, type=typ, default=None)
    ap.add_argument("--resume", action="store_true", default=False)
    ap.add_argument("--fresh_start", action="store_true", default=False)
    return ap.parse_args(argv)


def main(argv=None) -> int:
    args = parse(argv)
    import torch
    from torch import nn
    import sls_b200

    torch.manual_seed(args.seed)                                     # core_scripts/startup_config.py:54-57
    device = "cuda" if torch.cuda.is_available() else "cpu"
    print(f"Device: {device}")
    if not args.is_eval:
        print("Error: this replay covers the evaluation branch only (--is_eval); training is out of scope (DESIGN.md section 8)")
        return 2
    if args.use_window_topk:
        print(f"Using Window-based TopK (window_size={args.sae_window_size})")
        model = sls_b200.ModelWindowTopK(args=args, device=device, cp_path=args.cp_path, use_sae=True, use_sparse_features=args.use_sparse_features,
                                         sae_dict_size=args.sae_dict_size, sae_k=args.sae_k, sae_window_size=args.sae_window_size,
                                         sae_weight=args.sae_weight, precision=args.precision)
    else:
        print("Using Per-Timestep TopK")
        model = sls_b200.Model(args=args, device=device, cp_path=args.cp_path, use_sae=True, use_sparse_features=args.use_sparse_features,
                               sae_dict_size=args.sae_dict_size, sae_k=args.sae_k, sae_weight=args.sae_weight, precision=args.precision)
    model = nn.DataParallel(model).to(device)                        # main.py:518 (one visible GPU per process: see INTEGRATION.md)
    print(f"Total parameters: {sum(p.numel() for p in model.parameters()):,}")

    if not args.model_path:
        print("Error: --model_path is required for evaluation mode")
        return 1
    print(f"Loading: {args.model_path}")
    sls_b200.load_model_checkpoint(model, args.model_path)           # main.py:531-592 (state_dict lookup, prefix fix, strict -> non-strict)
    print("Loaded weights (fresh optimizer/epoch)")

    file_eval = sls_b200.genSpoof_list(dir_meta=args.protocols_path, is_train=False, is_eval=True)
    if args.track == "In-the-Wild":
        eval_set = sls_b200.Dataset_in_the_wild_eval(list_IDs=file_eval, base_dir=args.database_path)
    else:
        eval_set = sls_b200.Dataset_ASVspoof2021_eval(list_IDs=file_eval, base_dir=args.database_path)
    out = args.eval_output or os.path.join("scores", f"scores_{args.track}.txt")
    if os.path.dirname(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out):
        os.remove(out)                                               # main.py:646-647: never append to a stale file
    sls_b200.produce_evaluation_file(eval_set, model, device, out, quick_test=args.quick_test)
    print(f"Scores saved to: {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
