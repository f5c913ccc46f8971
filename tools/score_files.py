#!/usr/bin/env python
"""`main.py --is_eval --eval` on real audio files, the B200 way: trial list -> FLAC/WAV decode pool -> PCM shard -> scores ->
score.txt (-> EER when a key file is given).  One process per GPU under torchrun; every rank decodes and scores its own
contiguous block of the trial list, one all-gather of float32 scores at the end, rank 0 writes the file.

    python tools/score_files.py --protocol ASVspoof2021.DF.cm.eval.trl.txt --audio-dir /data/ASVspoof2021_DF_eval \\
        --model-path models/best.pth --out score.txt
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/score_files.py ...

Reference behaviour mirrored: trial list = one utterance id per line (`genSpoof_list(is_eval=True)`, data_utils_SSL.py:41-46),
audio at `<audio-dir>/flac/<utt>.flac` (`Dataset_ASVspoof2021_eval.__getitem__`, :109-115), first 64 600 samples or tile-repeat
padding (`pad`, :58-65), `"{utt} {score}\\n"` rows in protocol order (`produce_evaluation_file`, main.py:158-199).
`--shard-only` stops after the decode stage (no GPU needed): the shards can be re-used by later runs with `--reuse-shards`."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_trial_list(path):
    """data_utils_SSL.py:41-46 keeps the whole stripped line as the key; protocol files with several columns carry the
    utterance id in the 2nd column (:35, :49), which `--id-column` selects."""
    with open(path) as f:
        return [ln.strip() for ln in f if ln.strip()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--protocol", required=True, help="trial list, one utterance per line")
    ap.add_argument("--id-column", type=int, default=-1, help="take this whitespace-separated column as the id (-1: the whole line)")
    ap.add_argument("--audio-dir", required=True)
    ap.add_argument("--audio-pattern", default="{dir}/flac/{utt}.flac", help="path template (data_utils_SSL.py:112)")
    ap.add_argument("--shard-dir", default="", help="where the PCM shards go (default: <out>.shards)")
    ap.add_argument("--reuse-shards", action="store_true")
    ap.add_argument("--shard-only", action="store_true")
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    ap.add_argument("--samples", type=int, default=64600)
    ap.add_argument("--head", default="sae", choices=["sls", "sae", "window"])
    ap.add_argument("--ssl-checkpoint", default=None, help="fairseq xlsr2_300m.pt (model.py:113); omit for random init")
    ap.add_argument("--model-path", default=None, help="trained detector state_dict (main.py:531-592)")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--out", default="score.txt")
    ap.add_argument("--keys", default="", help="optional trial_metadata.txt for the EER (evaluate_2021_DF.py:21-39)")
    ap.add_argument("--key-id-column", type=int, default=1)
    ap.add_argument("--key-label-column", type=int, default=5)
    a = ap.parse_args()

    import sls_b200

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    utts = read_trial_list(a.protocol)
    if a.id_column >= 0:
        utts = [u.split()[a.id_column] for u in utts]
    n = len(utts)
    lo, hi = sls_b200.shard_range(n, rank, world)
    shard_root = a.shard_dir or (a.out + ".shards")
    my_dir = os.path.join(shard_root, f"rank{rank:03d}of{world:03d}")

    t0 = time.perf_counter()
    if not (a.reuse_shards and os.path.exists(os.path.join(my_dir, "pcm.npy"))):
        paths = [a.audio_pattern.format(dir=a.audio_dir, utt=u) for u in utts[lo:hi]]
        missing = [p for p in paths if not os.path.exists(p)]
        if missing:
            sys.exit(f"rank {rank}: {len(missing)} audio files missing, first: {missing[0]}")
        sls_b200.audio_files_to_shard(my_dir, utts[lo:hi], paths, workers=a.workers, max_samples=a.samples)
    shard = sls_b200.PcmShard(my_dir)
    if shard.utt_ids != utts[lo:hi]:
        sys.exit(f"rank {rank}: shard {my_dir} does not match trials {lo}..{hi} of the protocol (stale --reuse-shards?)")
    t_decode = time.perf_counter() - t0
    if a.shard_only:
        print(json.dumps({"rank": rank, "world": world, "trials": hi - lo, "decode_seconds": t_decode,
                          "clips_per_s": (hi - lo) / max(t_decode, 1e-9), "pcm_samples": int(shard.pcm.shape[0]), "shard": my_dir}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    geo = sls_b200.TrunkGeometry(layers=a.layers)
    cls = {"sls": sls_b200.ModelSLS, "sae": sls_b200.Model, "window": sls_b200.ModelWindowTopK}[a.head]
    model = cls(None, dev, cp_path=a.ssl_checkpoint, precision=a.precision, geometry=geo)
    if a.model_path:
        sls_b200.load_model_checkpoint(model, a.model_path)
    model = model.to(dev).eval()
    t1 = time.perf_counter()
    local_scores = sls_b200.score_pcm_shard(model, shard, batch=a.batch, samples=a.samples).to(dev)
    full = sls_b200.gather_scores(local_scores, n, rank, world)
    torch.cuda.synchronize()
    t_score = time.perf_counter() - t1
    if rank == 0:
        sls_b200.write_score_file(a.out, utts, full.cpu().tolist())
        rec = {"trials": n, "world": world, "head": a.head, "decode_seconds": t_decode, "score_seconds": t_score,
               "utt_per_s": n / max(t_score, 1e-9), "out": a.out, "checksum": float(full.double().sum())}
        if a.keys:
            lab = {}
            with open(a.keys) as f:                      # evaluate_2021_DF.py:21-39: column 1 = utterance, column 5 = bonafide / spoof
                for ln in f:
                    c = ln.split()
                    if len(c) > max(a.key_id_column, a.key_label_column):
                        lab[c[a.key_id_column]] = c[a.key_label_column]
            known = [i for i, u in enumerate(utts) if lab.get(u) in ("bonafide", "spoof")]
            if known:
                idx = torch.tensor(known, device=dev)
                is_bona = torch.tensor([lab[utts[i]] == "bonafide" for i in known], device=dev)
                eer, thr = sls_b200.compute_eer(full[idx], is_bona)
                rec.update({"eer": eer, "threshold": thr, "scored_with_keys": len(known)})
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
