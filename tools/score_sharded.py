#!/usr/bin/env python
"""BASELINE configs 3 / 5: data-parallel scoring of N synthetic clips, utterance-sharded over the ranks of one node
(torchrun, one process per GPU, NCCL), one all-gather of float32 scores at the end, rank 0 writes score.txt and the EER.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/score_sharded.py --utts 611829 --head sls --out /tmp/score.txt
    python tools/score_sharded.py --utts 4096 --verify        # 1 GPU; --verify re-scores a sample un-sharded and compares bits

Clips are generated on the device from a counter-based hash keyed by the utterance index, so any rank can produce any
clip and the gathered vector does not depend on the world size (asserted with --verify against a rank-0 re-score of the
head, the tail and the shard boundaries)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", dest="n", type=int, default=4096)
    ap.add_argument("--head", default="sls", choices=["sls", "sae", "window"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--out", default="")
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sls_b200

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                                   # same random-init weights on every rank
    geo = sls_b200.TrunkGeometry(layers=a.layers)
    cls = {"sls": sls_b200.ModelSLS, "sae": sls_b200.Model, "window": sls_b200.ModelWindowTopK}[a.head]
    model = cls(None, dev, cp_path=None, precision=a.precision, geometry=geo).to(dev).eval()
    lo, hi = sls_b200.shard_range(a.n, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    local_scores = sls_b200.score_synthetic_shard(model, lo, hi, batch=a.batch)
    full = sls_b200.gather_scores(local_scores, a.n, rank, world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = True
    if a.verify:
        # bit-stability across shardings: re-score boundary neighbourhoods as differently composed batches on this rank
        probes = sorted({i for r in range(world) for b in sls_b200.shard_range(a.n, r, world) for i in range(max(0, b - 3), min(a.n, b + 3))})
        again = torch.cat([sls_b200.score_synthetic_shard(model, i, i + 1, batch=1) for i in probes]) if probes else full[:0]
        ok = bool(torch.equal(again, full[torch.tensor(probes, device=dev)])) if probes else True
        flag = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    if rank == 0:
        labels = (torch.arange(a.n, device=dev) % 10) == 0             # synthetic protocol: every 10th trial is bonafide
        eer, thr = sls_b200.compute_eer(full, labels) if a.n >= 20 else (float("nan"), float("nan"))
        if a.out:
            sls_b200.write_score_file(a.out, [f"SYN_{i:07d}" for i in range(a.n)], full.cpu().tolist())
        print(json.dumps({"n": a.n, "world": world, "head": a.head, "seconds": dt, "utt_per_s": a.n / dt, "eer": eer, "threshold": thr,
                          "verify_bit_identical": ok if a.verify else None, "checksum": float(full.double().sum()),
                          "first": full[:4].cpu().tolist()}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
