"""Scoring loop, score.txt writer and utterance sharding (the callers either side of ``Model.forward``).

Mirrors ``/root/reference/main.py:158-199`` (``produce_evaluation_file``), ``:630-653`` (eval branch) and
``/root/reference/data_utils_SSL.py:58-65`` (``pad``).  Multi-GPU: one process per GPU, each rank scores a
contiguous block of the protocol list, one all-gather of float32 scores at the end, rank 0 writes the file in
protocol order.  No data-path collective exists or is needed.
"""
from __future__ import annotations

import math
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------
# data-format helpers
# ---------------------------------------------------------------------------------------------

def pad_clip(x: np.ndarray, max_len: int = 64600) -> np.ndarray:
    """data_utils_SSL.py:58-65: truncate the head, or tile-repeat then cut (no normalisation)."""
    n = x.shape[0]
    if n >= max_len:
        return x[:max_len]
    reps = int(max_len / n) + 1
    return np.tile(x, reps)[:max_len]


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def synth_clip_host(utt_index: int, samples: int = 64600) -> np.ndarray:
    """Host twin of ``slsb_synth_clips`` (csrc/frontend.cu): counter-based, keyed by utterance index, bit-exact."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(0x5EED0000 + utt_index) * np.uint64(0xD6E8FEB86659FD93) & _M64)
        h = _splitmix64(np.arange(samples, dtype=np.uint64) ^ key)
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48)))
    f = s.astype(np.float32) - np.float32(131070.0)
    return f * np.float32(1.0 / math.sqrt(4 * (65536.0 ** 2 - 1) / 12.0))


class SyntheticEvalSet(torch.utils.data.Dataset):
    """Stands in for ``Dataset_ASVspoof2021_eval`` (data_utils_SSL.py:96-115): ``(x_inp[64600], utt_id)``."""

    def __init__(self, n: int, samples: int = 64600, first: int = 0):
        self.n, self.samples, self.first = n, samples, first

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        u = self.first + i
        return torch.from_numpy(pad_clip(synth_clip_host(u, self.samples), self.samples)), f"SYN_{u:07d}"


def write_score_file(path: str, utt_ids: Sequence[str], scores: Iterable[float], fmt: str = "repr", append: bool = False) -> None:
    """``"{utt} {score}\\n"``, single space, no header (main.py:190-192; ``fmt='6f'`` = the stand-alone eval scripts)."""
    with open(path, "a+" if append else "w") as fh:
        for u, s in zip(utt_ids, scores):
            fh.write("{} {}\n".format(u, s) if fmt == "repr" else "{} {:.6f}\n".format(u, s))


# ---------------------------------------------------------------------------------------------
# the reference's eval loop
# ---------------------------------------------------------------------------------------------

def produce_evaluation_file(dataset, model, device, save_path, quick_test: bool = False, batch_size: int = 20,
                            num_workers: int = 0) -> None:
    """main.py:158-199 with the same observable behaviour: appends ``utt score`` lines batch by batch, score =
    ``exp(log_softmax)[:, 1]`` printed with Python float repr.  Host batches go through ``slsb_score_host``
    (pinned H2D, forward, D2H) so the loop has one synchronisation per batch, like the reference's ``.cpu()``."""
    loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, drop_last=False, pin_memory=True,
                                         num_workers=num_workers)
    model.eval()
    m = model.module if hasattr(model, "module") else model      # nn.DataParallel wrapper (main.py:518)
    eng = m.engine()
    head, prec = m._head() if hasattr(m, "_head") else 3, m._prec()
    for i, (batch_x, utt_id) in enumerate(loader):
        if quick_test and i >= 5:
            print("Quick test: breaking evaluation loop after 5 batches")
            break
        if batch_x.ndim == 3:
            batch_x = batch_x[:, :, 0]
        batch_x = batch_x.to(torch.float32).contiguous()
        scores = eng.score_host(batch_x, head, prec).tolist()
        write_score_file(save_path, list(utt_id), scores, append=True)
    print("Scores saved to {}".format(save_path))


# ---------------------------------------------------------------------------------------------
# utterance sharding (BASELINE config 3)
# ---------------------------------------------------------------------------------------------

def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; blocks differ by at most one utterance."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def score_synthetic_shard(model, lo: int, hi: int, batch: int = 64, samples: int = 64600) -> torch.Tensor:
    """Scores utterances [lo, hi) of the synthetic corpus, clips generated on the device (device-resident input)."""
    m = model.module if hasattr(model, "module") else model
    eng = m.engine()
    head, prec = (m._head() if hasattr(m, "_head") else 3), m._prec()
    out = torch.empty(hi - lo, device=eng.device, dtype=torch.float32)
    for s in range(lo, hi, batch):
        n = min(batch, hi - s)
        wav = eng.synth_clips(s, n, samples)
        logp = eng.forward(wav, head, prec)
        out[s - lo:s - lo + n] = torch.exp(logp[:, 1])
    return out


# ---------------------------------------------------------------------------------------------
# variable-length clips (BASELINE config 4): bucket by frame count, right-pad inside a bucket
# ---------------------------------------------------------------------------------------------

def bucket_by_frames(sample_lengths: Sequence[int], frames_of, bucket_frames: int = 64, max_batch: int = 64,
                     min_rows: Optional[int] = None) -> List[List[int]]:
    """Groups utterance indices so that every batch holds clips whose conv-stack frame counts fall into the same
    ``bucket_frames``-wide bucket (``frames_of(samples) -> frames``, e.g. ``Engine.frames``).  Inside a batch the
    indices are sorted by length (longest first: it defines the padded width); batches come out in bucket order and
    every index appears exactly once.  ``min_rows``: buckets of short clips take larger batches (up to 4 x ``max_batch``)
    so that a forward still has about ``min_rows`` frame rows for the GEMMs (a batch of 64 one-second clips is only
    3 200 rows: a quarter of the tiles 148 SMs need); scores do not depend on the batch a clip is scored in."""
    buckets = {}
    for i, n in enumerate(sample_lengths):
        buckets.setdefault((frames_of(int(n)) - 1) // bucket_frames, []).append(i)
    out: List[List[int]] = []
    for key in sorted(buckets):
        idx = sorted(buckets[key], key=lambda i: (-int(sample_lengths[i]), i))
        bs = max_batch
        if min_rows:
            bs = max(max_batch, min(4 * max_batch, int(min_rows) // ((key + 1) * bucket_frames)))
        out.extend(idx[j:j + bs] for j in range(0, len(idx), bs))
    return out


def batch_by_length(sample_lengths: Sequence[int], frames_of, max_pad_frames: int = 64, max_batch: int = 64,
                    min_rows: Optional[int] = None) -> List[List[int]]:
    """Length-sorted batching for clips of arbitrary length (BASELINE config 4): the indices are sorted longest first and cut into
    consecutive batches; a batch is closed when it holds ``max_batch`` clips (short clips: up to 4 x ``max_batch`` while the
    forward has fewer than ``min_rows`` frame rows) or when the next clip would be padded by ``max_pad_frames`` frames or more.
    Against fixed buckets (``bucket_by_frames``) there is one ragged batch per job instead of one per bucket and the padding of a
    batch is bounded by the length spread of ITS clips.  Every index appears exactly once; scores do not depend on the batching."""
    order = sorted(range(len(sample_lengths)), key=lambda i: (-int(sample_lengths[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    top = cap = 0
    for i in order:
        f = frames_of(int(sample_lengths[i]))
        if cur and (len(cur) >= cap or top - f >= max_pad_frames):
            out.append(cur)
            cur = []
        if not cur:
            top = f
            cap = max_batch
            if min_rows:
                cap = max(max_batch, min(4 * max_batch, int(min_rows) // max(top, 1)))
        cur.append(i)
    if cur:
        out.append(cur)
    return out


@torch.no_grad()
def score_variable_length(model, clips: Sequence[torch.Tensor], bucket_frames: int = 64, max_batch: int = 64,
                          min_rows: Optional[int] = None) -> torch.Tensor:
    """Scores 1-D float32 clips of different lengths: length-sorted batches (``batch_by_length``, at most ``bucket_frames`` frames of padding), zero right-padding to the longest clip
    of the batch, sample lengths passed down so that padded frames are masked (wav2vec2.py:567-586).  Returns the
    scores in the order of ``clips`` (CPU float32)."""
    m = model.module if hasattr(model, "module") else model
    eng = m.engine()
    head, prec = (m._head() if hasattr(m, "_head") else 3), m._prec()
    lens = [int(c.numel()) for c in clips]
    scores = torch.empty(len(clips), dtype=torch.float32)
    for batch in batch_by_length(lens, eng.frames, bucket_frames, max_batch, min_rows):
        S = lens[batch[0]]
        wav = torch.zeros(len(batch), S, dtype=torch.float32, pin_memory=True)
        for j, i in enumerate(batch):
            wav[j, :lens[i]] = clips[i]
        sl = torch.tensor([lens[i] for i in batch], dtype=torch.int32).pin_memory()
        scores[torch.tensor(batch)] = eng.score_host(wav, head, prec, lens_host=sl)
    return scores


# ---------------------------------------------------------------------------------------------
# EER on the device (closes main.py --is_eval -> score.txt -> evaluate_2021_DF.py)
# ---------------------------------------------------------------------------------------------

def compute_eer(scores: torch.Tensor, is_bonafide: torch.Tensor) -> Tuple[float, float]:
    """eval_metrics_DF.py:21-48 on whatever device ``scores`` lives on (600 k DF-eval scores never leave the GPU):
    targets-first concatenation, STABLE ascending sort, cumulative label sums in int64, rates in float64, EER = mean of
    FRR and FAR where they are closest.  Returns ``(eer, threshold)`` as Python floats."""
    scores = scores.reshape(-1)
    lab = is_bonafide.reshape(-1).to(device=scores.device, dtype=torch.bool)
    s = torch.cat((scores[lab], scores[~lab])).to(torch.float64)                   # :24-25 targets first
    n_t, n_n = int(lab.sum()), int((~lab).sum())
    if n_t == 0 or n_n == 0:
        raise ValueError("compute_eer needs at least one bonafide and one spoof trial")
    labels = torch.cat((torch.ones(n_t, dtype=torch.int64, device=s.device), torch.zeros(n_n, dtype=torch.int64, device=s.device)))
    srt, order = torch.sort(s, stable=True)                                        # :28 mergesort
    tar = torch.cumsum(labels[order], 0)                                           # :32
    non = n_n - (torch.arange(1, n_t + n_n + 1, device=s.device, dtype=torch.int64) - tar)   # :33
    zero, one = torch.zeros(1, dtype=torch.float64, device=s.device), torch.ones(1, dtype=torch.float64, device=s.device)
    frr = torch.cat((zero, tar.to(torch.float64) / n_t))                           # :35
    far = torch.cat((one, non.to(torch.float64) / n_n))                            # :36
    thr = torch.cat((srt[:1] - 0.001, srt))                                        # :37
    i = int(torch.argmin((frr - far).abs()))                                       # :46 (first minimum, like np.argmin)
    return float((frr[i] + far[i]) / 2), float(thr[i])


def read_score_file(path: str) -> Tuple[List[str], np.ndarray]:
    """Parses ``"{utt} {score}"`` lines (what evaluate_2021_DF.py:24 reads with pandas)."""
    utts, vals = [], []
    with open(path) as fh:
        for line in fh:
            u, v = line.split()
            utts.append(u)
            vals.append(float(v))
    return utts, np.asarray(vals, dtype=np.float64)


def gather_scores(local: torch.Tensor, n_total: int, rank: int, world: int) -> Optional[torch.Tensor]:
    """All-gather of per-rank score blocks (the only collective of the path).  Returns the [n_total] vector in
    protocol order on every rank.  Works with NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU tests)."""
    if world == 1:
        return local
    import torch.distributed as dist
    width = (n_total + world - 1) // world
    buf = torch.zeros(width, device=local.device, dtype=torch.float32)
    buf[:local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.empty(n_total, device=local.device, dtype=torch.float32)
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi] = parts[r][:hi - lo]
    return out
