"""Host side of the audio ingest: what replaces ``Dataset_*_eval.__getitem__`` + ``pad``
(data_utils_SSL.py:96-135, :58-65) and the 6-worker DataLoader (main.py:161-165) in front of a scorer that consumes
thousands of clips per second.

The reference decodes one FLAC per ``__getitem__`` with librosa, converts to float32 and tile-pads on the host.  Here the host
only ever handles **16-bit PCM**: clips are decoded once by a thread pool - FLAC with the library's own decoder
(csrc/flac_decode.cpp, ``slsb_flac_decode_mono16``: CRC-8 / CRC-16 / MD5 checked, stops after the 64 600 samples ``pad`` keeps),
RIFF/WAVE with the standard library - into a *PCM shard* (all clips back to back as int16 + an offset table), shards are
memory-mapped, and every batch travels to the device as 2 bytes per sample of the UN-padded clips; float conversion, truncation and tile-repeat padding happen in ``ingest_pcm16_kernel``
(csrc/frontend.cu) through ``slsb_score_pcm16_host`` (include/slsb200.h).

Pure host code: numpy + the standard library; the only GPU entry is ``score_pcm_shard``.
"""
from __future__ import annotations

import concurrent.futures as cf
import json
import os
import wave
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000          # ASVspoof audio; librosa.load(..., sr=16000) in the reference (data_utils_SSL.py:111)


class AudioFormatError(ValueError):
    pass


def read_wav_pcm16(path: str, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """RIFF/WAVE, 16-bit PCM -> int16 [n].  Multi-channel files are averaged to mono (what ``librosa.load(mono=True)`` does,
    here in integer arithmetic rounded half away from zero); other sample widths or rates are rejected - resampling is an
    off-line step, the scorer never guesses."""
    pcm, ch = _read_wav_frames(path, sample_rate)
    if ch > 1:
        s = pcm.reshape(-1, ch).astype(np.int32).sum(axis=1)
        pcm = (np.sign(s) * ((np.abs(s) * 2 + ch) // (2 * ch))).astype(np.int16)
    return np.ascontiguousarray(pcm, dtype=np.int16)


def _read_wav_frames(path: str, sample_rate: int = SAMPLE_RATE, max_frames: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """Interleaved int16 samples + channel count of a 16-bit PCM RIFF/WAVE file; every other flavour (24-bit, float,
    WAVE_FORMAT_EXTENSIBLE, other rates) raises ``AudioFormatError`` - including the ones the ``wave`` module itself refuses."""
    try:
        with wave.open(path, "rb") as w:
            if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
                raise AudioFormatError(f"{path}: need 16-bit PCM, got sample width {w.getsampwidth()} / {w.getcomptype()}")
            if w.getframerate() != sample_rate:
                raise AudioFormatError(f"{path}: need {sample_rate} Hz, got {w.getframerate()} Hz")
            ch = w.getnchannels()
            n = w.getnframes() if max_frames is None else min(w.getnframes(), int(max_frames))
            return np.frombuffer(w.readframes(n), dtype="<i2"), ch
    except (wave.Error, EOFError) as e:
        raise AudioFormatError(f"{path}: not a 16-bit PCM RIFF/WAVE file ({e})") from None


def write_wav_pcm16(path: str, pcm: np.ndarray, sample_rate: int = SAMPLE_RATE) -> None:
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


FLAC_ERRORS = {-2: "not a FLAC stream", -3: "truncated stream", -4: "bad frame header", -5: "frame header CRC-8 mismatch",
               -6: "frame CRC-16 mismatch", -7: "reserved bit pattern", -8: "MD5 of the decoded audio differs from STREAMINFO",
               -9: "unsupported stream (bit depth / mid-stream format change)", -10: "bad STREAMINFO", -11: "bad argument"}


def decode_flac_bytes(data: bytes, max_samples: Optional[int] = None, verify_md5: bool = True, sample_rate: Optional[int] = SAMPLE_RATE) -> np.ndarray:
    """A whole .flac file in memory -> mono int16 [n] (channels averaged like ``read_wav_pcm16``).  ``max_samples`` stops the
    decode early (the MD5 is then not checked unless the clip ends before - it covers the whole stream; frame CRCs always are)."""
    import ctypes as C
    from ._lib import load
    lib = load()
    buf = np.frombuffer(data, dtype=np.uint8)
    if max_samples:
        cap = int(max_samples)                                      # the common path (pad() keeps the head): one call, no probe
    else:
        probe = (C.c_int32 * 6)()
        rc = lib.slsb_flac_decode(buf.ctypes.data, buf.size, 0, 0, None, 0, probe)      # STREAMINFO only
        if rc < 0:
            raise AudioFormatError(f"FLAC: {FLAC_ERRORS.get(int(rc), rc)}")
        total = int(probe[3]) | (int(probe[5]) << 31)
        cap = total if total > 0 else 1 << 26
    out = np.empty(cap, dtype=np.int16)
    rate = C.c_int32(0)
    n = lib.slsb_flac_decode_mono16(buf.ctypes.data, buf.size, int(max_samples or 0), 1 if verify_md5 else 0, out.ctypes.data, cap, C.byref(rate))
    if sample_rate is not None and rate.value not in (0, sample_rate):
        raise AudioFormatError(f"FLAC: need {sample_rate} Hz, got {rate.value} Hz")
    if n == -9:
        raise AudioFormatError("FLAC: need 16-bit samples at a constant format (unsupported stream)")
    if n < 0:
        raise AudioFormatError(f"FLAC: {FLAC_ERRORS.get(int(n), n)}")
    return out[:n]


def read_flac_pcm16(path: str, max_samples: Optional[int] = None, verify_md5: bool = True, sample_rate: Optional[int] = SAMPLE_RATE) -> np.ndarray:
    with open(path, "rb") as f:
        return decode_flac_bytes(f.read(), max_samples, verify_md5, sample_rate)


def read_audio_pcm16(path: str, max_samples: Optional[int] = None) -> np.ndarray:
    """.flac (native decoder) or .wav by extension; ``max_samples`` keeps the head of the clip, which is all pad() uses."""
    if path.lower().endswith(".flac"):
        return read_flac_pcm16(path, max_samples)
    pcm = read_wav_pcm16(path)
    return pcm[:max_samples] if max_samples else pcm


def read_audio_float32(path: str, max_samples: Optional[int] = None) -> np.ndarray:
    """What ``librosa.load(path, sr=16000)`` hands the reference's ``__getitem__`` (data_utils_SSL.py:111) for 16-bit audio at
    16 kHz: float32 mono in [-1, 1).  Mono files: ``int16 / 32768`` (exact).  Multi-channel files: every channel converted to
    float32 first, THEN averaged in float32 - librosa's ``to_mono`` - so an odd channel sum keeps its half LSB (the int16 scorer
    path ``read_audio_pcm16`` rounds it; ASVspoof 2021 and In-the-Wild are mono, where the two agree bit for bit)."""
    if path.lower().endswith(".flac"):
        import ctypes as C
        from ._lib import load
        lib = load()
        with open(path, "rb") as f:
            buf = np.frombuffer(f.read(), dtype=np.uint8)
        info = (C.c_int32 * 6)()
        rc = lib.slsb_flac_decode(buf.ctypes.data, buf.size, 0, 0, None, 0, info)          # STREAMINFO
        if rc < 0:
            raise AudioFormatError(f"{path}: FLAC: {FLAC_ERRORS.get(int(rc), rc)}")
        rate, ch, bps = int(info[0]), int(info[1]), int(info[2])
        if rate != SAMPLE_RATE or bps != 16:
            raise AudioFormatError(f"{path}: need 16-bit samples at {SAMPLE_RATE} Hz, got {bps}-bit at {rate} Hz")
        if ch == 1:
            return read_flac_pcm16(path, max_samples).astype(np.float32) / np.float32(32768.0)
        total = int(info[3]) | (int(info[5]) << 31)
        cap = int(max_samples) if (max_samples and (total == 0 or max_samples < total)) else (total if total > 0 else 1 << 24)
        out = np.empty(cap * ch, dtype=np.int32)
        n = lib.slsb_flac_decode(buf.ctypes.data, buf.size, int(max_samples or 0), 1, out.ctypes.data, out.size, info)
        if n < 0:
            raise AudioFormatError(f"{path}: FLAC: {FLAC_ERRORS.get(int(n), n)}")
        frames = out[:n * ch].reshape(n, ch)
    else:
        pcm, ch = _read_wav_frames(path, SAMPLE_RATE, max_samples)
        if ch == 1:
            return pcm.astype(np.float32) / np.float32(32768.0)
        frames = pcm.reshape(-1, ch)
    return np.mean(frames.astype(np.float32) / np.float32(32768.0), axis=1, dtype=np.float32)


def decode_audio_files(paths: Sequence[str], workers: int = 6, max_samples: Optional[int] = None) -> List[np.ndarray]:
    """Thread pool over ``read_audio_pcm16`` (the ctypes call into the FLAC decoder releases the GIL); order preserved."""
    if workers <= 1 or len(paths) < 2:
        return [read_audio_pcm16(p, max_samples) for p in paths]
    with cf.ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(lambda p: read_audio_pcm16(p, max_samples), paths))


def decode_wav_files(paths: Sequence[str], workers: int = 6) -> List[np.ndarray]:
    """Thread pool over ``read_wav_pcm16`` (file reads and numpy release the GIL); order preserved."""
    if workers <= 1 or len(paths) < 2:
        return [read_wav_pcm16(p) for p in paths]
    with cf.ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(read_wav_pcm16, paths))


# ------------------------------------------------------------------------------------------------------------------
# PCM shard: <dir>/pcm.npy (int16, all clips back to back), <dir>/offsets.npy (int64 [n + 1]), <dir>/utts.json (ids)
# ------------------------------------------------------------------------------------------------------------------
def write_pcm_shard(directory: str, utt_ids: Sequence[str], clips: Iterable[np.ndarray]) -> None:
    clips = [np.ascontiguousarray(c, dtype=np.int16).reshape(-1) for c in clips]
    if len(clips) != len(utt_ids):
        raise ValueError(f"{len(utt_ids)} ids for {len(clips)} clips")
    if any(c.size == 0 for c in clips):
        raise AudioFormatError("empty clip (the reference's pad() would divide by zero, data_utils_SSL.py:62)")
    os.makedirs(directory, exist_ok=True)
    offsets = np.zeros(len(clips) + 1, dtype=np.int64)
    np.cumsum([c.size for c in clips], out=offsets[1:])
    np.save(os.path.join(directory, "pcm.npy"), np.concatenate(clips) if clips else np.zeros(0, np.int16))
    np.save(os.path.join(directory, "offsets.npy"), offsets)
    with open(os.path.join(directory, "utts.json"), "w") as f:
        json.dump(list(utt_ids), f)


def wav_files_to_shard(directory: str, utt_ids: Sequence[str], paths: Sequence[str], workers: int = 6) -> None:
    write_pcm_shard(directory, utt_ids, decode_wav_files(paths, workers))


def audio_files_to_shard(directory: str, utt_ids: Sequence[str], paths: Sequence[str], workers: int = 6,
                         max_samples: Optional[int] = 64600) -> None:
    """FLAC / WAV corpus -> PCM shard holding the head of every clip (``max_samples``; None keeps whole clips)."""
    write_pcm_shard(directory, utt_ids, decode_audio_files(paths, workers, max_samples))


class PcmShard:
    """Memory-mapped shard; ``batch(lo, hi)`` returns the three arrays ``slsb_score_pcm16_host`` takes."""

    def __init__(self, directory: str):
        self.pcm = np.load(os.path.join(directory, "pcm.npy"), mmap_mode="r")
        self.offsets = np.load(os.path.join(directory, "offsets.npy"))
        with open(os.path.join(directory, "utts.json")) as f:
            self.utt_ids: List[str] = json.load(f)
        if self.pcm.dtype != np.int16 or self.offsets.dtype != np.int64 or len(self.offsets) != len(self.utt_ids) + 1 \
                or int(self.offsets[-1]) != self.pcm.shape[0] or np.any(np.diff(self.offsets) <= 0):
            raise AudioFormatError(f"{directory}: inconsistent shard")

    def __len__(self) -> int:
        return len(self.utt_ids)

    def clip(self, i: int) -> np.ndarray:
        return np.asarray(self.pcm[self.offsets[i]:self.offsets[i + 1]])

    def batch(self, lo: int, hi: int, max_samples: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(pcm int16 [total], offsets int64 [B] relative to pcm, lens int32 [B]) of clips lo..hi-1.  With ``max_samples`` only
        the first ``max_samples`` of a longer clip are taken (pad() truncates to the head, data_utils_SSL.py:60-61), so a
        10-minute recording costs no more upload than a 4-second one."""
        starts, ends = self.offsets[lo:hi], self.offsets[lo + 1:hi + 1]
        lens = (ends - starts).astype(np.int64)
        if max_samples is not None:
            lens = np.minimum(lens, max_samples)
        if max_samples is None or np.all(ends - starts == lens):
            pcm = np.asarray(self.pcm[starts[0]:ends[-1]]) if hi > lo else np.zeros(0, np.int16)
            rel = (starts - starts[0]).astype(np.int64) if hi > lo else np.zeros(0, np.int64)
        else:
            pcm = np.concatenate([np.asarray(self.pcm[s:s + n]) for s, n in zip(starts, lens)])
            rel = np.zeros(hi - lo, dtype=np.int64)
            np.cumsum(lens[:-1], out=rel[1:])
        return np.ascontiguousarray(pcm), rel, lens.astype(np.int32)


def _score_batches(eng, head: int, prec: int, samples: int, batch: int, n_items: int, fetch) -> torch.Tensor:
    """Common loop of the file / shard scorers.  ``fetch(a, b)`` returns ``(pcm int16 [total], offsets int64 [b - a], lens int32
    [b - a])`` for items a..b-1.  Two pinned upload buffers: while the (synchronous, GIL-free) ``slsb_score_pcm16_host`` call of
    batch j runs, a helper thread fetches and stages batch j + 1 into the other buffer, so the device never waits for the host."""
    out = torch.empty(max(n_items, 0), dtype=torch.float32)
    pin = torch.cuda.is_available()
    stages = [torch.empty(batch * samples, dtype=torch.int16, pin_memory=pin) for _ in range(2)]
    starts = list(range(0, n_items, batch))

    def stage_batch(j):
        a = starts[j]
        b = min(a + batch, n_items)
        pcm, off, lens = fetch(a, b)
        up = stages[j & 1][:pcm.size]
        up.numpy()[:] = pcm                                       # the only host copy: decoder output / memory-mapped shard -> pinned buffer
        return a, b, up, torch.from_numpy(off), torch.from_numpy(lens)

    if not starts:
        return out
    with cf.ThreadPoolExecutor(max_workers=1) as ex:
        nxt = ex.submit(stage_batch, 0)
        for j in range(len(starts)):
            a, b, up, off, lens = nxt.result()
            if j + 1 < len(starts):
                nxt = ex.submit(stage_batch, j + 1)              # fills the other buffer while this batch is on the device
            out[a:b] = eng.score_pcm16_arrays(up, off, lens, head, prec, samples)
    return out


def score_pcm_shard(model, shard: PcmShard, batch: int = 64, samples: int = 64600, lo: int = 0, hi: Optional[int] = None) -> torch.Tensor:
    """Scores clips lo..hi-1 of a shard (a rank's ``shard_range``) with ``model`` (a ``Model`` / ``ModelSLS`` of this
    package): ``exp(logp[:, 1])`` per clip, float32 CPU tensor, protocol order (main.py:178-192)."""
    hi = len(shard) if hi is None else hi
    return _score_batches(model.engine(), model._head(), model._prec(), samples, batch, hi - lo,
                          lambda a, b: shard.batch(lo + a, lo + b, max_samples=samples))


def score_audio_files(model, paths: Sequence[str], batch: int = 64, samples: int = 64600, workers: int = 6, ahead: int = 4) -> torch.Tensor:
    """FLAC / WAV files -> scores without a shard on disk: what ``produce_evaluation_file`` does with its 6-worker DataLoader
    (main.py:161-193), as a pipeline - a decode pool runs ``ahead`` batches in front of the device, each batch is staged into a
    pinned buffer while the previous one is being scored, 2 bytes per sample of the un-padded heads are uploaded."""
    n = len(paths)
    pool = cf.ThreadPoolExecutor(max_workers=max(1, workers))
    futures = {}

    def submit_upto(j_last):
        for i in range(len(futures), min(n, (j_last + 1) * batch)):
            futures[i] = pool.submit(read_audio_pcm16, paths[i], samples)

    def fetch(a, b):
        submit_upto(b // batch + ahead)
        clips = [futures[i].result() for i in range(a, b)]
        for i in range(a, b):
            futures[i] = None                                     # drop the decoded clip, keep the index count
        if any(c.size == 0 for c in clips):
            raise AudioFormatError(f"empty clip in {paths[a:b]}")
        lens = np.array([c.size for c in clips], dtype=np.int32)
        off = np.zeros(len(clips), dtype=np.int64)
        np.cumsum(lens[:-1], out=off[1:])
        return np.concatenate(clips), off, lens

    try:
        submit_upto(ahead)
        return _score_batches(model.engine(), model._head(), model._prec(), samples, batch, n, fetch)
    finally:
        pool.shutdown(wait=False, cancel_futures=True)
