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


# ------------------------------------------------------------------------------------------------------------------
# Device FLAC decode: the host scans a stream once (frame boundaries + CRC-8 + CRC-16, csrc/flac_decode.cpp::slsb_flac_scan),
# the GPU decodes one frame per thread (csrc/flac_gpu.cu).  FRAME_DTYPE is the C struct slsb_flac_frame (include/slsb200.h).
# ------------------------------------------------------------------------------------------------------------------
FRAME_DTYPE = np.dtype([("byte_off", "<i8"), ("out_off", "<i8"), ("byte_len", "<i4"), ("keep", "<i4"), ("bps", "<i4"), ("reserved", "<i4")])
assert FRAME_DTYPE.itemsize == 32


class FlacScan:
    """Frame table of one FLAC stream: ``device_ok`` says whether the device decoder takes it (16-bit mono at ``sample_rate``)."""
    __slots__ = ("data", "rate", "channels", "bps", "total", "samples", "off", "len", "blocks", "device_ok")


def scan_flac_bytes(data, max_samples: Optional[int] = 64600, sample_rate: Optional[int] = SAMPLE_RATE) -> FlacScan:
    """One pass over the bytes of a .flac file: STREAMINFO + the frames that cover its first ``max_samples`` samples, header CRC-8
    and frame CRC-16 verified, nothing decoded.  Raises ``AudioFormatError`` on a damaged stream."""
    import ctypes as C
    from ._lib import load
    lib = load()
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    cap = 64 if max_samples else 4096
    while True:
        info = (C.c_int32 * 8)()
        off = np.empty(cap, dtype=np.int64); ln = np.empty(cap, dtype=np.int32); blocks = np.empty(cap, dtype=np.int32)
        n = int(lib.slsb_flac_scan(buf.ctypes.data, buf.size, int(max_samples or 0), info, off.ctypes.data, ln.ctypes.data, blocks.ctypes.data, cap))
        if n == -11 and cap < (1 << 22):
            cap *= 8                                              # more frames than slots (small block sizes): retry with a larger table
            continue
        break
    if n < 0:
        raise AudioFormatError(f"FLAC: {FLAC_ERRORS.get(n, n)}")
    r = FlacScan()
    r.data, r.rate, r.channels, r.bps = buf, int(info[0]), int(info[1]), int(info[2])
    r.total = int(info[3]) | (int(info[5]) << 31)
    r.samples = int(info[6])
    r.off, r.len, r.blocks = off[:n], ln[:n], blocks[:n]
    if sample_rate is not None and r.rate != sample_rate:
        raise AudioFormatError(f"FLAC: need {sample_rate} Hz, got {r.rate} Hz")
    r.device_ok = r.channels == 1 and r.bps == 16 and r.samples > 0
    return r


def pack_flac_batch(scans: Sequence[FlacScan], max_samples: int = 64600):
    """Frame jobs of a batch of scanned streams: (bytes uint8 [nbytes] - only the frames that are needed, back to back -, frames
    FRAME_DTYPE [n_frames], total_samples, offsets int64 [B], lens int32 [B])."""
    chunks, jobs, offsets, lens = [], [], [], []
    byte_pos = 0
    sample_pos = 0
    for sc in scans:
        keep_total = min(sc.samples, max_samples)
        a, b = int(sc.off[0]), int(sc.off[-1] + sc.len[-1])
        chunks.append(sc.data[a:b])
        left = keep_total
        first = sample_pos
        for o, l, bs in zip(sc.off.tolist(), sc.len.tolist(), sc.blocks.tolist()):
            k = min(bs, left)
            jobs.append((byte_pos + (o - a), sample_pos, l, k, sc.bps, 0))
            sample_pos += k
            left -= k
        offsets.append(first)
        lens.append(keep_total)
        byte_pos += b - a
    frames = np.array(jobs, dtype=FRAME_DTYPE)
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return data, frames, sample_pos, np.array(offsets, dtype=np.int64), np.array(lens, dtype=np.int32)


def decode_flac_frames_host(data: np.ndarray, frames: np.ndarray, total_samples: int):
    """CPU twin of the device decoder (the same frame core compiled for the host): (pcm int16 [total_samples], status int32 [n])."""
    from ._lib import load
    lib = load()
    padded = np.zeros(data.size + 8, dtype=np.uint8)
    padded[:data.size] = data
    pcm = np.zeros(total_samples, dtype=np.int16)
    status = np.zeros(len(frames), dtype=np.int32)
    rc = lib.slsb_flac_decode_frames_host(padded.ctypes.data, np.ascontiguousarray(frames).ctypes.data, len(frames), pcm.ctypes.data, status.ctypes.data)
    if rc != 0:
        raise RuntimeError("slsb_flac_decode_frames_host failed")
    return pcm, status


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


def score_flac_files_device(model, paths: Sequence[str], batch: int = 64, samples: int = 64600, workers: int = 6, ahead: int = 4,
                            stats: Optional[dict] = None) -> torch.Tensor:
    """FLAC files -> scores with the decode ON THE DEVICE: the worker pool only reads and scans the files (frame table, CRCs: ~10x
    cheaper than decoding), a batch travels as its compressed frames (about half the bytes of its PCM) and is decoded one frame per
    GPU thread in front of the forward (``slsb_score_flac_host``).  A batch holding a stream the device decoder does not take
    (stereo, not 16-bit) - or one it refuses - is decoded by the host decoder instead; ``stats`` counts both kinds."""
    n = len(paths)
    eng, head, prec = model.engine(), model._head(), model._prec()
    out = torch.empty(n, dtype=torch.float32)
    if stats is not None:
        stats.update(device_batches=0, host_batches=0, flac_bytes=0, pcm_bytes=0)

    def load(p):
        with open(p, "rb") as f:
            return scan_flac_bytes(f.read(), samples)

    def stage(j):
        a, b = j * batch, min((j + 1) * batch, n)
        scans = [futures.pop(i).result() for i in range(a, b)]
        if all(s.device_ok for s in scans):
            data, frames, total, off, lens = pack_flac_batch(scans, samples)
            return a, b, scans, (torch.from_numpy(data), torch.from_numpy(frames.view(np.uint8).reshape(-1)), total, torch.from_numpy(off), torch.from_numpy(lens))
        return a, b, scans, None

    def host_decode(a, b, scans):
        clips = [decode_flac_bytes(s.data.tobytes(), samples) for s in scans]
        lens = np.array([c.size for c in clips], dtype=np.int32)
        off = np.zeros(len(clips), dtype=np.int64)
        np.cumsum(lens[:-1], out=off[1:])
        return eng.score_pcm16_arrays(torch.from_numpy(np.concatenate(clips)), torch.from_numpy(off), torch.from_numpy(lens), head, prec, samples)

    pool = cf.ThreadPoolExecutor(max_workers=max(1, workers))
    futures = {}
    submitted = 0
    nb = (n + batch - 1) // batch
    try:
        with cf.ThreadPoolExecutor(max_workers=1) as stager:
            def submit_upto(j_last):
                nonlocal submitted
                hi = min(n, (j_last + 1) * batch)
                while submitted < hi:
                    futures[submitted] = pool.submit(load, paths[submitted])
                    submitted += 1
            submit_upto(ahead)
            nxt = stager.submit(stage, 0) if nb else None
            for j in range(nb):
                a, b, scans, packed = nxt.result()
                submit_upto(j + 1 + ahead)
                if j + 1 < nb:
                    nxt = stager.submit(stage, j + 1)            # scan results of the next batch are packed while this one is on the device
                scores = None
                if packed is not None:
                    scores, _ = eng.score_flac_arrays(*packed, head, prec, samples)
                    if stats is not None and scores is not None:
                        stats["device_batches"] += 1
                        stats["flac_bytes"] += int(packed[0].numel())
                        stats["pcm_bytes"] += int(packed[2]) * 2
                if scores is None:
                    scores = host_decode(a, b, scans)
                    if stats is not None:
                        stats["host_batches"] += 1
                out[a:b] = scores
        return out
    finally:
        pool.shutdown(wait=False, cancel_futures=True)


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
