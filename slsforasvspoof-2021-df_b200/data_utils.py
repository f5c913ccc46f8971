"""Eval-side mirror of the reference's ``data_utils_SSL.py`` (same names, same return values) on the native decoders, so that
``main.py``'s evaluation branch (:640-650: ``genSpoof_list`` -> ``Dataset_ASVspoof2021_eval`` -> ``produce_evaluation_file``)
runs unchanged against this package.  ``__getitem__`` returns what the reference returns - ``(float32 Tensor [64600], utt_id)``
- but decodes with ``ingest.read_audio_float32`` (FLAC: csrc/flac_decode.cpp, head of the clip only) instead of librosa.

For throughput use ``ingest.audio_files_to_shard`` + ``score_pcm_shard`` (2 bytes per sample uploaded, ``pad`` on the device);
this module is the compatibility surface.  Training-side pieces (RawBoost, ``Dataset_ASVspoof2019_train``) are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .ingest import read_audio_float32
from .scoring import pad_clip


def genSpoof_list(dir_meta, is_train=False, is_eval=False):
    """data_utils_SSL.py:26-53.  Eval: the stripped line is the key.  Train / dev: ``(labels, keys)`` from 5-column protocol
    lines ``<speaker> <key> <-> <attack> <bonafide|spoof>`` with bonafide = 1 (:38)."""
    with open(dir_meta, "r") as f:
        lines = f.readlines()
    if is_eval and not is_train:
        return [ln.strip() for ln in lines]
    d_meta, file_list = {}, []
    for ln in lines:
        _, key, _, _, label = ln.strip().split()
        file_list.append(key)
        d_meta[key] = 1 if label == "bonafide" else 0
    return d_meta, file_list


def pad(x, max_len=64600):
    """data_utils_SSL.py:58-65."""
    return pad_clip(np.asarray(x), max_len)


class _EvalClips(Dataset):
    cut = 64600      # ~4 s (data_utils_SSL.py:103)

    def __init__(self, list_IDs, base_dir):
        self.list_IDs = list_IDs
        self.base_dir = base_dir

    def __len__(self):
        return len(self.list_IDs)

    def _path(self, utt_id):
        raise NotImplementedError

    def __getitem__(self, index):
        utt_id = self.list_IDs[index]
        # pad() only ever looks at the first `cut` samples; float32 mono exactly as libsndfile + librosa deliver it for 16-bit
        # audio at 16 kHz (multi-channel: float mean of the channels, not an integer one).  Other rates / bit depths raise
        # AudioFormatError: the reference would resample (librosa), this path never guesses - resample off-line.
        x = read_audio_float32(self._path(utt_id), max_samples=self.cut)
        return torch.from_numpy(pad_clip(x, self.cut)), utt_id


class Dataset_ASVspoof2021_eval(_EvalClips):
    """data_utils_SSL.py:96-115: ``<base_dir>/flac/<utt>.flac``."""

    def _path(self, utt_id):
        return self.base_dir + "/flac/" + utt_id + ".flac"


class Dataset_in_the_wild_eval(_EvalClips):
    """data_utils_SSL.py:118-135: ``<base_dir><utt>`` (the list carries file names with their extension)."""

    def _path(self, utt_id):
        return self.base_dir + utt_id
