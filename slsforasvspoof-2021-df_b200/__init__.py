"""B200-native XLS-R-300M + {TopK-SAE | window-TopK | SLS} scoring path (drop-in for the reference's
``Model(args, device).forward(x)`` behind ``main.py --is_eval``).

The directory name carries the reference's name (with hyphens), so import it through the alias module at the
repo root::

    import sls_b200
    model = sls_b200.Model(None, "cuda", cp_path=None).to("cuda").eval()

Everything numerical happens in ``libslsb200.so`` (``csrc/``, C ABI in ``include/slsb200.h``); nothing here
imports ``oracle/``.
"""
from ._lib import SlsbError, LIB_PATH, EXPORTED_SYMBOLS, load as load_library
from .engine import Engine, make_config, PRECISIONS, HEAD_NONE, HEAD_SAE, HEAD_WINDOW, HEAD_SLS, HEAD_RETAIN, PREC_FP32, PREC_BF16
from .weights import TrunkGeometry, TrunkParams, pack_state_dict, load_checkpoint_tensors, load_model_checkpoint, fix_module_prefix
from .model import Model, ModelWindowTopK, ModelSLS, SSLModel, AutoEncoderTopK, getAttenF
from .scoring import (produce_evaluation_file, score_synthetic_shard, gather_scores, write_score_file, pad_clip,
                      SyntheticEvalSet, shard_range, bucket_by_frames, batch_by_length, score_variable_length, compute_eer, read_score_file,
                      synth_clip_host)
from .data_utils import genSpoof_list, pad, Dataset_ASVspoof2021_eval, Dataset_in_the_wild_eval
from .ingest import (read_wav_pcm16, write_wav_pcm16, decode_wav_files, write_pcm_shard, wav_files_to_shard, PcmShard,
                     score_pcm_shard, AudioFormatError, decode_flac_bytes, read_flac_pcm16, read_audio_pcm16, read_audio_float32, decode_audio_files,
                     audio_files_to_shard, score_audio_files, score_flac_files_device, scan_flac_bytes, pack_flac_batch,
                     decode_flac_frames_host, FRAME_DTYPE)

__all__ = ["Model", "ModelWindowTopK", "ModelSLS", "SSLModel", "AutoEncoderTopK", "getAttenF", "Engine", "make_config",
           "TrunkGeometry", "TrunkParams", "pack_state_dict", "load_checkpoint_tensors", "load_model_checkpoint", "fix_module_prefix", "produce_evaluation_file", "score_synthetic_shard",
           "gather_scores", "write_score_file", "pad_clip", "SyntheticEvalSet", "shard_range", "bucket_by_frames", "batch_by_length",
           "score_variable_length", "compute_eer", "read_score_file", "synth_clip_host", "SlsbError",
           "load_library", "LIB_PATH", "EXPORTED_SYMBOLS", "PRECISIONS", "HEAD_NONE", "HEAD_SAE", "HEAD_WINDOW",
           "HEAD_SLS", "HEAD_RETAIN", "PREC_FP32", "PREC_BF16", "read_wav_pcm16", "write_wav_pcm16", "decode_wav_files", "write_pcm_shard",
           "wav_files_to_shard", "PcmShard", "score_pcm_shard", "AudioFormatError", "decode_flac_bytes", "read_flac_pcm16",
           "read_audio_pcm16", "read_audio_float32", "decode_audio_files", "audio_files_to_shard", "score_audio_files", "score_flac_files_device", "scan_flac_bytes", "pack_flac_batch", "decode_flac_frames_host", "FRAME_DTYPE", "genSpoof_list", "pad", "Dataset_ASVspoof2021_eval",
           "Dataset_in_the_wild_eval"]
