"""Shared test helpers: special-token layout of the multilingual Whisper vocabulary,
synthetic model fixtures, and oracle <-> product option mapping."""
import numpy as np
import torch

from oracle import frontend as OF
from oracle.model import Dims, OracleWhisper
from tests.synth import speech_like


def special_tokens(n_vocab: int):
    s = 1 if n_vocab == 51866 else 0  # large-v3 adds <|yue|>
    n_lang = 100 if s else 99
    return dict(eot=50257, sot=50258, lang0=50259, n_lang=n_lang, translate=50358 + s, transcribe=50359 + s,
                sot_lm=50360 + s, sot_prev=50361 + s, no_speech=50362 + s, no_timestamps=50363 + s,
                timestamp_begin=50364 + s)


def decode_option_kwargs(n_vocab: int, beam_size=1, max_new_tokens=0, without_timestamps=False):
    sp = special_tokens(n_vocab)
    suppress = sorted(set([sp["sot"], sp["translate"], sp["transcribe"], sp["sot_lm"], sp["sot_prev"],
                           sp["no_speech"]] + list(range(sp["lang0"], sp["lang0"] + sp["n_lang"]))
                          + [1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92]))
    return dict(beam_size=beam_size, patience=1.0, length_penalty=1.0, max_new_tokens=max_new_tokens,
                suppress_blank=True, without_timestamps=without_timestamps, max_initial_timestamp_index=50,
                eot=sp["eot"], sot=sp["sot"], no_speech=sp["no_speech"], no_timestamps=sp["no_timestamps"],
                timestamp_begin=sp["timestamp_begin"], suppress_ids=suppress, blank_ids=[220, sp["eot"]])


def oracle_dims(md) -> Dims:
    return Dims(md.n_mels, md.d_model, md.n_heads, md.n_enc_layers, md.n_vocab, md.d_model, md.n_heads,
                md.n_dec_layers, md.n_audio_ctx, md.n_text_ctx)


def mel_windows(seeds, seconds, n_mels):
    """Oracle log-mel windows [B, n_mels, 3000] (zero-padded features) of synthetic clips."""
    mels = []
    for sd, sec in zip(seeds, seconds):
        m = OF.pad_or_trim(OF.log_mel(speech_like(sd, sec), n_mels)[:, :3000])
        mels.append(m)
    return np.stack(mels).astype(np.float32)


def windows_to_device(mel: np.ndarray, device) -> torch.Tensor:
    """[B, n_mels, 3000] fp32 -> bf16 [B, 3000, 128] time-major (what pack_windows produces)."""
    B, nm, T = mel.shape
    w = torch.zeros((B, T, 128), dtype=torch.float32)
    w[:, :, :nm] = torch.from_numpy(mel).permute(0, 2, 1)
    return w.to(torch.bfloat16).to(device).contiguous()
