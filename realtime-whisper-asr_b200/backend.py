"""``B200WhisperASR``: the streaming-ASR backend class that sits beside ``FasterWhisperASR``.

Contract (upstream ``whisper_online.ASRBase``) as the reference uses it:
  * constructor signature of ``CustomFasterWhisperASR`` (/root/reference/asr_components.py:195-198);
  * ``transcribe(audio, init_prompt="") -> list[Segment]``  (asr_components.py:267-289: merges
    ``transcribe_kargs`` with ``adaptive_params.get_transcribe_kwargs()``, forwards ``language``,
    ``initial_prompt``, ``beam_size`` (default 5), ``temperature`` (default 0.0),
    ``word_timestamps=True``, ``condition_on_previous_text=True`` and drains the generator);
  * ``ts_words(segments)`` -> [(start, end, word)]   (asr_components.py:291-297; like the reference
    override it does NOT drop segments with no_speech_prob > 0.9, unlike upstream);
  * ``segments_end_ts(segments)`` -> [end]             (asr_components.py:299-301);
  * ``set_translate_task()``, ``use_vad()``, ``sep = ""``, attributes ``original_language``,
    ``transcribe_kargs``, ``logfile``, ``model``.
Callers: upstream ``OnlineASRProcessor.process_iter`` (reached from
/root/reference/enhanced_asr_processor.py:357) and ``self.asr.sep`` at :322,339.

Two insertion points:
  B1  instantiate ``B200WhisperASR`` where the reference builds ``CustomFasterWhisperASR``
      (一键实时识别麦克风.py:1125); ``make_b200_asr_class(ASRBase)`` gives a true subclass when
      ``whisper_online`` is importable.
  B2  ``install_as_faster_whisper()`` registers a ``faster_whisper`` module whose ``WhisperModel`` is
      ``B200WhisperModel``; the UNCHANGED ``create_custom_faster_whisper_asr`` then loads this
      backend through its own ``load_model`` (asr_components.py:232-265).
"""
from __future__ import annotations

import sys
import types
from typing import List, Optional, Tuple

import numpy as np

from .transcribe import B200WhisperModel


class B200WhisperASR:
    """Drop-in ASR backend (B1).  Words carry their own leading space, hence ``sep = ""``."""

    sep = ""

    def __init__(self, lan, modelsize=None, cache_dir=None, model_dir=None, device="cuda",
                 compute_type="float16", device_index=0, num_workers=1, cpu_threads=None,
                 logfile=sys.stderr, adaptive_params=None, transcribe_kwargs=None, model=None,
                 **model_kwargs):
        self.device = device
        self.compute_type = compute_type
        self.device_index = device_index
        self.num_workers = num_workers
        self.cpu_threads = cpu_threads
        self.logfile = logfile
        self.transcribe_kargs = transcribe_kwargs if transcribe_kwargs else {}
        self.adaptive_params = adaptive_params
        self.original_language = None if lan == "auto" else lan
        self._model_kwargs = model_kwargs
        # one engine may serve many stream processors: pass an existing model to share it
        self.model = model if model is not None else self.load_model(modelsize, cache_dir, model_dir)

    def load_model(self, modelsize=None, cache_dir=None, model_dir=None):
        if model_dir is not None:
            model_size_or_path = model_dir
        elif modelsize is not None:
            model_size_or_path = modelsize
        else:
            raise ValueError("modelsize or model_dir parameter must be set")
        if self.device != "cuda":
            # the reference falls back to CPU int8 here (一键实时识别麦克风.py:1147-1164); this backend
            # is CUDA-only and says so, so that the caller's own fallback logic can take over
            raise RuntimeError("B200WhisperASR runs on CUDA only (no CPU fallback)")
        return B200WhisperModel(model_size_or_path, device="cuda", device_index=self.device_index,
                                compute_type=self.compute_type, download_root=cache_dir,
                                num_workers=self.num_workers, **self._model_kwargs)

    def _kwargs(self) -> dict:
        if self.adaptive_params:
            return {**self.transcribe_kargs, **self.adaptive_params.get_transcribe_kwargs()}
        return self.transcribe_kargs

    def transcribe(self, audio, init_prompt=""):
        kw = self._kwargs()
        segments, info = self.model.transcribe(
            audio, language=self.original_language, initial_prompt=init_prompt,
            beam_size=kw.get("beam_size", 5), temperature=kw.get("temperature", 0.0),
            word_timestamps=True, condition_on_previous_text=True,
            **{k: v for k, v in kw.items() if k not in ("beam_size", "temperature")})
        return list(segments)

    def ts_words(self, segments) -> List[Tuple[float, float, str]]:
        return [(w.start, w.end, w.word) for s in segments for w in s.words]

    def segments_end_ts(self, segments) -> List[float]:
        return [s.end for s in segments]

    def set_translate_task(self):
        self.transcribe_kargs["task"] = "translate"

    def use_vad(self):
        self.transcribe_kargs["vad_filter"] = True


def make_b200_asr_class(ASRBase):
    """A subclass of the caller's ``whisper_online.ASRBase`` with B200WhisperASR's behaviour
    (mirrors ``create_custom_faster_whisper_asr(FasterWhisperASR)``, asr_components.py:182)."""
    return type("B200WhisperASR", (B200WhisperASR, ASRBase), {"sep": ""})


def install_as_faster_whisper() -> types.ModuleType:
    """B2: make ``from faster_whisper import WhisperModel`` resolve to the B200 engine."""
    from . import transcribe as T
    mod = types.ModuleType("faster_whisper")
    mod.WhisperModel = T.B200WhisperModel
    mod.__version__ = "b200"
    tr = types.ModuleType("faster_whisper.transcribe")
    tr.Segment, tr.Word, tr.TranscriptionInfo = T.Segment, T.Word, T.TranscriptionInfo
    mod.transcribe = tr
    sys.modules["faster_whisper"] = mod
    sys.modules["faster_whisper.transcribe"] = tr
    return mod
