"""``B200WhisperModel``: the object the reference's backend wrapper holds as ``self.model``.

Shaped like ``faster_whisper.WhisperModel`` for exactly the surface the reference uses:
``create_custom_faster_whisper_asr`` builds it in ``load_model`` and calls
``self.model.transcribe(audio, language=, initial_prompt=, beam_size=, temperature=,
word_timestamps=True, condition_on_previous_text=True, **kw) -> (segments, info)``
(/root/reference/asr_components.py:232-289); it then reads ``Segment.words[*].{start,end,word}``,
``Segment.end`` and (upstream ``FasterWhisperASR.ts_words``) ``Segment.no_speech_prob``.

The pipeline below restates faster-whisper's ``transcribe`` / ``generate_segments`` (SURVEY.md
A.2): features -> 30 s windows from ``seek`` -> prompt -> generate -> no-speech skip -> split by
timestamp tokens -> word timestamps (align + DTW, duration heuristics, punctuation merge) ->
``Segment`` / ``Word``.  All tensor work runs in libb200whisper.so; what stays here is token-list
and string bookkeeping.  ``transcribe_batch`` runs many independent streams through the same
steps with one encoder / decoder / align launch per round (streams are the natural shard).
"""
from __future__ import annotations

import logging
import os
import zlib
from dataclasses import dataclass, field
from typing import Iterable, List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, frontend
from .engine import DecodeOptions, WhisperEngine
from .tokenizer import Tokenizer
from .weights import MODEL_DIMS, ModelDims, alignment_heads_for, pack_for_engine, synthetic_state_dict

_log = logging.getLogger("b200w")

SAMPLE_RATE = 16000
HOP = 160
N_FRAMES = 3000
FRAMES_PER_SECOND = 100
TOKENS_PER_SECOND = 50
TIME_PRECISION = 0.02
INPUT_STRIDE = 2
MAX_LENGTH = 448


class Word(NamedTuple):
    start: float
    end: float
    word: str
    probability: float


class Segment(NamedTuple):
    id: int
    seek: int
    start: float
    end: float
    text: str
    tokens: List[int]
    avg_logprob: float
    compression_ratio: float
    no_speech_prob: float
    words: Optional[List[Word]]
    temperature: float = 0.0


class TranscriptionOptions(NamedTuple):
    beam_size: int
    best_of: int
    patience: float
    length_penalty: float
    repetition_penalty: float
    no_repeat_ngram_size: int
    log_prob_threshold: Optional[float]
    no_speech_threshold: Optional[float]
    compression_ratio_threshold: Optional[float]
    condition_on_previous_text: bool
    prompt_reset_on_temperature: float
    temperatures: List[float]
    initial_prompt: Optional[Union[str, Iterable[int]]]
    prefix: Optional[str]
    suppress_blank: bool
    suppress_tokens: Optional[List[int]]
    without_timestamps: bool
    max_initial_timestamp: float
    word_timestamps: bool
    prepend_punctuations: str
    append_punctuations: str
    max_new_tokens: Optional[int]


class TranscriptionInfo(NamedTuple):
    language: str
    language_probability: float
    duration: float
    duration_after_vad: float
    all_language_probs: Optional[List[Tuple[str, float]]]
    transcription_options: TranscriptionOptions
    vad_options: None = None


@dataclass
class _Stream:
    """Per-stream state of the seek loop (faster-whisper generate_segments)."""
    sid: int
    content_frames: int
    tokenizer: Tokenizer
    options: TranscriptionOptions
    seek: int = 0
    idx: int = 0
    all_tokens: List[int] = field(default_factory=list)
    prompt_reset_since: int = 0
    last_speech_timestamp: float = 0.0
    segments: List[Segment] = field(default_factory=list)
    # per-round scratch
    segment_size: int = 0
    time_offset: float = 0.0
    prompt: List[int] = field(default_factory=list)


def get_compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def get_prompt(tokenizer: Tokenizer, previous_tokens: List[int], without_timestamps: bool = False,
               prefix: Optional[str] = None) -> List[int]:
    prompt: List[int] = []
    if previous_tokens:
        prompt.append(tokenizer.sot_prev)
        prompt.extend(previous_tokens[-(MAX_LENGTH // 2 - 1):])
    prompt.extend(tokenizer.sot_sequence)
    if without_timestamps:
        prompt.append(tokenizer.no_timestamps)
    if prefix:
        prefix_tokens = tokenizer.encode(" " + prefix.strip())
        if len(prefix_tokens) >= MAX_LENGTH // 2:
            prefix_tokens = prefix_tokens[: MAX_LENGTH // 2 - 1]
        if not without_timestamps:
            prompt.append(tokenizer.timestamp_begin)
        prompt.extend(prefix_tokens)
    return prompt


def split_segments_by_timestamps(tokenizer: Tokenizer, tokens: List[int], time_offset: float, segment_size: int,
                                 segment_duration: float, seek: int):
    tb = tokenizer.timestamp_begin
    current_segments = []
    single_timestamp_ending = len(tokens) >= 2 and tokens[-2] < tb <= tokens[-1]
    consecutive = [i for i in range(len(tokens)) if i > 0 and tokens[i] >= tb and tokens[i - 1] >= tb]
    if consecutive:
        slices = list(consecutive)
        if single_timestamp_ending:
            slices.append(len(tokens))
        last_slice = 0
        for current_slice in slices:
            sliced = tokens[last_slice:current_slice]
            start_pos = sliced[0] - tb
            end_pos = sliced[-1] - tb
            current_segments.append(dict(seek=seek, start=time_offset + start_pos * TIME_PRECISION,
                                         end=time_offset + end_pos * TIME_PRECISION, tokens=sliced))
            last_slice = current_slice
        if single_timestamp_ending:
            seek += segment_size  # no speech after the last timestamp
        else:
            seek += (tokens[last_slice - 1] - tb) * INPUT_STRIDE  # resume at the last timestamp
    else:
        duration = segment_duration
        timestamps = [t for t in tokens if t >= tb]
        if timestamps and timestamps[-1] != tb:
            duration = (timestamps[-1] - tb) * TIME_PRECISION
        current_segments.append(dict(seek=seek, start=time_offset, end=time_offset + duration, tokens=tokens))
        seek += segment_size
    return current_segments, seek, single_timestamp_ending


def merge_punctuations(alignment: List[dict], prepended: str, appended: str) -> None:
    i, j = len(alignment) - 2, len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous["word"].startswith(" ") and previous["word"].strip() in prepended:
            following["word"] = previous["word"] + following["word"]
            following["tokens"] = previous["tokens"] + following["tokens"]
            previous["word"] = ""
            previous["tokens"] = []
        else:
            j = i
        i -= 1
    i, j = 0, 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous["word"].endswith(" ") and following["word"] in appended:
            previous["word"] = previous["word"] + following["word"]
            previous["tokens"] = previous["tokens"] + following["tokens"]
            following["word"] = ""
            following["tokens"] = []
        else:
            i = j
        j += 1


def words_from_alignment(tokenizer: Tokenizer, text_tokens: List[int], text_indices: np.ndarray,
                         time_indices: np.ndarray, text_token_probs: np.ndarray) -> List[dict]:
    """faster-whisper find_alignment, after the device returned the DTW path."""
    words, word_tokens = tokenizer.split_to_word_tokens(text_tokens + [tokenizer.eot])
    if len(word_tokens) <= 1:
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    if len(word_boundaries) <= 1 or len(text_indices) == 0:
        return []
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    # (plain Python sums: np.mean on 1-3 element slices costs ~7 us per word, ~1600 words per bench step)
    tp = text_token_probs.tolist()
    probs = [(sum(tp[i:j]) / (j - i)) if j > i else float("nan")
             for i, j in zip(word_boundaries[:-1].tolist(), word_boundaries[1:].tolist())]
    return [dict(word=w, tokens=t, start=float(s), end=float(e), probability=p)
            for w, t, s, e, p in zip(words, word_tokens, start_times, end_times, probs)]


def attach_word_timestamps(tokenizer: Tokenizer, subsegments: List[dict], alignment: List[dict],
                           prepend_punctuations: str, append_punctuations: str,
                           last_speech_timestamp: float) -> float:
    """faster-whisper add_word_timestamps for one window (duration heuristics + punctuation)."""
    text_tokens_per_sub = [[t for t in s["tokens"] if t < tokenizer.eot] for s in subsegments]
    word_durations = np.array([w["end"] - w["start"] for w in alignment])
    word_durations = word_durations[word_durations.nonzero()]
    median_duration = float(np.median(word_durations)) if len(word_durations) > 0 else 0.0
    median_duration = min(0.7, median_duration)
    max_duration = median_duration * 2
    if len(word_durations) > 0:
        marks = ".。!！?？"
        for i in range(1, len(alignment)):
            if alignment[i]["end"] - alignment[i]["start"] > max_duration:
                if alignment[i]["word"] in marks:
                    alignment[i]["end"] = alignment[i]["start"] + max_duration
                elif alignment[i - 1]["word"] in marks:
                    alignment[i]["start"] = alignment[i]["end"] - max_duration
    merge_punctuations(alignment, prepend_punctuations, append_punctuations)
    word_index = 0
    time_offset = subsegments[0]["seek"] / FRAMES_PER_SECOND if subsegments else 0.0
    for si, sub in enumerate(subsegments):
        saved_tokens = 0
        words: List[dict] = []
        while word_index < len(alignment) and saved_tokens < len(text_tokens_per_sub[si]):
            timing = alignment[word_index]
            if timing["word"]:
                words.append(dict(word=timing["word"], start=round(time_offset + timing["start"], 2),
                                  end=round(time_offset + timing["end"], 2), probability=timing["probability"]))
            saved_tokens += len(timing["tokens"])
            word_index += 1
        if words:
            # the first words after a pause must not be longer than twice the median duration
            if words[0]["end"] - last_speech_timestamp > median_duration * 4 and (
                    words[0]["end"] - words[0]["start"] > max_duration
                    or (len(words) > 1 and words[1]["end"] - words[0]["start"] > max_duration * 2)):
                if len(words) > 1 and words[1]["end"] - words[1]["start"] > max_duration:
                    boundary = max(words[1]["end"] / 2, words[1]["end"] - max_duration)
                    words[0]["end"] = words[1]["start"] = boundary
                words[0]["start"] = max(0, words[0]["end"] - max_duration)
            # prefer the segment-level timestamps when the edge words are too long
            if sub["start"] < words[0]["end"] and sub["start"] - 0.5 > words[0]["start"]:
                words[0]["start"] = max(0, min(words[0]["end"] - median_duration, sub["start"]))
            else:
                sub["start"] = words[0]["start"]
            if sub["end"] > words[-1]["start"] and sub["end"] + 0.5 < words[-1]["end"]:
                words[-1]["end"] = max(words[-1]["start"] + median_duration, sub["end"])
            else:
                sub["end"] = words[-1]["end"]
            last_speech_timestamp = sub["end"]
        sub["words"] = words
    return last_speech_timestamp


class B200WhisperModel:
    """Drop-in for ``faster_whisper.WhisperModel`` on the surface the reference uses."""

    def __init__(self, model_size_or_path: str = "base", device: str = "cuda", device_index: int = 0,
                 compute_type: str = "bfloat16", cpu_threads: int = 0, num_workers: int = 1,
                 download_root: Optional[str] = None, local_files_only: bool = False,
                 state_dict=None, dims: Optional[ModelDims] = None, tokenizer_json: Optional[str] = None,
                 alignment_heads: Optional[Sequence[Tuple[int, int]]] = None, max_batch: int = 8,
                 max_beam: int = 5, seed: int = 0, synthetic: Optional[bool] = None, lanes: int = 1, **_ignored):
        """``model_size_or_path``: a model directory (CTranslate2 ``model.bin`` or ``model.safetensors``, see
        loader.py) or a size name that is looked up under ``download_root`` / the Hugging Face cache the way
        faster-whisper does.  There is no network: a size name without a checkpoint on disk is an ERROR unless
        ``synthetic=True`` (or the environment variable B200W_SYNTHETIC_WEIGHTS=1) asks for seeded random weights of
        that architecture with the synthetic byte vocabulary -- what tests and bench.py use, never a transcript.

        ``lanes`` > 1: a batch is cut into that many sub-batches that run CONCURRENTLY, each on its own engine
        (workspaces, KV cache, captured decode graphs; the weights are shared) and CUDA stream, driven by its own host
        thread.  A decode step is a chain of dependent launches (latency-bound), its cross attention is HBM-bound and
        the encoder tensor-bound: independent lanes fill each other's bubbles.  Streams are independent and every
        kernel is batch-invariant, so the results are identical to lanes=1."""
        if device not in ("cuda", "auto"):
            raise _lib.B200WError(f"device={device!r}: the B200 backend runs on CUDA only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise _lib.B200WError("CUDA device required: the B200 backend has no CPU fallback")
        self.device = torch.device("cuda", device_index if isinstance(device_index, int) else device_index[0])
        if compute_type not in ("bfloat16", "default", "auto"):
            _log.warning("compute_type=%r requested; this backend computes in bfloat16 (fp32 accumulation)", compute_type)
        self.compute_type = "bfloat16"  # the only arithmetic this backend implements
        self.tokenizer_json = tokenizer_json
        self.suppress_ids: Optional[List[int]] = None        # from a CTranslate2 config.json, when present
        self.suppress_ids_begin: Optional[List[int]] = None
        self.synthetic = False
        name = model_size_or_path
        if synthetic is None:
            synthetic = os.environ.get("B200W_SYNTHETIC_WEIGHTS", "") not in ("", "0")
        if state_dict is None:
            from .loader import load_model_dir, resolve_model_dir
            found = resolve_model_dir(name, download_root)
            if found is not None:
                loaded = load_model_dir(found)
                state_dict, dims = loaded.state_dict, loaded.dims
                self.tokenizer_json = self.tokenizer_json or loaded.tokenizer_json
                alignment_heads = alignment_heads or loaded.alignment_heads
                self.suppress_ids, self.suppress_ids_begin = loaded.suppress_ids, loaded.suppress_ids_begin
                _log.info("loaded %s from %s", loaded.source, found)
            else:
                base = str(name).replace(".en", "")
                if base not in MODEL_DIMS:
                    raise ValueError(f"{name!r} is neither a model directory nor a known model size")
                if not synthetic:
                    raise FileNotFoundError(
                        f"no checkpoint for model {name!r}: not a directory, nothing under download_root="
                        f"{download_root!r} or the Hugging Face cache, and there is no network to download one.  "
                        "Pass a CTranslate2 / safetensors model directory, or synthetic=True for seeded random "
                        "weights (tests and benchmarks only).")
                dims = MODEL_DIMS[base]
                # seeded synthetic weights of the exact architecture (large models are generated on the device:
                # ~1.5 G random numbers; the CUDA and CPU generators give different values, so parity tests pass
                # CPU-made state dicts)
                gen_dev = str(self.device) if dims.d_model >= 1024 else "cpu"
                state_dict = synthetic_state_dict(dims, seed=seed, device=gen_dev)
                self.synthetic = True
                _log.warning("model %r: SYNTHETIC random weights and byte vocabulary (no checkpoint) -- output is "
                             "well-formed but meaningless", name)
        else:
            self.synthetic = tokenizer_json is None
        assert dims is not None
        self.dims = dims
        self.alignment_heads = list(alignment_heads) if alignment_heads else alignment_heads_for(dims)
        self.is_multilingual = dims.multilingual
        self.n_lanes = max(1, min(int(lanes), max_batch))
        self.lane_batch = -(-max_batch // self.n_lanes)
        with torch.cuda.device(self.device):
            packed = pack_for_engine(state_dict, dims, self.device)
            self.engines = [WhisperEngine(dims, packed, max_batch=self.lane_batch, max_beam=max_beam, device=self.device)
                            for _ in range(self.n_lanes)]
            self.lane_streams = [torch.cuda.Stream(self.device) for _ in range(self.n_lanes)] if self.n_lanes > 1 else []
        self.engine = self.engines[0]
        self._pool = None
        self._sample_seed = 0  # advances per decode call: draws differ between calls, runs are reproducible
        self.max_batch = max_batch
        self.max_beam = max_beam
        self.last_timings = {}
        self.profile_phases = False
        self._vocab = None            # loaded once: tokenizer.json is a multi-MB parse
        self._tokenizers = {}         # (language, task) -> Tokenizer
        self._warned = set()

    def _warn_once(self, key: str, msg: str) -> None:
        if key not in self._warned:
            self._warned.add(key)
            _log.warning(msg)

    def tokenizer(self, language: Optional[str] = "en", task: str = "transcribe") -> Tokenizer:
        """Cached per (language, task); the vocabulary object is shared by all of them."""
        key = (language, task)
        tk = self._tokenizers.get(key)
        if tk is None:
            tk = Tokenizer(self.dims.n_vocab, self.is_multilingual, task=task, language=language,
                           tokenizer_json=self.tokenizer_json, vocab=self._vocab)
            self._vocab = tk.vocab
            self._tokenizers[key] = tk
        return tk

    @property
    def supported_languages(self) -> List[str]:
        return self.tokenizer("en").languages if self.is_multilingual else ["en"]

    # ------------------------------------------------------------------ public API
    def transcribe(self, audio: Union[np.ndarray, torch.Tensor], language: Optional[str] = None,
                   task: str = "transcribe", **kw):
        results = self.transcribe_batch([audio], languages=[language], task=task, **kw)
        segments, info = results[0]
        return iter(segments), info

    def transcribe_batch(self, audios: Sequence[np.ndarray], languages: Optional[Sequence[Optional[str]]] = None,
                         task: str = "transcribe", beam_size: int = 5, best_of: int = 5, patience: float = 1,
                         length_penalty: float = 1, repetition_penalty: float = 1, no_repeat_ngram_size: int = 0,
                         temperature: Union[float, Sequence[float]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
                         compression_ratio_threshold: Optional[float] = 2.4,
                         log_prob_threshold: Optional[float] = -1.0, no_speech_threshold: Optional[float] = 0.6,
                         condition_on_previous_text: bool = True, prompt_reset_on_temperature: float = 0.5,
                         initial_prompt: Union[None, str, Iterable[int]] = None,
                         initial_prompts: Optional[Sequence[Union[None, str, Iterable[int]]]] = None,
                         prefix: Optional[str] = None, suppress_blank: bool = True,
                         suppress_tokens: Optional[List[int]] = (-1,), without_timestamps: bool = False,
                         max_initial_timestamp: float = 1.0, word_timestamps: bool = False,
                         prepend_punctuations: str = "\"'“¿([{-", append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
                         vad_filter: bool = False, max_new_tokens: Optional[int] = None,
                         device_audio=None, max_rounds: Optional[int] = None, **unsupported):
        """Transcribe independent clips concurrently; returns [(segments, info)] per clip.

        ``initial_prompt`` has faster-whisper's meaning (a string, or token ids) and applies to every clip;
        ``initial_prompts`` gives one such value per clip.  ``device_audio=(audio_dev, offsets)`` skips the
        host->device copy: the clips are already concatenated in one fp32 device tensor (``audios`` is then only
        used for its length)."""
        if vad_filter:
            raise NotImplementedError("vad_filter: the VAD gate sits in the caller (Silero in the reference)")
        if repetition_penalty != 1 or no_repeat_ngram_size != 0:
            raise NotImplementedError("repetition_penalty / no_repeat_ngram_size are not implemented")
        temps = [float(temperature)] if isinstance(temperature, (int, float)) else [float(t) for t in temperature]
        if not temps:
            temps = [0.0]
        if best_of > self.max_beam and any(t > 0 for t in temps):
            self._warn_once("best_of", f"best_of {best_of} exceeds this engine's max_beam {self.max_beam}: clamped")
            best_of = self.max_beam
        if beam_size > self.max_beam:
            self._warn_once("beam", f"beam_size {beam_size} exceeds this engine's max_beam {self.max_beam}: clamped "
                                    "(build the model with max_beam=... to raise it, at most 8)")
            beam_size = self.max_beam
        n = len(audios) if device_audio is None else len(device_audio[1]) - 1
        languages = list(languages) if languages is not None else [None] * n
        if initial_prompts is not None:
            if len(initial_prompts) != n:
                raise ValueError("initial_prompts must hold one entry per clip")
            prompts_in = list(initial_prompts)
        else:
            prompts_in = [initial_prompt] * n
        if device_audio is None:
            audios = [np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1)) for a in audios]
            durations = [len(a) / SAMPLE_RATE for a in audios]
        else:
            durations = [float(x) / SAMPLE_RATE for x in np.diff(device_audio[1])]
        options = TranscriptionOptions(beam_size, best_of, patience, length_penalty, repetition_penalty,
                                       no_repeat_ngram_size, log_prob_threshold, no_speech_threshold,
                                       compression_ratio_threshold, condition_on_previous_text,
                                       prompt_reset_on_temperature, temps, None, prefix, suppress_blank,
                                       list(suppress_tokens) if suppress_tokens is not None else None,
                                       without_timestamps, max_initial_timestamp, word_timestamps,
                                       prepend_punctuations, append_punctuations, max_new_tokens)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.device(self.device):
            ev[0].record()
            if device_audio is None:
                audio_dev, offsets = frontend.upload_audio(audios, self.device)
            else:
                audio_dev, offsets = device_audio
            mel = frontend.log_mel(audio_dev, offsets, self.dims.n_mels)
            ev[1].record()
            # language detection on the first window of the clips that need it (faster-whisper A9)
            lang_info = [(languages[i], 1.0, None) for i in range(n)]
            need = [i for i in range(n) if languages[i] is None]
            if need:
                if not self.is_multilingual:
                    for i in need:
                        lang_info[i] = ("en", 1.0, None)
                else:
                    det = self._detect_language(mel, need)
                    for i, d in zip(need, det):
                        lang_info[i] = d
            streams: List[_Stream] = []
            for i in range(n):
                tk = self.tokenizer(lang_info[i][0], task)
                st = _Stream(i, mel.n_frames[i] - 1, tk, options)
                ip = prompts_in[i]
                if ip is not None:
                    if isinstance(ip, str):
                        st.all_tokens.extend(tk.encode(" " + ip.strip()))
                    else:
                        st.all_tokens.extend(int(t) for t in ip)
                streams.append(st)
            self._run_streams(mel, streams, max_rounds)
        out = []
        for i, st in enumerate(streams):
            info = TranscriptionInfo(lang_info[i][0], float(lang_info[i][1]), durations[i], durations[i],
                                     lang_info[i][2], options)
            out.append((st.segments, info))
        return out

    def detect_language(self, audio: np.ndarray):
        with torch.cuda.device(self.device):
            audio_dev, offsets = frontend.upload_audio([np.asarray(audio, dtype=np.float32)], self.device)
            mel = frontend.log_mel(audio_dev, offsets, self.dims.n_mels)
            lang, prob, allp = self._detect_language(mel, [0])[0]
        return lang, prob, allp

    # ------------------------------------------------------------------ internals
    def _detect_language(self, mel: frontend.MelBatch, which: List[int]):
        tk = self.tokenizer("en")
        out = []
        for c0 in range(0, len(which), self.lane_batch):
            ids = which[c0:c0 + self.lane_batch]
            windows = frontend.pack_windows(mel, ids, [0] * len(ids))
            enc = self.engine.encode(windows)
            ckv = self.engine.cross_kv(enc)
            probs = self.engine.detect_language(ckv, tk.sot, tk.language_ids)
            for p in probs:
                order = np.argsort(-p, kind="stable")
                allp = [(tk.languages[j], float(p[j])) for j in order]
                out.append((allp[0][0], allp[0][1], allp))
        return out

    def _decode_options(self, tk: Tokenizer, o: TranscriptionOptions, temperature: float = 0.0,
                        seed: int = 0) -> DecodeOptions:
        """Decoder options of a batch.  The suppress sets depend on the vocabulary, the model's special-token layout
        and the call's ``suppress_tokens`` -- never on a stream's language or task -- so one set serves the batch.
        A CTranslate2 config.json's ``suppress_ids`` / ``suppress_ids_begin`` take the place of the lists derived
        from the vocabulary (that is what faster-whisper feeds CTranslate2 for suppress_tokens=[-1])."""
        if o.suppress_tokens is None:
            suppress = []
        elif self.suppress_ids is not None and -1 in o.suppress_tokens:
            suppress = sorted(set([t for t in o.suppress_tokens if t >= 0] + list(self.suppress_ids) +
                                  [tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm, tk.no_speech]))
        else:
            suppress = tk.suppressed_tokens(o.suppress_tokens)
        blank = list(self.suppress_ids_begin) if self.suppress_ids_begin is not None else tk.blank_tokens()
        return DecodeOptions(
            beam_size=o.beam_size, patience=o.patience, length_penalty=o.length_penalty,
            max_new_tokens=o.max_new_tokens or 0, suppress_blank=o.suppress_blank,
            without_timestamps=o.without_timestamps,
            max_initial_timestamp_index=int(round(o.max_initial_timestamp / TIME_PRECISION)),
            eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech, no_timestamps=tk.no_timestamps,
            timestamp_begin=tk.timestamp_begin,
            suppress_ids=suppress, blank_ids=blank,
            temperature=temperature, best_of=max(1, min(o.best_of, self.max_beam)), seed=seed)

    def _run_streams(self, mel: frontend.MelBatch, streams: List[_Stream], max_rounds: Optional[int] = None) -> None:
        """Rounds of (pack windows -> encode -> cross K/V -> decode -> align) over the live streams.
        ``max_rounds`` (profiling aid) stops early."""
        rounds = 0
        while True:
            live = [s for s in streams if s.seek < s.content_frames]
            if not live or (max_rounds is not None and rounds >= max_rounds):
                return
            rounds += 1
            if self.n_lanes == 1 or self.profile_phases or len(live) == 1:
                for c0 in range(0, len(live), self.lane_batch):
                    self._round(mel, live[c0:c0 + self.lane_batch], self.engines[0])
                continue
            for c0 in range(0, len(live), self.max_batch):
                self._round_lanes(mel, live[c0:c0 + self.max_batch])

    def _round_lanes(self, mel: frontend.MelBatch, group: List[_Stream]) -> None:
        """One round of `group` (<= max_batch streams) as n_lanes concurrent sub-batches."""
        from concurrent.futures import ThreadPoolExecutor
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.n_lanes, thread_name_prefix="b200w-lane")
        n = min(self.n_lanes, len(group))
        per = -(-len(group) // n)
        parts = [group[i * per:(i + 1) * per] for i in range(n)]
        parts = [p for p in parts if p]
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)  # the log-mel features were produced on the caller's stream

        def run(i: int):
            with torch.cuda.device(self.device), torch.cuda.stream(self.lane_streams[i]):
                self.lane_streams[i].wait_event(ready)
                self._round(mel, parts[i], self.engines[i])
                done = torch.cuda.Event()
                done.record(self.lane_streams[i])
                return done

        futures = [self._pool.submit(run, i) for i in range(len(parts))]
        for f in futures:
            main.wait_event(f.result())  # re-raises a lane's exception in the caller

    def _phase(self, name: str, t0: float) -> float:
        """Accumulate wall time per phase when profiling is on (synchronises the device)."""
        if not self.profile_phases:
            return t0
        import time as _t
        torch.cuda.synchronize(self.device)
        now = _t.perf_counter()
        self.last_timings[name] = self.last_timings.get(name, 0.0) + (now - t0)
        return now

    def _decode_with_fallback(self, engine: WhisperEngine, ckv: torch.Tensor, batch: List["_Stream"],
                              o: TranscriptionOptions):
        """faster-whisper's generate_with_fallback over a batch: decode every chunk at the first temperature; the
        chunks whose result is too repetitive (compression ratio) or too unlikely (average log-probability) -- and is
        not silence -- are decoded again at the next temperature (sampling, best_of samples), and so on; a chunk that
        fails everywhere keeps its most likely result among those below the compression threshold (or among all)."""
        n = len(batch)
        final: List = [None] * n
        temp_of = [o.temperatures[0]] * n
        history: List[List] = [[] for _ in range(n)]  # per chunk: (result, avg_logprob, compression_ratio, temperature)
        pending = list(range(n))
        for ti, temp in enumerate(o.temperatures):
            if not pending:
                break
            if len(pending) == n:
                sub_ckv, prompts = ckv, [s.prompt for s in batch]
            else:
                idx = torch.tensor(pending, dtype=torch.long, device=ckv.device)
                sub_ckv, prompts = ckv.index_select(0, idx), [batch[i].prompt for i in pending]
            self._sample_seed += 1
            res = engine.decode(sub_ckv, prompts, self._decode_options(self.tokenizer("en"), o, temp, self._sample_seed))
            still = []
            for i, r in zip(pending, res):
                tk = batch[i].tokenizer
                avg_logprob = r.sum_logprob / (len(r.tokens) + 1)
                text = tk.decode(r.tokens).strip()
                cr = get_compression_ratio(text) if text else 0.0
                history[i].append((r, avg_logprob, cr, temp))
                needs_fallback = False
                if o.compression_ratio_threshold is not None and cr > o.compression_ratio_threshold:
                    needs_fallback = True
                if o.log_prob_threshold is not None and avg_logprob < o.log_prob_threshold:
                    needs_fallback = True
                if (o.no_speech_threshold is not None and r.no_speech_prob > o.no_speech_threshold
                        and o.log_prob_threshold is not None and avg_logprob < o.log_prob_threshold):
                    needs_fallback = False  # silence
                if needs_fallback and ti + 1 < len(o.temperatures):
                    still.append(i)
                elif needs_fallback:  # every temperature failed
                    ok = [h for h in history[i]
                          if o.compression_ratio_threshold is None or h[2] <= o.compression_ratio_threshold]
                    best = max(ok or history[i], key=lambda h: h[1])
                    final[i], temp_of[i] = best[0], best[3]
                else:
                    final[i], temp_of[i] = r, temp
            pending = still
        return final, temp_of

    def _round(self, mel: frontend.MelBatch, batch: List[_Stream], engine: WhisperEngine) -> None:
        import time as _t
        o = batch[0].options
        t = self._phase("_", _t.perf_counter())
        for s in batch:
            s.time_offset = s.seek * HOP / SAMPLE_RATE
            s.segment_size = min(N_FRAMES, s.content_frames - s.seek)
            s.prompt = get_prompt(s.tokenizer, s.all_tokens[s.prompt_reset_since:], o.without_timestamps, o.prefix)
        windows = frontend.pack_windows(mel, [s.sid for s in batch], [s.seek for s in batch])
        enc = engine.encode(windows)
        t = self._phase("encode", t)
        ckv = engine.cross_kv(enc)
        t = self._phase("cross_kv", t)
        # one set of decoder options per batch: they do not depend on a stream's language or task (see there)
        results, used_temp = self._decode_with_fallback(engine, ckv, batch, o)
        t = self._phase("decode", t)
        to_align = []  # (stream, subsegments, text_tokens, row in ckv)
        per_stream = []
        for row, (s, res) in enumerate(zip(batch, results)):
            tk = s.tokenizer
            tokens = res.tokens
            seq_len = len(tokens)
            avg_logprob = res.sum_logprob / (seq_len + 1)
            text = tk.decode(tokens).strip()
            compression_ratio = get_compression_ratio(text) if text else 0.0
            if o.no_speech_threshold is not None:
                should_skip = res.no_speech_prob > o.no_speech_threshold
                if o.log_prob_threshold is not None and avg_logprob > o.log_prob_threshold:
                    should_skip = False
                if should_skip:
                    s.seek += s.segment_size
                    per_stream.append(None)
                    continue
            segment_duration = s.segment_size * HOP / SAMPLE_RATE
            subs, new_seek, single_ending = split_segments_by_timestamps(tk, tokens, s.time_offset, s.segment_size,
                                                                        segment_duration, s.seek)
            per_stream.append((s, subs, new_seek, single_ending, avg_logprob, compression_ratio, res.no_speech_prob,
                               used_temp[row]))
            if o.word_timestamps:
                text_tokens = [t for sub in subs for t in sub["tokens"] if t < tk.eot]
                to_align.append((len(per_stream) - 1, row, text_tokens))
        if o.word_timestamps and to_align:
            self._align_batch(engine, ckv, batch, per_stream, to_align, o)
        t = self._phase("align", t)
        for entry in per_stream:
            if entry is None:
                continue
            s, subs, new_seek, single_ending, avg_logprob, compression_ratio, nsp, temp = entry
            tk = s.tokenizer
            if o.word_timestamps and not single_ending:
                last_word_end = None
                for sub in reversed(subs):
                    if sub.get("words"):
                        last_word_end = sub["words"][-1]["end"]
                        break
                if last_word_end is not None and last_word_end > s.time_offset:
                    new_seek = round(last_word_end * FRAMES_PER_SECOND)
            s.seek = new_seek
            for sub in subs:
                tokens = sub["tokens"]
                text = tk.decode(tokens)
                if sub["start"] == sub["end"] or not text.strip():
                    continue
                s.all_tokens.extend(tokens)
                s.idx += 1
                words = [Word(**w) for w in sub["words"]] if o.word_timestamps else None
                s.segments.append(Segment(s.idx, sub["seek"], sub["start"], sub["end"], text, tokens, avg_logprob,
                                          compression_ratio, nsp, words, temp))
            # (faster-whisper: text produced at a high temperature is not fed back as a prompt)
            if not o.condition_on_previous_text or temp > o.prompt_reset_on_temperature:
                s.prompt_reset_since = len(s.all_tokens)

    def _align_batch(self, engine, ckv, batch, per_stream, to_align, o: TranscriptionOptions) -> None:
        rows = [row for _, row, tt in to_align if tt]
        if rows:
            tk0 = batch[0].tokenizer
            seqs, nframes = [], []
            for _, row, tt in to_align:
                if not tt:
                    continue
                tk = batch[row].tokenizer
                seqs.append(list(tk.sot_sequence) + [tk.no_timestamps] + tt + [tk.eot])
                nframes.append(batch[row].segment_size)
            sub_ckv = ckv if rows == list(range(ckv.shape[0])) else ckv[rows].contiguous()
            aligned = engine.align(sub_ckv, seqs, nframes, len(tk0.sot_sequence), self.alignment_heads)
        k = 0
        for pi, row, tt in to_align:
            s, subs = per_stream[pi][0], per_stream[pi][1]
            if tt:
                ti, fi, tp = aligned[k]
                k += 1
                alignment = words_from_alignment(s.tokenizer, tt, ti, fi, tp)
            else:
                alignment = []
            s.last_speech_timestamp = attach_word_timestamps(s.tokenizer, subs, alignment, o.prepend_punctuations,
                                                             o.append_punctuations, s.last_speech_timestamp)
