"""Streaming policy for MANY streams at once.

The reference runs ONE stream: per iteration ``EnhancedOnlineASRProcessor.process_iter``
(/root/reference/enhanced_asr_processor.py:343-381, reaching upstream ``OnlineASRProcessor.process_iter``) builds the
prompt (:295-341), calls ``asr.transcribe(audio_buffer, init_prompt=prompt)`` once, feeds the timestamped words to the
LocalAgreement-n hypothesis buffer (``EnhancedHypothesisBuffer.insert/flush`` :54-156 over upstream's
``HypothesisBuffer``), commits the agreed prefix and trims the audio buffer at a segment boundary once it is longer
than ``buffer_trimming_sec`` (dynamic 5..30 s, ``DynamicBufferManager`` :159-236).  With hundreds of streams on a box
that per-stream Python loop -- one backend call each, ``np.append`` re-allocating the whole audio buffer on every
insert -- is the scaling limiter, not the GPU.

``BatchedOnlineASR`` keeps the state of every stream and advances any subset of them with ONE backend call
(``model.transcribe_batch``): same prompts, same commits, same trimming decisions as one reference processor per
stream (parity: tests/test_streaming_policy.py against recorded outputs of the UNCHANGED reference classes,
tests/golden/streaming_policy_golden.json), with
  * audio in per-stream growable buffers with a moving start (amortised O(1) per inserted sample; the reference's
    ``np.append`` copies the whole buffer on every insert),
  * one ``transcribe_batch`` per step for all streams that have new audio,
  * hypothesis buffers as plain lists handled in one thread (no GIL contention between stream threads).
Error convention of the reference (:369-381): a failing backend call is logged, the affected streams are re-initialised
at their current offset and return ``(None, None, "")``.
"""
from __future__ import annotations

import sys
import traceback
from typing import Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

SAMPLING_RATE = 16000
Word = Tuple[float, float, str]


class AudioBuffer:
    """Growable float32 buffer with a moving start: append and drop-front in amortised O(1) per sample."""

    def __init__(self, capacity: int = SAMPLING_RATE * 32):
        self._buf = np.empty(capacity, dtype=np.float32)
        self._start = 0
        self._end = 0

    def __len__(self) -> int:
        return self._end - self._start

    def append(self, audio) -> None:
        a = np.asarray(audio, dtype=np.float32).reshape(-1)
        n = len(a)
        if self._end + n > len(self._buf):
            live = self._end - self._start
            if live + n > len(self._buf) // 2:  # grow geometrically, else just compact in place
                new = np.empty(max(2 * len(self._buf), 2 * (live + n)), dtype=np.float32)
                new[:live] = self._buf[self._start:self._end]
                self._buf = new
            else:
                self._buf[:live] = self._buf[self._start:self._end].copy()
            self._start, self._end = 0, live
        self._buf[self._end:self._end + n] = a
        self._end += n

    def drop_front(self, n: int) -> None:
        self._start = min(self._end, self._start + max(0, int(n)))

    def clear(self) -> None:
        self._start = self._end = 0

    def view(self) -> np.ndarray:
        """The live samples (a view: valid until the next append)."""
        return self._buf[self._start:self._end]


class _Hypotheses:
    """LocalAgreement-n buffer of one stream: upstream ``HypothesisBuffer`` (n = 2) generalised as the reference's
    ``EnhancedHypothesisBuffer`` does for n > 2."""

    def __init__(self, agreement_n: int, last_commited_time: float):
        self.agreement_n = agreement_n
        self.commited_in_buffer: List[Word] = []
        self.buffer: List[Word] = []
        self.new: List[Word] = []
        self.last_commited_time = last_commited_time
        self.last_commited_word: Optional[str] = None
        self.history: List[List[Word]] = []

    def insert(self, new: Sequence[Word], offset: float) -> None:
        lct = self.last_commited_time
        self.new = [(a + offset, b + offset, t) for a, b, t in new if a + offset > lct - 0.1]
        if self.new and abs(self.new[0][0] - lct) < 1 and self.commited_in_buffer:
            # drop a leading 1..5-gram that repeats the tail of what is already committed
            cn, nn = len(self.commited_in_buffer), len(self.new)
            for i in range(1, min(cn, nn, 5) + 1):
                # (compared as space-joined strings, exactly as upstream does)
                if " ".join(w[2] for w in self.commited_in_buffer[-i:]) == " ".join(w[2] for w in self.new[:i]):
                    del self.new[:i]
                    break
        state = self.buffer + self.new
        if len(self.history) >= self.agreement_n:
            self.history.pop(0)
        self.history.append(list(state))

    def flush(self) -> List[Word]:
        if self.agreement_n == 2:
            commit: List[Word] = []
            k = 0
            while k < len(self.new) and k < len(self.buffer) and self.new[k][2] == self.buffer[k][2]:
                commit.append(self.new[k])
                k += 1
            if commit:
                self.last_commited_word = commit[-1][2]
                self.last_commited_time = commit[-1][1]
            self.buffer = self.new[k:]
            self.new = []
            self.commited_in_buffer.extend(commit)
            return commit
        if len(self.history) < self.agreement_n:
            self.buffer = self.buffer + self.new
            self.new = []
            return []
        recent = self.history[-self.agreement_n:]
        words = [[w[2] for w in h] for h in recent]
        n_common = 0
        for i in range(min(len(w) for w in words)):
            if all(w[i] == words[0][i] for w in words):
                n_common += 1
            else:
                break
        commit = list(recent[-1][:n_common])
        if commit:
            self.last_commited_word = commit[-1][2]
            self.last_commited_time = commit[-1][1]
            for h in recent:
                del h[:n_common]
            self.commited_in_buffer.extend(commit)
        self.buffer = list(recent[-1])
        self.new = []
        return commit

    def pop_commited(self, time: float) -> None:
        k = 0
        while k < len(self.commited_in_buffer) and self.commited_in_buffer[k][1] <= time:
            k += 1
        del self.commited_in_buffer[:k]


class _TrimController:
    """The reference's DynamicBufferManager (enhanced_asr_processor.py:159-236): +-2 s steps between 5 and 30 s."""

    def __init__(self, initial: float, lo: float = 5, hi: float = 30):
        self.current, self.lo, self.hi = initial, lo, hi
        self.delays: List[float] = []
        self.memory: List[float] = []

    def record_delay(self, d: float) -> None:
        self.delays.append(d)
        if len(self.delays) > 10:
            self.delays.pop(0)

    def adjust(self) -> bool:
        if not self.delays:
            return False  # (the reference returns the current value here, which is truthy only by accident: see below)
        avg_delay = sum(self.delays) / len(self.delays)
        avg_mem = sum(self.memory) / len(self.memory) if self.memory else 50
        new = self.current
        if avg_delay > 3.0 or avg_mem > 80.0:
            new = max(self.lo, self.current - 2.0)
        elif avg_delay < 1.5 and avg_mem < 56.0:
            new = min(self.hi, self.current + 2.0)
        if abs(new - self.current) > 0.5:
            self.current = new
            return True
        return False


class _Stream:
    def __init__(self, language, agreement_n, trimming_sec, dynamic):
        self.language = language
        self.agreement_n = agreement_n
        self.trimming_sec = trimming_sec
        self.ctl = _TrimController(trimming_sec) if dynamic else None
        self.audio = AudioBuffer()
        self.init(None)

    def init(self, offset: Optional[float]) -> None:
        self.audio.clear()
        self.offset = 0 if offset is None else offset
        self.hyp = _Hypotheses(self.agreement_n, self.offset)
        self.commited: List[Word] = []


class BatchedOnlineASR:
    """One object for many streams; ``process_iter(ids)`` = the reference's ``process_iter`` for each of them, with a
    single batched backend call.  ``model`` needs ``transcribe_batch(audios, languages=, initial_prompts=, **kw)``
    returning ``[(segments, info)]`` with faster-whisper-shaped segments (B200WhisperModel does)."""

    sep = ""  # words carry their own leading space (faster-whisper style), as CustomFasterWhisperASR.sep

    def __init__(self, model, agreement_n: int = 2, buffer_trimming: Tuple[str, float] = ("segment", 15),
                 enable_dynamic_buffer: bool = True, transcribe_kwargs: Optional[dict] = None, logfile=sys.stderr):
        self.model = model
        self.agreement_n = max(2, int(agreement_n))
        self.trimming_way, self.trimming_sec = buffer_trimming
        self.dynamic = enable_dynamic_buffer
        self.kw = dict(transcribe_kwargs or {})
        self.logfile = logfile
        self.streams: Dict[Hashable, _Stream] = {}
        self.calls = 0            # backend calls issued
        self.audio_seconds_submitted = 0.0

    # ------------------------------------------------------------------ stream management
    def add_stream(self, sid: Hashable, language: Optional[str] = None) -> None:
        self.streams[sid] = _Stream(language, self.agreement_n, self.trimming_sec, self.dynamic)

    def remove_stream(self, sid: Hashable) -> None:
        self.streams.pop(sid, None)

    def init(self, sid: Hashable, offset: Optional[float] = None) -> None:
        self.streams[sid].init(offset)

    def insert_audio_chunk(self, sid: Hashable, audio) -> None:
        self.streams[sid].audio.append(audio)

    def buffer_seconds(self, sid: Hashable) -> float:
        return len(self.streams[sid].audio) / SAMPLING_RATE

    # ------------------------------------------------------------------ the policy
    def prompt(self, sid: Hashable) -> Tuple[str, str]:
        """enhanced_asr_processor.py:295-341: newest committed words, whole words only, up to 300 characters."""
        st = self.streams[sid]
        parts, length = [], 0
        for item in reversed(st.commited):
            wl = len(item[2]) + 1
            if length + wl > 300:
                break
            parts.append(item[2])
            length += wl
        prompt = self.sep.join(reversed(parts))
        in_prompt = set(parts)
        rest, rl = [], 0
        for item in reversed(st.commited):
            if item[2] not in in_prompt:
                wl = len(item[2]) + 1
                if rl + wl > 500:
                    break
                rest.append(item[2])
                rl += wl
        return prompt, self.sep.join(reversed(rest))

    def process_iter(self, sids: Optional[Iterable[Hashable]] = None) -> Dict[Hashable, Tuple[Optional[float], Optional[float], str]]:
        ids = list(self.streams) if sids is None else list(sids)
        if not ids:
            return {}
        for sid in ids:  # dynamic trimming threshold, before the call (enhanced_asr_processor.py:349-352)
            st = self.streams[sid]
            if st.ctl is not None and st.ctl.adjust():
                st.trimming_sec = st.ctl.current
        prompts = [self.prompt(sid)[0] for sid in ids]
        audios = [self.streams[sid].audio.view() for sid in ids]
        try:
            kw = {k: v for k, v in self.kw.items() if k not in ("beam_size", "temperature")}
            results = self.model.transcribe_batch(
                audios, languages=[self.streams[sid].language for sid in ids], initial_prompts=prompts,
                beam_size=self.kw.get("beam_size", 5), temperature=self.kw.get("temperature", 0.0),
                word_timestamps=True, condition_on_previous_text=True, **kw)
            self.calls += 1
            self.audio_seconds_submitted += sum(len(a) for a in audios) / SAMPLING_RATE
        except Exception as e:  # the reference logs, resets the stream at its offset and carries on
            print(f"error in a batched iteration: {e}", file=self.logfile)
            traceback.print_exc(file=self.logfile)
            for sid in ids:
                st = self.streams[sid]
                st.init(st.offset)
            return {sid: (None, None, "") for sid in ids}
        out = {}
        for sid, (segments, _info) in zip(ids, results):
            out[sid] = self._after_call(self.streams[sid], list(segments))
        return out

    def _after_call(self, st: _Stream, segments) -> Tuple[Optional[float], Optional[float], str]:
        tsw = [(w.start, w.end, w.word) for s in segments for w in s.words]  # asr_components.py:291-297
        st.hyp.insert(tsw, st.offset)
        o = st.hyp.flush()
        st.commited.extend(o)
        limit = st.trimming_sec if self.trimming_way == "segment" else 30
        if len(st.audio) / SAMPLING_RATE > limit:
            self._chunk_completed_segment(st, [s.end for s in segments])  # asr_components.py:299-301
        res = self._to_flush(o)
        if st.ctl is not None and res[0] is not None:
            st.ctl.record_delay(len(st.audio) / SAMPLING_RATE)  # enhanced_asr_processor.py:360-365
        return res

    def _chunk_completed_segment(self, st: _Stream, ends: List[float]) -> None:
        if not st.commited:
            return
        t = st.commited[-1][1]
        if len(ends) > 1:
            e = ends[-2] + st.offset
            while len(ends) > 2 and e > t:
                ends.pop(-1)
                e = ends[-2] + st.offset
            if e <= t:
                st.hyp.pop_commited(e)
                st.audio.drop_front(int((e - st.offset) * SAMPLING_RATE))
                st.offset = e

    def _to_flush(self, words: Sequence[Word]) -> Tuple[Optional[float], Optional[float], str]:
        if not words:
            return (None, None, "")
        return (words[0][0], words[-1][1], self.sep.join(w[2] for w in words))

    def finish(self, sid: Hashable) -> Tuple[Optional[float], Optional[float], str]:
        st = self.streams[sid]
        res = self._to_flush(st.hyp.buffer)
        st.offset += len(st.audio) / SAMPLING_RATE
        return res
