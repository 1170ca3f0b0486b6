"""``B200AudioDeduplicator``: the reference's ``AudioDeduplicator`` with the feature extraction on
the GPU.

Same constructor, ``should_skip(audio, sample_rate, current_time)``, ``get_stats``,
``reset_stats``, ``reset`` and return values as /root/reference/audio_deduplicator.py:20-316, so it
can be constructed where the reference builds its deduplicator (一键实时识别麦克风.py:1391) and
called unchanged at :1488 / :1757.  ``_extract_features`` (:60-181) runs in
``b200w_dedup_features`` (whole-chunk DFT in shared memory); the <=100-entry history, the cosine
similarity and the 0.95 / 0.98 thresholds (:183-296) stay on the host, as in the reference.
``should_skip_batch`` checks many streams' chunks with one kernel launch.
"""
from __future__ import annotations

import time
from collections import deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import frontend


class B200AudioDeduplicator:
    def __init__(self, similarity_threshold: float = 0.95, time_window: float = 3.0,
                 min_audio_length: float = 0.1, enable: bool = True, device=None):
        self.similarity_threshold = similarity_threshold
        self.time_window = time_window
        self.min_audio_length = min_audio_length
        self.enabled = enable
        self.device = frontend.require_cuda(device)
        self.feature_history = deque(maxlen=100)
        self.reset_stats()

    # ---------------------------------------------------------------- features (GPU)
    def _extract_features(self, audio: np.ndarray, sample_rate: int = 16000) -> np.ndarray:
        if len(audio) == 0:
            return np.zeros(5)
        return self.extract_features_batch([audio], sample_rate)[0]

    def extract_features_batch(self, chunks: Sequence[np.ndarray], sample_rate: int = 16000) -> np.ndarray:
        audio_dev, offsets = frontend.upload_audio(chunks, self.device)
        feats = frontend.dedup_features(audio_dev, offsets, sample_rate).cpu().numpy()
        if np.isnan(feats).any():
            raise RuntimeError("dedup features: a chunk length is not supported by the GPU DFT")
        return feats

    # ---------------------------------------------------------------- host logic (as the reference)
    @staticmethod
    def _cosine_similarity(vec1: np.ndarray, vec2: np.ndarray) -> float:
        n1, n2 = np.linalg.norm(vec1), np.linalg.norm(vec2)
        if n1 < 1e-10 or n2 < 1e-10:
            return 0.0
        return float((np.dot(vec1, vec2) / (n1 * n2) + 1.0) / 2.0)

    def _clean_history(self, current_time: float):
        cutoff = current_time - self.time_window
        self.feature_history = deque([(f, ts, ln) for f, ts, ln in self.feature_history if ts > cutoff], maxlen=100)

    def should_skip(self, audio: np.ndarray, sample_rate: int = 16000,
                    current_time: Optional[float] = None) -> Tuple[bool, Optional[str], Optional[Dict]]:
        if not self.enabled:
            return False, None, None
        if current_time is None:
            current_time = time.time()
        self.stats["total_checked"] += 1
        audio_length = len(audio) / sample_rate
        if audio_length < self.min_audio_length:
            self.stats["passed"] += 1
            return False, None, None
        self._clean_history(current_time)
        try:
            feats = self._extract_features(audio, sample_rate)
        except Exception as e:
            # conservative like the reference (never drop audio on a failure), but never silently: chunks longer
            # than the shared-memory DFT takes (55 296 samples = 3.46 s; the reference's callers pass 1-2 s buffers,
            # 一键实时识别麦克风.py:1750-1760) are counted and reported once
            self.stats["passed"] += 1
            self.stats["feature_failures"] = self.stats.get("feature_failures", 0) + 1
            if not getattr(self, "_warned_failure", False):
                self._warned_failure = True
                import logging
                logging.getLogger("b200w").warning(
                    "audio dedup: feature extraction failed for a %.2f s chunk (%s); such chunks pass unchecked",
                    audio_length, e)
            return False, None, None
        return self._decide(feats, current_time, audio_length)

    def _decide(self, current_features, current_time, audio_length):
        max_similarity, most_similar_time = 0.0, None
        for hist_features, hist_time, _ in self.feature_history:
            sim = self._cosine_similarity(current_features, hist_features)
            if sim > max_similarity:
                max_similarity, most_similar_time = sim, hist_time
        if max_similarity >= self.similarity_threshold:
            reason = "duplicate" if max_similarity >= 0.98 else "similar"
            self.stats["skipped_duplicate" if reason == "duplicate" else "skipped_similar"] += 1
            self.stats["total_audio_time_skipped"] += audio_length
            details = {"type": reason, "similarity": max_similarity,
                       "time_since_similar": current_time - most_similar_time if most_similar_time else None,
                       "audio_length": audio_length}
            return True, reason, details
        self.feature_history.append((current_features, current_time, audio_length))
        self.stats["passed"] += 1
        return False, None, None

    def get_stats(self) -> Dict:
        return self.stats.copy()

    def reset_stats(self):
        self.stats = {"total_checked": 0, "skipped_duplicate": 0, "skipped_similar": 0, "passed": 0,
                      "total_audio_time_skipped": 0.0}

    def reset(self):
        self.feature_history.clear()
        self.reset_stats()


def should_skip_batch(dedups: Sequence[B200AudioDeduplicator], chunks: Sequence[np.ndarray],
                      sample_rate: int = 16000, current_time: Optional[float] = None) -> List[Tuple]:
    """One feature launch for one chunk of each of many streams (stream i uses dedups[i])."""
    now = time.time() if current_time is None else current_time
    idx = [i for i, (d, c) in enumerate(zip(dedups, chunks))
           if d.enabled and len(c) / sample_rate >= d.min_audio_length]
    feats = dedups[0].extract_features_batch([chunks[i] for i in idx], sample_rate) if idx else []
    out: List[Tuple] = [(False, None, None)] * len(chunks)
    for d, c in zip(dedups, chunks):
        if d.enabled:
            d.stats["total_checked"] += 1
            if len(c) / sample_rate < d.min_audio_length:
                d.stats["passed"] += 1
    for f, i in zip(feats, idx):
        d = dedups[i]
        d._clean_history(now)
        out[i] = d._decide(f, now, len(chunks[i]) / sample_rate)
    return out
