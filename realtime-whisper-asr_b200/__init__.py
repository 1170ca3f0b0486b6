"""B200-native Whisper hot path (log-mel + dedup features -> encoder -> decoder -> word timing)
behind the reference's ASR backend interface.  Host side: Python; compute: libb200whisper.so
(hand-written sm_100a CUDA, C ABI in include/b200w.h).  No CPU fallback."""

__version__ = "0.1.0"
