"""The chunk list of the token / word-timing parity sweep: a pure function of the chunk index, shared by the fixture
generator (tools/make_parity_golden.py, CPU oracle) and the GPU test (tests/test_gpu_parity_sweep.py).

Call shape = /root/reference/asr_components.py:279-288: one 16 kHz buffer (0.5 .. 30 s, the sizes the streaming
policy submits), the stream's language, and an init_prompt string (the reference's 300-character prompt,
/root/reference/enhanced_asr_processor.py:295-341, or a shorter / empty one)."""
from typing import List

SECONDS = [0.5, 1.0, 2.5, 5.0, 8.0, 12.0, 15.0, 20.0, 25.0, 29.5, 30.0]
LANGUAGES = ["en", "de", "zh", "ja", "fr", "es", "ru"]
_WORDS = ("the quick brown fox jumps over the lazy dog while seven wizards quietly mix potions and count every "
          "bright star above the old harbour town before dawn breaks again").split()


def _prompt(k: int, n_chars: int) -> str:
    if n_chars == 0:
        return ""
    out, i = "", 7 * k
    while len(out) < n_chars:
        out += _WORDS[i % len(_WORDS)] + " "
        i += 1 + (k % 3)
    return out[:n_chars].strip()


def chunk_spec(k: int) -> dict:
    return dict(seed=5000 + k, seconds=SECONDS[k % len(SECONDS)], language=LANGUAGES[k % len(LANGUAGES)],
                prompt=_prompt(k, [0, 50, 300][k % 3]))


def prompt_tokens(tk, init_prompt: str) -> List[int]:
    """Decoder prompt faster-whisper builds for the first window of a call with initial_prompt=init_prompt
    (SURVEY A.2 steps 4 and 8): the text is tokenised as " " + strip(), the last 223 tokens are sent as previous text."""
    prev = tk.encode(" " + init_prompt.strip())
    return [tk.sot_prev] + prev[-(448 // 2 - 1):] + list(tk.sot_sequence)
