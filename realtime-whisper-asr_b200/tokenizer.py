"""Whisper tokenizer wrapper: special-token layout, prompt pieces, word splitting.

Mirrors faster-whisper's ``Tokenizer`` (third-party; built in WhisperModel.transcribe, which the
reference reaches at /root/reference/asr_components.py:279-288).  Two vocabularies:
  * ``tokenizer.json`` (HF ``tokenizers``) when a model directory provides one;
  * a synthetic byte-level vocabulary otherwise (there is no Whisper vocabulary offline): ids
    0..255 are raw bytes, ids 256..50256 are deterministic pseudo word pieces, so that decode /
    word splitting / punctuation merging are exercised on synthetic-weight models.
"""
from __future__ import annotations

import string
from functools import cached_property, lru_cache
from typing import List, Optional, Sequence, Tuple

LANGUAGES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt "
    "la mi ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc "
    "ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()


class _ByteVocab:
    """Synthetic vocabulary: reversible for arbitrary text via the byte tokens."""
    N_TEXT = 50257

    def encode(self, text: str) -> List[int]:
        return list(text.encode("utf-8"))

    @staticmethod
    @lru_cache(maxsize=None)
    def _piece(i: int) -> bytes:
        if i < 256:
            return bytes([i])
        n, s = i - 256, []
        while True:
            s.append("abcdefghijklmnopqrstuvwxyz"[n % 26])
            n //= 26
            if n == 0:
                break
        word = "".join(s)
        if i % 29 == 0:
            word += ","
        elif i % 31 == 0:
            word += "."
        return ((" " if i % 3 != 1 else "") + word).encode()

    def decode(self, ids: Sequence[int]) -> str:
        return b"".join(self._piece(i) for i in ids).decode("utf-8", errors="replace")


class _HFVocab:
    def __init__(self, path: str):
        import tokenizers
        self.tk = tokenizers.Tokenizer.from_file(path)

    def encode(self, text: str) -> List[int]:
        return self.tk.encode(text, add_special_tokens=False).ids

    def decode(self, ids: Sequence[int]) -> str:
        return self.tk.decode(list(ids), skip_special_tokens=False)


class Tokenizer:
    def __init__(self, n_vocab: int = 51865, multilingual: bool = True, task: str = "transcribe",
                 language: Optional[str] = "en", tokenizer_json: Optional[str] = None, vocab=None):
        self.n_vocab = n_vocab
        self.multilingual = multilingual
        # `vocab`: an already loaded vocabulary object to share (tokenizer.json is a multi-MB parse)
        self.vocab = vocab if vocab is not None else (_HFVocab(tokenizer_json) if tokenizer_json else _ByteVocab())
        v3 = 1 if n_vocab >= 51866 else 0
        self.n_languages = 100 if v3 else 99
        self.eot = 50257 if multilingual else 50256
        self.sot = self.eot + 1
        self.lang0 = self.sot + 1
        if multilingual:
            self.translate = 50358 + v3
            self.transcribe = 50359 + v3
            self.sot_lm = 50360 + v3
            self.sot_prev = 50361 + v3
            self.no_speech = 50362 + v3
            self.no_timestamps = 50363 + v3
            self.timestamp_begin = 50364 + v3
        else:  # english-only vocabularies (*.en)
            self.translate, self.transcribe, self.sot_lm, self.sot_prev = 50357, 50358, 50359, 50360
            self.no_speech, self.no_timestamps, self.timestamp_begin = 50361, 50362, 50363
        if task not in ("transcribe", "translate"):
            raise ValueError(f"'{task}' is not a valid task (accepted tasks: transcribe, translate)")
        self.task_name = task
        self.task = self.transcribe if task == "transcribe" else self.translate
        self.language_code = language if multilingual else "en"
        if multilingual:
            if language not in self.languages:
                raise ValueError(f"'{language}' is not a valid language code")
            self.language = self.lang0 + self.languages.index(language)
        else:
            self.language = None

    @property
    def languages(self) -> List[str]:
        return LANGUAGES[: self.n_languages]

    @property
    def language_ids(self) -> List[int]:
        return [self.lang0 + i for i in range(self.n_languages)]

    @cached_property
    def sot_sequence(self) -> List[int]:
        return [self.sot, self.language, self.task] if self.multilingual else [self.sot]

    def encode(self, text: str) -> List[int]:
        return self.vocab.encode(text)

    def decode(self, tokens: Sequence[int]) -> str:
        return self.vocab.decode([t for t in tokens if t < self.eot])

    def decode_with_timestamps(self, tokens: Sequence[int]) -> str:
        outputs: List[list] = [[]]
        for t in tokens:
            if t >= self.timestamp_begin:
                outputs.append(f"<|{(t - self.timestamp_begin) * 0.02:.2f}|>")
                outputs.append([])
            else:
                outputs[-1].append(t)
        return "".join(s if isinstance(s, str) else self.vocab.decode([t for t in s if t < self.eot]) for s in outputs)

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        """faster-whisper Tokenizer.non_speech_tokens: what suppress_tokens=[-1] expands to."""
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        symbols += "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        miscellaneous = set("♩♪♫♬♭♮♯")
        result = {self.encode(" -")[0], self.encode(" '")[0]}
        for symbol in symbols + list(miscellaneous):
            for tokens in (self.encode(symbol), self.encode(" " + symbol)):
                if len(tokens) == 1 or symbol in miscellaneous:
                    result.add(tokens[0])
        return tuple(sorted(result))

    def suppressed_tokens(self, suppress_tokens: Optional[Sequence[int]]) -> List[int]:
        """faster-whisper get_suppressed_tokens."""
        sup = list(suppress_tokens or [])
        if -1 in sup:
            sup = [t for t in sup if t >= 0]
            sup.extend(self.non_speech_tokens)
        sup.extend([self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech])
        return sorted(set(sup))

    def blank_tokens(self) -> List[int]:
        """SuppressBlank: the tokens of " " plus EOT are barred at the first sampled position."""
        return self.encode(" ") + [self.eot]

    # ------------------------------------------------------------------ word splitting
    def split_to_word_tokens(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        if self.language_code in {"zh", "ja", "th", "lo", "my", "yue"}:
            # no spaces between words: every complete unicode code point is its own unit
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        decoded_full = self.decode_with_timestamps(tokens)
        rc = "�"
        words, word_tokens, current, offset = [], [], [], 0
        single = self.__dict__.setdefault("_single_token_text", {})  # most units are one token: decode each id once
        for token in tokens:
            current.append(token)
            if len(current) == 1:
                decoded = single.get(token)
                if decoded is None:
                    decoded = single[token] = self.decode_with_timestamps(current)
            else:
                decoded = self.decode_with_timestamps(current)
            try:
                idx = offset + decoded.index(rc)
                incomplete = decoded_full[idx] != rc if idx < len(decoded_full) else True
            except ValueError:
                incomplete = False
            if not incomplete:
                words.append(decoded)
                word_tokens.append(current)
                current = []
                offset += len(decoded)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        subwords, subword_tokens_list = self.split_tokens_on_unicode(tokens)
        words: List[str] = []
        word_tokens: List[List[int]] = []
        for subword, subword_tokens in zip(subwords, subword_tokens_list):
            special = subword_tokens[0] >= self.eot
            with_space = subword.startswith(" ")
            punctuation = subword.strip() in string.punctuation
            if special or with_space or punctuation or len(words) == 0:
                words.append(subword)
                word_tokens.append(list(subword_tokens))
            else:
                words[-1] = words[-1] + subword
                word_tokens[-1].extend(subword_tokens)
        return words, word_tokens
