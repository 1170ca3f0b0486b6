"""Stand-in for whisper_streaming's silero_vad_iterator.py plus an energy-based
fake of the Silero model (no torch.hub offline)."""
import numpy as np
import torch


class VADIterator:
    def __init__(self, model, threshold=0.5, sampling_rate=16000, min_silence_duration_ms=500, speech_pad_ms=100):
        self.model = model
        self.threshold = threshold
        self.sampling_rate = sampling_rate
        self.min_silence_samples = sampling_rate * min_silence_duration_ms / 1000
        self.speech_pad_samples = sampling_rate * speech_pad_ms / 1000
        self.reset_states()

    def reset_states(self):
        if hasattr(self.model, "reset_states"):
            self.model.reset_states()
        self.triggered = False
        self.temp_end = 0
        self.current_sample = 0

    def __call__(self, x, return_seconds=False):
        if not torch.is_tensor(x):
            x = torch.tensor(np.asarray(x), dtype=torch.float32)
        window = len(x[0]) if x.dim() == 2 else len(x)
        self.current_sample += window
        p = self.model(x, self.sampling_rate).item()
        if p >= self.threshold and self.temp_end:
            self.temp_end = 0
        if p >= self.threshold and not self.triggered:
            self.triggered = True
            start = self.current_sample - self.speech_pad_samples - window
            return {"start": int(start)}
        if p < self.threshold - 0.15 and self.triggered:
            if not self.temp_end:
                self.temp_end = self.current_sample
            if self.current_sample - self.temp_end < self.min_silence_samples:
                return None
            end = self.temp_end + self.speech_pad_samples - window
            self.temp_end = 0
            self.triggered = False
            return {"end": int(end)}
        return None


class FixedVADIterator(VADIterator):
    """Accepts chunks of any length by re-chunking to 512 samples."""

    def reset_states(self):
        super().reset_states()
        self.buffer = np.array([], dtype=np.float32)

    def __call__(self, x, return_seconds=False):
        self.buffer = np.append(self.buffer, x)
        ret = None
        while len(self.buffer) >= 512:
            r = super().__call__(self.buffer[:512], return_seconds=return_seconds)
            self.buffer = self.buffer[512:]
            if ret is None:
                ret = r
            elif r is not None:
                if "end" in r:
                    ret["end"] = r["end"]
                if "start" in r and "end" in ret:
                    del ret["end"]
        return ret if ret != {} else None


class FakeSileroModel:
    """Speech probability from the RMS of the 512-sample window (monotone, stateless)."""

    def __init__(self, rms_at_half=0.02):
        self.k = rms_at_half

    def reset_states(self):
        pass

    def __call__(self, x, sr):
        x = x.float()
        rms = float(torch.sqrt(torch.mean(x * x)))
        return torch.tensor(rms / (rms + self.k))
