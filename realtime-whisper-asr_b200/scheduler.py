"""Host scheduler for many concurrent streams.

The reference is single-stream (one OnlineASRProcessor on the main thread,
/root/reference/一键实时识别麦克风.py:1450-1513) and picks one GPU by free memory (:208-213).  At
scale the unit of parallelism is the stream: every ``transcribe`` call depends only on that
stream's buffer and prompt, so
  * across GPUs streams are SHARDED (sticky, by rank) with no data-path collective: one process
    per GPU, replicated weights; ``torch.distributed`` is used only for barriers and for the
    max-over-ranks of timings (NCCL on GPUs, gloo in the CPU tests);
  * inside a GPU concurrent calls are BATCHED: ``StreamBatcher`` gathers the calls that arrive
    within a short window and runs them as one ``transcribe_batch`` (one encoder / decoder / align
    launch sequence for all of them), each caller getting exactly what a solo call would return (every kernel on
    the path is batch-invariant: tests/test_gpu_fullsize.py::test_batch_invariance_full_size).
``BatchedASR`` is a ``B200WhisperASR`` whose ``transcribe`` goes through the batcher, so several
stream processors (one thread each, as the reference's loop would be per stream) share one engine.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Any, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from .backend import B200WhisperASR


def shard_streams(stream_ids: Sequence[Hashable], world_size: int) -> Dict[int, List[Hashable]]:
    """Sticky partition of streams over ranks: sorted ids dealt round-robin, so that a stream keeps
    its GPU (prompt / KV locality) as long as the set of streams does not change."""
    ids = sorted(stream_ids, key=lambda x: (str(type(x)), x))
    return {r: ids[r::world_size] for r in range(world_size)}


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction used by every multi-GPU number: the slowest rank defines the step."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


@dataclass
class _Request:
    audio: np.ndarray
    language: Optional[str]
    initial_prompt: Optional[str]
    key: Tuple
    kwargs: Dict[str, Any]
    done: threading.Event = field(default_factory=threading.Event)
    result: Any = None
    error: Optional[BaseException] = None


class StreamBatcher:
    """Dynamic batching of concurrent ``model.transcribe`` calls onto one engine."""

    def __init__(self, model, max_batch: Optional[int] = None, max_wait_ms: float = 2.0):
        self.model = model
        self.max_batch = max_batch or getattr(model, "max_batch", 8)
        self.max_wait = max_wait_ms / 1000.0
        self._lock = threading.Condition()
        self._pending: List[_Request] = []
        self._stop = False
        self.batches_run = 0
        self.requests_run = 0
        self.batch_sizes: List[int] = []      # size of every batch run (bench.py's histogram)
        self.busy_seconds = 0.0               # wall time spent inside model.transcribe_batch
        self._thread = threading.Thread(target=self._loop, name="b200w-batcher", daemon=True)
        self._thread.start()

    def transcribe(self, audio, language=None, initial_prompt=None, **kwargs):
        """Same contract as ``B200WhisperModel.transcribe``; blocks until this call's batch ran."""
        key = tuple(sorted((k, repr(v)) for k, v in kwargs.items()))
        req = _Request(np.asarray(audio, dtype=np.float32), language, initial_prompt, key, kwargs)
        with self._lock:
            if self._stop:
                raise RuntimeError("StreamBatcher is closed")
            self._pending.append(req)
            self._lock.notify_all()
        req.done.wait()
        if req.error is not None:
            raise req.error
        segments, info = req.result
        return iter(segments), info

    def close(self):
        with self._lock:
            self._stop = True
            self._lock.notify_all()
        self._thread.join(timeout=5)

    def _take_batch(self) -> List[_Request]:
        with self._lock:
            while not self._pending and not self._stop:
                self._lock.wait()
            if self._stop and not self._pending:
                return []
            deadline = time.monotonic() + self.max_wait
            while len(self._pending) < self.max_batch:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._lock.wait(timeout=left)
            key = self._pending[0].key
            batch = [r for r in self._pending if r.key == key][: self.max_batch]
            taken = set(id(r) for r in batch)
            self._pending = [r for r in self._pending if id(r) not in taken]
            return batch

    def _loop(self):
        while True:
            batch = self._take_batch()
            if not batch:
                return
            t0 = time.perf_counter()
            try:
                out = self.model.transcribe_batch([r.audio for r in batch], languages=[r.language for r in batch],
                                                  initial_prompts=[r.initial_prompt for r in batch], **batch[0].kwargs)
                for r, o in zip(batch, out):
                    r.result = o
            except BaseException as e:  # delivered to every caller of the batch, as a solo call would raise
                for r in batch:
                    r.error = e
            self.busy_seconds += time.perf_counter() - t0
            self.batches_run += 1
            self.requests_run += len(batch)
            self.batch_sizes.append(len(batch))
            for r in batch:
                r.done.set()


class _BatchedModel:
    """``model``-shaped view that routes ``transcribe`` through a StreamBatcher."""

    def __init__(self, batcher: StreamBatcher):
        self._batcher = batcher

    def transcribe(self, audio, language=None, initial_prompt=None, **kw):
        return self._batcher.transcribe(audio, language=language, initial_prompt=initial_prompt, **kw)

    def __getattr__(self, name):
        return getattr(self._batcher.model, name)


class BatchedASR(B200WhisperASR):
    """One per stream; all instances that share ``batcher`` are served by one engine in batches."""

    def __init__(self, lan, batcher: StreamBatcher, **kw):
        super().__init__(lan, model=_BatchedModel(batcher), **kw)


class MultiGpuModel:
    """ONE entry point for many streams on several GPUs of one box, in one process.

    The reference picks a single GPU by free memory (/root/reference/一键实时识别麦克风.py:208-213).  Streams are
    independent, so here they are sharded over the GPUs with no data-path collective: every GPU holds a full model
    (replicated bf16 weights), a stream is pinned to a GPU by its key (sticky: ``shard_streams``), and a call's clips
    are handed to their GPUs' worker threads concurrently (the C calls release the GIL; each engine has its own CUDA
    streams and mutex, kernel attributes and SM counts are per device).  ``transcribe_batch`` has the signature of
    ``B200WhisperModel.transcribe_batch`` -- so ``BatchedOnlineASR`` / ``StreamBatcher`` run on top of it unchanged --
    plus ``keys=`` to name the streams (default: position in the call, which is sticky when callers keep their order).
    One process per GPU under torchrun (bench.py) remains the other way to scale; results are identical either way."""

    def __init__(self, model_factory, devices: Sequence[int]):
        from concurrent.futures import ThreadPoolExecutor
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("MultiGpuModel needs at least one device")
        self.models = [model_factory(d) for d in self.devices]
        self.max_batch = sum(getattr(m, "max_batch", 8) for m in self.models)
        self.dims = self.models[0].dims
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices), thread_name_prefix="b200w-gpu")
        self.calls_per_device = [0] * len(self.devices)

    def __getattr__(self, name):  # tokenizer(), alignment_heads, ... come from the first replica
        return getattr(self.models[0], name)

    def placement(self, keys: Sequence[Hashable]) -> List[int]:
        """Replica index of every key: sorted keys dealt round-robin (the same rule as shard_streams)."""
        shards = shard_streams(list(dict.fromkeys(keys)), len(self.models))
        where = {k: r for r, ks in shards.items() for k in ks}
        return [where[k] for k in keys]

    def transcribe_batch(self, audios, languages=None, initial_prompts=None, keys: Optional[Sequence[Hashable]] = None,
                         **kw):
        n = len(audios)
        keys = list(range(n)) if keys is None else list(keys)
        languages = list(languages) if languages is not None else [None] * n
        prompts = list(initial_prompts) if initial_prompts is not None else None
        place = self.placement(keys)
        parts: Dict[int, List[int]] = {}
        for i, r in enumerate(place):
            parts.setdefault(r, []).append(i)

        def run(r: int, idx: List[int]):
            self.calls_per_device[r] += 1
            extra = {"initial_prompts": [prompts[i] for i in idx]} if prompts is not None else {}
            return self.models[r].transcribe_batch([audios[i] for i in idx], languages=[languages[i] for i in idx],
                                                   **extra, **kw)

        futures = {r: self._pool.submit(run, r, idx) for r, idx in parts.items()}
        out: List[Any] = [None] * n
        for r, idx in parts.items():
            for i, res in zip(idx, futures[r].result()):
                out[i] = res
        return out

    def transcribe(self, audio, language=None, **kw):
        segments, info = self.transcribe_batch([audio], languages=[language], **kw)[0]
        return iter(segments), info

    def close(self):
        self._pool.shutdown(wait=True)
