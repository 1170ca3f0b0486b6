"""CPU, world_size 2 over gloo: the multi-GPU path shards independent streams with no data-path
collective; only timings / counters are reduced (max / sum over ranks)."""
import os
import time

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realtime_whisper_asr_b200 import scheduler as S


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = [f"stream-{i:03d}" for i in range(13)]
    mine = S.shard_streams(streams, world)[rank]
    dist.barrier()
    t0 = time.perf_counter()
    time.sleep(0.05 * (rank + 1))  # rank 1 is the slow one
    elapsed = time.perf_counter() - t0
    slowest = S.max_over_ranks(elapsed)
    total = S.sum_over_ranks(len(mine) * 30.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ret[rank] = dict(mine=mine, elapsed=elapsed, slowest=slowest, total=total, gathered=gathered)
    dist.destroy_process_group()


def test_streams_shard_across_ranks_and_time_is_max_over_ranks():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    all_streams = sorted(r0["mine"] + r1["mine"])
    assert all_streams == [f"stream-{i:03d}" for i in range(13)]           # every stream exactly once
    assert not set(r0["mine"]) & set(r1["mine"]) and abs(len(r0["mine"]) - len(r1["mine"])) <= 1
    assert r0["slowest"] == r1["slowest"] >= max(r0["elapsed"], r1["elapsed"]) - 1e-9  # max over ranks
    assert r0["total"] == r1["total"] == 13 * 30.0                         # whole-job audio seconds
    assert r0["gathered"] == r1["gathered"]


def test_sharding_is_sticky_and_balanced():
    ids = list(range(256))
    a = S.shard_streams(ids, 8)
    assert all(len(v) == 32 for v in a.values()) and sorted(sum(a.values(), [])) == ids
    assert a == S.shard_streams(list(reversed(ids)), 8)  # independent of arrival order
    assert S.shard_streams(ids, 1)[0] == ids


class _FakeModel:
    max_batch = 4

    def __init__(self):
        self.batches = []

    def transcribe_batch(self, audios, languages=None, initial_prompts=None, **kw):
        self.batches.append(len(audios))
        time.sleep(0.01)
        return [([("seg", len(a), p)], ("info", l)) for a, l, p in zip(audios, languages, initial_prompts)]


def test_stream_batcher_batches_concurrent_calls():
    import threading
    import numpy as np
    model = _FakeModel()
    b = S.StreamBatcher(model, max_batch=4, max_wait_ms=50)
    results = {}

    def call(i):
        asr = S.BatchedASR("en", b, transcribe_kwargs={"beam_size": 1})
        results[i] = asr.transcribe(np.zeros(100 + i, np.float32), init_prompt=f"p{i}")

    ts = [threading.Thread(target=call, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    b.close()
    assert sorted(results) == list(range(8))
    for i in range(8):
        assert results[i] == [("seg", 100 + i, f"p{i}")]  # each caller gets its own result
    assert sum(model.batches) == 8 and max(model.batches) > 1 and b.requests_run == 8


def test_multi_gpu_model_shards_and_restores_order():
    import numpy as np
    """MultiGpuModel (one process, one replica per device): sticky placement, every clip reaches exactly one replica,
    results come back in input order, prompts follow their clips.  (Fake replicas: no GPU needed.)"""
    from realtime_whisper_asr_b200.scheduler import MultiGpuModel

    class Fake:
        max_batch = 4
        dims = "dims"

        def __init__(self, dev):
            self.dev, self.seen = dev, []

        def transcribe_batch(self, audios, languages=None, initial_prompts=None, **kw):
            self.seen.append(len(audios))
            ps = initial_prompts or [None] * len(audios)
            return [([("seg", self.dev, float(a[0]), p, l)], ("info", kw.get("beam_size"))) for a, p, l in zip(audios, ps, languages)]

    mg = MultiGpuModel(Fake, [0, 1, 2])
    auds = [np.full(10, float(i), np.float32) for i in range(10)]
    out = mg.transcribe_batch(auds, languages=["en"] * 10, initial_prompts=[f"p{i}" for i in range(10)], beam_size=5)
    assert [o[0][0][2] for o in out] == [float(i) for i in range(10)]          # input order restored
    assert [o[0][0][3] for o in out] == [f"p{i}" for i in range(10)]          # prompts travel with their clips
    assert sorted(sum(m.seen) for m in mg.models) == [3, 3, 4] and mg.max_batch == 12
    place = mg.placement(list(range(10)))
    assert place == mg.placement(list(range(10))) and [o[0][0][1] for o in out] == place   # sticky, and used
    keyed = mg.transcribe_batch(auds[:4], languages=["en"] * 4, keys=["b", "a", "d", "c"])
    assert [o[0][0][1] for o in keyed] == mg.placement(["b", "a", "d", "c"])
    mg.close()
