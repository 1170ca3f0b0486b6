#!/usr/bin/env python
"""Headline benchmark: audio-sec/sec (RTFx) of the Whisper hot path on B200.

Workload (BASELINE.json configs[3], per GPU): Whisper large-v3 bf16 (128 mel bins), beam 5, 32
concurrent synthetic 16 kHz streams per GPU (256 across 8 GPUs, weak scaling); one "step" = every
stream submits one 30 s chunk and gets timestamped words back:
    log-mel -> encoder -> cross K/V -> beam-5 decode (Whisper logit rules) -> word timing (DTW).
Weights are seeded synthetic tensors of the exact large-v3 architecture (no checkpoints offline);
the "peaked" synthetic model emits ~100-token transcripts ending in EOT, like real speech.

  value     RTFx with the audio already resident in HBM when the timed region starts
  e2e       RTFx through the public host API (numpy in, Python Segment/Word objects out):
            pinned H2D of every chunk and D2H of the results inside the timed region
  roofline  the kernel with the largest share of the step: the decode-step cross attention
            (cross_attn_sp_kernel, HBM-bound: every live (stream, head) reads its K and V once per
            layer per step).  Decode steps are replayed as CUDA graphs; the engine records CUDA events
            around this kernel's launch in the middle decoder layer as nodes of that graph
            (b200w_step_probe_*), so it is timed INSIDE the timed region, every decode step;
            achieved = algorithmic bytes / that time, peak = MEASURED_PEAKS.json hbm_gbs.  An extra
            eager pass of the same step with events around every launch of every layer
            (b200w_op_profile_enable) is reported next to it as a cross-check
  roofline_gemm  the encoder's tcgen05 GEMM (128x256 tiles, tensor-bound): algorithmic flops /
            CUDA-event time of every launch INSIDE the timed region, vs the measured bf16 peak
  decode_ops  per-op milliseconds of that eager pass (prefill + all decode steps of one step)
  cpu_baseline  the CPU oracle (restatement of the reference's faster-whisper path, fp32) on the
            host cores, on a bounded sample of the same workload (rank 0, N=1 only)

`--impl reference` times that CPU path as the whole run (rank 0 only under torchrun).
One process per GPU; streams are independent, so there is no collective on the data path (NCCL is
used only for the barrier and the max-over-ranks of the elapsed time).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--streams", type=int, default=32, help="concurrent streams per GPU")
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--seconds", type=float, default=30.0, help="audio seconds per chunk")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the front-end roofline and base-latency legs")
    ap.add_argument("--cpu-sample-streams", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=1, help="concurrent sub-batches per GPU (engines + CUDA streams)")
    ap.add_argument("--config", default="large32", choices=["large32", "small64"],
                    help="large32: BASELINE configs[3] per-GPU share (default, one-shot 30 s chunks); small64: configs[2], "
                         "whisper small, 64 streams per GPU, 60 s of audio per stream in 1 s increments through the "
                         "batched streaming policy")
    ap.add_argument("--stream-seconds", type=float, default=None, help="audio per stream in the streaming legs")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="time budget of the reference arm")
    a = ap.parse_args()
    if a.config == "small64":
        a.model, a.streams = "small", 64
    return a


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_streams(n, seconds, rank):
    from tests.synth import speech_like
    return [speech_like(1234 + rank * 100000 + i, seconds) for i in range(n)]


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------ CPU path
def cpu_oracle_pass(model_name, audios, beam, threads=None, precision="fp32"):
    """One pass of the hot path on the host cores with the oracle; returns seconds elapsed.
    (bench.py's cpu_baseline / reference arm: the only place the oracle is executed here.)
    precision: "fp32", or "int8" = Linear layers as dynamically quantised s8 GEMMs (what the reference's CPU path
    does through CTranslate2 compute_type="int8", /root/reference/一键实时识别麦克风.py:1147-1164)."""
    import torch
    from oracle import align as OA
    from oracle import decoding as OD
    from oracle import frontend as OF
    from oracle.model import Dims, OracleWhisper
    from realtime_whisper_asr_b200 import weights as W
    from realtime_whisper_asr_b200.tokenizer import Tokenizer

    if threads:
        torch.set_num_threads(threads)
    md = W.MODEL_DIMS[model_name]
    cache = cpu_oracle_pass.__dict__.setdefault("cache", {})
    if (model_name, "sd") not in cache:
        cache[(model_name, "sd")] = W.synthetic_state_dict(md, seed=0)
    if (model_name, precision) not in cache:
        dims = Dims(md.n_mels, md.d_model, md.n_heads, md.n_enc_layers, md.n_vocab, md.d_model, md.n_heads,
                    md.n_dec_layers)
        cache[(model_name, precision)] = OracleWhisper(cache[(model_name, "sd")], dims, precision)
    model = cache[(model_name, precision)]
    tk = Tokenizer(md.n_vocab, True, language="en")
    opts = OD.DecodeOptions(beam_size=beam, eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech,
                            no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
                            suppress_ids=tk.suppressed_tokens([-1]), blank_ids=tk.blank_tokens())
    heads = W.alignment_heads_for(md)
    t0 = time.perf_counter()
    with torch.no_grad():
        for a in audios:
            mel = OF.log_mel(a, md.n_mels)
            nfr = mel.shape[-1] - 1
            win = OF.pad_or_trim(mel[:, :3000])
            enc = model.encode(torch.from_numpy(win[None]))
            res = OD.decode_one(model, enc[0], tk.sot_sequence, opts)
            text = [t for t in res.tokens if t < tk.eot]
            if text:
                OA.align(model, enc[0], tk.sot_sequence, tk.no_timestamps, tk.eot, text, min(nfr, 3000), heads)
    return time.perf_counter() - t0


def cpu_config1(threads, precision, calls=4):
    """BASELINE configs[0] shape on the host: whisper tiny, ONE stream fed in 1 s increments with the buffer trimmed at
    15 s (what OnlineASRProcessor submits), greedy.  `calls` consecutive calls from the steady state (15 s buffers) are
    timed; returns audio-s/s counting every re-submitted second, as the product's own RTFx does."""
    from tests.synth import speech_like
    audio = speech_like(4321, 30.0)
    bufs = [audio[(i - 15) * SR: i * SR] for i in range(30 - calls + 1, 31)]
    cpu_oracle_pass("tiny", bufs[:1], 1, threads=threads, precision=precision)  # warm-up: weights, packing
    t = cpu_oracle_pass("tiny", bufs, 1, threads=threads, precision=precision)
    return sum(len(b) for b in bufs) / SR / t


def cpu_baseline_report(args, audios):
    """cpu_baseline object: the configured workload on all cores in fp32 and int8 (the faster one is `value`), plus
    the configs[0] shape at the reference-faithful 4 threads (CTranslate2's default intra-op count: the reference
    passes its thread count as num_workers, /root/reference/asr_components.py:259-261) and at all cores."""
    cores = os.cpu_count() or 1
    sample = audios[: args.cpu_sample_streams]
    out, best = {}, None
    for prec in ("fp32", "int8"):
        t = cpu_oracle_pass(args.model, sample, args.beam, threads=cores, precision=prec)
        v = len(sample) * args.seconds / t
        out[f"{prec}_{cores}t"] = {"audio_s_per_s": v, "seconds_of_cpu_work": t}
        if best is None or v > best[0]:
            best = (v, prec, t)
    cfg1 = {}
    for prec in ("fp32", "int8"):
        for th in sorted({4, cores}):
            cfg1[f"{prec}_{th}t"] = cpu_config1(th, prec)
    return {"value": best[0], "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": f"{len(sample)} of {args.streams} streams, one {args.seconds:g} s chunk, full path (log-mel, "
                      f"encoder, beam-{args.beam} decode, word timing), PyTorch oracle on the host, {best[1]} Linear "
                      f"layers ({best[2]:.1f} s of CPU work)",
            "variants": out,
            "config1_tiny_streaming_audio_s_per_s": cfg1,
            "note": "oracle port, not faster-whisper/CTranslate2 (not installable offline); int8 = torch dynamic "
                    "quantisation (s8 GEMMs on the x86 engine), the stand-in BASELINE.md section 4 names"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    audios = make_streams(args.cpu_sample_streams, args.seconds, 0)
    # which arithmetic is faster on this host (fp32 or int8 Linear layers)?  Probe on whisper tiny, then use it.
    probe = {p: cpu_oracle_pass("tiny", audios[:1], 1, threads=cores, precision=p) for p in ("fp32", "int8")}
    probe = {p: cpu_oracle_pass("tiny", audios[:1], 1, threads=cores, precision=p) for p in ("fp32", "int8")}
    prec = min(probe, key=probe.get)
    t_start = time.perf_counter()
    for _ in range(min(args.warmup, 1)):  # one warm-up pass builds the weights and pages them in
        cpu_oracle_pass(args.model, audios, args.beam, precision=prec)
    t, done = 0.0, 0
    for _ in range(args.steps):  # bounded: a step is one chunk of one stream (~20-40 s for large-v3); stop at the budget
        t += cpu_oracle_pass(args.model, audios, args.beam, precision=prec)
        done += 1
        if time.perf_counter() - t_start > args.cpu_budget_s:
            break
    audio_s = done * len(audios) * args.seconds
    value = audio_s / t
    sample = (f"{len(audios)} of {args.streams} streams per step, one {args.seconds:g} s chunk each, "
              f"full path (log-mel, encoder, beam-{args.beam} decode, word timing), PyTorch oracle on the host with "
              f"{prec} Linear layers; {done} of {args.steps} steps timed inside the {args.cpu_budget_s:g} s budget")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "steps_timed": done, "warmup": args.warmup,
        "ms_per_step": 1000 * t / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8" if prec == "int8" else "f32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port (faster-whisper/CTranslate2 are not installable offline); reported, not a target",
    }
    print(json.dumps(line), flush=True)


METRIC = "audio-sec/sec (RTFx), whole job over all GPUs (per_gpu_value = value / n_gpus)"


def workload_config(args):
    if args.config == "small64":
        return {"workload": f"whisper small bf16, beam {args.beam}, 64 concurrent 16 kHz streams per GPU, each "
                            f"{args.stream_seconds or 60:g} s of audio submitted in 1 s increments through the batched "
                            "streaming policy (growing buffers, trimmed at 15 s), word timestamps on "
                            "(BASELINE.json configs[2])",
                "streams_per_gpu": 64, "beam_size": args.beam,
                "weights": "seeded synthetic (peaked transcript circuit), exact architecture",
                "cache_note": "inputs larger than L2 (0.5 GB of weights, fresh buffers every step)",
                "parallelism": "independent streams sharded over GPUs, no collective"}
    return {"workload": f"whisper {args.model} bf16, beam {args.beam}, {args.streams} concurrent 16 kHz streams "
                        f"per GPU x {args.seconds:g} s chunk per step, word timestamps on "
                        "(BASELINE.json configs[3] per-GPU share)",
            "streams_per_gpu": args.streams, "chunk_seconds": args.seconds, "beam_size": args.beam,
            "weights": "seeded synthetic (peaked transcript circuit), exact architecture",
            "cache_note": "inputs larger than L2 (3.1 GB of weights + 61 MB of audio per step)",
            "parallelism": "independent streams sharded over GPUs, no collective"}


# ------------------------------------------------------------------------------------ side legs
def frontend_leg(dev, peaks):
    """BASELINE.json configs[4]: 1 h of synthetic audio as 120 x 30 s rows -> log-mel (80 and 128
    bins) + dedup features of 3600 x 1 s chunks; HBM roofline with the algorithmic bytes of
    SURVEY.md section 8d (4 B/sample in + 4 B x n_mels per 160 samples out)."""
    import torch
    from realtime_whisper_asr_b200 import frontend as FE
    n_rows, n = 120, 480000
    g = torch.Generator(device=dev).manual_seed(5)
    audio_dev = (torch.rand(n_rows * n, generator=g, device=dev) - 0.5) * 0.2
    offsets = np.arange(n_rows + 1, dtype=np.int64) * n
    chunk_offsets = np.arange(3600 + 1, dtype=np.int64) * 16000
    out = {}
    hbm = (peaks.get("hbm_gbs") or 6650.0)
    for name, fn, alg_bytes in (
            ("logmel80", lambda: FE.log_mel(audio_dev, offsets, 80), n_rows * n * 4 + n_rows * 3001 * 80 * 4),
            ("logmel128", lambda: FE.log_mel(audio_dev, offsets, 128), n_rows * n * 4 + n_rows * 3001 * 128 * 4),
            ("dedup_1s_chunks", lambda: FE.dedup_features(audio_dev, chunk_offsets), n_rows * n * 4 + 3600 * 20)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000.0 / reps
        gbs = alg_bytes / (us * 1e-6) / 1e9
        out[name] = {"us_per_hour_of_audio": us, "algorithmic_bytes": alg_bytes, "achieved_GBs": gbs,
                     "frac_of_hbm_peak": gbs / hbm, "audio_s_per_s": 3600.0 / (us * 1e-6)}
    out["note"] = ("input 230 MB > L2; 3 launches per log-mel call (init, mel, finalize); the finalize pass "
                   "re-reads the features (counted once in algorithmic bytes)")
    return out


def latency_leg(local):
    """BASELINE.json configs[1]: Whisper base, one live-style stream, 1 s increments (buffer trimmed
    at 15 s as the reference's OnlineASRProcessor does), beam 5: latency of one transcribe() call."""
    import torch
    from realtime_whisper_asr_b200.backend import B200WhisperASR
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel
    from tests.synth import speech_like
    model = B200WhisperModel("base", device="cuda", device_index=local, max_batch=1, max_beam=5, synthetic=True)
    asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 5, "temperature": 0.0})
    audio = speech_like(4321, 30.0)
    lat = []
    for i in range(1, 31):
        buf = audio[max(0, i - 15) * 16000: i * 16000]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = asr.transcribe(buf, init_prompt="previous words")
        asr.ts_words(res)
        lat.append((time.perf_counter() - t0) * 1000.0)
    lat = np.array(lat[3:])
    del asr, model
    torch.cuda.empty_cache()
    return {"model": "base", "beam_size": 5, "calls": int(len(lat)), "p50_ms": float(np.percentile(lat, 50)),
            "p90_ms": float(np.percentile(lat, 90)),
            "what": "wall time of one backend transcribe()+ts_words() call on a 1..15 s buffer, host numpy in"}


def plugin_leg(model, audios, args, steps, warmup, barrier, max_over_ranks):
    """e2e through the reference-facing plugin call: one backend object per stream (BatchedASR, the class that sits
    where the reference builds CustomFasterWhisperASR), one host thread per stream as the reference's loop would be,
    every step each thread calls asr.transcribe(audio, init_prompt) + asr.ts_words(res)
    (/root/reference/asr_components.py:267-297) on a HOST numpy buffer; the StreamBatcher gathers the concurrent calls
    into one batch.  Returns (ms per step, batch-size histogram, words returned in the last step)."""
    import torch
    from realtime_whisper_asr_b200.scheduler import BatchedASR, StreamBatcher
    n = len(audios)
    batcher = StreamBatcher(model, max_batch=n, max_wait_ms=50)
    asrs = [BatchedASR("en", batcher, transcribe_kwargs={"beam_size": args.beam, "temperature": 0.0}) for _ in range(n)]
    go, done = threading.Barrier(n + 1), threading.Barrier(n + 1)
    words = [0] * n
    errors = []

    def worker(i):
        try:
            for _ in range(warmup + steps):
                go.wait()
                res = asrs[i].transcribe(audios[i], init_prompt="")
                words[i] = len(asrs[i].ts_words(res))
                done.wait()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            go.abort()
            done.abort()

    ts = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n)]
    [t.start() for t in ts]
    try:
        for _ in range(warmup):
            go.wait()
            done.wait()
        barrier()
        batcher.batch_sizes.clear()
        t0 = time.perf_counter()
        for _ in range(steps):
            go.wait()
            done.wait()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1000.0
    except threading.BrokenBarrierError:
        raise errors[0] if errors else RuntimeError("plugin leg: a stream thread failed")
    [t.join(timeout=10) for t in ts]
    hist = {}
    for b in batcher.batch_sizes:
        hist[str(b)] = hist.get(str(b), 0) + 1
    batcher.close()
    return max_over_ranks(ms) / steps, hist, sum(words)


def streaming_leg(model, args, n_streams, seconds, rank, policy="batched"):
    """Growing-buffer streaming (SURVEY 8d cfg 3/4): every stream receives `seconds` of audio in 1 s increments and
    re-submits its buffer (trimmed at 15 s on segment boundaries) after every increment.
      policy "batched": realtime_whisper_asr_b200.streaming.BatchedOnlineASR -- one thread, one transcribe_batch per
                        increment for all streams;
      policy "threads": one upstream-style OnlineASRProcessor (tests/harness stand-in of whisper_online) per stream on
                        its own host thread, calls gathered by the StreamBatcher -- the reference's loop, replicated.
    RTFx per SUBMITTED audio second (every re-submitted buffer second counts: the product's own definition,
    /root/reference/performance_monitor.py:147-175) and per UNIQUE audio second."""
    import torch
    from tests.synth import speech_like
    auds = [speech_like(777 + rank * 100000 + i, seconds) for i in range(n_streams)]
    n_inc = int(seconds)
    kw = {"beam_size": args.beam, "temperature": 0.0}
    out = {"policy": policy, "streams": n_streams, "seconds_per_stream": seconds, "increments": n_inc}
    inner = {"s": 0.0}
    orig = model.transcribe_batch

    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            inner["s"] += time.perf_counter() - t0

    model.transcribe_batch = timed
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if policy == "batched":
            from realtime_whisper_asr_b200.streaming import BatchedOnlineASR
            pol = BatchedOnlineASR(model, agreement_n=2, buffer_trimming=("segment", 15), transcribe_kwargs=kw)
            for i in range(n_streams):
                pol.add_stream(i, "en")
            committed = 0
            for it in range(n_inc):
                for i in range(n_streams):
                    pol.insert_audio_chunk(i, auds[i][it * SR:(it + 1) * SR])
                for _, (b, e, text) in pol.process_iter().items():
                    committed += len(text)
            submitted, calls, hist = pol.audio_seconds_submitted, pol.calls, {str(n_streams): pol.calls}
        else:
            from realtime_whisper_asr_b200.scheduler import BatchedASR, StreamBatcher
            from tests import harness
            harness.install()
            from whisper_online import OnlineASRProcessor
            batcher = StreamBatcher(model, max_batch=n_streams, max_wait_ms=50)
            sub = [0.0] * n_streams
            com = [0] * n_streams

            def worker(i):
                asr = BatchedASR("en", batcher, transcribe_kwargs=dict(kw))
                online = OnlineASRProcessor(asr, buffer_trimming=("segment", 15))
                for it in range(n_inc):
                    online.insert_audio_chunk(auds[i][it * SR:(it + 1) * SR])
                    sub[i] += len(online.audio_buffer) / SR
                    com[i] += len(online.process_iter()[2])

            ts = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n_streams)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            hist = {}
            for b in batcher.batch_sizes:
                hist[str(b)] = hist.get(str(b), 0) + 1
            submitted, calls, committed = sum(sub), batcher.batches_run, sum(com)
            batcher.close()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    finally:
        model.transcribe_batch = orig
    out.update({"wall_s": wall, "backend_calls": calls, "batch_size_histogram": hist,
                "audio_s_submitted": submitted, "audio_s_unique": n_streams * seconds,
                "rtfx_submitted": submitted / wall, "rtfx_unique": n_streams * seconds / wall,
                "host_policy_share": max(0.0, 1.0 - inner["s"] / wall), "committed_chars": committed})
    return out


# ------------------------------------------------------------------------------------ GPU path
def run_b200(args, rank, world, local):
    import torch
    from realtime_whisper_asr_b200 import _lib
    from realtime_whisper_asr_b200 import frontend as FE
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    model = B200WhisperModel(args.model, device="cuda", device_index=local, synthetic=True, max_batch=args.streams,
                             max_beam=max(args.beam, 1), lanes=args.lanes)
    audios = make_streams(args.streams, args.seconds, rank)
    kw = dict(languages=["en"] * args.streams, beam_size=args.beam, temperature=0.0, word_timestamps=True,
              condition_on_previous_text=True)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if args.config == "small64":
        run_small64(args, rank, world, local, model, barrier, max_over_ranks, sum_over_ranks)
        if use_dist:
            dist.destroy_process_group()
        return

    # ---- device-resident leg: audio uploaded once, outside the timed region
    audio_dev, offsets = FE.upload_audio(audios, dev)
    torch.cuda.synchronize()
    n_words = 0
    for _ in range(args.warmup):
        out = model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    import ctypes as C
    _lib.check(lib.b200w_gemm_profile_begin(256))
    _lib.check(lib.b200w_attn_profile_begin())
    _lib.check(lib.b200w_step_probe_reset(model.engine._h))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
    ev1.record()
    barrier()
    pr_ms, pr_n, pr_bytes = C.c_double(), C.c_int64(), C.c_double()
    _lib.check(lib.b200w_step_probe_read(model.engine._h, C.byref(pr_ms), C.byref(pr_n), C.byref(pr_bytes)))
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = _lib.launch_count() - launches0
    tms, tfl, tn = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(lib.b200w_gemm_profile_end(C.byref(tms), C.byref(tfl), C.byref(tn)))
    ams, afl, an = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(lib.b200w_attn_profile_end(C.byref(ams), C.byref(afl), C.byref(an)))
    # one extra, untimed pass with per-phase synchronisation: where the step goes (stderr only)
    model.profile_phases = True
    model.last_timings = {}
    model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
    model.profile_phases = False
    if rank == 0:
        print("phase seconds (1 step, synchronised): " +
              ", ".join(f"{k}={v:.3f}" for k, v in model.last_timings.items() if k != "_"), file=sys.stderr)
    n_words = sum(len(s.words) for segs, _ in out for s in segs)
    n_tokens = sum(len(s.tokens) for segs, _ in out for s in segs)

    # ---- end-to-end legs.  (1) THE e2e number: the reference-facing plugin call, one backend object and one host
    # thread per stream, host numpy in, timestamped words out (H2D of every chunk and D2H of the results inside the
    # timed region).  (2) the same step through one direct transcribe_batch call, for comparison.
    e2e_ms_step, batch_hist, e2e_words = plugin_leg(model, audios, args, args.steps, max(1, args.warmup // 3), barrier,
                                                    max_over_ranks)
    e2e_ms = e2e_ms_step * args.steps
    for _ in range(max(1, args.warmup // 3)):
        model.transcribe_batch(audios, **kw)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out2 = model.transcribe_batch(audios, **kw)
    ev1.record()
    barrier()
    direct_ms = max_over_ranks(ev0.elapsed_time(ev1))
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    audio_s = world * args.streams * args.seconds * args.steps
    value = audio_s / (dev_ms / 1000.0)
    e2e = audio_s / (e2e_ms / 1000.0)
    h2d = int(sum(len(a) for a in audios) * 4)
    # D2H per step: sampled tokens + log-probs + no-speech (decode), DTW paths + token probs (align)
    d2h = int(args.streams * (448 * 4 + 8 + (448 + 1502) * 8 + 448 * 4))

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    achieved = (tfl.value / 1e12) / (tms.value / 1e3) if tms.value > 0 else None

    def traffic_of(name):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f).get("dram_bytes_per_launch")
        except Exception:
            return None

    roofline_gemm = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<256,4,1>", "achieved": achieved,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                     "traffic": traffic_of("gemm_traffic.json"), "peak_source": peak_src,
                     "launches_timed": int(tn.value), "kernel_ms_per_step": tms.value / args.steps,
                     "share_of_step": tms.value / dev_ms if dev_ms else None,
                     "timing": "CUDA events around every launch inside the timed region"}
    a_ach = (afl.value / 1e12) / (ams.value / 1e3) if ams.value > 0 else None
    roofline_attn = {"bound": "tensor", "kernel": "enc_attn4_kernel", "achieved": a_ach, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (a_ach / peak_tf) if a_ach else None, "traffic": None, "peak_source": peak_src,
                     "launches_timed": int(an.value), "kernel_ms_per_step": ams.value / args.steps,
                     "share_of_step": ams.value / dev_ms if dev_ms else None,
                     "algorithmic_flops_per_launch": (afl.value / an.value) if an.value else None,
                     "timing": "CUDA events around every launch inside the timed region"}
    # ---- decode ops: one extra pass of the same step, decode steps launched eagerly with events per op
    _lib.check(lib.b200w_op_profile_enable(model.engine._h, 1))
    model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
    torch.cuda.synchronize()
    _lib.check(lib.b200w_op_profile_enable(model.engine._h, 0))
    ops = {}
    for op in range(16):
        nm, ms_, n_, wk_ = C.c_char_p(), C.c_double(), C.c_int64(), C.c_double()
        _lib.check(lib.b200w_op_profile_read(model.engine._h, op, C.byref(nm), C.byref(ms_), C.byref(n_), C.byref(wk_)))
        if n_.value:
            ops[nm.value.decode()] = {"ms": ms_.value, "launches": int(n_.value), "work": wk_.value}
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)" if peaks.get("hbm_gbs") else "fallback 6.65 TB/s"
    x = ops.get("cross_attn")
    step_ms = dev_ms / args.steps
    n_dec_layers = model.dims.n_dec_layers
    if pr_n.value > 0 and pr_ms.value > 0:
        # timed INSIDE the timed region: events recorded by nodes of the captured step graph around the launch of
        # the middle decoder layer, every decode step; the other layers launch the same kernel on the same shapes
        gbs = pr_bytes.value / 1e9 / (pr_ms.value / 1e3)
        k_ms_step = pr_ms.value * n_dec_layers / args.steps
        roofline = {"bound": "hbm", "kernel": "cross_attn_sp_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbs / hbm_peak, "traffic": traffic_of("cross_tc_traffic.json"), "peak_source": hbm_src,
                    "launches_timed": int(pr_n.value), "kernel_ms_per_step": k_ms_step,
                    "share_of_step": k_ms_step / step_ms,
                    "algorithmic_bytes_per_launch": pr_bytes.value / pr_n.value,
                    "timing": "CUDA events around the launch in the middle decoder layer of every decode step, recorded "
                              f"inside the captured step graph during the timed region; kernel_ms_per_step = that x "
                              f"{n_dec_layers} layers",
                    "eager_pass_check": ({"achieved": x["work"] / 1e9 / (x["ms"] / 1e3), "launches": x["launches"],
                                          "kernel_ms_per_step": x["ms"]} if x and x["ms"] > 0 else None)}
    elif x and x["ms"] > 0:
        gbs = x["work"] / 1e9 / (x["ms"] / 1e3)
        roofline = {"bound": "hbm", "kernel": "cross_attn_sp_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbs / hbm_peak, "traffic": traffic_of("cross_tc_traffic.json"), "peak_source": hbm_src,
                    "launches_timed": x["launches"], "kernel_ms_per_step": x["ms"],
                    "share_of_step": x["ms"] / step_ms,
                    "algorithmic_bytes_per_launch": x["work"] / x["launches"],
                    "timing": "CUDA events around every launch, decode steps run eagerly in one extra pass of the "
                              "same step (inside the timed region they are CUDA-graph replays)"}
    else:
        roofline = dict(roofline_gemm)
    decode_ops = {k: round(v["ms"], 3) for k, v in ops.items()}

    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "per_gpu_value": value / world,
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "path": "one BatchedASR backend object + host thread per stream: asr.transcribe(host numpy, init_prompt) "
                        "+ asr.ts_words(), concurrent calls gathered by the StreamBatcher",
                "batch_size_histogram": batch_hist, "words_per_step": e2e_words},
        "e2e_transcribe_batch": {"value": audio_s / (direct_ms / 1000.0), "unit": "audio-s/s",
                                 "ms_per_step": direct_ms / args.steps,
                                 "path": "one direct model.transcribe_batch(host numpy list) call per step"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "roofline_gemm": roofline_gemm,
        "roofline_attn": roofline_attn,
        "decode_ops_ms_per_step": decode_ops,
        "tokens_per_step": n_tokens, "words_per_step": n_words,
    }
    if world == 1 and not args.no_side_legs:
        try:  # streaming (growing-buffer) workload on the same model: batched policy vs one thread per stream
            sec = args.stream_seconds or 12.0
            line["streaming"] = {p: streaming_leg(model, args, args.streams, sec, rank, p) for p in ("batched", "threads")}
        except Exception as ex:  # side legs must never cost the headline number
            line["streaming_error"] = repr(ex)
        del model
        torch.cuda.empty_cache()
        try:
            line["frontend"] = frontend_leg(dev, peaks)
            line["latency"] = latency_leg(local)
        except Exception as ex:
            line["side_leg_error"] = repr(ex)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_report(args, audios)
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def run_small64(args, rank, world, local, model, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE configs[2]: whisper small, 64 streams per GPU, 60 s of audio per stream in 1 s increments.  One step =
    one increment of every stream through the batched streaming policy (host buffers in, committed text out)."""
    import torch
    from realtime_whisper_asr_b200 import _lib
    sec = args.stream_seconds or 60.0
    streaming_leg(model, args, args.streams, 3.0, rank, "batched")  # warm-up: graphs, allocator
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    res = streaming_leg(model, args, args.streams, sec, rank, "batched")
    barrier()
    launches = _lib.launch_count() - launches0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    wall = max_over_ranks(res["wall_s"])
    submitted = sum_over_ranks(res["audio_s_submitted"])
    unique = sum_over_ranks(res["audio_s_unique"])
    threads = streaming_leg(model, args, args.streams, min(sec, 20.0), rank, "threads") if world == 1 else None
    if rank != 0:
        return
    n_inc = res["increments"]
    line = {
        "metric": METRIC + "; every re-submitted buffer second counts", "value": submitted / wall, "unit": "audio-s/s",
        "n_gpus": world, "steps": n_inc, "warmup": 3, "ms_per_step": 1000.0 * wall / n_inc, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "per_gpu_value": submitted / wall / world,
        "rtfx_unique_audio": unique / wall,
        "e2e": {"value": submitted / wall, "unit": "audio-s/s",
                "h2d_bytes_per_step": int(4 * SR * res["audio_s_submitted"] / n_inc),
                "d2h_bytes_per_step": int(args.streams * (448 * 4 + 8 + (448 + 1502) * 8 + 448 * 4)),
                "path": "BatchedOnlineASR.process_iter (batched streaming policy) over host numpy buffers"},
        "gpu_launches": int(launches), "clocks": sampler.summary(),
        "streaming": {"batched": res, "threads": threads},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local)


if __name__ == "__main__":
    main()
