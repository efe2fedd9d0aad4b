#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): audio-seconds generated per second per B200 at batch 256, plus the
batch-1 per-frame latency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input = BASELINE config 4: 256 utterances
of 60 synthetic token ids each, shared 125-frame voice prefix, Mimi warm-up frame, text prefill, then 275
autoregressive frames (FlowLM step + flow head + Mimi decode) per utterance with EOS disabled
(eos_threshold=+1e30) = 5632 audio-seconds.  Random-init weights of the b6369a24 architecture (seed 0).

  value : device-resident (Philox noise on the device, no host copies), timed with CUDA events on the library's
          own stream, max over ranks.
  e2e   : the same job through the reference-facing C-ABI call ptts_batch_step with HOST buffers: noise
          host->device and latents/EOS logits/1920-sample frames device->host inside the timed region.
  --impl reference : the reference's algorithm on the host cores.  MLX cannot be installed in this image, so
          this arm runs the NumPy restatement (oracle/, pinned to the reference's own Python) with all BLAS
          threads, batch 1 (the only batch size the reference supports), on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

N_SEQ, N_TOK, VOICE_FRAMES = 256, 60, 125
FRAME_SEC = 0.08
PIPELINED = os.environ.get("PTTS_BENCH_SEQUENTIAL", "0") != "1"


def _peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def _init_pg(world, rank, local):
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return dist


def _barrier_max(dist, local, value: float) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def load_model(device: int, kv_pool_tokens: int):
    from pocket_tts_mlx_b200 import TTSModel
    from pocket_tts_mlx_b200.synthetic import default_bundle_dir, write_synthetic_bundle
    yml = write_synthetic_bundle(default_bundle_dir(), seed=0)
    return TTSModel.load_model(str(yml), eos_threshold=1e30, precision="bf16", device_id=device,
                               kv_pool_tokens=kv_pool_tokens), yml


def one_job(model, state, ids, frames, host_io: bool, rng, cap_frames=None):
    """One step of the benchmark: create the batch, warm up Mimi, prefill text, generate `frames` frames.
    cap_frames sizes the KV reservation (warm-up jobs run fewer frames on the same reservation)."""
    from pocket_tts_mlx_b200 import _native
    n = len(ids)
    req = [state["prompt_len"] + len(t) + (cap_frames or frames) for t in ids]
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n, req)
    h2d = d2h = 0
    try:
        batch.seed(1234)
        batch.set_pipelined(PIPELINED)   # Mimi decode of frame t-1 overlaps the FlowLM step of frame t
        batch.warmup_mimi(1)
        batch.prefill_text(ids)
        h2d += sum(len(t) for t in ids) * 4
        if host_io:
            # host buffers = the library's pinned staging arrays (zero-copy C-ABI variant): the host writes this
            # frame's noise, the graph copies it in, runs, and copies latents / EOS logits / 1920-sample frames out
            sink = 0.0
            if PIPELINED:
                # two sets of staging buffers: the host draws the noise of frame t+1 and enqueues it while frame t
                # is still running, then waits for frame t and reads its results (every frame still does its own
                # H2D of noise and D2H of latents / EOS logits / audio inside the timed region)
                batch.set_async_staging(True)
                sets = batch.staging_sets()
                pending = None
                for f in range(frames):
                    z, lat, logit, audio = sets[f & 1]
                    rng.standard_normal(z.shape, dtype=np.float32, out=z)
                    k = batch.step_staged_async()
                    if pending is not None:
                        batch.staged_wait(pending)
                        _, plat, plogit, paudio = sets[pending]
                        sink += float(plogit[0]) + float(paudio[0, 0]) + float(plat[0, 0])     # the host reads the results
                    pending = k
                    h2d += z.nbytes
                    d2h += lat.nbytes + logit.nbytes + audio.nbytes
                batch.staged_wait(pending)
                _, plat, plogit, paudio = sets[pending]
                sink += float(plogit[0]) + float(paudio[0, 0]) + float(plat[0, 0])
                batch.flush()               # audio of the last frame (the first step returned an empty frame)
            else:
                z, lat, logit, audio = batch.staging()
                for f in range(frames):
                    rng.standard_normal(z.shape, dtype=np.float32, out=z)
                    batch.step_staged()
                    sink += float(logit[0]) + float(audio[0, 0]) + float(lat[0, 0])      # the host reads the results
                    h2d += z.nbytes
                    d2h += lat.nbytes + logit.nbytes + audio.nbytes
        else:
            for f in range(frames):
                batch.step_device()
            if PIPELINED:
                batch.flush(want_audio=False)
        model._ctx.sync()
    finally:
        batch.close()
    return h2d, d2h


def latency_bs1(model, state, rng, frames=200, n_tok=42):
    """BASELINE config 2: batch-1 streaming generation, per-frame device latency (CUDA events)."""
    from pocket_tts_mlx_b200 import _native
    ids = [rng.integers(0, 4000, size=n_tok).astype(np.int32)]
    batch = _native.Batch(model._ctx, [state["voice_id"]], [state["prompt_len"] + n_tok + frames + 8])
    batch.seed(7)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(5):
        batch.step_device()
    model._ctx.sync()
    ms = []
    for _ in range(frames - 5):
        model._ctx.timer_begin()
        batch.step_device()
        ms.append(model._ctx.timer_end())
    t0 = time.perf_counter()
    e2e = []
    for _ in range(3):
        t0 = time.perf_counter()
        batch.step(rng.standard_normal((1, 32), dtype=np.float32), want_audio=True)
        e2e.append((time.perf_counter() - t0) * 1e3)
    batch.close()
    # the same utterance on the two-branch frame graph: the step time is then the frame PERIOD (the audio of frame t
    # leaves with step t + 1), i.e. the streaming rate rather than the latency of one frame
    batch = _native.Batch(model._ctx, [state["voice_id"]], [state["prompt_len"] + n_tok + frames + 8])
    batch.seed(7)
    batch.set_pipelined(True)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(5):
        batch.step_device()
    model._ctx.sync()
    per = []
    for _ in range(frames - 5):
        model._ctx.timer_begin()
        batch.step_device()
        per.append(model._ctx.timer_end())
    batch.close()
    return {"p50_ms": float(np.percentile(ms, 50)), "p95_ms": float(np.percentile(ms, 95)),
            "e2e_host_p50_ms": float(np.median(e2e)), "frames": len(ms), "config": "batch 1, 42 tokens, 200 frames",
            "pipelined_period_p50_ms": float(np.percentile(per, 50)),
            "note": "p50_ms: one frame from latent to waveform on the sequential frame graph; pipelined_period: step time "
                    "of the two-branch graph (audio one step later)"}


def roofline_from_profile(model, state, ids, at_frame: int, peaks):
    """Eager per-kernel CUDA-event profile of ONE frame in the middle of the job; the dominant kernel's
    algorithmic bytes / flops come from the launch parameters."""
    from pocket_tts_mlx_b200 import _native
    n = len(ids)
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n, [state["prompt_len"] + len(t) + at_frame + 8 for t in ids])
    batch.seed(1)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(at_frame):
        batch.step_device()
    model._ctx.sync()
    batch.profile_step()                      # warm the eager path
    rows = batch.profile_step()
    batch.close()
    total = sum(r["ms"] for r in rows)
    # One entry per (kernel function, call site): launches of one entry share a problem shape, so "bytes per
    # launch / average launch duration" means something.  The tcgen05 GEMM template serves 30 call sites that
    # range from weight streaming (HBM-bound) to dense contractions (tensor-bound); its summed share is reported
    # separately as gemm_tc_all_share.
    per = {}
    for r in rows:
        a = per.setdefault(r["kernel"], {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
        for f in ("ms", "launches", "flops", "bytes"):
            a[f] += r[f]
    gemm_all = sum(v["ms"] for k, v in per.items() if k.startswith("gemm_tc"))
    top = max(per.items(), key=lambda kv: kv[1]["ms"])
    name, a = top
    sec = a["ms"] / 1e3
    gbs = a["bytes"] / sec / 1e9 if sec > 0 else 0.0
    tfs = a["flops"] / sec / 1e12 if sec > 0 else 0.0
    t_hbm = a["bytes"] / (peaks["hbm_gbs"] * 1e9)
    t_tc = a["flops"] / (peaks["tf_sustained"] * 1e12)
    if t_hbm >= t_tc:
        roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]}
    else:
        roof = {"bound": "tensor", "achieved": tfs, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": tfs / peaks["tf_sustained"]}
    traffic = None
    tp = REPO / "profiles" / "r02_ncu_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(name, {}).get("dram_bytes_per_launch")
    roof.update({"kernel": name, "share_of_frame": a["ms"] / total if total else None,
                 "launches_per_frame": a["launches"], "avg_launch_ms": a["ms"] / max(1, a["launches"]),
                 "algorithmic_bytes_per_launch": a["bytes"] / max(1, a["launches"]),
                 "algorithmic_flops_per_launch": a["flops"] / max(1, a["launches"]),
                 "traffic": traffic, "peak_source": peaks["source"], "frame_ms_eager": total,
                 "gemm_tc_all_share": gemm_all / total if total else None,
                 "note": "eager frame, every launch bracketed by CUDA events on the library stream; shared "
                         "voice-prefix pages are re-read by all sequences and mostly served from L2"})
    breakdown = sorted(({"kernel": k, "ms": v["ms"], "share": v["ms"] / total} for k, v in per.items()),
                       key=lambda r: -r["ms"])[:10]
    return roof, breakdown, rows


def tensor_pipe_evidence(model, peaks):
    """Kernel-level tcgen05 GEMM throughput on the dense contractions of the path (text prefill of the whole
    batch, M = 256 x 60 rows; Mimi ffn, M = 4096 rows), L2 flushed between launches."""
    out = []
    for name, (nb, t, taps, c, n, epi) in {
        "prefill.qkv": (1, N_SEQ * N_TOK, 1, 1024, 3072, 0), "prefill.ff1": (1, N_SEQ * N_TOK, 1, 1024, 4096, 1),
        "prefill.ff2": (1, N_SEQ * N_TOK, 1, 4096, 1024, 0), "mimi.ff2": (N_SEQ, 16, 1, 2048, 512, 0),
        "sn.conv0": (N_SEQ, 16, 7, 512, 512, 1),
    }.items():
        us, cfg = model._ctx.gemm_bench(nb, t, taps, c, n, epi, reps=5)
        tf = 2.0 * nb * t * n * taps * c / us / 1e6
        out.append({"gemm": name, "us": us, "tflops": tf, "frac_of_sustained_peak": tf / peaks["tf_sustained"],
                    "config": {"bn": cfg[0], "stages": cfg[1], "splits": cfg[2], "persistent": cfg[3]}})
    return out


def section_times(model, state, ids, at_frame):
    """In-graph time of the frame's sections (each replayed as its own CUDA graph), microseconds."""
    from pocket_tts_mlx_b200 import _native
    n = len(ids)
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n, [state["prompt_len"] + len(t) + at_frame + 64 for t in ids])
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(at_frame):
        batch.step_device()
    model._ctx.sync()
    sec = batch.profile_sections()
    batch.close()
    return {k: round(v * 1e3, 1) for k, v in sec.items()}


def parity_check(model, state, ids, frames=3, sample=(0, 255)):
    """Outside the timed region: the job's own batch shape (all sequences, pipelined frame graph) stepped `frames`
    frames with host noise, teacher-forced, and the sampled sequences compared with the oracle (the checker, never the
    thing measured).  -> dict for the bench line."""
    from oracle.ptts_oracle import Oracle
    from pocket_tts_mlx_b200 import _native
    from pocket_tts_mlx_b200.config import load_config
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors
    from pocket_tts_mlx_b200.synthetic import default_bundle_dir, write_synthetic_bundle
    yml = write_synthetic_bundle(default_bundle_dir(), seed=0)
    cfg = load_config(yml)
    n = len(ids)
    sample = [b for b in sample if b < n]
    rng = np.random.Generator(np.random.PCG64(4242))
    noise = rng.standard_normal((1 + frames, n, 32)).astype(np.float32)
    orc = Oracle(read_safetensors(cfg.weights_path), cfg, dtype=np.float32, eos_threshold=1e30)
    voice = read_safetensors(Path(yml).parent / "embeddings" / "alba.safetensors")["audio_prompt"]
    st = orc.new_flow_state()
    orc.prefill_audio(st, voice[0])
    refs = {b: orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames) for b in sample}
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n, [state["prompt_len"] + len(t) + frames + 4 for t in ids])
    worst, audio = 0.0, {b: [] for b in sample}
    try:
        batch.set_pipelined(PIPELINED)
        batch.warmup_mimi(1)
        batch.prefill_text(ids)
        for f in range(frames):
            lat, _, au = batch.step(noise[1 + f])
            forced = lat.copy()
            for b in sample:
                r = refs[b]["latents"][f]
                worst = max(worst, float(np.linalg.norm(lat[b] - r) / np.linalg.norm(r)))
                forced[b] = r
                if not PIPELINED or f >= 1:
                    audio[b].append(au[b].copy())
            batch.set_prev_latent(forced)
        if PIPELINED:
            last = batch.flush()
            for b in sample:
                audio[b].append(last[b].copy())
    finally:
        batch.close()
    snr = []
    for b in sample:
        a, r = np.concatenate(audio[b]).astype(np.float64), refs[b]["audio"].astype(np.float64)
        snr.append(float(10 * np.log10((r ** 2).sum() / max(((a - r) ** 2).sum(), 1e-300))))
    ok = bool(worst < 1e-2 and min(snr) > 30.0)
    return {"ok": ok, "latent_rel_l2_max": worst, "waveform_snr_db_min": min(snr), "frames": frames, "sequences": sample,
            "batch": n, "tolerance": "latents 1e-2 rel-L2 (bf16, teacher-forced), waveform 30 dB", "against": "oracle (fp32 NumPy)"}


def workload3(model, state, peaks, n_seq=256, frames=250, steps=2):
    """BASELINE config 3: Mimi decoder only, n_seq synthetic latent sequences x 250 frames (20 s) -> 24 kHz waveform.
    `value`: latents resident on the device side of the call (the small H2D of the latents is inside, the waveform stays
    on the device); `e2e`: the same call returning the waveform to host memory."""
    from pocket_tts_mlx_b200 import _native
    rng = np.random.Generator(np.random.PCG64(33))
    lat = rng.standard_normal((n_seq, frames, 32)).astype(np.float32)
    audio_sec = n_seq * frames * FRAME_SEC
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n_seq, [state["prompt_len"] + 8] * n_seq)
    try:
        batch.warmup_mimi(1)
        batch.mimi_decode(lat[:, :8], want_audio=False)
        batch.mimi_decode(lat, want_audio=False)                    # captures the graph for this frame count
        model._ctx.sync()
        model._ctx.timer_begin()
        for _ in range(steps):
            batch.mimi_decode(lat, want_audio=False)
        ms = model._ctx.timer_end() / steps
        out = batch.mimi_decode(lat, want_audio=True)               # allocates the result array (first touch of 0.5 GB)
        t0 = time.perf_counter()
        out = batch.mimi_decode(lat, want_audio=True, out=out)      # a service re-uses its buffers
        e2e_s = time.perf_counter() - t0
    finally:
        batch.close()
    # algorithmic work of one Mimi frame for one sequence: transformer 201.6 MFLOP, SEANet 323.7 MFLOP
    flops = (51.6e9 + 82.9e9) / 256 * n_seq * frames
    tf = flops / (ms / 1e3) / 1e12
    return {"workload": f"config 3: Mimi decoder only, {n_seq} latent sequences x {frames} frames -> 24 kHz waveform",
            "value": audio_sec / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms, "ms_per_frame": ms / frames,
            "e2e": {"value": audio_sec / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": int(lat.nbytes),
                    "d2h_bytes_per_step": int(out.nbytes)},
            "tensor": {"achieved_tflops": tf, "peak": peaks["tf_sustained"], "frac": tf / peaks["tf_sustained"],
                       "flops_per_frame": flops / frames}}


def workload5(model, state, world, rank, n_total=4096, n_tok=174, slots=256, max_frames=None):
    """BASELINE config 5: n_total independent long-form utterances (174 tokens -> 750 frames = 60 s, KV up to 1049
    tokens), sharded over the ranks (strong scaling: the set is fixed, each rank decodes its share through the
    continuous-batching scheduler of the public API).  Returns (audio seconds this rank produced, wall seconds)."""
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    ids = list(synthetic_token_ids(11, n_total, n_tok))
    t0 = time.perf_counter()
    mine, waves = model.generate_audio_sharded([state] * n_total, ids, rank=rank, world_size=world, slots=slots,
                                               seed=17, max_frames=max_frames)
    dt = time.perf_counter() - t0
    return sum(len(w) for w in waves) / 24000.0, dt, len(mine)


def cpu_baseline(frames: int, seed: int = 0):
    """The oracle (NumPy restatement of the reference's path), batch 1, on this box's host cores."""
    _use_all_host_threads()
    from oracle.ptts_oracle import Oracle
    from pocket_tts_mlx_b200.config import load_config
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors
    from pocket_tts_mlx_b200.synthetic import default_bundle_dir, synthetic_token_ids, write_synthetic_bundle
    yml = write_synthetic_bundle(default_bundle_dir(), seed=0)
    cfg = load_config(yml)
    orc = Oracle(read_safetensors(cfg.weights_path), cfg, dtype=np.float32, eos_threshold=1e30)
    voice = read_safetensors(Path(yml).parent / "embeddings" / "alba.safetensors")["audio_prompt"]
    st = orc.new_flow_state()
    orc.prefill_audio(st, voice[0])
    ids = synthetic_token_ids(2 + seed, 1, N_TOK)[0]
    rng = np.random.Generator(np.random.PCG64(3 + seed))
    noise = rng.standard_normal((1 + frames, 32)).astype(np.float32)
    t0 = time.perf_counter()
    res = orc.generate(st, ids, noise, frames_after_eos=3, max_frames=frames)
    dt = time.perf_counter() - t0
    assert res["n_frames"] == frames
    return frames * FRAME_SEC / dt, dt


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms (rank 0 only) are meant to use the whole host."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    except Exception:
        pass


def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def config4(n_seq: int, frames: int) -> dict:
    """The workload both arms are quoted on (BASELINE config 4)."""
    return {"workload": f"config 4: {n_seq} x {N_TOK}-token utterances, shared {VOICE_FRAMES}-frame voice prefix, "
                        f"{frames} frames each, random-init b6369a24 weights",
            "per_gpu_batch": n_seq, "frames": frames,
            "l2": "inputs larger than L2 (per-frame KV + activations > 126 MB)"}


def run_reference(args, world, rank):
    _use_all_host_threads()
    if rank != 0:
        return
    frames = args.frames          # one whole utterance of the workload per step (275 frames: about 6 s of CPU work)
    for _ in range(args.warmup):
        cpu_baseline(4)
    t_all, audio = 0.0, 0.0
    for k in range(args.steps):
        _, dt = cpu_baseline(frames, seed=k)
        t_all += dt
        audio += frames * FRAME_SEC
    v = audio / t_all
    cores = _blas_threads()
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(config4(args.batch, frames), storage="fp32 (NumPy port of the reference)",
                       note=f"reference is batch-1 only: each step = 1 utterance of the workload ({frames} frames incl. text "
                            "prefill), the job's utterances would run one after the other at this rate"),
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x (1 utterance, 60 tokens, {frames} frames), NumPy/OpenBLAS fp32; "
                                   "MLX itself is not installable here"},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=275, help="frames per utterance (275 = max_gen_len of 60 tokens)")
    ap.add_argument("--batch", type=int, default=N_SEQ)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-latency", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel frame profile (JSON) here")
    ap.add_argument("--workload", type=int, default=4, choices=[3, 4, 5],
                    help="BASELINE config: 4 = headline (256 x 60-token utterances), 3 = Mimi decoder only, "
                         "5 = 4096 long-form utterances sharded over the GPUs (strong scaling)")
    ap.add_argument("--utterances", type=int, default=4096, help="workload 5: size of the utterance set")
    ap.add_argument("--skip-parity", action="store_true")
    args = ap.parse_args()
    world, rank, local = _dist()
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = _init_pg(world, rank, local)
    peaks = _peaks()
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    n_seq, frames = args.batch, args.frames
    if args.workload in (3, 5):
        return run_other_workload(args, world, rank, local, dist, peaks)
    kv_tokens = n_seq * (VOICE_FRAMES + N_TOK + frames + 64) + 4096
    long_ctx = world == 1 and not args.skip_latency          # also time the attention at config-5 context lengths
    if long_ctx:
        kv_tokens = max(kv_tokens, n_seq * (VOICE_FRAMES + 174 + 740 + 16) + 4096)
    if rank != 0:
        _barrier(dist, local)            # rank 0 writes the synthetic bundle first (same files for all ranks)
    model, _ = load_model(local, kv_tokens)
    if rank == 0:
        _barrier(dist, local)
    state = model.get_state_for_audio_prompt("alba")
    ids = list(synthetic_token_ids(2 + rank, n_seq, N_TOK))
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    audio_sec = n_seq * frames * FRAME_SEC

    # ---- value: device-resident ------------------------------------------------------------------
    for _ in range(args.warmup):
        one_job(model, state, ids, min(frames, 16), False, rng, cap_frames=frames)
    sampler = ClockSampler(local)
    _barrier(dist, local)
    model._ctx.sync()
    model._ctx.launch_count(reset=True)
    sampler.start()
    model._ctx.timer_begin()
    for _ in range(args.steps):
        one_job(model, state, ids, frames, False, rng)
    ms = model._ctx.timer_end()
    clocks = sampler.stop()
    launches = model._ctx.launch_count()
    _barrier(dist, local)
    ms = _barrier_max(dist, local, ms)
    value = world * audio_sec * args.steps / (ms / 1e3)

    # ---- e2e: host buffers through the C-ABI step call ---------------------------------------------
    one_job(model, state, ids, min(frames, 8), True, rng, cap_frames=frames)
    _barrier(dist, local)
    model._ctx.sync()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        a, b = one_job(model, state, ids, frames, True, rng)
        h2d += a
        d2h += b
    model._ctx.sync()
    e2e_s = time.perf_counter() - t0
    _barrier(dist, local)
    e2e_s = _barrier_max(dist, local, e2e_s)
    e2e_value = world * audio_sec * e2e_steps / e2e_s

    if rank == 0:
        at_frame = int(os.environ.get("PTTS_BENCH_AT_FRAME", min(frames // 2, 137)))    # where the eager profile is taken
        roof, breakdown, rows = roofline_from_profile(model, state, ids, at_frame, peaks)
        lat = None if args.skip_latency else latency_bs1(model, state, rng)
        sections = section_times(model, state, ids, at_frame)
        tensor = tensor_pipe_evidence(model, peaks)
        # the two branches of the pipelined frame graph (in-graph section times) against the pipelined frame itself
        br_flow = sections.get("flow_backbone", 0.0) + sections.get("eos_flow_head", 0.0)
        br_mimi = sections.get("mimi_transformer_capped", sections.get("mimi_transformer", 0.0)) + \
            sections.get("seanet_capped", sections.get("seanet", 0.0))
        frame_us = ms / args.steps / frames * 1e3
        branches = {"flow_branch_us": round(br_flow, 1), "mimi_branch_us": round(br_mimi, 1),
                    "pipelined_frame_us": round(frame_us, 1),
                    "overlap_fraction": round(max(0.0, br_flow + br_mimi - frame_us) / max(1e-9, min(br_flow, br_mimi)), 3),
                    "note": "overlap = (flow + mimi - frame) / min(flow, mimi); mimi = the branch on the SM share it gets in "
                            "the pipelined graph (PTTS_MIMI_GRID, default 74 SMs); frame = whole job / frames, incl. prefill"}
        parity = None if args.skip_parity else parity_check(model, state, ids)
        extra3 = extra5 = None
        if world == 1 and not args.skip_latency:
            extra3 = workload3(model, state, peaks, steps=2)
        roof_long = None
        if long_ctx:
            # the same kernel where the contexts are long (BASELINE config 5: 174 tokens, 750 frames, KV up to 1049):
            # per-launch set-up and the short items of config 4 no longer weigh on it
            r5 = roofline_from_profile(model, state, list(synthetic_token_ids(9, n_seq, 174)), 740, peaks)[0]
            roof_long = {k: r5[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms",
                                            "algorithmic_bytes_per_launch", "share_of_frame")}
            roof_long["context"] = (f"{n_seq} sequences x ({VOICE_FRAMES} voice + 174 text + 740 generated) keys, "
                                    "eager frame 740 of a config-5 utterance")
        api = None
        if not args.skip_latency and world == 1:
            # the call a user of the reference's API makes: one TTSModel.generate_audio_batch over the same workload
            # (host RNG, per-frame H2D/D2H, EOS bookkeeping, waveforms returned as NumPy arrays); second of two calls
            api_s, waves = None, None
            for _ in range(3):
                waves = None                    # the previous call's 0.5 GB of waveforms are released before the next call
                t0 = time.perf_counter()
                waves = model.generate_audio_batch([state] * n_seq, ids, seed=5, max_frames=frames)
                api_s = time.perf_counter() - t0
            api = {"value": sum(len(w) for w in waves) / 24000.0 / api_s, "unit": "audio-s/s",
                   "call": "TTSModel.generate_audio_batch (pipelined, asynchronous staged steps), third of three calls"}
        cpu = None
        if not args.skip_cpu_baseline and world == 1:        # reported at N = 1 only (it is the same host either way)
            n_utt, dt = 2, 0.0
            for k in range(n_utt):                      # about 12 s of CPU work: two whole utterances of the workload
                dt += cpu_baseline(275, seed=k)[1]
            v = n_utt * 275 * FRAME_SEC / dt
            cpu = {"value": v, "unit": "audio-s/s", "cores": _blas_threads(), "kind": "port",
                   "sample": f"{n_utt} utterances one after the other (batch 1: the reference's only batch size), 60 tokens "
                             f"each, text prefill + 275 frames, NumPy/OpenBLAS fp32 oracle, {dt:.1f} s"}
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config4(n_seq, frames), storage="bf16 weights + bf16 paged KV, fp32 accumulate"),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d // e2e_steps,
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps},
            "public_api": api, "roofline": roof, "roofline_long_context": roof_long, "frame_breakdown": breakdown, "frame_sections_us": sections,
            "branches": branches, "parity_checked": bool(parity and parity["ok"]), "parity": parity,
            "tensor_pipe": tensor, "cpu_baseline": cpu, "latency_bs1": lat,
            "pipelined": PIPELINED,
            "ms_per_frame": ms / args.steps / frames,
            "config3_mimi_only": extra3,
        }
        if args.profile_out:
            Path(args.profile_out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.profile_out).write_text(json.dumps({"rows": rows, "roofline": roof}, indent=1))
        print(json.dumps(line), flush=True)
    _barrier(dist, local)
    model.close()
    if dist is not None:
        dist.destroy_process_group()


def run_other_workload(args, world, rank, local, dist, peaks):
    """--workload 3 (Mimi decoder only, one GPU) and --workload 5 (long-form utterances, strong scaling over GPUs)."""
    n_tok5 = 174
    if args.workload == 3:
        kv_tokens = args.batch * 64 + VOICE_FRAMES + 4096
    else:
        kv_tokens = 256 * (VOICE_FRAMES + n_tok5 + 750 + 40) + 4096
    if rank != 0:
        _barrier(dist, local)
    model, _ = load_model(local, kv_tokens)
    if rank == 0:
        _barrier(dist, local)
    state = model.get_state_for_audio_prompt("alba")
    sampler = ClockSampler(local)
    if args.workload == 3:
        if rank == 0:
            workload3(model, state, peaks, n_seq=args.batch, frames=64, steps=1)           # warm-up
            sampler.start()
            r = workload3(model, state, peaks, n_seq=args.batch, frames=250, steps=max(1, args.steps))
            clocks = sampler.stop()
            line = {"metric": "audio_seconds_per_second", "value": r["value"], "unit": "audio-s/s", "n_gpus": 1,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": r["workload"], "l2": "activations of one frame (> 1 GB at batch 256) exceed L2"},
                    "clocks": clocks, "e2e": r["e2e"],
                    "roofline": {"bound": "tensor", "achieved": r["tensor"]["achieved_tflops"], "peak": r["tensor"]["peak"],
                                 "unit": "TFLOP/s", "frac": r["tensor"]["frac"], "traffic": None,
                                 "kernel": "whole Mimi frame (all GEMMs + attention)",
                                 "peak_source": peaks["source"]},
                    "ms_per_frame": r["ms_per_frame"]}
            print(json.dumps(line), flush=True)
    else:
        # warm-up: graphs, arenas (a short run of the same shape), then --warmup whole steps (at most 2: a step is the
        # whole utterance set) so that the host allocator has seen the result arrays once, as in a running service
        workload5(model, state, world, rank, n_total=256 * world, n_tok=n_tok5, max_frames=24)
        n_warm, n_steps = min(2, max(0, args.warmup)), max(1, args.steps)
        for _ in range(n_warm):
            workload5(model, state, world, rank, n_total=args.utterances, n_tok=n_tok5)
        _barrier(dist, local)
        model._ctx.launch_count(reset=True)
        if rank == 0:
            sampler.start()
        dt = 0.0
        for _ in range(n_steps):
            audio, dt1, n_mine = workload5(model, state, world, rank, n_total=args.utterances, n_tok=n_tok5)
            dt += _barrier_max(dist, local, dt1)
        dt /= n_steps
        clocks = sampler.stop() if rank == 0 else None
        launches = model._ctx.launch_count()
        total_audio = args.utterances * 750 * FRAME_SEC
        if rank == 0:
            line = {"metric": "audio_seconds_per_second", "value": total_audio / dt, "unit": "audio-s/s", "n_gpus": world,
                    "steps": n_steps, "warmup": n_warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": f"config 5: {args.utterances} independent {n_tok5}-token utterances x 750 frames "
                                           f"(60 s, KV up to 1049 tokens), sharded over {world} GPU(s) by frame budget, 256 "
                                           "slots per GPU through TTSModel.generate_audio_sharded (continuous batching)",
                               "utterances_rank0": n_mine, "l2": "inputs larger than L2"},
                    "clocks": clocks, "gpu_launches": int(launches),
                    "e2e": {"value": total_audio / dt, "unit": "audio-s/s",
                            "h2d_bytes_per_step": int(n_mine * 750 * 32 * 4),
                            "d2h_bytes_per_step": int(n_mine * 750 * (1920 + 33) * 4), "steps": n_steps},
                    "note": "timed on the host around the public API call (host RNG, per-frame H2D/D2H, EOS bookkeeping, "
                            "waveforms returned), max over ranks; value == e2e for this workload"}
            print(json.dumps(line), flush=True)
    _barrier(dist, local)
    model.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
