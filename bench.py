#!/usr/bin/env python
"""Headline benchmark: audio-seconds transcribed per wall-second (BASELINE.json `metric`).

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): whisper
large-v3 (128 mel bins, 32+32 layers), bf16, one hour of synthetic 16 kHz audio cut into
120 independent 30-s windows, greedy decode with the reference's parameter block
(src-tauri/src/whisper.rs:88-124: temperature fallback on, suppress_blank, thresholds).
Random-init weights of the named architecture written as a ggml f16 file (no network here).

One "step" = one pass of the hot path (mel -> encoder -> cross-KV -> decoder loop with
fallback -> segment text) over the rank's 120 windows.  N > 1: one process per GPU, every rank
transcribes its own 120 windows (weak scaling, no collective on the data path; only a barrier
and a max-reduce of the times), `value` = all ranks' audio-seconds / max-over-ranks time.
`--scaling strong` splits ONE hour (120 windows in total, BASELINE config 4 as written) over the ranks
with sharding.shard_range instead (15 windows per GPU at N = 8: a latency-bound decoder batch).

  value  : inputs already resident in HBM (device PCM), timed on the device-side wall
           (barrier + cuda synchronize both sides)
  e2e    : the same through the public host API (WhisperEngine.transcribe_batch over pinned
           host PCM): H2D of the PCM and D2H of the transcripts are inside the timed region.
  roofline: the dominant kernel (decoder cross-attention stream over the cross-KV panels, HBM-bound):
           algorithmic K/V bytes / summed per-launch CUDA-event durations recorded inside the timed
           steps (on the decode lane's own stream).  Two decode lanes share the GPU, so a launch is
           timed while the other lane's projection kernels run beside it; `isolated` is the same
           kernel timed alone in this process.  Encoder GEMM / attention are listed as other_kernels.
  latency : BASELINE.json's second metric — p50 / p99 end-to-end latency of one 5-s utterance on
           large-v3-turbo through WhisperEngine.transcribe (host PCM in, text out), wall clock.
  cpu_baseline: the oracle (CPU restatement of the reference path) on a bounded sample.

`--impl reference` times the reference's own CPU path.  whisper-rs / whisper.cpp are not
vendored in /root/reference and cannot be built offline, so this arm runs the oracle port
(oracle/) on the host cores — the one place besides tests/ and smoke() that executes oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_DIR = os.environ.get("NOBS_BENCH_MODEL_DIR", "/tmp/nobs_whisper_models")
WINDOW_S = 30.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tf_burst": d.get("bf16_tflops", 1590.0),
                "tf_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def gemm_flops_per_window(a) -> float:
    """Algorithmic GEMM FLOPs of one window on the encoder side: both stem
    convolutions, QKV/out/MLP of every encoder layer, cross-KV projections of every decoder layer."""
    d, nm = a.n_audio_state, a.n_mels
    return (2.0 * 3000 * 3 * nm * d + 2.0 * 1500 * 3 * d * d + a.n_audio_layer * 24.0 * 1500 * d * d
            + a.n_text_layer * 4.0 * 1500 * d * d)


def attn_flops_per_window(a) -> float:
    return a.n_audio_layer * 4.0 * 1500 * 1500 * a.n_audio_state


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ensure_model(arch: str, rank: int, world: int, barrier) -> str:
    from nobs_whisper_b200 import ggml_synth
    path = ggml_synth.model_path(MODEL_DIR, arch, seed=0, ftype=1, init="survey")
    if rank == 0 and not os.path.exists(path):
        ggml_synth.ensure_model(MODEL_DIR, arch, seed=0, ftype=1, init="survey")
    barrier()
    return path


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def make_audio(n_windows: int, first: int):
    from nobs_whisper_b200 import synth_audio
    return [synth_audio.synth_clip(first + i, WINDOW_S) for i in range(n_windows)]


def turbo_latency(n_clips: int):
    """BASELINE.json configs[4]: large-v3-turbo, hotkey-style 5-s utterances, one at a time through the
    reference-facing call (WhisperEngine.transcribe: host PCM in, text out); wall clock per utterance."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(MODEL_DIR, "large-v3-turbo", seed=0, ftype=1, init="survey")
    eng = nw.WhisperEngine()
    eng.load_model(path)
    clips = [synth_audio.synth_clip(5000 + i, 5.0) for i in range(n_clips + 3)]
    ms = []
    for i, c in enumerate(clips):
        t0 = time.perf_counter()
        eng.transcribe(c, "en", None, None)
        if i >= 3:   # three warm-up utterances
            ms.append(1e3 * (time.perf_counter() - t0))
    eng.close()
    return {"workload": "whisper large-v3-turbo (32 + 4 layers) bf16, 5-s synthetic utterances, greedy + temperature fallback, one utterance per call",
            "n": len(ms), "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(np.mean(ms)),
            "timing": "host wall clock around WhisperEngine.transcribe (H2D of the PCM and D2H of the text included)"}


def beam5_latency(n_clips: int):
    """BASELINE.json configs[1]: whisper base, beam_size 5 with an initial_prompt custom vocabulary, one 30-s window per call
    (whisper-rs shaped API: FullParams BeamSearch + set_initial_prompt); wall clock per window.  The five beams of a window share one
    audio slot, so the tcgen05 cross-attention streams their K / V panels once per group of up to four rows."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    vocab = ("Claude Code, Anthropic, Supabase, Vercel, shadcn, tRPC, Drizzle, Zod, pnpm, Bun, Deno, Turso, Neon, PlanetScale, Turborepo, Tauri, "
             "SvelteKit, Nuxt, Astro, Vite, Zustand, TanStack, LangChain, LlamaIndex, Ollama, Cursor, Neovim, Vitest, Playwright, Prisma")
    ctx = nw.WhisperContext.new_with_params(ggml_synth.ensure_model(MODEL_DIR, "base", ftype=1, init="fanin"), nw.WhisperContextParameters.default(),
                                            precision="bf16")
    ms, rows = [], 0
    for i in range(n_clips + 2):
        pcm = synth_audio.synth_clip(7000 + i, WINDOW_S)
        p = nw.FullParams.new(nw.SamplingStrategy.BeamSearch(beam_size=5))
        p.set_language("en")
        p.set_initial_prompt(vocab)
        p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
        st = ctx.create_state()
        t0 = time.perf_counter()
        st.full(p, pcm)
        n_seg = len(st.segments())
        dt = 1e3 * (time.perf_counter() - t0)
        if i >= 2:   # two warm-up windows
            ms.append(dt)
            rows += n_seg
        st.close()
    ctx.close()
    return {"workload": "whisper base bf16, beam_size 5 + initial_prompt custom vocabulary (130 prompt tokens), one 30-s synthetic window per call, "
                        "temperature fallback as the reference inherits it",
            "n": len(ms), "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(np.mean(ms)),
            "segments": int(rows), "timing": "host wall clock around WhisperState.full + segment read-back (H2D of the PCM and D2H of the text included)"}


def single_pass_child(args):
    """Child process of the default run (`--single-pass-child`): the SAME clips and model with `temperature_inc = 0`, i.e. one decoding
    pass per window — what the workload costs when the first pass is accepted, as trained weights do on clean speech.  NOT the
    reference's parameters (it inherits the 0.2 ladder; on random-init weights every window then runs all six passes): an auxiliary
    figure next to the headline, printed as one JSON object."""
    import nobs_whisper_b200 as nw
    path = ensure_model(args.model, 0, 1, lambda: None)
    ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision=args.precision)
    audio = make_audio(args.windows, 0)

    def params():
        p = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=1))
        p.set_language("en")
        p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
        p.set_temperature_inc(0.0)
        return p

    secs, windows, segments = [], 0, 0
    for i in range(1 + args.single_pass_steps):   # one warm-up step
        states = [ctx.create_state() for _ in audio]
        t0 = time.perf_counter()
        rc = nw.full_batch(ctx, states, params(), audio)
        n_seg = sum(st.full_n_segments() for st in states)
        dt = time.perf_counter() - t0
        if any(rc):
            raise RuntimeError(f"full_batch rc {rc}")
        if i >= 1:
            secs.append(dt)
            windows += sum(int(st.stats().n_windows) for st in states)
            segments += n_seg
        for st in states:
            st.close()
    ctx.close()
    n = len(secs)
    print(json.dumps({"value": args.windows * WINDOW_S * n / sum(secs), "unit": "audio-s/s", "ms_per_step": 1e3 * sum(secs) / n, "steps": n,
                      "windows_encoded_per_step": windows / n, "segments_per_step": segments / n,
                      "workload": f"whisper {args.model} {args.precision}, the same {args.windows} x 30-s clips, greedy best_of=1 with temperature_inc = 0: ONE decoding pass "
                                  "per window (the cost of the workload when the first pass is accepted); not the reference's parameters — auxiliary figure",
                      "timing": "host wall clock around whisper_b200_full_batch + segment counts, pageable host PCM (H2D inside)"}), flush=True)


def single_pass_figure(args):
    """Runs single_pass_child in its own process (its own context; a failure there cannot take the headline line with it)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--single-pass-child", "--model", args.model, "--windows", str(args.windows),
           "--precision", args.precision, "--single-pass-steps", str(args.single_pass_steps)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.single_pass_timeout_s)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"rc {r.returncode}: " + (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
        return json.loads(lines[-1])
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


def run_reference(args, rank, world, arch_name):
    """--impl reference: the oracle port on the host cores, bounded sample per step."""
    if rank != 0:
        return
    from nobs_whisper_b200 import ggml_synth
    from oracle import oracle
    path = ensure_model(arch_name, 0, 1, lambda: None)
    orc = oracle.Oracle(path)
    # all the host threads this process may use; torchrun exports OMP_NUM_THREADS=1 to its workers, which must not
    # decide the CPU arm's speed (rank 0 is the only rank that works here)
    cores = host_cores()
    oracle.lib().wo_set_threads(cores)
    n_sample = args.cpu_windows
    audio = make_audio(n_sample, 0)
    prm = oracle.reference_params("en")
    # One step = n_sample whole windows (the unit of the metric cannot be cut smaller: every window runs the full 1500-frame
    # encoder and the whole temperature ladder).  A window of large-v3 is ~1 minute of CPU work, so the arm is bounded in TIME:
    # at most one warm-up step (a CPU needs none beyond touching the weights), then as many of the K requested steps as fit
    # --ref-budget-s (at least one); the JSON line reports the steps actually timed.
    times = []
    t_start = time.perf_counter()
    n_warm = min(args.warmup, 1)
    steps_run = 0
    for it in range(n_warm + args.steps):
        t0 = time.perf_counter()
        for a in audio:
            orc.full(prm, a)
        dt = time.perf_counter() - t0
        if it >= n_warm:
            times.append(dt)
            steps_run += 1
        if it + 1 >= n_warm + 1 and (time.perf_counter() - t_start) + dt > args.ref_budget_s:
            break
    args_steps_requested = args.steps
    args.steps = steps_run
    total = sum(times)
    value = args.steps * n_sample * WINDOW_S / total
    sample = (f"{n_sample} window(s) of {WINDOW_S:.0f} s per step (of the 120-window workload), oracle port of the whisper-rs CPU path, fp32, "
              f"{cores} OpenMP threads; the port's GEMM is a plain OpenMP-SIMD loop, not ggml's tuned CPU kernels")
    line = {
        "impl": "reference", "metric": "audio-seconds/sec", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper {arch_name}, 120 x 30-s windows (1 h synthetic 16 kHz audio), greedy + temperature fallback; CPU arm times a bounded sample",
                   "windows_per_step": n_sample, "steps_requested": args_steps_requested, "warmup_run": n_warm, "time_budget_s": args.ref_budget_s},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default=os.environ.get("NOBS_BENCH_MODEL", "large-v3"))
    ap.add_argument("--windows", type=int, default=120)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-windows", type=int, default=1, help="bounded sample for the CPU arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: stop starting new timed steps once this much wall time is spent")
    ap.add_argument("--no-cpu-4threads", action="store_true", help="skip the second cpu_baseline pass at the upstream default of 4 threads")
    ap.add_argument("--latency-clips", type=int, default=200, help="5-s utterances for the large-v3-turbo latency figure (0: skip); BASELINE config 5 asks for >= 200")
    ap.add_argument("--beam-clips", type=int, default=20, help="30-s windows for the base / beam 5 / vocabulary-prompt latency figure of BASELINE config 2 (0: skip)")
    ap.add_argument("--single-pass-steps", type=int, default=2, help="timed steps of the auxiliary one-pass-per-window figure (temperature_inc = 0; 0: skip)")
    ap.add_argument("--single-pass-timeout-s", type=float, default=240.0)
    ap.add_argument("--single-pass-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU transcribes its own --windows windows; strong: --windows windows in total (BASELINE config 4: one hour = 120 windows), "
                         "rank r takes shard_range(windows, world, r)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.single_pass_child:
        single_pass_child(args)
        return
    if args.impl == "reference":
        run_reference(args, rank, world, args.model)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth

    arch = ggml_synth.ARCHS[args.model]
    path = ensure_model(args.model, rank, world, barrier)
    os.environ["NOBS_WHISPER_PRECISION"] = args.precision
    ctxp_device = local_rank

    # the public host API (mirror of the reference's WhisperEngine); the context lands on this rank's GPU
    os.environ["NOBS_WHISPER_DEVICE"] = str(ctxp_device)
    eng = nw.WhisperEngine()
    eng.load_model(path)
    eng.set_profiling(True)

    from nobs_whisper_b200.sharding import shard_range
    if args.scaling == "strong":
        mine = shard_range(args.windows, world, rank)      # contiguous block of the ONE workload
        n_win, first = len(mine), (mine[0] if len(mine) else 0)
        total_windows = args.windows
    else:
        n_win, first = args.windows, rank * args.windows
        total_windows = args.windows * world
    audio = make_audio(n_win, first)
    n_samples = [len(a) for a in audio]
    # pinned host copies (e2e arm) and device copies (HBM-resident arm)
    host = [torch.from_numpy(a).pin_memory() for a in audio]
    dev = [h.cuda(non_blocking=True) for h in host]
    torch.cuda.synchronize()
    h2d_bytes = int(sum(n_samples) * 4)

    def step(tensors):
        texts = eng.transcribe_batch_ptrs([t.data_ptr() for t in tensors], n_samples, language="en")
        return texts, eng.last_stats()

    def timed(tensors, steps, warmup):
        for _ in range(warmup):
            step(tensors)
        barrier()
        torch.cuda.synchronize()
        agg = {"launches": 0, "gemm_ms": 0.0, "gemm_n": 0, "attn_ms": 0.0, "attn_n": 0, "cross_ms": 0.0, "cross_n": 0, "cross_bytes": 0.0, "rows": 0, "samples": 0, "windows": 0, "fallbacks": 0,
               "ms_mel": 0.0, "ms_enc": 0.0, "ms_dec": 0.0, "rounds": 0, "d2h": 0}
        t0 = time.perf_counter()
        eng.event_record(0)  # CUDA events on the stream the kernels are launched on
        for _ in range(steps):
            texts, s = step(tensors)
            agg["launches"] += s.n_kernel_launches; agg["gemm_ms"] += s.gpu_ms_enc_gemm; agg["gemm_n"] += s.n_enc_gemm
            agg["attn_ms"] += s.gpu_ms_enc_attn; agg["attn_n"] += s.n_enc_attn; agg["rows"] += s.n_decode_rows
            agg["cross_ms"] += s.gpu_ms_dec_cross; agg["cross_n"] += s.n_dec_cross; agg["cross_bytes"] += s.dec_cross_bytes
            agg["samples"] += s.n_sample_rows; agg["windows"] += s.n_windows; agg["fallbacks"] += s.n_fallbacks
            agg["ms_mel"] += s.gpu_ms_mel; agg["ms_enc"] += s.gpu_ms_encode; agg["ms_dec"] += s.gpu_ms_decode
            agg["rounds"] += s.n_decode_rounds
            agg["d2h"] += sum(len(t.encode()) for t in texts) + 48 * s.n_sample_rows  # transcripts + per-step sample results
        eng.event_record(1)
        dt_dev_ms = eng.event_elapsed_ms(0, 1)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dt = dt_dev_ms / 1e3 if dt_dev_ms > 0 else wall
        agg["wall_s"] = wall
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        barrier()
        return dt, agg, texts

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dt_dev, agg_dev, _ = timed(dev, args.steps, args.warmup)
    dt_e2e, agg_e2e, texts = timed(host, args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None

    audio_s_per_step = total_windows * WINDOW_S
    value = args.steps * audio_s_per_step / dt_dev
    e2e = args.steps * audio_s_per_step / dt_e2e

    if rank == 0:
        peaks = load_peaks()
        gemm_flops = gemm_flops_per_window(arch) * agg_dev["windows"]
        gemm_s = agg_dev["gemm_ms"] / 1e3
        achieved = gemm_flops / gemm_s / 1e12 if gemm_s > 0 else 0.0
        peak = peaks["tf_sustained"]
        attn_tf = attn_flops_per_window(arch) * agg_dev["windows"] / (agg_dev["attn_ms"] / 1e3) / 1e12 if agg_dev["attn_ms"] > 0 else 0.0
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:   # the reported CPU baseline is an N = 1 figure
            from oracle import oracle
            orc = oracle.Oracle(path)
            prm = oracle.reference_params("en")
            cores = host_cores()

            def cpu_run(threads):
                oracle.lib().wo_set_threads(threads)
                t0 = time.perf_counter()
                for a in audio[: args.cpu_windows]:
                    orc.full(prm, a)
                return time.perf_counter() - t0

            cdt = cpu_run(cores)
            cpu_baseline = {"value": args.cpu_windows * WINDOW_S / cdt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                            "sample": f"first {args.cpu_windows} of the {n_win} windows ({args.cpu_windows * WINDOW_S:.0f} s of audio), oracle port of the "
                                      f"whisper-rs CPU path in fp32 (plain OpenMP-SIMD GEMM loops, not ggml's tuned kernels), {cdt:.1f} s of CPU work"}
            if not args.no_cpu_4threads and cores > 4:
                # The reference never calls set_n_threads (whisper.rs:88-124): it inherits upstream's n_threads = min(4, hw).  A second full
                # pass at 4 threads would add minutes to every run, so the 4-thread / all-core TIME RATIO is measured on the same window
                # with temperature_inc = 0 (mel + encoder + one decoding pass instead of the whole ladder) and applied to the full figure.
                prm1 = oracle.reference_params("en", temperature_inc=0.0)

                def one_pass(threads):
                    oracle.lib().wo_set_threads(threads)
                    t0 = time.perf_counter()
                    orc.full(prm1, audio[0])
                    return time.perf_counter() - t0

                t_all, t_4 = one_pass(cores), one_pass(4)
                oracle.lib().wo_set_threads(cores)
                cpu_baseline["at_4_threads"] = {"value": cpu_baseline["value"] * t_all / t_4, "unit": "audio-s/s", "cores": 4,
                                                "note": "upstream default n_threads = min(4, hw), which the reference inherits; all-core figure scaled by the "
                                                        "4-thread / all-core time ratio measured on the same window with a single decoding pass",
                                                "one_pass_s_all_cores": t_all, "one_pass_s_4_threads": t_4}
            orc.close()
        steps = args.steps
        # Dominant kernel of the step: the decoder's cross-attention stream over the cross-KV panels
        # (HBM-bound).  achieved = algorithmic K/V bytes of the timed launches / summed CUDA-event durations.
        cross_s = agg_dev["cross_ms"] / 1e3
        cross_gbs = agg_dev["cross_bytes"] / cross_s / 1e9 if cross_s > 0 else 0.0
        roofline = {
            "bound": "hbm", "kernel": "dec_cross_attention_tc_kernel (decoder cross-attention: TMA ring -> tcgen05 over the head-major cross-KV panels)",
            "achieved": cross_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": cross_gbs / peaks["hbm_gbs"] if peaks["hbm_gbs"] else None,
            # algorithmic bytes count every cross-KV panel ONCE per launch: rows of the same audio (a pass and its speculative successor)
            # stream the same panels, the second read is served by L2
            # DRAM bytes per launch: ncu --set full of this kernel (profiles/r1_ncu_full_dec_cross_attention_tc.csv, 8 rows per
            # launch) read 63.0-63.7 MB + wrote 1.8-3.7 MB against 61.4 MB algorithmic (the 12th TMA box of a panel covers 36 pad
            # rows; the writes are the attention output and the query tiles); scaled by that ratio to this run's launch size
            "traffic": 1.06 * agg_dev["cross_bytes"] / agg_dev["cross_n"] if agg_dev["cross_n"] else None,
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum at 8 rows per launch (x1.06 of algorithmic), scaled to this launch size",
            "peak_source": peaks["source"] + " hbm_gbs", "launches": agg_dev["cross_n"],
            "algorithmic_bytes_per_launch": agg_dev["cross_bytes"] / agg_dev["cross_n"] if agg_dev["cross_n"] else None,
            "avg_launch_us": 1e3 * agg_dev["cross_ms"] / agg_dev["cross_n"] if agg_dev["cross_n"] else None,
            # every 8th layer's launch is timed (engine.cu kCrossSample): extrapolated share of the decode lanes' stream time
            "launches_timed_of": 8, "share_of_decode_lane_time": 8.0 * agg_dev["cross_ms"] / agg_dev["ms_dec"] if agg_dev["ms_dec"] else None,
            "other_kernels": {
                "gemm_bf16_sm100_kernel (encoder side: conv stem, QKV/out/MLP, cross-KV projection)": {
                    "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                    "launches": agg_dev["gemm_n"], "share_of_step": agg_dev["gemm_ms"] / (1000.0 * dt_dev) if dt_dev else None},
                "enc_attention_sm100_kernel": {
                    "bound": "tensor", "achieved": attn_tf, "peak": peak, "unit": "TFLOP/s", "frac": attn_tf / peak if peak else None,
                    "launches": agg_dev["attn_n"], "share_of_step": agg_dev["attn_ms"] / (1000.0 * dt_dev) if dt_dev else None},
            },
        }
        # the same kernel alone on the GPU (micro-benchmark hook of the library, rows of one lane)
        try:
            import ctypes as C
            from nobs_whisper_b200 import _lib
            us = np.zeros(16, np.float32)
            r_iso = max(1, min(128, int(round(agg_dev["rows"] / max(agg_dev["rounds"], 1) / max(eng.n_lanes(), 1)))))
            if _lib.lib().whisper_b200_debug_time_decode_kernels(r_iso, arch.n_text_state, 50, us.ctypes.data_as(C.POINTER(C.c_float))) == 0 and us[12] > 0:
                iso_bytes = 2.0 * r_iso * 1500 * arch.n_text_state * 2
                iso_us = float(us[12])
                roofline["isolated"] = {"rows": int(r_iso), "avg_launch_us": iso_us, "achieved": float(iso_bytes / (iso_us * 1e-6) / 1e9),
                                        "frac": float(iso_bytes / (iso_us * 1e-6) / 1e9 / peaks["hbm_gbs"])}
        except Exception as exc:  # the figure is auxiliary
            roofline["isolated"] = {"error": str(exc)}
        latency = None
        if args.latency_clips > 0 and world == 1:
            latency = turbo_latency(args.latency_clips)
        latency_beam5 = None
        if args.beam_clips > 0 and world == 1:
            try:
                latency_beam5 = beam5_latency(args.beam_clips)
            except Exception as exc:  # auxiliary figure: never lose the headline line to it
                latency_beam5 = {"error": str(exc)}
        single_pass = None
        if args.single_pass_steps > 0 and world == 1 and not args.no_cpu_baseline:   # part of the default run only
            single_pass = single_pass_figure(args)
        line = {
            "metric": "audio-seconds/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt_dev / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {
                "workload": f"whisper {args.model} ({arch.n_mels} mel bins, {arch.n_audio_layer}+{arch.n_text_layer} layers) {args.precision}, "
                            + (f"{n_win} x 30-s windows per GPU (1 h of synthetic 16 kHz audio each; weak scaling)" if args.scaling == "weak" else
                               f"{total_windows} x 30-s windows in total split over {world} GPU(s) ({n_win} on rank 0; strong scaling of the one-hour workload)")
                            + ", greedy best_of=1 with the reference's temperature fallback; every clip ends in a sub-second remainder window "
                              "(whisper.cpp 1.7.6 delta_min = 100 ms), so a step encodes and decodes ~2 windows per clip — twice the work of the round-1 "
                              "figure, which stopped at 1 s (see DESIGN.md section 6)",
                "windows_per_gpu": n_win, "windows_total": total_windows, "weights": "random-init N(0,0.02) ggml f16 file, seed 0", "l2": "inputs and weights exceed L2 (3.1 GB weights, 30 GB cross-KV)",
                "decoder_rows_per_step": agg_dev["rows"] / steps, "sampled_tokens_per_step": agg_dev["samples"] / steps,
                "decoder_rounds_per_step": agg_dev["rounds"] / steps, "fallbacks_per_step": agg_dev["fallbacks"] / steps,
                "stage_ms_per_step": {"mel": agg_dev["ms_mel"] / steps, "encode": agg_dev["ms_enc"] / steps,
                                      "decode_lane_sum": agg_dev["ms_dec"] / steps, "decode_lanes": eng.n_lanes()},
                "speculative_first_fallback": os.environ.get("NOBS_WHISPER_SPECULATE", "1") != "0",
                "x_realtime": value, "timing": "CUDA events on the library stream around the K steps, max over ranks",
                "host_wall_ms_per_step": 1000.0 * agg_dev["wall_s"] / steps,
            },
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(total_windows * WINDOW_S * 16000 * 4), "d2h_bytes_per_step": int(agg_e2e["d2h"] / steps) * world,
                    "ms_per_step": 1000.0 * dt_e2e / steps},
            "gpu_launches": int(agg_dev["launches"] + agg_e2e["launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "latency": latency,
            "latency_beam5": latency_beam5,
            "single_pass": single_pass,
        }
        print(json.dumps(line, default=float), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
