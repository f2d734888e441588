#!/usr/bin/env python
"""Headline benchmark: Griffin-Lim audio-seconds per second (64 iterations, 24 kHz) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|gl|gl_sharded|frontend]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json `configs`):

* ``gl`` (default at N = 1) -- config 2: a Fisher-test-shaped ragged batch of 256 synthetic log-mel utterances
  (T ~ U{56..400}, seed 0, length-sorted) through inverse-mel, the initial inverse and 64 fused STFT/iSTFT
  Griffin-Lim iterations.  One step = one pass over that batch.  Extra keys time config 1 (one 500-frame utterance
  through ``GriffinLimVocoder.forward``) and config 5 (a 60 s utterance, 64 / 256 iterations).
* ``gl_sharded`` (default at N > 1) -- config 4: ONE global list of 10 000 utterances (same length law, seed 0)
  sharded by utterance over the ranks (longest-processing-time-first on frames x iterations), cut into length
  buckets per rank, synthesised bucket by bucket, and gathered to rank 0 over NCCL at the end (strong scaling; the
  gather is the only collective).  One step = the whole list once.
* ``frontend`` -- config 3: fbank80 + fused global CMVN over 10 000 synthetic utterances of 8-20 s at 16 kHz
  (logmelspec80 at 24 kHz and fbank80 at 8 kHz on a share of it beside it).

Prints ONE JSON line (rank 0).  ``value`` is device-timed with inputs resident in HBM; ``e2e`` goes through the
public API from pinned host buffers (H2D of the inputs incl. the seeded initial phase, D2H of the results inside the
timed region); ``roofline`` is the dominant kernel against the measured HBM peak; ``cpu_baseline`` is the CPU port
timed on this box's host cores on a bounded sample.  ``--impl reference`` times the reference's own CPU formulation
with all host cores instead (the reference is Python/PyTorch CPU code that cannot travel to the GPU box).
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "speech-to-speech-translation_b200"

SR, N_FFT, WIN, HOP, N_MELS, F_MIN, F_MAX, N_ITER = 24000, 2048, 1200, 300, 80, 20.0, 8000.0, 64
N_BINS = N_FFT // 2 + 1
N_UTTS = 256
N_UTTS_SHARDED = 10000
BUCKET_FRAMES = 100000                  # frames per synthesis call of the sharded workload (several buckets per rank:
                                        # the gather of one overlaps the synthesis of the next)
ALGO_BYTES_PER_FRAME_ITER = 6500        # 1200 B wave in + 4100 B magnitude + 1200 B wave out
ALGO_BYTES_PER_FRAME_ONCE = 13820       # inverse-mel + initial inverse
FBANK_BYTES_PER_FRAME = 960             # 160 new samples + 80 features
LOGMEL_BYTES_PER_FRAME = 1520           # 300 new samples + 80 features
WORKLOAD = ("Fisher-test-shaped batch: 256 synthetic log-mel utterances, 56-400 frames length-bucketed, "
            "64 Griffin-Lim iters per GPU")
WORKLOAD_SHARDED = ("Griffin-Lim 64 iters over ONE list of 10k synthetic utterances (56-400 frames) sharded by "
                    "utterance over the GPUs (LPT on frames x iterations), length-bucketed per rank, final gather of "
                    "waveforms to rank 0")
WORKLOAD_FRONTEND = ("source fbank80 + global CMVN extraction over 10k synthetic utterances (8-20 s, 16 kHz) on 1 B200")
GL_NCU_SUMMARY = os.path.join("profiles", "r02_glpass_ncu_summary.txt")
GL_NCU_SUMMARY_OLD = os.path.join("profiles", "r01_glpass_current_ncu_summary.txt")


# ---- synthetic workloads (shared with tests/) -----------------------------------------------------------------
def batch_frames(seed, n_utts=N_UTTS):
    rng = np.random.RandomState(seed)
    return sorted(int(t) for t in rng.randint(56, 401, size=n_utts))


def sharded_frames(seed=0):
    """Config 4: the global utterance list in its (unsorted) corpus order."""
    rng = np.random.RandomState(seed)
    return [int(t) for t in rng.randint(56, 401, size=N_UTTS_SHARDED)]


def synth_logmel_np(T, seed):
    """SURVEY 8(d) speech-like log-mel: smooth random walk in time + spectral tilt, clamped."""
    rng = np.random.RandomState(seed)
    x = 0.1 * np.cumsum(rng.randn(T, N_MELS), axis=0) + np.linspace(0, -4, N_MELS)[None, :] - 2.0
    return np.clip(x, np.log(1e-5), 2.0).astype(np.float32)


def config2_batch(rank=0):
    """(frames, log-mel [sum T, 80], initial phase [sum T, 1025] frame-major) of the config-2 batch of one rank."""
    frames = batch_frames(0)
    logmel = np.concatenate([synth_logmel_np(T, 1234 + 1000 * rank + i) for i, T in enumerate(frames)])
    rng = np.random.RandomState(100 + rank)
    phase = np.angle(np.exp(2j * np.pi * rng.rand(int(sum(frames)), N_BINS))).astype(np.float32)
    return frames, logmel, phase


def sharded_utterance_inputs(i, T):
    """Log-mel [T, 80] and initial phase [T, 1025] (frame-major) of utterance i of the config-4 list: a function of
    the utterance alone, so every sharding of the list synthesises the same thing."""
    x = synth_logmel_np(T, 50000 + i)
    rng = np.random.RandomState(90000 + i)
    u = rng.rand(T, N_BINS).astype(np.float32)
    return x, ((2.0 * u - 1.0) * np.float32(np.pi)).astype(np.float32)


def measured_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_gl_pass launch on the config-2 batch, from the committed
    `ncu --set full` capture (ncu cannot run inside the benchmark); (bytes, source) or (None, None)."""
    for rel in (GL_NCU_SUMMARY, GL_NCU_SUMMARY_OLD):
        try:
            tot = 0.0
            for line in open(os.path.join(ROOT, rel)):
                if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
                    val, unit = line.split("=")[1].split()[:2]
                    tot += float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
            if tot:
                return tot, rel
        except (OSError, KeyError, ValueError, IndexError):
            pass
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1620.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU through NVML (what the nvidia-smi --query-gpu clocks line
    reports).  sample() is called by the benchmark itself once every step of the timed region is enqueued and
    the GPU is still executing them -- a polling thread, a subprocess, or calls between steps were measurably
    delaying kernel launches on this driver."""

    def __init__(self, gpu_index):
        self.samples, self.reasons, self.err, self.sm_max = [], set(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err, self.nv = repr(e), None

    def sample(self):
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for n, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                           ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                           ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                           ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
                if mask & bit:
                    self.reasons.add(n)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample_while(self, event, max_samples=8):
        """Sample while `event` (recorded after the last timed step) has not completed: everything of the timed region
        is enqueued, the GPU is still working through the queue, and no launch can be delayed by a slow NVML call."""
        while not event.query() and len(self.samples) < max_samples:
            self.sample()
            time.sleep(0.005)
        self.under_load = len(self.samples)
        if not self.samples:
            self.sample()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                "samples_under_load": getattr(self, "under_load", len(sm)), "reasons": sorted(self.reasons),
                "source": "NVML, sampled inside the timed region while the GPU drains the enqueued steps"}


# ---- CPU legs (the only places that execute oracle/) ---------------------------------------------------------------
def cpu_port_time(frames_subset, seed, n_iter, basis):
    """Time the numpy oracle (the CPU port of the reference path) on the given utterances, 1 core."""
    from oracle import griffin_lim as ogl
    t0 = time.perf_counter()
    audio = 0.0
    for i, T in enumerate(frames_subset):
        x = synth_logmel_np(T, seed + i)
        np.random.seed(seed + i)
        phase = ogl.random_phase((N_BINS, T))
        y = ogl.vocoder_forward(x, phase, n_iter, basis=basis)
        audio += y.shape[0] / SR
    return audio, time.perf_counter() - t0


def _cpu_worker(args):
    frames, seed, n_iter = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import griffin_lim as ogl
    basis = ogl.pinv_mel_basis(SR, N_FFT, N_MELS, F_MIN, F_MAX)
    return cpu_port_time(frames, seed, n_iter, basis)[0]


def _fft_port_all_cores(frames, cores):
    """The numpy FFT oracle, one process per core, on a length-stratified sample: audio-s/s."""
    import multiprocessing as mp
    n_sample = max(1, min(len(frames), 2 * cores, 32))
    sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, n_sample)]
    chunks = [(sample[i::cores], 1000 + i, N_ITER) for i in range(min(cores, n_sample))]
    with mp.get_context("spawn").Pool(len(chunks)) as pool:
        pool.map(_cpu_worker, [(c[0][:1], c[1], 1) for c in chunks])  # start the workers, import numpy
        t0 = time.perf_counter()
        audio = sum(pool.map(_cpu_worker, chunks))
        dt = time.perf_counter() - t0
    return audio / dt, len(chunks), n_sample, sum(sample)


def synth_audio_np(n, sr, seed):
    """SURVEY 8(d) audio for config 3: white noise x 0.1 + 3 random sinusoids, in [-1, 1]."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / sr
    x = 0.1 * rng.randn(n)
    for _ in range(3):
        x += rng.uniform(0.05, 0.3) * np.sin(2 * np.pi * rng.uniform(80, 0.45 * sr) * t + rng.uniform(0, 6.28))
    return np.clip(x, -1, 1).astype(np.float32)


def frontend_cpu_time(durations_s, sr, threads, min_seconds=10.0):
    """The reference's own fbank80 + global CMVN path on the host cores: torchaudio.compliance.kaldi.fbank
    (audio_utils.py:141-147) then (x - mean) / std in numpy (global_cmvn.py:26-29), one utterance per call.  The
    utterances of the sample are processed again and again until ``min_seconds`` of CPU work have been timed."""
    import torch
    import torchaudio.compliance.kaldi as ta_kaldi
    torch.set_num_threads(threads)
    rng = np.random.RandomState(7)
    mean, std = (rng.randn(80) - 4).astype(np.float32), rng.uniform(0.5, 2, 80).astype(np.float32)
    waves = [torch.from_numpy(synth_audio_np(int(d * sr), sr, 300 + i) * (2 ** 15))[None] for i, d in enumerate(durations_s)]
    audio, t0 = 0.0, time.perf_counter()
    while time.perf_counter() - t0 < min_seconds:
        for w, d in zip(waves, durations_s):
            f = ta_kaldi.fbank(w, num_mel_bins=80, sample_frequency=sr).numpy()
            np.divide(np.subtract(f, mean), std)
            audio += float(d)
    return audio, time.perf_counter() - t0


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, bounded sample per step.

    Griffin-Lim workloads: what is timed is oracle/conv_formulation.py, the reference's own formulation (dense-basis
    conv1d / conv_transpose1d, per-call window-sum-square loop, one utterance per call like
    speech_generator_for_s2st.py:115-124, torch intra-op threads = all cores), which reproduces the reference's golden
    waveforms bit for bit (tests/test_oracle_golden.py).  The initial phases are pre-drawn outside the timed loop
    (the GPU arm's are uploaded from pre-drawn host buffers too).  The cheaper numpy FFT oracle on all cores is
    reported beside it.  Front-end workload: torchaudio's kaldi.fbank + numpy CMVN, which IS the reference's code."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    workload = resolve_workload(args)
    if workload == "frontend":
        durs = list(np.random.RandomState(0).uniform(8, 20, 10000)[:24])
        times, audio = [], 0.0
        for step in range(args.warmup + args.steps):
            audio, dt = frontend_cpu_time(durs, 16000, cores, min_seconds=5.0)
            if step >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        value = audio / (ms / 1e3)
        print(json.dumps({
            "impl": "reference", "metric": "fbank80_cmvn_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_FRONTEND, "sample_rate": 16000, "n_bins": 80},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "reference",
                             "sample": f"the first {len(durs)} utterances of the list, repeated ({audio:.0f} audio-s per step), "
                                       "torchaudio compliance.kaldi.fbank + numpy global CMVN, one utterance per call"},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    from oracle import conv_formulation as ocf
    from oracle import griffin_lim as ogl
    sharded = workload == "gl_sharded"
    frames = sorted(sharded_frames(0)) if sharded else batch_frames(0)
    n_sample = 4
    sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, n_sample)]
    basis = ogl.pinv_mel_basis(SR, N_FFT, N_MELS, F_MIN, F_MAX)
    gl = ocf.ConvGriffinLim(N_FFT, WIN, HOP, N_ITER)
    inputs = []
    for i, T in enumerate(sample):
        np.random.seed(1000 + i)
        inputs.append((synth_logmel_np(T, 1000 + i), ogl.random_phase((N_BINS, T))))
    times, audio = [], 0.0
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        audio = 0.0
        with torch.no_grad():
            for x, ph in inputs:
                audio += ocf.vocoder_forward(x, ph, N_ITER, basis, gl=gl).shape[0] / SR
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = audio / (ms / 1e3)
    fft_value, fft_procs, fft_n, fft_frames = _fft_port_all_cores(frames, cores)
    sample_desc = (f"{n_sample} of the {len(frames)} utterances (length-stratified, {sum(sample)} frames), {N_ITER} iters, one "
                   f"utterance per call, the reference's dense-convolution formulation (oracle/conv_formulation.py, "
                   f"bit-identical to the reference's golden waveforms), torch CPU with {torch.get_num_threads()} threads")
    print(json.dumps({
        "impl": "reference", "metric": "griffin_lim_audio_seconds_per_second", "value": value,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_SHARDED if sharded else WORKLOAD, "n_iter": N_ITER, "sample_rate": SR,
                   "n_fft": N_FFT, "hop": HOP, "win": WIN},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample_desc,
                         "fft_oracle_all_cores": {"value": fft_value, "unit": "audio-s/s", "cores": fft_procs,
                                                  "sample": f"{fft_n} utterances ({fft_frames} frames), numpy FFT oracle, "
                                                            "one process per core"}},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- shared GPU plumbing -------------------------------------------------------------------------------------------
class Ctx:
    """Process-group / device context of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == args.gpus or self.world == 1, (self.world, args.gpus)
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        import __graft_entry__
        __graft_entry__.build()
        self.pkg = importlib.import_module(PKG)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world > 1:
            t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def gather_floats(self, xs):
        """Every rank's list of floats on every rank: [world, len(xs)]."""
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return t[None].cpu().numpy()
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return self.torch.stack(out).cpu().numpy()

    def timed(self, fn, steps, sampler=None):
        """EXACTLY `steps` calls of fn between two events, barrier + synchronize on both sides; ms per step (max over
        ranks) and the host wall time per step."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if sampler:
            sampler.sample_while(e1)
        self.barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0) / steps
        return self.max_over_ranks(e0.elapsed_time(e1) / steps), self.max_over_ranks(wall_ms)

    def warm(self, fn, steps, min_seconds=1.0):
        """`steps` untimed warm-up calls, then more of them until the GPU has been busy for `min_seconds`: a fresh box
        starts at idle clocks and needs several hundred milliseconds of load to reach its boost clock -- five 15 ms
        steps are not enough, and a timed region that starts on a ramping clock measures the ramp (seen as 17-20 ms
        per step against 15.3 ms a second later).  Returns the number of warm-up calls made."""
        torch = self.torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        took = self.max_over_ranks(time.perf_counter() - t0)  # the same number on every rank: fn may hold collectives
        extra = int(np.ceil(max(0.0, min_seconds - took) / max(took / max(steps, 1), 1e-4)))
        for _ in range(extra):
            fn()
        torch.cuda.synchronize()
        return steps + extra

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_vocoder(ctx, n_iter=N_ITER):
    voc = ctx.pkg.GriffinLimVocoder(SR, WIN, HOP, N_FFT, N_MELS, F_MIN, F_MAX, ctx.torch.hann_window,
                                    spec_bwd_max_iter=n_iter).to(ctx.dev)
    return voc, voc._plan(ctx.dev)


def pass_profile(ctx, plan, run_once, steps):
    """Per-launch device time of the Griffin-Lim passes: CUDA events recorded by the library on the launching stream
    around every pass, over `steps` extra runs (kept out of `value`'s region so the event records cannot perturb it)."""
    plan.set_pass_timing(True)
    per_step = []
    ctx.barrier()
    for _ in range(steps):
        run_once()
        per_step.append(plan.pass_times_ms())
    plan.set_pass_timing(False)
    return np.mean(np.stack(per_step), axis=0)


# ---- workload: config 2 on every GPU (the N = 1 headline) ---------------------------------------------------------
def run_gl(args, ctx):
    torch = ctx.torch
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    voc, plan = make_vocoder(ctx)
    frames, logmel_np, phase_np = config2_batch(rank)  # every rank the same lengths, its own content (weak scaling)
    total = int(sum(frames))
    audio_s = sum((T - 1) * HOP for T in frames) / SR
    logmel_h = torch.from_numpy(logmel_np).pin_memory()
    phase_h = torch.from_numpy(phase_np).pin_memory()
    n_samples = (total - len(frames)) * HOP
    wave_h = torch.empty(n_samples, dtype=torch.float32).pin_memory()
    logmel_d, phase_d = logmel_h.to(dev), phase_h.to(dev)

    # ---- device-resident timing (value) ----------------------------------------------------------
    sampler = ClockSampler(ctx.local_rank) if (rank == 0 and not os.environ.get("BENCH_NO_CLOCKS")) else None
    keep = {}

    def step():
        keep["wave"] = voc.synthesize_flat(logmel_d, frames, phase_d)

    warm_steps = ctx.warm(step, args.warmup)

    ms_step, _ = ctx.timed(step, args.steps, sampler)
    clocks = sampler.stop() if sampler else None
    assert torch.isfinite(keep["wave"]).all()
    # sustained check: the same step back to back for >= 2 s, so a power-cap droop would show
    n_sus = max(args.steps, int(np.ceil(2000.0 / ms_step)))
    ms_sus, _ = ctx.timed(step, n_sus)
    last_pass_ms = pass_profile(ctx, plan, step, min(args.steps, 20))
    # one launch per iteration: [initial inverse, iteration 1, ..., iteration n]
    iters_per_launch = 1
    iter_ms = float(np.mean(last_pass_ms[1:])) if len(last_pass_ms) > 1 else float("nan")  # per LAUNCH
    launches = plan.gl_launch_count(N_ITER, True) * args.steps

    # ---- end to end through the public API, host buffers in and out (e2e) -------------------------
    # Every step: H2D of that step's log-mel AND of the seeded initial phase from pinned memory (north_star: "seeded
    # initial phase supplied from the host"), synthesis, D2H of the waveforms.  Steps are fed back to back like
    # generate_waveform.py feeds batches: synthesize_host() uploads / downloads on copy streams, so step i's D2H
    # overlaps step i+1's H2D and kernels; two host output buffers alternate, and the timed region ends only when every
    # download has landed.  The variant that lets the library draw the phase on the device is reported beside it.
    wave_h2 = torch.empty_like(wave_h).pin_memory()
    e2e_bufs, e2e_events = [wave_h, wave_h2], [None, None]
    e2e_count = [0]

    def e2e_step(host_phase):
        i = e2e_count[0] % 2
        e2e_count[0] += 1
        if e2e_events[i] is not None:
            e2e_events[i].synchronize()  # the buffer's previous download must be complete before it is reused
        e2e_events[i] = voc.synthesize_host(logmel_h, frames, e2e_bufs[i], device=dev,
                                            phase_host=phase_h if host_phase else None, seed=1234)

    def time_e2e(host_phase):
        for _ in range(min(args.warmup, 3)):
            e2e_step(host_phase)
        dev_ms, wall_ms = ctx.timed(lambda: e2e_step(host_phase), args.steps)
        return max(dev_ms, wall_ms)

    e2e_dev_ms = time_e2e(False)
    e2e_host_ms = time_e2e(True)
    assert torch.isfinite(wave_h).all() and torch.isfinite(wave_h2).all()
    e2e_count[0] = 0
    e2e_step(True)  # leave the host-phase result in wave_h for the parity spot check below
    torch.cuda.synchronize()

    extras = {}
    if rank == 0 and world == 1 and not os.environ.get("BENCH_NO_EXTRAS"):
        extras = gl_extras(ctx, voc)

    if rank != 0:
        ctx.close()
        return

    hbm_peak, peak_src = peaks()
    algo_bytes_iter = ALGO_BYTES_PER_FRAME_ITER * total * iters_per_launch  # per launch
    achieved = algo_bytes_iter / (iter_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic_bytes()
    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ---------------------
    from oracle import griffin_lim as ogl
    basis = voc.inv_mel_transform.basis.cpu().numpy()
    sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, 6)]
    cpu_audio, cpu_s = cpu_port_time(sample, 4321, N_ITER, basis) if world == 1 else (0.0, 0.0)  # N = 1 only
    # parity spot check on the way (checker only): the LONGEST utterance of the batch (most strips, 64 iterations,
    # automatic strip length) through the e2e path with the host phase, against the oracle
    Tl = frames[-1]
    f0 = total - Tl
    ref = ogl.vocoder_forward(logmel_np[f0:], np.ascontiguousarray(phase_np[f0:].T), N_ITER, basis=basis)
    parity = ogl.rel_l2(wave_h[n_samples - (Tl - 1) * HOP: n_samples].numpy(), ref)

    out = {
        "metric": "griffin_lim_audio_seconds_per_second", "value": world * audio_s / (ms_step * 1e-3),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warm_steps,
        "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_gpu": N_UTTS, "frames_per_gpu": total,
                   "audio_seconds_per_gpu": audio_s, "n_iter": N_ITER, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP,
                   "win": WIN, "l2_policy": "per-step working set (magnitudes + phase + waveform buffers, "
                   f"{(total * (684 + 1025) * 4 + 4 * n_samples * 4) / 1e6:.0f} MB) exceeds the 126 MB L2"},
        "e2e": {"value": world * audio_s / (e2e_host_ms * 1e-3), "unit": "audio-s/s",
                "h2d_bytes_per_step": int(logmel_h.numel() * 4 + phase_h.numel() * 4),
                "d2h_bytes_per_step": int(wave_h.numel() * 4), "ms_per_step": e2e_host_ms,
                "api": "GriffinLimVocoder.synthesize_host(pinned log-mel + pinned host-drawn seeded initial phase in, "
                       "pinned waveforms out; H2D and D2H run on copy streams and overlap the kernels of the "
                       "neighbouring steps)",
                "with_device_drawn_phase": {"value": world * audio_s / (e2e_dev_ms * 1e-3), "ms_per_step": e2e_dev_ms,
                                            "h2d_bytes_per_step": int(logmel_h.numel() * 4)}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_gl_pass<19,false,true,true> (fused iSTFT+OLA+normalise+STFT+magnitude re-imposition)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic * iters_per_launch if traffic else None,
                     "traffic_source": ("ncu --set full capture of one iteration on this workload, committed as %s "
                                        "(ncu cannot run inside the timed benchmark)" % traffic_src) if traffic else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes_iter,
                     "launch_ms": iter_ms, "launches_per_step": N_ITER // iters_per_launch,
                     "iterations_per_launch": iters_per_launch, "ms_per_iteration": iter_ms / iters_per_launch,
                     "share_of_step": float(np.sum(last_pass_ms[1:]) / ms_step) if len(last_pass_ms) > 1 else None,
                     "first_pass_ms": float(last_pass_ms[0])},
        "sustained": {"steps": n_sus, "seconds": n_sus * ms_sus * 1e-3, "ms_per_step": ms_sus,
                      "value": world * audio_s / (ms_sus * 1e-3)},
        "cpu_baseline": None if world > 1 else {
            "value": cpu_audio / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port",
            "sample": f"6 length-stratified utterances of the batch ({sum(sample)} frames), {N_ITER} iters, "
                      f"numpy FFT oracle, {cpu_s:.1f} s"},
        "clocks": clocks,
        "parity_rel_l2_vs_oracle": parity,
        "parity_checked": f"longest utterance of the batch ({Tl} frames) through the e2e path with the host phase",
    }
    out.update(extras)
    print(json.dumps(out))
    ctx.close()


def gl_extras(ctx, voc):
    """Bounded extra measurements for the driver's view (N = 1): BASELINE config 1 through the reference's own call
    shape (one utterance per GriffinLimVocoder.forward), config 5 (one 60 s utterance, 64 and 256 iterations), and the
    tensor-core inverse-mel projection."""
    torch = ctx.torch
    dev = ctx.dev
    out = {}

    def timeit(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / n

    # config 1: T = 500, 64 iterations.  (a) forward(): phase from numpy's global RNG on the host like vocoder.py:103;
    # (b) the same utterance with the phase already on the device (what the kernels cost)
    T = 500
    x = torch.from_numpy(synth_logmel_np(T, 77)).to(dev)
    audio = (T - 1) * HOP / SR
    np.random.seed(0)
    ms_fwd = timeit(lambda: voc(x), 10)
    ph = torch.from_numpy(config2_batch(0)[2][:T].copy()).to(dev)
    ms_dev = timeit(lambda: voc.synthesize_flat(x, [T], ph), 20)
    out["config1"] = {"workload": "one 500-frame utterance, 64 iters", "forward_ms": ms_fwd,
                      "forward_audio_s_per_s": audio / (ms_fwd * 1e-3), "device_phase_resident_ms": ms_dev,
                      "device_phase_resident_audio_s_per_s": audio / (ms_dev * 1e-3),
                      "note": "small calls run all iterations in ONE cooperative launch, a warp per frame (k_gl_frames); "
                              "forward() continues numpy's global MT19937 on the device (vocoder.py:103 draws 1025 x T "
                              "float64 uniforms from it) instead of drawing on the host; round 1: 3.16 / 1.65 ms"}
    # config 5: one 60 s utterance
    T = 4800
    x5 = torch.from_numpy(synth_logmel_np(T, 78)).to(dev)
    ph5 = ((torch.rand(T, N_BINS, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 2 - 1) * np.pi).contiguous()
    audio = (T - 1) * HOP / SR
    c5 = {"workload": "one 60 s utterance (4800 frames)"}
    for n_iter in (64, 256):
        ms = timeit(lambda: voc.synthesize_flat(x5, [T], ph5, n_iter=n_iter), 3)
        c5[f"{n_iter}_iters"] = {"ms": ms, "audio_s_per_s": audio / (ms * 1e-3),
                                 "hbm_frac": (ALGO_BYTES_PER_FRAME_ITER * n_iter + ALGO_BYTES_PER_FRAME_ONCE) * T / (ms * 1e-3) / 1e9 / peaks()[0]}
    out["config5"] = c5
    # the two dense contractions of the path on the tensor cores (north_star job 4): device time of the stand-alone
    # entries on the config-2 batch, algorithmic FLOPs (2 * n_mels * n_bins per frame) against the measured bf16 peak;
    # the 3 x TF32 split executes three times those FLOPs at the TF32 rate (half the bf16 rate)
    frames, logmel_np, _ = config2_batch(0)
    total = int(sum(frames))
    lm = torch.from_numpy(logmel_np).to(dev)
    lib, ptr, sptr = ctx.pkg._lib.load(), ctx.pkg._lib.ptr, ctx.pkg._lib.stream_ptr
    plan = voc._plan(dev)
    mag = torch.empty(total, N_BINS, device=dev)
    ms_inv = timeit(lambda: ctx.pkg._lib.check(lib.s2st_inverse_mel(plan.handle, total, ptr(lm), 1, ptr(mag), sptr(dev)), "s2st_inverse_mel"), 20)
    plans = importlib.import_module(PKG + ".plans")
    mplan = plans.get_stft_plan(dev, N_FFT, N_FFT, N_FFT // 4, N_MELS, torch.ones(N_FFT),
                                mel=ctx.pkg.get_mel_filters(SR, N_FFT, N_MELS, F_MIN, F_MAX))
    mel_out = torch.empty(total, N_MELS, device=dev)
    ms_mel = timeit(lambda: ctx.pkg._lib.check(lib.s2st_mel_project(mplan.handle, total, ptr(mag), ptr(mel_out), sptr(dev)), "s2st_mel_project"), 20)
    tpeak, tsrc = tensor_peak()
    flops = 2.0 * N_MELS * N_BINS * total
    out["tensor"] = {
        "peak_tflops": tpeak, "peak_source": tsrc, "frames": total,
        "note": "3 x TF32 split: executed tensor FLOPs = 3 x algorithmic, at the TF32 rate (half the bf16 peak); both kernels "
                "are bound by their HBM traffic, not by the tensor pipe (ncu sm__pipe_tensor_cycles_active in profiles/)",
        "inverse_mel": {"kernel": "k_inverse_mel_tc (tcgen05 kind::tf32, exp + pinv[1025x80] + clamp)", "ms": ms_inv,
                        "algorithmic_tflops": flops / (ms_inv * 1e-3) / 1e12, "frac_of_peak": flops / (ms_inv * 1e-3) / 1e12 / tpeak,
                        "hbm_frac": total * (320 + 4 * N_BINS) / (ms_inv * 1e-3) / 1e9 / peaks()[0]},
        "mel_project": {"kernel": "k_mel_project_tc (tcgen05 kind::tf32, mel[80x1025])", "ms": ms_mel,
                        "algorithmic_tflops": flops / (ms_mel * 1e-3) / 1e12, "frac_of_peak": flops / (ms_mel * 1e-3) / 1e12 / tpeak,
                        "hbm_frac": total * (320 + 4 * N_BINS) / (ms_mel * 1e-3) / 1e9 / peaks()[0]}}
    return out


# ---- workload: config 4, one list sharded over the ranks -----------------------------------------------------------
def run_gl_sharded(args, ctx):
    torch, dist = ctx.torch, ctx.dist
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    sh = importlib.import_module(PKG + ".sharding")
    voc, plan = make_vocoder(ctx)
    frames_all = sharded_frames(0)
    owned = sh.shard_utterances(frames_all, world, N_ITER)
    mine = owned[rank]
    all_buckets = [sh.length_buckets(frames_all, owned[r], BUCKET_FRAMES, balanced=True) for r in range(world)]
    buckets = all_buckets[rank]
    # every rank knows every rank's buckets (the sharding is a pure function of the list): the gather needs no metadata
    # exchange, and bucket b of all ranks travels while bucket b + 1 is synthesised
    same_count = all(len(b) == len(buckets) for b in all_buckets)
    layouts = [[(all_buckets[r][b], [(frames_all[i] - 1) * HOP for i in all_buckets[r][b]]) for r in range(world)]
               for b in range(len(buckets))] if same_count else None
    audio_all = sum((T - 1) * HOP for T in frames_all) / SR
    my_frames = sum(frames_all[i] for i in mine)

    # inputs per bucket: pinned host (e2e) and device-resident (value)
    host, devb = [], []
    for b in buckets:
        fr = [frames_all[i] for i in b]
        xs, ps = zip(*(sharded_utterance_inputs(i, frames_all[i]) for i in b))
        lm = torch.from_numpy(np.concatenate(xs)).pin_memory()
        ph = torch.from_numpy(np.concatenate(ps)).pin_memory()
        wh = torch.empty((sum(fr) - len(fr)) * HOP, dtype=torch.float32).pin_memory()
        host.append((fr, lm, ph, wh))
        devb.append((fr, lm.to(dev), ph.to(dev)))
    del xs, ps

    keep = {}

    def synth():
        keep["waves"] = [voc.synthesize_flat(lm, fr, ph) for fr, lm, ph in devb]

    ids_local = [i for b in buckets for i in b]
    lens_local = [(frames_all[i] - 1) * HOP for i in ids_local]

    def gather():
        """All of this rank's waveforms in one exchange, after the synthesis (the non-overlapped form)."""
        flat = keep["waves"][0] if len(keep["waves"]) == 1 else torch.cat(keep["waves"])
        st = {}
        keep["gathered"] = [sh.gather_waveforms(ids_local, flat, len(frames_all), dst=0, stats=st, local_lengths=lens_local)]
        keep["gather_stats"] = st

    def step():
        """The whole job: synthesise bucket by bucket; the waveforms of bucket b start for rank 0 as soon as they exist
        and travel while bucket b + 1 is synthesised; the step ends when everything has landed."""
        if world == 1 or layouts is None:
            synth()
            if world > 1:
                gather()
            return
        handles, waves, st = [], [], {}
        for b, (fr, lm, ph) in enumerate(devb):
            w = voc.synthesize_flat(lm, fr, ph)
            waves.append(w)
            s_b = {}
            handles.append(sh.gather_waveforms(buckets[b], w, len(frames_all), dst=0, stats=s_b, local_lengths=layouts[b][rank][1],
                                               layout=layouts[b], async_op=True))
            st["bytes_to_dst"] = st.get("bytes_to_dst", 0) + s_b["bytes_to_dst"]
        keep["waves"] = waves
        keep["gathered"] = [h.wait() for h in handles]
        keep["gather_stats"] = st

    warm_steps = ctx.warm(step, args.warmup)
    sampler = ClockSampler(ctx.local_rank) if (rank == 0 and not os.environ.get("BENCH_NO_CLOCKS")) else None
    ms_step, _ = ctx.timed(step, args.steps, sampler)
    clocks = sampler.stop() if sampler else None
    ms_nogather, _ = ctx.timed(synth, args.steps)
    # the gather alone (device time on each rank, max over ranks)
    ms_gather = ctx.timed(gather, max(3, args.steps // 4))[0] if world > 1 else 0.0
    # per-rank compute time of one pass over its shard (no barrier inside): the imbalance the sharding leaves
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    synth()
    e1.record()
    torch.cuda.synchronize()
    per_rank = ctx.gather_floats([my_frames * (N_ITER + 1), e0.elapsed_time(e1), len(mine), len(buckets)])
    # correctness of what arrived on rank 0: utterances synthesised on other ranks equal rank 0's own synthesis of the
    # same inputs (spot check on three of them; run-to-run determinism makes this bitwise when the strip length agrees)
    gather_check = None
    if world > 1:
        step()  # every rank takes part; rank 0 keeps what it received
        ctx.barrier()
    if rank == 0 and world > 1:
        picks = [owned[r][0] for r in range(1, world)][:3]
        errs = []
        for i in picks:
            x, p = sharded_utterance_inputs(i, frames_all[i])
            alone = voc.synthesize_flat(torch.from_numpy(x).to(dev), [frames_all[i]], torch.from_numpy(p).to(dev))
            got = [g for g in keep["gathered"] if i in g._where][0][i]
            errs.append(float((alone - got).norm() / alone.norm()))
        gather_check = {"utterances": picks, "max_rel_l2_vs_local_resynthesis": max(errs)}

    # roofline of the iteration kernel over this rank's buckets
    iter_ms_sum, iter_bytes, first_ms_sum = 0.0, 0.0, 0.0
    for fr, lm, ph in devb:
        prof = pass_profile(ctx, plan, lambda: voc.synthesize_flat(lm, fr, ph), 2) if True else None
        iter_ms_sum += float(np.sum(prof[1:]))
        first_ms_sum += float(prof[0])
        iter_bytes += ALGO_BYTES_PER_FRAME_ITER * sum(fr) * N_ITER
    launches = sum(plan.gl_launch_count(N_ITER, True) for _ in devb) * args.steps

    # ---- e2e: pinned host in, per-rank pinned host out (what the reference's shards do: every shard keeps its own
    # waveforms), uploads / downloads on copy streams
    events = []

    def e2e_step():
        for e in events:
            e.synchronize()
        events.clear()
        for fr, lm, ph, wh in host:
            events.append(voc.synthesize_host(lm, fr, wh, device=dev, phase_host=ph))

    for _ in range(2):
        e2e_step()
    e2e_dev_ms, e2e_wall_ms = ctx.timed(e2e_step, max(2, args.steps // 2))
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)
    h2d = sum(lm.numel() * 4 + ph.numel() * 4 for _, lm, ph, _ in host)
    d2h = sum(wh.numel() * 4 for *_, wh in host)
    tot_bytes = ctx.gather_floats([h2d, d2h]).sum(axis=0)

    if rank != 0:
        ctx.close()
        return
    hbm_peak, peak_src = peaks()
    achieved = iter_bytes / (iter_ms_sum * 1e-3) / 1e9
    loads = per_rank[:, 0]
    out = {
        "metric": "griffin_lim_audio_seconds_per_second", "value": audio_all / (ms_step * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warm_steps, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_SHARDED, "utterances": len(frames_all), "frames": int(sum(frames_all)),
                   "audio_seconds": audio_all, "n_iter": N_ITER, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP, "win": WIN,
                   "bucket_frames": BUCKET_FRAMES, "buckets_per_rank": [int(x) for x in per_rank[:, 3]],
                   "l2_policy": "per-call working set (>= 1 GB per bucket) exceeds the 126 MB L2"},
        "collective": {"name": "gather_waveforms: exact-size point-to-point sends to rank 0, one NCCL group per bucket, started "
                               "as soon as the bucket is synthesised and overlapped with the next bucket (no data-path "
                               "collective, no metadata exchange: every rank knows the sharding)" if world > 1 else "none (single rank)",
                       "bytes_to_rank0": keep.get("gather_stats", {}).get("bytes_to_dst", 0),
                       "ms_not_overlapped": ms_gather, "ms_exposed_in_step": ms_step - ms_nogather,
                       "check": gather_check},
        "without_gather": {"ms_per_step": ms_nogather, "value": audio_all / (ms_nogather * 1e-3)},
        "sharding": {"policy": "LPT on frames x (n_iter + 1), then balanced length buckets per rank",
                     "frame_iterations_per_rank": [int(x) for x in loads],
                     "imbalance_max_over_mean": float(loads.max() / loads.mean()),
                     "compute_ms_per_rank": [float(x) for x in per_rank[:, 1]],
                     "utterances_per_rank": [int(x) for x in per_rank[:, 2]]},
        "e2e": {"value": audio_all / (e2e_ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": int(tot_bytes[0]),
                "d2h_bytes_per_step": int(tot_bytes[1]), "ms_per_step": e2e_ms,
                "api": "per rank and bucket GriffinLimVocoder.synthesize_host(pinned log-mel + pinned seeded initial phase "
                       "in, pinned waveforms out on the rank that synthesised them -- the reference's shards keep their "
                       "own outputs too, generate_waveform.py:166-167)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_gl_pass<19,false,true,true> (fused iSTFT+OLA+normalise+STFT+magnitude "
                                               "re-imposition), rank 0's buckets",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": iter_bytes, "iteration_ms_per_step": iter_ms_sum,
                     "share_of_step": iter_ms_sum / ms_nogather, "first_pass_ms_per_step": first_ms_sum},
        "cpu_baseline": None,
        "clocks": clocks,
    }
    print(json.dumps(out))
    ctx.close()


# ---- workload: config 3, the feature front-end ----------------------------------------------------------------------
def run_frontend(args, ctx):
    torch = ctx.torch
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    pkg = ctx.pkg
    plans = importlib.import_module(PKG + ".plans")
    lib = pkg._lib.load()
    ptr, check, sptr = pkg._lib.ptr, pkg._lib.check, pkg._lib.stream_ptr
    n_utts = int(os.environ.get("BENCH_FRONTEND_UTTS", "10000"))
    rng = np.random.RandomState(0 + rank)
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    mean = (torch.randn(80, device=dev, generator=g) - 4).contiguous()
    std = (torch.rand(80, device=dev, generator=g) * 1.5 + 0.5).contiguous()

    def synth_audio_dev(lens, sr, scale):
        """The SURVEY 8(d) signal built on the device (setup plumbing): noise x 0.1 + three sinusoids, clipped."""
        n = int(lens.sum())
        x = 0.1 * torch.randn(n, device=dev, generator=g)
        t = torch.arange(n, device=dev, dtype=torch.float32) / sr
        for _ in range(3):
            x += float(rng.uniform(0.05, 0.3)) * torch.sin(2 * np.pi * float(rng.uniform(80, 0.45 * sr)) * t + float(rng.uniform(0, 6.28)))
        return (x.clamp_(-1, 1) * scale).contiguous()

    def offsets(lens, frames):
        fo = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)).to(dev)
        wo = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
        return fo, wo

    def fbank_case(sr, n):
        lens = (rng.uniform(8, 20, n) * sr).astype(np.int64)
        plan = plans.get_fbank_plan(dev, sr, 80)
        frames = [1 + (int(k) - plan.win) // plan.shift for k in lens]
        flat = synth_audio_dev(lens, sr, 2.0 ** 15)
        fo, wo = offsets(lens, frames)
        total = int(sum(frames))
        o = torch.empty(total, 80, device=dev)

        def run(src=flat, dst=o):
            check(lib.s2st_fbank(plan.handle, n, total, ptr(wo), ptr(fo), ptr(src), ptr(mean), ptr(std), None, ptr(dst),
                                 sptr(dev)), "s2st_fbank")
        return run, float(lens.sum()) / sr, total, flat, o

    res = {}
    hbm_peak, peak_src = peaks()
    # headline: fbank80 + CMVN at 16 kHz over the whole list
    run16, audio16, frames16, flat16, out16 = fbank_case(16000, n_utts)
    ctx.warm(run16, args.warmup)
    sampler = ClockSampler(ctx.local_rank) if (rank == 0 and not os.environ.get("BENCH_NO_CLOCKS")) else None
    ms16, _ = ctx.timed(run16, args.steps, sampler)
    clocks = sampler.stop() if sampler else None
    assert torch.isfinite(out16).all()
    # e2e: pinned host waveform in, pinned host features out, chunked so that upload, kernel and download overlap
    flat_h = flat16.cpu().pin_memory()
    out_h = torch.empty(out16.shape, dtype=torch.float32).pin_memory()
    # (the front-end C entry takes one ragged batch; the host shim pipelines equal shares of the list)
    n_chunks = 8
    e2e_ms = frontend_e2e(ctx, lib, pkg, plans, flat_h, out_h, n_utts, 16000, mean, std, n_chunks, args, rng_seed=rank)
    # the same audio as the 16-bit PCM it would be on disk (get_fbank feeds Kaldi's fbank int16-range samples)
    flat_h16 = flat_h.round().clamp_(-32768, 32767).to(torch.int16).pin_memory()
    e2e16_ms = frontend_e2e(ctx, lib, pkg, plans, flat_h16, out_h, n_utts, 16000, mean, std, n_chunks, args, rng_seed=rank)
    del flat_h16
    res16 = {"ms": ms16, "audio_s_per_s": audio16 / (ms16 * 1e-3), "frames": frames16,
             "hbm_frac": frames16 * FBANK_BYTES_PER_FRAME / (ms16 * 1e-3) / 1e9 / hbm_peak}
    del flat16, out16, run16
    torch.cuda.empty_cache()
    # beside it: fbank80 at 8 kHz and logmelspec80 at 24 kHz on a fifth of the list
    n_side = max(1, n_utts // 5)
    run8, audio8, frames8, _f8, _o8 = fbank_case(8000, n_side)
    for _ in range(3):
        run8()
    ms8, _ = ctx.timed(run8, max(3, args.steps // 2))
    res["fbank80_cmvn_8k"] = {"utterances": n_side, "ms": ms8, "audio_s_per_s": audio8 / (ms8 * 1e-3), "frames": frames8,
                              "hbm_frac": frames8 * FBANK_BYTES_PER_FRAME / (ms8 * 1e-3) / 1e9 / hbm_peak}
    del run8, _f8, _o8
    lens = (rng.uniform(8, 20, n_side) * SR).astype(np.int64)
    flat = synth_audio_dev(lens, SR, 1.0)
    frames = [1 + int(k) // HOP for k in lens]
    total = int(sum(frames))
    planl = plans.get_stft_plan(dev, N_FFT, WIN, HOP, N_MELS, torch.hann_window(WIN),
                                mel=pkg.get_mel_filters(SR, N_FFT, N_MELS, F_MIN, F_MAX))
    fo, wo = offsets(lens, frames)
    o = torch.empty(total, N_MELS, device=dev)

    def runl():
        check(lib.s2st_logmel(planl.handle, n_side, total, ptr(wo), ptr(fo), ptr(flat), 1e-5, ptr(mean), ptr(std),
                              None, ptr(o), sptr(dev)), "s2st_logmel")
    for _ in range(3):
        runl()
    msl, _ = ctx.timed(runl, max(3, args.steps // 2))
    res["logmelspec80_cmvn_24k"] = {"utterances": n_side, "ms": msl, "audio_s_per_s": float(lens.sum()) / SR / (msl * 1e-3),
                                    "frames": total, "hbm_frac": total * LOGMEL_BYTES_PER_FRAME / (msl * 1e-3) / 1e9 / hbm_peak}
    if rank != 0:
        ctx.close()
        return
    cores = os.cpu_count() or 1
    durs = list(np.random.RandomState(0).uniform(8, 20, 10000)[:16])
    cpu_audio, cpu_s = frontend_cpu_time(durs, 16000, cores) if world == 1 else (0.0, 1.0)
    achieved = frames16 * FBANK_BYTES_PER_FRAME / (ms16 * 1e-3) / 1e9
    out = {
        "metric": "fbank80_cmvn_audio_seconds_per_second", "value": world * audio16 / (ms16 * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms16, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_FRONTEND, "utterances_per_gpu": n_utts, "frames_per_gpu": frames16,
                   "audio_seconds_per_gpu": audio16, "sample_rate": 16000, "n_bins": 80,
                   "l2_policy": f"inputs + outputs ({(audio16 * 16000 * 4 + frames16 * 320) / 1e9:.1f} GB) exceed the 126 MB L2"},
        "e2e": {"value": world * audio16 / (e2e_ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": int(flat_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": e2e_ms,
                "api": f"fbank80 + fused CMVN, pinned host waveforms in, pinned host features out, {n_chunks} chunks "
                       "pipelined over copy streams (PCIe-bound: 4 B in per sample)",
                "pcm16_in": {"value": world * audio16 / (e2e16_ms * 1e-3), "ms_per_step": e2e16_ms,
                             "h2d_bytes_per_step": int(flat_h.numel() * 2),
                             "api": "the same, with the waveforms uploaded as the 16-bit PCM they are on disk and converted "
                                    "on the device (s2st_pcm16_to_wave)"}},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "kernel": "k_fbank_fast<0> (16 kHz: DC removal, pre-emphasis, povey window, FFT-512, "
                                               "power, mel, log, CMVN)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": frames16 * FBANK_BYTES_PER_FRAME,
                     "launch_ms": ms16},
        "cpu_baseline": None if world > 1 else {
            "value": cpu_audio / cpu_s, "unit": "audio-s/s", "cores": cores, "kind": "reference",
            "sample": f"the first {len(durs)} utterances of the list, repeated for {cpu_s:.1f} s ({cpu_audio:.0f} audio-s): "
                      "torchaudio compliance.kaldi.fbank + numpy CMVN (the reference's own code path, audio_utils.py:141-147, "
                      "global_cmvn.py:26-29)"},
        "clocks": clocks,
        "fbank80_cmvn_16k": res16,
    }
    out.update(res)
    print(json.dumps(out))
    ctx.close()


def frontend_e2e(ctx, lib, pkg, plans, flat_h, out_h, n_utts, sr, mean, std, n_chunks, args, rng_seed):
    """Pinned host waveforms -> fbank80 + CMVN -> pinned host features, utterance chunks pipelined over an upload
    stream, the compute stream and a download stream.  flat_h float32, or int16 PCM (the samples as they are on disk:
    2 bytes each over PCIe, converted on the device by s2st_pcm16_to_wave)."""
    pcm16 = flat_h.dtype == ctx.torch.int16
    torch = ctx.torch
    dev = ctx.dev
    ptr, check = pkg._lib.ptr, pkg._lib.check
    rng = np.random.RandomState(0 + rng_seed)
    lens = (rng.uniform(8, 20, n_utts) * sr).astype(np.int64)  # the same draw as fbank_case(16000, n_utts)
    plan = plans.get_fbank_plan(dev, sr, 80)
    frames = np.array([1 + (int(k) - plan.win) // plan.shift for k in lens])
    bounds = np.linspace(0, n_utts, n_chunks + 1).astype(int)
    wo_all = np.concatenate([[0], np.cumsum(lens)])
    fo_all = np.concatenate([[0], np.cumsum(frames)])
    chunks = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        fo = torch.from_numpy((fo_all[a:b + 1] - fo_all[a]).astype(np.int32)).to(dev)
        wo = torch.from_numpy((wo_all[a:b + 1] - wo_all[a]).astype(np.int64)).to(dev)
        chunks.append((int(b - a), int(fo_all[b] - fo_all[a]), fo, wo, int(wo_all[a]), int(wo_all[b]), int(fo_all[a]), int(fo_all[b])))
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    def step():
        for n, total, fo, wo, w0, w1, f0, f1 in chunks:
            with torch.cuda.stream(up):
                src = flat_h[w0:w1].to(dev, non_blocking=True)
            main.wait_stream(up)
            src.record_stream(main)
            if pcm16:
                raw, src = src, torch.empty(src.numel(), dtype=torch.float32, device=dev)
                check(lib.s2st_pcm16_to_wave(raw.numel(), ptr(raw), 1.0, ptr(src), ctypes_stream(main)), "s2st_pcm16_to_wave")
            dst = torch.empty(total, 80, device=dev)
            check(lib.s2st_fbank(plan.handle, n, total, ptr(wo), ptr(fo), ptr(src), ptr(mean), ptr(std), None, ptr(dst),
                                 ctypes_stream(main)), "s2st_fbank")
            down.wait_stream(main)
            with torch.cuda.stream(down):
                out_h[f0:f1].copy_(dst, non_blocking=True)
            dst.record_stream(down)
        main.wait_stream(down)

    for _ in range(2):
        step()
    dev_ms, wall_ms = ctx.timed(step, max(2, args.steps // 4))
    return max(dev_ms, wall_ms)


def ctypes_stream(stream):
    import ctypes
    return ctypes.c_void_p(stream.cuda_stream)


def resolve_workload(args):
    if args.workload != "auto":
        return args.workload
    return "gl" if args.gpus == 1 else "gl_sharded"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "gl", "gl_sharded", "frontend"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    workload = resolve_workload(args)
    ctx = Ctx(args)
    {"gl": run_gl, "gl_sharded": run_gl_sharded, "frontend": run_frontend}[workload](args, ctx)


if __name__ == "__main__":
    main()
