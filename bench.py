#!/usr/bin/env python
"""Headline benchmark: Griffin-Lim audio-seconds per second (64 iterations, 24 kHz) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: BASELINE.json config 2, a Fisher-test-shaped
ragged batch of 256 synthetic log-mel utterances (T ~ U{56..400}, seed 0, length-sorted) taken
through inverse-mel, the initial inverse and 64 fused STFT/iSTFT Griffin-Lim iterations.  With N > 1
every rank runs its own batch of that shape (weak scaling, no collective on the data path).

Prints ONE JSON line (rank 0).  ``value`` is device-timed with inputs resident in HBM; ``e2e`` goes
through the public API from pinned host buffers (H2D of log-mel + initial phase, D2H of waveforms
inside the timed region); ``roofline`` is the fused iteration kernel against the measured HBM peak;
``cpu_baseline`` is the numpy oracle timed on this box's host cores on a bounded sample.
``--impl reference`` times that CPU port with all host cores instead (the reference is Python/PyTorch
CPU code that cannot travel to the GPU box).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "speech-to-speech-translation_b200"

SR, N_FFT, WIN, HOP, N_MELS, F_MIN, F_MAX, N_ITER = 24000, 2048, 1200, 300, 80, 20.0, 8000.0, 64
N_UTTS = 256
ALGO_BYTES_PER_FRAME_ITER = 6500        # 1200 B wave in + 4100 B magnitude + 1200 B wave out
ALGO_BYTES_PER_FRAME_ONCE = 13820       # inverse-mel + initial inverse
WORKLOAD = ("Fisher-test-shaped batch: 256 synthetic log-mel utterances, 56-400 frames length-bucketed, "
            "64 Griffin-Lim iters per GPU")


def batch_frames(seed):
    rng = np.random.RandomState(seed)
    return sorted(int(t) for t in rng.randint(56, 401, size=N_UTTS))


def synth_logmel_np(T, seed):
    """SURVEY 8(d) speech-like log-mel: smooth random walk in time + spectral tilt, clamped."""
    rng = np.random.RandomState(seed)
    x = 0.1 * np.cumsum(rng.randn(T, N_MELS), axis=0) + np.linspace(0, -4, N_MELS)[None, :] - 2.0
    return np.clip(x, np.log(1e-5), 2.0).astype(np.float32)


def measured_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_gl_pass launch on this workload, from the
    committed `ncu --set full` capture (profiles/r01_glpass_current_ncu_summary.txt); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01_glpass_current_ncu_summary.txt")
    try:
        tot = 0.0
        for line in open(path):
            if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
                val, unit = line.split("=")[1].split()[:2]
                tot += float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
        return tot or None
    except (OSError, KeyError, ValueError, IndexError):
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                "samples_under_load": getattr(self, "under_load", len(sm)), "reasons": sorted(self.reasons),
                "source": "NVML, sampled inside the timed region while the GPU drains the enqueued steps"}


# ------------------------------------------------------------------------------------------------
def cpu_port_time(frames_subset, seed, n_iter, basis):
    """Time the numpy oracle (the CPU port of the reference path) on the given utterances, 1 core."""
    from oracle import griffin_lim as ogl
    t0 = time.perf_counter()
    audio = 0.0
    for i, T in enumerate(frames_subset):
        x = synth_logmel_np(T, seed + i)
        np.random.seed(seed + i)
        phase = ogl.random_phase((N_FFT // 2 + 1, T))
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


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, bounded sample per step.

    What is timed is oracle/conv_formulation.py: the reference's own formulation (dense-basis conv1d /
    conv_transpose1d, per-call window-sum-square loop, one utterance per call like speech_generator_for_s2st.py:115-124,
    torch intra-op threads = all cores), which reproduces the reference's golden waveforms bit for bit
    (tests/test_oracle_golden.py).  The cheaper numpy FFT oracle on all cores is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import conv_formulation as ocf
    from oracle import griffin_lim as ogl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = batch_frames(0)
    n_sample = 4
    sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, n_sample)]
    basis = ogl.pinv_mel_basis(SR, N_FFT, N_MELS, F_MIN, F_MAX)
    gl = ocf.ConvGriffinLim(N_FFT, WIN, HOP, N_ITER)
    inputs = []
    for i, T in enumerate(sample):
        np.random.seed(1000 + i)
        inputs.append((synth_logmel_np(T, 1000 + i), ogl.random_phase((N_FFT // 2 + 1, T))))
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
    sample_desc = (f"{n_sample} of the {N_UTTS} utterances (length-stratified, {sum(sample)} frames), {N_ITER} iters, one "
                   f"utterance per call, the reference's dense-convolution formulation (oracle/conv_formulation.py, "
                   f"bit-identical to the reference's golden waveforms), torch CPU with {torch.get_num_threads()} threads")
    print(json.dumps({
        "impl": "reference", "metric": "griffin_lim_audio_seconds_per_second", "value": value,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_iter": N_ITER, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP, "win": WIN},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample_desc,
                         "fft_oracle_all_cores": {"value": fft_value, "unit": "audio-s/s", "cores": fft_procs,
                                                  "sample": f"{fft_n} utterances ({fft_frames} frames), numpy FFT oracle, "
                                                            "one process per core"}},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, (world, args.gpus)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    __graft_entry__.build()
    pkg = importlib.import_module(PKG)
    voc = pkg.GriffinLimVocoder(SR, WIN, HOP, N_FFT, N_MELS, F_MIN, F_MAX, torch.hann_window,
                                spec_bwd_max_iter=N_ITER).to(dev)
    plan = voc._plan(dev)

    frames = batch_frames(0)  # config 2 length law; every rank the same lengths, its own content (weak scaling)
    total = int(sum(frames))
    n_bins = N_FFT // 2 + 1
    audio_s = sum((T - 1) * HOP for T in frames) / SR
    # host (pinned) inputs: denormalised log-mel, frame-major, and the seeded initial phase, frame-major
    logmel_h = torch.from_numpy(np.concatenate([synth_logmel_np(T, 1234 + 1000 * rank + i) for i, T in enumerate(frames)])).pin_memory()
    rng = np.random.RandomState(100 + rank)
    phase_h = torch.from_numpy(np.angle(np.exp(2j * np.pi * rng.rand(total, n_bins))).astype(np.float32)).pin_memory()
    n_samples = (total - len(frames)) * HOP
    wave_h = torch.empty(n_samples, dtype=torch.float32).pin_memory()
    logmel_d = logmel_h.to(dev)
    phase_d = phase_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timing (value) ----------------------------------------------------------
    for _ in range(args.warmup):
        voc.synthesize_flat(logmel_d, frames, phase_d)
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("BENCH_NO_CLOCKS")) else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for step in range(args.steps):
        wave_d = voc.synthesize_flat(logmel_d, frames, phase_d)
    ev1.record()
    # Clocks / throttle reasons are sampled NOW: everything of the timed region is enqueued, the GPU is still
    # working through the queue (the host runs up to ~15 steps ahead), and no launch can be delayed by a slow
    # NVML call (on some boxes one call takes tens of ms and, issued between steps, drained the launch queue).
    if sampler:
        while not ev1.query() and len(sampler.samples) < 8:
            sampler.sample()
            time.sleep(0.005)
        sampler.under_load = len(sampler.samples)
        if not sampler.samples:
            sampler.sample()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    assert torch.isfinite(wave_d).all()
    # per-launch device time of the fused iteration kernel: CUDA events recorded by the library on the
    # launching stream around every pass, over the same steps again (kept out of `value`'s region so the
    # extra event records cannot perturb it)
    plan.set_pass_timing(True)
    per_step = []
    barrier()
    for _ in range(args.steps):
        voc.synthesize_flat(logmel_d, frames, phase_d)
        per_step.append(plan.pass_times_ms())
    plan.set_pass_timing(False)
    last_pass_ms = np.mean(np.stack(per_step), axis=0)
    # one launch per iteration: [initial inverse, iteration 1, ..., iteration n]; persistent mode (all iterations in one
    # launch, the default when every strip gets a resident warp): [initial inverse, the persistent launch]
    persistent = len(last_pass_ms) == 2 and N_ITER > 1
    iters_per_launch = N_ITER if persistent else 1
    iter_ms = float(np.mean(last_pass_ms[1:])) if len(last_pass_ms) > 1 else float("nan")  # per LAUNCH

    # ---- end to end through the public API, host buffers in and out (e2e) -------------------------
    # Every step: H2D of that step's log-mel from pinned memory, synthesis, D2H of the waveforms.  The
    # initial phase is drawn inside the timed call in both arms: the reference draws it with numpy on the
    # host (vocoder.py:103), the library draws the same distribution on the device (phase_fm=None).  The
    # variant that uploads a host-drawn phase (4.1 KB per frame, what the parity tests use) is reported too.
    # Steps are fed back to back like generate_waveform.py feeds batches: synthesize_host() downloads on a copy stream,
    # so step i's D2H overlaps step i+1's H2D and kernels; two host output buffers alternate, and the timed region ends
    # only when every download has landed (barrier() synchronises the whole device).
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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            e2e_step(host_phase)
        e1.record()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t_host0) / args.steps
        return max_over_ranks(max(e0.elapsed_time(e1) / args.steps, wall_ms))

    e2e_host_ms = time_e2e(True)
    e2e_ms = time_e2e(False)
    assert torch.isfinite(wave_h).all() and torch.isfinite(wave_h2).all()
    e2e_count[0] = 0
    e2e_step(True)  # leave the host-phase result in wave_h for the parity spot check below
    torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    algo_bytes_iter = ALGO_BYTES_PER_FRAME_ITER * total * iters_per_launch  # per launch
    achieved = algo_bytes_iter / (iter_ms * 1e-3) / 1e9
    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ---------------------
    from oracle import griffin_lim as ogl
    basis = voc.inv_mel_transform.basis.cpu().numpy()
    sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, 6)]
    cpu_audio, cpu_s = cpu_port_time(sample, 4321, N_ITER, basis) if world == 1 else (0.0, 0.0)  # N = 1 only
    # parity spot check on the way (checker only): first utterance of the batch vs the oracle
    T0 = frames[0]
    ref0 = ogl.vocoder_forward(logmel_h[:T0].numpy(), np.ascontiguousarray(phase_h[:T0].numpy().T), N_ITER, basis=basis)
    parity = ogl.rel_l2(wave_h[: (T0 - 1) * HOP].numpy(), ref0)

    launches = plan.gl_launch_count(N_ITER, True) * args.steps
    out = {
        "metric": "griffin_lim_audio_seconds_per_second", "value": world * audio_s / (ms_step * 1e-3),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_gpu": N_UTTS, "frames_per_gpu": total,
                   "audio_seconds_per_gpu": audio_s, "n_iter": N_ITER, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP,
                   "win": WIN, "l2_policy": "per-step working set (magnitudes + phase + waveform buffers, "
                   f"{(total * (684 + 1025) * 4 + 4 * n_samples * 4) / 1e6:.0f} MB) exceeds the 126 MB L2"},
        "e2e": {"value": world * audio_s / (e2e_ms * 1e-3), "unit": "audio-s/s",
                "h2d_bytes_per_step": int(logmel_h.numel() * 4),
                "d2h_bytes_per_step": int(wave_h.numel() * 4), "ms_per_step": e2e_ms,
                "api": "GriffinLimVocoder.synthesize_host(pinned log-mel in, pinned waveforms out; initial phase drawn on "
                       "the device; H2D and D2H run on copy streams and overlap the kernels of the neighbouring steps)",
                "with_host_drawn_phase": {"value": world * audio_s / (e2e_host_ms * 1e-3), "ms_per_step": e2e_host_ms,
                                          "h2d_bytes_per_step": int(logmel_h.numel() * 4 + phase_h.numel() * 4)}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": ("k_gl_pass<19,false,true,true,PERSIST> (all %d iterations in one cooperative launch: " % N_ITER
                                                 if persistent else "k_gl_pass<19,false,true,true> (") +
                                                "fused iSTFT+OLA+normalise+STFT+magnitude re-imposition)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     # ncu capture of ONE iteration (profiles/r01_glpass_current_ncu_summary.txt) x iterations per launch
                     "traffic": (measured_traffic_bytes() or 0) * iters_per_launch or None, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes_iter,
                     "launch_ms": iter_ms, "launches_per_step": N_ITER // iters_per_launch,
                     "iterations_per_launch": iters_per_launch, "ms_per_iteration": iter_ms / iters_per_launch,
                     "share_of_step": float(np.sum(last_pass_ms[1:]) / ms_step) if len(last_pass_ms) > 1 else None,
                     "first_pass_ms": float(last_pass_ms[0])},
        "cpu_baseline": None if world > 1 else {
            "value": cpu_audio / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port",
            "sample": f"6 length-stratified utterances of the batch ({sum(sample)} frames), {N_ITER} iters, "
                      f"numpy FFT oracle, {cpu_s:.1f} s"},
        "clocks": clocks,
        "parity_rel_l2_vs_oracle": parity,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
