"""Drop-in for ``fairseq/models/text_to_speech/vocoder.py`` (Griffin-Lim part), B200-native.

Same classes, constructor arguments, ``forward`` contract and error behaviour as the reference
(``PseudoInverseMelScale`` :24-46, ``GriffinLim`` :49-110, ``GriffinLimVocoder`` :113-158,
``get_vocoder`` :191-197).  All arithmetic runs in the sm_100a CUDA library through the C ABI
(``include/s2st_b200.h``); there is no CPU path.  On top of the per-utterance ``forward`` the
vocoder has a ragged batched entry, ``synthesize_batch``, which is the data-parallel hot path:
one kernel launch per Griffin-Lim iteration for the whole batch.
"""
import logging
from typing import List, Optional, Sequence

import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib
from .audio_utils import TTSSpectrogram, get_mel_filters
from .plans import get_stft_plan, require_cuda, upload_small

logger = logging.getLogger(__name__)


def draw_initial_phase(shape) -> np.ndarray:
    """The reference's initial phase (vocoder.py:103): consumes the GLOBAL numpy RNG in float64."""
    return np.angle(np.exp(2j * np.pi * np.random.rand(*shape))).astype(np.float32)


def _draw_initial_phase_host_rng(shape, dev) -> torch.Tensor:
    """The draw with the RNG on the host: ``np.random.rand(*shape)`` consumes numpy's GLOBAL generator exactly like
    vocoder.py:103, the float64 uniforms are uploaded from pinned memory, and ``angle(exp(2j pi u))`` becomes the closed
    form theta / theta - 2 pi in float64 on the device, written frame-major.  Used when numpy's global generator is not
    the legacy MT19937 (it always is unless somebody replaced it)."""
    u = np.random.rand(*shape)
    B = shape[0] if len(shape) == 3 else 1
    F, T = shape[-2], shape[-1]
    staged = torch.empty(u.shape, dtype=torch.float64, pin_memory=True)
    staged.numpy()[...] = u
    phase = torch.empty(B * T, F, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        u_d = staged.to(dev, non_blocking=True)
        rc = _lib.load().s2st_phase_from_uniform(B, F, T, _lib.ptr(u_d), _lib.ptr(phase), _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_phase_from_uniform")
    return phase


class _NumpyStreamOnDevice:
    """Per-device resources of ``draw_initial_phase_device``: a side stream and the small buffers that only that stream
    touches (key up / down, the raw output words), reused from call to call -- the stream orders their uses, and
    ``finish()`` of a call has synchronised with its last kernel before the next call writes the pinned buffers."""
    _by_device = {}

    @classmethod
    def get(cls, dev):
        if dev.index not in cls._by_device:
            cls._by_device[dev.index] = cls(dev)
        return cls._by_device[dev.index]

    def __init__(self, dev):
        self.stream = torch.cuda.Stream(device=dev)
        self.key_stage = torch.empty(624, dtype=torch.int32, pin_memory=True)
        self.key_host = torch.empty(624, dtype=torch.int32, pin_memory=True)
        with torch.cuda.stream(self.stream):
            self.key_d = torch.empty(624, dtype=torch.int32, device=dev)
            self.key_out = torch.empty(624, dtype=torch.int32, device=dev)
            self.words = torch.empty(0, dtype=torch.int32, device=dev)
        self.done = torch.cuda.Event()
        self.pending = None  # finish() of a draw whose state has not been handed back to numpy yet

    def words_buffer(self, n_words, dev):
        if self.words.numel() < n_words:
            with torch.cuda.stream(self.stream):
                self.words = torch.empty(n_words, dtype=torch.int32, device=dev)
        return self.words


def draw_initial_phase_device(shape, dev):
    """The reference's initial phase (vocoder.py:103: ``angle(exp(2j pi np.random.rand(*shape)))``) for shape = (B,) F, T,
    frame-major [B * T, F] float32 on ``dev`` -- with numpy's GLOBAL generator continued on the device.

    ``np.random.seed(s)`` before ``forward`` must reproduce the reference, so the draw has to be numpy's stream; but
    512 500 doubles for a 500-frame utterance cost ~1.2 ms on the host plus a 4 MB upload -- more than the synthesis.
    Instead the generator state (MT19937 key + position, 2.5 KB) goes to the device, ``s2st_phase_from_mt19937``
    produces exactly the outputs ``rand`` would consume and the state it would leave, and that state is put back.

    Returns ``(phase, finish)``: call ``finish()`` once the rest of the work has been enqueued -- it waits for the
    generator kernel (not for the synthesis) and hands the advanced state to ``np.random.set_state``."""
    for res in _NumpyStreamOnDevice._by_device.values():
        if res.pending is not None:  # a caller that drew without finishing: numpy's state is still the old one
            res.pending()
    st = np.random.get_state()
    n = int(np.prod(shape))
    if st[0] != "MT19937" or n == 0:
        return _draw_initial_phase_host_rng(shape, dev), (lambda: None)
    B = shape[0] if len(shape) == 3 else 1
    F, T = shape[-2], shape[-1]
    pos = int(st[2])
    # The generator runs on a side stream: it is one thread block walking a sequential recurrence, so in a loop of
    # forward() calls it overlaps the previous call's synthesis kernels instead of queueing behind them (and finish()
    # waits for the generator only).  The phase is allocated on that stream; the caller's stream waits for it.
    res = _NumpyStreamOnDevice.get(dev)
    main = torch.cuda.current_stream(dev)
    side = res.stream
    words = res.words_buffer(2 * n, dev)
    res.key_stage.numpy()[...] = np.asarray(st[1], np.uint32).view(np.int32)
    with torch.cuda.stream(side):
        res.key_d.copy_(res.key_stage, non_blocking=True)
        phase = torch.empty(B * T, F, dtype=torch.float32, device=dev)
        rc = _lib.load().s2st_phase_from_mt19937(B, F, T, _lib.ptr(res.key_d), pos, _lib.ptr(words), _lib.ptr(phase),
                                                 _lib.ptr(res.key_out), ctypes.c_void_p(side.cuda_stream))
        _lib.check(rc, "s2st_phase_from_mt19937")
        res.key_host.copy_(res.key_out, non_blocking=True)
        res.done.record(side)
    main.wait_stream(side)
    phase.record_stream(main)
    end = pos + 2 * n
    new_pos = end - 624 * ((end - 1) // 624)

    def finish():
        if res.pending is finish:
            res.pending = None
            res.done.synchronize()
            np.random.set_state((st[0], res.key_host.numpy().view(np.uint32).copy(), new_pos, st[3], st[4]))

    res.pending = finish
    return phase, finish


class PseudoInverseMelScale(torch.nn.Module):
    def __init__(self, n_stft, n_mels, sample_rate, f_min, f_max) -> None:
        super().__init__()
        self.n_mels, self.n_stft = n_mels, n_stft
        self.n_fft = (n_stft - 1) * 2
        _check_synthesis_geometry(self.n_fft)
        mel = get_mel_filters(sample_rate, self.n_fft, n_mels, f_min, f_max)
        self.register_buffer("basis", torch.pinverse(mel))  # F x F_mel, as the reference builds it

    def _plan(self, device):
        return get_stft_plan(device, self.n_fft, self.n_fft, self.n_fft // 4, self.n_mels,
                             torch.ones(self.n_fft), inv_mel=self.basis)

    def forward(self, melspec: torch.Tensor) -> torch.Tensor:
        """melspec [..., F_mel, T] (linear mel magnitudes) -> [..., F, T], clamped at 0."""
        shape = melspec.shape
        n_mels, time = shape[-2], shape[-1]
        assert self.n_mels == n_mels, (self.n_mels, n_mels)
        dev = require_cuda(melspec.device)
        x = melspec.detach().to(dev, torch.float32).reshape(-1, n_mels, time).transpose(1, 2).contiguous()
        rows = x.shape[0] * time
        out = torch.empty(rows, self.n_stft, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_inverse_mel(self._plan(dev).handle, rows, _lib.ptr(x), 0, _lib.ptr(out),
                                              _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_inverse_mel")
        out = out.view(-1, time, self.n_stft).transpose(1, 2)
        return out.reshape(shape[:-2] + (self.n_stft, time)).to(melspec.device, melspec.dtype)


class GriffinLim(torch.nn.Module):
    """Same constructor / ``forward`` / ``inverse`` / ``get_window_sum_square`` as the reference (vocoder.py:49-110).

    One deliberate deviation, visible only with a ``window_fn`` other than ``torch.hann_window`` (the recipe and every
    config of the reference use hann): the reference builds its ANALYSIS transform and its window-sum-square with the
    default hann window whatever ``window_fn`` is (vocoder.py:54, :90) and applies ``window_fn`` only to the synthesis
    basis (:62) -- an inconsistency of that code, not a design.  Here ``window_fn`` is used for analysis, synthesis and
    the normalisation alike, so a non-hann window gives a consistent (perfect-reconstruction) STFT pair but not the
    reference's numbers.  The module's constant is registered as ``window`` ([win_length]); the reference's dense
    ``basis`` buffer ([2050, 1, 2048]) does not exist because the transforms are FFTs (state dicts carry no learned
    parameters either way)."""

    def __init__(self, n_fft: int, win_length: int, hop_length: int, n_iter: int, window_fn=torch.hann_window):
        super().__init__()
        _check_synthesis_geometry(n_fft)
        self.transform = TTSSpectrogram(n_fft, win_length, hop_length, window_fn=window_fn, return_phase=True)
        self.register_buffer("window", window_fn(win_length).float())
        self.n_fft, self.win_length, self.hop_length, self.n_iter = n_fft, win_length, hop_length, n_iter
        self.tiny = 1.1754944e-38

    @classmethod
    def get_window_sum_square(cls, n_frames, hop_length, win_length, n_fft, window_fn=torch.hann_window) -> torch.Tensor:
        """sum_t w^2[n - t*hop], length n_fft + hop*(n_frames-1) (host helper of the C library)."""
        import ctypes
        w = np.ascontiguousarray(window_fn(win_length).float().numpy())
        out = np.empty(n_fft + hop_length * (n_frames - 1), np.float32)
        rc = _lib.load().s2st_window_sum_square(n_frames, hop_length, win_length, n_fft,
                                                ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data))
        _lib.check(rc, "s2st_window_sum_square")
        return torch.from_numpy(out)

    def _plan(self, device):
        return get_stft_plan(device, self.n_fft, self.win_length, self.hop_length, 1, self.window)

    def _run(self, mag_fm, phase_fm, n_utts, frames_per_utt, n_iter, dev):
        """mag_fm / phase_fm: frame-major [B*T, F] float32 on dev -> [B, L]."""
        plan = self._plan(dev)
        total = n_utts * frames_per_utt
        fo_h = np.arange(0, total + 1, frames_per_utt, dtype=np.int32)
        fo = upload_small(fo_h, dev)
        L = (frames_per_utt - 1) * self.hop_length
        wave = torch.empty(n_utts, L, dtype=torch.float32, device=dev)
        if L == 0:
            return wave
        ws = plan.workspace(n_utts, total)
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_gl_synthesize(plan.handle, n_utts, total, _lib.ptr(fo), fo_h.ctypes.data, None, _lib.ptr(mag_fm),
                                                _lib.ptr(phase_fm), 0, n_iter, _lib.ptr(wave), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_gl_synthesize")
        return wave

    def inverse(self, magnitude: torch.Tensor, phase) -> torch.Tensor:
        """magnitude, phase [B, F, T] -> [B, 1, (T-1)*hop] (iSTFT with window-sum-square normalisation)."""
        dev = require_cuda(magnitude.device)
        B, F, T = magnitude.shape
        mag = magnitude.detach().to(dev, torch.float32).transpose(1, 2).reshape(B * T, F).contiguous()
        ph = torch.as_tensor(phase).detach().to(dev, torch.float32).transpose(1, 2).reshape(B * T, F).contiguous()
        wave = self._run(mag, ph, B, T, 0, dev)
        return wave.unsqueeze(1).to(magnitude.device, magnitude.dtype)

    def forward(self, specgram: torch.Tensor) -> torch.Tensor:
        """specgram [F, T] or [B, F, T] linear magnitudes -> waveform(s); random initial phase from the
        global numpy RNG exactly like the reference."""
        dev = require_cuda(specgram.device)
        ph, finish_rng = draw_initial_phase_device(tuple(specgram.shape), dev)  # consumes numpy's global RNG first, like the reference
        try:
            spec = specgram.detach().reshape(-1, specgram.shape[-2], specgram.shape[-1])
            B, F, T = spec.shape
            _check_length(T, self.hop_length, self.n_fft, self.n_iter)
            mag = spec.to(dev, torch.float32).transpose(1, 2).reshape(B * T, F).contiguous()
            wave = self._run(mag, ph, B, T, self.n_iter, dev)
        finally:
            finish_rng()
        return wave.squeeze(0).to(specgram.device, specgram.dtype)


def _check_synthesis_geometry(n_fft):
    """Griffin-Lim synthesis (inverse-mel + fused STFT / iSTFT iterations) is built for the recipe's 2048-point
    transform; the analysis side (TTSSpectrogram, TTSMelScale, log-mel extraction) takes any power of two."""
    if n_fft != 2048:
        raise ValueError(f"Griffin-Lim synthesis supports n_fft = 2048 only (got n_fft = {n_fft}); the STFT / log-mel / mel "
                         "projection entry points accept any power-of-two n_fft in [64, 4096]")


def _check_length(T, hop, n_fft, n_iter):
    # the reference's STFT reflect-pads n_fft//2 and F.pad raises when padding >= length
    L = (T - 1) * hop
    if n_iter > 0 and L <= n_fft // 2:
        raise RuntimeError(
            f"Argument #4: Padding size should be less than the corresponding input dimension, but got: padding "
            f"({n_fft // 2}, {n_fft // 2}) at dimension 2 of input [1, 1, {L}] ({T} frames are too few for "
            f"n_fft {n_fft}, hop {hop})")


class GriffinLimVocoder(nn.Module):
    def __init__(self, sample_rate, win_size, hop_size, n_fft, n_mels, f_min, f_max, window_fn,
                 spec_bwd_max_iter=32, fp16=False):
        super().__init__()
        self.inv_mel_transform = PseudoInverseMelScale(n_stft=n_fft // 2 + 1, n_mels=n_mels, sample_rate=sample_rate,
                                                       f_min=f_min, f_max=f_max)
        self.gl_transform = GriffinLim(n_fft=n_fft, win_length=win_size, hop_length=hop_size, window_fn=window_fn,
                                       n_iter=spec_bwd_max_iter)
        self.sample_rate, self.n_mels = sample_rate, n_mels
        # the kernels compute in fp32; with fp16=True inputs / outputs are half like the reference's
        # (the S2ST recipe never passes --fp16 at synthesis, run_baseline.sh:143-150)
        self.fp16 = fp16
        self.float()

    # -- plans ----------------------------------------------------------------------------------
    def _plan(self, device):
        # The plan registry keys on a digest of the constants, which needs them on the host: for CUDA-resident buffers
        # that is a device-to-host copy, i.e. a full stream synchronisation.  Doing it per call serialised back-to-back
        # synthesis calls (the host could not enqueue step i+1 before step i had finished), so the lookup is memoised
        # on the identity / version of the buffers and only repeated after .cuda() / .half() / in-place edits.
        g = self.gl_transform
        w, b = g.window, self.inv_mel_transform.basis
        key = (device.index, g.n_fft, g.win_length, g.hop_length, self.n_mels, w.data_ptr(), w._version, w.dtype,
               b.data_ptr(), b._version, b.dtype)
        memo = self.__dict__.setdefault("_plan_memo", {})
        if memo.get("key") != key:
            memo["key"] = key
            memo["plan"] = get_stft_plan(device, g.n_fft, g.win_length, g.hop_length, self.n_mels, w.float(),
                                         inv_mel=b.float())
        return memo["plan"]

    def _device(self, x):
        if x.device.type == "cuda":
            return require_cuda(x.device)
        b = self.inv_mel_transform.basis
        return require_cuda(b.device if b.device.type == "cuda" else None)

    # -- the reference API ----------------------------------------------------------------------
    def forward(self, x):
        """x: (B x) T x n_mels denormalised log-mel -> (B x) (T-1)*hop waveform on x's device / dtype."""
        if self.training:  # (the reference calls self.eval() on every forward; the recursive call costs 30 us of a 0.5 ms call)
            self.eval()
        g = self.gl_transform
        batched = x.dim() == 3
        feats = x.detach()
        B = feats.shape[0] if batched else 1
        T = feats.shape[-2]
        assert feats.shape[-1] == self.n_mels, (self.n_mels, feats.shape[-1])
        # initial phase: one draw of shape (B x) F x T from numpy's global RNG (vocoder.py:103)
        shape = ((B,) if batched else ()) + (g.n_fft // 2 + 1, T)
        dev = self._device(feats)
        phase_fm, finish_rng = draw_initial_phase_device(shape, dev)
        try:
            _check_length(T, g.hop_length, g.n_fft, g.n_iter)
            waves = self._synthesize_flat(feats.reshape(B * T, self.n_mels).to(dev, torch.float32).contiguous(),
                                          [T] * B, phase_fm, g.n_iter, dev)
        finally:
            finish_rng()  # numpy's generator is where the reference's draw would have left it
        out = waves.view(B, -1)
        out = out.squeeze(0) if (not batched or B == 1) else out
        return out.to(x.device, x.dtype)

    @classmethod
    def from_data_cfg(cls, args, data_cfg):
        feat_cfg = data_cfg.config["features"]
        window_fn = getattr(torch, feat_cfg["window_fn"] + "_window")
        return cls(sample_rate=feat_cfg["sample_rate"],
                   win_size=int(feat_cfg["win_len_t"] * feat_cfg["sample_rate"]),
                   hop_size=int(feat_cfg["hop_len_t"] * feat_cfg["sample_rate"]),
                   n_fft=feat_cfg["n_fft"], n_mels=feat_cfg["n_mels"],
                   f_min=feat_cfg["f_min"], f_max=feat_cfg["f_max"],
                   window_fn=window_fn, spec_bwd_max_iter=args.spec_bwd_max_iter, fp16=args.fp16)

    # -- the data-parallel entry ----------------------------------------------------------------
    def _synthesize_flat(self, logmel_flat, frames: Sequence[int], phase_fm, n_iter, dev, seed=0):
        plan = self._plan(dev)
        n_utts, total = len(frames), int(sum(frames))
        fo = np.zeros(n_utts + 1, np.int32)
        fo[1:] = np.cumsum(frames)
        fo_d = upload_small(fo, dev)
        n_samples = (total - n_utts) * self.gl_transform.hop_length
        wave = torch.empty(max(n_samples, 0), dtype=torch.float32, device=dev)
        if n_samples <= 0:
            return wave
        ws = plan.workspace(n_utts, total)
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_gl_synthesize(plan.handle, n_utts, total, _lib.ptr(fo_d), fo.ctypes.data, _lib.ptr(logmel_flat), None,
                                                _lib.ptr(phase_fm), int(seed) & 0xFFFFFFFFFFFFFFFF, n_iter, _lib.ptr(wave), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_gl_synthesize")
        return wave

    def synthesize_batch(self, feats: List[torch.Tensor], init_phase: Optional[List] = None,
                         n_iter: Optional[int] = None, device=None) -> List[torch.Tensor]:
        """Ragged batch: feats[i] is [T_i, n_mels] denormalised log-mel -> list of [(T_i-1)*hop] CUDA tensors.

        init_phase[i] (optional) is the reference-layout [F, T_i] initial phase (numpy or tensor); when
        omitted, phases are drawn from numpy's global RNG utterance by utterance, i.e. exactly what
        calling ``forward`` on each utterance in order would consume.
        """
        g = self.gl_transform
        n_iter = g.n_iter if n_iter is None else n_iter
        F = g.n_fft // 2 + 1
        frames = [int(f.shape[0]) for f in feats]
        for T in frames:
            _check_length(T, g.hop_length, g.n_fft, n_iter)
        dev = require_cuda(device) if device is not None else self._device(feats[0])
        if init_phase is None:
            init_phase = [draw_initial_phase((F, T)) for T in frames]
        ph = []
        for p, T in zip(init_phase, frames):
            if isinstance(p, torch.Tensor):
                assert p.shape == (F, T)
                ph.append(p.to(dev, torch.float32).t())
            else:
                p = np.asarray(p, np.float32)
                assert p.shape == (F, T)
                ph.append(torch.from_numpy(np.ascontiguousarray(p.T)).to(dev, non_blocking=True))
        phase_fm = torch.cat(ph).contiguous()
        flat = torch.cat([f.detach().to(dev, torch.float32) for f in feats]).contiguous()
        wave = self._synthesize_flat(flat, frames, phase_fm, n_iter, dev)
        lens = [(T - 1) * g.hop_length for T in frames]
        return list(torch.split(wave, lens))

    def synthesize_flat(self, logmel_flat: torch.Tensor, frames: Sequence[int], phase_fm: Optional[torch.Tensor],
                        n_iter: Optional[int] = None, seed: int = 0) -> torch.Tensor:
        """Lowest-overhead entry: everything already resident and frame-major on one CUDA device:
        logmel_flat [sum T, n_mels], phase_fm [sum T, F] -> concatenated waveforms [sum (T_i-1)*hop].
        phase_fm=None draws the initial phase on the device (U[-pi, pi), counter-based generator keyed by
        ``seed``): same distribution as the reference's numpy draw, no 4 KB/frame upload."""
        n_iter = self.gl_transform.n_iter if n_iter is None else n_iter
        dev = require_cuda(logmel_flat.device)
        assert logmel_flat.is_cuda and logmel_flat.dtype == torch.float32
        if phase_fm is not None:
            assert phase_fm.is_cuda and phase_fm.dtype == torch.float32
            phase_fm = phase_fm.contiguous()
        for T in frames:
            _check_length(T, self.gl_transform.hop_length, self.gl_transform.n_fft, n_iter)
        return self._synthesize_flat(logmel_flat.contiguous(), frames, phase_fm, n_iter, dev, seed)


    def synthesize_host(self, logmel_host: torch.Tensor, frames: Sequence[int], wave_host: torch.Tensor,
                        device=None, phase_host: Optional[torch.Tensor] = None, n_iter: Optional[int] = None,
                        seed: int = 0) -> torch.cuda.Event:
        """Host buffers in, host buffer out, pipelined: uploads ``logmel_host`` [sum T, n_mels] (pinned), synthesises
        on ``device`` and downloads the concatenated waveforms into ``wave_host`` (pinned, sum (T_i-1)*hop floats).

        Upload and download run on private copy streams, so they overlap the kernels of the neighbouring calls (this is
        how generate_waveform.py would feed batch after batch).  Returns the CUDA event that marks ``wave_host`` complete;
        callers that reuse ``wave_host`` must synchronise on it (or on the device) first."""
        dev = require_cuda(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        assert logmel_host.dtype == torch.float32 and wave_host.dtype == torch.float32
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = {}
        if dev.index not in self._copy_streams:
            self._copy_streams[dev.index] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        cs, up = self._copy_streams[dev.index]  # download / upload streams
        main = torch.cuda.current_stream(dev)
        # uploads run on their own stream too: the inputs of this call cross PCIe while the kernels of the previous
        # call are still running (with a host-drawn phase that is 4.1 KB per frame, 257 MB for the config-2 batch)
        with torch.cuda.stream(up):
            lm = logmel_host.to(dev, non_blocking=True)
            ph = phase_host.to(dev, non_blocking=True) if phase_host is not None else None
        main.wait_stream(up)
        lm.record_stream(main)
        if ph is not None:
            ph.record_stream(main)
        wave = self.synthesize_flat(lm, frames, ph, n_iter=n_iter, seed=seed)
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            wave_host[: wave.numel()].copy_(wave, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cs)
        wave.record_stream(cs)  # keep the device buffer alive until the download has read it
        return done


def get_vocoder(args, data_cfg):
    if args.vocoder == "griffin_lim":
        return GriffinLimVocoder.from_data_cfg(args, data_cfg)
    elif args.vocoder == "hifigan":
        raise NotImplementedError("the HiFi-GAN neural vocoder is outside this package's scope (Griffin-Lim hot path only)")
    else:
        raise ValueError("Unknown vocoder")
