"""Plan handles: device-resident constants of the CUDA library, cached per configuration.

A plan corresponds to what ``GriffinLimVocoder.__init__`` / ``TTSSpectrogram.__init__`` /
``TTSMelScale.__init__`` precompute in the reference (vocoder.py:114-134,
audio_utils.py:246-257,275-282): window, transform tables, mel / pseudo-inverse mel matrices.
"""
import ctypes
import hashlib
import weakref

import numpy as np
import torch

from . import _lib


def _as_f32(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _np_ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def require_cuda(device=None):
    if torch.cuda._is_in_bad_fork():
        # the reference runs its numpy transforms in FORKED DataLoader workers; a forked child cannot use the CUDA context
        # of its parent.  Fail with the two supported set-ups instead of CUDA's "Cannot re-initialize CUDA in forked
        # subprocess": (1) keep the dataset transform-free and run CompositeAudioFeatureTransform.apply_cuda_from_host on
        # the collated batch in the main process (one launch per transform per batch: the fast way), or (2) start the
        # workers with DataLoader(..., multiprocessing_context="spawn").  INTEGRATION.md, "DataLoader workers".
        raise RuntimeError("s2st_b200 was called in a forked worker process after CUDA was initialised in the parent. Apply "
                           "the feature transforms after collation (CompositeAudioFeatureTransform.apply_cuda_from_host) "
                           "or create the DataLoader with multiprocessing_context='spawn'; see INTEGRATION.md")
    if not torch.cuda.is_available():
        raise RuntimeError("s2st_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None or torch.device(device).type != "cuda":
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    return torch.device("cuda", torch.cuda.current_device() if device.index is None else device.index)


def upload_small(a: np.ndarray, device) -> torch.Tensor:
    """Host array -> device tensor without stalling the host: the array is staged in pinned memory (torch's caching
    host allocator recycles the block once the copy has run) and copied with a truly asynchronous H2D.
    ``torch.from_numpy(a).to(device)`` from pageable memory makes the driver synchronise the stream before the copy, i.e.
    every call would wait for all GPU work enqueued before it -- which serialises back-to-back synthesis calls."""
    a = np.ascontiguousarray(a)
    staged = torch.empty(a.shape, dtype=torch.from_numpy(a.reshape(-1)[:0]).dtype, pin_memory=True)
    staged.numpy()[...] = a
    return staged.to(device, non_blocking=True)


class StftPlan:
    """Owns one ``s2st_plan``."""

    def __init__(self, device, n_fft, win_length, hop_length, n_mels, window, inv_mel=None, mel=None):
        lib = _lib.load()
        self.device = require_cuda(device)
        self.n_fft, self.win_length, self.hop_length, self.n_mels = int(n_fft), int(win_length), int(hop_length), int(n_mels)
        window = _as_f32(window)
        inv_mel = _as_f32(inv_mel)
        mel = _as_f32(mel)
        assert window.shape == (self.win_length,)
        n_bins = self.n_fft // 2 + 1
        if inv_mel is not None:
            assert inv_mel.shape == (n_bins, self.n_mels), inv_mel.shape
        if mel is not None:
            assert mel.shape == (self.n_mels, n_bins), mel.shape
        handle = ctypes.c_void_p()
        rc = lib.s2st_plan_create(ctypes.byref(handle), self.device.index, self.n_fft, self.win_length,
                                  self.hop_length, self.n_mels, _np_ptr(window), _np_ptr(inv_mel), _np_ptr(mel))
        _lib.check(rc, "s2st_plan_create")
        self.handle = handle
        kb = ctypes.c_int()
        _lib.check(lib.s2st_plan_active_bins(self.handle, ctypes.byref(kb)), "s2st_plan_active_bins")
        self.active_bins = kb.value

    def workspace(self, n_utts, total_frames):
        """Scratch for ONE synthesis call (strip tables, magnitudes, two rotating waveform buffers), allocated per call
        from torch's caching allocator on the caller's current stream.  The allocator hands a freed block back only to
        work of the same stream, so calls on different streams / threads never share scratch, and a repeated call of
        the same size gets its block back without a cudaMalloc."""
        lib = _lib.load()
        need = ctypes.c_size_t()
        _lib.check(lib.s2st_gl_workspace_bytes(self.handle, n_utts, total_frames, ctypes.byref(need)),
                   "s2st_gl_workspace_bytes")
        with torch.cuda.device(self.device):
            return torch.empty(need.value, dtype=torch.uint8, device=self.device)

    def set_option(self, option, value):
        """s2st_plan_set_option (keys: _lib.OPT_*)."""
        _lib.check(_lib.load().s2st_plan_set_option(self.handle, int(option), int(value)), "s2st_plan_set_option")

    def gl_launch_count(self, n_iter, from_logmel=True):
        n = ctypes.c_int()
        _lib.check(_lib.load().s2st_gl_launch_count(self.handle, n_iter, int(from_logmel), ctypes.byref(n)),
                   "s2st_gl_launch_count")
        return n.value

    def set_strip_frames(self, frames):
        """0 = automatic strip length per call; > 0 pins it (bitwise batch-invariant results)."""
        _lib.check(_lib.load().s2st_plan_set_strip_frames(self.handle, int(frames)), "s2st_plan_set_strip_frames")

    def set_pass_timing(self, enabled):
        _lib.check(_lib.load().s2st_plan_set_pass_timing(self.handle, int(enabled)), "s2st_plan_set_pass_timing")

    def pass_times_ms(self):
        """Device time of every Griffin-Lim pass of the last call (synchronises on its last event)."""
        buf = np.zeros(1026, np.float32)
        n = ctypes.c_int()
        _lib.check(_lib.load().s2st_plan_get_pass_times(self.handle, ctypes.c_void_p(buf.ctypes.data), buf.size,
                                                        ctypes.byref(n)), "s2st_plan_get_pass_times")
        return buf[: n.value].copy()

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().s2st_plan_destroy(h)
            except Exception:
                pass


class FbankPlan:
    """Owns one ``s2st_fbank_plan`` (Kaldi fbank constants for one sample rate)."""

    def __init__(self, device, sample_rate, n_bins=80):
        lib = _lib.load()
        self.device = require_cuda(device)
        self.sample_rate, self.n_bins = int(sample_rate), int(n_bins)
        handle = ctypes.c_void_p()
        _lib.check(lib.s2st_fbank_plan_create(ctypes.byref(handle), self.device.index, self.sample_rate, self.n_bins),
                   "s2st_fbank_plan_create")
        self.handle = handle
        w, s, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.s2st_fbank_frame_params(self.handle, ctypes.byref(w), ctypes.byref(s), ctypes.byref(p)),
                   "s2st_fbank_frame_params")
        self.win, self.shift, self.padded = w.value, s.value, p.value

    def set_option(self, option, value):
        _lib.check(_lib.load().s2st_fbank_plan_set_option(self.handle, int(option), int(value)),
                   "s2st_fbank_plan_set_option")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().s2st_fbank_plan_destroy(h)
            except Exception:
                pass


_cache = {}


_digest_memo = {}


def _digest(a):
    """Content digest of a constant (window, mel basis, ...).  For tensors the result is memoised on the tensor OBJECT
    (weak reference + version counter): hashing needs the data on the host, and for a CUDA-resident module buffer
    that is a device-to-host copy -- a full stream synchronisation -- which must not happen on every call."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        hit = _digest_memo.get(id(a))
        if hit is not None and hit[0]() is a and hit[1] == a._version:
            return hit[2]
        d = hashlib.sha1(_as_f32(a).tobytes()).hexdigest()
        if len(_digest_memo) > 256:
            _digest_memo.clear()
        try:
            _digest_memo[id(a)] = (weakref.ref(a), a._version, d)
        except TypeError:
            pass
        return d
    return hashlib.sha1(_as_f32(a).tobytes()).hexdigest()


def get_stft_plan(device, n_fft, win_length, hop_length, n_mels, window, inv_mel=None, mel=None):
    device = require_cuda(device)
    key = ("stft", device.index, n_fft, win_length, hop_length, n_mels, _digest(window), _digest(inv_mel), _digest(mel))
    if key not in _cache:
        _cache[key] = StftPlan(device, n_fft, win_length, hop_length, n_mels, window, inv_mel, mel)
    return _cache[key]


def get_fbank_plan(device, sample_rate, n_bins=80):
    device = require_cuda(device)
    key = ("fbank", device.index, int(sample_rate), int(n_bins))
    if key not in _cache:
        _cache[key] = FbankPlan(device, sample_rate, n_bins)
    return _cache[key]
