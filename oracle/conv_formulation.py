"""Second CPU restatement of the Griffin-Lim path, in the reference's OWN formulation (torch, float32): the STFT is a
strided convolution with a dense windowed DFT basis and the inverse a transposed convolution with the pseudo-inverse
basis, exactly the operations ``fairseq/models/text_to_speech/vocoder.py:50-110`` and
``fairseq/data/audio/audio_utils.py:246-271`` execute.  TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists next to the numpy FFT oracle (griffin_lim.py):
* it cross-checks that oracle with an independent formulation (tests/test_oracle_golden.py pins both against the
  reference's golden vectors);
* it is what ``bench.py --impl reference`` times: the FFT oracle is ~3.5x cheaper than what the reference actually
  runs (8.4 MFLOP per frame and transform as a dense convolution, plus a per-call Python loop for the
  window-sum-square, vocoder.py:71-82, kept here on purpose because it is ~20 % of the reference's time), so timing
  the FFT oracle would understate the reference arm's cost;
* ``tests/measure/conv_formulation_on_gpu.py`` runs the same code on the GPU (cuDNN convolutions): the "existing kernels" bar of
  SURVEY section 8d.
"""
import numpy as np
import torch
import torch.nn.functional as F

TINY = 1.1754944e-38  # vocoder.py:69


def padded_window(n_fft, win_length, window_fn=torch.hann_window):
    """audio_utils.py:218-223: the window centred in an n_fft frame."""
    pad = n_fft - win_length
    return F.pad(window_fn(win_length), (pad // 2, pad - pad // 2))


def dft_rows(n_fft):
    """audio_utils.py:226-231: [real rows 0..n_fft/2 ; imaginary rows 0..n_fft/2] of the DFT matrix, float32."""
    full = np.fft.fft(np.eye(n_fft))[: n_fft // 2 + 1]
    return torch.from_numpy(np.concatenate([full.real, full.imag], axis=0)).float()


class ConvGriffinLim:
    """Dense-basis Griffin-Lim (analysis = conv1d, synthesis = conv_transpose1d).  ``device`` may be a CUDA device."""

    def __init__(self, n_fft=2048, win_length=1200, hop_length=300, n_iter=64, device="cpu"):
        self.n_fft, self.win_length, self.hop, self.n_iter = n_fft, win_length, hop_length, n_iter
        self.device = torch.device(device)
        w = padded_window(n_fft, win_length)
        rows = dft_rows(n_fft)
        self.window = w
        self.analysis = (rows[:, None, :] * w).to(self.device)                       # [2F, 1, n_fft]  (audio_utils.py:254-256)
        self.synthesis = (torch.pinverse(n_fft / hop_length * rows).T[:, None, :] * w).to(self.device)  # vocoder.py:58-61

    def window_sum_square(self, n_frames):
        """vocoder.py:71-82, including its per-frame Python loop (part of the reference's cost per call)."""
        w2 = self.window ** 2
        n = self.n_fft + self.hop * (n_frames - 1)
        acc = torch.zeros(n, dtype=torch.float32)
        for t in range(n_frames):
            lo = t * self.hop
            acc[lo: min(n, lo + self.n_fft)] += w2[: max(0, min(self.n_fft, n - lo))]
        return acc

    def stft(self, wave):
        """audio_utils.py:259-271: wave [B, L] -> (magnitude, phase) [B, F, T]."""
        half = self.n_fft // 2
        x = F.conv1d(F.pad(wave[:, None, :], (half, half), mode="reflect"), self.analysis, stride=self.hop)
        re, im = x[:, : half + 1], x[:, half + 1:]
        return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)

    def istft(self, mag, phase):
        """vocoder.py:84-100: [B, F, T] x 2 -> [B, (T - 1) * hop]."""
        half = self.n_fft // 2
        x = F.conv_transpose1d(torch.cat([mag * torch.cos(phase), mag * torch.sin(phase)], dim=1), self.synthesis,
                               stride=self.hop)
        wss = self.window_sum_square(mag.shape[-1]).to(x.device)
        ok = wss > TINY
        x[:, :, ok] /= wss[ok]
        x *= self.n_fft / self.hop
        return x[:, 0, half: x.shape[-1] - half]

    def __call__(self, mag, phase0):
        """vocoder.py:102-110 with the initial phase supplied: mag, phase0 [B, F, T] -> [B, L]."""
        mag, phase0 = mag.to(self.device), phase0.to(self.device)
        y = self.istft(mag, phase0)
        for _ in range(self.n_iter):
            _, ph = self.stft(y)
            y = self.istft(mag, ph)
        return y


def vocoder_forward(logmel, phase0, n_iter, basis, device="cpu", gl=None):
    """GriffinLimVocoder.forward (vocoder.py:136-144) for one utterance: logmel [T, 80], phase0 [F, T] (numpy),
    basis = pinv mel [F, 80] -> waveform [(T - 1) * hop] (numpy float32)."""
    gl = gl or ConvGriffinLim(n_iter=n_iter, device=device)
    gl.n_iter = n_iter
    x = torch.from_numpy(np.ascontiguousarray(logmel, np.float32))
    mag = torch.from_numpy(np.asarray(basis, np.float32)).matmul(x.exp().T).clamp(min=0)[None]   # vocoder.py:42, 141
    y = gl(mag, torch.from_numpy(np.ascontiguousarray(phase0, np.float32))[None])
    return y[0].cpu().numpy()
